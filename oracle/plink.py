"""PLINK .bed/.bim/.fam readers (oracle; test infrastructure only).

Restates rgcgithub/regenie v4.1.2:
  read_bim                 src/Geno.cpp:518-611   (chrStrToInt src/Regenie.cpp:1583-1594)
  read_fam                 src/Geno.cpp:643-691
  prep_bed (magic bytes)   src/Geno.cpp:735-752
  buildLookupTable         src/Geno.cpp:2833-2857 (2-bit code -> {2, NA, 1, 0})
  readChunkFromBedFileToG  src/Geno.cpp:1702-1768
  mean_impute_g            src/Geno.cpp:3183-3193
"""
import re
from dataclasses import dataclass, field

import numpy as np

NCHROM = 23          # src/Regenie.hpp:217
MISSING_G = -3.0     # src/Geno.cpp:2843


def chr_str_to_int(s: str) -> int:
    """src/Regenie.cpp:1583-1594."""
    s = re.sub(r"^chr", "", s)
    if s and s[0].isdigit():
        m = re.match(r"\d+", s)
        c = int(m.group(0))
        if 1 <= c <= NCHROM:
            return c
    elif s in ("X", "XY", "Y", "PAR1", "PAR2"):
        return NCHROM
    return -1


@dataclass
class Bim:
    chrom: np.ndarray      # int
    ids: list
    pos: np.ndarray        # int
    allele0: list          # reference allele (ALLELE0)
    allele1: list          # effect allele (ALLELE1)
    offset: np.ndarray     # 0-based row in the .bed
    chr_read: list = field(default_factory=list)


def read_bim(path: str, ref_first: bool = False, exclude=None) -> Bim:
    """src/Geno.cpp:518-611.  Default is ref-last: ALLELE1 = col 5, ALLELE0 = col 6."""
    chrom, ids, pos, a0, a1, off = [], [], [], [], [], []
    chr_read = []
    exclude = set(exclude or ())
    with open(path) as fh:
        for lineno, line in enumerate(fh):
            t = line.rstrip("\r\n").split()
            if len(t) < 6:
                raise ValueError(f"incorrectly formatted bim file at line {lineno + 1}")
            c = chr_str_to_int(t[0])
            if c == -1:
                raise ValueError(f"unknown chromosome code in bim file at line {lineno + 1}")
            if not chr_read or c != chr_read[-1]:
                if chr_read and c <= max(chr_read):
                    raise ValueError("chromosomes in bim file are not in ascending order.")
                chr_read.append(c)
            if t[1] in exclude:
                continue
            chrom.append(c)
            ids.append(t[1])
            pos.append(int(t[3], 0))
            if ref_first:
                a0.append(t[4]); a1.append(t[5])   # allele1(ref)=col5, allele2=col6
            else:
                a0.append(t[5]); a1.append(t[4])
            off.append(lineno)
    return Bim(np.array(chrom), ids, np.array(pos), a0, a1, np.array(off, dtype=np.int64), chr_read)


def read_fam(path: str):
    """src/Geno.cpp:643-691.  Returns (list of 'FID_IID' keys in file order, sex array)."""
    keys, sex = [], []
    seen = set()
    with open(path) as fh:
        for line in fh:
            t = line.rstrip("\r\n").split()
            if len(t) < 6:
                raise ValueError("incorrectly formatted fam file")
            k = t[0] + "_" + t[1]
            if k in seen:
                raise ValueError("duplicate individual in fam file : " + k)
            seen.add(k)
            keys.append(k)
            sex.append(int(t[4]) if t[4] in ("0", "1", "2") else 0)
    return keys, np.array(sex)


# byte -> 4 genotype values, sample k of the byte in bits 2k..2k+1
_MAP = np.array([2.0, MISSING_G, 1.0, 0.0])                 # src/Geno.cpp:2843
_LUT = np.zeros((256, 4))
for _b in range(256):
    for _j in range(4):
        _LUT[_b, _j] = _MAP[(_b >> (2 * _j)) & 3]


def read_bed_rows(path: str, n_file: int, rows) -> np.ndarray:
    """Raw packed rows [len(rows)][ceil(n_file/4)] (src/Geno.cpp:1714-1719)."""
    stride = (n_file + 3) // 4
    rows = np.asarray(rows, dtype=np.int64)
    out = np.empty((len(rows), stride), dtype=np.uint8)
    with open(path, "rb") as fh:
        magic = fh.read(3)
        if magic != b"\x6c\x1b\x01":                         # src/Geno.cpp:744-746
            raise ValueError("invalid bed file (SNP-major magic bytes expected)")
        for i, r in enumerate(rows):
            fh.seek(3 + int(r) * stride)
            out[i] = np.frombuffer(fh.read(stride), dtype=np.uint8)
    return out


def decode_bed(packed: np.ndarray, n_file: int, keep=None, ref_first: bool = False) -> np.ndarray:
    """Packed rows -> hard calls in {0,1,2,-3}, shape [bs][n_kept] (src/Geno.cpp:1727-1747).

    `keep` is the boolean complement of `ind_ignore` over the n_file samples.
    """
    g = _LUT[packed].reshape(packed.shape[0], -1)[:, :n_file]
    if keep is not None:
        g = g[:, np.asarray(keep, dtype=bool)]
    g = g.copy()
    if ref_first:
        nz = g != MISSING_G
        g[nz] = 2.0 - g[nz]
    return g


def mean_impute_block(g: np.ndarray, in_analysis: np.ndarray):
    """Step-1 imputation (src/Geno.cpp:1749-1762, 3183-3188).

    mean over analysed, non-missing samples; masked samples -> 0, missing -> mean.
    Returns (imputed block, per-SNP mean).
    """
    a = np.asarray(in_analysis, dtype=bool)[None, :]
    ok = a & (g != MISSING_G)
    tot = np.where(ok, g, 0.0).sum(axis=1)
    ns = ok.sum(axis=1)
    mu = tot / ns
    out = np.where(g == MISSING_G, mu[:, None], g)
    out = np.where(a, out, 0.0)
    return out, mu
