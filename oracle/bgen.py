"""BGEN v1.2 (layout 2, zlib, 8-bit, unphased, biallelic, diploid) reader (oracle; test infrastructure).

The reference parses the file header / variant identifying data through the un-vendored BGEN library
v1.1.7 (`genfile::bgen`, src/bgen_to_vcf.hpp:92-120, src/Geno.cpp:38-178) and the probability blocks by
hand (readChunkFromBGEN src/Geno.cpp:2122-2170, parseSnpfromBGEN :2186-2345).  The header layout below is
restated from the public BGEN v1.2 specification; the dosage / INFO arithmetic follows parseSnpfromBGEN.
"""
import struct
import zlib

import numpy as np


def _zstd_decompress(buf, dl):
    import ctypes
    lib = ctypes.CDLL("libzstd.so.1")
    lib.ZSTD_decompress.restype = ctypes.c_size_t
    lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    out = ctypes.create_string_buffer(dl)
    got = lib.ZSTD_decompress(out, dl, bytes(buf), len(buf))
    assert got == dl
    return out.raw


class Bgen:
    def __init__(self, path):
        self.data = open(path, "rb").read()
        d = self.data
        (offset,) = struct.unpack_from("<I", d, 0)
        lh, m, n = struct.unpack_from("<III", d, 4)
        if d[16:20] not in (b"bgen", b"\0\0\0\0"):
            raise ValueError("not a bgen file")
        (flags,) = struct.unpack_from("<I", d, 4 + lh - 4)
        self.compression = flags & 3
        self.layout = (flags >> 2) & 0xF
        self.has_ids = bool(flags >> 31)
        self.n_samples, self.n_variants = n, m
        if self.layout != 2 or self.compression not in (0, 1, 2):
            raise ValueError("oracle reader supports layout 2 only")
        self.sample_ids = []
        pos = 4 + lh
        if self.has_ids:
            lsi, ns = struct.unpack_from("<II", d, pos)
            p = pos + 8
            for _ in range(ns):
                (l,) = struct.unpack_from("<H", d, p)
                self.sample_ids.append(d[p + 2:p + 2 + l].decode())
                p += 2 + l
        self.start = offset + 4

    def variants(self):
        """Yield (chrom str, pos, rsid, alleles, p0 u8[N], p1 u8[N], missing bool[N]) per variant."""
        d, p = self.data, self.start
        for _ in range(self.n_variants):
            (l,) = struct.unpack_from("<H", d, p); p += 2 + l                     # SNPID (unused)
            (l,) = struct.unpack_from("<H", d, p); rsid = d[p + 2:p + 2 + l].decode(); p += 2 + l
            (l,) = struct.unpack_from("<H", d, p); chrom = d[p + 2:p + 2 + l].decode(); p += 2 + l
            (pos,) = struct.unpack_from("<I", d, p); p += 4
            (k,) = struct.unpack_from("<H", d, p); p += 2
            alleles = []
            for _a in range(k):
                (l,) = struct.unpack_from("<I", d, p); alleles.append(d[p + 4:p + 4 + l].decode()); p += 4 + l
            if self.compression == 0:
                (c,) = struct.unpack_from("<I", d, p); p += 4
                raw = d[p:p + c]; p += c
            else:
                c, dl = struct.unpack_from("<II", d, p); p += 8
                raw = zlib.decompress(d[p:p + c - 4]) if self.compression == 1 else _zstd_decompress(d[p:p + c - 4], dl)
                p += c - 4
                assert len(raw) == dl
            n, ka, pmin, pmax = struct.unpack_from("<IHBB", raw, 0)
            assert ka == 2 and pmin == 2 and pmax == 2
            ploidy = np.frombuffer(raw, dtype=np.uint8, count=n, offset=8)
            phased, bits = raw[8 + n], raw[9 + n]
            assert phased == 0 and bits == 8
            probs = np.frombuffer(raw, dtype=np.uint8, count=2 * n, offset=10 + n).reshape(n, 2)
            yield chrom, pos, rsid, alleles, probs[:, 0].copy(), probs[:, 1].copy(), (ploidy & 0x80) != 0


def dosage(p0, p1, missing, ref_first=False):
    """parseSnpfromBGEN (src/Geno.cpp:2273-2295): dosage and the per-sample INFO numerator term."""
    a = p0 / 255.0
    b = p1 / 255.0
    c = np.maximum(1 - a - b, 0.0)
    if ref_first:
        g = b + 2 * c
        ival = 4 * c + b - g * g
    else:
        g = b + 2 * a
        ival = 4 * a + b - g * g
    g = np.where(missing, -3.0, g)
    return g, np.where(missing, 0.0, ival)
