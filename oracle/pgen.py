"""PLINK 2 .pgen / .pvar / .psam reader for biallelic hard calls (oracle; test infrastructure).

The reference reads these files through the vendored pgenlib (`ReadHardcalls(..., allele_idx = 1)`,
src/Geno.cpp:1773-1821; .pvar / .psam parsing src/Geno.cpp:771-1011).  This module restates the on-disk layout of
the formats pgenlib documents (external_libs/pgenlib/include/pgenlib_read.cc: header PgfiInitPhase1/2 :684-1420,
difflists :2177-2268, 1-bit and difflist records :2597-2731, LD-compressed records) for storage modes 0x01 (.bed
inside a .pgen), 0x02 (fixed-width 2-bit) and 0x10 (variable-width records, 4/8-bit record types), hard calls only.
Value of a sample = ALT allele count 0/1/2, 3 = missing.
"""
import numpy as np

VBLOCK = 65536
GROUP = 64


def _vint(d, p):
    v, shift = 0, 0
    while True:
        b = d[p]; p += 1
        v |= (b & 0x7F) << shift
        if not (b & 0x80):
            return v, p
        shift += 7


def _unpack2(buf, n):
    a = np.frombuffer(buf, dtype=np.uint8)
    out = np.empty((len(a), 4), dtype=np.uint8)
    for k in range(4):
        out[:, k] = (a >> (2 * k)) & 3
    return out.reshape(-1)[:n]


class Pgen:
    def __init__(self, path):
        self.d = d = open(path, "rb").read()
        if d[:2] != b"\x6c\x1b":
            raise ValueError("not a .pgen file")
        self.mode = d[2]
        if self.mode == 0x01:
            raise ValueError("PLINK 1 .bed inside .pgen: sample count comes from the .psam; use the bed reader")
        self.m = int.from_bytes(d[3:7], "little")
        self.n = int.from_bytes(d[7:11], "little")
        ctrl = d[11]
        n4 = (self.n + 3) // 4
        if self.mode == 0x02:
            if ctrl & 63:
                raise ValueError("inconsistent fixed-width header")
            off = 12 + ((self.m + 7) // 8 if (ctrl >> 6) == 3 else 0)
            self.vrtype = np.zeros(self.m, dtype=np.uint8)
            self.fpos = off + n4 * np.arange(self.m + 1, dtype=np.int64)
            return
        if self.mode != 0x10:
            raise ValueError("unsupported .pgen storage mode 0x%02x (dosage / extensions are outside the oracle)" % self.mode)
        storage = ctrl & 15
        if storage >= 8:
            raise ValueError("single-sample fused record-type headers are not supported")
        if (ctrl >> 4) & 3:
            raise ValueError("multiallelic .pgen files are not supported")
        nonref_stored = (ctrl >> 6) == 3
        nblk = (self.m - 1) // VBLOCK + 1
        p = 12
        blk_fpos = [int.from_bytes(d[p + 8 * b:p + 8 * b + 8], "little") for b in range(nblk)]
        p += 8 * nblk
        lb = 1 + (storage & 3)
        vrt, fpos = [], []
        for b in range(nblk):
            cnt = min(VBLOCK, self.m - b * VBLOCK)
            if storage < 4:
                raw = np.frombuffer(d[p:p + (cnt + 1) // 2], dtype=np.uint8); p += (cnt + 1) // 2
                t = np.empty(2 * len(raw), dtype=np.uint8)
                t[0::2] = raw & 15; t[1::2] = raw >> 4
                vrt.append(t[:cnt])
            else:
                vrt.append(np.frombuffer(d[p:p + cnt], dtype=np.uint8)); p += cnt
            lens = np.frombuffer(d[p:p + cnt * lb], dtype=np.uint8).reshape(cnt, lb).astype(np.int64); p += cnt * lb
            ln = sum(lens[:, k] << (8 * k) for k in range(lb))
            fpos.append(blk_fpos[b] + np.concatenate([[0], np.cumsum(ln)[:-1]]))
            last_end = blk_fpos[b] + int(ln.sum())
            if nonref_stored:
                p += (cnt + 7) // 8
        self.vrtype = np.concatenate(vrt)
        self.fpos = np.concatenate(fpos + [[last_end]]).astype(np.int64)
        self._base = (-1, None)

    def _difflist(self, p):
        """-> (sample ids, 2-bit values, next position)."""
        d = self.d
        ln, p = _vint(d, p)
        if ln == 0:
            return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.uint8), p
        ng = (ln + GROUP - 1) // GROUP
        sb = 1 if self.n <= 0xFF else 2 if self.n <= 0xFFFF else 3 if self.n <= 0xFFFFFF else 4
        first = [int.from_bytes(d[p + g * sb:p + (g + 1) * sb], "little") for g in range(ng)]
        p += ng * (sb + 1) - 1
        vals = _unpack2(d[p:p + (ln + 3) // 4], ln); p += (ln + 3) // 4
        ids = np.empty(ln, dtype=np.int64)
        k = 0
        for g in range(ng):
            cur = first[g]
            ids[k] = cur; k += 1
            for _ in range(min(GROUP, ln - g * GROUP) - 1):
                dv, p = _vint(d, p)
                cur += dv
                ids[k] = cur; k += 1
        return ids, vals, p

    def _nonld(self, v):
        t = int(self.vrtype[v]) & 7
        p = int(self.fpos[v])
        n = self.n
        if t == 0:
            return _unpack2(self.d[p:p + (n + 3) // 4], n).copy()
        if t == 1:
            code = self.d[p]
            bits = np.unpackbits(np.frombuffer(self.d[p + 1:p + 1 + (n + 7) // 8], dtype=np.uint8), bitorder="little")[:n]
            g = ((code >> 2) + bits * (code & 3)).astype(np.uint8)
            ids, vals, _ = self._difflist(p + 1 + (n + 7) // 8)
            g[ids] = vals
            return g
        if t == 5:
            return np.zeros(n, dtype=np.uint8)
        g = np.full(n, t & 3, dtype=np.uint8)               # 4: all 0, 6: all 2, 7: all missing
        ids, vals, _ = self._difflist(p)
        g[ids] = vals
        return g

    def read(self, v):
        """ALT allele counts of variant v (0/1/2, 3 = missing) for all samples in the file."""
        vt = int(self.vrtype[v])
        if vt & 0xF8 & ~0x10:
            raise ValueError("multiallelic or dosage track in variant record %d" % v)
        t = vt & 7
        if (t & 6) != 2:
            g = self._nonld(v)
            self._base = (v, g)
            return g.copy()
        b = v - 1
        while (int(self.vrtype[b]) & 6) == 2:                # most recent record that is not LD-compressed
            b -= 1
        if self._base[0] != b:
            self._base = (b, self._nonld(b))
        g = self._base[1].copy()
        ids, vals, _ = self._difflist(int(self.fpos[v]))
        g[ids] = vals
        if t == 3:                                           # inverted: 0 <-> 2
            g = np.where(g == 0, 2, np.where(g == 2, 0, g)).astype(np.uint8)
        return g


def read_pvar(path):
    """src/Geno.cpp:771-800: skip '##' lines, header '#CHROM POS ID REF ALT'; ALLELE0 = REF, ALLELE1 = ALT."""
    rows = []
    cols = None
    for line in open(path):
        if line.startswith("##"):
            continue
        t = line.split()
        if cols is None:
            if not t or t[0] != "#CHROM":
                raise ValueError("header of pvar file does not have correct format.")
            cols = {k: t.index(k) for k in ("POS", "ID", "REF", "ALT")}
            continue
        if len(t) < 5:
            raise ValueError("incorrectly formatted pvar file")
        rows.append((t[0], int(t[cols["POS"]]), t[cols["ID"]], t[cols["REF"]], t[cols["ALT"]]))
    return rows


def read_psam(path):
    """src/Geno.cpp:941-1011: header must start '#FID IID'; optional SEX column. -> (keys, sex)."""
    keys, sex = [], []
    si = None
    for ln, line in enumerate(open(path)):
        t = line.split()
        if ln == 0:
            if len(t) < 2 or t[0] != "#FID" or t[1] != "IID":
                raise ValueError("header does not have the correct format (must start with #FID IID).")
            si = t.index("SEX") if "SEX" in t else None
            continue
        if not t:
            continue
        keys.append(t[0] + "_" + t[1])
        sex.append(int(t[si]) if si is not None and t[si] in ("0", "1", "2") else 0)
    return keys, np.array(sex)
