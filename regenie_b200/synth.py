"""Deterministic synthetic PLINK panels: per-SNP MAF ~ U(0.01, 0.5),
calls ~ Binomial(2, MAF), a fraction of missing calls, packed as real .bed rows."""
import numpy as np

SEED = 20260924
# PLINK 2-bit codes (ref-last): dosage 2 -> 00, missing -> 01, 1 -> 10, 0 -> 11
_CODE = np.array([3, 2, 0, 1], dtype=np.uint8)   # index: 0,1,2 dosage, 3 = missing


def pack_bed(g):
    """g: int array [M, N] with values 0/1/2 and 3 for missing -> uint8 [M, ceil(N/4)]."""
    M, N = g.shape
    pad = (-N) % 4
    c = _CODE[g]
    if pad:
        c = np.concatenate([c, np.zeros((M, pad), dtype=np.uint8)], axis=1)
    c = c.reshape(M, -1, 4)
    return (c[:, :, 0] | (c[:, :, 1] << 2) | (c[:, :, 2] << 4) | (c[:, :, 3] << 6)).astype(np.uint8)


def genotypes(N, M, seed=SEED, miss=0.01, maf_lo=0.01, maf_hi=0.5):
    rng = np.random.default_rng(seed)
    maf = rng.uniform(maf_lo, maf_hi, size=M)
    g = rng.binomial(2, maf[:, None], size=(M, N)).astype(np.uint8)
    # no monomorphic SNPs (reference throws on sd < 1e-6, src/Data.cpp:207)
    for i in np.where(g.min(axis=1) == g.max(axis=1))[0]:
        g[i, rng.integers(0, N, size=3)] = [0, 1, 2]
    if miss > 0:
        g[rng.random(size=(M, N)) < miss] = 3
    return g


def phenotypes(g, P, C, seed=SEED, h2=0.2, n_causal=100, na_frac=0.0):
    """Y = G beta + noise, covariates ~ N(0,1) (+ intercept added by the caller)."""
    M, N = g.shape
    rng = np.random.default_rng(seed + 1)
    gg = np.where(g == 3, 0, g).astype(np.float64)
    gg = (gg - gg.mean(axis=1, keepdims=True))
    sd = gg.std(axis=1, keepdims=True); sd[sd == 0] = 1
    gg /= sd
    Y = np.empty((N, P))
    for p in range(P):
        idx = rng.choice(M, size=min(n_causal, M), replace=False)
        b = rng.normal(size=len(idx)) * np.sqrt(h2 / len(idx))
        Y[:, p] = b @ gg[idx] + rng.normal(size=N) * np.sqrt(1 - h2)
    cov = rng.normal(size=(N, C - 1))
    na = rng.random(size=(N, P)) < na_frac
    return Y, cov, na


# ---------------------------------------------------------------------------------------- synthetic .pgen files
# A writer of PLINK 2 .pgen hard-call files for tests and bench.py (there is no plink2 in the image).  tests/test_pgen_cpu.py
# checks its output with the reference's own pgenlib (PgrValidate + ReadHardcalls).
def _vint_bytes(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _pack2(vals):
    vals = np.asarray(vals, dtype=np.uint8)
    pad = (-len(vals)) % 4
    v = np.concatenate([vals, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    return (v[:, 0] | (v[:, 1] << 2) | (v[:, 2] << 4) | (v[:, 3] << 6)).astype(np.uint8).tobytes()


def _difflist_bytes(ids, vals, n):
    ids = list(map(int, ids))
    out = bytearray(_vint_bytes(len(ids)))
    if not ids:
        return bytes(out)
    sb = 1 if n <= 0xFF else 2 if n <= 0xFFFF else 3 if n <= 0xFFFFFF else 4
    ng = (len(ids) + 63) // 64
    deltas = []
    for g in range(ng):
        grp = ids[g * 64:(g + 1) * 64]
        out += grp[0].to_bytes(sb, "little")
        deltas.append(b"".join(_vint_bytes(b - a) for a, b in zip(grp[:-1], grp[1:])))
    for g in range(ng - 1):
        out.append((len(deltas[g]) - 63) & 0xFF)          # group byte counts (only used for random access)
    out += _pack2(vals)
    for dl in deltas:
        out += dl
    return bytes(out)


def write_pgen(prefix, g, storage=1, records_out=None):
    """Write <prefix>.pgen (mode 0x10) for g [M, N] with values 0/1/2 (ALT counts) and 3 (missing), choosing for every
    variant the most compressed record type that applies, like plink2 does: all-zero (5), difflist over a constant
    (4/6/7), LD difflist against the previous non-LD record (2) or its inversion (3), 1-bit + difflist (1), plain (0).
    Returns the list of record types used."""
    M, N = g.shape
    recs, types = [], []
    base = None
    inv = np.array([2, 1, 0, 3], dtype=np.uint8)
    for v in range(M):
        x = g[v].astype(np.uint8)
        cands = []
        if not x.any():
            cands.append((5, b""))
        for const, t in ((0, 4), (2, 6), (3, 7)):
            ids = np.nonzero(x != const)[0]
            if len(ids) <= N // 8:
                cands.append((t, _difflist_bytes(ids, x[ids], N)))
        if base is not None:
            for t, tgt in ((2, x), (3, inv[x])):
                ids = np.nonzero(tgt != base)[0]
                if len(ids) <= N // 8:
                    cands.append((t, _difflist_bytes(ids, tgt[ids], N)))
        cnt = np.bincount(x, minlength=4)
        top = sorted(np.argsort(-cnt, kind="stable")[:2])
        lo, hi = int(top[0]), int(top[1])
        ids = np.nonzero((x != lo) & (x != hi))[0]
        if len(ids) < N // 16:                                # pgenlib's validator: 1-bit exceptions < n / 16
            bits = np.packbits((x == hi).astype(np.uint8), bitorder="little").tobytes()
            cands.append((1, bytes([(lo << 2) | (hi - lo)]) + bits + _difflist_bytes(ids, x[ids], N)))
        cands.append((0, _pack2(x)))
        t, rec = min(cands, key=lambda c: (len(c[1]), c[0]))
        if (t & 6) != 2:
            base = x.copy()
        recs.append(rec); types.append(t)
    if records_out is not None:
        records_out.extend(recs)
    lb = 1 + (storage & 3)
    hdr = bytearray(b"\x6c\x1b\x10") + M.to_bytes(4, "little") + N.to_bytes(4, "little") + bytes([0x80 | storage])
    body_hdr = bytearray()
    if storage < 4:
        tt = types + [0] * (len(types) % 2)
        body_hdr += bytes(tt[i] | (tt[i + 1] << 4) for i in range(0, len(tt), 2))
    else:
        body_hdr += bytes(types)
    for r in recs:
        body_hdr += len(r).to_bytes(lb, "little")
    first = len(hdr) + 8 + len(body_hdr)
    with open(prefix + ".pgen", "wb") as fh:
        fh.write(hdr + first.to_bytes(8, "little") + body_hdr + b"".join(recs))
    return types




def gather_pgen_records(record_of, type_of, variants):
    """The tables of include/rg_b200.h's rg_pgen_block for the given variants (what host/pgen.cpp PgenFile::gather builds):
    record_of(v) -> bytes, type_of(v) -> low 3 bits of the record type; bases of LD-compressed records are added once."""
    data = bytearray()
    rec_off, rec_len, rec_type, own, base = [], [], [], [], []
    last = (-1, -1)

    def add(v):
        while len(data) % 16:
            data.append(0)
        r = record_of(v)
        rec_off.append(len(data)); rec_len.append(len(r)); rec_type.append(type_of(v))
        data.extend(r)
        return len(rec_off) - 1
    for v in variants:
        t = type_of(v)
        own.append(add(v))
        if (t & 6) != 2:
            last = (v, own[-1]); base.append(-1)
            continue
        b = v - 1
        while (type_of(b) & 6) == 2:
            b -= 1
        if last[0] != b:
            last = (b, add(b))
        base.append(last[1])
    if not data:
        data.extend(b"\0" * 16)
    return dict(data=np.frombuffer(bytes(data), dtype=np.uint8), rec_off=rec_off, rec_len=rec_len, rec_type=rec_type,
                own=own, base=base)
