"""Throughput of the on-device BGEN inflate (rg_bgen_inflate) next to host zlib, on synthetic payloads of BGEN v1.2
layout-2 shape (10 + 3N bytes per variant: header, ploidy bytes, 8-bit probability pairs of imputed-looking data).

    python tools/inflate_bench.py [N] [bs] [reps]

Wall-clock around the C-ABI call (it returns after the device finished): includes the H2D copy of the compressed bytes.
RG_B200_INFLATE=window in the environment selects the ring-buffer kernel (inflate_zlib_window) instead of the direct one.
"""
import os
import sys
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regenie_b200 import capi  # noqa: E402


def payload(rng, n, maf):
    g = rng.binomial(2, maf, n)
    conf = rng.random(n) < 0.85                                    # 85 % of calls are certain (255 / 0)
    p = np.zeros((n, 2), dtype=np.uint8)
    p[g == 2, 0] = 255
    p[g == 1, 1] = 255
    u = ~conf
    a = rng.integers(0, 256, u.sum())
    b = (rng.random(u.sum()) * (255 - a)).astype(np.int64)
    p[u, 0], p[u, 1] = a, b
    hdr = np.zeros(8, dtype=np.uint8)
    hdr[:4] = np.frombuffer(np.uint32(n).tobytes(), dtype=np.uint8)
    hdr[4], hdr[6], hdr[7] = 2, 2, 2
    raw = np.concatenate([hdr, np.full(n, 2, dtype=np.uint8), np.array([0, 8], dtype=np.uint8), p.reshape(-1)])
    return raw.tobytes()


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    bs = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    rng = np.random.default_rng(1)
    t0 = time.time()
    raws = [payload(rng, N, m) for m in rng.uniform(0.01, 0.5, bs)]
    with ThreadPoolExecutor(16) as ex:
        comps = list(ex.map(lambda r: zlib.compress(r, 6), raws))
    offs = np.zeros(bs + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(c) for c in comps])
    comp = np.frombuffer(b"".join(comps), dtype=np.uint8)
    raw_bytes = bs * (10 + 3 * N)
    print("setup %.1fs: N=%d bs=%d raw %.1f MB compressed %.1f MB (ratio %.2f)" %
          (time.time() - t0, N, bs, raw_bytes / 1e6, comp.size / 1e6, raw_bytes / comp.size), flush=True)
    X = np.ones((N, 1)) / np.sqrt(N)
    s2 = capi.Step2(X, np.ones((N, 1), dtype=np.uint8), np.ones(N, dtype=np.uint8), N, bs)
    s2.bgen_inflate(comp, offs, N)                                  # warm-up (allocations)
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        s2.bgen_inflate(comp, offs, N)
        ts.append(time.perf_counter() - t)
    dev = min(ts)
    print("device inflate: %.2f ms per block -> %.1f GB/s of inflated bytes (%.0f variants/s)" %
          (dev * 1e3, raw_bytes / dev / 1e9, bs / dev), flush=True)
    for threads in (1, 32):
        with ThreadPoolExecutor(threads) as ex:
            t = time.perf_counter()
            out = list(ex.map(zlib.decompress, comps))
            host = time.perf_counter() - t
        assert out[0] == raws[0]
        print("host zlib, %2d threads: %.2f ms per block -> %.1f GB/s" % (threads, host * 1e3, raw_bytes / host / 1e9), flush=True)


if __name__ == "__main__":
    main()
