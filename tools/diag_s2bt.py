"""Where does the Step-2 BT/BGEN leg spend its time?  (host wall clock around the ABI calls)"""
import math
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from regenie_b200 import capi, hostprep, synth  # noqa: E402

N, nvar = 100_000, 400
rng = np.random.default_rng(1)
cov = rng.normal(size=(N, 2))
X, _, _, in_an, _ = hostprep.prepare_qt(rng.normal(size=(N, 1)), cov)
y = (rng.random(N) < 0.1).astype(np.float64)
mask = np.ones((N, 1), dtype=np.uint8)
p0 = float(y.mean()); eta = math.log(p0 / (1 - p0)); w = math.sqrt(p0 * (1 - p0))
st = capi.Step2(X, mask, in_an, N, nvar)
st.set_chr_bt(np.full((N, 1), w), np.full((N, 1), w), ((y - p0) / w)[:, None], [X], y[:, None], np.full((N, 1), eta))
g = torch.Generator().manual_seed(3)
maf = 0.01 + 0.49 * torch.rand((nvar, 1), generator=g)
u = torch.rand((nvar, N), generator=g)
hom = u < maf * maf
het = (u < 2 * maf - maf * maf) & ~hom
probs = torch.stack([hom.to(torch.uint8) * 255, het.to(torch.uint8) * 255], dim=2).contiguous().pin_memory().numpy()
miss = torch.full((nvar, N), 2, dtype=torch.uint8).pin_memory().numpy()
dprobs = torch.from_numpy(probs).cuda(); dmiss = torch.from_numpy(miss).cuda()
for name, a, b in (("host pinned", probs, miss), ("host, no ploidy bytes", probs, None)):
    st.block_bgen8_bt(a, b)
    t0 = time.perf_counter()
    for _ in range(5):
        o = st.block_bgen8_bt(a, b)
    print("%-24s score only: %.2f ms/block" % (name, (time.perf_counter() - t0) / 5 * 1e3))
sel = np.nonzero((np.abs(o["stat"][:, 0]) > 1.96) & ((o["flags"] & 17) == 0))[0]
t0 = time.perf_counter()
for _ in range(5):
    r = st.firth(sel, np.zeros(len(sel), dtype=np.int32))
print("firth on %d variants: %.2f ms/call" % (len(sel), (time.perf_counter() - t0) / 5 * 1e3))
st.set_timing(1) if hasattr(st, "set_timing") else None
