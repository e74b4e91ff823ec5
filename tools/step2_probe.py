"""Run a few Step-2 blocks with inputs resident in HBM (QT on .bed rows, BT on 8-bit dosages) - the command wrapped by
`ncu --metrics gpu__time_duration.sum` / `ncu --set full` to get the launch list and the roofline traffic of the s2_* kernels."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from regenie_b200 import capi, hostprep  # noqa: E402

N, P, C, bs, nb = 100_000, 10, 3, 1000, int(os.environ.get("RG_PROBE_BLOCKS", "3"))
REPS = int(os.environ.get("RG_PROBE_REPS", "2"))
dev = torch.device("cuda", 0)
Yr, cov, na = bench.gen_pheno(N, P, C, bench.SEED)
X, Y, mask, in_an, neff = hostprep.prepare_qt(Yr, cov, na)
panel = bench.gen_panel_gpu(torch, N, nb * bs, bs, bench.SEED, dev, 0.01)
stride = panel.shape[1]
rng = np.random.default_rng(1)
res = np.asfortranarray(rng.normal(size=(N, P)) * mask)
st = capi.Step2(X, mask, in_an, N, bs)
st.set_chr(res, np.ones(P))
out = st._out(bs)
for rep in range(REPS):
    for b in range(nb):
        st.block_bed_raw(panel.data_ptr() + b * bs * stride, bs, stride, out)
st.close()
# binary trait on dosages
nvar = 400
y = (rng.random(N) < 0.1).astype(np.float64)
p0 = float(y.mean()); w = (p0 * (1 - p0)) ** 0.5
gsm = np.full((N, 1), w); yres = ((y - p0) / w)[:, None]
m1 = np.ones((N, 1), dtype=np.uint8)
st = capi.Step2(X, m1, in_an, N, nvar)
st.set_chr_bt(gsm, gsm, yres, [X], y[:, None], np.full((N, 1), np.log(p0 / (1 - p0))))
probs = torch.randint(0, 120, (nvar, N, 2), dtype=torch.uint8, device=dev)
miss = torch.full((nvar, N), 2, dtype=torch.uint8, device=dev)
o = st._out(nvar, with_info=True)
for rep in range(REPS + 1):
    st.block_bgen8_bt_raw(probs.data_ptr(), miss.data_ptr(), N, nvar, o)
st.close()
print("step2 probe done")
