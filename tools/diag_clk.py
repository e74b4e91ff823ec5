import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from regenie_b200 import capi, hostprep
os.environ["RG_DBG_CLK"] = "1"; os.environ["RG_B200_LANES"] = "1"
N, M, bs = 100000, 3000, 1000
dev = torch.device("cuda", 0)
Yr, cov, na = bench.gen_pheno(N, 10, 3, 1)
X, Y, mask, in_an, neff = hostprep.prepare_qt(Yr, cov, na)
fsz = hostprep.fold_sizes(N, 5); h = hostprep.ridge_grid(5); lam = 50000 * (1 - h) / h
panel = bench.gen_panel_gpu(torch, N, M, bs, 5, dev, 0.01); stride = panel.shape[1]
st = capi.Step1(X, Y, mask, in_an, fsz, lam, neff, N, bs, 3)
for b in range(3):
    st.l0_block_bed(panel.data_ptr() + b * bs * stride, bs, b, row_stride=stride)
st.sync()
d = st.debug("dbg_clk", np.int64, 790 * 4).reshape(-1, 4)
d = d[d[:, 0] > 0]
main = d[:, 1] - d[:, 0]; epi = d[:, 2] - d[:, 1]
print("ctas", len(d), "mainloop cycles median/min/max", np.median(main), main.min(), main.max())
print("epilogue cycles median/min/max", np.median(epi), epi.min(), epi.max())
print("span of all CTAs (cycles)", d[:, 2].max() - d[:, 0].min())
