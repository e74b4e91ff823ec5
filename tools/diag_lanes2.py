import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from regenie_b200 import capi, hostprep
N, M, bs = 100000, int(sys.argv[1]), 1000
dev = torch.device("cuda", 0)
Yr, cov, na = bench.gen_pheno(N, 10, 3, bench.SEED)
X, Y, mask, in_an, neff = hostprep.prepare_qt(Yr, cov, na)
fsz = hostprep.fold_sizes(N, 5)
h = hostprep.ridge_grid(5); lam = 50000 * (1 - h) / h
panel = bench.gen_panel_gpu(torch, N, M, bs, bench.SEED, dev, 0.01)
stride = panel.shape[1]
host_panel = torch.empty(panel.shape, dtype=torch.uint8, pin_memory=True); host_panel.copy_(panel); torch.cuda.synchronize()
blocks = bench.blocks_of(M, bs)
os.environ["RG_B200_LANES"] = sys.argv[2]
for mode in sys.argv[3].split(","):
    st = capi.Step1(X, Y, mask, in_an, fsz, lam, neff, N, bs, len(blocks))
    if "timing" in mode: st.set_timing(True)
    base = host_panel.data_ptr() if mode == "host" else panel.data_ptr()
    for rep in range(6):
        for b, (s, n) in enumerate(blocks):
            st.l0_block_bed(base + s * stride, n, b, row_stride=stride)
        if "fence" in mode: st.fence()
        code = st.status()
        dbg = st.debug("dbg_counter", np.uint64, 1) if os.environ.get("RG_DBG_CHECK_DIAG") else None
        print(mode, "pass", rep, "status", code, "diag mismatches", dbg, flush=True)
        if code: break
    st.close()
