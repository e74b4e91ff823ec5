"""GPU diagnostic: lanes=1 vs lanes=N determinism at bench size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from regenie_b200 import capi, hostprep

N, M, bs = int(sys.argv[1]), int(sys.argv[2]), 1000
dev = torch.device("cuda", 0)
Yr, cov, na = bench.gen_pheno(N, 10, 3, 1)
X, Y, mask, in_an, neff = hostprep.prepare_qt(Yr, cov, na)
fsz = hostprep.fold_sizes(N, 5)
h = hostprep.ridge_grid(5); lam = 50000 * (1 - h) / h
panel = bench.gen_panel_gpu(torch, N, M, bs, 5, dev, 0.01)
stride = panel.shape[1]
blocks = bench.blocks_of(M, bs)
res = {}
for lanes, sync_each in ((1, False), (4, True), (4, False)):
    os.environ["RG_B200_LANES"] = str(lanes)
    st = capi.Step1(X, Y, mask, in_an, fsz, lam, neff, N, bs, len(blocks))
    for rep in range(2):
        for b, (s, n) in enumerate(blocks):
            st.l0_block_bed(panel.data_ptr() + s * stride, n, b, row_stride=stride)
            if sync_each:
                code = st.status()
                if code: print("lanes", lanes, "sync_each block", b, "status", code)
    code = st.status()
    print("lanes", lanes, "sync_each", sync_each, "final status", code, capi.lib().rg_last_error().decode() if code else "")
    res[(lanes, sync_each)] = [st.fetch_W(b, 0) for b in range(len(blocks))]
    st.close()
ref = res[(1, False)]
for k, v in res.items():
    bad = [b for b in range(len(blocks)) if not np.array_equal(v[b], ref[b])]
    print(k, "blocks differing from lanes=1:", bad, [float(np.abs(v[b] - ref[b]).max()) for b in bad[:4]])
