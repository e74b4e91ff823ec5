"""Time the rgb200 CLI end to end (files in, files out) on a synthetic fileset of BASELINE configs[1] shape
(N = 100k, bsize 1000, 10 traits; M = 10k variants by default).  Prints the driver's own phase timings."""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

N, M, P = 100_000, int(os.environ.get("M", "10000")), 10
d = tempfile.mkdtemp()
panel = bench.gen_panel_gpu(torch, N, M, 1000, 7, torch.device("cuda"), 0.01).cpu().numpy()
with open(d + "/syn.bed", "wb") as fh:
    fh.write(b"\x6c\x1b\x01")
    fh.write(panel.tobytes())
with open(d + "/syn.bim", "w") as fh:
    for i in range(M):
        fh.write("%d rs%d 0 %d A G\n" % (1 + i * 22 // M, i, 1000 + i))
with open(d + "/syn.fam", "w") as fh:
    for s in range(N):
        fh.write("F%d I%d 0 0 0 -9\n" % (s, s))
rng = np.random.default_rng(1)
Y = rng.standard_normal((N, P)); cov = rng.standard_normal((N, 2))
with open(d + "/pheno.txt", "w") as fh:
    fh.write("FID IID " + " ".join("Y%d" % (p + 1) for p in range(P)) + "\n")
    for s in range(N):
        fh.write("F%d I%d " % (s, s) + " ".join("%.6g" % v for v in Y[s]) + "\n")
with open(d + "/covar.txt", "w") as fh:
    fh.write("FID IID V1 V2\n")
    for s in range(N):
        fh.write("F%d I%d %.6g %.6g\n" % (s, s, cov[s, 0], cov[s, 1]))
rgb = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "regenie_b200", "rgb200")
for step, extra in ((1, []), (2, ["--pred", d + "/fit_pred.list"])):
    t0 = time.perf_counter()
    r = subprocess.run([rgb, "--step", str(step), "--bed", d + "/syn", "--phenoFile", d + "/pheno.txt", "--covarFile",
                        d + "/covar.txt", "--bsize", "1000", "--out", d + ("/fit" if step == 1 else "/s2")] + extra,
                       capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stdout[-2000:]
    keep = [l for l in r.stdout.splitlines() if "Level 0 done" in l or "Elapsed" in l]
    print("step %d: wall %.2f s (%d variants -> %.0f variants/s incl. file parsing and output) | %s" %
          (step, dt, M, M / dt, " | ".join(keep)))
