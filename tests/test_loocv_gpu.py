"""GPU parity for the --loocv path (BASELINE.json configs[0]: example.bed, 2 QTs, --bsize 100 --loocv)."""
import os
import subprocess

import numpy as np
import pytest

import helpers
from oracle import step1

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


def run_all(pb):
    st = pb.gpu_step1()
    for b in range(len(pb.blocks)):
        pb.gpu_l0_block(st, b)
    assert st.status() == 0
    P = pb.prep.Y.shape[1]
    W = [[st.fetch_W(b, ph) for ph in range(P)] for b in range(len(pb.blocks))]
    h1 = np.array([0.01, 0.25, 0.5, 0.75, 0.99])
    B = len(pb.blocks) * 5
    cs, best = st.l1_fit(np.tile(B * (1 - h1) / h1, (P, 1)))
    loco = st.loco([c for c, _, _ in pb.blocks])
    assert st.status() == 0
    return W, cs, best, loco


def oracle(pb):
    def gen():
        for b in range(len(pb.blocks)):
            yield pb.oracle_block(b)[0]
    return step1.run_step1_qt(gen(), pb.blocks, pb.prep, pb.fold_sizes, pb.M, loocv=True)


def check(pb):
    W, cs, best, loco = run_all(pb)
    o = oracle(pb)
    P = pb.prep.Y.shape[1]
    R = 5
    for ph in range(P):
        for b in range(len(pb.blocks)):
            assert rel(W[b][ph], o["W"][ph][:, b * R:(b + 1) * R]) < 1e-8
        assert rel(cs[:, ph, :], o["cs"][ph]) < 1e-7
        assert best[ph] == o["best"][ph]
        assert rel(loco[ph], o["loco"][ph]) < 1e-6


def test_loocv_synthetic(tmp_path):
    check(helpers.synthetic_problem(tmp_path, N=700, M=260, bsize=100, miss=0.02, loocv=True))


def test_loocv_example_config0(golden_dir):
    pb = helpers.Problem(golden_dir + "/example", golden_dir + "/phenotype.txt", golden_dir + "/covariates.txt", 100,
                         loocv=True)
    check(pb)


def test_loocv_example_3chr_driver(tmp_path, golden_dir):
    """rgb200 --step 1 --loocv on example_3chr: .loco files vs the oracle."""
    prefix = golden_dir + "/example_3chr"
    out = str(tmp_path / "fit")
    r = subprocess.run([os.path.join(ROOT, "regenie_b200", "rgb200"), "--step", "1", "--bed", prefix, "--phenoFile",
                        golden_dir + "/phenotype.txt", "--covarFile", golden_dir + "/covariates.txt", "--bsize", "100",
                        "--loocv", "--out", out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    pb = helpers.Problem(prefix, golden_dir + "/phenotype.txt", golden_dir + "/covariates.txt", 100, loocv=True)
    o = oracle(pb)
    for ph in range(2):
        ref = str(tmp_path / ("oracle_%d.loco" % (ph + 1)))
        step1.write_loco(ref, pb.keys, pb.prep.in_analysis, pb.prep.mask[:, ph], o["loco"][ph])
        la, lb = open(out + "_%d.loco" % (ph + 1)).read().splitlines(), open(ref).read().splitlines()
        assert la[0] == lb[0] and len(la) == len(lb) == 24
        for x, y in zip(la[1:], lb[1:]):
            for a, b in zip(x.split(), y.split()):
                assert a == b or abs(float(a) - float(b)) <= 2e-5 * max(abs(float(a)), abs(float(b))) + 1e-9
