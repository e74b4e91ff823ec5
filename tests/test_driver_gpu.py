"""End to end through the C++ host driver `rgb200` (regenie's CLI surface) vs the oracle:
.loco / _pred.list / .regenie files produced from the reference's own example filesets."""
import os
import subprocess

import numpy as np
import pytest

import helpers
from oracle import step1

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RGB = os.path.join(ROOT, "regenie_b200", "rgb200")


def run(args):
    r = subprocess.run([RGB] + args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def close(a, b, rtol=1.5e-5):
    """Printed 6-significant-digit values: equal up to the reference's own rounding noise."""
    if a == b:
        return True
    if "NA" in (a, b):
        return False
    x, y = float(a), float(b)
    return abs(x - y) <= rtol * max(abs(x), abs(y)) + 1e-12


def compare_token_files(fa, fb, exact_cols=0):
    la, lb = open(fa).read().splitlines(), open(fb).read().splitlines()
    assert len(la) == len(lb)
    assert la[0] == lb[0]                      # header: byte-identical
    for x, y in zip(la[1:], lb[1:]):
        tx, ty = x.split(), y.split()
        assert len(tx) == len(ty)
        assert tx[:exact_cols] == ty[:exact_cols]
        for a, b in zip(tx[exact_cols:], ty[exact_cols:]):
            assert close(a, b), (a, b, x[:80])


@pytest.mark.parametrize("fileset,bsize,remove", [("example_3chr", 100, False), ("example", 100, True)])
def test_step1_then_step2_files(tmp_path, golden_dir, fileset, bsize, remove):
    prefix = os.path.join(golden_dir, fileset)
    pheno, covar = golden_dir + "/phenotype.txt", golden_dir + "/covariates.txt"
    out1 = str(tmp_path / "fit")
    args = ["--step", "1", "--bed", prefix, "--phenoFile", pheno, "--covarFile", covar, "--bsize", str(bsize),
            "--lowmem", "--out", out1]
    rm = None
    if remove:
        args += ["--remove", golden_dir + "/fid_iid_to_remove.txt"]
        rm = {"_".join(l.split()[:2]) for l in open(golden_dir + "/fid_iid_to_remove.txt")}
    log = run(args)
    assert "<- min value" in log
    # oracle Step 1
    pb = helpers.Problem(prefix, pheno, covar, bsize, remove=rm)

    def gen():
        for b in range(len(pb.blocks)):
            yield pb.oracle_block(b)[0]
    o = step1.run_step1_qt(gen(), pb.blocks, pb.prep, pb.fold_sizes, pb.M)
    for ph in range(2):
        ref = str(tmp_path / ("oracle_%d.loco" % (ph + 1)))
        step1.write_loco(ref, pb.keys, pb.prep.in_analysis, pb.prep.mask[:, ph], o["loco"][ph])
        compare_token_files(out1 + "_%d.loco" % (ph + 1), ref, exact_cols=1)
    pl = [l.split() for l in open(out1 + "_pred.list")]
    assert [p[0] for p in pl] == ["Y1", "Y2"] and all(os.path.isabs(p[1]) for p in pl)

    # Step 2 on the same fileset, reading the driver's own .loco files
    out2 = str(tmp_path / "test")
    args2 = ["--step", "2", "--bed", prefix, "--phenoFile", pheno, "--covarFile", covar, "--bsize", "200",
             "--pred", out1 + "_pred.list", "--out", out2]
    if remove:
        args2 += ["--remove", golden_dir + "/fid_iid_to_remove.txt"]
    run(args2)
    rows = helpers.oracle_step2_rows(prefix, pheno, covar, out1 + "_pred.list", 200, remove=rm)
    for nm in ("Y1", "Y2"):
        got = open(out2 + "_%s.regenie" % nm).read().splitlines()
        exp = ["CHROM GENPOS ID ALLELE0 ALLELE1 A1FREQ N TEST BETA SE CHISQ LOG10P EXTRA"] + [r.rstrip("\n") for r in rows[nm]]
        assert len(got) == len(exp)
        assert got[0] == exp[0]
        for x, y in zip(got[1:], exp[1:]):
            tx, ty = x.split(), y.split()
            assert tx[:8] == ty[:8], (x, y)          # CHROM GENPOS ID ALLELE0 ALLELE1 A1FREQ N TEST: byte-identical
            for a, b in zip(tx[8:12], ty[8:12]):
                assert close(a, b), (x, y)
            assert tx[12] == ty[12]


def test_step2_htp_rows_match_the_oracle(tmp_path, golden_dir):
    """--htp COHORT (print_sum_stats_htp, src/Step2_Models.cpp:2542-2646) for quantitative traits: Name .. Model, AAF and the
    genotype counts byte-identical to the oracle restatement; Effect / CI / Pval and the numbers of the Info column within
    1e-5 relative."""
    prefix = os.path.join(golden_dir, "example")
    pheno, covar = golden_dir + "/phenotype.txt", golden_dir + "/covariates.txt"
    out1, out2 = str(tmp_path / "fit"), str(tmp_path / "htp")
    run(["--step", "1", "--bed", prefix, "--phenoFile", pheno, "--covarFile", covar, "--bsize", "100", "--out", out1])
    run(["--step", "2", "--bed", prefix, "--phenoFile", pheno, "--covarFile", covar, "--bsize", "200", "--pred",
         out1 + "_pred.list", "--htp", "TESTCOHORT", "--out", out2])
    rows = helpers.oracle_step2_rows(prefix, pheno, covar, out1 + "_pred.list", 200, htp="TESTCOHORT")
    from oracle import step2
    for nm in ("Y1", "Y2"):
        got = open(out2 + "_%s.regenie" % nm).read().splitlines()
        exp = [step2.HTP_HEADER.rstrip("\n")] + [r.rstrip("\n") for r in rows[nm]]
        assert got[0] == exp[0] and len(got) == len(exp) and len(got) > 900
        for x, y in zip(got[1:], exp[1:]):
            tx, ty = x.split("\t"), y.split("\t")
            assert len(tx) == 22 and tx[:8] == ty[:8] and tx[12:21] == ty[12:21], (x, y)   # ids, model; AAF + counts: identical
            for a, b in zip(tx[8:12], ty[8:12]):
                assert close(a, b), (x, y)
            ia, ib = tx[21].split(";"), ty[21].split(";")
            assert [t.split("=")[0] for t in ia] == [t.split("=")[0] for t in ib], (x, y)
            for a, b in zip(ia, ib):
                assert close(a.split("=")[1], b.split("=")[1]), (x, y)


def test_driver_rejects_out_of_scope_options(tmp_path):
    r = subprocess.run([RGB, "--step", "2", "--bed", "x", "--phenoFile", "y", "--bsize", "10", "--out",
                        str(tmp_path / "o"), "--pred", "z", "--spa"], capture_output=True, text=True)
    assert r.returncode != 0 and "ERROR" in r.stdout


def _write_zero_loco(tmp_path, golden_dir, traits):
    """LOCO files as the reference's BT Step 1 writes them for example/ (one chromosome): the row of chr 1
    is identically 0 (src/Data.cpp:1847-1858); the other rows are never read for this fileset."""
    from oracle import bgen
    b = bgen.Bgen(golden_dir + "/example.bgen")
    lst = tmp_path / "fit_bin_out_pred.list"
    with open(lst, "w") as fl:
        for j, t in enumerate(traits):
            f = tmp_path / ("fit_bin_out_%d.loco" % (j + 1))
            with open(f, "w") as fh:
                fh.write("FID_IID " + " ".join(b.sample_ids) + "\n")
                for c in range(1, 24):
                    fh.write(str(c) + " " + " ".join(["0"] * len(b.sample_ids)) + "\n")
            fl.write("%s %s\n" % (t, f))
    return str(lst)


def test_step2_bt_firth_bgen_reproduces_reference_golden_file(tmp_path, golden_dir):
    """The reference's documented Step-2 command (docs/docs/options.md:20-51, test/test_bash.sh) through rgb200:
    --step 2 --bgen example.bgen --bt --firth --approx --pThresh 0.01 --remove ... ; its output for Y1 must
    reproduce the golden file example/test_bin_out_firth_Y1.regenie that the reference ships: text columns
    (CHROM GENPOS ID ALLELE0 ALLELE1 A1FREQ INFO N TEST ... EXTRA) exactly, BETA/SE/CHISQ/LOG10P to the
    printed 6 significant digits (1e-4: the golden file was produced by an -ffast-math build)."""
    d = golden_dir
    pred = _write_zero_loco(tmp_path, d, ["Y1", "Y2"])
    out = str(tmp_path / "test_bin_out_firth")
    run(["--step", "2", "--bgen", d + "/example.bgen", "--covarFile", d + "/covariates.txt", "--phenoFile",
         d + "/phenotype_bin.txt", "--remove", d + "/fid_iid_to_remove.txt", "--bsize", "200", "--bt", "--firth",
         "--approx", "--pThresh", "0.01", "--pred", pred, "--out", out])
    got = open(out + "_Y1.regenie").read().splitlines()
    ref = open(d + "/test_bin_out_firth_Y1.regenie").read().splitlines()
    assert got[0] == ref[0]
    assert len(got) == len(ref) == 1001
    for x, y in zip(got[1:], ref[1:]):
        tx, ty = x.split(), y.split()
        assert tx[:9] == ty[:9], (x, y)
        assert tx[13] == ty[13]
        for a, c in zip(tx[9:13], ty[9:13]):
            assert close(a, c, rtol=1e-4), (x, y)
    assert os.path.getsize(out + "_Y2.regenie") > 0


def test_step2_bt_on_bed_hard_calls(tmp_path, golden_dir):
    """--bt --firth --approx on a .bed fileset (hard calls go through the same dosage kernels) vs the oracle."""
    import math

    from oracle import plink, prep, step2, step2_bt
    d = golden_dir
    pred = _write_zero_loco(tmp_path, d, ["Y1", "Y2"])
    out = str(tmp_path / "bt_bed")
    run(["--step", "2", "--bed", d + "/example", "--covarFile", d + "/covariates.txt", "--phenoFile",
         d + "/phenotype_bin.txt", "--bsize", "300", "--bt", "--firth", "--approx", "--pThresh", "0.05",
         "--pred", pred, "--out", out])
    bim = plink.read_bim(d + "/example.bim")
    keys, _ = plink.read_fam(d + "/example.fam")
    G = plink.decode_bed(plink.read_bed_rows(d + "/example.bed", len(keys), bim.offset), len(keys))
    pr = prep.prepare(keys, d + "/phenotype_bin.txt", d + "/covariates.txt", bt=True, step=2)
    z_thr = math.sqrt(3.841458820694124)
    for j, name in enumerate(["Y1", "Y2"]):
        y, mask = pr.Y_raw[:, j], pr.mask[:, j]
        st = step2_bt.BtChrom(y, pr.X, np.zeros(len(keys)), mask)
        got = open(out + "_%s.regenie" % name).read().splitlines()
        assert got[0] == step2.HEADER.strip()
        k = 1
        for v in range(len(bim.ids)):
            g = G[v]                                        # counts of ALLELE1, -3 = missing
            r = step2_bt.score_bt(g, np.zeros(len(g)), pr.in_analysis, mask, y, st, z_thr, len(keys))
            if r is None:
                continue
            row = step2.sumstats_row(int(bim.chrom[v]), int(bim.pos[v]), bim.ids[v], bim.allele0[v], bim.allele1[v], r["af"], r["n"],
                                     r["beta"], r["se"], r["chisq"], r["logp"], test_pass=not r["test_fail"]).split()
            tx = got[k].split()
            k += 1
            assert tx[:8] == row[:8], (tx, row)
            for a, c in zip(tx[8:12], row[8:12]):
                assert close(a, c), (tx, row)
            assert tx[12] == row[12]
        assert k == len(got)


def test_step2_qt_on_bgen_dosages(tmp_path, golden_dir):
    """Quantitative traits on BGEN dosages (BASELINE configs[2] shape): rgb200 --step 2 --bgen vs the oracle
    (compute_score_qt on the parseSnpfromBGEN dosages, INFO column included)."""
    from oracle import bgen, prep, step2
    d = golden_dir
    pred = _write_zero_loco(tmp_path, d, ["Y1", "Y2"])
    out = str(tmp_path / "qt_bgen")
    run(["--step", "2", "--bgen", d + "/example.bgen", "--covarFile", d + "/covariates.txt", "--phenoFile",
         d + "/phenotype.txt", "--remove", d + "/fid_iid_to_remove.txt", "--bsize", "400", "--pred", pred, "--out", out])
    rm = {"_".join(l.split()[:2]) for l in open(d + "/fid_iid_to_remove.txt") if l.strip()}
    b = bgen.Bgen(d + "/example.bgen")
    keep = np.array([k not in rm for k in b.sample_ids])
    keys = [k for k in b.sample_ids if k not in rm]
    pr = helpers.prepare_step2_with_mask(keys, d + "/phenotype.txt", d + "/covariates.txt",
                                         np.ones((len(keys), 2), dtype=bool))
    res, p_sd, scf = step2.compute_res(pr.Y, np.zeros_like(pr.Y), pr.mask, pr.neff, pr.ncov, pr.scale_Y)
    YtX = res.T @ pr.X
    rows = {nm: [] for nm in pr.pheno_names}
    for chrom, pos, rsid, alleles, p0, p1, miss in b.variants():
        g, iv = bgen.dosage(p0[keep], p1[keep], miss[keep])
        vs = step2.variant_stats(g, pr.in_analysis, pr.mask)
        if vs["ignored"]:
            continue
        sc = step2.score_qt(vs["g"], pr.X, res, pr.mask, pr.in_analysis, pr.n_analyzed, pr.ncov, scf, YtX, False)
        if sc is None:
            continue
        for ph, nm in enumerate(pr.pheno_names):
            if vs["ignored_trait"][ph]:
                continue
            okp = pr.in_analysis & (g != -3.0) & pr.mask[:, ph]
            af = vs["af"][ph]
            info = 1.0 if af in (0.0, 1.0) else 1 - iv[okp].sum() / (2 * vs["ns"][ph] * af * (1 - af))
            rows[nm].append(step2.sumstats_row(int(chrom), pos, rsid, alleles[1], alleles[0], af, vs["ns"][ph],
                                               sc["beta"][ph], sc["se"][ph], sc["chisq"][ph], sc["logp"][ph], info=info))
    for nm in pr.pheno_names:
        got = open(out + "_%s.regenie" % nm).read().splitlines()
        assert got[0] == step2.HEADER_INFO.strip()
        assert len(got) - 1 == len(rows[nm]) > 900
        for x, y in zip(got[1:], rows[nm]):
            tx, ty = x.split(), y.split()
            assert tx[:6] == ty[:6] and tx[7:9] == ty[7:9], (x, y)
            assert close(tx[6], ty[6]), (x, y)             # INFO: formed from sums in a different order
            for a, c in zip(tx[9:13], ty[9:13]):
                assert close(a, c), (x, y)


def test_step1_lowmem_keep_l0_files(tmp_path, golden_dir):
    """--lowmem --keep-l0: <prefix>_l0_Y<k> holds, per block, the N x R column-major f64 slab the reference's
    write_l0_file appends (src/Step1_Models.cpp:728-733; size check of read_l0_chunk :1962)."""
    prefix = os.path.join(golden_dir, "example_3chr")
    pheno, covar = golden_dir + "/phenotype.txt", golden_dir + "/covariates.txt"
    out = str(tmp_path / "fit")
    run(["--step", "1", "--bed", prefix, "--phenoFile", pheno, "--covarFile", covar, "--bsize", "100", "--lowmem",
         "--lowmem-prefix", str(tmp_path / "tmp_rg"), "--keep-l0", "--out", out])
    pb = helpers.Problem(prefix, pheno, covar, 100)
    n, R, nb = len(pb.keys), 5, len(pb.blocks)
    for ph in range(pb.prep.Y.shape[1]):
        f = str(tmp_path / ("tmp_rg_l0_Y%d" % (ph + 1)))
        assert os.path.getsize(f) == 8 * n * R * nb
        data = np.fromfile(f, dtype=np.float64).reshape(nb, R, n)
        for b in (0, nb - 1):
            W = pb.oracle_l0(b)[0]                      # [P][N x R]
            np.testing.assert_allclose(data[b].T, W[ph], rtol=1e-7, atol=1e-9)


def test_bt_step1_known_answer_then_step2_golden_file_end_to_end(tmp_path, golden_dir):
    """The reference's two documented commands (docs/docs/options.md:20-51, test/test_bash.sh:58-89) end to end
    through rgb200:  Step 1 `--bt --lowmem` (LOOCV is forced, N_analyzed = 494) must print the known answer
    `0.4504 ... <- min value`, its table / .loco files must match the oracle, and Step 2 on its _pred.list must
    reproduce the golden file the reference ships."""
    import test_oracle_golden as tog
    d = golden_dir
    out1 = str(tmp_path / "fit_bin_out")
    log = run(["--step", "1", "--bed", d + "/example", "--exclude", d + "/snplist_rm.txt", "--covarFile",
               d + "/covariates.txt", "--phenoFile", d + "/phenotype_bin.txt", "--remove", d + "/fid_iid_to_remove.txt",
               "--bsize", "100", "--bt", "--lowmem", "--lowmem-prefix", str(tmp_path / "tmp_rg"), "--out", out1])
    mins = [l for l in log.splitlines() if "min value" in l]
    assert len(mins) == 2 and any("0.4504" in l for l in mins), mins           # test/test_bash.sh:87
    tabs = tog.bt_step1_tables(d, with_loco=True)
    rows_got = [l for l in log.splitlines() if "Rsq =" in l]
    rows_ref = [r for t in tabs for r in t[1]]
    assert len(rows_got) == len(rows_ref) == 10
    for a, b in zip(rows_got, rows_ref):
        ta, tb = a.replace("<-", " <-").split(), b.replace("<-", " <-").split()
        assert len(ta) == len(tb), (a, b)
        for x, y in zip(ta, tb):
            x, y = x.rstrip(","), y.rstrip(",")
            try:
                float(y)
            except ValueError:
                assert x == y, (a, b)
                continue
            assert close(x, y), (a, b)
    for ph, (best, rows, loco, keys, pr) in enumerate(tabs):
        ref_file = str(tmp_path / ("ref_%d.loco" % (ph + 1)))
        step1.write_loco(ref_file, keys, pr.in_analysis, pr.mask[:, ph], loco)
        compare_token_files(out1 + "_%d.loco" % (ph + 1), ref_file, exact_cols=1)
    # Step 2 with the LOCO predictions rgb200 just wrote
    out2 = str(tmp_path / "test_bin_out_firth")
    run(["--step", "2", "--bgen", d + "/example.bgen", "--covarFile", d + "/covariates.txt", "--phenoFile",
         d + "/phenotype_bin.txt", "--remove", d + "/fid_iid_to_remove.txt", "--bsize", "200", "--bt", "--firth",
         "--approx", "--pThresh", "0.01", "--pred", out1 + "_pred.list", "--out", out2])
    got = open(out2 + "_Y1.regenie").read().splitlines()
    ref = open(d + "/test_bin_out_firth_Y1.regenie").read().splitlines()
    assert got[0] == ref[0] and len(got) == len(ref) == 1001
    for x, y in zip(got[1:], ref[1:]):
        tx, ty = x.split(), y.split()
        assert tx[:9] == ty[:9] and tx[13] == ty[13], (x, y)
        for a, c in zip(tx[9:13], ty[9:13]):
            assert close(a, c, rtol=1e-4), (x, y)


def test_step2_chr_jobs_concatenate_to_the_full_run(tmp_path, golden_dir):
    """Step-2 jobs split by chromosome (--chr / --chrList, one process per GPU) give, concatenated, exactly the
    rows of the single run - the way regenie users shard Step 2."""
    prefix = os.path.join(golden_dir, "example_3chr")
    pheno, covar = golden_dir + "/phenotype.txt", golden_dir + "/covariates.txt"
    out1 = str(tmp_path / "fit")
    run(["--step", "1", "--bed", prefix, "--phenoFile", pheno, "--covarFile", covar, "--bsize", "100", "--out", out1])
    base = ["--step", "2", "--bed", prefix, "--phenoFile", pheno, "--covarFile", covar, "--bsize", "200",
            "--pred", out1 + "_pred.list"]
    run(base + ["--out", str(tmp_path / "all")])
    run(base + ["--chr", "1", "--out", str(tmp_path / "c1")])
    run(base + ["--chrList", "2,3", "--out", str(tmp_path / "c23")])
    for nm in ("Y1", "Y2"):
        full = open(str(tmp_path / ("all_%s.regenie" % nm))).read().splitlines()
        a = open(str(tmp_path / ("c1_%s.regenie" % nm))).read().splitlines()
        b = open(str(tmp_path / ("c23_%s.regenie" % nm))).read().splitlines()
        assert a[0] == b[0] == full[0]
        assert a[1:] + b[1:] == full[1:]
        assert len(a) > 1 and len(b) > 1


def test_pgen_input_equals_bed_input(tmp_path, golden_dir):
    """--pgen vs --bed on the reference's own fixture pair (example.pgen / example.bed hold the same calls): Step 1 and
    Step 2 outputs must be byte-identical; then a synthetic .pgen that uses every hard-call record type (plain, 1-bit,
    difflists over 0 / 2 / missing, all-zero, LD and inverted LD) against the .bed written from the same calls."""
    import test_pgen_cpu as tp
    from regenie_b200 import synth
    d = golden_dir
    pheno, covar = d + "/phenotype.txt", d + "/covariates.txt"
    outs = {}
    for kind, arg in (("bed", ["--bed", d + "/example"]), ("pgen", ["--pgen", d + "/example"])):
        o1 = str(tmp_path / ("fit_" + kind))
        run(["--step", "1"] + arg + ["--phenoFile", pheno, "--covarFile", covar, "--bsize", "100", "--out", o1])
        o2 = str(tmp_path / ("s2_" + kind))
        run(["--step", "2"] + arg + ["--phenoFile", pheno, "--covarFile", covar, "--bsize", "300", "--pred",
                                     o1 + "_pred.list", "--out", o2])
        outs[kind] = (o1, o2)
    for k in (1, 2):
        assert open(outs["bed"][0] + "_%d.loco" % k).read() == open(outs["pgen"][0] + "_%d.loco" % k).read()
    for nm in ("Y1", "Y2"):
        assert open(outs["bed"][1] + "_%s.regenie" % nm).read() == open(outs["pgen"][1] + "_%s.regenie" % nm).read()
    # synthetic: all record types
    g = tp.synthetic_calls(N=700, M=160, seed=3)
    Y, cov, na = synth.phenotypes(np.where(g == 3, 0, g).astype(np.uint8), 2, 3, seed=2)
    prefix = helpers.write_fileset(str(tmp_path / "syn"), g, Y, cov, na)          # .bed/.bim/.fam + pheno/covar
    bim = [l.split() for l in open(prefix + ".bim")]
    keys = ["_".join(l.split()[:2]) for l in open(prefix + ".fam")]
    types = helpers.write_pgen(prefix, g, storage=5)
    assert set(types) == set(range(8))
    helpers.write_pvar_psam(prefix, [b[0] for b in bim], [b[1] for b in bim], [int(b[3]) for b in bim],
                            [b[5] for b in bim], [b[4] for b in bim], keys)           # REF = .bim col 6, ALT = col 5
    sd = str(tmp_path / "syn")
    res = {}
    for kind, arg in (("bed", ["--bed", prefix]), ("pgen", ["--pgen", prefix])):
        o2 = str(tmp_path / ("syn_s2_" + kind))
        lst = _write_zero_loco_keys(tmp_path, keys, ["Y1", "Y2"], kind)
        run(["--step", "2"] + arg + ["--phenoFile", sd + "/pheno.txt", "--covarFile", sd + "/covar.txt", "--bsize", "64",
                                     "--pred", lst, "--out", o2])
        res[kind] = o2
    for nm in ("Y1", "Y2"):
        a, b = open(res["bed"] + "_%s.regenie" % nm).read(), open(res["pgen"] + "_%s.regenie" % nm).read()
        assert a == b and a.count("\n") > 60


def _write_zero_loco_keys(tmp_path, keys, traits, tag):
    lst = tmp_path / ("zero_%s_pred.list" % tag)
    with open(lst, "w") as fl:
        for j, t in enumerate(traits):
            f = tmp_path / ("zero_%s_%d.loco" % (tag, j + 1))
            with open(f, "w") as fh:
                fh.write("FID_IID " + " ".join(keys) + "\n")
                for c in range(1, 24):
                    fh.write(str(c) + " " + " ".join(["0"] * len(keys)) + "\n")
            fl.write("%s %s\n" % (t, f))
    return str(lst)


def test_bgen_zstd_and_uncompressed_payloads(tmp_path, golden_dir):
    """BGEN compression flags 0 (none) and 2 (zstd) decode to the same probabilities as the zlib file the reference
    ships: identical Step-2 output (the reference reads both, src/Geno.cpp:1608-1610, :2207-2209)."""
    d = golden_dir
    pred = _write_zero_loco(tmp_path, d, ["Y1", "Y2"])
    outs = []
    for tag, mode in (("zlib", None), ("none", 0), ("zstd", 2)):
        f = d + "/example.bgen"
        if mode is not None:
            f = str(tmp_path / ("example_%s.bgen" % tag))
            helpers.recompress_bgen(d + "/example.bgen", f, mode)
        out = str(tmp_path / ("o_" + tag))
        run(["--step", "2", "--bgen", f, "--covarFile", d + "/covariates.txt", "--phenoFile", d + "/phenotype_bin.txt",
             "--remove", d + "/fid_iid_to_remove.txt", "--bsize", "200", "--bt", "--firth", "--approx", "--pThresh", "0.01",
             "--pred", pred, "--out", out])
        outs.append(open(out + "_Y1.regenie").read())
    assert outs[0] == outs[1] == outs[2] and outs[0].count("\n") == 1001


def test_split_l0_run_l0_run_l1_equals_single_run(tmp_path, golden_dir):
    """The reference's multi-job protocol (test/test_bash.sh:127-137): --split-l0 into 3 jobs, --run-l0 per job
    (one rgb200 process per job / GPU), --run-l1 -> .loco files byte-identical to the single run.  The job files have the
    layout of write_l0_file, so the jobs could equally be reference CPU jobs."""
    prefix = os.path.join(golden_dir, "example_3chr")
    pheno, covar = golden_dir + "/phenotype.txt", golden_dir + "/covariates.txt"
    base = ["--step", "1", "--bed", prefix, "--phenoFile", pheno, "--covarFile", covar, "--bsize", "100"]
    run(base + ["--out", str(tmp_path / "single")])
    mp = str(tmp_path / "par")
    run(base + ["--split-l0", mp + ",3", "--out", str(tmp_path / "split")])
    master = open(mp + ".master").read().split("\n")
    assert master[0].split()[1] == "100" and len([l for l in master[1:] if l.strip()]) == 3
    for j in (1, 2, 3):
        log = run(base + ["--run-l0", mp + ".master,%d" % j, "--out", str(tmp_path / ("job%d" % j))])
        assert "Done writing level 0 predictions to file." in log
        assert os.path.exists(mp + "_job%d_l0_Y1" % j)
    run(base + ["--run-l1", mp + ".master", "--out", str(tmp_path / "par_l1")])
    for k in (1, 2):
        assert open(str(tmp_path / ("single_%d.loco" % k))).read() == open(str(tmp_path / ("par_l1_%d.loco" % k))).read()
    assert not os.path.exists(mp + "_job1_l0_Y1")            # removed after level 1 like the reference (no --keep-l0)


def test_step2_bt_spa_driver_vs_oracle(tmp_path, golden_dir):
    """rgb200 --step 2 --bgen --bt --spa --pThresh 0.05 vs the oracle rows (saddlepoint correction)."""
    import math

    from oracle import bgen, prep, step2, step2_bt
    d = golden_dir
    pred = _write_zero_loco(tmp_path, d, ["Y1", "Y2"])
    out = str(tmp_path / "spa")
    run(["--step", "2", "--bgen", d + "/example.bgen", "--covarFile", d + "/covariates.txt", "--phenoFile",
         d + "/phenotype_bin.txt", "--remove", d + "/fid_iid_to_remove.txt", "--bsize", "400", "--bt", "--spa",
         "--pThresh", "0.05", "--pred", pred, "--out", out])
    rm = {"_".join(l.split()[:2]) for l in open(d + "/fid_iid_to_remove.txt") if l.strip()}
    b = bgen.Bgen(d + "/example.bgen")
    keep = np.array([k not in rm for k in b.sample_ids])
    keys = [k for k in b.sample_ids if k not in rm]
    pr = prep.prepare(keys, d + "/phenotype_bin.txt", d + "/covariates.txt", bt=True, step=2)
    y, mask = pr.Y_raw[:, 0], pr.mask[:, 0]
    st = step2_bt.BtChrom(y, pr.X, np.zeros(len(keys)), mask)
    z_thr = math.sqrt(3.841458820694124)
    got = open(out + "_Y1.regenie").read().splitlines()
    k = 1
    n_spa = 0
    for chrom, pos, rsid, alleles, p0, p1, miss in b.variants():
        g, iv = bgen.dosage(p0[keep], p1[keep], miss[keep])
        r = step2_bt.score_bt(g, iv, pr.in_analysis, mask, y, st, z_thr, len(keys), correction="spa")
        if r is None:
            continue
        n_spa += abs(r["stat"]) > z_thr
        row = step2.sumstats_row(int(chrom), pos, rsid, alleles[1], alleles[0], r["af"], r["n"], r["beta"], r["se"],
                                 r["chisq"], r["logp"], info=r["info"], test_pass=not r["test_fail"]).split()
        tx = got[k].split()
        k += 1
        assert tx[:6] == row[:6] and tx[7:9] == row[7:9] and tx[13] == row[13], (tx, row)
        assert close(tx[6], row[6])
        for a, c in zip(tx[9:13], row[9:13]):
            assert close(a, c), (tx, row)
    assert k == len(got) and n_spa > 20


def test_column_selection_ignore_pred_and_min_info(tmp_path, golden_dir):
    """--phenoColList / --covarColList pick columns like the reference (a run restricted to Y2 / V1 equals the Y2 file of a
    run on tables holding only those columns); --ignore-pred runs Step 2 with zero LOCO offsets (== the all-zero .loco
    files); --minINFO drops rows below the threshold."""
    d = golden_dir
    # tables with only Y2 and V1
    def cut(src, dst, cols):
        rows = [l.split() for l in open(src)]
        idx = [0, 1] + [rows[0].index(c) for c in cols]
        with open(dst, "w") as fh:
            for r in rows:
                fh.write(" ".join(r[i] for i in idx) + "\n")
    cut(d + "/phenotype.txt", str(tmp_path / "ph_y2.txt"), ["Y2"])
    cut(d + "/covariates.txt", str(tmp_path / "cv_v1.txt"), ["V1"])
    base = ["--step", "2", "--bgen", d + "/example.bgen", "--bsize", "500", "--ignore-pred"]
    run(base + ["--phenoFile", d + "/phenotype.txt", "--covarFile", d + "/covariates.txt", "--phenoColList", "Y2",
                "--covarColList", "V1", "--out", str(tmp_path / "sel")])
    run(base + ["--phenoFile", str(tmp_path / "ph_y2.txt"), "--covarFile", str(tmp_path / "cv_v1.txt"), "--out",
                str(tmp_path / "cutf")])
    a = open(str(tmp_path / "sel_Y2.regenie")).read()
    assert a == open(str(tmp_path / "cutf_Y2.regenie")).read() and a.count("\n") > 900
    assert not os.path.exists(str(tmp_path / "sel_Y1.regenie"))
    # --ignore-pred == zero LOCO predictions
    pred = _write_zero_loco(tmp_path, d, ["Y1", "Y2"])
    run(["--step", "2", "--bgen", d + "/example.bgen", "--bsize", "500", "--phenoFile", d + "/phenotype.txt", "--covarFile",
         d + "/covariates.txt", "--pred", pred, "--out", str(tmp_path / "zero")])
    run(base + ["--phenoFile", d + "/phenotype.txt", "--covarFile", d + "/covariates.txt", "--out", str(tmp_path / "ign")])
    for nm in ("Y1", "Y2"):
        assert open(str(tmp_path / ("zero_%s.regenie" % nm))).read() == open(str(tmp_path / ("ign_%s.regenie" % nm))).read()
    # --minINFO
    full = open(str(tmp_path / "ign_Y1.regenie")).read().splitlines()[1:]
    for thr, tag in (("0.99", "mi_lo"), ("1.01", "mi_hi")):          # example.bgen was made from hard calls: INFO == 1
        run(base + ["--phenoFile", d + "/phenotype.txt", "--covarFile", d + "/covariates.txt", "--minINFO", thr, "--out",
                    str(tmp_path / tag)])
        kept = open(str(tmp_path / (tag + "_Y1.regenie"))).read().splitlines()[1:]
        assert kept == [l for l in full if float(l.split()[6]) >= float(thr)]
    assert len(kept) == 0
    r = subprocess.run([RGB] + base + ["--phenoFile", d + "/phenotype.txt", "--phenoCol", "nope", "--out", str(tmp_path / "x")],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "column 'nope' was not found" in r.stdout


def test_apply_rint(tmp_path, golden_dir):
    """--apply-rint (rank-based inverse normal transform with average ranks for ties, src/Pheno.cpp:1937-2010): the Step-1
    .loco files vs the oracle run on transformed phenotypes."""
    prefix = os.path.join(golden_dir, "example_3chr")
    pheno, covar = golden_dir + "/phenotype.txt", golden_dir + "/covariates.txt"
    out1 = str(tmp_path / "fit")
    run(["--step", "1", "--bed", prefix, "--phenoFile", pheno, "--covarFile", covar, "--bsize", "100", "--apply-rint",
         "--out", out1])
    pb = helpers.Problem(prefix, pheno, covar, 100, rint=True)
    def gen():
        for b in range(len(pb.blocks)):
            yield pb.oracle_block(b)[0]
    ref = step1.run_step1_qt(gen(), pb.blocks, pb.prep, pb.fold_sizes, pb.M)
    for ph in range(pb.prep.Y.shape[1]):
        rf = str(tmp_path / ("ref_%d.loco" % (ph + 1)))
        step1.write_loco(rf, pb.keys, pb.prep.in_analysis, pb.prep.mask[:, ph], ref["loco"][ph])
        compare_token_files(out1 + "_%d.loco" % (ph + 1), rf, exact_cols=1)
    # and it changes the result
    run(["--step", "1", "--bed", prefix, "--phenoFile", pheno, "--covarFile", covar, "--bsize", "100", "--out",
         str(tmp_path / "plain")])
    assert open(out1 + "_1.loco").read() != open(str(tmp_path / "plain_1.loco")).read()


def test_categorical_covariates(tmp_path, golden_dir):
    """--catCovarList: a 3-level string covariate gives exactly the run with its K-1 indicator columns written by hand;
    too many levels are refused like the reference (check_categories, src/Pheno.cpp:985-1010)."""
    d = golden_dir
    rows = [l.split() for l in open(d + "/covariates.txt")]
    rng = np.random.default_rng(2)
    lv = ["siteA", "siteB", "siteC"]
    cat = [lv[k] for k in rng.integers(0, 3, len(rows) - 1)]
    cat[0], cat[1], cat[2] = "siteA", "siteB", "siteC"                 # order of first appearance = A, B, C
    with open(tmp_path / "cov_cat.txt", "w") as fh:
        fh.write(" ".join(rows[0] + ["SITE"]) + "\n")
        for r, c in zip(rows[1:], cat):
            fh.write(" ".join(r + [c]) + "\n")
    with open(tmp_path / "cov_dummy.txt", "w") as fh:
        fh.write(" ".join(rows[0] + ["SITE_1", "SITE_2"]) + "\n")
        for r, c in zip(rows[1:], cat):
            fh.write(" ".join(r + [str(int(c == "siteB")), str(int(c == "siteC"))]) + "\n")
    base = ["--step", "2", "--bed", d + "/example", "--phenoFile", d + "/phenotype.txt", "--bsize", "500", "--ignore-pred"]
    run(base + ["--covarFile", str(tmp_path / "cov_cat.txt"), "--catCovarList", "SITE", "--out", str(tmp_path / "cat")])
    run(base + ["--covarFile", str(tmp_path / "cov_dummy.txt"), "--out", str(tmp_path / "dum")])
    for nm in ("Y1", "Y2"):
        a = open(str(tmp_path / ("cat_%s.regenie" % nm))).read()
        assert a == open(str(tmp_path / ("dum_%s.regenie" % nm))).read() and a.count("\n") > 900
    r = subprocess.run([RGB] + base + ["--covarFile", str(tmp_path / "cov_cat.txt"), "--catCovarList", "SITE", "--maxCatLevels", "2",
                                       "--out", str(tmp_path / "x")], capture_output=True, text=True)
    assert r.returncode != 0 and "too many categories for covariate: SITE (=3)" in r.stdout


def test_gz_prs_write_samples_and_gz_inputs(tmp_path, golden_dir):
    """--gz / --print-prs / --use-prs / --write-samples and .gz inputs (src/Files.cpp:39-150, src/Data.cpp:1795-1922,
    src/Pheno.cpp:1238-1345, :1538-1576): compressed outputs hold exactly the text of the plain run, the .prs row is the
    all-chromosome sum (= the LOCO row of a chromosome that is absent from the data), Step 2 reads .gz prediction and
    phenotype files, and --use-prs equals a run on .loco files whose rows all hold the PRS."""
    import gzip
    import shutil
    d = golden_dir
    pheno, covar = d + "/phenotype.txt", d + "/covariates.txt"
    base = ["--step", "1", "--bed", d + "/example_3chr", "--phenoFile", pheno, "--covarFile", covar, "--bsize", "100"]
    o_plain, o_gz = str(tmp_path / "plain"), str(tmp_path / "gz")
    run(base + ["--out", o_plain])
    log = run(base + ["--out", o_gz, "--gz", "--print-prs"])
    assert "List of files with whole genome PRS written to" in log
    for k in (1, 2):
        plain = open(o_plain + "_%d.loco" % k).read()
        assert gzip.open(o_gz + "_%d.loco.gz" % k, "rt").read() == plain
        rows = {l.split()[0]: l.split()[1:] for l in plain.splitlines()}
        prs = gzip.open(o_gz + "_%d.prs.gz" % k, "rt").read().splitlines()
        assert prs[0] == plain.splitlines()[0] and len(prs) == 2
        t = prs[1].split()
        assert t[0] == "0" and t[1:] == rows["7"]               # chromosome 7 is not in example_3chr: its LOCO row is the full sum
        assert t[1:] != rows["2"]
    pl = [l.split() for l in open(o_gz + "_pred.list")]
    assert [os.path.basename(p[1]) for p in pl] == ["gz_1.loco.gz", "gz_2.loco.gz"]
    pr = [l.split() for l in open(o_gz + "_prs.list")]
    assert [p[0] for p in pr] == ["Y1", "Y2"] and [os.path.basename(p[1]) for p in pr] == ["gz_1.prs.gz", "gz_2.prs.gz"]

    # Step 2: plain inputs vs .gz inputs + --gz output + --write-samples
    for f in ("phenotype.txt", "covariates.txt"):
        with open(d + "/" + f, "rb") as src, gzip.open(tmp_path / (f + ".gz"), "wb") as dst:
            shutil.copyfileobj(src, dst)
    s2 = ["--step", "2", "--bed", d + "/example_3chr", "--bsize", "200"]
    run(s2 + ["--phenoFile", pheno, "--covarFile", covar, "--pred", o_plain + "_pred.list", "--out", str(tmp_path / "s2a")])
    run(s2 + ["--phenoFile", str(tmp_path / "phenotype.txt.gz"), "--covarFile", str(tmp_path / "covariates.txt.gz"), "--pred",
              o_gz + "_pred.list", "--out", str(tmp_path / "s2b"), "--gz", "--write-samples", "--print-pheno"])
    for nm in ("Y1", "Y2"):
        assert gzip.open(str(tmp_path / "s2b") + "_%s.regenie.gz" % nm, "rt").read() == open(str(tmp_path / "s2a") + "_%s.regenie" % nm).read()
        ids = open(str(tmp_path / "s2b") + "_%s.regenie.ids" % nm).read().split("\n")
        assert ids[0] == nm + "\tNA" and len(ids) == 501 and ids[1] == "1\t1" and not ids[-1].endswith("\n")

    # --use-prs == LOCO files whose every row is the PRS row
    lst = tmp_path / "fake_pred.list"
    with open(lst, "w") as fl:
        for k, nm in ((1, "Y1"), (2, "Y2")):
            prs = gzip.open(o_gz + "_%d.prs.gz" % k, "rt").read().splitlines()
            f = tmp_path / ("fake_%d.loco" % k)
            with open(f, "w") as fh:
                fh.write(prs[0] + "\n")
                for c in range(1, 24):
                    fh.write(str(c) + " " + prs[1].split(" ", 1)[1] + "\n")
            fl.write("%s %s\n" % (nm, f))
    run(s2 + ["--phenoFile", pheno, "--covarFile", covar, "--pred", str(lst), "--out", str(tmp_path / "s2c")])
    log = run(s2 + ["--phenoFile", pheno, "--covarFile", covar, "--pred", o_gz + "_prs.list", "--use-prs", "--out", str(tmp_path / "s2d")])
    assert " * PRS predictions : [" in log
    for nm in ("Y1", "Y2"):
        a = open(str(tmp_path / "s2c") + "_%s.regenie" % nm).read()
        assert a == open(str(tmp_path / "s2d") + "_%s.regenie" % nm).read()
        assert a != open(str(tmp_path / "s2a") + "_%s.regenie" % nm).read()


def test_bgen_index_file_is_used_and_gives_the_scan_result(tmp_path, golden_dir):
    """example_3chr.bgen with its bgenix index (example_3chr.bgen.bgi, next to it) vs a copy of the file without one."""
    import shutil
    d = golden_dir
    shutil.copy(d + "/example_3chr.bgen", tmp_path / "noidx.bgen")
    common = ["--step", "2", "--phenoFile", d + "/phenotype.txt", "--covarFile", d + "/covariates.txt", "--bsize", "200",
              "--ignore-pred", "--sample", d + "/example_3chr.sample"]
    la = run(common + ["--bgen", d + "/example_3chr.bgen", "--out", str(tmp_path / "a")])
    lb = run(common + ["--bgen", str(tmp_path / "noidx.bgen"), "--out", str(tmp_path / "b")])
    lc = run(common + ["--bgen", str(tmp_path / "noidx.bgen"), "--bgi", d + "/example_3chr.bgen.bgi", "--chr", "2", "--out", str(tmp_path / "c")])
    assert "-index bgi file [" in la and "-index bgi file [" not in lb and "-index bgi file [" in lc
    for nm in ("Y1", "Y2"):
        a = open(str(tmp_path / "a") + "_%s.regenie" % nm).read()
        assert a == open(str(tmp_path / "b") + "_%s.regenie" % nm).read() and len(a.splitlines()) > 400
        c = open(str(tmp_path / "c") + "_%s.regenie" % nm).read().splitlines()
        assert c[1:] == [l for l in a.splitlines()[1:] if l.startswith("2 ")]


def test_gpu_inflate_equals_host_inflate(tmp_path, golden_dir):
    """--gpu-inflate (rg_bgen_inflate: one warp per zlib stream, csrc/inflate_core.h) must give byte-identical results
    to the default host-zlib path: the reference's documented BT / Firth command on example.bgen and a QT run on the
    three-chromosome fileset with a sample subset."""
    d = golden_dir
    pred = _write_zero_loco(tmp_path, d, ["Y1", "Y2"])
    bt = ["--step", "2", "--bgen", d + "/example.bgen", "--covarFile", d + "/covariates.txt", "--phenoFile",
          d + "/phenotype_bin.txt", "--remove", d + "/fid_iid_to_remove.txt", "--bsize", "200", "--bt", "--firth",
          "--approx", "--pThresh", "0.01", "--pred", pred]
    run(bt + ["--out", str(tmp_path / "host")])
    log = run(bt + ["--out", str(tmp_path / "dev"), "--gpu-inflate"])
    assert "inflated on the GPU" in log
    for nm in ("Y1", "Y2"):
        a = open(str(tmp_path / "host") + "_%s.regenie" % nm).read()
        assert a == open(str(tmp_path / "dev") + "_%s.regenie" % nm).read() and len(a.splitlines()) == 1001
    qt = ["--step", "2", "--bgen", d + "/example_3chr.bgen", "--sample", d + "/example_3chr.sample", "--phenoFile",
          d + "/phenotype.txt", "--covarFile", d + "/covariates.txt", "--bsize", "77", "--ignore-pred", "--remove",
          d + "/fid_iid_to_remove.txt"]
    run(qt + ["--out", str(tmp_path / "qh")])
    run(qt + ["--out", str(tmp_path / "qd"), "--gpu-inflate"])
    for nm in ("Y1", "Y2"):
        a = open(str(tmp_path / "qh") + "_%s.regenie" % nm).read()
        assert a == open(str(tmp_path / "qd") + "_%s.regenie" % nm).read() and len(a.splitlines()) > 400


def test_step1_on_bgen_dosages_matches_oracle(tmp_path, golden_dir):
    """`rgb200 --step 1 --bgen example.bgen` (readChunkFromBGENFileToG_fast, src/Geno.cpp:1574-1699 -> the dense FP64
    level-0 route) vs the oracle run on the dosages the oracle's own BGEN reader decodes: .loco files token by token."""
    from oracle import bgen as obgen
    from oracle import plink, prep
    pheno, covar = golden_dir + "/phenotype.txt", golden_dir + "/covariates.txt"
    out1 = str(tmp_path / "fit_bgen")
    log = run(["--step", "1", "--bgen", golden_dir + "/example.bgen", "--phenoFile", pheno, "--covarFile", covar,
               "--bsize", "200", "--out", out1])
    assert "<- min value" in log
    bg = obgen.Bgen(golden_dir + "/example.bgen")
    vs = list(bg.variants())
    keys = bg.sample_ids if bg.sample_ids else None
    # sample keys as the driver derives them for a .bgen without --sample: the ids embedded in the file ("FID_IID" form below)
    pb = helpers.Problem(golden_dir + "/example", pheno, covar, 200)          # same 500 samples / phenotypes / folds as example.bed
    chrom = np.array([plink.chr_str_to_int(v[0]) for v in vs])
    blocks = prep.set_blocks(chrom, 200)
    pr = pb.prep

    def gen():
        for c, s, bs in blocks:
            g = np.stack([obgen.dosage(v[4].astype(np.float64), v[5].astype(np.float64), v[6])[0] for v in vs[s:s + bs]])
            yield plink.mean_impute_block(g, pr.in_analysis)[0]
    o = step1.run_step1_qt(gen(), blocks, pr, pb.fold_sizes, len(vs))
    for ph in range(2):
        ref = str(tmp_path / ("oracle_bgen_%d.loco" % (ph + 1)))
        step1.write_loco(ref, pb.keys, pr.in_analysis, pr.mask[:, ph], o["loco"][ph])
        compare_token_files(out1 + "_%d.loco" % (ph + 1), ref, exact_cols=1)
