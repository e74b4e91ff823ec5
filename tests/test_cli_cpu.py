"""Host driver logic that needs no GPU: option parsing and the reference's failure shape (`ERROR: <msg>` on stdout + log,
non-zero exit, src/Regenie.cpp:67-92), and the refusal to run without a CUDA device (no CPU fallback)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RGB = os.path.join(ROOT, "regenie_b200", "rgb200")


def run(args, cwd):
    return subprocess.run([RGB] + args, capture_output=True, text=True, timeout=60, cwd=cwd)


def test_cli_errors_have_the_reference_shape(tmp_path, golden_dir):
    d = golden_dir
    base = ["--bed", d + "/example", "--phenoFile", d + "/phenotype.txt", "--bsize", "100", "--out", str(tmp_path / "o")]
    cases = [
        (base, "specify which mode regenie should be running using option '--step'"),
        (["--step", "1"] + base[:-4] + ["--out", str(tmp_path / "o")], "must specify the block size using '--bsize'"),
        (["--step", "2"] + base, "must specify --pred if using --step 2"),
        (["--step", "1", "--skat"] + base, "outside the hot path covered by rgb200"),
        (["--step", "2", "--pred", "x", "--bed", d + "/example", "--bgen", d + "/example.bgen"] + base[2:], "specify only one genotype input"),
        (["--step", "2", "--firth", "--bt", "--pred", "x"] + base, "exact Firth"),
        (["--step", "2", "--split-l0", "p,2", "--pred", "x"] + base, "only work in step 1"),
        (["--step", "2", "--pred", "x", "--range", "1:100"] + base, "wrong format for --range (must be CHR:MINPOS-MAXPOS)."),
        (["--step", "2", "--pred", "x", "--range", "Z:1-100"] + base, "unrecognized chromosome in --range."),
        (["--step", "1", "--setl0", "0,0.5"] + base, "must specify values for --l0 in (0,1)."),
        (["--step", "2", "--pred", "x", "--test", "overdominant"] + base, "unrecognized argument for option --test"),
        (["--step", "1", "--test", "dominant"] + base, "can only use --test in step 2"),
        (["--step", "1", "--covarColList", "V{1:x}"] + base, "invalid string expansion (=V{1:x})."),
        (["--step", "1", "--setl1", "0.5,1"] + base, "must specify values for --l1 in (0,1)."),
        (["--step", "2", "--pred", "x", "--write-samples", "--bgen", d + "/example.bgen"] + base[2:],
         "must specify sample file (using --sample) if writing sample IDs to file."),
    ]
    for args, msg in cases:
        r = run(args, str(tmp_path))
        assert r.returncode != 0, args
        assert ("ERROR: " in r.stdout) and (msg in r.stdout), (args, r.stdout[-400:])


def test_version_and_help(tmp_path):
    r = run(["--version"], str(tmp_path))
    assert r.returncode == 0 and r.stdout.startswith("rgb200 (")
    r = run(["--help"], str(tmp_path))
    assert r.returncode == 0 and "--gpu-inflate" in r.stdout and "--range" in r.stdout


def test_no_cpu_fallback(tmp_path, golden_dir):
    """Without a CUDA device the driver stops before touching any data (this test only asserts it on GPU-less hosts)."""
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "regenie_b200", "librg_b200.so"))
    if lib.rg_device_count() > 0:
        return
    d = golden_dir
    r = run(["--step", "1", "--bed", d + "/example", "--phenoFile", d + "/phenotype.txt", "--bsize", "100", "--out",
             str(tmp_path / "o")], str(tmp_path))
    assert r.returncode != 0 and "no CUDA device available" in r.stdout and "no CPU fallback" in r.stdout
