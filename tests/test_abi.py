"""CPU-side check: the C-ABI library loads and exports every symbol include/rg_b200.h declares."""
import os
import re

from regenie_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "rg_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[A-Za-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported():
    L = capi.lib()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), "librg_b200.so does not export " + s
    assert set(capi.EXPORTS) <= set(syms)


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must fail loudly (never fall back)."""
    import numpy as np
    L = capi.lib()
    assert L.rg_version().decode().startswith("regenie_b200")
    if L.rg_device_count() > 0:
        return
    X = np.ones((8, 1)); Y = np.zeros((8, 1)); m = np.ones((8, 1), dtype=np.uint8)
    try:
        capi.Step1(X, Y, m, np.ones(8, dtype=np.uint8), [4, 4], [1.0], [8.0], 8, 4, 1)
    except capi.RgError as e:
        assert "no CUDA device" in str(e)
    else:
        raise AssertionError("Step1 creation succeeded without a GPU")
