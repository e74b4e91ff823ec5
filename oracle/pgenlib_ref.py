"""ctypes face of oracle/_ref/libpgenlib_ref.so: the REFERENCE's vendored pgenlib (external_libs/pgenlib, compiled from the
sources where they lie by oracle/build_native.py) behind the two calls oracle/ref_pgenlib/pgen_ref_shim.cpp exports.
Test infrastructure: used by tests/ (to pin oracle/pgen.py, host/pgen.cpp and csrc/pgen_core.h on reference code) and by
bench.py's cpu_baseline leg of the .pgen decode; never by the product path."""
import ctypes as C
import os

import numpy as np

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libpgenlib_ref.so")
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.pgref_read_hardcalls.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                           C.c_void_p, C.c_char_p, C.c_int]
        L.pgref_validate.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        _lib = L
    return _lib


def read_hardcalls(path, n_raw, v0, nv, subset=None, timing=False):
    """[nv, n] float64 as PgenReader::ReadHardcalls(..., allele_idx = 1) fills them (src/Geno.cpp:1798): ALT counts, -3 =
    missing; subset = 0-based file indices of the samples to keep (the reference passes them 1-based)."""
    sub = None if subset is None else np.ascontiguousarray(np.asarray(subset) + 1, dtype=np.int32)
    n = n_raw if sub is None else sub.size
    out = np.empty((nv, n), dtype=np.float64)
    err = C.create_string_buffer(512)
    sec = C.c_double(0.0)
    rc = lib().pgref_read_hardcalls(path.encode(), n_raw, None if sub is None else sub.ctypes.data, 0 if sub is None else sub.size,
                                    v0, nv, out.ctypes.data, C.addressof(sec), err, 512)
    if rc != 0:
        raise RuntimeError("pgenlib: " + err.value.decode())
    return (out, sec.value) if timing else out


def validate(path):
    """pgenlib's PgrValidate over the whole file; raises with the library's message when the file is not well-formed."""
    err = C.create_string_buffer(512)
    rc = lib().pgref_validate(path.encode(), err, 512)
    if rc != 0:
        raise RuntimeError("pgenlib validation failed (%d): %s" % (rc, err.value.decode()))
