"""GPU parity of the on-device .pgen decode (rg_pgen_decode): the rows the kernels write must equal, bit for
bit, the PLINK 1 coding of the calls that were written (oracle/pgen.py reads the same file back), for every hard-call record
type, difflists of more than 32 groups, LD records whose base is fetched as an extra record, and a partial last word; the
level-0 predictors and the Step-2 statistics computed from device-decoded rows must be IDENTICAL to those from host rows."""
import numpy as np
import pytest

import helpers
from oracle import pgen, prep

pytestmark = pytest.mark.gpu

BED_OF = np.array([3, 2, 0, 1], dtype=np.uint8)          # ALT count 0 / 1 / 2 / missing -> PLINK 1 code (ref-last)


def pack_rows(g):
    M, N = g.shape
    c = np.zeros((M, (N + 3) // 4 * 4), dtype=np.uint8)
    c[:, :N] = BED_OF[g]
    c = c.reshape(M, -1, 4)
    return (c[:, :, 0] | (c[:, :, 1] << 2) | (c[:, :, 2] << 4) | (c[:, :, 3] << 6)).astype(np.uint8)


def make_file(tmp_path, N, M, storage):
    from test_host_cpu import big_pgen_calls
    g = big_pgen_calls(N, M)
    pfx = str(tmp_path / ("p%d" % N))
    types = helpers.write_pgen(pfx, g, storage=storage)
    assert set(types) >= set(range(8))
    return g, pgen.Pgen(pfx + ".pgen")


def step2_handle(N, P, bs, seed=3):
    from regenie_b200 import capi
    rng = np.random.default_rng(seed)
    X = np.linalg.qr(np.column_stack([np.ones(N), rng.normal(size=(N, 2))]))[0]
    mask = np.ones((N, P), dtype=np.uint8)
    mask[rng.random((N, P)) < 0.03] = 0
    st = capi.Step2(X, mask, np.ones(N, dtype=np.uint8), N, bs)
    res = rng.normal(size=(N, P)) * mask
    res -= X @ (X.T @ res)
    st.set_chr(res * mask, np.ones(P))
    return st


@pytest.mark.parametrize("N,M,bs,storage", [(701, 40, 7, 5), (33333, 40, 9, 6), (70001, 30, 30, 2)])
def test_pgen_rows_decoded_on_the_device_are_bit_exact(tmp_path, N, M, bs, storage):
    from regenie_b200 import capi
    g, pg = make_file(tmp_path, N, M, storage)
    want = pack_rows(g)
    st = step2_handle(N, 2, 32)
    for first in range(0, M, bs):
        vs = list(range(first, min(M, first + bs)))
        b = helpers.gather_pgen(pg, vs)
        rows, stride = capi.pgen_decode(st, n_file=N, block_id=first // bs, **b)
        assert stride % 16 == 0 and stride >= (N + 3) // 4
        got = capi.debug_fetch(st, "pgen_rows", np.uint8, len(vs) * stride).reshape(len(vs), stride)
        assert np.array_equal(got[:, :want.shape[1]], want[vs]), first
        assert not got[:, want.shape[1]:].any()
        # the score test on the device rows == on host rows of the same calls
        a = st.block_bed(rows, row_stride=stride, bs=len(vs))
        c = st.block_bed(want[vs])
        for k in a:
            assert np.array_equal(a[k], c[k], equal_nan=True), k
    st.close()


def test_level0_from_device_decoded_pgen_rows_is_identical(tmp_path):
    """Step-1 handle: decode on the lane's stream, consumed by the next rg_l0_block_bed; W bit-identical to host rows."""
    from regenie_b200 import capi
    N, M, bs = 4099, 60, 20
    g, pg = make_file(tmp_path, N, M, 5)
    g = g.copy()
    keep = [v for v in range(M) if len(np.unique(g[v][g[v] < 3])) > 1 and (g[v] == 3).mean() < 0.5]   # polymorphic ones
    want = pack_rows(g)
    rng = np.random.default_rng(1)
    X = np.linalg.qr(np.column_stack([np.ones(N), rng.normal(size=(N, 2))]))[0]
    Y = rng.normal(size=(N, 2)); Y -= X @ (X.T @ Y); Y /= np.linalg.norm(Y, axis=0) / np.sqrt(N - 3)
    mask = np.ones((N, 2), dtype=np.uint8)
    folds = [N // 5 + (1 if k < N % 5 else 0) for k in range(5)]
    lam = [bs * (1 - h) / h for h in (0.01, 0.25, 0.5, 0.75, 0.99)]
    blocks = [keep[i:i + bs] for i in range(0, len(keep) - bs + 1, bs)]
    assert len(blocks) >= 2
    Ws = []
    for mode in ("host", "device"):
        st = capi.Step1(X, Y, mask, np.ones(N, dtype=np.uint8), folds, lam, [float(N)] * 2, N, bs, len(blocks))
        for bi, vs in enumerate(blocks):
            if mode == "host":
                st.l0_block_bed(want[vs], len(vs), bi)
            else:
                rows, stride = capi.pgen_decode(st, n_file=N, block_id=bi, **helpers.gather_pgen(pg, vs))
                st.l0_block_bed(rows, len(vs), bi, row_stride=stride)
        assert st.status() == 0, capi.lib().rg_last_error().decode()
        Ws.append([st.fetch_W(bi, ph) for bi in range(len(blocks)) for ph in range(2)])
        st.close()
    for a, b in zip(*Ws):
        assert np.array_equal(a, b)


def test_malformed_pgen_records_are_reported(tmp_path):
    from regenie_b200 import capi
    N, M = 5003, 20
    g, pg = make_file(tmp_path, N, M, 5)
    st = step2_handle(N, 1, 32)
    b = helpers.gather_pgen(pg, list(range(M)))
    # a difflist record cut short: the kernel must flag it, not read past the record
    t = [i for i, ty in enumerate(b["rec_type"]) if ty in (2, 3, 4, 6, 7) and b["rec_len"][i] > 40][0]
    bad = dict(b); bad["rec_len"] = list(b["rec_len"]); bad["rec_len"][t] = b["rec_len"][t] // 2
    with pytest.raises(capi.RgError, match="malformed .pgen record"):
        capi.pgen_decode(st, n_file=N, **bad)
    # a sample index beyond the file's sample count
    with pytest.raises(capi.RgError, match="malformed .pgen record"):
        capi.pgen_decode(st, n_file=N // 2, **b)
    # host-side table checks
    bad = dict(b); bad["own"] = list(b["own"]); bad["own"][0] = len(b["rec_off"])
    with pytest.raises(capi.RgError, match="record index out of range"):
        capi.pgen_decode(st, n_file=N, **bad)
    rows, stride = capi.pgen_decode(st, n_file=N, **b)       # the handle still works afterwards
    got = capi.debug_fetch(st, "pgen_rows", np.uint8, M * stride).reshape(M, stride)
    assert np.array_equal(got[:, :(N + 3) // 4], pack_rows(g))
    st.close()
