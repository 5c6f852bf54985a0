"""CPU: pin the oracle (oracle/pq_oracle.py) against fixtures produced by the reference's own code
(tests/golden/make_golden.py -> reference_golden.npz)."""
import os

import numpy as np
import pytest

from oracle import pq_oracle as O


def test_l2Ns_and_nbits(golden):
    assert [O.l2Ns(int(l)) for l in golden["l2Ns_l"]] == list(golden["l2Ns_v"])
    assert [np.dtype(O.nbits2dtype(int(b))).itemsize for b in golden["nbits"]] == list(golden["nbits_itemsize"])
    with pytest.raises(ValueError):
        O.nbits2dtype(65)


@pytest.mark.parametrize("i", range(5))
def test_encode_matches_reference_keops_path(golden, i):
    X, cent = golden[f"enc{i}_X"], golden[f"enc{i}_cent"]
    ref = golden[f"enc{i}_codes_keops"]
    got = O.pq_encode(X, cent)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert np.array_equal(got, ref)                      # bit-exact vs sa_encode_4d_keops


@pytest.mark.parametrize("i", range(4))
def test_cdist_twin_differs_only_on_near_ties(golden, i):
    X, cent = golden[f"enc{i}_X"], golden[f"enc{i}_cent"]
    a, b = golden[f"enc{i}_codes_keops"], golden[f"enc{i}_codes_cdist"]
    n_mis, n_bad = O.encode_mismatch_is_near_tie(X, cent, a, b, rel=1e-5)
    assert n_bad == 0
    assert n_mis <= a.size * 1e-3


@pytest.mark.parametrize("i", range(5))
def test_decode_matches_reference(golden, i):
    got = O.pq_decode(golden[f"enc{i}_codes_keops"], golden[f"enc{i}_cent"])
    assert got.dtype == golden[f"enc{i}_decoded"].dtype
    assert np.array_equal(got, golden[f"enc{i}_decoded"])


def test_tie_goes_to_lowest_index(golden):
    got = O.pq_encode(golden["tie_X"], golden["tie_cent"])
    assert np.array_equal(got, golden["tie_codes"])
    assert (got == 2).all()


def test_wide_codes(golden):
    got = O.pq_encode(golden["wide_X"], golden["wide_cent"], out_dtype=np.int32)
    assert np.array_equal(got, golden["wide_codes"])
    assert got.max() > 255


@pytest.mark.parametrize("i", range(4))
def test_decode_attention_matches_reference_invariant(golden, i):
    g = lambda k: golden[f"att{i}_{k}"]
    bs, nh, nh_k, nk, r, d, M, C = [int(v) for v in g("shape")]
    got = O.pq_decode_attn(g("q"), g("kc"), g("vc"), g("kcent"), g("vcent"), g("kres"), g("vres"), r)
    np.testing.assert_allclose(got, g("out"), atol=2e-6, rtol=1e-5)


def test_split_merge_equals_full(golden):
    g = lambda k: golden[f"att0_{k}"]
    nk, r = int(g("shape")[3]), int(g("shape")[4])
    args = (g("q"), g("kc"), g("vc"), g("kcent"), g("vcent"), g("kres"), g("vres"), r)
    full, lse = O.pq_decode_attn(*args, return_lse=True)
    for bounds in ([(0, nk)], [(0, 7), (7, 7), (7, nk)], [(i * 19, min(nk, (i + 1) * 19)) for i in range(16)]):
        outs, lses = O.split_partials(*args, bounds)
        np.testing.assert_allclose(O.lse_merge(outs, lses), full[:, :, 0], atol=1e-6, rtol=1e-5)
    # total lse is the lse of lses
    m = lses.max(-1)
    np.testing.assert_allclose(np.log(np.exp(lses - m[..., None]).sum(-1)) + m, lse, atol=1e-5)


def test_page_pool_round_trip():
    rng = np.random.default_rng(0)
    for T in (0, 1, 63, 64, 65, 200):
        vc = rng.integers(0, 256, (2, 3, T, 64), dtype=np.uint8)
        pool, table = O.build_page_pool(vc, 64)
        assert pool.shape == (((T + 63) // 64) * 6, 64, 64)
        assert np.array_equal(O.gather_pages(pool, table, T), vc)
        if T:
            assert table[0, 0, 0] == 0 and table[0, 1, 0] == 1 and table[1, 0, 0] == 3   # chunk-major, b, h
            if T % 64:
                assert (pool[table[0, 0, -1], :, T % 64:] == 0).all()                    # zero padded tail


def test_cache_policy_oracle():
    rng = np.random.default_rng(1)
    kw = dict(bs=1, nh=4, num_key_value_heads=2, M=16, layer_num=1, d=32)
    c = O.DynamicPQCacheOracle(**kw)
    cent = rng.standard_normal((16, 256, 2)).astype(np.float32)
    c.set_cent(cent, cent)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    out = c.prefill(f(1, 4, 10, 32), f(1, 2, 10, 32), f(1, 2, 10, 32), 0)
    assert out.shape == (1, 4, 10, 32) and c.key_cache[0].shape == (1, 2, 10, 16) and c.residualed_tokens[0] == 0
    for step in range(40):
        o = c.decoding(f(1, 4, 1, 32), f(1, 2, 1, 32), f(1, 2, 1, 32), 0)
        assert o.shape == (1, 4, 1, 32) and np.isfinite(o).all()
    # window is d=32 long: one flush happened at step 32
    assert c.key_cache[0].shape[2] == 10 + 32 and c.residualed_tokens[0] == 8 and c.seen_tokens[0] == 50
    p = O.PagedPQCacheOracle(page_size=8, extended_residual_size=16, **kw)
    p.set_cent(cent, cent)
    p.prefill(f(1, 4, 10, 32), f(1, 2, 10, 32), f(1, 2, 10, 32), 0)
    assert p.value_cache[0].shape == (1, 2, 16, 10)
    for step in range(20):
        p.decoding_with_pages(f(1, 4, 1, 32), f(1, 2, 1, 32), f(1, 2, 1, 32), 0)
    assert p.key_cache[0].shape[2] == 10 + 8 and p.value_cache[0].shape[3] == 18 and p.residualed_tokens[0] == 12
    assert p.seen_tokens[0] == 30


def test_c_oracle_agrees_with_numpy_oracle_and_golden(golden):
    from oracle import c_oracle as CO
    for i in range(4):
        X, cent = golden[f"enc{i}_X"], golden[f"enc{i}_cent"]
        assert np.array_equal(CO.pq_encode(X, cent), golden[f"enc{i}_codes_keops"])
        assert np.array_equal(CO.pq_decode(golden[f"enc{i}_codes_keops"], cent),
                              golden[f"enc{i}_decoded"].astype(np.float32))
    for i in range(4):
        g = lambda k: golden[f"att{i}_{k}"]
        got = CO.pq_decode_attn(g("q"), g("kc"), g("vc"), g("kcent"), g("vcent"), g("kres"), g("vres"), int(g("shape")[4]))
        np.testing.assert_allclose(got, g("out"), atol=3e-6, rtol=1e-5)


# ------------------------------------------------------------------------------------------------ outlier side store
# (extension; no reference counterpart: these pin the oracle's own definition, section A.6 of oracle/pq_oracle.py)


def test_outlier_oracle_definition():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((1, 2, 40, 128)).astype(np.float16)
    X[..., 3] = 40.0
    C = rng.standard_normal((64, 256, 2)).astype(np.float16).astype(np.float32)
    codes, idx, val = O.pq_encode_outliers(X, C, 2)
    assert idx.shape == (1, 2, 40, 2) and idx.dtype == np.uint8 and val.dtype == np.float16
    assert (idx[..., 0] == 3).all()                                   # the spiked channel is always the largest
    # codes are the plain encoder's codes of the masked vectors
    Xm = X.astype(np.float32).copy()
    np.put_along_axis(Xm, idx.astype(np.int64), 0.0, axis=-1)
    assert np.array_equal(codes, O.pq_encode(Xm, C))
    # outlier dims are reconstructed to fp16 rounding of the delta, other dims exactly as plain PQ of the masked vector
    xh = O.pq_decode_outliers(codes, C, idx, val)
    at = np.take_along_axis(xh, idx.astype(np.int64), -1) - np.take_along_axis(X.astype(np.float32), idx.astype(np.int64), -1)
    assert np.abs(at).max() <= 2.0 ** -5                              # |delta| < 64 -> half an fp16 ulp <= 2^-5
    rest = xh.copy(); base = O.pq_decode(codes, C).copy()
    np.put_along_axis(rest, idx.astype(np.int64), 0.0, -1); np.put_along_axis(base, idx.astype(np.int64), 0.0, -1)
    assert np.array_equal(rest, base)
    # k_out = 0: the plain path
    c0, i0, v0 = O.pq_encode_outliers(X, C, 0)
    assert np.array_equal(c0, O.pq_encode(X, C)) and i0.shape[-1] == 0
    assert np.array_equal(O.pq_decode_outliers(c0, C, i0, v0), O.pq_decode(c0, C))


def test_outlier_oracle_attention_reduces_to_plain():
    inp = O.make_inputs(bs=1, nh=4, nh_k=2, nk=200, seed=3)
    a = O.pq_decode_attn(inp["q"], inp["kc"], inp["vc"], inp["kcent"], inp["vcent"], inp["kres"], inp["vres"], 9)
    zi = np.zeros((1, 2, 200, 2), np.uint8); zv = np.zeros((1, 2, 200, 2), np.float16)
    b = O.pq_decode_attn_outliers(inp["q"], inp["kc"], inp["vc"], inp["kcent"], inp["vcent"], inp["kres"], inp["vres"], 9,
                                  kout=(zi, zv), vout=(zi, zv))
    np.testing.assert_allclose(a, b, atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize("name,k_out", [("m64k2", 2), ("m32k1", 1), ("m64k4", 4)])
def test_outlier_oracle_reproduces_committed_vectors(name, k_out):
    """tests/golden/outlier_golden.npz (made by tests/golden/make_outlier_golden.py from this oracle): the definition must not drift."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "outlier_golden.npz"))
    codes, idx, val = O.pq_encode_outliers(g[f"{name}_X"], g[f"{name}_C"], k_out)
    assert np.array_equal(codes, g[f"{name}_codes"]) and np.array_equal(idx, g[f"{name}_idx"]) and np.array_equal(val, g[f"{name}_val"])
    assert idx[0, 0, 0].tolist() == list(range(k_out)) and idx[0, 0, 1].tolist() == list(range(k_out))      # all-equal rows: lowest dims
    assert np.array_equal(O.pq_decode_outliers(codes, g[f"{name}_C"], idx, val), g[f"{name}_recon"])
    attn = O.pq_decode_attn_outliers(g[f"{name}_q"], codes, g[f"{name}_vcodes"], g[f"{name}_C"], g[f"{name}_C"], g[f"{name}_kres"],
                                     g[f"{name}_vres"], 9, kout=(idx, val), vout=(g[f"{name}_vidx"], g[f"{name}_vval"]))
    np.testing.assert_allclose(attn, g[f"{name}_attn"], atol=1e-6, rtol=1e-6)
