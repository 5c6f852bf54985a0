"""GPU parity of the outlier side store (extension — the reference has no outlier code, SURVEY.md section 0.1, so parity
here is against OUR OWN restatement in oracle/pq_oracle.py section A.6: "parity unpinned" for this feature).

Bars: record dims, deltas and codes bit-exact; attention max-abs 2e-3 / rel 1e-2; with the store disabled the path is
bit-identical to the plain one."""
import numpy as np
import pytest
import torch

from oracle import pq_oracle as O

pytestmark = pytest.mark.gpu
ATOL, RTOL = 2e-3, 1e-2


@pytest.fixture(scope="module")
def M():
    import million_b200.ops as ops
    from million_b200 import _lib
    _lib.lib()
    return ops


def _data(seed, bs, nh_k, n, d=128, Mm=64, heavy=True, dtype=torch.float16):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((bs, nh_k, n, d), dtype=np.float32)
    if heavy:                                   # a few fixed outlier channels plus 1 % random spikes
        X[..., [5, d - 13]] *= 12
        X *= np.where(rng.random(X.shape) < 0.01, 20.0, 1.0).astype(np.float32)
    X = torch.from_numpy(X).to(dtype)
    cent = torch.from_numpy(rng.standard_normal((Mm, 256, d // Mm), dtype=np.float32)).to(dtype).float()
    return X, cent


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("k_out", [1, 2, 4, 8])
@pytest.mark.parametrize("d,Mm", [(128, 64), (128, 32), (64, 32)])
def test_split_encode_bit_exact(M, dtype, k_out, d, Mm):
    X, cent = _data(3, 2, 3, 257, d, Mm, dtype=dtype)
    io = np.float16 if dtype == torch.float16 else 'bf16'
    codes, idx, val = O.pq_encode_outliers(X.float().numpy(), cent.numpy(), k_out, io_dtype=io)
    gc, gi, gv = M.pq_encode_outliers(X.cuda(), cent.cuda(), k_out)
    assert np.array_equal(gi.cpu().numpy(), idx)
    assert np.array_equal(gc.cpu().numpy(), codes)
    assert np.array_equal(gv.float().cpu().numpy(), np.asarray(val, dtype=np.float32))


def test_split_ties_pick_lowest_dim(M):
    X = torch.zeros(1, 1, 4, 128, dtype=torch.float16)
    X[0, 0, 0, [9, 40, 100]] = torch.tensor([3.0, -3.0, 3.0], dtype=torch.float16)   # three equal |x|
    X[0, 0, 1, :] = 1.0                                                             # all equal
    X[0, 0, 2, 127] = -2.0
    _, cent = _data(1, 1, 1, 1)
    _, gi, _ = M.pq_encode_outliers(X.cuda(), cent.cuda(), 2)
    assert gi.cpu().numpy()[0, 0].tolist() == [[9, 40], [0, 1], [127, 0], [0, 1]]
    assert np.array_equal(gi.cpu().numpy(), O.outlier_select(X.float().numpy(), 2))


def test_reconstruct_with_outliers(M):
    X, cent = _data(5, 1, 2, 300)
    gc, gi, gv = M.pq_encode_outliers(X.cuda(), cent.cuda(), 2)
    got = M.pq_decode_outliers(gc, cent.half().cuda(), gi, gv).float().cpu().numpy()
    ref = O.pq_decode_outliers(gc.cpu().numpy(), cent.numpy(), gi.cpu().numpy(), gv.float().cpu().numpy())
    np.testing.assert_allclose(got, ref, atol=0, rtol=2e-3)      # one fp16 rounding of (centroid + delta)
    # the point of the store: the reconstruction error of spiky vectors drops
    plain = O.pq_decode(O.pq_encode(X.float().numpy(), cent.numpy()), cent.numpy())
    assert np.abs(got - X.float().numpy()).max() < 0.25 * np.abs(plain - X.float().numpy()).max()


def _attn_case(M, bs, nh, nh_k, nk, r, k_out, v_out, impl, dtype=torch.float16, seed=11, Mm=64):
    rng = np.random.default_rng(seed)
    K, cent = _data(seed, bs, nh_k, nk, Mm=Mm, dtype=dtype)
    V, vcent = _data(seed + 1, bs, nh_k, nk, Mm=Mm, heavy=False, dtype=dtype)
    q = torch.from_numpy(rng.standard_normal((bs, nh, 1, 128), dtype=np.float32)).to(dtype)
    kres = torch.from_numpy(rng.standard_normal((bs, nh_k, 128, 128), dtype=np.float32)).to(dtype)
    vres = torch.from_numpy(rng.standard_normal((bs, nh_k, 128, 128), dtype=np.float32)).to(dtype)
    if k_out:
        kc, ki, kv = M.pq_encode_outliers(K.cuda(), cent.cuda(), k_out)
    else:
        kc, ki, kv = M.pq_encode(K.cuda(), cent.cuda()), None, None
    if v_out:
        vc, vi, vv = M.pq_encode_outliers(V.cuda(), vcent.cuda(), v_out)
    else:
        vc, vi, vv = M.pq_encode(V.cuda(), vcent.cuda()), None, None
    got = M.pq_decode_attn(q.cuda(), kc, vc, cent.to(dtype).cuda(), vcent.to(dtype).cuda(), kres.cuda(), vres.cuda(), r, impl=impl,
                           k_outliers=(ki, kv) if k_out else None, v_outliers=(vi, vv) if v_out else None)
    f = lambda t: t.float().cpu().numpy()
    ref = O.pq_decode_attn_outliers(f(q), kc.cpu().numpy(), vc.cpu().numpy(), cent.to(dtype).float().numpy(), vcent.to(dtype).float().numpy(),
                                    f(kres), f(vres), r, kout=(ki.cpu().numpy(), f(kv)) if k_out else None,
                                    vout=(vi.cpu().numpy(), f(vv)) if v_out else None)
    return f(got), ref


@pytest.mark.parametrize("k_out,v_out", [(2, 0), (1, 1), (4, 2), (0, 2)])
@pytest.mark.parametrize("nh,nh_k", [(8, 2), (4, 4)])
def test_attn_generic_with_outliers(M, k_out, v_out, nh, nh_k):
    from million_b200 import _lib as L
    got, ref = _attn_case(M, 2, nh, nh_k, 700, 17, k_out, v_out, L.IMPL_GENERIC)
    np.testing.assert_allclose(got, ref, atol=ATOL, rtol=RTOL)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("k_out", [1, 2, 3, 4])
@pytest.mark.parametrize("nh,nh_k,nk", [(8, 2, 1500), (4, 2, 333), (4, 4, 2049), (16, 2, 64)])
def test_attn_fast_k_outliers(M, dtype, k_out, nh, nh_k, nk):
    from million_b200 import _lib as L
    got, ref = _attn_case(M, 2, nh, nh_k, nk, 128, k_out, 0, L.IMPL_FAST, dtype=dtype)
    np.testing.assert_allclose(got, ref, atol=ATOL, rtol=RTOL)


@pytest.mark.parametrize("nh,nh_k,nk,k_out", [(8, 2, 1500, 2), (4, 4, 777, 1), (4, 2, 2100, 4)])
def test_attn_fast_k_outliers_m32(M, nh, nh_k, nk, k_out):
    """M = 32 (d_m = 4, the reference's "2-bit"): attn_fast_dm4.cu"""
    from million_b200 import _lib as L
    got, ref = _attn_case(M, 2, nh, nh_k, nk, 100, k_out, 0, L.IMPL_FAST, Mm=32)
    np.testing.assert_allclose(got, ref, atol=ATOL, rtol=RTOL)


def test_attn_fast_k_outliers_long_batch(M):
    """flat scheduling (batch 8, several pieces per CTA) with the side store"""
    from million_b200 import _lib as L
    got, ref = _attn_case(M, 8, 8, 2, 4100, 77, 2, 0, L.IMPL_FAST)
    np.testing.assert_allclose(got, ref, atol=ATOL, rtol=RTOL)


@pytest.mark.parametrize("Mm", [64, 32])
@pytest.mark.parametrize("nh,nh_k", [(8, 8), (8, 4), (8, 2), (16, 2)])
def test_attn_fast_v_outliers_every_group_size(M, Mm, nh, nh_k):
    """V-side records on the fast kernels (round 2): per-warp shared-memory reduction in the QK phase.  M=32: every group size;
    M=64: G <= 2 directly, GQA-4/8 as 2-head sub-groups.  AUTO and FAST agree with the oracle; only v_out > 4 is refused."""
    from million_b200 import _lib as L
    for impl in (L.IMPL_AUTO, L.IMPL_FAST):
        got, ref = _attn_case(M, 2, nh, nh_k, 900, 5, 2, 2, impl, Mm=Mm)
        np.testing.assert_allclose(got, ref, atol=ATOL, rtol=RTOL)
    got, ref = _attn_case(M, 1, nh, nh_k, 700, 33, 0, 1, L.IMPL_FAST, Mm=Mm)      # a single V record: the aligned pair loads
    np.testing.assert_allclose(got, ref, atol=ATOL, rtol=RTOL)
    with pytest.raises(L.MillionError) as e:
        _attn_case(M, 1, nh, nh_k, 300, 5, 0, 5, L.IMPL_FAST, Mm=Mm)
    assert e.value.status == L.MILLION_ERR_UNSUPPORTED


def test_disabled_store_is_bit_identical(M):
    from million_b200 import _lib as L
    rng = np.random.default_rng(2)
    inp = O.make_inputs(bs=2, nh=8, nh_k=2, nk=1200, seed=9)
    t = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    a = M.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], 33)
    zi = torch.zeros(2, 2, 1200, 2, dtype=torch.uint8, device="cuda")
    zv = torch.zeros(2, 2, 1200, 2, dtype=t["q"].dtype, device="cuda")
    b = M.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], 33, k_outliers=(zi, zv))
    assert torch.equal(a, b)          # zero deltas add exactly nothing


def test_cache_with_outliers_prefill_decode(M):
    """DynamicPQCache(outliers=(2, 0)): prefill, decode across a window flush (async and sync), against the oracle attention on
    the cache's own state; and the state against the oracle encoder."""
    from million_b200.pq_utils import DynamicPQCache, Singleton
    outs = {}
    for async_flush in (True, False):
        Singleton.clear_instance()
        torch.manual_seed(0)
        bs, nh, nh_k, T = 1, 8, 2, 300
        cache = DynamicPQCache(bs=bs, nh=nh, num_key_value_heads=nh_k, M=64, layer_num=1, scalar_t=torch.float16, outliers=(2, 0),
                               async_flush=async_flush)
        cent = torch.randn(64, 256, 2, device="cuda").half()
        cache.set_cent(cent, cent)
        k = torch.randn(bs, nh_k, T, 128, device="cuda").half(); k[..., 7] *= 15
        v = torch.randn(bs, nh_k, T, 128, device="cuda").half()
        q = torch.randn(bs, nh, T, 128, device="cuda").half()
        cache.prefill(q, k, v, 0)
        codes, idx, val = O.pq_encode_outliers(k.float().cpu().numpy(), cent.float().cpu().numpy(), 2)
        assert np.array_equal(cache.key_cache[0].cpu().numpy(), codes)
        ki, kv = cache._ko[0].view(T)
        assert np.array_equal(ki.cpu().numpy(), idx) and np.array_equal(kv.float().cpu().numpy(), val.astype(np.float32))
        steps = []
        for step in range(140):                 # crosses one flush of the 128-token window
            qs = torch.randn(bs, nh, 1, 128, device="cuda").half()
            kn = torch.randn(bs, nh_k, 1, 128, device="cuda").half(); kn[..., 7] *= 15
            vn = torch.randn(bs, nh_k, 1, 128, device="cuda").half()
            out = cache.decoding(qs, kn, vn, 0)
            steps.append(out)
            if step in (0, 127, 128, 139):
                n, r = cache._k[0].len, cache.residualed_tokens[0]
                if cache._pending[0] is not None:
                    torch.cuda.synchronize()
                ki, kv = cache._ko[0].view(n)
                f = lambda t: t.float().cpu().numpy()
                ref = O.pq_decode_attn_outliers(f(qs), cache.key_cache[0][:, :, :n].cpu().numpy(), cache.value_cache[0][:, :, :n].cpu().numpy(),
                                                f(cent), f(cent), f(cache.key_residual_cache[0]), f(cache.value_residual_cache[0]), r,
                                                kout=(ki.cpu().numpy(), f(kv)))
                np.testing.assert_allclose(f(out), ref, atol=ATOL, rtol=RTOL)
        assert cache._k[0].len + (cache._pending[0][1] if cache._pending[0] else 0) == T + 128
        assert cache.outlier_store_size > 0
        outs[async_flush] = torch.stack(steps)
    assert torch.equal(outs[True], outs[False])
    Singleton.clear_instance()


def test_outlier_store_reduces_attention_error(M):
    """What the store is for: keys with a few heavy channels (heavy-tailed key channels).  Against EXACT fp32 attention
    on the unquantized K/V, PQ + 2 K-side records per token must be several times closer than plain PQ with the same codebook."""
    from million_b200 import _lib as L
    rng = np.random.default_rng(5)
    bs, nh, nh_k, nk = 1, 8, 2, 4096
    K = rng.standard_normal((bs, nh_k, nk, 128), dtype=np.float32)
    K[..., [11, 90]] *= 10.0                                             # two heavy channels (per-token spread: a common offset would cancel in the softmax)
    V = rng.standard_normal((bs, nh_k, nk, 128), dtype=np.float32)
    q = rng.standard_normal((bs, nh, 1, 128), dtype=np.float32)
    cent = torch.from_numpy(rng.standard_normal((64, 256, 2), dtype=np.float32)).half()
    Kh, Vh, qh = (torch.from_numpy(a).half().cuda() for a in (K, V, q))
    zres = torch.zeros(bs, nh_k, 128, 128, dtype=torch.float16, device="cuda")
    c32, c16 = cent.float().cuda(), cent.cuda()
    # exact attention on the fp16 inputs
    G = nh // nh_k
    Kf, Vf = Kh.float().repeat_interleave(G, 1), Vh.float().repeat_interleave(G, 1)
    exact = torch.softmax(qh.float() @ Kf.transpose(-1, -2) / 128 ** 0.5, -1) @ Vf
    vc = M.pq_encode(Vh, c32)
    plain = M.pq_decode_attn(qh, M.pq_encode(Kh, c32), vc, c16, c16, zres, zres, 0)
    kc, ki, kv = M.pq_encode_outliers(Kh, c32, 2)
    withs = M.pq_decode_attn(qh, kc, vc, c16, c16, zres, zres, 0, k_outliers=(ki, kv), impl=L.IMPL_FAST)
    # V is quantized identically in both runs: compare against exact attention over the SAME quantized V
    Vq = M.pq_decode(vc, c16).float().repeat_interleave(G, 1)
    exact_q = torch.softmax(qh.float() @ Kf.transpose(-1, -2) / 128 ** 0.5, -1) @ Vq
    e_plain = (plain.float() - exact_q).abs().max().item()
    e_with = (withs.float() - exact_q).abs().max().item()
    assert (ki.cpu().numpy()[..., :2] == 11).any() and e_with < 0.5 * e_plain, (e_plain, e_with)


def test_paged_cache_with_outliers(M):
    """PagedPQCache(outliers=(2, 1)): K records through the fast paged kernel when only K has them, K + V records through the generic
    one; prefill + decode across page flushes against the oracle attention on the cache's own state."""
    from million_b200.paged_pq_utils import PagedPQCache
    from million_b200.pq_utils import Singleton
    for outl in ((2, 0), (2, 1)):
        Singleton.clear_instance()
        torch.manual_seed(1)
        bs, nh, nh_k, T = 1, 8, 2, 256
        cache = PagedPQCache(bs=bs, nh=nh, num_key_value_heads=nh_k, M=64, layer_num=1, scalar_t=torch.float16, outliers=outl)
        cent = torch.randn(64, 256, 2, device="cuda").half()
        cache.set_cent(cent, cent)
        k = torch.randn(bs, nh_k, T, 128, device="cuda").half(); k[..., 7] *= 15
        v = torch.randn(bs, nh_k, T, 128, device="cuda").half(); v[..., 100] *= 10
        cache.prefill(torch.randn(bs, nh, T, 128, device="cuda").half(), k, v, 0)
        codes, idx, val = O.pq_encode_outliers(k.float().cpu().numpy(), cent.float().cpu().numpy(), 2)
        assert np.array_equal(cache.key_cache[0].cpu().numpy(), codes)
        f = lambda t: t.float().cpu().numpy()
        for step in range(200):                 # window 128, flush granularity 64: several page flushes
            qs = torch.randn(bs, nh, 1, 128, device="cuda").half()
            kn = torch.randn(bs, nh_k, 1, 128, device="cuda").half(); kn[..., 7] *= 15
            vn = torch.randn(bs, nh_k, 1, 128, device="cuda").half(); vn[..., 100] *= 10
            out = cache.decoding_with_pages(qs, kn, vn, 0)
            if step in (0, 63, 64, 130, 199):
                torch.cuda.synchronize()
                n, r = cache._k[0].len, cache.residualed_tokens[0]
                vcodes = cache.value_cache[0][..., :n].transpose(2, 3).contiguous().cpu().numpy()
                ki, kv = cache._ko[0].view(n)
                vout = None
                if outl[1]:
                    vi, vv = cache._vo[0].view(n)
                    vout = (vi.cpu().numpy(), f(vv))
                ref = O.pq_decode_attn_outliers(f(qs), cache.key_cache[0][:, :, :n].cpu().numpy(), vcodes, f(cent), f(cent),
                                                f(cache.key_residual_cache[0]), f(cache.value_residual_cache[0]), r,
                                                kout=(ki.cpu().numpy(), f(kv)), vout=vout)
                np.testing.assert_allclose(f(out), ref, atol=ATOL, rtol=RTOL)
    Singleton.clear_instance()


@pytest.mark.parametrize("name,k_out,Mm", [("m64k2", 2, 64), ("m32k1", 1, 32), ("m64k4", 4, 64)])
def test_committed_outlier_vectors(M, name, k_out, Mm):
    """the CUDA path against tests/golden/outlier_golden.npz: records and codes bit-exact, attention within the north-star tolerance
    (generic kernel: K and V records; fast kernel: K records only)"""
    import os
    from million_b200 import _lib as L
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "outlier_golden.npz"))
    X, C = torch.from_numpy(g[f"{name}_X"]).cuda(), torch.from_numpy(g[f"{name}_C"]).cuda()
    codes, idx, val = M.pq_encode_outliers(X, C, k_out)
    assert np.array_equal(codes.cpu().numpy(), g[f"{name}_codes"]) and np.array_equal(idx.cpu().numpy(), g[f"{name}_idx"])
    assert np.array_equal(val.cpu().numpy(), g[f"{name}_val"])
    dv = lambda k: torch.from_numpy(g[f"{name}_{k}"]).cuda()
    out = M.pq_decode_attn(dv("q"), codes, dv("vcodes"), C.half(), C.half(), dv("kres"), dv("vres"), 9, impl=L.IMPL_GENERIC,
                           k_outliers=(idx, val), v_outliers=(dv("vidx"), dv("vval")))
    np.testing.assert_allclose(out.float().cpu().numpy(), g[f"{name}_attn"], atol=ATOL, rtol=RTOL)
    ref_k = O.pq_decode_attn_outliers(g[f"{name}_q"], g[f"{name}_codes"], g[f"{name}_vcodes"], g[f"{name}_C"], g[f"{name}_C"],
                                      g[f"{name}_kres"], g[f"{name}_vres"], 9, kout=(g[f"{name}_idx"], g[f"{name}_val"]))
    fast = M.pq_decode_attn(dv("q"), codes, dv("vcodes"), C.half(), C.half(), dv("kres"), dv("vres"), 9, impl=L.IMPL_FAST, k_outliers=(idx, val))
    np.testing.assert_allclose(fast.float().cpu().numpy(), ref_k, atol=ATOL, rtol=RTOL)
