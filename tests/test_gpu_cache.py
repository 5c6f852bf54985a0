"""GPU: the Python cache API (million_b200.pq_utils.DynamicPQCache / paged_pq_utils.PagedPQCache) against the oracle's
restatement of the reference policy, step by step."""
import numpy as np
import pytest
import torch

from oracle import pq_oracle as O

pytestmark = pytest.mark.gpu
ATOL, RTOL = 2e-3, 1e-2


def _mk(cls, **kw):
    from million_b200.pq_utils import Singleton
    Singleton.clear_instance()
    return cls(**kw)


def _run(cache, oracle, steps, T0, nh, nh_k, d, decode_name, seed=0):
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.standard_normal(s, dtype=np.float32).astype(np.float16)
    q, k, v = f(1, nh, T0, d), f(1, nh_k, T0, d), f(1, nh_k, T0, d)
    out = cache.prefill(torch.from_numpy(q).cuda(), torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda(), 0)
    ref = oracle.prefill(q, k, v, 0)
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=ATOL, rtol=RTOL)
    for step in range(steps):
        q, k, v = f(1, nh, 1, d), f(1, nh_k, 1, d), f(1, nh_k, 1, d)
        out = getattr(cache, decode_name)(torch.from_numpy(q).cuda(), torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda(), 0)
        ref = getattr(oracle, decode_name)(q, k, v, 0)
        np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=ATOL, rtol=RTOL, err_msg=f"step {step}")
        assert cache.residualed_tokens[0] == oracle.residualed_tokens[0]
        assert cache.seen_tokens[0] == oracle.seen_tokens[0]
    torch.cuda.synchronize()


@pytest.mark.parametrize("async_flush", [False, True])
def test_dynamic_cache_follows_reference_policy(async_flush):
    from million_b200.pq_utils import DynamicPQCache
    kw = dict(bs=1, nh=8, num_key_value_heads=2, M=64, layer_num=1, d=128)
    rng = np.random.default_rng(1)
    kc = rng.standard_normal((64, 256, 2), dtype=np.float32).astype(np.float16)
    vc = rng.standard_normal((64, 256, 2), dtype=np.float32).astype(np.float16)
    cache = _mk(DynamicPQCache, scalar_t=torch.float16, async_flush=async_flush, **kw)
    cache.set_cent(torch.from_numpy(kc).cuda(), torch.from_numpy(vc).cuda())
    oracle = O.DynamicPQCacheOracle(**kw)
    oracle.set_cent(kc, vc)
    _run(cache, oracle, 270, 37, 8, 2, 128, "decoding")
    # codes are bit-exact, window flushed twice (37 + 2*128 coded)
    assert np.array_equal(cache.key_cache[0].cpu().numpy(), oracle.key_cache[0])
    assert np.array_equal(cache.value_cache[0].cpu().numpy(), oracle.value_cache[0])
    assert cache.key_cache[0].shape == (1, 2, 37 + 256, 64)
    assert cache.pq_cache_size == 2 * 2 * (37 + 256) * 64


@pytest.mark.parametrize("async_flush", [False, True])
def test_paged_cache_follows_reference_policy(async_flush):
    from million_b200.paged_pq_utils import PagedPQCache
    kw = dict(bs=1, nh=8, num_key_value_heads=2, M=64, layer_num=1, d=128)
    rng = np.random.default_rng(2)
    kc = rng.standard_normal((64, 256, 2), dtype=np.float32).astype(np.float16)
    vc = rng.standard_normal((64, 256, 2), dtype=np.float32).astype(np.float16)
    cache = _mk(PagedPQCache, scalar_t=torch.float16, async_flush=async_flush, **kw)
    cache.set_cent(torch.from_numpy(kc).cuda(), torch.from_numpy(vc).cuda())
    oracle = O.PagedPQCacheOracle(**kw)
    oracle.set_cent(kc, vc)
    _run(cache, oracle, 300, 128, 8, 2, 128, "decoding_with_pages")
    assert np.array_equal(cache.key_cache[0].cpu().numpy(), oracle.key_cache[0])
    assert np.array_equal(cache.value_cache[0].cpu().numpy(), oracle.value_cache[0])     # (bs, nh_k, M, T) face
    # block table: chunk-major, then b, then h, ids ascending from a fresh pool
    assert cache.value_page_ids[0][0][0][:3] == [0, 2, 4] and cache.value_page_ids[0][0][1][:3] == [1, 3, 5]
    pool_ref, table_ref = O.build_page_pool(np.ascontiguousarray(oracle.value_cache[0].transpose(0, 1, 3, 2)), 64)
    n = table_ref.shape[2]
    assert np.array_equal(cache._table[0][:, :, :n].cpu().numpy(), table_ref)
    assert np.array_equal(cache.page_managers[0].page_pool[:pool_ref.shape[0]].cpu().numpy(), pool_ref)
    st = cache.get_cache_stats()
    assert st['total_pages_allocated'] == pool_ref.shape[0]


def test_update_path_matches_oracle():
    from million_b200.pq_utils import DynamicPQCache
    kw = dict(bs=2, nh=4, num_key_value_heads=4, M=32, layer_num=2, d=128)
    rng = np.random.default_rng(3)
    cent = rng.standard_normal((32, 256, 4), dtype=np.float32).astype(np.float16)
    cache = _mk(DynamicPQCache, scalar_t=torch.float16, **kw)
    cache.set_cent(torch.from_numpy(cent).cuda(), torch.from_numpy(cent).cuda())
    oracle = O.DynamicPQCacheOracle(**kw)
    oracle.set_cent(cent, cent)
    f = lambda *s: rng.standard_normal(s, dtype=np.float32).astype(np.float16)
    for n, distort in ((5, False), (1, False), (3, True)):
        k, v = f(2, 4, n, 128), f(2, 4, n, 128)
        K, V = cache.update(torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda(), 1, distort_recent=distort)
        Kr, Vr = oracle.update(k, v, 1, distort_recent=distort)
        assert np.array_equal(K.float().cpu().numpy(), np.asarray(Kr, np.float32))
        assert np.array_equal(V.float().cpu().numpy(), np.asarray(Vr, np.float32))
    assert cache.seen_tokens == [0, 9]


def test_synthetic_token_perplexity_parity():
    """Perplexity path (pq_utils.py:243-260 with distort_recent=True): a tiny random-init Llama-like stack run with
    the PQ cache vs the oracle's PQ path; NLL must agree within 0.5 % (north_star)."""
    from million_b200.pq_utils import DynamicPQCache
    torch.manual_seed(0)
    L_, nh, nh_k, d, T, V = 2, 8, 8, 128, 96, 64
    hid = nh * d
    kw = dict(bs=1, nh=nh, num_key_value_heads=nh_k, M=64, layer_num=L_, d=d)
    cent = torch.randn(64, 256, 2).half()
    cache = _mk(DynamicPQCache, scalar_t=torch.float16, **kw)
    cache.set_cent(cent.cuda(), cent.cuda())
    oracle = O.DynamicPQCacheOracle(**kw)
    oracle.set_cent(cent.numpy(), cent.numpy())
    emb = torch.randn(V, hid) * 0.5
    W = [dict(q=torch.randn(hid, hid) / hid ** 0.5, k=torch.randn(hid, nh_k * d) / hid ** 0.5,
              v=torch.randn(hid, nh_k * d) / hid ** 0.5, o=torch.randn(hid, hid) / hid ** 0.5) for _ in range(L_)]
    head = torch.randn(hid, V) / hid ** 0.5
    ids = torch.randint(0, V, (1, T))

    def run(attn):
        x = emb[ids]
        for l in range(L_):
            h = torch.nn.functional.layer_norm(x, (hid,))
            q = (h @ W[l]['q']).view(1, T, nh, d).transpose(1, 2).half()
            k = (h @ W[l]['k']).view(1, T, nh_k, d).transpose(1, 2).half()
            v = (h @ W[l]['v']).view(1, T, nh_k, d).transpose(1, 2).half()
            a = attn(q, k, v, l)
            x = x + a.transpose(1, 2).reshape(1, T, hid) @ W[l]['o']
        logits = torch.nn.functional.layer_norm(x, (hid,)) @ head
        return torch.nn.functional.cross_entropy(logits[0, :-1], ids[0, 1:]).item()

    nll_gpu = run(lambda q, k, v, l: cache.prefill(q.cuda(), k.cuda(), v.cuda(), l, distort_recent=True).float().cpu())
    nll_ref = run(lambda q, k, v, l: torch.from_numpy(oracle.prefill(q.numpy(), k.numpy(), v.numpy(), l, distort_recent=True)))
    assert abs(np.exp(nll_gpu) / np.exp(nll_ref) - 1) < 5e-3


@pytest.mark.parametrize("dtype,bs,nh,nh_k,M", [(torch.bfloat16, 2, 8, 2, 64), (torch.float16, 2, 4, 4, 32), (torch.float16, 1, 16, 2, 64)])
def test_dynamic_cache_batched_gqa_bf16(dtype, bs, nh, nh_k, M):
    """bs > 1 (the reference hard-wires 1, pq_utils.py:56), GQA 4 and 8, bf16, the 2-bit shape: prefill + decode across a flush."""
    from million_b200.pq_utils import DynamicPQCache
    kw = dict(bs=bs, nh=nh, num_key_value_heads=nh_k, M=M, layer_num=2, d=128)
    rng = np.random.default_rng(5)
    mk = lambda *s: torch.from_numpy(rng.standard_normal(s, dtype=np.float32)).to(dtype)
    kc, vc = mk(M, 256, 128 // M), mk(M, 256, 128 // M)
    cache = _mk(DynamicPQCache, scalar_t=dtype, **kw)
    cache.set_cent(kc.cuda(), vc.cuda())
    oracle = O.DynamicPQCacheOracle(**kw)
    oracle.set_cent(kc.float().numpy(), vc.float().numpy())
    atol = ATOL
    for layer in (0, 1):
        q, k, v = mk(bs, nh, 70, 128), mk(bs, nh_k, 70, 128), mk(bs, nh_k, 70, 128)
        out = cache.prefill(q.cuda(), k.cuda(), v.cuda(), layer)
        ref = oracle.prefill(q.float().numpy(), k.float().numpy(), v.float().numpy(), layer)
        # prefill attention is torch's SDPA in the model dtype — library code, the same call the reference makes
        # (pq_utils.py:253-260): in bf16 it rounds P to 8 bits before PV, so ITS tolerance applies here, not the PQ kernel's
        np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=atol if dtype == torch.float16 else 1.6e-2, rtol=RTOL)
    for step in range(135):
        for layer in (0, 1):
            q, k, v = mk(bs, nh, 1, 128), mk(bs, nh_k, 1, 128), mk(bs, nh_k, 1, 128)
            out = cache.decoding(q.cuda(), k.cuda(), v.cuda(), layer)
            ref = oracle.decoding(q.float().numpy(), k.float().numpy(), v.float().numpy(), layer)
            np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=atol, rtol=RTOL, err_msg=f"step {step} layer {layer}")
    assert cache.key_cache[1].shape == (bs, nh_k, 70 + 128, M) and cache.residualed_tokens == [7, 7]
    assert np.array_equal(cache.key_cache[0].cpu().numpy(), oracle.key_cache[0])
    assert np.array_equal(cache.value_cache[1].cpu().numpy(), oracle.value_cache[1])


def test_errors_are_exceptions_not_exit():
    """The reference prints and exit()s on a launch error (Interface.cu:3-11); here bad arguments raise MillionError."""
    from million_b200 import _lib as L, ops
    q = torch.randn(1, 8, 1, 128, device="cuda").half()
    kc = torch.zeros(1, 2, 10, 64, dtype=torch.uint8, device="cuda")
    cent = torch.randn(64, 256, 2, device="cuda").half()
    res = torch.randn(1, 2, 128, 128, device="cuda").half()
    with pytest.raises(L.MillionError, match="r"):
        ops.pq_decode_attn(q, kc, kc, cent, cent, res, res, 200)            # r > window length
    with pytest.raises(L.MillionError):
        ops.pq_decode_attn(q, kc, kc, cent, cent, res, res, 5, impl=L.IMPL_FAST, prepared=None, v_layout=L.V_PAGED,
                           v_page_ids=torch.zeros(1, 2, 0, dtype=torch.int64, device="cuda"), page_size=64)   # 0 pages for 10 tokens
    out = ops.pq_decode_attn(q, kc, kc, cent, cent, res, res, 5)           # the library is still usable afterwards
    assert torch.isfinite(out).all()
