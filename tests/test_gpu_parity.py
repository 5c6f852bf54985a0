"""GPU parity tests proper: the CUDA path, called through the C ABI (million_b200.ops / .bindings), against the
CPU oracle (oracle/) on the same seeded inputs and against the committed golden fixtures.

Bars (BASELINE.json north_star): codes bit-exact; attention max-abs 2e-3 / rel 1e-2.
"""
import numpy as np
import pytest
import torch

from oracle import pq_oracle as O

pytestmark = pytest.mark.gpu

ATOL, RTOL = 2e-3, 1e-2   # north_star tolerance for fp16/bf16 I/O with fp32 accumulation


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


@pytest.fixture(scope="module")
def M():
    import million_b200.ops as ops
    from million_b200 import _lib
    _lib.lib()
    return ops


def impls():
    from million_b200 import _lib as L
    return [L.IMPL_GENERIC, L.IMPL_FAST]


# ------------------------------------------------------------------------------------------------ encode


@pytest.mark.parametrize("i", range(5))
@pytest.mark.parametrize("impl", [1, 2])
def test_encode_golden_bit_exact(M, golden, i, impl):
    from million_b200 import _lib as L
    X, cent = golden[f"enc{i}_X"], golden[f"enc{i}_cent"]
    try:
        got = M.pq_encode(dev(X), dev(cent, torch.float32), impl=impl)
    except L.MillionError as e:
        if impl == L.IMPL_FAST and e.status == L.MILLION_ERR_UNSUPPORTED:
            pytest.skip("shape not covered by the fast encoder")
        raise
    assert np.array_equal(got.cpu().numpy(), golden[f"enc{i}_codes_keops"])


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("d,Mm,C", [(128, 64, 256), (128, 32, 256), (128, 16, 256), (64, 32, 128), (64, 16, 256)])
def test_encode_random_bit_exact(M, dtype, d, Mm, C):
    rng = np.random.default_rng(7)
    X = torch.from_numpy(rng.standard_normal((2, 3, 301, d), dtype=np.float32)).to(dtype)
    cent = torch.from_numpy(rng.standard_normal((Mm, C, d // Mm), dtype=np.float32)).to(dtype)
    ref = O.pq_encode(X.float().numpy(), cent.float().numpy())
    got = M.pq_encode(X.cuda(), cent.float().cuda())
    assert np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_encode_grid_bit_exact(M, dtype, golden):
    """candidate-grid encoder (d/M = 2): random, heavy-tailed, degenerate codebooks, non-finite inputs; golden vectors"""
    from million_b200 import _lib as L
    rng = np.random.default_rng(21)
    X = rng.standard_normal((2, 3, 1501, 128), dtype=np.float32)
    X *= np.where(rng.random(X.shape) < 0.02, 25.0, 1.0).astype(np.float32)
    X[0, 0, :50] = np.nan; X[0, 0, 50:60] = np.inf; X[0, 0, 60:70] = 0.0
    X = torch.from_numpy(X).to(dtype)
    cent = torch.from_numpy(rng.standard_normal((64, 256, 2), dtype=np.float32)).half().float()
    cent[1, 100] = cent[1, 7]                                  # duplicate centroid: first index must win
    cent[2] = cent[2] * 0.01 + 1.0                             # tight cluster (cells overflow -> full scan)
    cent[3, :, 1] = 0.0                                        # collinear
    cent[4] = cent[4, 0]                                       # all identical
    X[1, 1, :256] = cent.permute(1, 0, 2).reshape(256, 128).to(dtype)     # points exactly on centroids
    with np.errstate(invalid="ignore"):
        ref = O.pq_encode(X.float().numpy(), cent.numpy())
    got = M.pq_encode(X.cuda(), cent.cuda(), impl=L.IMPL_GRID)
    assert np.array_equal(got.cpu().numpy(), ref)
    assert np.array_equal(M.pq_encode(X.cuda(), cent.cuda()).cpu().numpy(), ref)     # AUTO takes the same path
    for i in range(5):
        Xg, cg = golden[f"enc{i}_X"], golden[f"enc{i}_cent"]
        if cg.shape[2] == 2 and cg.shape[1] <= 256 and dtype == torch.float16:
            g = M.pq_encode(dev(Xg), dev(cg, torch.float32), impl=L.IMPL_GRID)
            assert np.array_equal(g.cpu().numpy(), golden[f"enc{i}_codes_keops"])


def test_encode_ties_and_wide_codes(M, golden):
    got = M.pq_encode(dev(golden["tie_X"]), dev(golden["tie_cent"], torch.float32))
    assert np.array_equal(got.cpu().numpy(), golden["tie_codes"])
    got = M.pq_encode(dev(golden["wide_X"]), dev(golden["wide_cent"], torch.float32), out_dtype=torch.int32)
    assert np.array_equal(got.cpu().numpy(), golden["wide_codes"])


def test_encode_layouts_strided_transposed_paged(M):
    rng = np.random.default_rng(3)
    bs, nh, Lt, d, Mm = 2, 3, 128, 128, 64
    win = torch.from_numpy(rng.standard_normal((bs, nh, Lt, d), dtype=np.float32)).half().cuda()
    cent = torch.from_numpy(rng.standard_normal((Mm, 256, 2), dtype=np.float32)).half()
    ref = O.pq_encode(win[:, :, :64].float().cpu().numpy(), cent.float().numpy())
    c32 = cent.float().cuda()
    # strided source slice, append at t0 into a larger cache
    cache = torch.zeros(bs, nh, 200, Mm, dtype=torch.uint8, device="cuda")
    M.pq_encode_into(win[:, :, :64], c32, cache, t0=100)
    assert np.array_equal(cache[:, :, 100:164].cpu().numpy(), ref) and int(cache[:, :, :100].sum()) == 0
    # transposed store (live PagedPQCache layout)
    cache_t = torch.zeros(bs, nh, Mm, 256, dtype=torch.uint8, device="cuda")
    M.pq_encode_into(win[:, :, :64], c32, cache_t, t0=64, layout="transposed")
    assert np.array_equal(cache_t[:, :, :, 64:128].transpose(2, 3).cpu().numpy(), ref)
    # paged store: pages handed out chunk-major, b, h
    pool_ref, table_ref = O.build_page_pool(np.concatenate([ref, ref[:, :, :10]], axis=2), 64)
    pool = torch.zeros(pool_ref.shape, dtype=torch.uint8, device="cuda")
    table = dev(table_ref)
    M.pq_encode_paged(win[:, :, :64], c32, pool, table, t0=0)
    M.pq_encode_paged(win[:, :, :10], c32, pool, table, t0=64)
    assert np.array_equal(pool.cpu().numpy(), pool_ref)


def test_encode_empty(M):
    X = torch.zeros(1, 2, 0, 128, dtype=torch.float16, device="cuda")
    cent = torch.randn(64, 256, 2, device="cuda")
    assert M.pq_encode(X, cent).shape == (1, 2, 0, 64)


# ------------------------------------------------------------------------------------------------ reconstruct


@pytest.mark.parametrize("i", range(5))
def test_decode_golden_exact(M, golden, i):
    got = M.pq_decode(dev(golden[f"enc{i}_codes_keops"]), dev(golden[f"enc{i}_cent"]))
    assert got.dtype == torch.float16
    assert np.array_equal(got.cpu().numpy(), golden[f"enc{i}_decoded"])


def test_encode_decode_idempotent(M):
    """decode(encode(x)) is a fixed point of encode (distinct centroids)."""
    torch.manual_seed(0)
    cent = torch.randn(64, 256, 2, device="cuda").half()
    X = torch.randn(1, 4, 1000, 128, device="cuda").half()
    codes = M.pq_encode(X, cent.float())
    again = M.pq_encode(M.pq_decode(codes, cent), cent.float())
    assert torch.equal(codes, again)


# ------------------------------------------------------------------------------------------------ attention


def _attn_case(M, g, impl, dtype=torch.float16, v_layout=0, n_splits=0):
    from million_b200 import _lib as L
    bs, nh, nh_k, nk, r, d, Mm, C = [int(v) for v in g("shape")]
    cast = lambda a: dev(a).to(dtype)
    q, kcent, vcent, kres, vres = cast(g("q")), cast(g("kcent")), cast(g("vcent")), cast(g("kres")), cast(g("vres"))
    kc, vc = dev(g("kc")), dev(g("vc"))
    kw = {}
    if v_layout == L.V_TRANSPOSED:
        vt = torch.zeros(bs, nh_k, Mm, nk + 37, dtype=torch.uint8, device="cuda")
        vt[..., :nk] = vc.transpose(2, 3)
        vc_arg = vt
    elif v_layout == L.V_PAGED:
        pool, table = O.build_page_pool(g("vc"), 64)
        perm = np.random.default_rng(5).permutation(pool.shape[0] + 3)      # scatter the pages
        pool2 = np.zeros((pool.shape[0] + 3,) + pool.shape[1:], np.uint8)
        pool2[perm[:pool.shape[0]]] = pool
        table = perm[table] if table.size else table
        vc_arg, kw = dev(pool2), dict(v_page_ids=dev(table.astype(np.int64)), page_size=64)
    else:
        vc_arg = vc
    try:
        out = M.pq_decode_attn(q, kc, vc_arg, kcent, vcent, kres, vres, r, nk=nk, v_layout=v_layout, impl=impl,
                               n_splits=n_splits, **kw)
    except L.MillionError as e:
        if impl == L.IMPL_FAST and e.status == L.MILLION_ERR_UNSUPPORTED:
            pytest.skip("shape not covered by the fast kernel")
        raise
    ref = O.pq_decode_attn(q.float().cpu().numpy(), g("kc"), g("vc"), kcent.float().cpu().numpy(), vcent.float().cpu().numpy(),
                           kres.float().cpu().numpy(), vres.float().cpu().numpy(), r)
    return out.float().cpu().numpy(), ref


@pytest.mark.parametrize("i", range(4))
@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("v_layout", [0, 1, 2])
def test_attn_golden(M, golden, i, impl, v_layout):
    g = lambda k: golden[f"att{i}_{k}"]
    out, ref = _attn_case(M, g, impl, v_layout=v_layout)
    np.testing.assert_allclose(out, ref, atol=ATOL, rtol=RTOL)
    # and against the fixture computed by the reference's own functions (fp16 inputs)
    np.testing.assert_allclose(out, g("out"), atol=ATOL, rtol=RTOL)


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_attn_dtypes_and_splits(M, golden, impl, dtype):
    g = lambda k: golden[f"att0_{k}"]
    for n_splits in (0, 1, 3, 19):
        out, ref = _attn_case(M, g, impl, dtype=dtype, n_splits=n_splits)
        np.testing.assert_allclose(out, ref, atol=ATOL, rtol=RTOL)


def _rand_case(bs, nh, nh_k, nk, r, d=128, Mm=64, C=256, seed=0, dtype=torch.float16):
    inp = O.make_inputs(bs=bs, nh=nh, nh_k=nh_k, nk=nk, d=d, M=Mm, C=C, Lt=d, seed=seed)
    t = {k: dev(v) for k, v in inp.items()}
    for k in ("q", "kcent", "vcent", "kres", "vres"):
        t[k] = t[k].to(dtype)
    return inp, t


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("shape", [
    (1, 8, 8, 4096 - 17, 17),     # BASELINE config 1 (tiny Llama, MHA 8 heads, 4K ctx)
    (2, 32, 8, 777, 128),         # GQA 4, ragged length, full window
    (1, 32, 32, 130, 1),          # MHA, r = 1
    (1, 16, 2, 1, 3),             # GQA 8, a single coded token
    (3, 4, 4, 0, 9),              # no coded tokens at all
    (1, 8, 4, 2000, 0),           # empty window (the reference forbids r = 0; we define it)
])
def test_attn_random_shapes(M, impl, shape):
    from million_b200 import _lib as L
    bs, nh, nh_k, nk, r = shape
    inp, t = _rand_case(bs, nh, nh_k, nk, r, seed=sum(shape))
    try:
        out = M.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], r, impl=impl)
    except L.MillionError as e:
        if impl == L.IMPL_FAST and e.status == L.MILLION_ERR_UNSUPPORTED:
            pytest.skip("shape not covered by the fast kernel")
        raise
    if nk + r == 0:
        assert torch.count_nonzero(out) == 0
        return
    ref = O.pq_decode_attn(inp["q"], inp["kc"], inp["vc"], inp["kcent"], inp["vcent"], inp["kres"], inp["vres"], r)
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=ATOL, rtol=RTOL)


@pytest.mark.parametrize("d,Mm,C", [(128, 32, 256), (128, 16, 256), (64, 32, 256), (64, 16, 128), (128, 64, 128)])
def test_attn_template_grid_of_the_reference(M, d, Mm, C):
    """setup.py:10-15 compiles d in {64,128} x M in {16,32,64} x C in {128,256}; all must work (AUTO)."""
    inp, t = _rand_case(1, 8, 4, 333, 21, d=d, Mm=Mm, C=C, seed=d + Mm + C)
    out = M.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], 21)
    ref = O.pq_decode_attn(inp["q"], inp["kc"], inp["vc"], inp["kcent"], inp["vcent"], inp["kres"], inp["vres"], 21)
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=ATOL, rtol=RTOL)


def test_attn_peaked_distribution(M):
    """A few keys dominate (real models have sinks): output ~ those tokens' values; checks max tracking / rescale and states the
    precision the kernels reach as the logits grow.  Measured derivation: tools/precision_budget.py -> profiles/r02_precision_budget.txt.
      * logits with sigma ~ 4 (q x4): every kernel meets north_star's 2e-3 / 1e-2;
      * sigma ~ 12 (q x12): the fp32 all-shapes kernel and the one-head-per-KV-head fast kernel (fp32 LUT entries) still do; the
        GQA fast kernel keeps its K LUT as fp16 (4 heads in 8 bytes — the shared memory it has), i.e. 2^-11 relative on entries
        of magnitude |q_m||c|, ~3e-3 on a logit: softmax weights move by ~0.3 %, which shows where two tokens share the mass.
        Bound asserted for that case: atol 1.2e-2, and at most 2 % of the elements beyond north_star's tolerance.  (The
        reference's kernel accumulates everything in fp16: tests/test_gpu_vs_reference_kernel.py.)  The packed-half PV sums are
        NOT the limiter: flushing them twice as often changes nothing in that table."""
    for qscale in (4.0, 12.0):
        inp, t = _rand_case(1, 8, 2, 3000, 5, seed=11)
        t["q"] = (t["q"].float() * qscale).half()
        ref = O.pq_decode_attn(t["q"].float().cpu().numpy(), inp["kc"], inp["vc"], inp["kcent"], inp["vcent"], inp["kres"], inp["vres"], 5)
        for impl in (1, 0):
            out = M.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], 5, impl=impl).float().cpu().numpy()
            if impl == 1 or qscale <= 4.0:
                np.testing.assert_allclose(out, ref, atol=ATOL, rtol=RTOL)
            else:
                np.testing.assert_allclose(out, ref, atol=1.2e-2, rtol=RTOL)
                beyond = (np.abs(out - ref) > ATOL + RTOL * np.abs(ref)).mean()
                assert beyond <= 0.02, f"{beyond:.1%} of the elements beyond 2e-3 / 1e-2"
    # one head per KV head: fp32 LUT entries, north_star's tolerance at sigma ~ 12 too
    inp, t = _rand_case(1, 8, 8, 3000, 5, seed=12)
    t["q"] = (t["q"].float() * 12.0).half()
    ref = O.pq_decode_attn(t["q"].float().cpu().numpy(), inp["kc"], inp["vc"], inp["kcent"], inp["vcent"], inp["kres"], inp["vres"], 5)
    out = M.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], 5, impl=2)
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=ATOL, rtol=RTOL)


def test_attn_deterministic_and_counter_reset(M):
    inp, t = _rand_case(2, 32, 8, 5000, 77, seed=5)
    outs = [M.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], 77).clone() for _ in range(5)]
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


def test_named_bindings_signature(M, golden):
    """The reference's 10-argument call (pq_utils.py:83-94) through the manufactured name."""
    from million_b200 import bindings
    g = lambda k: golden[f"att3_{k}"]
    bs, nh, nh_k, nk, r, d, Mm, C = [int(v) for v in g("shape")]
    fn = getattr(bindings, f"flash_decoding_allocated_buffer_f16u8_Ns16Lt{d}d{d}M{Mm}C{C}")
    po = torch.empty(bs, nh, 17, d, dtype=torch.float16, device="cuda")
    pl = torch.empty(bs, nh, 17, dtype=torch.float16, device="cuda")
    out = fn(dev(g("q")), dev(g("kc")), dev(g("vc")), dev(g("kcent")), dev(g("vcent")), dev(g("kres")), dev(g("vres")), r, po, pl)
    assert out.shape == (bs, nh, 1, d) and out.dtype == torch.float16
    np.testing.assert_allclose(out.float().cpu().numpy(), g("out"), atol=ATOL, rtol=RTOL)
    with pytest.raises(TypeError):
        fn(dev(g("q")).bfloat16(), dev(g("kc")), dev(g("vc")), dev(g("kcent")), dev(g("vcent")), dev(g("kres")), dev(g("vres")), r, po, pl)


def test_partial_and_lse_merge_equals_single_call(M):
    """Split-KV across 'ranks': every rank attends to its token slice (window on the last), partial states are
    merged with million_lse_merge; must equal the single-call result."""
    inp, t = _rand_case(1, 32, 8, 4000, 50, seed=9)
    full = M.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], 50)
    G = 4
    parts = torch.empty(G, 32, 130, dtype=torch.float32, device="cuda")
    per = 4000 // G
    for g in range(G):
        kc = t["kc"][:, :, g * per:(g + 1) * per].contiguous()
        vc = t["vc"][:, :, g * per:(g + 1) * per].contiguous()
        M.pq_decode_attn(t["q"], kc, vc, t["kcent"], t["vcent"], t["kres"], t["vres"], 50 if g == G - 1 else 0, partial=parts[g])
    merged = M.lse_merge(parts, 128, torch.float16).view(1, 32, 1, 128)
    np.testing.assert_allclose(merged.float().cpu().numpy(), full.float().cpu().numpy(), atol=1e-3, rtol=5e-3)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("bs,nh,nh_k,nk", [(1, 32, 8, 33000), (1, 8, 2, 3000), (8, 8, 2, 4100), (3, 16, 4, 0)])
def test_fused_splitkv_world1_equals_plain_call(M, dtype, bs, nh, nh_k, nk):
    """MILLION_ATTN_FUSED_SPLITKV with a world of one rank (the protocol degenerates to push-to-self, per-group flag, merge of one
    part): same result as the plain call, on consecutive launches (sequence parity flips), no timeout, with and without PDL.
    The multi-rank run is bench.py --gpus N (extra.splitkv_128k) and tools/splitkv_nccl.py --p2p --fused."""
    from million_b200 import _lib as L
    inp = O.make_inputs(bs=bs, nh=nh, nh_k=nh_k, nk=nk, seed=5)
    t = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    for k in ("q", "kcent", "vcent", "kres", "vres"):
        t[k] = t[k].to(dtype)
    rows, d = bs * nh, 128
    buf = torch.zeros(L.lib().million_splitkv_symmetric_bytes(1, rows, d), dtype=torch.uint8, device="cuda")
    state = M.splitkv_state([buf.data_ptr()], 0, rows, buf.device)
    n = 0
    for r in (128, 17, 1):
        for pdl in (False, True):
            plain = M.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], r)
            fused = M.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], r, p2p=state, pdl=pdl)
            n += 1
            np.testing.assert_allclose(fused.float().cpu().numpy(), plain.float().cpu().numpy(), atol=2e-3, rtol=1e-2)
    st = state.view(torch.int32).cpu().numpy()
    assert st[0] == n and st[1] == 0 and st[2] == 0          # n completed calls, ticket back to zero, no timeout


def test_fused_splitkv_unsupported_shapes_say_so(M):
    from million_b200 import _lib as L
    inp = O.make_inputs(bs=1, nh=8, nh_k=8, nk=500, seed=5)           # MHA: the fused exchange is compiled for nh/nh_k = 4
    t = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    buf = torch.zeros(L.lib().million_splitkv_symmetric_bytes(1, 8, 128), dtype=torch.uint8, device="cuda")
    state = M.splitkv_state([buf.data_ptr()], 0, 8, buf.device)
    with pytest.raises(L.MillionError) as e:
        M.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], 5, p2p=state)
    assert e.value.status == L.MILLION_ERR_UNSUPPORTED


def test_pdl_launches_match_plain_launches(M):
    """MILLION_ATTN_PDL: back-to-back launches that overlap their prologue with the predecessor's tail give bit-identical results
    (fast M=64, fast M=32 and generic kernels), also inside a CUDA graph."""
    for Mm, impl in ((64, 0), (32, 0), (64, 1)):
        inp = O.make_inputs(bs=2, nh=32, nh_k=8, nk=3000, M=Mm, seed=8)
        t = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
        outs = [torch.empty(2, 32, 1, 128, dtype=torch.float16, device="cuda") for _ in range(6)]
        want = M.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], 77, impl=impl).clone()

        def run():
            for o in outs:
                M.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], 77, impl=impl, out=o, pdl=True)
        run()
        torch.cuda.synchronize()
        for o in outs:
            assert torch.equal(o, want)
            o.zero_()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            run()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        for o in outs:
            o.zero_()
        g.replay()
        torch.cuda.synchronize()
        for o in outs:
            assert torch.equal(o, want)


def test_full_size_llama31_8b_32k_properties(M):
    """BASELINE config 2 at full size (bs 1, 32q/8kv heads, 32K ctx): checked against a torch fp32 restatement on the
    GPU (de-quantise + softmax), and the size-independent property 'all value codes equal -> output is that centroid
    row mixed only with the window'."""
    torch.manual_seed(42)
    bs, nh, nh_k, nk, r, d, Mm, C = 1, 32, 8, 32768 - 128, 128, 128, 64, 256
    kcent, vcent = torch.randn(Mm, C, 2, device="cuda").half(), torch.randn(Mm, C, 2, device="cuda").half()
    kc = torch.randint(0, C, (bs, nh_k, nk, Mm), dtype=torch.uint8, device="cuda")
    vc = torch.randint(0, C, (bs, nh_k, nk, Mm), dtype=torch.uint8, device="cuda")
    q = torch.randn(bs, nh, 1, d, device="cuda").half()
    kres, vres = torch.randn(bs, nh_k, 128, d, device="cuda").half(), torch.randn(bs, nh_k, 128, d, device="cuda").half()
    out = M.pq_decode_attn(q, kc, vc, kcent, vcent, kres, vres, r)
    ar = torch.arange(Mm, device="cuda")
    Kh = kcent.float()[ar, kc.long()].reshape(bs, nh_k, nk, d)
    Vh = vcent.float()[ar, vc.long()].reshape(bs, nh_k, nk, d)
    K = torch.cat([Kh, kres.float()], 2).repeat_interleave(4, 1)
    V = torch.cat([Vh, vres.float()], 2).repeat_interleave(4, 1)
    ref = torch.softmax((q.float() @ K.transpose(-1, -2)) / d ** 0.5, -1) @ V
    torch.testing.assert_close(out.float(), ref, atol=ATOL, rtol=RTOL)
    vc.fill_(7)
    out2 = M.pq_decode_attn(q, kc, vc, kcent, vcent, kres, vres, 0)
    want = vcent[:, 7, :].reshape(1, 1, 1, d).float().expand(bs, nh, 1, d)
    torch.testing.assert_close(out2.float(), want, atol=ATOL, rtol=RTOL)


def test_full_size_encode_grid_equals_exact_encoder(M):
    """BASELINE config 2/3 prefill shape (8 kv-heads x 32768 tokens, M=64): the candidate-grid encoder, the tensor-core encoder and
    the exact CUDA-core encoder must agree on every one of the 16.8 M codes (the oracle is too slow at this size: the exact
    encoder is itself pinned to the oracle and the golden vectors by the tests above)."""
    from million_b200 import _lib as L
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    X = torch.randn(1, 8, 32768, 128, device="cuda", generator=g).half()
    X *= torch.where(torch.rand(X.shape, device="cuda", generator=g) < 0.01, 20.0, 1.0).half()      # 1 % heavy entries
    cent = torch.randn(64, 256, 2, device="cuda", generator=g).half().float().contiguous()
    exact = M.pq_encode(X, cent, impl=L.IMPL_GENERIC)
    assert torch.equal(M.pq_encode(X, cent, impl=L.IMPL_GRID), exact)
    assert torch.equal(M.pq_encode(X, cent, impl=L.IMPL_FAST), exact)
    # idempotence through the reconstruction: encode(decode(codes)) == codes (centroids encode to themselves or an equal-distance twin)
    again = M.pq_encode(M.pq_decode(exact, cent.half()), cent, impl=L.IMPL_GRID)
    assert torch.equal(M.pq_decode(again, cent.half()), M.pq_decode(exact, cent.half()))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("bs,nh,nh_k,nk,r", [(1, 32, 8, 5000, 128), (2, 8, 8, 777, 17), (1, 16, 8, 64, 1), (1, 32, 4, 2100, 40), (2, 4, 4, 0, 9)])
def test_attn_two_bit_config_fast_path(M, dtype, bs, nh, nh_k, nk, r):
    """M=32, d_m=4 ('2-bit', BASELINE config 5): the dedicated fast kernel (attn_fast_dm4.cu) against the oracle."""
    from million_b200 import _lib as L
    inp, t = _rand_case(bs, nh, nh_k, nk, r, Mm=32, seed=bs + nh + nk, dtype=dtype)
    out = M.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], r, impl=L.IMPL_FAST)
    f = lambda k: t[k].float().cpu().numpy()
    ref = O.pq_decode_attn(f("q"), inp["kc"], inp["vc"], f("kcent"), f("vcent"), f("kres"), f("vres"), r)
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, atol=ATOL, rtol=RTOL)
