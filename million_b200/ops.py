"""torch-tensor front end of the C ABI: device pointers, strides and the current CUDA stream go straight
into libmillion_b200.so.  torch is used for memory and streams only."""
import ctypes

import torch

from . import _lib as L

_DT = {torch.float16: L.MILLION_F16, torch.bfloat16: L.MILLION_BF16, torch.float32: L.MILLION_F32}


def _dt(t):
    try:
        return _DT[t.dtype if isinstance(t, torch.Tensor) else t]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype if isinstance(t, torch.Tensor) else t}")


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("million_b200 ops run on CUDA tensors only (there is no CPU path)")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


# ---------------------------------------------------------------------------------------------- encode


def _x_view(X):
    """(bs, nh_k, n, d) with contiguous rows -> (ptr tensor, head stride in elements)."""
    bs, nh, n, d = X.shape
    if n and (X.stride(3) != 1 or X.stride(2) != d or (nh > 1 and bs > 1 and X.stride(0) != nh * X.stride(1))):
        X = X.contiguous()
    hs = X.stride(1) if nh > 1 else (X.stride(0) if bs > 1 else n * d)
    return X, hs


_enc_prepared = {}


def prepare_encoder(cent_f32, x_dtype):
    """Operand tiles of the tensor-core encoder for (codebook, input dtype); None when the shape is not covered or the
    centroids are not representable in the input format (then the exact CUDA-core encoder is used)."""
    if x_dtype not in (torch.float16, torch.bfloat16):
        return None
    M, C, dm = cent_f32.shape
    nbytes = L.lib().million_pq_encoder_prepared_bytes(M * dm, M, C)
    if nbytes == 0:
        return None
    key = (cent_f32.data_ptr(), cent_f32._version, x_dtype, cent_f32.device)
    if key not in _enc_prepared:
        if len(_enc_prepared) > 64:
            _enc_prepared.clear()
        buf = torch.empty(nbytes, dtype=torch.uint8, device=cent_f32.device)
        st = L.lib().million_pq_encoder_prepare(_ptr(cent_f32), _DT[x_dtype], M * dm, M, C, _ptr(buf), _stream(cent_f32))
        if st == L.MILLION_ERR_UNSUPPORTED:
            buf = None
        else:
            L.check(st)
        _enc_prepared[key] = (buf, cent_f32)
    return _enc_prepared[key][0]


_grid_prepared = {}


def prepare_encoder_grid(cent_f32):
    """Candidate-grid tables of the exact 2-D encoder (d/M = 2, C <= 256), or None for other shapes.  One asynchronous launch per
    codebook (tensor + version), cached."""
    M, C, dm = cent_f32.shape
    nbytes = L.lib().million_pq_encoder_grid_prepared_bytes(M * dm, M, C)
    if nbytes == 0:
        return None
    key = (cent_f32.data_ptr(), cent_f32._version, cent_f32.device)
    hit = _grid_prepared.get(key)
    if hit is None:
        if len(_grid_prepared) > 64:
            _grid_prepared.clear()
        buf = torch.empty(nbytes, dtype=torch.uint8, device=cent_f32.device)
        L.check(L.lib().million_pq_encoder_grid_prepare(_ptr(cent_f32), M * dm, M, C, _ptr(buf), _stream(cent_f32)))
        hit = (buf, cent_f32)
        _grid_prepared[key] = hit
    return hit[0]


def prepare_encoder_auto(cent_f32, x_dtype):
    """The tables MILLION_IMPL_AUTO expects for this codebook shape (million_pq_encoder_auto_prepare): the candidate grid for
    two-dimensional sub-spaces, the tensor-core tiles otherwise; None = no fast encoder for the shape (AUTO then runs the exact
    CUDA-core encoder).  Same caches as the two specific functions."""
    M, C, dm = cent_f32.shape
    if L.lib().million_pq_encoder_grid_prepared_bytes(M * dm, M, C) > 0:
        return prepare_encoder_grid(cent_f32)
    return prepare_encoder(cent_f32, x_dtype)


def _encoder_choice(X, cent_f32, impl, code_bytes):
    """(impl, prepared) for one encode call.  AUTO is decided INSIDE the library (c_api.cu: grid encoder for d/M = 2, tensor-core
    encoder for d/M = 4, exact CUDA-core encoder otherwise / for unaligned inputs); this only supplies the matching tables."""
    if code_bytes != 1 or impl == L.IMPL_GENERIC:
        return impl, None
    if impl == L.IMPL_GRID:
        grid = prepare_encoder_grid(cent_f32)
        if grid is None or X.data_ptr() % (4 * X.element_size()):
            raise L.MillionError(L.MILLION_ERR_UNSUPPORTED, "grid encoder: shape or alignment not covered")
        return impl, grid
    if impl == L.IMPL_FAST:
        return impl, prepare_encoder(cent_f32, X.dtype)
    return impl, prepare_encoder_auto(cent_f32, X.dtype)


def pq_encode_into(X, cent_f32, codes, *, t0=0, layout="rowmajor", impl=L.IMPL_AUTO):
    """Encode X (bs, nh_k, n, d) and write the codes at token offset t0 of a preallocated cache.
    layout 'rowmajor': codes (bs, nh_k, cap, M); 'transposed': codes (bs, nh_k, M, cap)."""
    _need_cuda(X, cent_f32, codes)
    bs, nh, n, d = X.shape
    M, C, dm = cent_f32.shape
    assert cent_f32.dtype == torch.float32 and cent_f32.is_contiguous() and M * dm == d
    X, xhs = _x_view(X)
    if layout == "rowmajor":
        assert codes.shape[0] == bs and codes.shape[1] == nh and codes.shape[3] == M and codes.stride(3) == 1
        hs, ts, ms = codes.stride(1), codes.stride(2), 1
        assert bs == 1 or codes.stride(0) == nh * codes.stride(1)
    else:
        assert codes.shape[0] == bs and codes.shape[1] == nh and codes.shape[2] == M and codes.stride(3) == 1
        hs, ts, ms = codes.stride(1), 1, codes.stride(2)
        assert bs == 1 or codes.stride(0) == nh * codes.stride(1)
    impl, prep = _encoder_choice(X, cent_f32, impl, codes.element_size())
    L.check(L.lib().million_pq_encode(_ptr(X), _dt(X), xhs, _ptr(cent_f32), _ptr(prep), _ptr(codes), codes.element_size(),
                                      hs, ts, ms, t0, bs * nh, n, d, M, C, impl, _stream(X)))
    return codes


def pq_encode(X, cent_f32, *, out_dtype=torch.uint8, impl=L.IMPL_AUTO):
    """sa_encode_4d_keops semantics: (bs, nh_k, n, d) -> (bs, nh_k, n, M) codes."""
    bs, nh, n, d = X.shape
    M = cent_f32.shape[0]
    store = torch.uint8 if out_dtype == torch.uint8 else torch.int16
    codes = torch.empty(bs, nh, n, M, dtype=store, device=X.device)
    pq_encode_into(X, cent_f32, codes, impl=impl)
    if store != out_dtype:
        codes = (codes.to(torch.int32) & 0xFFFF).to(out_dtype)
    return codes


def pq_encode_paged(X, cent_f32, page_pool, page_ids, *, t0, impl=L.IMPL_AUTO):
    """Encode X (bs, nh_k, n, d) into pool (pages, M, page_size) through page_ids (bs, nh_k, n_pages) int64."""
    _need_cuda(X, cent_f32, page_pool, page_ids)
    bs, nh, n, d = X.shape
    M, C, dm = cent_f32.shape
    assert page_pool.dtype == torch.uint8 and page_pool.is_contiguous() and page_pool.shape[1] == M
    assert page_ids.dtype == torch.int64 and page_ids.is_contiguous() and page_ids.shape[:2] == (bs, nh)
    X, xhs = _x_view(X)
    impl, prep = _encoder_choice(X, cent_f32, impl, 1)
    L.check(L.lib().million_pq_encode_paged(_ptr(X), _dt(X), xhs, _ptr(cent_f32), _ptr(prep), _ptr(page_pool), _ptr(page_ids),
                                            page_ids.shape[2], page_pool.shape[2], t0, bs * nh, n, d, M, C, impl, _stream(X)))


def pq_decode(codes, cent):
    """sa_decode_4d semantics: codes (bs, nh_k, n, M) [any strides with unit M stride], cent (M, C, d_m) -> (bs, nh_k, n, d)."""
    _need_cuda(codes, cent)
    bs, nh, n, M = codes.shape
    Mc, C, dm = cent.shape
    cent = cent.contiguous()
    if codes.dtype not in (torch.uint8, torch.int16, torch.uint16):
        codes = codes.to(torch.int16)
    if n and (codes.stride(3) != 1 or (bs > 1 and nh > 1 and codes.stride(0) != nh * codes.stride(1))):
        codes = codes.contiguous()
    out = torch.empty(bs, nh, n, M * dm, dtype=cent.dtype, device=codes.device)
    hs = codes.stride(1) if nh > 1 else (codes.stride(0) if bs > 1 else n * M)
    L.check(L.lib().million_pq_decode(_ptr(codes), codes.element_size(), hs, codes.stride(2) if n else M, 1,
                                      _ptr(cent), _ptr(out), _dt(cent), n * M * dm, bs * nh, n, M * dm, M, C, _stream(codes)))
    return out


# ---------------------------------------------------------------------------------------------- outlier side store
# Extension (the reference has no outlier code): semantics in DESIGN.md section 3.5.


def outlier_split_into(X, cent_f32, idx_store, val_store, *, t0=0):
    """Top-k_out |x| entries of every head-vector of X (bs, nh_k, n, d) -> records (dim, delta) written at token offset t0 of
    idx_store (bs, nh_k, cap, k_out) uint8 / val_store (same shape, X's dtype).  Returns X with those entries zeroed — the
    tensor to hand to pq_encode*()."""
    _need_cuda(X, cent_f32, idx_store, val_store)
    bs, nh, n, d = X.shape
    M, C, dm = cent_f32.shape
    k_out = idx_store.shape[3]
    assert cent_f32.dtype == torch.float32 and cent_f32.is_contiguous() and M * dm == d
    assert idx_store.dtype == torch.uint8 and val_store.dtype == X.dtype and idx_store.shape == val_store.shape
    assert idx_store.shape[:2] == (bs, nh) and idx_store.stride(3) == 1 and idx_store.stride(2) == k_out
    assert idx_store.stride() == val_store.stride() and (bs == 1 or idx_store.stride(0) == nh * idx_store.stride(1))
    assert t0 + n <= idx_store.shape[2]
    X, xhs = _x_view(X)
    Xm = torch.empty(bs, nh, n, d, dtype=X.dtype, device=X.device)
    L.check(L.lib().million_pq_outlier_split(_ptr(X), _dt(X), xhs, _ptr(cent_f32), _ptr(Xm), _ptr(idx_store), _ptr(val_store),
                                             idx_store.stride(1), t0, bs * nh, n, d, M, C, k_out, _stream(X)))
    return Xm


def pq_encode_outliers(X, cent_f32, k_out, *, impl=L.IMPL_AUTO):
    """(codes (bs, nh_k, n, M) uint8, idx (bs, nh_k, n, k_out) uint8, val (bs, nh_k, n, k_out) X.dtype)."""
    bs, nh, n, d = X.shape
    idx = torch.empty(bs, nh, n, k_out, dtype=torch.uint8, device=X.device)
    val = torch.empty(bs, nh, n, k_out, dtype=X.dtype, device=X.device)
    Xm = outlier_split_into(X, cent_f32, idx, val)
    return pq_encode(Xm, cent_f32, impl=impl), idx, val


def outlier_apply(out, idx, val, *, t0=0):
    """out (bs, nh_k, n, d) += side store records [t0, t0 + n): the reconstruction with outliers (in place)."""
    _need_cuda(out, idx, val)
    bs, nh, n, d = out.shape
    k_out = idx.shape[3]
    assert out.is_contiguous() and idx.dtype == torch.uint8 and idx.stride() == val.stride() and idx.stride(3) == 1 and idx.stride(2) == k_out
    assert bs == 1 or idx.stride(0) == nh * idx.stride(1)
    L.check(L.lib().million_pq_outlier_apply(_ptr(out), _dt(out), n * d, _ptr(idx), _ptr(val), _dt(val), idx.stride(1), t0,
                                             bs * nh, n, d, k_out, _stream(out)))
    return out


def pq_decode_outliers(codes, cent, idx, val):
    """sa_decode_4d + the side store: (bs, nh_k, n, d) in cent's dtype."""
    return outlier_apply(pq_decode(codes, cent), idx, val)


# ---------------------------------------------------------------------------------------------- attention

_workspaces = {}


def attn_workspace(device, bs, nh, nh_k, d, splits):
    """Zero-initialised scratch, cached per device and grown on demand (the kernel leaves its counters zero)."""
    need = L.lib().million_pq_decode_attn_workspace_bytes(bs, nh, nh_k, d, splits)
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def default_splits(bs, nh_k, nk):
    return L.lib().million_pq_decode_attn_default_splits(bs, nh_k, nk)


_prepared = {}


def prepare_codebooks(k_cent, v_cent):
    """fp16 gather tables for the fast kernel (None when the shape is not covered).  Cached per codebook tensor and
    torch version counter, so in-place edits of the centroids are noticed."""
    M, C, dm = k_cent.shape
    nbytes = L.lib().million_pq_codebook_prepared_bytes(M * dm, M, C)
    if nbytes == 0:
        return None
    key = (k_cent.data_ptr(), v_cent.data_ptr(), k_cent._version, v_cent._version, k_cent.dtype, k_cent.device)
    hit = _prepared.get(key)
    if hit is None:
        if len(_prepared) > 64:
            _prepared.clear()
        buf = torch.empty(nbytes, dtype=torch.uint8, device=k_cent.device)
        L.check(L.lib().million_pq_codebook_prepare(_ptr(k_cent), _ptr(v_cent), _dt(k_cent), M * dm, M, C, _ptr(buf), _stream(k_cent)))
        hit = (buf, k_cent, v_cent)          # keep the sources alive so data_ptr keys cannot be recycled
        _prepared[key] = hit
    return hit[0]


def pq_decode_attn(q, k_codes, v_codes, k_cent, v_cent, k_res, v_res, r, *, nk=None, v_layout=L.V_ROWMAJOR,
                   v_page_ids=None, page_size=0, out=None, partial=None, n_splits=0, impl=L.IMPL_AUTO, workspace=None,
                   prepared=None, k_outliers=None, v_outliers=None, p2p=None, k_new=None, v_new=None, r_dev=None, pdl=False):
    """One decode-attention call (include/million_b200.h: million_pq_decode_attn).

    q (bs, nh, 1, d) | (bs, nh, d); k_codes (bs, nh_k, >=nk, M) uint8 (head stride taken from the tensor);
    v_codes: rowmajor (bs, nh_k, >=nk, M) | transposed (bs, nh_k, M, >=nk) | paged pool (pages, M, page_size);
    k_res/v_res (bs, nh_k, Lt, d).  Returns (bs, nh, 1, d) in q's dtype, or fills `partial` (bs, nh, d+2) fp32.
    k_outliers / v_outliers: optional (idx, val) side stores (bs, nh_k, >=nk, k_out) uint8 / q's dtype (extension).
    p2p = state tensor prepared by splitkv_state(): split-KV across GPUs fused into this launch (sharding.SplitKVPeerGroup).
    k_new / v_new (bs, nh_k, 1, d) | (bs, nh_k, d): the token being decoded; the kernel stores it into window row r-1 itself (r
    counts it) — the append of pq_utils.py:304-311 without a copy launch.  r_dev: int32 device scalar, the kernel then uses
    r = *r_dev + r (CUDA-graph replay; see include/million_b200.h).
    """
    _need_cuda(q, k_codes, v_codes, k_cent, v_cent, k_res, v_res)
    bs, nh = q.shape[0], q.shape[1]
    d = q.shape[-1]
    M, C, dm = k_cent.shape
    nh_k = k_res.shape[1] if k_res is not None else k_codes.shape[1]
    if nk is None:
        nk = k_codes.shape[2]
    assert q.is_contiguous() and k_cent.is_contiguous() and v_cent.is_contiguous()
    assert k_cent.dtype == q.dtype and v_cent.dtype == q.dtype, "centroids must be in the query dtype"
    p = L.AttnParams()
    p.struct_size = ctypes.sizeof(L.AttnParams)
    p.io_dtype, p.impl, p.flags = _dt(q), impl, (L.ATTN_PARTIAL_ONLY if partial is not None else 0)
    if pdl:
        p.flags |= L.ATTN_PDL       # programmatic dependent launch: see MILLION_ATTN_PDL in include/million_b200.h
    if p2p is not None:
        assert partial is None
        p.flags |= L.ATTN_FUSED_SPLITKV
        p.p2p_state = p2p.data_ptr()
    p.bs, p.nh, p.nh_k, p.d, p.M, p.C, p.nk, p.r = bs, nh, nh_k, d, M, C, nk, r
    p.q = q.data_ptr()
    if nk:
        # one-byte codes (uint8) or two-byte codes (uint16 / int16 storage: nbits2dtype for nbits > 8); ABI strides are in bytes
        es = k_codes.element_size()
        assert k_codes.dtype in (torch.uint8, torch.uint16, torch.int16) and v_codes.dtype == k_codes.dtype
        assert k_codes.stride(3) == 1 and k_codes.stride(2) == M
        assert bs == 1 or k_codes.stride(0) == nh_k * k_codes.stride(1)
        p.code_bytes = es
        p.k_codes, p.k_head_stride = k_codes.data_ptr(), es * (k_codes.stride(1) if nh_k > 1 or bs > 1 else k_codes.shape[2] * M)
        p.v_codes = v_codes.data_ptr()
        if v_layout == L.V_ROWMAJOR:
            assert v_codes.stride(3) == 1 and v_codes.stride(2) == M
            p.v_head_stride = es * (v_codes.stride(1) if nh_k > 1 or bs > 1 else v_codes.shape[2] * M)
        elif v_layout == L.V_TRANSPOSED:
            assert v_codes.stride(3) == 1
            p.v_head_stride, p.v_ld = es * v_codes.stride(1), es * v_codes.stride(2)
        else:
            assert v_codes.is_contiguous() and v_page_ids.dtype == torch.int64 and v_page_ids.is_contiguous()
            p.v_page_ids, p.n_pages, p.page_size = v_page_ids.data_ptr(), v_page_ids.shape[2], page_size or v_codes.shape[2]
    p.v_layout = v_layout
    if nk:
        for side, store in (("k", k_outliers), ("v", v_outliers)):
            if store is None:
                continue
            idx, val = store
            _need_cuda(idx, val)
            ko = idx.shape[3]
            assert idx.dtype == torch.uint8 and val.dtype == q.dtype and idx.stride() == val.stride()
            assert idx.stride(3) == 1 and idx.stride(2) == ko and idx.shape[2] >= nk
            assert bs == 1 or idx.stride(0) == nh_k * idx.stride(1)
            hs = idx.stride(1) if nh_k > 1 or bs > 1 else idx.shape[2] * ko
            if side == "k":
                p.k_out, p.k_out_idx, p.k_out_val, p.k_out_head_stride = ko, idx.data_ptr(), val.data_ptr(), hs
            else:
                p.v_out, p.v_out_idx, p.v_out_val, p.v_out_head_stride = ko, idx.data_ptr(), val.data_ptr(), hs
    p.k_cent, p.v_cent = k_cent.data_ptr(), v_cent.data_ptr()
    if impl != L.IMPL_GENERIC:
        if prepared is None:
            prepared = prepare_codebooks(k_cent, v_cent)
        if prepared is not None:
            p.prepared_codebook = prepared.data_ptr()
    p.res_len = k_res.shape[2] if k_res is not None else 0
    if k_new is not None:
        _need_cuda(k_new, v_new)
        assert k_new.is_contiguous() and v_new.is_contiguous() and k_new.dtype == q.dtype and v_new.dtype == q.dtype
        assert k_new.numel() == bs * nh_k * d and v_new.numel() == bs * nh_k * d and r >= 1
        p.k_new, p.v_new = k_new.data_ptr(), v_new.data_ptr()
    if r_dev is not None:
        _need_cuda(r_dev)
        assert r_dev.dtype == torch.int32
        p.r_dev = r_dev.data_ptr()
    if r or r_dev is not None:
        assert k_res.dtype == q.dtype and v_res.dtype == q.dtype
        # rows must be contiguous; a [:, :, :r] slice of a larger window is taken in place (res_len from the stride)
        def rows_ok(t):
            return t.stride(3) == 1 and t.stride(2) == d and t.stride(1) % d == 0 and (bs == 1 or t.stride(0) == nh_k * t.stride(1))
        if not (rows_ok(k_res) and rows_ok(v_res) and k_res.stride(1) == v_res.stride(1)):
            assert k_new is None, "the kernel appends to the window in place: it needs contiguous rows"
            k_res, v_res = k_res.contiguous(), v_res.contiguous()
        p.res_len = k_res.stride(1) // d if (nh_k > 1 or bs > 1) else (k_res.shape[2] if (k_new is not None or r_dev is not None) else max(k_res.shape[2], r))
        p.k_res, p.v_res = k_res.data_ptr(), v_res.data_ptr()
    splits = n_splits or default_splits(bs, nh_k, nk)
    ws = workspace if workspace is not None else attn_workspace(q.device, bs, nh, nh_k, d, splits)
    p.workspace, p.workspace_bytes, p.n_splits = ws.data_ptr(), ws.numel(), n_splits   # 0 = library's choice (may schedule flat)
    if partial is not None:
        assert partial.dtype == torch.float32 and partial.is_contiguous() and partial.numel() == bs * nh * (d + 2)
        p.partial = partial.data_ptr()
        ret = partial
    else:
        if out is None:
            out = torch.empty(bs, nh, 1, d, dtype=q.dtype, device=q.device)
        assert out.is_contiguous() and out.dtype == q.dtype
        p.out = out.data_ptr()
        ret = out
    L.check(L.lib().million_pq_decode_attn(ctypes.byref(p), _stream(q)))
    return ret


def splitkv_state(peer_ptrs, rank, rows, device):
    """Protocol state block of one rank for million_splitkv_push_merge / the fused split-KV attention (zeroed; rank, world, rows = bs * nh
    and the peers' symmetric-buffer pointers recorded)."""
    world = len(peer_ptrs)
    state = torch.zeros(L.lib().million_splitkv_state_bytes(), dtype=torch.uint8, device=device)
    arr = (ctypes.c_void_p * world)(*peer_ptrs)
    L.check(L.lib().million_splitkv_state_init(_ptr(state), ctypes.cast(arr, ctypes.c_void_p), rank, world, rows, _stream(state)))
    torch.cuda.current_stream(state.device).synchronize()
    return state


def lse_merge(parts, d, out_dtype):
    """parts (n_parts, rows, d+2) fp32 -> (rows, d) out_dtype."""
    _need_cuda(parts)
    assert parts.dtype == torch.float32 and parts.is_contiguous()
    n_parts, rows = parts.shape[0], parts.shape[1]
    out = torch.empty(rows, d, dtype=out_dtype, device=parts.device)
    L.check(L.lib().million_lse_merge(_ptr(parts), n_parts, rows, d, _ptr(out), _dt(out), _stream(parts)))
    return out


def window_append(k_win, v_win, k_new, v_new, r0):
    """window[:, :, r0:r0+n] = new for K and V in one launch (pq_utils.py:304-311)."""
    _need_cuda(k_win, v_win, k_new, v_new)
    bs, nh, Lt, d = k_win.shape
    n = k_new.shape[2]
    assert r0 + n <= Lt
    if not (k_new.stride(3) == 1 and k_new.stride(2) == d and (bs == 1 or nh == 1 or k_new.stride(0) == nh * k_new.stride(1))):
        k_new = k_new.contiguous()
    if not (v_new.stride(3) == 1 and v_new.stride(2) == d and (bs == 1 or nh == 1 or v_new.stride(0) == nh * v_new.stride(1))):
        v_new = v_new.contiguous()
    if k_new.stride(1) != v_new.stride(1) or k_new.stride(0) != v_new.stride(0):
        k_new, v_new = k_new.contiguous(), v_new.contiguous()
    shs = k_new.stride(1) if nh > 1 else (k_new.stride(0) if bs > 1 else n * d)
    L.check(L.lib().million_window_append(_ptr(k_win), _ptr(v_win), Lt * d, _ptr(k_new), _ptr(v_new), shs, bs * nh, r0, n, d,
                                          _dt(k_win), _stream(k_win)))


def counter_add(ctr, delta):
    """ctr (int32, CUDA) += delta in one tiny launch of the library (graph capturable)."""
    _need_cuda(ctr)
    assert ctr.dtype == torch.int32 and ctr.is_contiguous()
    L.check(L.lib().million_counter_add(_ptr(ctr), ctr.numel(), delta, _stream(ctr)))


def rope_qk(q, k, cos, sin, q_out=None, k_out=None):
    """Rotary embedding of ONE new token per sequence in one launch (million_rope_qk): q (bs, nh, 1, d) or (bs, nh, d),
    k (bs, nh_k, 1, d) or (bs, nh_k, d), cos / sin (bs, 1, d) or (bs, d) — the tensors transformers' apply_rotary_pos_emb gets
    at q_len = 1 (reference modeling_llama.py:500-512).  Returns (q_rot, k_rot) with the shapes of q / k, bit-identical to it."""
    _need_cuda(q)
    d = q.shape[-1]
    bs, nh, nh_k = q.shape[0], q.shape[1], k.shape[1]
    if q.numel() != bs * nh * d or k.numel() != bs * nh_k * d or cos.numel() != bs * d or sin.numel() != bs * d:
        raise ValueError("rope_qk handles one token per sequence (q_len = 1)")
    if not (q.dtype == k.dtype == cos.dtype == sin.dtype):
        raise ValueError("rope_qk: q, k, cos, sin must share a dtype")
    qc, kc = q.reshape(bs, nh, d), k.reshape(bs, nh_k, d)           # views at q_len = 1 (a (bs, 1, nh, d) projection transposed)
    qc, kc = (qc if qc.is_contiguous() else qc.contiguous()), (kc if kc.is_contiguous() else kc.contiguous())
    cc, sc = cos.reshape(bs, d).contiguous(), sin.reshape(bs, d).contiguous()
    q_out = torch.empty(q.shape, dtype=q.dtype, device=q.device) if q_out is None else q_out
    k_out = torch.empty(k.shape, dtype=k.dtype, device=k.device) if k_out is None else k_out
    assert q_out.is_contiguous() and k_out.is_contiguous() and q_out.numel() == qc.numel() and k_out.numel() == kc.numel()
    L.check(L.lib().million_rope_qk(_ptr(qc), _ptr(kc), _ptr(cc), _ptr(sc), _ptr(q_out), _ptr(k_out), _dt(q), bs, nh, nh_k, d, _stream(q)))
    return q_out, k_out


def window_shift(k_win, v_win, shift, rem):
    _need_cuda(k_win, v_win)
    bs, nh, Lt, d = k_win.shape
    L.check(L.lib().million_window_shift(_ptr(k_win), _ptr(v_win), Lt * d, bs * nh, shift, rem, d, _dt(k_win), _stream(k_win)))
