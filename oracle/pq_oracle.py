"""TEST INFRASTRUCTURE ONLY — numpy/torch-CPU oracle for MILLION's PQ KV-cache hot path.

This file is the *checker*, never the thing measured or shipped: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs import it.

Every function restates one piece of the reference (paths relative to /root/reference):

  l2Ns                       scripts/utils/pq_utils.py:8-22
  nbits2dtype                scripts/utils/pq_utils.py:542-552
  pq_encode                  scripts/utils/pq_utils.py:451-499  (sa_encode_4d_keops; the arithmetic
                             itself lives in pykeops==2.2.3 (requirements.txt:80), not vendored: the
                             published semantics of `((x_i - c_j)**2).sum(-1).argmin(dim=2)` are an
                             fp32 squared-L2 arg-min; ties -> lowest index is OUR definition)
  pq_encode_cdist            scripts/utils/pq_utils.py:410-449  (sa_encode_4d, torch.cdist twin)
  pq_decode                  scripts/utils/pq_utils.py:501-540  (sa_decode_4d)
  pq_decode_attn             scripts/utils/pq_utils.py:360-368 + bindings/test_kernel.py:86-90 (the
                             invariant the reference states for its kernel, without the stray
                             is_causal flag) == bindings/Kernel.cuh:11-166,1038-1209,1211-1270
  split_partials/lse_merge   bindings/Kernel.cuh:161-165, 1249-1269
  DynamicPQCacheOracle       scripts/utils/pq_utils.py:98-327 (cache policy)
  PagedPQCacheOracle         scripts/utils/paged_pq_utils.py:130-210, 216-320, 341-385
  paged V layout             scripts/utils/dynamic_paged_pq_utils.py:46-48, 768-811

Pinning: `tests/golden/make_golden.py` imports the reference's own pq_utils.py (pykeops replaced by
a dense torch stand-in) in the build container and stores its outputs; tests/test_oracle_golden.py
checks this oracle against those fixtures.  The decode-attention *kernel* of the reference is CUDA
fp16 and cannot run in the build container: the oracle there follows the invariant the reference
itself writes down, in fp32.  The outlier side store has no reference counterpart (parity unpinned
for that feature only).
"""
from __future__ import annotations

import math
import numpy as np

# ----------------------------------------------------------------------------------------------
# small helpers


def l2Ns(l: int) -> int:
    """Split-count heuristic, scripts/utils/pq_utils.py:8-22."""
    if l > 2048:
        return 32
    elif l > 256:
        return 16
    elif l > 128:
        return 4
    elif l > 64:
        return 2
    return 1


def nbits2dtype(nbits: int):
    """scripts/utils/pq_utils.py:542-552 (numpy dtypes)."""
    if nbits <= 8:
        return np.uint8
    elif nbits <= 16:
        return np.uint16
    elif nbits <= 32:
        return np.uint32
    elif nbits <= 64:
        return np.uint64
    raise ValueError("nbits must be <= 64")


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


# ----------------------------------------------------------------------------------------------
# A.2 encode


def pq_encode(X, C, out_dtype=np.uint8, chunk: int = 4096) -> np.ndarray:
    """Nearest-centroid codes, restating sa_encode_4d_keops (pq_utils.py:451-499).

    X (bs, nh_k, n, d) any float dtype; C (M, c, d_m).  Both are cast to fp32 (pq_utils.py:483-484),
    D[m, t, c] = sum_k fl32((x - c)^2) accumulated left-to-right in fp32, code = first arg-min.
    Returns (bs, nh_k, n, M) in `out_dtype`.
    """
    X = _f32(X)
    C = _f32(C)
    bs, nh, n, d = X.shape
    M, c, d_m = C.shape
    assert d == M * d_m
    Xr = X.reshape(bs * nh * n, M, d_m)
    codes = np.empty((bs * nh * n, M), dtype=np.int64)
    for s in range(0, Xr.shape[0], chunk):
        xs = Xr[s:s + chunk]                                    # (t, M, d_m)
        acc = None
        for k in range(d_m):
            diff = xs[:, :, None, k] - C[None, :, :, k]         # (t, M, c) fp32
            sq = diff * diff
            acc = sq if acc is None else (acc + sq)             # sequential fp32 sum
        codes[s:s + chunk] = np.argmin(acc, axis=-1)            # first minimum
    return codes.reshape(bs, nh, n, M).astype(out_dtype)


def pq_encode_distances(x_vec, C) -> np.ndarray:
    """fp32 distances (M, c) of ONE head-vector to every centroid (same arithmetic as pq_encode)."""
    C = _f32(C)
    M, c, d_m = C.shape
    xs = _f32(x_vec).reshape(M, d_m)
    acc = None
    for k in range(d_m):
        diff = xs[:, None, k] - C[:, :, k]
        sq = diff * diff
        acc = sq if acc is None else (acc + sq)
    return acc


def pq_encode_cdist(X, C, out_dtype=np.uint8) -> np.ndarray:
    """sa_encode_4d (pq_utils.py:410-449): Euclidean distance (sqrt) then arg-min.  Done in fp64 here;
    used only to show the two reference encoders agree except on near-ties."""
    X = np.asarray(X, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    bs, nh, n, d = X.shape
    M, c, d_m = C.shape
    Xr = X.reshape(bs * nh * n, M, 1, d_m)
    dis = np.sqrt(((Xr - C[None]) ** 2).sum(-1))
    return np.argmin(dis, axis=-1).reshape(bs, nh, n, M).astype(out_dtype)


def encode_mismatch_is_near_tie(X, C, codes_a, codes_b, rel: float = 1e-6) -> tuple[int, int]:
    """For every position where two code tensors differ, check the documented carve-out
    (BASELINE.json north_star): the fp32 distances of the two chosen centroids differ by less than
    `rel` relative.  Returns (n_mismatch, n_violations)."""
    X = _f32(X)
    C = _f32(C)
    bs, nh, n, d = X.shape
    M, c, d_m = C.shape
    a = np.asarray(codes_a).astype(np.int64).reshape(-1, M)
    b = np.asarray(codes_b).astype(np.int64).reshape(-1, M)
    Xr = X.reshape(-1, M, d_m)
    idx = np.argwhere(a != b)
    bad = 0
    for t, m in idx:
        x = Xr[t, m].astype(np.float64)
        da = float(((x - C[m, a[t, m]].astype(np.float64)) ** 2).sum())
        db = float(((x - C[m, b[t, m]].astype(np.float64)) ** 2).sum())
        if abs(da - db) > rel * max(da, db, 1e-30):
            bad += 1
    return len(idx), bad


# ----------------------------------------------------------------------------------------------
# A.3 reconstruct


def pq_decode(codes, C) -> np.ndarray:
    """sa_decode_4d (pq_utils.py:501-540): x_hat[m*d_m + k] = C[m, code[m], k]; keeps C's dtype."""
    codes = np.asarray(codes).astype(np.int64)
    C = np.asarray(C)
    bs, nh, n, M = codes.shape
    Mc, c, d_m = C.shape
    assert M == Mc
    out = C[np.arange(M)[None, None, None, :], codes]          # (bs, nh, n, M, d_m)
    return out.reshape(bs, nh, n, M * d_m)


# ----------------------------------------------------------------------------------------------
# A.4 decode attention


def _scores(q, Kc, Kcent, Kres, r):
    """fp32 logits (bs, nh, nk + r): LUT formulation, Interface.cu:49-50 + Kernel.cuh:93-103."""
    q = _f32(q)
    if q.ndim == 4:
        q = q[:, :, 0, :]
    bs, nh, d = q.shape
    Kc = np.asarray(Kc)
    nh_k, nk, M = Kc.shape[1], Kc.shape[2], Kc.shape[3]
    G = nh // nh_k
    Kcent = _f32(Kcent)
    d_m = d // M
    scale = np.float32(1.0 / math.sqrt(d))
    # LUT[b, h, m, c] = <q[b,h,m-chunk], Kcent[m, c]>
    lut = np.einsum('bhmk,mck->bhmc', q.reshape(bs, nh, M, d_m), Kcent).astype(np.float32)
    s = np.empty((bs, nh, nk + r), dtype=np.float32)
    for b in range(bs):
        for h in range(nh):
            hk = h // G                                         # Kernel.cuh:52
            if nk:
                codes = Kc[b, hk].astype(np.int64)             # (nk, M)
                s[b, h, :nk] = lut[b, h][np.arange(M)[None, :], codes].sum(-1, dtype=np.float32)
            if r:
                s[b, h, nk:] = _f32(Kres)[b, hk, :r] @ q[b, h]
    return s * scale


def pq_decode_attn(q, Kc, Vc, Kcent, Vcent, Kres, Vres, r, return_lse: bool = False):
    """fp32 decode attention over nk PQ-coded tokens plus the r most recent fp16 tokens, no mask.

    q (bs, nh, 1, d) | (bs, nh, d); Kc/Vc (bs, nh_k, nk, M) uint8; *cent (M, c, d_m);
    Kres/Vres (bs, nh_k, Lt, d) with rows [0, r) valid.  Returns (bs, nh, 1, d) float32.
    Follows pq_utils.py:360-368 / test_kernel.py:86-90 (the kernel's stated invariant).
    """
    s = _scores(q, Kc, Kcent, Kres, r)
    bs, nh, T = s.shape
    Vc = np.asarray(Vc)
    nh_k, nk = Vc.shape[1], Vc.shape[2]
    G = nh // nh_k
    mx = s.max(-1, keepdims=True)
    p = np.exp(s - mx)
    l = p.sum(-1, keepdims=True, dtype=np.float32)
    Vhat = _f32(pq_decode(Vc, _f32(Vcent))) if nk else np.zeros((bs, nh_k, 0, _f32(Vcent).shape[0] * _f32(Vcent).shape[2]), np.float32)
    d = Vhat.shape[-1]
    out = np.empty((bs, nh, 1, d), dtype=np.float32)
    for b in range(bs):
        for h in range(nh):
            hk = h // G
            V = np.concatenate([Vhat[b, hk], _f32(Vres)[b, hk, :r]], axis=0) if r else Vhat[b, hk]
            out[b, h, 0] = (p[b, h] @ V) / l[b, h]
    if return_lse:
        return out, (np.log(l) + mx)[..., 0]
    return out


def split_partials(q, Kc, Vc, Kcent, Vcent, Kres, Vres, r, bounds):
    """Per-part (o_normalised, lse) over token ranges of the *quantized* tokens plus a final part for
    the fp16 window — the form the reference kernel keeps in its partial buffers
    (Kernel.cuh:161-165, 1204-1208).  `bounds` = list of (start, end) token ranges.
    Empty parts yield (0, -inf) (the reference would produce NaN there: SURVEY Appendix B.7)."""
    s = _scores(q, Kc, Kcent, Kres, r)
    bs, nh, T = s.shape
    Vc = np.asarray(Vc)
    nh_k, nk = Vc.shape[1], Vc.shape[2]
    G = nh // nh_k
    Vhat = _f32(pq_decode(Vc, _f32(Vcent)))
    d = Vhat.shape[-1] if nk else _f32(Vres).shape[-1]
    parts = list(bounds) + [(nk, nk + r)]
    outs = np.zeros((bs, nh, len(parts), d), np.float32)
    lses = np.full((bs, nh, len(parts)), -np.inf, np.float32)
    for b in range(bs):
        for h in range(nh):
            hk = h // G
            V = np.concatenate([Vhat[b, hk], _f32(Vres)[b, hk, :r]], axis=0)
            for i, (a, e) in enumerate(parts):
                if e <= a:
                    continue
                si = s[b, h, a:e]
                mx = si.max()
                p = np.exp(si - mx)
                l = p.sum(dtype=np.float32)
                outs[b, h, i] = (p @ V[a:e]) / l
                lses[b, h, i] = np.log(l) + mx
    return outs, lses


def lse_merge(outs, lses):
    """flash_decoding_reduce_kernel (Kernel.cuh:1211-1270): out = sum_i softmax_i(lse) * out_i.
    outs (..., P, d), lses (..., P).  Parts with lse = -inf get weight 0."""
    outs = _f32(outs)
    lses = _f32(lses)
    L = lses.max(-1, keepdims=True)
    w = np.exp(lses - L)
    w = np.where(np.isfinite(lses), w, 0.0).astype(np.float32)
    den = w.sum(-1, keepdims=True, dtype=np.float32)
    return ((w / den)[..., None] * outs).sum(-2, dtype=np.float32)


# ----------------------------------------------------------------------------------------------
# paged V layout (dynamic_paged_pq_utils.py:46-48, 768-811; paged_pq_utils.py:173-175, 283-284)


def v_codes_transposed(Vc) -> np.ndarray:
    """(bs, nh_k, T, M) -> (bs, nh_k, M, T): the live PagedPQCache value layout."""
    return np.ascontiguousarray(np.asarray(Vc).transpose(0, 1, 3, 2))


def build_page_pool(Vc, page_size: int = 64):
    """Pages (n_pages_total, M, page_size) uint8, zero padded tail, plus the block table
    (bs, nh_k, n_chunks) int64.  Allocation order: chunk-major, then b, then h, ids ascending from
    a fresh pool (dynamic_paged_pq_utils.py:206-225, 776-811)."""
    Vc = np.asarray(Vc)
    bs, nh_k, T, M = Vc.shape
    n_chunks = (T + page_size - 1) // page_size
    pool = np.zeros((n_chunks * bs * nh_k, M, page_size), dtype=Vc.dtype)
    table = np.zeros((bs, nh_k, n_chunks), dtype=np.int64)
    pid = 0
    for ch in range(n_chunks):
        a, e = ch * page_size, min((ch + 1) * page_size, T)
        for b in range(bs):
            for h in range(nh_k):
                pool[pid, :, :e - a] = Vc[b, h, a:e].T
                table[b, h, ch] = pid
                pid += 1
    return pool, table


def gather_pages(pool, table, T: int) -> np.ndarray:
    """Inverse of build_page_pool: (bs, nh_k, T, M)."""
    bs, nh_k, n_chunks = table.shape
    _, M, ps = pool.shape
    out = np.zeros((bs, nh_k, n_chunks * ps, M), dtype=pool.dtype)
    for b in range(bs):
        for h in range(nh_k):
            for ch in range(n_chunks):
                out[b, h, ch * ps:(ch + 1) * ps] = pool[table[b, h, ch]].T
    return out[:, :, :T]


# ----------------------------------------------------------------------------------------------
# A.6 outlier side store — OUR definition, no reference counterpart (parity unpinned: the reference has no
# outlier code at all; these functions are checked only against themselves and the kernels)
#
#   1. outliers of a head-vector x = its k_out entries of largest |x| (fp32), ties -> lowest dim; listed in that order
#   2. x' = x with those entries zeroed; codes = pq_encode(x') (A.2, unchanged)
#   3. delta_i = round_to_io( fp32(x[dim_i]) - C[m_i, code[m_i], k_i] )          (C fp32, the encoder's codebook)
#   4. reconstruction  x_hat = pq_decode(codes, C) ; x_hat[dim_i] += delta_i
#   5. decode attention uses x_hat for K and V:  s_j += scale * sum_i q[kdim_i] * kdelta_i,
#                                                o   += sum_j p_j * sum_i vdelta_i * e[vdim_i]
# Side store layout: idx (bs, nh_k, n, k_out) uint8, val (bs, nh_k, n, k_out) io dtype.  k_out = 0 is the plain path.


def outlier_select(X, k_out: int) -> np.ndarray:
    """(…, d) -> (…, k_out) uint8: dims of the k_out largest |x|, largest first, ties -> lowest dim."""
    a = np.abs(_f32(X))
    order = np.argsort(-a, axis=-1, kind='stable')              # stable: equal |x| keep ascending dim order
    return order[..., :k_out].astype(np.uint8)


def pq_encode_outliers(X, C, k_out: int, io_dtype=np.float16, out_dtype=np.uint8):
    """Returns (codes (bs,nh,n,M), idx (bs,nh,n,k_out) uint8, val (bs,nh,n,k_out) io_dtype)."""
    X = _f32(X)
    C = _f32(C)
    M, c, d_m = C.shape
    idx = outlier_select(X, k_out)
    Xm = X.copy()
    np.put_along_axis(Xm, idx.astype(np.int64), np.float32(0), axis=-1)
    codes = pq_encode(Xm, C, out_dtype)
    ii = idx.astype(np.int64)
    m_i, k_i = ii // d_m, ii % d_m
    code_i = np.take_along_axis(codes.astype(np.int64), m_i, axis=-1)
    cval = C[m_i, code_i, k_i]                                  # fp32
    xval = np.take_along_axis(X, ii, axis=-1)
    val = (xval - cval).astype(np.float32)
    return codes, idx, _round_io(val, io_dtype)


def _round_io(a, io_dtype):
    if io_dtype == 'bf16':
        import torch
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).float().numpy()
    return np.asarray(a, dtype=np.float32).astype(io_dtype)


def pq_decode_outliers(codes, C, idx, val) -> np.ndarray:
    """fp32 reconstruction with the side store applied (step 4)."""
    xh = _f32(pq_decode(codes, _f32(C))).copy()
    ii = np.asarray(idx).astype(np.int64)
    if ii.shape[-1]:
        cur = np.take_along_axis(xh, ii, axis=-1)
        np.put_along_axis(xh, ii, cur + _f32(val), axis=-1)     # dims of one vector are distinct: no collisions
    return xh


def pq_decode_attn_outliers(q, Kc, Vc, Kcent, Vcent, Kres, Vres, r, kout=None, vout=None):
    """fp32 decode attention on the reconstruction with outliers (step 5); kout/vout = (idx, val) or None."""
    q32 = _f32(q)
    if q32.ndim == 4:
        q32 = q32[:, :, 0, :]
    bs, nh, d = q32.shape
    Kc = np.asarray(Kc)
    nh_k, nk = Kc.shape[1], Kc.shape[2]
    G = nh // nh_k
    Khat = pq_decode_outliers(Kc, Kcent, *kout) if kout is not None else _f32(pq_decode(Kc, _f32(Kcent)))
    Vhat = pq_decode_outliers(Vc, Vcent, *vout) if vout is not None else _f32(pq_decode(Vc, _f32(Vcent)))
    scale = np.float32(1.0 / math.sqrt(d))
    out = np.empty((bs, nh, 1, d), dtype=np.float32)
    for b in range(bs):
        for h in range(nh):
            hk = h // G
            K = np.concatenate([Khat[b, hk], _f32(Kres)[b, hk, :r]], axis=0)
            V = np.concatenate([Vhat[b, hk], _f32(Vres)[b, hk, :r]], axis=0)
            s = (K @ q32[b, h]) * scale
            p = np.exp(s - s.max())
            out[b, h, 0] = (p @ V) / p.sum(dtype=np.float32)
    return out


# ----------------------------------------------------------------------------------------------
# A.5 cache policy


class DynamicPQCacheOracle:
    """State machine of DynamicPQCache (pq_utils.py:98-327) on numpy arrays, fp32 attention."""

    def __init__(self, *, bs, nh, num_key_value_heads, M, layer_num, d=128, nbits=8):
        self.bs, self.nh, self.nh_k, self.M, self.L, self.d = bs, nh, num_key_value_heads, M, layer_num, d
        self.dtype = nbits2dtype(nbits)
        self.max_residual_length = d                                   # pq_utils.py:111
        self.init_cache()

    def init_cache(self):                                              # pq_utils.py:116-138
        z = lambda *s: np.zeros(s, dtype=self.dtype)
        self.key_cache = [z(self.bs, self.nh_k, 0, self.M) for _ in range(self.L)]
        self.value_cache = [z(self.bs, self.nh_k, 0, self.M) for _ in range(self.L)]
        zr = lambda: np.zeros((self.bs, self.nh_k, self.max_residual_length, self.d), np.float32)
        self.key_residual_cache = [zr() for _ in range(self.L)]
        self.value_residual_cache = [zr() for _ in range(self.L)]
        self.seen_tokens = [0] * self.L
        self.residualed_tokens = [0] * self.L

    def set_cent(self, key_cent, value_cent):                          # pq_utils.py:149-159
        self.key_cent, self.value_cent = _f32(key_cent), _f32(value_cent)

    def cat_codes(self, kc, vc, layer):                                # pq_utils.py:140-147
        self.key_cache[layer] = np.concatenate([self.key_cache[layer], kc], axis=2)
        self.value_cache[layer] = np.concatenate([self.value_cache[layer], vc], axis=2)
        self.seen_tokens[layer] += kc.shape[2]

    @staticmethod
    def _causal_sdpa(q, k, v):
        q, k, v = _f32(q), _f32(k), _f32(v)
        bs, nh, T, d = q.shape
        G = nh // k.shape[1]
        k = np.repeat(k, G, axis=1)
        v = np.repeat(v, G, axis=1)
        s = np.einsum('bhqd,bhkd->bhqk', q, k) * np.float32(1.0 / math.sqrt(d))
        s = np.where(np.tril(np.ones((T, T), bool))[None, None], s, -np.inf)
        s = s - s.max(-1, keepdims=True)
        p = np.exp(s)
        p /= p.sum(-1, keepdims=True)
        return np.einsum('bhqk,bhkd->bhqd', p, v).astype(np.float32)

    def prefill(self, q, k, v, layer, distort_recent=False):           # pq_utils.py:222-260
        kc = pq_encode(k, self.key_cent, self.dtype)
        vc = pq_encode(v, self.value_cent, self.dtype)
        self.cat_codes(kc, vc, layer)
        if distort_recent is True:
            k, v = pq_decode(kc, self.key_cent), pq_decode(vc, self.value_cent)
        return self._causal_sdpa(q, k, v)

    def decoding(self, q, k, v, layer):                                # pq_utils.py:281-327
        if self.residualed_tokens[layer] == self.max_residual_length:
            kc = pq_encode(self.key_residual_cache[layer], self.key_cent, self.dtype)
            vc = pq_encode(self.value_residual_cache[layer], self.value_cent, self.dtype)
            self.key_cache[layer] = np.concatenate([self.key_cache[layer], kc], axis=2)
            self.value_cache[layer] = np.concatenate([self.value_cache[layer], vc], axis=2)
            self.residualed_tokens[layer] = 0
        r, n = self.residualed_tokens[layer], k.shape[2]
        self.key_residual_cache[layer][:, :, r:r + n] = _f32(k)
        self.value_residual_cache[layer][:, :, r:r + n] = _f32(v)
        self.residualed_tokens[layer] += n
        self.seen_tokens[layer] += n
        return pq_decode_attn(q, self.key_cache[layer], self.value_cache[layer], self.key_cent, self.value_cent,
                              self.key_residual_cache[layer], self.value_residual_cache[layer],
                              self.residualed_tokens[layer])

    def update(self, k, v, layer, distort_recent=False):               # pq_utils.py:166-220
        past = self.key_cache[layer].shape[2]
        if distort_recent:
            self.cat_codes(pq_encode(k, self.key_cent, self.dtype), pq_encode(v, self.value_cent, self.dtype), layer)
            return pq_decode(self.key_cache[layer], self.key_cent), pq_decode(self.value_cache[layer], self.value_cent)
        if past > 0:
            pk = pq_decode(self.key_cache[layer], self.key_cent)
            pv = pq_decode(self.value_cache[layer], self.value_cent)
        self.cat_codes(pq_encode(k, self.key_cent, self.dtype), pq_encode(v, self.value_cent, self.dtype), layer)
        if past > 0:
            return np.concatenate([pk, _f32(k)], axis=2), np.concatenate([pv, _f32(v)], axis=2)
        return _f32(k), _f32(v)


class PagedPQCacheOracle(DynamicPQCacheOracle):
    """Policy of the live PagedPQCache (paged_pq_utils.py): window 128, flush oldest 64 and shift,
    K codes row-major, V codes kept transposed (bs, nh_k, M, T).  The decode output is the intended
    one (no-mask attention over codes + window), not the reference's broken fallback (Appendix B.2).
    seen_tokens counts each token once (the reference double-counts on flush, Appendix B.5)."""

    def __init__(self, *, page_size=64, extended_residual_size=128, **kw):
        self.page_size, self.extended_residual_size = page_size, extended_residual_size
        super().__init__(**kw)

    def init_cache(self):                                              # paged_pq_utils.py:68-116
        self.max_residual_length = self.extended_residual_size
        super().init_cache()
        self.value_cache = [np.zeros((self.bs, self.nh_k, self.M, 0), self.dtype) for _ in range(self.L)]

    def prefill(self, q, k, v, layer, distort_recent=False):           # paged_pq_utils.py:216-320
        kc = pq_encode(k, self.key_cent, self.dtype)
        vc = pq_encode(v, self.value_cent, self.dtype)
        self.key_cache[layer] = np.concatenate([self.key_cache[layer], kc], axis=2)
        self.value_cache[layer] = np.concatenate([self.value_cache[layer], v_codes_transposed(vc)], axis=3)
        self.seen_tokens[layer] += k.shape[2]
        if distort_recent:
            k, v = pq_decode(kc, self.key_cent), pq_decode(vc, self.value_cent)
        return self._causal_sdpa(q, k, v)

    def flush_to_pages(self, layer):                                   # paged_pq_utils.py:130-210
        ps = self.page_size
        if self.residualed_tokens[layer] < ps:
            return
        kc = pq_encode(self.key_residual_cache[layer][:, :, :ps], self.key_cent, self.dtype)
        vc = pq_encode(self.value_residual_cache[layer][:, :, :ps], self.value_cent, self.dtype)
        self.key_cache[layer] = np.concatenate([self.key_cache[layer], kc], axis=2)
        self.value_cache[layer] = np.concatenate([self.value_cache[layer], v_codes_transposed(vc)], axis=3)
        rem = self.residualed_tokens[layer] - ps
        for cache in (self.key_residual_cache[layer], self.value_residual_cache[layer]):
            src = cache[:, :, ps:ps + rem].copy()
            cache[...] = 0
            cache[:, :, :rem] = src
        self.residualed_tokens[layer] = rem

    def decoding_with_pages(self, q, k, v, layer):                     # paged_pq_utils.py:341-385
        if self.residualed_tokens[layer] >= self.extended_residual_size:
            self.flush_to_pages(layer)
        r, n = self.residualed_tokens[layer], k.shape[2]
        self.key_residual_cache[layer][:, :, r:r + n] = _f32(k)
        self.value_residual_cache[layer][:, :, r:r + n] = _f32(v)
        self.residualed_tokens[layer] += n
        self.seen_tokens[layer] += n
        vc = np.ascontiguousarray(self.value_cache[layer].transpose(0, 1, 3, 2))
        return pq_decode_attn(q, self.key_cache[layer], vc, self.key_cent, self.value_cent,
                              self.key_residual_cache[layer], self.value_residual_cache[layer],
                              self.residualed_tokens[layer])


# ----------------------------------------------------------------------------------------------
# synthetic inputs shared by tests, smoke() and bench (seed 42 = configs/default.json:19)


def make_inputs(*, bs, nh, nh_k, nk, d=128, M=64, C=256, Lt=128, seed=42, self_consistent=False,
                dtype=np.float16):
    """Seeded inputs mirroring bindings/test_kernel.py:59-69.  With `self_consistent`, the codes come
    from encoding randn K/V (so attention is tested on data the encoder would really produce)."""
    rng = np.random.default_rng(seed)
    d_m = d // M
    f = lambda *s: rng.standard_normal(s, dtype=np.float32).astype(dtype)
    kcent, vcent = f(M, C, d_m), f(M, C, d_m)
    q = f(bs, nh, 1, d)
    kres, vres = f(bs, nh_k, Lt, d), f(bs, nh_k, Lt, d)
    if self_consistent:
        K, V = f(bs, nh_k, nk, d), f(bs, nh_k, nk, d)
        kc, vc = pq_encode(K, kcent), pq_encode(V, vcent)
    else:
        ct = np.uint8 if C <= 256 else np.uint16              # nbits2dtype
        kc = rng.integers(0, C, (bs, nh_k, nk, M), dtype=ct)
        vc = rng.integers(0, C, (bs, nh_k, nk, M), dtype=ct)
    return dict(q=q, kc=kc, vc=vc, kcent=kcent, vcent=vcent, kres=kres, vres=vres)


# ----------------------------------------------------------------------------------------------
# Codebook training.  The reference calls faiss (pq_utils.py:586-609: IndexPQ.train, niter 25); faiss is a
# third-party dependency absent from the reference tree (faiss-cpu 1.9.0.post1, requirements.txt:18).  Restated here: ONE Lloyd
# iteration of its per-sub-space k-means (assignment = fp32 squared-L2 arg-min, lowest index on ties — A.2; update = mean of the
# assigned points, an empty cluster keeps its centroid).  Parity unpinned by the reference (no vectors, RNG not reproducible).


def rope_qk(q, k, cos, sin):
    """Rotary position embedding of the new token's q and k (transformers' apply_rotary_pos_emb as called by the reference,
    scripts/modeldb/models/modeling_llama.py:500-512): x * cos + rotate_half(x) * sin with rotate_half(x) = cat(-x[d/2:], x[:d/2]),
    evaluated like torch evaluates it elementwise in the tensors' dtype: each product and the sum in fp32, rounded to the dtype.
    q (bs, nh, d), k (bs, nh_k, d), cos / sin (bs, d); numpy float16 / float32 (bf16 callers pass float32 views of bf16 values and
    round with their own helper)."""
    def one(x):
        dt = x.dtype
        h = x.shape[-1] // 2
        rot = np.concatenate([-x[..., h:], x[..., :h]], axis=-1)
        a = (x.astype(np.float32) * cos[:, None, :].astype(np.float32)).astype(dt)
        b = (rot.astype(np.float32) * sin[:, None, :].astype(np.float32)).astype(dt)
        return (a.astype(np.float32) + b.astype(np.float32)).astype(dt)
    return one(q), one(k)


def kmeans_step(X, cent):
    """X (n, d) fp32, cent (M, C, dm) fp32 -> (new cent, codes (n, M), objective w.r.t. the old centroids, counts (M, C))."""
    X = _f32(X)
    cent = _f32(cent)
    n, d = X.shape
    M, C, dm = cent.shape
    codes = pq_encode(X.reshape(1, 1, n, d), cent, out_dtype=np.uint8 if C <= 256 else np.uint16)[0, 0].astype(np.int64)
    new = cent.copy()
    counts = np.zeros((M, C), dtype=np.int64)
    obj = 0.0
    for m in range(M):
        sub = X[:, m * dm:(m + 1) * dm]
        obj += float(((sub - cent[m, codes[:, m]]) ** 2).sum(dtype=np.float64))
        counts[m] = np.bincount(codes[:, m], minlength=C)
        sums = np.zeros((C, dm), dtype=np.float64)
        np.add.at(sums, codes[:, m], sub)
        nz = counts[m] > 0
        new[m, nz] = (sums[nz] / counts[m, nz, None]).astype(np.float32)
    return new, codes, obj, counts
