"""B200-native counterpart of the reference's scripts/utils/pq_utils.py — same names, arguments and
semantics; different machinery:

  * `sa_encode_4d_keops` / `sa_encode_4d` / `sa_decode_4d` call the CUDA library (pq_utils.py:410-540),
  * `DynamicPQCache` keeps codes in a PREALLOCATED, geometrically grown cache and appends in place
    (the reference re-allocates the whole cache with torch.cat on every flush, pq_utils.py:145-146,
    297-298), and quantizes a full window ASYNCHRONOUSLY on a side stream as soon as it fills
    (the reference encodes synchronously inside the next decode step, pq_utils.py:287-302),
  * `KernelRegistry` resolves the same kernel names (pq_utils.py:64) through `million_b200.bindings`.

There is no CPU path: every compute entry point needs CUDA tensors and the built extension.
"""
import contextlib
import gc

import torch

from . import _lib as L
from . import bindings as _bindings
from . import ops


# ------------------------------------------------------------------------------------------------ helpers


def l2Ns(l):
    """Split-count heuristic of the reference (pq_utils.py:8-22).  Only used to form kernel NAMES; the
    B200 kernel picks its own split count from the SM count."""
    if l > 2048:
        return 32
    elif l > 256:
        return 16
    elif l > 128:
        return 4
    elif l > 64:
        return 2
    else:
        return 1


def scalarTypeToStr(scalar_t):
    """pq_utils.py:24-30, plus bf16 (new)."""
    if scalar_t == torch.float32:
        return 'f32'
    elif scalar_t == torch.float16:
        return 'f16'
    elif scalar_t == torch.bfloat16:
        return 'bf16'
    else:
        raise ValueError(f"Unknown scalar type: {scalar_t}")


def nbits2dtype(nbits):
    """pq_utils.py:542-552."""
    if nbits <= 8:
        return torch.uint8
    elif nbits <= 16:
        return torch.uint16
    elif nbits <= 32:
        return torch.uint32
    elif nbits <= 64:
        return torch.uint64
    else:
        raise ValueError("nbits must be <= 64")


class Singleton(type):
    """Per-class singleton metaclass (scripts/utils/Singleton.py:3-27).  `has_instance` is per class here;
    the reference's is global across every singleton class — not replicated."""
    _instances = {}

    def __call__(cls, *args, **kwargs):
        if cls not in Singleton._instances:
            Singleton._instances[cls] = super().__call__(*args, **kwargs)
        return Singleton._instances[cls]

    def has_instance(cls):
        return cls in Singleton._instances

    def clear_instance(cls=None):
        if cls is None or cls is Singleton:
            Singleton._instances.clear()
        else:
            Singleton._instances.pop(cls, None)


# ------------------------------------------------------------------------------------------------ codec


def _cent_f32(C):
    if C.dim() != 3:
        raise NotImplementedError("per-head codebooks (5-D C) are not used by the reference's callers")
    return C.to(torch.float32).contiguous()


def sa_encode_4d_keops(X, C, target_dtype=torch.uint8):
    """Nearest-codeword indices, (bs, num_heads, n, d) x (M, c, d//M) -> (bs, num_heads, n, M)
    (pq_utils.py:451-499).  fp32 squared-L2 arg-min on the GPU; lowest index on ties."""
    return ops.pq_encode(X, _cent_f32(C), out_dtype=target_dtype)


def sa_encode_4d(X, C, target_dtype=torch.uint8):
    """pq_utils.py:410-449 (torch.cdist twin).  Same arg-min; the sqrt of cdist is monotone, so the codes
    agree with sa_encode_4d_keops except on float near-ties of the reference's own arithmetic."""
    return ops.pq_encode(X, _cent_f32(C), out_dtype=target_dtype)


def sa_decode_4d(codes, C):
    """Reconstruct (bs, num_heads, n, M) codes -> (bs, num_heads, n, d) in C's dtype (pq_utils.py:501-540)."""
    if C.dim() != 3:
        raise NotImplementedError("per-head codebooks (5-D C) are not used by the reference's callers")
    return ops.pq_decode(codes, C)


# ------------------------------------------------------------------------------------------------ registry


class KernelRegistry:
    """pq_utils.py:32-96: resolves `flash_decoding_allocated_buffer_<...>` by name and returns a closure that
    hides the partial buffers.  The partial buffers are still allocated (attribute compatibility,
    pq_utils.py:73-77) but the B200 kernel keeps its fp32 partials in its own workspace."""

    def __init__(self, *, M=64, d=128, nbits=8, nh=32, scalar_t=torch.float32, bs=1):
        self.kernels = {}
        self.partial_out_buffers = {}
        self.partial_lse_buffers = {}
        self.M, self.d, self.nbits, self.nh, self.scalar_t, self.bs = M, d, nbits, nh, scalar_t, bs

    def get_kernel(self, l=4096):
        Ns = l2Ns(l)
        if Ns not in self.kernels:
            self.kernels[Ns] = self.get_custom_kernel_with_allocated_buffer(l)
        return self.kernels[Ns]

    def get_custom_kernel_with_allocated_buffer(self, l=4096):
        M, d, nbits, nh, scalar_t = self.M, self.d, self.nbits, self.nh, self.scalar_t
        if nbits != 8:
            raise NotImplementedError("Only uint8 code type is supported for now")
        Ns = l2Ns(l)
        scalar_str = scalarTypeToStr(scalar_t)
        fname = f"flash_decoding_allocated_buffer_{scalar_str}u8_Ns{Ns}Lt{d}d{d}M{M}C{2**nbits}"
        try:
            func = getattr(_bindings, fname)
        except Exception as e:
            print(f"{fname} not available in million_b200.bindings (supported scalars: f16, bf16).")
            raise e
        partial_out_buffer = torch.empty(self.bs, nh, Ns + 1, d, dtype=scalar_t, device='cuda')
        partial_lse_buffer = torch.empty(self.bs, nh, Ns + 1, dtype=scalar_t, device='cuda')
        self.partial_out_buffers[Ns] = partial_out_buffer
        self.partial_lse_buffers[Ns] = partial_lse_buffer

        def flash_decoding(query, key_codes, value_codes, key_cents, value_cents, key_residuals, value_residuals, r):
            return func(query, key_codes, value_codes, key_cents, value_cents, key_residuals, value_residuals, r,
                        partial_out_buffer, partial_lse_buffer)

        return flash_decoding


# ------------------------------------------------------------------------------------------------ cache


class _CodeStore:
    """Preallocated (bs, nh_k, cap, M) code cache with amortised-doubling growth and in-place append."""

    def __init__(self, bs, nh_k, M, dtype, device, transposed=False):
        self.bs, self.nh_k, self.M, self.dtype, self.device, self.transposed = bs, nh_k, M, dtype, device, transposed
        self.cap, self.len = 0, 0
        self.buf = self._alloc(0)

    def _alloc(self, cap):
        shape = (self.bs, self.nh_k, self.M, cap) if self.transposed else (self.bs, self.nh_k, cap, self.M)
        return torch.zeros(shape, dtype=self.dtype, device=self.device)

    def reserve(self, n):
        if n <= self.cap:
            return
        cap = max(n, 2 * self.cap, 1024)
        cap = (cap + 127) // 128 * 128
        new = self._alloc(cap)
        if self.len:
            if self.transposed:
                new[:, :, :, :self.len] = self.buf[:, :, :, :self.len]
            else:
                new[:, :, :self.len] = self.buf[:, :, :self.len]
        self.buf, self.cap = new, cap

    def view(self, n=None):
        n = self.len if n is None else n
        return self.buf[:, :, :, :n] if self.transposed else self.buf[:, :, :n]


class _OutlierStore:
    """Side store of one layer and one of K/V: (dim uint8, delta fp16/bf16) records, (bs, nh_k, cap, k_out), grown with
    the code store it belongs to.  Extension — the reference has no outlier code."""

    def __init__(self, bs, nh_k, k_out, val_dtype, device):
        self.bs, self.nh_k, self.k_out, self.val_dtype, self.device = bs, nh_k, k_out, val_dtype, device
        self.cap = 0
        self.idx = torch.zeros((bs, nh_k, 0, k_out), dtype=torch.uint8, device=device)
        self.val = torch.zeros((bs, nh_k, 0, k_out), dtype=val_dtype, device=device)

    def reserve(self, cap, keep):
        if cap <= self.cap:
            return
        idx = torch.zeros((self.bs, self.nh_k, cap, self.k_out), dtype=torch.uint8, device=self.device)
        val = torch.zeros((self.bs, self.nh_k, cap, self.k_out), dtype=self.val_dtype, device=self.device)
        if keep:
            idx[:, :, :keep] = self.idx[:, :, :keep]
            val[:, :, :keep] = self.val[:, :, :keep]
        self.idx, self.val, self.cap = idx, val, cap

    def view(self, n):
        return self.idx[:, :, :n], self.val[:, :, :n]


class DynamicPQCache(metaclass=Singleton):
    """pq_utils.py:98-408.  Same constructor, attributes and methods.

    key_cache[layer] / value_cache[layer] are views (bs, nh_k, T, M) into preallocated stores.
    `async_flush=True` (default) starts quantizing a full window on a side stream right after the step
    that filled it; the next step only waits on an event.  Results are identical either way."""

    def __init__(self, *, bs, nh, num_key_value_heads, M, layer_num, dtype=torch.uint8, nbits=8, d=128,
                 scalar_t=torch.float32, async_flush=True, device='cuda', outliers=(0, 0)):
        """`outliers=(k_out_K, k_out_V)` (extension, default off): that many largest-|x| entries of every key / value vector
        bypass PQ and are kept as (dim, delta) records; see DESIGN.md section 3.5 for the exact semantics."""
        self.bs, self.nh, self.num_key_value_heads, self.M = bs, nh, num_key_value_heads, M
        self.k_out, self.v_out = (int(outliers), int(outliers)) if isinstance(outliers, int) else (int(outliers[0]), int(outliers[1]))
        if (self.k_out or self.v_out) and scalar_t not in (torch.float16, torch.bfloat16):
            raise ValueError("the outlier side store needs scalar_t float16 or bfloat16")
        self.layer_num, self.dtype, self.nbits, self.d, self.scalar_t = layer_num, dtype, nbits, d, scalar_t
        self.device = torch.device(device)
        self.async_flush = async_flush
        self.max_residual_length = d  # Lt = d (pq_utils.py:111)
        self.registery = KernelRegistry(M=M, d=d, nbits=nbits, nh=nh, scalar_t=scalar_t, bs=bs)
        self._side = None
        self.init_cache()

    # ---- storage
    def _new_store(self, transposed=False):
        return _CodeStore(self.bs, self.num_key_value_heads, self.M, self.dtype, self.device, transposed)

    def init_cache(self):
        """pq_utils.py:116-138."""
        if not self.device.type == 'cuda':
            raise RuntimeError("DynamicPQCache needs a CUDA device (no CPU path)")
        # A reset (main_pq.py:363 uses init_cache as cache_clear_func) must not free memory an in-flight flush still writes,
        # and must drop every cached raw pointer into the old windows / stores.
        for pend in getattr(self, '_pending', ()):
            if pend is not None:
                pend[0].synchronize()
        self._plans = {}
        self._step_graph = None
        self._k = [self._new_store() for _ in range(self.layer_num)]
        self._v = [self._new_store() for _ in range(self.layer_num)]
        mk = lambda ko: [_OutlierStore(self.bs, self.num_key_value_heads, ko, self.scalar_t, self.device) if ko else None
                         for _ in range(self.layer_num)]
        self._ko, self._vo = mk(self.k_out), mk(self.v_out)
        z = lambda: torch.zeros((self.bs, self.num_key_value_heads, self.max_residual_length, self.d),
                                dtype=self.scalar_t, device=self.device)
        self.key_residual_cache = [z() for _ in range(self.layer_num)]
        self.value_residual_cache = [z() for _ in range(self.layer_num)]
        self.seen_tokens = [0 for _ in range(self.layer_num)]
        self.residualed_tokens = [0 for _ in range(self.layer_num)]
        self._pending = [None for _ in range(self.layer_num)]   # (event, n_tokens) of an in-flight async flush

    @property
    def key_cache(self):
        return [s.view() for s in self._k]

    @property
    def value_cache(self):
        return [s.view() for s in self._v]

    def set_cent(self, key_cent, value_cent):
        """cent is of shape (M, c, d//M) (pq_utils.py:149-159).  fp32 copies feed the encoder (the reference
        casts per call, pq_utils.py:484)."""
        if key_cent is value_cent:
            self._cent = key_cent.contiguous()
            self.key_cent = self._cent
            self.value_cent = self._cent
        else:
            self.key_cent = key_cent.contiguous()
            self.value_cent = value_cent.contiguous()
        self._key_cent_f32 = self.key_cent.to(device=self.device, dtype=torch.float32).contiguous()
        self._value_cent_f32 = self.value_cent.to(device=self.device, dtype=torch.float32).contiguous()
        # new codebooks: the attention-side casts, their gather tables and the per-layer launch plans are stale
        self._attn_dtype = None
        self._plans = {}
        self._step_graph = None
        # encoder tables are built here, on the caller's stream, not lazily inside the first flush (which may run on the side
        # stream): every later use is ordered after this point
        for c in {id(self._key_cent_f32): self._key_cent_f32, id(self._value_cent_f32): self._value_cent_f32}.values():
            ops.prepare_encoder_grid(c)

    def _attn_cents(self, dtype):
        """Centroids in the query dtype for the attention kernel (main_pq.py:258-260 casts them to the model dtype)."""
        if getattr(self, '_attn_dtype', None) != dtype:
            self._kc_attn = self.key_cent.to(device=self.device, dtype=dtype).contiguous()
            self._vc_attn = self.value_cent.to(device=self.device, dtype=dtype).contiguous()
            self._attn_dtype = dtype
        return self._kc_attn, self._vc_attn

    # ---- code append
    def _reserve(self, layer_idx, n):
        ks, vs = self._k[layer_idx], self._v[layer_idx]
        ks.reserve(ks.len + n)
        vs.reserve(vs.len + n)
        if self._ko[layer_idx] is not None:
            self._ko[layer_idx].reserve(ks.cap, ks.len)
        if self._vo[layer_idx] is not None:
            self._vo[layer_idx].reserve(vs.cap, vs.len)

    def _encode_kv(self, key_states, value_states, layer_idx):
        """Encode at the current end of the stores (which must already be reserved); with the side store on, the outliers are
        split off first and the encoder sees the masked vectors."""
        ks, vs = self._k[layer_idx], self._v[layer_idx]
        ko, vo = self._ko[layer_idx], self._vo[layer_idx]
        if ko is not None:
            key_states = ops.outlier_split_into(key_states, self._key_cent_f32, ko.idx, ko.val, t0=ks.len)
        if vo is not None:
            value_states = ops.outlier_split_into(value_states, self._value_cent_f32, vo.idx, vo.val, t0=vs.len)
        ops.pq_encode_into(key_states, self._key_cent_f32, ks.buf, t0=ks.len)
        ops.pq_encode_into(value_states, self._value_cent_f32, vs.buf, t0=vs.len,
                           layout="transposed" if vs.transposed else "rowmajor")

    def _reconstruct(self, layer_idx, which, n_last=None, dtype=None):
        """sa_decode_4d of the last n_last (default: all) tokens of a store, side store applied."""
        st, cent, out = (self._k, self.key_cent, self._ko) if which == 'k' else (self._v, self.value_cent, self._vo)
        st, out = st[layer_idx], out[layer_idx]
        n = st.len if n_last is None else n_last
        codes = st.view()[:, :, :, st.len - n:].transpose(2, 3) if st.transposed else st.view()[:, :, st.len - n:]
        x = sa_decode_4d(codes, cent if dtype is None else cent.to(dtype))
        if out is not None and n:
            ops.outlier_apply(x, out.idx, out.val, t0=st.len - n)
        return x

    def _encode_append(self, key_states, value_states, layer_idx, count_seen=True):
        self._complete_pending_flush(layer_idx)
        n = key_states.size(2)
        ks, vs = self._k[layer_idx], self._v[layer_idx]
        self._reserve(layer_idx, n)
        self._encode_kv(key_states, value_states, layer_idx)
        ks.len += n
        vs.len += n
        if count_seen:
            self.seen_tokens[layer_idx] += n
        return ks.view()[:, :, ks.len - n:], n

    def cat_codes(self, key_codes, value_codes, layer_idx):
        """pq_utils.py:140-147 — kept for API compatibility: appends ready-made codes in place."""
        self._complete_pending_flush(layer_idx)
        n = key_codes.size(2)
        ks, vs = self._k[layer_idx], self._v[layer_idx]
        ks.reserve(ks.len + n)
        vs.reserve(vs.len + n)
        ks.buf[:, :, ks.len:ks.len + n] = key_codes
        vs.buf[:, :, vs.len:vs.len + n] = value_codes
        ks.len += n
        vs.len += n
        self.seen_tokens[layer_idx] += n

    # ---- torch-PQ path
    def update(self, key_states, value_states, layer_idx, distort_recent=False):
        """pq_utils.py:166-220."""
        assert hasattr(self, 'key_cent') and hasattr(self, 'value_cent')
        past_length = self._k[layer_idx].len
        if distort_recent:
            self._encode_append(key_states, value_states, layer_idx)
            return self._reconstruct(layer_idx, 'k'), self._reconstruct(layer_idx, 'v')
        if past_length > 0:
            past_key_states = self._reconstruct(layer_idx, 'k')
            past_value_states = self._reconstruct(layer_idx, 'v')
        self._encode_append(key_states, value_states, layer_idx)
        if past_length > 0:
            key_states = torch.cat([past_key_states.to(key_states.dtype), key_states], dim=2)
            value_states = torch.cat([past_value_states.to(value_states.dtype), value_states], dim=2)
        return key_states, value_states

    # ---- prefill
    def _prefill_attention(self, query_states, key_states, value_states):
        from torch.nn.functional import scaled_dot_product_attention as sdpa
        return sdpa(query_states, key_states, value_states, is_causal=True,
                    enable_gqa=query_states.size(1) != key_states.size(1))

    def prefill(self, query_states, key_states, value_states, layer_idx, distort_recent=False):
        """pq_utils.py:222-260: quantize ALL prefill tokens, then causal SDPA on the fp16 (or, with
        distort_recent, the reconstructed) K/V.  GQA is handled inside SDPA instead of repeat_kv."""
        n = key_states.size(2)
        self._encode_append(key_states, value_states, layer_idx)
        if distort_recent is True:
            key_states = self._reconstruct(layer_idx, 'k', n, key_states.dtype)
            value_states = self._reconstruct(layer_idx, 'v', n, value_states.dtype)
        return self._prefill_attention(query_states, key_states, value_states)

    def residual_attention(self, query_states, layer_idx):
        """pq_utils.py:262-279 (deprecated in the reference: done by the kernel).  Kept for parity checks."""
        key_states = self.key_residual_cache[layer_idx]
        value_states = self.value_residual_cache[layer_idx]
        G = self.nh // self.num_key_value_heads
        k = key_states.repeat_interleave(G, dim=1).float()
        v = value_states.repeat_interleave(G, dim=1).float()
        S = (self.d ** -0.5) * query_states.float() @ k.transpose(-2, -1)
        S[:, :, :, self.residualed_tokens[layer_idx]:] = float('-inf')
        out = torch.softmax(S, dim=-1) @ v
        lse = torch.logsumexp(S, dim=-1, keepdim=False)
        return out, lse

    # ---- flush machinery
    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def _flush_window(self, layer_idx, n):
        """Encode rows [0, n) of the window and append them (pq_utils.py:288-301).  seen_tokens already counts them."""
        self._encode_append(self.key_residual_cache[layer_idx][:, :, :n], self.value_residual_cache[layer_idx][:, :, :n],
                            layer_idx, count_seen=False)

    def _start_async_flush(self, layer_idx, n):
        """Launch the flush of a full window on the side stream; the code stores are grown on the main stream
        first so the side stream never allocates memory the main stream is using."""
        self._reserve(layer_idx, n)
        main = torch.cuda.current_stream(self.device)
        side = self._side_stream()
        side.wait_stream(main)                      # window rows are complete, caches are allocated
        with torch.cuda.stream(side):
            self._encode_kv(self.key_residual_cache[layer_idx][:, :, :n], self.value_residual_cache[layer_idx][:, :, :n], layer_idx)
            ev = torch.cuda.Event()
            ev.record(side)
        self._pending[layer_idx] = (ev, n)

    def _finish_async_flush(self, layer_idx):
        ev, n = self._pending[layer_idx]
        torch.cuda.current_stream(self.device).wait_event(ev)   # device-side wait; the host does not block
        self._k[layer_idx].len += n
        self._v[layer_idx].len += n
        self._pending[layer_idx] = None

    def _window_after_flush(self, layer_idx):
        """Window bookkeeping that belongs to a completed flush: the whole window was retired (pq_utils.py:299-301)."""
        self.residualed_tokens[layer_idx] = 0

    def _complete_pending_flush(self, layer_idx):
        """The ONE place an in-flight flush is consumed: its codes become visible and the window forgets the flushed rows,
        whoever asks first (the next decode step, a prefill/update/cat_codes that follows, a reset).  Returns True if there was one."""
        if self._pending[layer_idx] is None:
            return False
        self._finish_async_flush(layer_idx)
        self._window_after_flush(layer_idx)
        return True

    # ---- decode
    def _window_full(self, layer_idx):
        return self.residualed_tokens[layer_idx] == self.max_residual_length

    def _flush_len(self):
        return self.max_residual_length

    def _retire_window(self, layer_idx):
        """pq_utils.py:288-301: the full window becomes codes (already in flight when async_flush started it)."""
        if not self._complete_pending_flush(layer_idx):
            self._flush_window(layer_idx, self._flush_len())
            self._window_after_flush(layer_idx)

    def _fused_ok(self, q, k, v):
        """One C call per layer (append fused into the attention launch): single token, cache dtype, plain aligned tensors."""
        return (k.size(2) == 1 and self.nbits == 8 and q.dtype == self.scalar_t and k.dtype == self.scalar_t and v.dtype == self.scalar_t
                and q.is_contiguous() and k.is_contiguous() and v.is_contiguous()
                and (k.data_ptr() | v.data_ptr()) % 16 == 0 and type(self)._fast_decode_ok)

    def decoding(self, query_states, key_states, value_states, layer_idx, out=None):
        """pq_utils.py:281-327: flush a full window, append the new token to the window, run the kernel.
        `out` (extension): a preallocated (bs, nh, 1, d) result buffer."""
        if self._window_full(layer_idx):
            self._retire_window(layer_idx)
        r = self.residualed_tokens[layer_idx]
        n = key_states.size(2)
        if self._fused_ok(query_states, key_states, value_states):
            out = self._decode_fast(query_states, key_states, value_states, layer_idx, r, out)
        else:
            ops.window_append(self.key_residual_cache[layer_idx], self.value_residual_cache[layer_idx], key_states, value_states, r)
            self.residualed_tokens[layer_idx] += n
            self.seen_tokens[layer_idx] += n
            res = self._decode_registry(query_states, layer_idx)
            out = res if out is None else out.copy_(res)
        if self.async_flush and self._window_full(layer_idx):
            self._start_async_flush(layer_idx, self._flush_len())
        return out

    def _decode_registry(self, query_states, layer_idx):
        """The reference's route: KernelRegistry -> named binding (pq_utils.py:313-327); side stores go through ops directly."""
        kc, vc = self._attn_cents(query_states.dtype)
        ks = self._k[layer_idx]
        if self.k_out or self.v_out or self.nbits != 8:
            # the reference's 10-argument kernel signature (pq_utils.py:83-94) has no place for the side store, and its registry
            # refuses nbits != 8 ("Only uint8 code type is supported for now", pq_utils.py:58-59): two-byte codes go straight
            # to the C ABI (code_bytes = 2, all-shapes kernel)
            return ops.pq_decode_attn(query_states, ks.view(), self._v[layer_idx].view(), kc, vc,
                                      self.key_residual_cache[layer_idx], self.value_residual_cache[layer_idx],
                                      self.residualed_tokens[layer_idx],
                                      k_outliers=self._ko[layer_idx].view(ks.len) if self.k_out else None,
                                      v_outliers=self._vo[layer_idx].view(ks.len) if self.v_out else None)
        kernel = self.registery.get_kernel(l=self.seen_tokens[layer_idx])
        return kernel(query_states, ks.view(), self._v[layer_idx].view(), kc, vc,
                      self.key_residual_cache[layer_idx], self.value_residual_cache[layer_idx],
                      self.residualed_tokens[layer_idx])

    # ---- the same step as ONE C-ABI call per layer (window append fused into the attention launch), without the Python layers
    _fast_decode_ok = True

    def _fill_v(self, p, layer_idx):
        vs = self._v[layer_idx]
        p.v_layout, p.v_codes, p.v_head_stride = L.V_ROWMAJOR, vs.buf.data_ptr(), vs.cap * self.M * vs.buf.element_size()

    def _plan(self, layer_idx, dtype, device):
        """Per-layer, pre-filled million_attn_params (the per-call Python work is what bounds batch-1 decode).  Dropped by
        init_cache() and set_cent(): it holds raw pointers into the windows and the codebook tables."""
        import ctypes
        plan = self._plans.get(layer_idx)
        kc, vc = self._attn_cents(dtype)
        if plan is None or plan['cent'] is not kc:
            lib = L.lib()
            p = L.AttnParams()
            p.struct_size = ctypes.sizeof(L.AttnParams)
            p.io_dtype = ops._DT[dtype]
            # PDL: the launch may begin (table copies, first code tiles) while its stream predecessor drains; nothing the
            # predecessor can write — q, k/v of the new token, window, workspace — is read before it has completed
            p.impl, p.flags = L.IMPL_AUTO, L.ATTN_PDL
            p.bs, p.nh, p.nh_k, p.d, p.M, p.C = self.bs, self.nh, self.num_key_value_heads, self.d, self.M, kc.shape[1]
            p.v_layout = L.V_ROWMAJOR
            p.k_cent, p.v_cent = kc.data_ptr(), vc.data_ptr()
            p.k_res, p.v_res = self.key_residual_cache[layer_idx].data_ptr(), self.value_residual_cache[layer_idx].data_ptr()
            p.res_len = self.max_residual_length
            prepared = ops.prepare_codebooks(kc, vc)
            if prepared is not None:
                p.prepared_codebook = prepared.data_ptr()
            max_splits = ops.default_splits(self.bs, self.num_key_value_heads, 1 << 30)
            ws = ops.attn_workspace(device, self.bs, self.nh, self.num_key_value_heads, self.d, max_splits)
            p.workspace, p.workspace_bytes, p.n_splits = ws.data_ptr(), ws.numel(), 0
            plan = dict(p=p, ref=ctypes.byref(p), cent=kc, keep=(prepared, ws, vc), attn=lib.million_pq_decode_attn)
            self._plans[layer_idx] = plan
        return plan

    def _fill_stores(self, p, layer_idx):
        """Current code stores (and side stores) of a layer -> the launch parameters."""
        ks = self._k[layer_idx]
        p.nk = ks.len
        if ks.len:
            p.k_codes, p.k_head_stride = ks.buf.data_ptr(), ks.cap * self.M * ks.buf.element_size()
            self._fill_v(p, layer_idx)
            if self.k_out:
                ko = self._ko[layer_idx]
                p.k_out, p.k_out_idx, p.k_out_val, p.k_out_head_stride = self.k_out, ko.idx.data_ptr(), ko.val.data_ptr(), ko.cap * self.k_out
            if self.v_out:
                vo = self._vo[layer_idx]
                p.v_out, p.v_out_idx, p.v_out_val, p.v_out_head_stride = self.v_out, vo.idx.data_ptr(), vo.val.data_ptr(), vo.cap * self.v_out

    def _decode_fast(self, q, k, v, layer_idx, r, out=None):
        """One decode step of one layer: what decoding() does through window_append + KernelRegistry -> bindings -> ops, as one
        launch (k_new / v_new of million_attn_params: the kernel stores the new token into window row r itself)."""
        import ctypes
        plan = self._plan(layer_idx, q.dtype, q.device)
        p = plan['p']
        if out is None:
            out = torch.empty(self.bs, self.nh, 1, self.d, dtype=q.dtype, device=q.device)
        p.q, p.out, p.k_new, p.v_new, p.r, p.r_dev = q.data_ptr(), out.data_ptr(), k.data_ptr(), v.data_ptr(), r + 1, None
        self._fill_stores(p, layer_idx)
        st = plan['attn'](plan['ref'], ctypes.c_void_p(torch.cuda.current_stream(q.device).cuda_stream))
        if st:
            L.check(st)
        self.residualed_tokens[layer_idx] = r + 1
        self.seen_tokens[layer_idx] += 1
        return out

    def decode_step_graph(self, q, k, v, out):
        """CUDA-graph replay of one decode step of ALL layers over static buffers (extension; the reference calls decoding()
        layer by layer from Python, modeling_llama.py:529-547): q/out (layers, bs, nh, 1, d), k/v (layers, bs, nh_k, 1, d).
        Returns a DecodeStepGraph; call .step() once per token after filling q, k, v."""
        return DecodeStepGraph(self, q, k, v, out)

    # ---- size properties (pq_utils.py:383-408)
    @property
    def pq_cache_size(self):
        return sum(s.len * self.bs * self.num_key_value_heads * self.M * s.buf.element_size() for s in self._k + self._v)

    @property
    def outlier_store_size(self):
        """bytes of (dim, delta) records held for the tokens quantized so far (extension)."""
        per_tok = self.bs * self.num_key_value_heads * 3
        return sum(s.len for s in self._k) * per_tok * self.k_out + sum(s.len for s in self._v) * per_tok * self.v_out

    @property
    def codebook_size(self):
        if hasattr(self, '_cent'):
            return self._cent.numel() * self._cent.element_size()
        return (self.key_cent.numel() * self.key_cent.element_size()
                + self.value_cent.numel() * self.value_cent.element_size())

    @property
    def residual_cache_size(self):
        return sum(c.numel() * c.element_size() for c in self.key_residual_cache + self.value_residual_cache)

    @property
    def registry_size(self):
        return (sum(b.numel() * b.element_size() for b in self.registery.partial_out_buffers.values())
                + sum(b.numel() * b.element_size() for b in self.registery.partial_lse_buffers.values()))


@contextlib.contextmanager
def _quiet_gc():
    """No cyclic garbage collection while a stream captures.  Destroying a CUDA graph (cudaGraphExecDestroy) is not permitted
    during a capture and INVALIDATES it; a dropped decoder's graphs sit in reference cycles, so Python may free them at any
    allocation — e.g. inside the model forward being captured (seen as a 1-in-6 `cudaErrorStreamCaptureInvalidated` when several
    decoders were created in one process; torch 2.11's `torch.cuda.graph` no longer collects on entry).  Collect first, then hold."""
    gc.collect()
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


class _GraphStep:
    """Host side of a CUDA-graph decode step over a PQ cache: the cache POLICY of decoding() (retire a full window, count tokens,
    start the asynchronous flush) runs on the host around a graph replay; the captured launches read the window length from a
    device counter (million_attn_params.r_dev).  The graph is re-captured when the code stores changed (every Lt = 128 tokens,
    64 for the paged cache): their lengths and addresses are launch parameters.  Subclasses define `_run()` = what is captured
    (it must launch one attention call per layer through `_launch_layer` and end with `_advance()`)."""

    def __init__(self, cache, device, dtype):
        self.cache, self.device, self.dtype = cache, device, dtype
        self.r_dev = torch.zeros(1, dtype=torch.int32, device=device)
        self._r_host = 0
        self.graph, self.sig, self.captures = None, None, 0

    def _launch_layer(self, l, q, k, v, out):
        """Append-and-attend of layer l on the current stream with the device-resident window length; q/k/v/out are tensors
        whose addresses are baked into the capture."""
        import ctypes
        c = self.cache
        plan = c._plan(l, self.dtype, self.device)
        p = plan['p']
        p.q, p.out, p.k_new, p.v_new = q.data_ptr(), out.data_ptr(), k.data_ptr(), v.data_ptr()
        p.r, p.r_dev = 1, self.r_dev.data_ptr()
        c._fill_stores(p, l)
        st = plan['attn'](plan['ref'], ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        p.r_dev = None
        if st:
            L.check(st)

    def _advance(self):
        ops.counter_add(self.r_dev, 1)

    def _signature(self):
        c = self.cache
        sig = []
        for l in range(c.layer_num):
            p = c._plan(l, self.dtype, self.device)['p']
            c._fill_stores(p, l)
            sig.append((p.nk, p.k_codes, p.k_head_stride, p.v_codes, p.v_head_stride, p.v_page_ids, p.n_pages, p.k_out_idx, p.v_out_idx, p.k_res))
        return sig

    def _run(self):
        raise NotImplementedError

    def _capture(self, r):
        # warm-up on the current stream (first-use attributes, library workspaces), then capture; both write window row r from
        # the static inputs, which the replay writes again with the same values
        self.r_dev.fill_(r)
        self._run()
        self.r_dev.fill_(r)
        torch.cuda.current_stream(self.device).synchronize()
        g = torch.cuda.CUDAGraph()
        with _quiet_gc():
            with torch.cuda.graph(g):
                self._run()
        self.sig = self._signature()
        self.graph = g
        self.captures += 1
        self._r_host = r

    def _step(self):
        c = self.cache
        for l in range(c.layer_num):
            if c._window_full(l):
                c._retire_window(l)
        r = c.residualed_tokens[0]
        assert all(x == r for x in c.residualed_tokens), "a graph decode step advances all layers together"
        if self.graph is None or c._step_graph is not self or self._signature() != self.sig:
            c._step_graph = self
            self._capture(r)
        elif self._r_host != r:
            self.r_dev.fill_(r)
            self._r_host = r
        self.graph.replay()
        self._r_host = r + 1
        for l in range(c.layer_num):
            c.residualed_tokens[l] = r + 1
            c.seen_tokens[l] += 1
        if c.async_flush and c._window_full(0):
            for l in range(c.layer_num):
                c._start_async_flush(l, c._flush_len())


class DecodeStepGraph(_GraphStep):
    """The attention of every layer of one decode step as a CUDA graph over the caller's static buffers: each layer is ONE
    launch (window append fused), q/out (layers, bs, nh, 1, d), k/v (layers, bs, nh_k, 1, d)."""

    def __init__(self, cache, q, k, v, out):
        c = cache
        Ln = c.layer_num
        assert q.shape == (Ln, c.bs, c.nh, 1, c.d) and out.shape == q.shape and k.shape == (Ln, c.bs, c.num_key_value_heads, 1, c.d) and v.shape == k.shape
        for t in (q, k, v, out):
            assert t.is_cuda and t.is_contiguous() and t.dtype == c.scalar_t
        assert c._fused_ok(q[0], k[0], v[0]), "decode_step_graph needs nbits=8 and fp16/bf16 buffers in the cache dtype"
        super().__init__(c, q.device, q.dtype)
        self.q, self.k, self.v, self.out = q, k, v, out

    def _run(self):
        for l in range(self.cache.layer_num):
            self._launch_layer(l, self.q[l], self.k[l], self.v[l], self.out[l])
        self._advance()

    def step(self):
        self._step()
        return self.out
