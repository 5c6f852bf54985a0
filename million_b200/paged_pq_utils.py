"""B200-native counterpart of the reference's paged caches:

  * `PageManager` — page pool + free list, the format of the (dead) scripts/utils/dynamic_paged_pq_utils.py:10-321:
    pool (pages, M, page_size) uint8 holding V codes TRANSPOSED per page, deterministic allocation.
  * `PagedPQCache(DynamicPQCache)` — constructor and methods of the live scripts/utils/paged_pq_utils.py:10-1156
    (window of `extended_residual_size`=128 tokens, flush the oldest `page_size`=64 and shift), but V codes really
    live in pages addressed by a block table (bs, nh_k, n_pages) that the CUDA kernel consumes directly — the
    reference asks for such a kernel (`flash_decoding_paged_v_*`, paged_pq_utils.py:547, 621-635), never has one,
    and falls back to de-quantize + SDPA every step.

The flush of the oldest page is launched on a side stream right after the step that filled the window and is
joined by an event at the next step ("asynchronous quantization overlapped with decode").
"""
import logging
from typing import Dict, Set

import torch

from . import _lib as L
from . import ops
from .pq_utils import DynamicPQCache, KernelRegistry, l2Ns, sa_decode_4d

logger = logging.getLogger(__name__)


class PageManager:
    """dynamic_paged_pq_utils.py:10-321.  Pages are handed out lowest-id-first from a fresh pool, so prefill
    allocates chunk-major, then b, then h (dynamic_paged_pq_utils.py:776-811).  The pool grows by doubling
    (re-allocation + copy, :104-121) unless `max_pages` caps it."""

    def __init__(self, page_size: int = 64, initial_pages: int = 100, max_pages: int = None, M: int = 64, device='cuda'):
        self.page_size, self.M, self.device = page_size, M, torch.device(device)
        self.initial_pages, self.max_pages = initial_pages, max_pages
        self.preallocated_size = min(initial_pages * 2, max_pages) if max_pages is not None else max(initial_pages * 2, 1)
        self.current_active_pages = min(initial_pages, self.preallocated_size)
        self.page_pool = torch.zeros((self.preallocated_size, M, page_size), dtype=torch.uint8, device=self.device)
        self._free = list(range(self.current_active_pages - 1, -1, -1))   # stack, lowest id on top
        self._allocated: Set[int] = set()         # allocated page ids; `allocated_pages` is the reference's dict view of it
        self.total_expansions = 0
        self.page_reuse_count = 0
        self.total_allocations = 0
        self._ever_used: Set[int] = set()

    @property
    def free_pages(self) -> Set[int]:
        return set(self._free)

    @property
    def allocated_pages(self) -> Dict[int, Dict]:
        """The reference's bookkeeping dict (dynamic_paged_pq_utils.py:52-63), built on demand: a page that is allocated was
        allocated exactly once since it was last freed."""
        return {pid: {'allocation_count': 1} for pid in self._allocated}

    def _expand_page_pool(self, additional_pages: int = None, force_expansion: bool = False):
        if additional_pages is None:
            additional_pages = max(self.current_active_pages, 100) if force_expansion else max(self.current_active_pages // 2, 50)
        if self.max_pages is not None:
            room = self.max_pages - self.current_active_pages
            if room <= 0:
                raise RuntimeError(f"Cannot expand page pool: reached max_pages limit {self.max_pages}")
            additional_pages = min(additional_pages, room)
        new_active = self.current_active_pages + additional_pages
        if new_active > self.preallocated_size:
            new_size = max(new_active, self.preallocated_size * 2)
            if self.max_pages is not None:
                new_size = min(new_size, self.max_pages)
            new_pool = torch.zeros((new_size, self.M, self.page_size), dtype=torch.uint8, device=self.device)
            new_pool[:self.current_active_pages] = self.page_pool[:self.current_active_pages]
            self.page_pool, self.preallocated_size = new_pool, new_size
        self._free = list(range(new_active - 1, self.current_active_pages - 1, -1)) + self._free
        self._free.sort(reverse=True)
        self.current_active_pages = new_active
        self.total_expansions += 1

    def allocate_page(self) -> int:
        if not self._free:
            try:
                self._expand_page_pool()
            except RuntimeError as e:
                raise RuntimeError(f"No free pages available and cannot expand: {e}")
        pid = self._free.pop()
        if pid in self._ever_used:
            self.page_reuse_count += 1
        self._ever_used.add(pid)
        self._allocated.add(pid)
        self.total_allocations += 1
        return pid

    allocate_reused_page = allocate_page

    def allocate_pages(self, n: int) -> list:
        if n <= 0:
            return []
        if len(self._free) < n:
            try:
                self._expand_page_pool(additional_pages=n - len(self._free), force_expansion=True)
            except RuntimeError as e:
                raise RuntimeError(f"Cannot bulk-allocate pages: {e}")
        if len(self._free) < n:
            raise RuntimeError(f"Cannot bulk-allocate {n} pages: max_pages={self.max_pages}")
        # bulk form of n allocate_page() calls (same order: lowest id first), without n trips through the interpreter
        ids = self._free[-n:][::-1]
        del self._free[-n:]
        self.page_reuse_count += len(self._ever_used.intersection(ids))
        self._ever_used.update(ids)
        self._allocated.update(ids)
        self.total_allocations += n
        return ids

    def free_page(self, page_id: int):
        if page_id not in self._allocated:
            logger.warning(f"Attempting to free unallocated page {page_id}")
            return
        self._allocated.discard(page_id)
        self._free.append(page_id)
        self._free.sort(reverse=True)

    def _is_valid_page_id(self, page_id: int) -> bool:
        return isinstance(page_id, int) and 0 <= page_id < self.current_active_pages

    def get_page(self, page_id: int) -> torch.Tensor:
        if not self._is_valid_page_id(page_id):
            raise ValueError(f"Invalid page_id {page_id}: must be within [0, {self.current_active_pages})")
        if page_id not in self._allocated:
            raise ValueError(f"Page {page_id} is not allocated")
        return self.page_pool[page_id]

    safe_page_access = get_page

    def get_stats(self) -> Dict:
        return {
            'initial_pages': self.initial_pages, 'current_active_pages': self.current_active_pages,
            'preallocated_size': self.preallocated_size, 'max_pages': self.max_pages,
            'allocated_pages': len(self._allocated), 'free_pages': len(self._free),
            'utilization': len(self._allocated) / self.current_active_pages if self.current_active_pages else 0,
            'memory_usage_mb': self.page_pool.numel() / (1024 * 1024),
            'page_reuse_count': self.page_reuse_count, 'total_allocations': self.total_allocations,
            'total_expansions': self.total_expansions,
        }


class PagedPQCache(DynamicPQCache):
    """paged_pq_utils.py:10-1156: same constructor and public methods; V codes in pages + block table."""

    def __init__(self, *, bs, nh, num_key_value_heads, M, layer_num, dtype=torch.uint8, nbits=8, d=128,
                 scalar_t=torch.float32, page_size=64, extended_residual_size=128, max_pages_per_layer=None,
                 async_flush=True, device='cuda', outliers=(0, 0)):
        self.page_size = page_size
        self.extended_residual_size = extended_residual_size
        self.max_pages_per_layer = max_pages_per_layer
        self._stats = {'flushes': 0, 'paged_kernel_calls': 0, 'prefill_tokens': 0}
        super().__init__(bs=bs, nh=nh, num_key_value_heads=num_key_value_heads, M=M, layer_num=layer_num, dtype=dtype,
                         nbits=nbits, d=d, scalar_t=scalar_t, async_flush=async_flush, device=device, outliers=outliers)

    def init_cache(self):
        """paged_pq_utils.py:68-128."""
        self.max_residual_length = self.extended_residual_size
        super().init_cache()
        self.registery = KernelRegistry(M=self.M, d=self.d, nbits=self.nbits, nh=self.nh, scalar_t=self.scalar_t, bs=self.bs)
        self.page_managers = [PageManager(self.page_size, initial_pages=64, max_pages=self.max_pages_per_layer, M=self.M,
                                          device=self.device) for _ in range(self.layer_num)]
        # block tables: host list-of-lists [layer][b][h] -> page ids, mirrored on the device as (bs, nh_k, cap) int64
        self._grow_log = [[] for _ in range(self.layer_num)]    # page ids of every allocation, in allocation order (host mirror)
        self._table = [torch.zeros((self.bs, self.num_key_value_heads, 0), dtype=torch.int64, device=self.device)
                       for _ in range(self.layer_num)]
        self._n_pages = [0 for _ in range(self.layer_num)]
        self._v_tokens = [0 for _ in range(self.layer_num)]

    @property
    def value_page_ids(self):
        """[layer][b][h] -> page ids (dynamic_paged_pq_utils.py:453-456), materialised on demand: filling these Python lists
        eagerly cost ~25 ms of host time per 32K-token prefill (4096 appends per layer)."""
        out = []
        for l in range(self.layer_num):
            per = [[[] for _ in range(self.num_key_value_heads)] for _ in range(self.bs)]
            for ids in self._grow_log[l]:
                it = iter(ids)
                for _ in range(len(ids) // (self.bs * self.num_key_value_heads)):
                    for b in range(self.bs):
                        for h in range(self.num_key_value_heads):
                            per[b][h].append(next(it))
            out.append(per)
        return out

    # value_cache keeps the live class's (bs, nh_k, M, T) face, materialised from the pages on demand
    @property
    def value_cache(self):
        out = []
        for l in range(self.layer_num):
            T, n = self._v_tokens[l], self._n_pages[l]
            if n == 0:
                out.append(torch.zeros((self.bs, self.num_key_value_heads, self.M, 0), dtype=self.dtype, device=self.device))
                continue
            pages = self.page_managers[l].page_pool[self._table[l][:, :, :n]]       # (bs, nh_k, n, M, ps)
            out.append(pages.permute(0, 1, 3, 2, 4).reshape(self.bs, self.num_key_value_heads, self.M, n * self.page_size)[..., :T])
        return out

    def _grow_table(self, layer_idx, n_new_chunks):
        """Allocate pages for n_new_chunks more chunks: chunk-major, then b, then h (dynamic_paged_pq_utils.py:776-811)."""
        pm = self.page_managers[layer_idx]
        ids = pm.allocate_pages(n_new_chunks * self.bs * self.num_key_value_heads)
        if ids[-1] - ids[0] == len(ids) - 1:      # the usual case: a consecutive run -> build the table slice on the device, no H2D copy
            new = torch.arange(ids[0], ids[0] + len(ids), dtype=torch.int64, device=self.device)
        else:
            new = torch.tensor(ids, dtype=torch.int64)
        new = new.view(n_new_chunks, self.bs, self.num_key_value_heads).permute(1, 2, 0)
        self._grow_log[layer_idx].append(ids)     # value_page_ids (the reference's list-of-lists view) is built from this on demand
        n_old = self._n_pages[layer_idx]
        tab = self._table[layer_idx]
        if n_old + n_new_chunks > tab.shape[2]:
            cap = max(2 * tab.shape[2], n_old + n_new_chunks, 64)
            grown = torch.zeros((self.bs, self.num_key_value_heads, cap), dtype=torch.int64, device=self.device)
            grown[:, :, :n_old] = tab[:, :, :n_old]
            self._table[layer_idx] = tab = grown
        tab[:, :, n_old:n_old + n_new_chunks] = new.to(self.device)
        self._n_pages[layer_idx] = n_old + n_new_chunks

    def _reserve_k(self, layer_idx, n):
        ks = self._k[layer_idx]
        ks.reserve(ks.len + n)
        if self._ko[layer_idx] is not None:
            self._ko[layer_idx].reserve(ks.cap, ks.len)
        if self._vo[layer_idx] is not None:       # V records are indexed by token like K's, whatever page the codes live in
            self._vo[layer_idx].reserve(ks.cap, self._v_tokens[layer_idx])

    def _encode_k(self, key_states, layer_idx):
        ks, ko = self._k[layer_idx], self._ko[layer_idx]
        if ko is not None:
            key_states = ops.outlier_split_into(key_states, self._key_cent_f32, ko.idx, ko.val, t0=ks.len)
        ops.pq_encode_into(key_states, self._key_cent_f32, ks.buf, t0=ks.len)

    def _encode_v(self, value_states, layer_idx):
        vo = self._vo[layer_idx]
        if vo is not None:
            value_states = ops.outlier_split_into(value_states, self._value_cent_f32, vo.idx, vo.val, t0=self._v_tokens[layer_idx])
        ops.pq_encode_paged(value_states, self._value_cent_f32, self.page_managers[layer_idx].page_pool,
                            self._table[layer_idx], t0=self._v_tokens[layer_idx])

    def _reconstruct_v(self, layer_idx, n_last=None, dtype=None):
        T = self._v_tokens[layer_idx]
        n = T if n_last is None else n_last
        vt = self.value_cache[layer_idx][..., T - n:].transpose(2, 3)
        x = sa_decode_4d(vt, self.value_cent if dtype is None else self.value_cent.to(dtype))
        if self._vo[layer_idx] is not None and n:
            ops.outlier_apply(x, self._vo[layer_idx].idx, self._vo[layer_idx].val, t0=T - n)
        return x

    def _encode_append(self, key_states, value_states, layer_idx, count_seen=True):
        """K codes row-major in place; V codes into pages through the block table.  Requires the V token count to
        be page aligned before the append (true for prefill-from-empty and for page-sized flushes)."""
        self._complete_pending_flush(layer_idx)
        n = key_states.size(2)
        ks = self._k[layer_idx]
        self._reserve_k(layer_idx, n)
        self._encode_k(key_states, layer_idx)
        self._encode_v_pages(value_states, layer_idx, n)
        ks.len += n
        self._v_tokens[layer_idx] += n
        if count_seen:
            self.seen_tokens[layer_idx] += n
        return ks.view()[:, :, ks.len - n:], n

    def _reserve_pages(self, layer_idx, n):
        t0 = self._v_tokens[layer_idx]
        need = (t0 + n + self.page_size - 1) // self.page_size - self._n_pages[layer_idx]
        if need > 0:
            self._grow_table(layer_idx, need)

    def _encode_v_pages(self, value_states, layer_idx, n):
        self._reserve_pages(layer_idx, n)
        self._encode_v(value_states, layer_idx)

    def cat_codes(self, key_codes, value_codes, layer_idx):
        raise NotImplementedError("PagedPQCache stores V in pages; use prefill/update/decoding_with_pages")

    def update(self, key_states, value_states, layer_idx, distort_recent=False):
        """paged_pq_utils.py:322-339 -> parent update (torch-PQ path) on the paged stores."""
        past = self._k[layer_idx].len
        if distort_recent:
            self._encode_append(key_states, value_states, layer_idx)
            return self._reconstruct(layer_idx, 'k'), self._reconstruct_v(layer_idx)
        if past > 0:
            pk = self._reconstruct(layer_idx, 'k')
            pv = self._reconstruct_v(layer_idx)
        self._encode_append(key_states, value_states, layer_idx)
        if past > 0:
            key_states = torch.cat([pk.to(key_states.dtype), key_states], dim=2)
            value_states = torch.cat([pv.to(value_states.dtype), value_states], dim=2)
        return key_states, value_states

    def prefill(self, query_states, key_states, value_states, layer_idx, distort_recent=False):
        """paged_pq_utils.py:216-320: encode everything (K row-major, V into pages), then causal SDPA.
        `distort_recent` is tested for truthiness here as in the reference (:290)."""
        n = key_states.size(2)
        self._stats['prefill_tokens'] += n
        self._encode_append(key_states, value_states, layer_idx)
        if distort_recent is True:
            key_states = self._reconstruct(layer_idx, 'k', n, key_states.dtype)
            value_states = self._reconstruct_v(layer_idx, n, value_states.dtype)
        return self._prefill_attention(query_states, key_states, value_states)

    # ---- flush (paged_pq_utils.py:130-210)
    def flush_to_pages(self, layer_idx: int):
        """Quantize the oldest page_size window tokens into a fresh page per (b, h) and shift the window."""
        if self._complete_pending_flush(layer_idx):
            return
        if self.residualed_tokens[layer_idx] < self.page_size:
            return
        self._flush_window(layer_idx, self.page_size)
        self._after_flush_shift(layer_idx)

    def _window_after_flush(self, layer_idx):
        self._after_flush_shift(layer_idx)

    def _after_flush_shift(self, layer_idx):
        ps = self.page_size
        rem = self.residualed_tokens[layer_idx] - ps
        if rem > 0:
            ops.window_shift(self.key_residual_cache[layer_idx], self.value_residual_cache[layer_idx], ps, rem)
        self.residualed_tokens[layer_idx] = rem
        self._stats['flushes'] += 1

    def _start_async_flush(self, layer_idx, n):
        self._reserve_k(layer_idx, n)
        self._reserve_pages(layer_idx, n)
        main = torch.cuda.current_stream(self.device)
        side = self._side_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._encode_k(self.key_residual_cache[layer_idx][:, :, :n], layer_idx)
            self._encode_v(self.value_residual_cache[layer_idx][:, :, :n], layer_idx)
            ev = torch.cuda.Event()
            ev.record(side)
        self._pending[layer_idx] = (ev, n)

    def _finish_async_flush(self, layer_idx):
        ev, n = self._pending[layer_idx]
        torch.cuda.current_stream(self.device).wait_event(ev)
        self._k[layer_idx].len += n
        self._v_tokens[layer_idx] += n
        self._pending[layer_idx] = None

    # ---- decode (paged_pq_utils.py:341-397)
    def _window_full(self, layer_idx):
        return self.residualed_tokens[layer_idx] >= self.extended_residual_size

    def _flush_len(self):
        return self.page_size

    def _retire_window(self, layer_idx):
        self.flush_to_pages(layer_idx)

    def decoding_with_pages(self, query_states, key_states, value_states, layer_idx, out=None):
        if self._window_full(layer_idx):
            self._retire_window(layer_idx)
        r = self.residualed_tokens[layer_idx]
        n = key_states.size(2)
        self._stats['paged_kernel_calls'] += 1
        if self._fused_ok(query_states, key_states, value_states):
            out = self._decode_fast(query_states, key_states, value_states, layer_idx, r, out)    # one C call, pre-filled params
        else:
            ops.window_append(self.key_residual_cache[layer_idx], self.value_residual_cache[layer_idx], key_states, value_states, r)
            self.residualed_tokens[layer_idx] += n
            self.seen_tokens[layer_idx] += n          # each token counted once (the reference double-counts, Appendix B.5)
            res = self._call_paged_kernel(query_states, layer_idx, r + n)
            out = res if out is None else out.copy_(res)
        if self.async_flush and self._window_full(layer_idx):
            self._start_async_flush(layer_idx, self.page_size)
        return out

    decoding = decoding_with_pages

    def _fill_v(self, p, layer_idx):
        tab = self._table[layer_idx]
        p.v_layout, p.v_codes = L.V_PAGED, self.page_managers[layer_idx].page_pool.data_ptr()
        p.v_page_ids, p.n_pages, p.page_size = tab.data_ptr(), tab.shape[2], self.page_size

    def _call_paged_kernel(self, query_states, layer_idx, residual_length):
        """paged_pq_utils.py:399-681 — the 13-argument paged kernel call, for every (b, h) (the reference builds
        its pool from batch 0 / head 0 only, Appendix B.3)."""
        if self.k_out or self.v_out:
            # the reference's 13-argument paged call (paged_pq_utils.py:621-635) has no place for the side store
            ks = self._k[layer_idx]
            kc, vc = self._attn_cents(query_states.dtype)
            return ops.pq_decode_attn(query_states, ks.view(), self.page_managers[layer_idx].page_pool, kc, vc,
                                      self.key_residual_cache[layer_idx], self.value_residual_cache[layer_idx], residual_length,
                                      v_layout=L.V_PAGED, v_page_ids=self._table[layer_idx], page_size=self.page_size,
                                      k_outliers=self._ko[layer_idx].view(ks.len) if self.k_out else None,
                                      v_outliers=self._vo[layer_idx].view(ks.len) if self.v_out else None)
        Ns = l2Ns(self.seen_tokens[layer_idx])
        name = f"flash_decoding_paged_v_{'bf16' if query_states.dtype == torch.bfloat16 else 'f16'}u8_Ns{Ns}Lt{self.extended_residual_size}d{self.d}M{self.M}C256"
        from . import bindings
        kernel = getattr(bindings, name)
        kc, vc = self._attn_cents(query_states.dtype)
        n_pages = self._n_pages[layer_idx]
        ks = self._k[layer_idx]
        # pages may hold a partial tail: the kernel is told nk through key_codes, the table through n_pages
        return kernel(query_states, ks.view(), kc, self.key_residual_cache[layer_idx],
                      self._table[layer_idx],   # full-capacity table: its row length is the stride, entries >= n_pages are unused
                      self.page_managers[layer_idx].page_pool, vc, self.value_residual_cache[layer_idx],
                      residual_length, n_pages, self.page_size, None, None)

    # ---- stats API (paged_pq_utils.py:898-1078)
    def get_cache_stats(self) -> Dict:
        pages = sum(len(pm._allocated) for pm in self.page_managers)
        return {
            'layer_num': self.layer_num, 'page_size': self.page_size, 'extended_residual_size': self.extended_residual_size,
            'seen_tokens': list(self.seen_tokens), 'residualed_tokens': list(self.residualed_tokens),
            'quantized_tokens': [s.len for s in self._k], 'total_pages_allocated': pages,
            'pages_per_layer': [len(pm._allocated) for pm in self.page_managers],
            'pq_cache_bytes': self.pq_cache_size, 'residual_cache_bytes': self.residual_cache_size,
        }

    def get_performance_stats(self) -> Dict:
        return dict(self._stats)

    def get_memory_usage_summary(self) -> Dict:
        pool = sum(pm.page_pool.numel() for pm in self.page_managers)
        kbytes = sum(s.buf.numel() for s in self._k)
        return {'page_pool_mb': pool / 2**20, 'key_codes_mb': kbytes / 2**20, 'residual_mb': self.residual_cache_size / 2**20,
                'total_mb': (pool + kbytes + self.residual_cache_size) / 2**20}

    def print_performance_summary(self):
        print(f"PagedPQCache: {self._stats}")

    def show_current_status(self):
        print(f"PagedPQCache status: {self.get_cache_stats()}")

    @property
    def pq_cache_size(self):
        per = self.bs * self.num_key_value_heads * self.M
        return sum(s.len * per for s in self._k) + sum(t * per for t in self._v_tokens)

    def cleanup(self):
        """paged_pq_utils.py:1082-1114."""
        self.init_cache()     # waits for in-flight flushes and drops every cached pointer (DynamicPQCache.init_cache)
