"""Multi-GPU partitioning of the PQ decode-attention path (one process per GPU, torch.distributed).

The reference is single-GPU (scripts/modeldb/main_pq.py:74 "TODO: support multi-gpu"); the path shards in two ways:

  * by KV head — codes, windows, block tables are independent per (batch, kv-head): rank g owns kv-heads
    [g*nh_k/G, (g+1)*nh_k/G) and their query heads.  No collective on the data path.
  * by sequence (split-KV) for batch-1 long contexts — rank g owns a page-aligned token range, computes the
    un-normalised (o, m, l) state of its range with the same kernel (MILLION_ATTN_PARTIAL_ONLY), the states are
    all-gathered (bs*nh*(d+2) fp32 = 16.6 KB per rank for Llama-3.1-8B shapes) over NCCL/NVLink and merged with the
    log-sum-exp algebra of flash_decoding_reduce_kernel (scripts/modeldb/bindings/Kernel.cuh:1249-1269).
    The fp16 window lives on the last rank, or — balanced — is replicated and dealt out row-wise (split_window_rows).
"""
from typing import List, Tuple

import torch
import torch.distributed as dist


def kv_head_shard(nh: int, nh_k: int, world: int, rank: int) -> Tuple[range, range]:
    """(kv-head range, query-head range) owned by `rank`.  nh_k must be divisible by world."""
    if nh_k % world:
        raise ValueError(f"num_key_value_heads={nh_k} is not divisible by world size {world}")
    per = nh_k // world
    G = nh // nh_k
    return range(rank * per, (rank + 1) * per), range(rank * per * G, (rank + 1) * per * G)


def split_kv_ranges(nk: int, world: int, page: int = 64) -> List[Tuple[int, int]]:
    """Page-aligned token ranges [(start, end)] of the nk quantized tokens, one per rank, sizes differ by <= 1 page."""
    n_pages = (nk + page - 1) // page
    base, extra = divmod(n_pages, world)
    out, p = [], 0
    for g in range(world):
        q = base + (1 if g < extra else 0)
        out.append((min(p * page, nk), min((p + q) * page, nk)))
        p += q
    return out


def split_window_rows(r: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [start, end) of the r-row fp16 window that `rank` attends over when the window is replicated on every rank (each
    rank computes the new token's k/v anyway under tensor parallelism) instead of living on the last rank only: the ranks'
    work stays balanced, nobody waits for the one that also owns the window.  Every rank then appends the new token to ITS replica
    with ops.window_append: the append fused into the attention launch (k_new / v_new) only stores a row the call attends over."""
    return r * rank // world, r * (rank + 1) // world


def allgather_partials(partial: torch.Tensor, group=None) -> torch.Tensor:
    """partial (rows, d+2) fp32 -> (world, rows, d+2): one small all-gather (NCCL over NVLink on GPUs, gloo on CPU)."""
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(partial.shape), dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(out.view(-1), partial.contiguous().view(-1), group=group)
    return out


def splitkv_decode_attn(q, k_codes_local, v_codes_local, k_cent, v_cent, k_res, v_res, r_local, group=None, merge_fn=None,
                        attn_fn=None, **attn_kw):
    """Split-KV decode attention across the ranks of `group`.

    Every rank passes its LOCAL slice of the code caches; `r_local` is the window length on this rank (non-zero on one
    rank only).  Returns the full (bs, nh, 1, d) output on every rank."""
    from . import ops
    attn_fn = attn_fn or ops.pq_decode_attn
    merge_fn = merge_fn or ops.lse_merge
    bs, nh, d = q.shape[0], q.shape[1], q.shape[-1]
    partial = torch.empty(bs * nh, d + 2, dtype=torch.float32, device=q.device)
    attn_fn(q, k_codes_local, v_codes_local, k_cent, v_cent, k_res, v_res, r_local, partial=partial, **attn_kw)
    parts = allgather_partials(partial, group)
    return merge_fn(parts, d, q.dtype).view(bs, nh, 1, d)


class SplitKVPeerGroup:
    """Split-KV decode attention whose cross-GPU exchange is the library's own NVLink kernel (million_splitkv_push_merge)
    instead of NCCL: partial states are stored directly into the peers' symmetric buffers and merged in the same launch.
    Symmetric memory comes from torch.distributed._symmetric_memory (CUDA IPC / fabric handles); torch is plumbing only."""

    def __init__(self, rows, d, dtype, group=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.rows, self.d, self.dtype = rows, d, dtype
        nbytes = L.lib().million_splitkv_symmetric_bytes(self.world, rows, d)
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, self.group)
        ptrs = list(self.handle.buffer_ptrs)
        self._peer_array = (ctypes.c_void_p * self.world)(*ptrs)
        from . import ops
        self.state = ops.splitkv_state(ptrs, self.rank, rows, self.buf.device)
        self.partial = torch.empty(rows, d + 2, dtype=torch.float32, device=self.buf.device)
        torch.cuda.synchronize()
        dist.barrier(self.group)

    def merge(self, out=None, pdl=False):
        """Push self.partial to every rank and merge everybody's state; returns (rows, d)."""
        import ctypes
        from . import _lib as L
        from .ops import _DT
        if out is None:
            out = torch.empty(self.rows, self.d, dtype=self.dtype, device=self.buf.device)
        st = ctypes.c_void_p(torch.cuda.current_stream(self.buf.device).cuda_stream)
        L.check(L.lib().million_splitkv_push_merge(ctypes.c_void_p(self.partial.data_ptr()), ctypes.cast(self._peer_array, ctypes.c_void_p),
                                                   self.rank, self.world, self.rows, self.d, ctypes.c_void_p(out.data_ptr()), _DT[self.dtype],
                                                   ctypes.c_void_p(self.state.data_ptr()), L.ATTN_PDL if pdl else 0, st))
        return out

    def decode_attn(self, q, k_codes_local, v_codes_local, k_cent, v_cent, k_res, v_res, r_local, out=None, fused=False, pdl=False, **attn_kw):
        """fused=True: ONE launch per layer — the last CTA of every (b, kv-head) group pushes the group's rows to the peers, waits
        for theirs and writes the merged rows (MILLION_ATTN_FUSED_SPLITKV; M=64, nh/nh_k = 4, row-major V); otherwise attention +
        million_splitkv_push_merge.  pdl=True launches with programmatic dependent launch (MILLION_ATTN_PDL)."""
        from . import ops
        bs, nh, d = q.shape[0], q.shape[1], q.shape[-1]
        if fused:
            assert bs * nh == self.rows
            return ops.pq_decode_attn(q, k_codes_local, v_codes_local, k_cent, v_cent, k_res, v_res, r_local, out=out,
                                      p2p=self.state, pdl=pdl, **attn_kw)
        ops.pq_decode_attn(q, k_codes_local, v_codes_local, k_cent, v_cent, k_res, v_res, r_local, partial=self.partial, pdl=pdl, **attn_kw)
        o = self.merge(None if out is None else out.view(bs * nh, d), pdl=pdl)
        return o.view(bs, nh, 1, d)

    def set_timeout(self, microseconds):
        """How long a wait for the peers may last before the kernel gives up (NaN rows + error word); 0 = the default (~0.8 s)."""
        import ctypes
        from . import _lib as L
        L.check(L.lib().million_splitkv_set_timeout(ctypes.c_void_p(self.state.data_ptr()), int(microseconds),
                                                    ctypes.c_void_p(torch.cuda.current_stream(self.buf.device).cuda_stream)))

    def timed_out(self):
        return bool(self.state.view(torch.int32)[2].item())

    def check(self):
        """Raise if any exchange since the last check gave up waiting for a peer (synchronises the device: call it at a
        convenient point — e.g. once per generated token — not per layer).  The affected output rows are NaN as well."""
        if self.timed_out():
            self.state.view(torch.int32)[2] = 0
            raise RuntimeError(f"split-KV exchange on rank {self.rank}: a peer did not publish its state in time (dead or stalled rank)")
