"""CPU, world_size 2, gloo: host-side logic of the multi-GPU paths (range planning, the all-gather of partial states and
the log-sum-exp merge).  The per-rank attention is played by the ORACLE here (there is no CPU product path); on GPUs the
same code runs with the CUDA kernel (tests/test_gpu_parity.py::test_partial_and_lse_merge_equals_single_call)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from million_b200 import sharding
from oracle import pq_oracle as O


def test_range_planning():
    r8 = sharding.split_kv_ranges(131072 - 128, 8)
    assert r8[0] == (0, 16384) and r8[-1][1] == 130944 and max(e - a for a, e in r8) - min(e - a for a, e in r8) <= 64
    r = sharding.split_kv_ranges(1000, 3, page=64)
    assert r[0][0] == 0 and r[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    assert all(s % 64 == 0 for s, _ in r)
    assert sharding.split_kv_ranges(0, 4) == [(0, 0)] * 4
    kv, qh = sharding.kv_head_shard(32, 8, 4, 3)
    assert list(kv) == [6, 7] and list(qh) == list(range(24, 32))
    with pytest.raises(ValueError):
        sharding.kv_head_shard(32, 8, 3, 0)
    # the replicated window dealt out row-wise: a partition of [0, r) for every world size
    for r, world in ((128, 8), (128, 3), (17, 4), (1, 2), (0, 4)):
        rows = [sharding.split_window_rows(r, world, g) for g in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == r and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        assert max(e - a for a, e in rows) - min(e - a for a, e in rows) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_partial(q, kc, vc, kcent, vcent, kres, vres, r, partial=None, **kw):
    """(o_unnormalised | m | l) of one rank's slice, natural-log units — what the CUDA kernel writes with PARTIAL_ONLY."""
    s = O._scores(q.numpy(), kc.numpy(), kcent.numpy(), kres.numpy(), r)
    bs, nh, T = s.shape
    nk = kc.shape[2]
    G = nh // kc.shape[1]
    d = q.shape[-1]
    Vh = O.pq_decode(vc.numpy(), vcent.numpy().astype(np.float32)).astype(np.float32)
    out = np.zeros((bs * nh, d + 2), np.float32)
    for b in range(bs):
        for h in range(nh):
            V = np.concatenate([Vh[b, h // G], vres.numpy().astype(np.float32)[b, h // G, :r]], 0)
            if T == 0:
                out[b * nh + h, d] = -np.inf
                continue
            m = s[b, h].max()
            p = np.exp(s[b, h] - m)
            out[b * nh + h, :d] = p @ V
            out[b * nh + h, d] = m
            out[b * nh + h, d + 1] = p.sum()
    partial.copy_(torch.from_numpy(out))
    return partial


def _oracle_merge(parts, d, dtype):
    p = parts.numpy()
    o, m, l = p[..., :d], p[..., d], p[..., d + 1]
    mstar = np.where(l > 0, m, -np.inf).max(0)
    w = np.where(l > 0, np.exp(m - mstar[None]), 0.0)
    return torch.from_numpy(((w[..., None] * o).sum(0) / (w * l).sum(0)[:, None]).astype(np.float32))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        inp = O.make_inputs(bs=1, nh=8, nh_k=2, nk=700, seed=3)
        t = {k: torch.from_numpy(v) for k, v in inp.items()}
        # ---- split-KV
        a, e = sharding.split_kv_ranges(700, world)[rank]
        r_local = 50 if rank == world - 1 else 0
        out = sharding.splitkv_decode_attn(t["q"], t["kc"][:, :, a:e], t["vc"][:, :, a:e], t["kcent"], t["vcent"], t["kres"], t["vres"],
                                           r_local, attn_fn=_oracle_partial, merge_fn=_oracle_merge)
        ref = O.pq_decode_attn(inp["q"], inp["kc"], inp["vc"], inp["kcent"], inp["vcent"], inp["kres"], inp["vres"], 50)
        err_split = float(np.abs(out.numpy() - ref).max())
        # ---- KV-head sharding: disjoint output slices, no collective on the data path; gather only to compare
        kv, qh = sharding.kv_head_shard(8, 2, world, rank)
        sl = O.pq_decode_attn(inp["q"][:, qh.start:qh.stop], inp["kc"][:, kv.start:kv.stop], inp["vc"][:, kv.start:kv.stop], inp["kcent"],
                              inp["vcent"], inp["kres"][:, kv.start:kv.stop], inp["vres"][:, kv.start:kv.stop], 50)
        gathered = [torch.empty_like(torch.from_numpy(sl)) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(sl))
        err_head = float(np.abs(torch.cat(gathered, 1).numpy() - ref).max())
        ret[rank] = (err_split, err_head)
    finally:
        dist.destroy_process_group()


def test_two_rank_splitkv_and_head_sharding():
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for rank in range(world):
        err_split, err_head = ret[rank]
        assert err_split < 1e-5 and err_head < 1e-6
