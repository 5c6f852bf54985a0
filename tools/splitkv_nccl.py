"""Split-KV decode attention across ranks (NCCL): Llama-3.1-8B shapes, batch 1, 128K context (BASELINE config 4).
torchrun --nproc-per-node N tools/splitkv_nccl.py [--ctx 131072] — checks the merged result against a single-GPU call
(rank 0 holds the whole cache for the check) and times 32 layers per step on the device (max over ranks)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from million_b200 import ops, sharding

ap = argparse.ArgumentParser(); ap.add_argument("--ctx", type=int, default=131072); ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--steps", type=int, default=20); ap.add_argument("--graph", action="store_true"); ap.add_argument("--p2p", action="store_true"); ap.add_argument("--fused", action="store_true", help="with --p2p: exchange inside the attention kernel (one launch per layer)"); a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
NH, NHK, D, M, C, LT = 32, 8, 128, 64, 256, 128
nk, r = a.ctx - LT, LT
g = torch.Generator(device=dev); g.manual_seed(42)          # same seed on every rank -> identical full tensors
kcent = torch.randn(M, C, 2, device=dev, generator=g).half(); vcent = torch.randn(M, C, 2, device=dev, generator=g).half()
q = torch.randn(1, NH, 1, D, device=dev, generator=g).half()
kc = torch.randint(0, C, (1, NHK, nk, M), dtype=torch.uint8, device=dev, generator=g)
vc = torch.randint(0, C, (1, NHK, nk, M), dtype=torch.uint8, device=dev, generator=g)
kres = torch.randn(1, NHK, LT, D, device=dev, generator=g).half(); vres = torch.randn(1, NHK, LT, D, device=dev, generator=g).half()
s, e = sharding.split_kv_ranges(nk, world)[rank]
kcl, vcl = kc[:, :, s:e].contiguous(), vc[:, :, s:e].contiguous()
r_local = r if rank == world - 1 else 0
full = ops.pq_decode_attn(q, kc, vc, kcent, vcent, kres, vres, r)
peer = sharding.SplitKVPeerGroup(NH, D, torch.float16) if (a.p2p and world > 1) else None
if peer is not None:
    out = peer.decode_attn(q, kcl, vcl, kcent, vcent, kres, vres, r_local, fused=a.fused).clone()
elif world > 1:
    out = sharding.splitkv_decode_attn(q, kcl, vcl, kcent, vcent, kres, vres, r_local)
else:
    out = full
err = (out.float() - full.float()).abs().max().item()
# timing: `layers` calls per step (per-layer slices of the same size), device events, max over ranks
layers = [(kcl.clone(), vcl.clone()) for _ in range(min(a.layers, 8))]
outbuf = torch.empty(1, NH, 1, D, dtype=torch.float16, device=dev)
def step():
    for i in range(a.layers):
        k_, v_ = layers[i % len(layers)]
        if peer is not None: peer.decode_attn(q, k_, v_, kcent, vcent, kres, vres, r_local, out=outbuf, fused=a.fused)
        elif world > 1: sharding.splitkv_decode_attn(q, k_, v_, kcent, vcent, kres, vres, r_local)
        else: ops.pq_decode_attn(q, k_, v_, kcent, vcent, kres, vres, r)
for _ in range(3): step()
if a.graph and (world == 1 or peer is not None):   # NCCL collectives inside a captured graph hung on this pool; the P2P kernel is plain launches
    # capture the whole 32-layer step (kernels + NCCL all-gathers + merges) in one CUDA graph: no per-layer launch cost
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        step()
    torch.cuda.current_stream().wait_stream(st); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        step()
    eager_step = step
    step = gr.replay
    for _ in range(3): step()
if world > 1: dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps): step()
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if peer is not None:
    # the exchange alone: push + flag + wait + merge of a fixed partial state
    for _ in range(5): peer.merge(outbuf.view(NH, D))
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(200): peer.merge(outbuf.view(NH, D))
    e1.record(); torch.cuda.synchronize()
    tm = torch.tensor([e0.elapsed_time(e1) / 200 * 1e3], device=dev); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    # the local attention alone (partial only)
    e0.record()
    for _ in range(50):
        for k_, v_ in layers: ops.pq_decode_attn(q, k_, v_, kcent, vcent, kres, vres, r_local, partial=peer.partial)
    e1.record(); torch.cuda.synchronize()
    ta = torch.tensor([e0.elapsed_time(e1) / (50 * len(layers)) * 1e3], device=dev); dist.all_reduce(ta, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"  exchange+merge kernel alone: {tm.item():.1f} us per call (eager launches); local partial attention alone: {ta.item():.1f} us per layer")
if rank == 0:
    ms = t.item(); alg = 2 * NHK * nk * M + 2 * NHK * r * D * 2
    print(f"split-KV world={world} ctx={a.ctx} graph={a.graph} p2p={a.p2p} fused={a.fused} timed_out={peer.timed_out() if peer is not None else None}: max|merged - single| = {err:.2e}; {ms:.3f} ms per {a.layers}-layer token -> {1e3/ms:.1f} tok/s; "
          f"{alg * a.layers / (ms * 1e-3) / 1e9:.0f} GB/s aggregate algorithmic")
if world > 1: dist.destroy_process_group()
