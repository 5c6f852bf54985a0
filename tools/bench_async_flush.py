"""BASELINE config 3: Llama-3.1-8B shapes, PagedPQCache — paged prefill at 32K, then decode with the quantization of the
retiring window block either synchronous (reference behaviour, paged_pq_utils.py:357-360) or asynchronous on a side stream.
Reports per-step device time of the 32-layer attention path (CUDA events), mean and worst step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from million_b200.paged_pq_utils import PagedPQCache
from million_b200.pq_utils import DynamicPQCache, Singleton

L, NH, NHK, D, M = 32, 32, 8, 128, 64
T0 = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(0)
cent_k, cent_v = torch.randn(M, 256, 2, device="cuda").half(), torch.randn(M, 256, 2, device="cuda").half()
k0 = torch.randn(1, NHK, T0, D, device="cuda").half(); v0 = torch.randn(1, NHK, T0, D, device="cuda").half()
qs = torch.randn(max(steps, 140), L, 1, NH, 1, D, device="cuda").half()
ks = torch.randn(max(steps, 140), L, 1, NHK, 1, D, device="cuda").half(); vs = torch.randn(max(steps, 140), L, 1, NHK, 1, D, device="cuda").half()

def run(cls, async_flush, name):
    Singleton.clear_instance()
    cache = cls(bs=1, nh=NH, num_key_value_heads=NHK, M=M, layer_num=L, d=D, scalar_t=torch.float16, async_flush=async_flush)
    cache.set_cent(cent_k, cent_v)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for l in range(L):
        cache._encode_append(k0, v0, l)           # the quantize-and-store part of prefill (SDPA itself is library code)
    e1.record(); torch.cuda.synchronize()
    t_prefill = e0.elapsed_time(e1)
    dec = cache.decoding_with_pages if cls is PagedPQCache else cache.decoding
    outs = None
    for s in range(140):                               # warm-up: passes the first flush (lazy kernel loading, pool growth)
        for l in range(L):
            outs = dec(qs[s, l], ks[s, l], vs[s, l], l)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for s in range(steps):
        for l in range(L):
            outs = dec(qs[s, l], ks[s, l], vs[s, l], l)
        evs[s + 1].record()
    torch.cuda.synchronize()
    ts = torch.tensor([evs[i].elapsed_time(evs[i + 1]) for i in range(steps)])
    print(f"{name:34s} prefill-encode {T0} tok x {L} layers: {t_prefill:7.1f} ms ({T0 * L / t_prefill / 1e3:.2f} Mtok-layers/s) | "
          f"decode step (32 layers): mean {ts.mean():.3f} ms  p50 {ts.median():.3f}  max {ts.max():.3f} ms; checksum {outs.float().sum().item():.4f}")
    return ts

a = run(PagedPQCache, False, "PagedPQCache sync flush")
b = run(PagedPQCache, True, "PagedPQCache async flush")
c = run(DynamicPQCache, False, "DynamicPQCache sync flush")
d = run(DynamicPQCache, True, "DynamicPQCache async flush")
