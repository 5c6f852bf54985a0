"""µs per launch of the decode-attention kernel inside a CUDA graph of 32 launches (Llama-3.1-8B shapes), for A/B runs of library
variants: MILLION_B200_LIB=variants/x.so python tools/graph_rate.py [ctx] [bs ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from million_b200 import ops, _lib
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
Mm = int(os.environ.get("RATE_M", "64"))          # RATE_M=32: the two-bit kernel (attn_fast_dm4.cu)
NH, NHK = int(os.environ.get("RATE_NH", "32")), int(os.environ.get("RATE_NHK", "8"))
KOUT = int(os.environ.get("RATE_KOUT", "0"))    # K-side outlier records per token (the bench's synthetic store)
bss = [int(x) for x in sys.argv[2:]] or [1, 8]
dev = torch.device("cuda", 0)
for bs in bss:
    for pdl in (False, True):
        g, layers, cents = bench.resident_graph(torch, ops, bs, NH, NHK, ctx - 128, 128, dev, pdl=pdl, m=Mm, k_out=KOUT)
        for _ in range(5): g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize(); e0.record()
            for _ in range(20): g.replay()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / (20 * 32) * 1e3)
        print(f"{os.path.basename(_lib.LIB_PATH):28s} M {Mm} kout {KOUT} heads {NH}/{NHK} ctx {ctx} bs {bs} pdl {int(pdl)}: {best:7.2f} us per launch")
        del g, layers, cents
        torch.cuda.empty_cache()
