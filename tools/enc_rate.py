"""Encoder throughput + bit-exactness against the exact CUDA-core encoder, for A/B of library variants.
MILLION_B200_LIB=variants/x.so python tools/enc_rate.py [M ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from million_b200 import ops, _lib as L
Ms = [int(x) for x in sys.argv[1:]] or [64, 32]
n = 32768
for m in Ms:
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    X = torch.randn(1, 8, n, 128, device="cuda", generator=g).half()
    cent = torch.randn(m, 256, 128 // m, device="cuda", generator=g).half().float().contiguous()
    codes = torch.empty(1, 8, n, m, dtype=torch.uint8, device="cuda")
    exact = ops.pq_encode(X, cent, impl=L.IMPL_GENERIC)
    ops.pq_encode_into(X, cent, codes)
    ok = torch.equal(codes, exact)
    for _ in range(3): ops.pq_encode_into(X, cent, codes)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): ops.pq_encode_into(X, cent, codes)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    print(f"{os.path.basename(L.LIB_PATH):24s} M={m}: {best*1e3:7.1f} us per (8 heads x 32768 tokens) = {8*n/best/1e3:7.1f} M head-vectors/s, {8*n/best/1e3/(2*8*32):.2f} Mtok/s (32 layers, K+V); bit-exact vs exact encoder: {ok}")
