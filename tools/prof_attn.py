"""Tiny driver for ncu: a few decode-attention launches on Llama-3.1-8B shapes (one layer), then CUDA-event timing."""
import argparse, os, sys
sys.path.insert(0, os.environ.get("MILLION_PKG_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # MILLION_PKG_ROOT: A/B against another tree (e.g. variants/r1)
import torch
from million_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--bs", type=int, default=8)
ap.add_argument("--ctx", type=int, default=32768)
ap.add_argument("--nh", type=int, default=32)
ap.add_argument("--nhk", type=int, default=8)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--impl", type=int, default=0)
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--M", type=int, default=64)
ap.add_argument("--kout", type=int, default=0, help="K-side outlier records per token")
a = ap.parse_args()
torch.manual_seed(0)
nk, r = a.ctx - 128, 128
kcent = torch.randn(a.M, 256, 128 // a.M, device="cuda").half(); vcent = torch.randn(a.M, 256, 128 // a.M, device="cuda").half()
L = []
for _ in range(a.layers):
    L.append((torch.randn(a.bs, a.nh, 1, 128, device="cuda").half(),
              torch.randint(0, 256, (a.bs, a.nhk, nk, a.M), dtype=torch.uint8, device="cuda"),
              torch.randint(0, 256, (a.bs, a.nhk, nk, a.M), dtype=torch.uint8, device="cuda"),
              torch.randn(a.bs, a.nhk, 128, 128, device="cuda").half(), torch.randn(a.bs, a.nhk, 128, 128, device="cuda").half()))
out = torch.empty(a.bs, a.nh, 1, 128, device="cuda", dtype=torch.float16)
KO = [(torch.randint(0, 128, (a.bs, a.nhk, nk, a.kout), dtype=torch.uint8, device="cuda"),
       torch.randn(a.bs, a.nhk, nk, a.kout, device="cuda").half()) for _ in range(a.layers)] if a.kout else [None] * a.layers
def run():
    for (q, kc, vc, kr, vr), ko in zip(L, KO):
        ops.pq_decode_attn(q, kc, vc, kcent, vcent, kr, vr, r, out=out, impl=a.impl, k_outliers=ko)
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters): run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (a.iters * a.layers)
alg = 2 * a.bs * a.nhk * nk * a.M + 3 * a.kout * a.bs * a.nhk * nk + 2 * a.bs * a.nhk * r * 128 * 2 + 2 * 64 * 256 * 2 * 2 + 2 * a.bs * a.nh * 128 * 2
print(f"kout={a.kout} M={a.M} bs={a.bs} ctx={a.ctx} nh={a.nh}/{a.nhk}: {us:.1f} us/launch, {alg/us/1e3:.0f} GB/s ({alg/us/1e3/6554.6*100:.1f}% of measured HBM peak)")
