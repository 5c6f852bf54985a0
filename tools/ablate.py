"""Ablation timing of attn_fast phases (temporary debug hook million_debug_set_mode).
Needs a -DMILLION_DEBUG build: MILLION_NVCC_EXTRA=-DMILLION_DEBUG python -m million_b200._build --out variants/debug.so; MILLION_B200_LIB=variants/debug.so python tools/ablate.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from million_b200 import ops, _lib
bs, ctx, layers = int(sys.argv[1]), int(sys.argv[2]), 4
nh, nhk = 32, 8
nk, r = ctx - 128, 128
torch.manual_seed(0)
kcent = torch.randn(64, 256, 2, device="cuda").half(); vcent = torch.randn(64, 256, 2, device="cuda").half()
L = [(torch.randn(bs, nh, 1, 128, device="cuda").half(),
      torch.randint(0, 256, (bs, nhk, nk, 64), dtype=torch.uint8, device="cuda"), torch.randint(0, 256, (bs, nhk, nk, 64), dtype=torch.uint8, device="cuda"),
      torch.randn(bs, nhk, 128, 128, device="cuda").half(), torch.randn(bs, nhk, 128, 128, device="cuda").half()) for _ in range(layers)]
out = torch.empty(bs, nh, 1, 128, device="cuda", dtype=torch.float16)
h = _lib.lib(); h.million_debug_set_mode.argtypes = [ctypes.c_int]
def run():
    for q, kc, vc, kr, vr in L: ops.pq_decode_attn(q, kc, vc, kcent, vcent, kr, vr, r, out=out)
for mode in [0, 1, 2, 3] + [int(x) for x in sys.argv[3:]]:
    h.million_debug_set_mode(mode)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    print(f"mode {mode}: {e0.elapsed_time(e1) * 1e3 / (5 * layers):.1f} us/launch")
h.million_debug_set_mode(0)
