"""Main-loop cost of the decode kernel in clk per coded token and SM: slope of the launch time between two context lengths
(fixed per-CTA costs cancel).  python tools/loop_rate.py [bs] [mode ...]"""
import ctypes, os, sys
sys.path.insert(0, os.environ.get("MILLION_PKG_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # MILLION_PKG_ROOT: A/B against another tree (e.g. variants/r1)
import torch
from million_b200 import ops, _lib
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
modes = [int(x) for x in sys.argv[2:]] or [0]
nh, nhk, layers = 32, 8, 4
torch.manual_seed(0)
kcent = torch.randn(64, 256, 2, device="cuda").half(); vcent = torch.randn(64, 256, 2, device="cuda").half()
h = _lib.lib()
has_mode = hasattr(h, "million_debug_set_mode")
if has_mode: h.million_debug_set_mode.argtypes = [ctypes.c_int]
def timeit(ctx, mode):
    nk, r = ctx - 128, 128
    L = [(torch.randn(bs, nh, 1, 128, device="cuda").half(),
          torch.randint(0, 256, (bs, nhk, nk, 64), dtype=torch.uint8, device="cuda"), torch.randint(0, 256, (bs, nhk, nk, 64), dtype=torch.uint8, device="cuda"),
          torch.randn(bs, nhk, 128, 128, device="cuda").half(), torch.randn(bs, nhk, 128, 128, device="cuda").half()) for _ in range(layers)]
    out = torch.empty(bs, nh, 1, 128, device="cuda", dtype=torch.float16)
    if has_mode: h.million_debug_set_mode(mode)
    def run():
        for q, kc, vc, kr, vr in L: ops.pq_decode_attn(q, kc, vc, kcent, vcent, kr, vr, r, out=out)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8): run()
    e1.record(); torch.cuda.synchronize()
    if has_mode: h.million_debug_set_mode(0)
    return e0.elapsed_time(e1) * 1e3 / (8 * layers)
for mode in modes:
    a, b = timeit(32768 + 128, mode), timeit(65536 + 128, mode)
    per_sm = bs * nhk * 32768 / 148
    print(f"bs={bs} mode {mode}: 32K {a:.1f} us, 64K {b:.1f} us; loop {(b - a) * 1e-6 * 1.965e9 / per_sm:.2f} clk/token/SM ({(b - a):.1f} us per 32K), fixed {2 * a - b:.1f} us")
