"""Per-CTA phase timing of attn_fast (debug hook million_debug_set_timing_buffer)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from million_b200 import ops, _lib
bs, ctx = int(sys.argv[1]), int(sys.argv[2])
nh, nhk = 32, 8
nk, r = ctx - 128, 128
torch.manual_seed(0)
kcent = torch.randn(64, 256, 2, device="cuda").half(); vcent = torch.randn(64, 256, 2, device="cuda").half()
q = torch.randn(bs, nh, 1, 128, device="cuda").half()
kc = torch.randint(0, 256, (bs, nhk, nk, 64), dtype=torch.uint8, device="cuda"); vc = torch.randint(0, 256, (bs, nhk, nk, 64), dtype=torch.uint8, device="cuda")
kr = torch.randn(bs, nhk, 128, 128, device="cuda").half(); vr = torch.randn(bs, nhk, 128, 128, device="cuda").half()
buf = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
for _ in range(3): ops.pq_decode_attn(q, kc, vc, kcent, vcent, kr, vr, r)
h = _lib.lib(); h.million_debug_set_timing_buffer.argtypes = [ctypes.c_void_p]
h.million_debug_set_timing_buffer(ctypes.c_void_p(buf.data_ptr()))
ops.pq_decode_attn(q, kc, vc, kcent, vcent, kr, vr, r); torch.cuda.synchronize()
h.million_debug_set_timing_buffer(None)
S = ops.default_splits(bs, nhk, nk); n = S * nhk * bs
t = buf[:n * 8].view(n, 8).cpu().double()
t0 = t[:, 0].min()
names = ["start", "prologue", "main", "partial", "window", "ticket", "merge"]
print(f"bs={bs} ctx={ctx} S={S} CTAs={n}; kernel span {(t[:, :7].max() - t0) / 1e3:.1f} us")
d = t[:, 1:7] - t[:, 0:6]
for i, nm in enumerate(names[1:]):
    print(f"  {nm:9s} mean {d[:, i].mean() / 1e3:7.2f} us   max {d[:, i].max() / 1e3:7.2f} us")
print(f"  start skew: max {(t[:, 0].max() - t0) / 1e3:.2f} us;  end: mean {(t[:, 6] - t0).mean() / 1e3:.1f} max {(t[:, 6] - t0).max() / 1e3:.1f} us")
