"""Needs a -DMILLION_DEBUG build (see tools/attn_phases.py)."""
import ctypes, os, sys
sys.path.insert(0, "/root/repo")
import torch
from million_b200 import ops, _lib
bs, ctx = 8, 32768
nh, nhk = 32, 8
nk, r = ctx - 128, 128
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
t = buf.view(-1, 8).cpu().double()
used = (t[:, 0] > 0).sum().item()
tt = t[:used]
t0 = tt[:, 0].min()
print("CTAs with stamps:", used, "span us", (tt[:, :7].max() - t0) / 1e3)
print("end per CTA (us): min %.1f mean %.1f max %.1f" % (((tt[:,6]-t0).min())/1e3, ((tt[:,6]-t0).mean())/1e3, ((tt[:,6]-t0).max())/1e3))
d = tt[:, 1:7] - tt[:, 0:6]
for i, nm in enumerate(["prologue", "main", "partial", "window", "ticket", "merge"]):
    print(f"  {nm:9s} mean {d[:, i].mean() / 1e3:7.2f} us   max {d[:, i].max() / 1e3:7.2f} us")
