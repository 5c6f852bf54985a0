"""Per-CTA phase timing of attn_fast (debug hook million_debug_set_timing_buffer): raw globaltimer stamps per piece.
Needs a -DMILLION_DEBUG build: MILLION_NVCC_EXTRA=-DMILLION_DEBUG python -m million_b200._build --out variants/debug.so; MILLION_B200_LIB=$PWD/variants/debug.so python tools/attn_phases.py 8 32768"""
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
buf = torch.zeros(4096 * 64, dtype=torch.int64, device="cuda")
for _ in range(3): ops.pq_decode_attn(q, kc, vc, kcent, vcent, kr, vr, r)
torch.cuda.synchronize()
h = _lib.lib(); h.million_debug_set_timing_buffer.argtypes = [ctypes.c_void_p]
h.million_debug_set_timing_buffer(ctypes.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.pq_decode_attn(q, kc, vc, kcent, vcent, kr, vr, r); e1.record(); torch.cuda.synchronize()
h.million_debug_set_timing_buffer(None)
t = buf.view(4096, 4, 16).cpu().double()
t = t[t[:, 0, 0] > 0]
t0 = t[:, 0, 0].min()
pieces = (t[:, :, 0] > 0).sum(1).double()
end = torch.where(t[:, :, 6] > 0, t[:, :, 6], torch.zeros(())).max(1).values
print(f"bs={bs} ctx={ctx}: {len(t)} CTAs, {pieces.mean():.2f} pieces per CTA; events {e0.elapsed_time(e1) * 1e3:.1f} us; first start -> last end {(end.max() - t0) / 1e3:.1f} us")
names = ["prologue", "main loop", "window", "combine", "ticket", "merge"]
for i, nm in enumerate(names):
    d = torch.where(t[:, :, 0] > 0, t[:, :, i + 1] - t[:, :, i], torch.zeros(())).sum(1)
    print(f"  {nm:10s} mean {d.mean() / 1e3:7.2f} us   max {d.max() / 1e3:7.2f} us   (per CTA, summed over its pieces)")
d = torch.where(t[:, :, 0] > 0, t[:, :, 7] - t[:, :, 2], torch.zeros(())).sum(1)
print(f"  (window: warp 0's own window work {d.mean() / 1e3:.2f} us mean, {d.max() / 1e3:.2f} max; the rest is waiting at the barrier for the other warps)")
m = t[:, :, 10].reshape(-1) > 0
if m.any():
    tt = t.reshape(-1, 16)[m]
    print(f"  merging pieces ({int(m.sum())}): total {((tt[:, 6] - tt[:, 5]).mean()) / 1e3:.2f} us = call+issue {(tt[:, 11] - tt[:, 5]).mean() / 1e3:.2f}, wait {(tt[:, 12] - tt[:, 11]).mean() / 1e3:.2f}, "
          f"m/l {(tt[:, 9] - tt[:, 12]).mean() / 1e3:.2f}, weights {(tt[:, 10] - tt[:, 9]).mean() / 1e3:.2f}, sum+store {(tt[:, 6] - tt[:, 10]).mean() / 1e3:.2f}")
print(f"  start skew max {(t[:, 0, 0].max() - t0) / 1e3:.2f} us; end mean {(end - t0).mean() / 1e3:.1f} min {(end - t0).min() / 1e3:.1f} max {(end - t0).max() / 1e3:.1f} us")
# where the tail skew comes from: CTAs by number of pieces (a run that crosses a group boundary pays a second prologue/epilogue)
ml = torch.where(t[:, :, 0] > 0, t[:, :, 2] - t[:, :, 1], torch.zeros(())).sum(1)
fx = (end - t[:, 0, 0]) - ml
for p in sorted(set(pieces.tolist())):
    sel = pieces == p
    print(f"  CTAs with {int(p)} piece(s): {int(sel.sum()):3d}; end mean {(end[sel] - t0).mean() / 1e3:6.1f} (min {(end[sel] - t0).min() / 1e3:6.1f}, max {(end[sel] - t0).max() / 1e3:6.1f}) us; "
          f"main loop mean {ml[sel].mean() / 1e3:6.1f} (min {ml[sel].min() / 1e3:6.1f}, max {ml[sel].max() / 1e3:6.1f}); everything else {fx[sel].mean() / 1e3:5.1f} us")
q10 = torch.quantile(end - t0, torch.tensor([0.1, 0.5, 0.9], dtype=torch.float64)) / 1e3
print(f"  end time quantiles 10/50/90 %: {q10[0]:.1f} / {q10[1]:.1f} / {q10[2]:.1f} us")
