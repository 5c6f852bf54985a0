"""Evidence that the asynchronous quantization overlaps decode: CUDA events on BOTH streams around the
flush of every layer (side stream) and around every layer's attention launch (main stream) during the decode step that fills the
window, all on one device clock.  Llama-3.1-8B shapes, batch 8, 32K prefill, DynamicPQCache (window 128 -> 128-token flush).
Prints, per layer, when its flush ran relative to the main stream's attention launches, and the step times of the flush step with
the flush synchronous / asynchronous."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from million_b200.pq_utils import DynamicPQCache, Singleton

L, NH, NHK, D, M, BS = 32, 32, 8, 128, 64, 8
T0 = 32768
torch.manual_seed(0)
ck, cv = torch.randn(M, 256, 2, device="cuda").half(), torch.randn(M, 256, 2, device="cuda").half()
k0 = torch.randn(BS, NHK, T0, D, device="cuda").half(); v0 = torch.randn(BS, NHK, T0, D, device="cuda").half()
q = torch.randn(L, BS, NH, 1, D, device="cuda").half(); k = torch.randn(L, BS, NHK, 1, D, device="cuda").half(); v = torch.randn(L, BS, NHK, 1, D, device="cuda").half()

def run(async_flush):
    Singleton.clear_instance()
    c = DynamicPQCache(bs=BS, nh=NH, num_key_value_heads=NHK, M=M, layer_num=L, d=D, scalar_t=torch.float16, async_flush=async_flush)
    c.set_cent(ck, cv)
    for l in range(L):
        c._encode_append(k0, v0, l)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    enc = {}
    orig = c._encode_kv
    def timed_encode(ks, vs, layer):                 # runs on whatever stream the flush uses
        a, b = ev(), ev()
        a.record(torch.cuda.current_stream()); orig(ks, vs, layer); b.record(torch.cuda.current_stream())
        enc[layer] = (a, b)
    steps = []
    for s in range(2 * 128 + 3):
        if s == 128 + 127:                           # the second window fill: everything is warm
            c._encode_kv = timed_encode
        t0 = ev(); t0.record()
        att = []
        for l in range(L):
            a = ev(); a.record()
            c.decoding(q[l], k[l], v[l], l)
            b = ev(); b.record()
            att.append((a, b))
        t1 = ev(); t1.record()
        steps.append((s, t0, t1, att))
        if s == 128 + 128:
            c._encode_kv = orig
    torch.cuda.synchronize()
    return steps, enc

for mode in (False, True):
    steps, enc = run(mode)
    ms = {s: t0.elapsed_time(t1) for s, t0, t1, _ in steps}
    fill, nxt = 128 + 127, 128 + 128
    print(f"async_flush={mode}: step that fills the window {ms[fill]:.3f} ms, the next step {ms[nxt]:.3f} ms, median step {sorted(ms.values())[len(ms) // 2]:.3f} ms")
    if not enc:
        continue
    ref = steps[fill][1]
    rows = []
    for l in sorted(enc):
        a, b = enc[l]
        s_, t0_, t1_, att = steps[fill] if mode else steps[nxt]
        la, lb = att[l]
        rows.append((l, ref.elapsed_time(a), ref.elapsed_time(b), ref.elapsed_time(la), ref.elapsed_time(lb)))
    print("  layer: flush encode [start, end] ms | that layer's attention launch [start, end] ms (same clock, t = 0 at the start of the step that fills the window)")
    for l, ea, eb, la, lb in rows[:4] + rows[-2:]:
        print(f"  {l:5d}: [{ea:7.3f}, {eb:7.3f}] | [{la:7.3f}, {lb:7.3f}]")
    if mode:
        s_, t0_, t1_, att = steps[fill]
        end_of_step = ref.elapsed_time(t1_)
        inside = sum(1 for l, ea, eb, la, lb in rows if eb <= end_of_step + 1e-3)
        over = sum(min(eb, end_of_step) - ea for l, ea, eb, la, lb in rows if ea < end_of_step)
        print(f"  {inside} of {len(rows)} flushes completed before the main stream finished the same step ({end_of_step:.3f} ms); "
              f"{over:.3f} ms of encode time ran concurrently with that step's attention launches")
