"""Where the decode kernel's error comes from on a PEAKED score distribution (q scaled x12: logits with sigma ~ 12, a handful of
tokens carry the softmax mass) — the case tests/test_gpu_parity.py::test_attn_peaked_distribution checks.  Prints max |err| and
the north_star violation count (atol 2e-3, rtol 1e-2) against the fp32 oracle for:
  G = 4 (fp16 LUT entries, packed-half PV sums), G = 1 (fp32 LUT entries, packed-half PV sums), the all-shapes fp32 kernel,
for fp16 and bf16 I/O.  Run with libraries built with -DMILLION_PV_FLUSH_TILES=1 to see the effect of flushing the half sums
twice as often (MILLION_B200_LIB=variants/flush1.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from million_b200 import ops, _lib as L
from oracle import pq_oracle as O
print("library:", L.LIB_PATH)
for qscale in (1.0, 4.0, 12.0):
    for dtype in (torch.float16, torch.bfloat16):
        for nh, nh_k in ((8, 2), (8, 8)):
            worst = []
            for seed in (11, 12, 13):
                inp = O.make_inputs(bs=1, nh=nh, nh_k=nh_k, nk=3000, d=128, M=64, C=256, Lt=128, seed=seed)
                t = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
                t["q"] = (t["q"].float() * qscale)
                for k in ("q", "kcent", "vcent", "kres", "vres"):
                    t[k] = t[k].to(dtype)
                f = lambda k: t[k].float().cpu().numpy()
                ref = O.pq_decode_attn(f("q"), inp["kc"], inp["vc"], f("kcent"), f("vcent"), f("kres"), f("vres"), 5)
                row = []
                for impl in (0, 1):
                    out = ops.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], 5, impl=impl).float().cpu().numpy()
                    err = np.abs(out - ref)
                    viol = int((err > 2e-3 + 1e-2 * np.abs(ref)).sum())
                    row.append((float(err.max()), viol))
                worst.append(row)
            fast = max(w[0][0] for w in worst), sum(w[0][1] for w in worst)
            gen = max(w[1][0] for w in worst), sum(w[1][1] for w in worst)
            print(f"q x{qscale:4.1f} {str(dtype):15s} G={nh // nh_k}: fast kernel max|err| {fast[0]:.2e} ({fast[1]} of {3 * nh * 128} beyond 2e-3/1e-2) | all-shapes fp32 kernel {gen[0]:.2e} ({gen[1]})")
