"""Debug: fused window append vs append-then-attend over r0 and shapes; prints the first mismatches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from million_b200 import ops
from oracle import pq_oracle as O
for (bs, nh, nh_k, nk) in [(1, 8, 2, 165), (2, 32, 8, 1500), (2, 8, 8, 1500), (1, 8, 2, 37)]:
    inp = O.make_inputs(bs=bs, nh=nh, nh_k=nh_k, nk=nk, d=128, M=64, C=256, Lt=128, seed=3)
    t = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    bad = []
    for r0 in range(128):
        k_new = torch.randn(bs, nh_k, 1, 128, device="cuda", generator=g).half()
        v_new = torch.randn(bs, nh_k, 1, 128, device="cuda", generator=g).half()
        kw1, vw1 = t["kres"].clone(), t["vres"].clone()
        ops.window_append(kw1, vw1, k_new, v_new, r0)
        want = ops.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], kw1, vw1, r0 + 1)
        kw2, vw2 = t["kres"].clone(), t["vres"].clone()
        got = ops.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], kw2, vw2, r0 + 1, k_new=k_new, v_new=v_new)
        torch.cuda.synchronize()
        eo, ek, ev = torch.equal(got, want), torch.equal(kw2, kw1), torch.equal(vw2, vw1)
        if not (eo and ek and ev):
            rows = (kw2 != kw1).any(-1).nonzero().tolist()[:4]
            bad.append((r0, eo, ek, ev, rows, float((got.float() - want.float()).abs().max())))
    print(f"shape bs={bs} nh={nh} nh_k={nh_k} nk={nk}: splits {ops.default_splits(bs, nh_k, nk)}; {len(bad)} bad r0:", bad[:6])
