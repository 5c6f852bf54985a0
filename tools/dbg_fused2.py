import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from million_b200 import ops
from oracle import pq_oracle as O
bs, nh, nh_k, nk = 1, 8, 2, 165
inp = O.make_inputs(bs=bs, nh=nh, nh_k=nh_k, nk=nk, d=128, M=64, C=256, Lt=128, seed=3)
t = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
g = torch.Generator(device="cuda"); g.manual_seed(1)
for r0 in (7, 8, 20, 40):
    for ns in (0, 1, 2, 3):
        k_new = torch.randn(bs, nh_k, 1, 128, device="cuda", generator=g).half()
        v_new = torch.randn(bs, nh_k, 1, 128, device="cuda", generator=g).half()
        kw2, vw2 = t["kres"].clone(), t["vres"].clone()
        got = ops.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], kw2, vw2, r0 + 1, k_new=k_new, v_new=v_new, n_splits=ns)
        torch.cuda.synchronize()
        diff = (kw2 != t["kres"]).any(-1)[0, 0].nonzero().flatten().tolist()
        desc = []
        for row in diff:
            x = kw2[0, 0, row]
            if torch.equal(x, k_new[0, 0, 0]): desc.append((row, "k_new"))
            elif torch.equal(x, v_new[0, 0, 0]): desc.append((row, "v_new"))
            else:
                m = [j for j in range(128) if torch.equal(x, t["kres"][0, 0, j])]
                desc.append((row, f"kres row {m}" if m else "other", x[:4].tolist(), t["kres"][0,0,row,:4].tolist()))
        print(f"r0={r0} n_splits={ns}: changed rows {desc}")
