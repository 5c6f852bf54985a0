#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_outliers.py -m gpu -q -p no:cacheprovider -k "outlier" > gpurun_out/tests_vout.log 2>&1; tail -4 gpurun_out/tests_vout.log
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from million_b200 import ops
bs, nh, nhk, nk = 8, 32, 8, 32640
g = torch.Generator(device="cuda"); g.manual_seed(0)
kcent = torch.randn(64, 256, 2, device="cuda", generator=g).half(); vcent = torch.randn(64, 256, 2, device="cuda", generator=g).half()
L = [(torch.randn(bs, nh, 1, 128, device="cuda").half(), torch.randint(0, 256, (bs, nhk, nk, 64), dtype=torch.uint8, device="cuda"), torch.randint(0, 256, (bs, nhk, nk, 64), dtype=torch.uint8, device="cuda"),
      torch.randn(bs, nhk, 128, 128, device="cuda").half(), torch.randn(bs, nhk, 128, 128, device="cuda").half(),
      (torch.randint(0, 128, (bs, nhk, nk, 2), dtype=torch.uint8, device="cuda"), (0.1 * torch.randn(bs, nhk, nk, 2, device="cuda")).half()),
      (torch.randint(0, 128, (bs, nhk, nk, 2), dtype=torch.uint8, device="cuda"), (0.1 * torch.randn(bs, nhk, nk, 2, device="cuda")).half())) for _ in range(4)]
out = torch.empty(bs, nh, 1, 128, device="cuda", dtype=torch.float16)
for name, kw in (("no store", lambda x: {}), ("K records (2)", lambda x: dict(k_outliers=x[5])), ("K+V records (2,2)", lambda x: dict(k_outliers=x[5], v_outliers=x[6]))):
    def run():
        for x in L: ops.pq_decode_attn(x[0], x[1], x[2], kcent, vcent, x[3], x[4], 128, out=out, **kw(x))
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    print(f"8B shapes 32K bs 8, {name}: {e0.elapsed_time(e1) * 1e3 / 20:.1f} us per launch")
PY
