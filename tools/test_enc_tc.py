import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from million_b200 import ops, _lib as L
torch.manual_seed(0)
for dtype in (torch.float16, torch.bfloat16):
    for (M, dm) in ((64, 2), (32, 4)):
        cent = torch.randn(M, 256, dm, device="cuda").to(dtype)
        c32 = cent.float().contiguous()
        for shape in ((1, 2, 300), (2, 8, 4096), (1, 1, 1)):
            X = torch.randn(*shape, 128, device="cuda").to(dtype)
            ref = ops.pq_encode(X, c32, impl=L.IMPL_GENERIC)
            got = ops.pq_encode(X, c32, impl=L.IMPL_FAST)
            torch.cuda.synchronize()
            nm = int((ref != got).sum())
            print(dtype, M, dm, shape, "mismatches", nm, "of", ref.numel())
# near-duplicate centroids / ties
cent = torch.randn(64, 256, 2, device="cuda").half()
cent[:, 100] = cent[:, 7]
cent[:, 200] = (cent[:, 9].float() * (1 + 2 ** -10)).half()
c32 = cent.float().contiguous()
X = torch.cat([cent[:, 7].reshape(1, 1, 1, 128), cent[:, 9].reshape(1, 1, 1, 128), torch.randn(1, 1, 2000, 128, device="cuda").half()], 2).contiguous()
ref = ops.pq_encode(X, c32, impl=L.IMPL_GENERIC); got = ops.pq_encode(X, c32, impl=L.IMPL_FAST)
print("ties: mismatches", int((ref != got).sum()), "code of duplicate", int(got[0, 0, 0, 0]))
# timing at Llama-3.1-8B prefill shape: 8 kv heads x 32768 tokens
X = torch.randn(1, 8, 32768, 128, device="cuda").half()
cent = torch.randn(64, 256, 2, device="cuda").half(); c32 = cent.float().contiguous()
codes = torch.empty(1, 8, 32768, 64, dtype=torch.uint8, device="cuda")
for impl, name in ((L.IMPL_GENERIC, "generic"), (L.IMPL_FAST, "tcgen05")):
    ops.pq_encode_into(X, c32, codes, impl=impl); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.pq_encode_into(X, c32, codes, impl=impl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    nvec = 8 * 32768
    print(f"{name}: {ms:.3f} ms per K tensor of one layer ({nvec/ms/1e3:.1f} M head-vectors/s, {nvec/ms/1e3/16/32:.3f} Mtok/s for K+V x 32 layers, {nvec*65536*2/ms/1e9:.1f} TFLOP/s algorithmic)")
