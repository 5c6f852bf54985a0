"""Candidate-grid encoder (encode_grid.cu) against the exact CUDA-core encoder on random and adversarial inputs, then timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from million_b200 import ops, _lib as L
torch.manual_seed(0)
def check(name, X, c32):
    ref = ops.pq_encode(X, c32, impl=L.IMPL_GENERIC); got = ops.pq_encode(X, c32, impl=L.IMPL_GRID)
    torch.cuda.synchronize()
    print(f"{name}: mismatches {int((ref != got).sum())} of {ref.numel()}")
for dtype in (torch.float16, torch.bfloat16, torch.float32):
    cent = torch.randn(64, 256, 2, device="cuda").to(dtype if dtype != torch.float32 else torch.float16)
    c32 = cent.float().contiguous()
    for shape in ((1, 2, 300), (2, 8, 4096), (1, 1, 1), (3, 5, 1777)):
        check(f"{dtype} randn {shape}", torch.randn(*shape, 128, device="cuda").to(dtype), c32)
    check(f"{dtype} heavy tails", (torch.randn(2, 4, 3000, 128, device="cuda") * torch.where(torch.rand(2, 4, 3000, 128, device="cuda") < 0.02, 30.0, 1.0)).to(dtype), c32)
    check(f"{dtype} tiny", (torch.randn(1, 4, 3000, 128, device="cuda") * 1e-3).to(dtype), c32)
# duplicates, near duplicates, clustered and collinear codebooks, points ON centroids and on midpoints
cent = torch.randn(64, 256, 2, device="cuda").half()
cent[:, 100] = cent[:, 7]; cent[:, 200] = (cent[:, 9].float() * (1 + 2 ** -10)).half()
cent[3] = (torch.randn(256, 2, device="cuda") * 0.01 + 1.0).half()          # one tight cluster: cells overflow -> full scan
cent[4, :, 1] = 0                                                            # collinear
cent[5] = cent[5, 0]                                                         # all centroids identical
c32 = cent.float().contiguous()
X = torch.randn(1, 2, 4000, 128, device="cuda").half()
X[0, 0, :256] = cent.permute(1, 0, 2).reshape(256, 128)                      # exactly on centroids
X[0, 0, 256:511] = ((cent[:, :-1].float() + cent[:, 1:].float()) / 2).permute(1, 0, 2).reshape(255, 128).half()   # midpoints
X[0, 1, :100] = float("nan"); X[0, 1, 100:200] = float("inf"); X[0, 1, 200:300] = 65504.0; X[0, 1, 300:400] = 0.0
check("adversarial codebook", X, c32)
grid = [i for i in range(8, 17)]
for n in (1 << 10, 1 << 14):
    Xg = torch.stack(torch.meshgrid(torch.linspace(-5, 5, 128, device="cuda"), torch.linspace(-5, 5, n // 128, device="cuda"), indexing="ij"), -1).reshape(1, 1, n, 2).half()
    check(f"regular lattice {n}", Xg.repeat(1, 1, 1, 64).contiguous(), c32)
# paged / transposed destinations through the same kernel
X = torch.randn(2, 4, 640, 128, device="cuda").half(); cent = torch.randn(64, 256, 2, device="cuda").half(); c32 = cent.float().contiguous()
ref = ops.pq_encode(X, c32, impl=L.IMPL_GENERIC)
tr = torch.zeros(2, 4, 64, 1024, dtype=torch.uint8, device="cuda"); ops.pq_encode_into(X, c32, tr, t0=128, layout="transposed", impl=L.IMPL_GRID)
print("transposed dst mismatches", int((tr[:, :, :, 128:768].transpose(2, 3) != ref).sum()))
# timing at the Llama-3.1-8B prefill shape: 8 kv heads x 32768 tokens
X = torch.randn(1, 8, 32768, 128, device="cuda").half()
codes = torch.empty(1, 8, 32768, 64, dtype=torch.uint8, device="cuda")
for impl, name in ((L.IMPL_GENERIC, "generic"), (L.IMPL_FAST, "tcgen05"), (L.IMPL_GRID, "grid")):
    ops.pq_encode_into(X, c32, codes, impl=impl); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.pq_encode_into(X, c32, codes, impl=impl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nvec = 8 * 32768
    print(f"{name}: {ms:.3f} ms per K tensor of one layer ({nvec/ms/1e3:.1f} M head-vectors/s, {nvec/ms/1e3/16/32:.3f} Mtok/s for K+V x 32 layers; "
          f"input+codes {nvec*320/ms/1e6:.0f} GB/s)")
for n in (128, 1024):
    Xs = torch.randn(1, 8, n, 128, device="cuda").half(); cs_ = torch.empty(1, 8, n, 64, dtype=torch.uint8, device="cuda")
    for impl, name in ((L.IMPL_FAST, "tcgen05"), (L.IMPL_GRID, "grid")):
        ops.pq_encode_into(Xs, c32, cs_, impl=impl); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): ops.pq_encode_into(Xs, c32, cs_, impl=impl)
        e1.record(); torch.cuda.synchronize()
        print(f"  {n} tokens x 8 heads, {name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call")
