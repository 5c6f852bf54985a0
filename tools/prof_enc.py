import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from million_b200 import ops, _lib as L
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
torch.manual_seed(0)
X = torch.randn(1, 8, n, 128, device="cuda").half()
cent = torch.randn(64, 256, 2, device="cuda").half(); c32 = cent.float().contiguous()
codes = torch.empty(1, 8, n, 64, dtype=torch.uint8, device="cuda")
for _ in range(3): ops.pq_encode_into(X, c32, codes, impl=impl)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.pq_encode_into(X, c32, codes, impl=impl)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"impl {impl}: {ms:.3f} ms for {8*n} head-vectors -> {8*n/ms/1e3:.1f} M vec/s")
