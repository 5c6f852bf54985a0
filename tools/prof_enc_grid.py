import os, sys
sys.path.insert(0, "/root/repo")
import torch
from million_b200 import ops, _lib as L
torch.manual_seed(0)
X = torch.randn(1, 8, 32768, 128, device="cuda").half()
c32 = torch.randn(64, 256, 2, device="cuda").half().float().contiguous()
codes = torch.empty(1, 8, 32768, 64, dtype=torch.uint8, device="cuda")
for _ in range(3): ops.pq_encode_into(X, c32, codes, impl=L.IMPL_GRID)
torch.cuda.synchronize()
