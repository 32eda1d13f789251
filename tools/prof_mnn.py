"""ncu driver: the matcher alone at N=M=8192, D=128."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import posfeat_b200 as P

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator().manual_seed(1)
a = torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=1)
b = torch.nn.functional.normalize(a[torch.randperm(n, generator=g)] + 0.5 * torch.randn(n, 128, generator=g), dim=1)
a, b = a.cuda(), b.cuda()
for _ in range(iters):
    m, nm, _, _ = P.mnn_match(a, b, algo=2)
torch.cuda.synchronize()
print("n", n, "matches", int(nm.item()))
