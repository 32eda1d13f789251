"""Time the matcher kernels with the library's own per-kernel CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import posfeat_b200 as P
from posfeat_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
g = torch.Generator().manual_seed(1)
a = torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=1)
b = torch.nn.functional.normalize(a[torch.randperm(n, generator=g)] + 0.5 * torch.randn(n, 128, generator=g), dim=1)
a, b = a.cuda(), b.cuda()
for _ in range(3):
    P.mnn_match(a, b, algo=2)
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(iters):
    m, nm, _, _ = P.mnn_match(a, b, algo=2)
torch.cuda.synchronize()
pr = _lib.profile_read()
_lib.profile_enable(False)
print("n", n, "debug", os.environ.get("POSFEAT_TC_DEBUG", "0"), "matches", int(nm.item()),
      {k: round(1e3 * v[0] / v[1], 1) for k, v in pr.items()})
