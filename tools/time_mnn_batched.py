"""Time the batched matches-only matcher (the pipeline's call) with per-kernel events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from posfeat_b200 import _lib
from posfeat_b200.pairs import PairPipeline

P = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
g = torch.Generator().manual_seed(1)
a = torch.nn.functional.normalize(torch.randn(P, n, 128, generator=g), dim=-1)
b = torch.nn.functional.normalize(a + 0.5 * torch.randn(P, n, 128, generator=g), dim=-1)
desc = torch.stack([a, b], 1).reshape(2 * P, n, 128).cuda()
pipe = PairPipeline(dict(nms_radius=1, num_pts=n))
for _ in range(2):
    m, nm = pipe.match(desc)
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(5):
    m, nm = pipe.match(desc)
torch.cuda.synchronize()
pr = _lib.profile_read()
print("P", P, "n", n, "debug", os.environ.get("POSFEAT_TC_DEBUG", "0"), "mean matches", float(nm.float().mean()),
      {k: round(1e3 * v[0] / v[1] / P, 2) for k, v in pr.items()}, "us per pair")
