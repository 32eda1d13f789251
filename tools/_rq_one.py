import os, sys
sys.path.insert(0, "/root/repo")
import torch
from posfeat_b200 import _lib
from posfeat_b200.pairs import PairPipeline
P, n = 16, 8192
g = torch.Generator().manual_seed(0)
a = torch.nn.functional.normalize(torch.randn(P, n, 128, generator=g), dim=-1)
b = torch.nn.functional.normalize(a[:, torch.randperm(n, generator=g)] + 0.3 * torch.randn(P, n, 128, generator=g), dim=-1)
desc = torch.stack([a, b], 1).reshape(2 * P, n, 128).cuda()
pipe = PairPipeline(dict(nms_radius=1, num_pts=n), mnn_algo=_lib.MNN_TC)
for _ in range(3):
    pipe.match(desc)
torch.cuda.synchronize()
