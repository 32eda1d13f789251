"""Small driver for ncu: window expectation forward+backward and the epipolar line search (C4 shapes)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from posfeat_b200 import preprocess as PP

B, n, D, h, w = 8, 512, 128, 120, 160
g = torch.Generator().manual_seed(7)
q = torch.nn.functional.normalize(torch.randn(B, n, D, generator=g), dim=-1).cuda().requires_grad_(True)
fmap = torch.nn.functional.normalize(torch.randn(B, D, h, w, generator=g), dim=1).cuda().requires_grad_(True)
centre = (torch.rand(B, n, 2, generator=g) * 1.8 - 0.9).cuda()
Fm = torch.randn(B, 3, 3, generator=g).cuda()
coord = torch.stack([torch.rand(B, n, generator=g) * 639, torch.rand(B, n, generator=g) * 479], -1).cuda()
jit = torch.rand(B, n, 2, generator=g).cuda()
for _ in range(2):
    e, _, s, _ = PP.get_expected_correspondence_within_window(q, fmap, centre, 0.1, with_std=True)
    (e.square().sum() + s.sum()).backward()
    PP.epipolar_line_search(coord, Fm, q.detach(), fmap.detach(), 480, 640, jitter=jit)
torch.cuda.synchronize()
print("ok")
