"""Small driver for ncu: one dense-fine and one dense-coarse forward+backward (C4 shapes), the map normalised by the
package's one-pass scale * normalize -> channels_last kernels (csrc/normalize.cu) as in the training step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from posfeat_b200 import preprocess as PP

B, n, D = 8, 512, 128
g = torch.Generator().manual_seed(7)
for h, w in ((120, 160), (30, 40)):
    q = torch.nn.functional.normalize(torch.randn(B, n, D, generator=g), dim=-1).cuda().requires_grad_(True)
    fmap = torch.randn(B, D, h, w, generator=g).cuda().requires_grad_(True)
    for _ in range(2):
        out = PP.get_expected_correspondence_locs(q, PP.normalize_scale_channels_last(fmap, 20.0))
        out.square().sum().backward()
torch.cuda.synchronize()
print("ok")
