"""Small driver for ncu: two DiskLoss forward+backward passes at the training shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from posfeat_b200.kploss import DiskLoss

B, H, W, D = 8, 480, 640, 128
g = torch.Generator().manual_seed(3)
kp1 = torch.randn(B, 1, H, W, generator=g).cuda().requires_grad_(True)
kp2 = torch.randn(B, 1, H, W, generator=g).cuda().requires_grad_(True)
xf1 = torch.randn(B, D, H // 4, W // 4, generator=g).cuda()
xf2 = (xf1.cpu() + 0.3 * torch.randn(B, D, H // 4, W // 4, generator=g)).cuda()
F1 = torch.tensor([[0., 0., 0.], [0., 0., -1.], [0., 1., 0.]]).repeat(B, 1, 1).cuda()
cfg = dict(grid_size=8, loss_distance="cos", temperature_base=60, temperature_max=60, epipolar_reward="constant_reward",
           reward_config=dict(reward_thr=2, rescale_thr=False), cor_detach=True, good_reward=1, bad_reward=-0.25,
           kp_penalty=-0.001, match_grad=False)
mod = DiskLoss(cfg)
inputs = {"F1": F1, "F2": F1.transpose(1, 2).contiguous()}
outputs = {"epoch": 0, "preds1": {"local_point": kp1, "local_map": xf1}, "preds2": {"local_point": kp2, "local_map": xf2}}
for _ in range(2):
    loss, comp = mod(inputs, outputs, None)
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
