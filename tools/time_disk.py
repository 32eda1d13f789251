"""DiskLoss dense part at the training shape (B=8, 480x640 maps, grid 8 -> 4800 points per image, D=128,
T=60): the fused tensor-core path vs the reference's sequence of tensor ops (losses/kploss.py:158-182)
written with stock PyTorch on the same GPU."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from posfeat_b200.kploss import DiskLoss, _epipolar_lines

torch.backends.cuda.matmul.allow_tf32 = False
B, H, W, D, G, T = 8, 480, 640, 128, 8, 60.0
g = torch.Generator().manual_seed(3)
kp1 = torch.randn(B, 1, H, W, generator=g).cuda().requires_grad_(True)
kp2 = torch.randn(B, 1, H, W, generator=g).cuda().requires_grad_(True)
xf1 = torch.randn(B, D, H // 4, W // 4, generator=g).cuda()
xf2 = (xf1.cpu() + 0.3 * torch.randn(B, D, H // 4, W // 4, generator=g)).cuda()
F1 = torch.tensor([[0., 0., 0.], [0., 0., -1.], [0., 1., 0.]]).repeat(B, 1, 1).cuda()
F2 = F1.transpose(1, 2).contiguous()
cfg = dict(grid_size=G, loss_distance="cos", temperature_base=60, temperature_max=60, epipolar_reward="constant_reward",
           reward_config=dict(reward_thr=2, rescale_thr=False), cor_detach=True, good_reward=1, bad_reward=-0.25,
           kp_penalty=-0.001, match_grad=False)
mod = DiskLoss(cfg)
torch.manual_seed(1)
s1, s2 = mod.point_sample(kp1), mod.point_sample(kp2)
# the draws are replayed: their log-probabilities become leaves so that every timed call can run backward
s1 = (s1[0], s1[1].detach().requires_grad_(True), s1[2])
s2 = (s2[0], s2[1].detach().requires_grad_(True), s2[2])
inputs = {"F1": F1, "F2": F2}
outputs = {"epoch": 0, "preds1": {"local_point": kp1, "local_map": xf1}, "preds2": {"local_point": kp2, "local_map": xf2}}


def ours():
    loss, comp = mod(inputs, outputs, None, samples=(s1, s2))
    loss.backward()
    return loss


def eager():
    from posfeat_b200.preprocess_utils import normalize_coords, sample_feat_by_coord
    (c1, lp1, a1), (c2, lp2, a2) = s1, s2
    b = B
    c1, c2 = c1.reshape(b, -1, 2).float(), c2.reshape(b, -1, 2).float()
    f1 = sample_feat_by_coord(xf1, normalize_coords(c1, H, W), True)
    f2 = sample_feat_by_coord(xf2, normalize_coords(c2, H, W), True)
    with torch.no_grad():
        aff = -T * (1 - f1 @ f2.transpose(1, 2))
    ci, ct = torch.distributions.Categorical(logits=aff), torch.distributions.Categorical(logits=aff.transpose(1, 2))
    dense_p = ci.probs * ct.probs.transpose(1, 2)
    dense_logp = ci.logits + ct.logits.transpose(1, 2)
    l1, l2 = _epipolar_lines(F1, c1), _epipolar_lines(F2, c2)
    ch1 = torch.cat([c1, torch.ones_like(c1[..., :1])], -1)
    ch2 = torch.cat([c2, torch.ones_like(c2[..., :1])], -1)
    d1 = (l1 @ ch2.transpose(1, 2)).abs()
    d2 = (l2 @ ch1.transpose(1, 2)).abs().transpose(1, 2)
    good = (d1 < 2) & (d2 < 2)
    reward = 1.0 * good + (-0.25) * (~good)
    kl = lp1.reshape(b, 1, -1).transpose(1, 2) + lp2.reshape(b, 1, -1)
    plogp = dense_p.detach() * (dense_logp + kl)
    acc = a1.reshape(b, 1, -1).transpose(1, 2) * a2.reshape(b, 1, -1)
    reinforce = (reward[acc] * plogp[acc]).sum()
    pen = -0.001 * (lp1[a1].sum() + lp2[a2].sum())
    loss = -reinforce - pen
    loss.backward()
    return loss


def timeit(fn, iters=5):
    for _ in range(2):
        v = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        v = fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters, float(v)


t_o, l_o = timeit(ours)
t_e, l_e = timeit(eager)
print(json.dumps({"case": "disk_loss_B8_4800x4800_D128_T60", "ours_fwd_bwd_us": t_o, "torch_fwd_bwd_us": t_e,
                  "loss_ours": l_o, "loss_torch": l_e, "rel_diff": abs(l_o - l_e) / abs(l_e)}))
