"""Ratio-test matchers vs the same maths in stock PyTorch on this GPU (ETH config: 8k keypoints, ratio 0.75)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from posfeat_b200.matchers import mutual_nn_ratio_matcher

torch.backends.cuda.matmul.allow_tf32 = False


def torch_mutual_ratio(d1, d2, ratio):
    sim = d1 @ d2.t()
    s12, n12 = torch.topk(sim, 2, dim=1)
    r12 = torch.sqrt(2 - 2 * s12[:, 0]) / (torch.sqrt(2 - 2 * s12[:, 1]) + 1e-8)
    s21, n21 = torch.topk(sim.t(), 2, dim=1)
    r21 = torch.sqrt(2 - 2 * s21[:, 0]) / (torch.sqrt(2 - 2 * s21[:, 1]) + 1e-8)
    nn12, nn21 = n12[:, 0], n21[:, 0]
    ids = torch.arange(sim.shape[0], device=d1.device)
    mask = (ids == nn21[nn12]) & (r12 <= ratio) & (r21[nn12] <= ratio)
    return torch.stack([ids[mask], nn12[mask]], -1).cpu().numpy()


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters, r


for n in (1024, 4096, 8192):
    g = torch.Generator().manual_seed(n)
    a = torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=1)
    lvl = 0.05 + 0.2 * torch.rand(n, 1, generator=g)
    b = torch.nn.functional.normalize(a[torch.randperm(n, generator=g)] + lvl * torch.randn(n, 128, generator=g), dim=1)
    a, b = a.cuda(), b.cuda()
    t_ours, m_ours = timeit(lambda: mutual_nn_ratio_matcher(a, b, ratio=0.75))
    t_torch, m_torch = timeit(lambda: torch_mutual_ratio(a, b, 0.75))
    so, st = {tuple(x) for x in m_ours.tolist()}, {tuple(x) for x in m_torch.tolist()}
    print(json.dumps({"N": n, "ours_us": t_ours, "torch_us": t_torch, "matches": len(so), "differing_vs_torch": len(so ^ st)}), flush=True)
