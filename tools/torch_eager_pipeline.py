"""The reference's post-backbone pipeline written with stock PyTorch ops, to be timed on the
same B200 next to the CUDA path (the "PyTorch-on-B200" baseline BASELINE.json's north star
names).  It follows what the reference executes per image pair -- generate_kpts_single
(losses/preprocess_utils.py:215-267, grid built on the host and copied, as :217-221 does),
sample_feat_by_coord (:40-53) and mnn_matcher (evaluations/hpatches/evaluation.py:27-38) --
but is a measurement aid only: nothing under posfeat_b200/ imports it."""
import torch
import torch.nn.functional as F


def detect(score, nms_radius, num_pts, thr):
    b, _, h, w = score.shape
    r = int(nms_radius)
    ys = torch.linspace(-1, 1, h).view(1, 1, h, 1).to(score)          # host-built grid, copied per call
    xs = torch.linspace(-1, 1, w).view(1, 1, 1, w).to(score)
    inner = score[:, :, 1:-1, 1:-1]
    padded = F.pad(inner, (r, r, r, r), mode="reflect")
    _, arg = F.max_pool2d(padded, 2 * r + 1, stride=1, return_indices=True)
    hp, wp = h - 2 + 2 * r, w - 2 + 2 * r
    lin = torch.arange(hp * wp, device=score.device).view(1, 1, hp, wp)[:, :, r:hp - r, r:wp - r]
    keep = (arg == lin) & (inner > thr)
    weight = F.avg_pool2d(score, 3, stride=1)
    cx = F.avg_pool2d(score * xs, 3, stride=1) / weight
    cy = F.avg_pool2d(score * ys, 3, stride=1) / weight
    peak = F.max_pool2d(score, 3, stride=1)
    n = int(min(num_pts, int(keep.view(b, -1).sum(1).min())))
    n = max(n, 128)
    _, idx = (keep * inner).view(b, -1).topk(n)
    kps = torch.stack([cx.view(b, -1).gather(1, idx), cy.view(b, -1).gather(1, idx)], -1)
    return kps, peak.view(b, -1).gather(1, idx), idx


def sample(fmap, kps):
    out = F.grid_sample(fmap, kps[:, :, None, :], mode="bilinear", align_corners=False)[..., 0]
    return F.normalize(out.transpose(1, 2), p=2, dim=2)


def mnn(a, b):
    sim = a @ b.t()
    nn12 = sim.max(dim=1)[1]
    nn21 = sim.max(dim=0)[1]
    ids = torch.arange(a.shape[0], device=a.device)
    mask = nn21[nn12] == ids
    return torch.stack([ids[mask], nn12[mask]], -1)


def run_pair(score2, fmap2, cfg):
    """score2 [2,1,H,W], fmap2 [2,D,h,w] on the device -> (idx [2,n], matches [K,2])."""
    kps, _, idx = detect(score2, cfg["nms_radius"], cfg["num_pts"], cfg["thr"])
    desc = sample(fmap2, kps)
    return idx, mnn(desc[0], desc[1])


def time_pairs(score, fmap, cfg, n_pairs, warmup=2):
    """Pairs/s of the eager pipeline on the first n_pairs pairs (CUDA events, fp32 matmul)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False            # the reference runs plain fp32
    try:
        for i in range(min(warmup, n_pairs)):
            run_pair(score[2 * i:2 * i + 2], fmap[2 * i:2 * i + 2], cfg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        first = None
        for i in range(n_pairs):
            r = run_pair(score[2 * i:2 * i + 2], fmap[2 * i:2 * i + 2], cfg)
            if first is None:
                first = r
        e1.record()
        torch.cuda.synchronize()
        return n_pairs / (e0.elapsed_time(e1) * 1e-3), first
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
