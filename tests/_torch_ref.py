"""Plain PyTorch restatements used ONLY by the tests of the floating-point
training-side kernels (autograd gives the reference gradients).  Each function
cites the reference lines it restates."""
import torch
import torch.nn.functional as F


def corr_expect_ref(q, k, v, scale):
    """softmax(scale * q k^T) @ v -- losses/preprocess.py:59-63 / preprocess_utils.py:100-106."""
    p = torch.softmax(scale * q @ k.transpose(1, 2), dim=-1)
    return p @ (v if v.dim() == 3 else v[None])


def window_ref(feat1, fmap, centre, offsets):
    """losses/preprocess_utils.py:721-758 (with_std=True)."""
    B, d, h, w = fmap.shape
    cg = centre.unsqueeze(-2) + offsets[None, None]
    win = F.grid_sample(fmap, cg, padding_mode="zeros", align_corners=False).permute(0, 2, 3, 1)
    sim = (feat1.unsqueeze(-2) * win).sum(-1)
    prob = torch.softmax(sim, -1)
    e = (cg * prob.unsqueeze(-1)).sum(2)
    var = (cg ** 2 * prob.unsqueeze(-1)).sum(2) - e ** 2
    return e, torch.sqrt(torch.clamp(var, min=1e-10)).sum(-1), prob


def homogenize(c):
    return torch.cat((c, torch.ones_like(c[..., [0]])), -1)


def epipolar_loss_full(inputs, processed, cfg):
    """losses/epipolarloss.py:38-101 (consumer of subsystem 4; stays PyTorch in the product)."""
    def cost(c1, c2, Fm):
        line = Fm.bmm(homogenize(c1).transpose(1, 2))
        line = line / torch.clamp(torch.norm(line[:, :2, :], dim=1, keepdim=True), min=1e-8)
        return torch.abs(torch.sum(homogenize(c2).transpose(1, 2) * line, dim=1))

    def weight(inv_std, mask):
        wgt = (inv_std / torch.mean(inv_std)).detach()
        wgt = wgt * mask.float()
        return wgt / (torch.mean(wgt) + 1e-8)
    c1, c2 = processed["coord1"], processed["coord2"]
    short = min(inputs["im1"].shape[2:])
    terms = []
    for a, cc, key, Fm, valid in ((c1, "feat1g_corloc", "feat1g_std", inputs["F1"], processed["valid_epi1"]),
                                  (c1, "feat1w_corloc", "feat1w_std", inputs["F1"], processed["valid_epi1"]),
                                  (c2, "feat2g_corloc", "feat2g_std", inputs["F2"], processed["valid_epi2"]),
                                  (c2, "feat2w_corloc", "feat2w_std", inputs["F2"], processed["valid_epi2"])):
        cst = cost(a, processed[cc], Fm)
        thr = cfg["grid_cost_thr"] if "g_" in cc else cfg["win_cost_thr"]
        mask = (cst < short * thr) & valid
        wg = weight(1 / processed[key].clamp(min=1e-10), mask)
        terms.append((wg * cst).mean())
    lg1, lw1, lg2, lw2 = terms
    return cfg["weight_grid"] * (lg1 + lg2) + cfg["weight_window"] * (lw1 + lw2)
