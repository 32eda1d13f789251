"""DiskLoss (second training stage) with the dense affinity part on the B200 kernels.

Mirrors losses/kploss.py:7-196 of the reference: same constructor keys (`DiskLoss_config` of
configs/train_kp.yaml), same `forward(inputs, outputs, processed) -> (loss, components)`.

What runs where
  * keypoint proposal sampling (`point_sample`, kploss.py:20-50) is a handful of tensor ops on the
    [b,1,h/g,w/g,g*g] unfolded score logits and stays in torch (it owns the random draws);
  * descriptors at the sampled points: the bilinear sampler kernel;
  * everything that touches the dense [b,m,n] affinity (kploss.py:158-182: two Categorical
    normalisations, their product and log, the epipolar reward, the accept mask, the REINFORCE sum and
    the statistics reported as components) is fused: two tensor-core passes give the row and column
    log-sum-exp, two more (roles swapped) give per point the sums the loss and its gradient need.
    No [b,m,n] tensor is ever materialised (the reference holds about ten of them).

The fused path serves `match_grad: False`, `cor_detach: True`, `rescale_thr: False` (the shipped
config): the affinity then carries no gradient and the loss is differentiable only with respect to the
keypoint log-probabilities.  The other settings need what the fused kernels never produce -- gradients
THROUGH the [b,m,n] affinity (`match_grad: True`, `cor_detach: False`) or the mean epipolar distance before
the rewards (`rescale_thr: True`) -- and take `_forward_dense`: the same mathematics as [b,m,n] tensors with
autograd (kploss.py:52-129, :158-182), descriptors still from the bilinear sampler kernel (with its backward
kernel when the affinity carries gradient).  Ablation settings only: at the training shape (b=8, 4800 points)
this path holds a handful of 737 MB tensors, which is exactly what the fused path avoids.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._runtime import check, lib, require_cuda, stream_ptr, workspace
from .preprocess import corr_expect, homogenize
from .preprocess_utils import normalize_coords, sample_feat_by_coord


def _unfold(t, g):
    """losses/preprocess_utils.py:344-350: [b,c,h,w] -> [b,c,h/g,w/g,g*g] non-overlapping cells."""
    u = t.unfold(2, g, g).unfold(3, g, g)
    b, c, h, w, g1, g2 = u.shape
    return u.reshape(b, c, h, w, g1 * g2)


def _epipolar_lines(Fmat, coord):
    """Normalised lines F x~ (kploss.py:59-62): [B,n,3] with the first two components of unit norm."""
    line = Fmat.float().bmm(homogenize(coord).transpose(1, 2))
    line = line / torch.clamp(torch.norm(line[:, :2, :], p=2, dim=1, keepdim=True), min=1e-8)
    return line.transpose(1, 2).contiguous()


def _rows(feat_a, feat_b, tab_a, tab_b, T, thr_a, thr_b, good, bad, dynamic):
    """[B,n,4] = per point of side a: reinforce part, reward-weighted p, sum p, max p."""
    B, n, D = feat_a.shape
    m = feat_b.shape[1]
    dev = feat_a.device
    out = torch.empty((B, n, 4), dtype=torch.float32, device=dev)
    L = lib()
    with torch.cuda.device(dev):
        ws = workspace("disk", L.posfeat_dual_softmax_reward_workspace_bytes(B, n, m, D), dev)
        check(L.posfeat_dual_softmax_reward_f32(feat_a.data_ptr(), feat_b.data_ptr(), tab_a.data_ptr(), tab_b.data_ptr(),
                                                B, n, m, D, float(T), float(thr_a), float(thr_b), float(good), float(bad),
                                                int(dynamic), out.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(dev)))
    return out


class _Reinforce(torch.autograd.Function):
    """reinforce = sum_ij acc r p (log p + logp1_i + logp2_j); only logp1 / logp2 carry gradient."""

    @staticmethod
    def forward(ctx, logp1, logp2, value, r1, r2):
        ctx.save_for_backward(r1, r2)
        return value.clone()

    @staticmethod
    def backward(ctx, g):
        r1, r2 = ctx.saved_tensors
        return g * r1, g * r2, None, None, None


class DiskLoss(nn.Module):
    def __init__(self, configs, device=None):
        super().__init__()
        self.__lossname__ = "DiskLoss"
        self.config = configs
        self.unfold_size = configs["grid_size"]
        self.t_base = configs["temperature_base"]
        self.t_max = configs["temperature_max"]
        if configs["epipolar_reward"] not in ("constant_reward", "dynamic_reward"):
            raise AttributeError(configs["epipolar_reward"])          # getattr(self, ...) in the reference
        self.dynamic = configs["epipolar_reward"] == "dynamic_reward"
        self.good_reward = configs["good_reward"]
        self.bad_reward = configs["bad_reward"]
        self.kp_penalty = configs["kp_penalty"]
        self.match_grad = bool(configs.get("match_grad"))
        self.cor_detach = bool(configs.get("cor_detach", True))
        self.rescale_thr = bool(configs["reward_config"].get("rescale_thr"))
        self.dense = self.match_grad or not self.cor_detach or self.rescale_thr       # see _forward_dense

    def name(self):
        return self.__lossname__

    # ---- kploss.py:20-50 ------------------------------------------------------------------------
    def point_distribution(self, logits):
        from torch.distributions import Bernoulli, Categorical
        prop = Categorical(logits=logits)
        proposals = prop.sample()
        accept_logits = torch.gather(logits, dim=-1, index=proposals[..., None]).squeeze(-1)
        acc = Bernoulli(logits=accept_logits)
        accept = acc.sample()
        return proposals, accept == 1., prop.log_prob(proposals) + acc.log_prob(accept)

    def point_sample(self, kp_map):
        g = self.unfold_size
        proposals, accept_mask, logp = self.point_distribution(_unfold(kp_map, g))
        b, _, h, w = kp_map.shape
        ys, xs = torch.meshgrid(torch.arange(h, device=kp_map.device, dtype=kp_map.dtype),
                                torch.arange(w, device=kp_map.device, dtype=kp_map.dtype), indexing="ij")
        grid = torch.stack([xs, ys], 0)[None].expand(b, 2, h, w)                 # pixel (x, y) of every cell entry
        kps = _unfold(grid, g).gather(dim=4, index=proposals.unsqueeze(-1).repeat(1, 2, 1, 1, 1))
        return kps.squeeze(4).permute(0, 2, 3, 1), logp, accept_mask

    # ---- kploss.py:132-196 ----------------------------------------------------------------------
    def forward(self, inputs, outputs, processed=None, samples=None):
        """``samples`` = ((coord1, logp1, accept1), (coord2, logp2, accept2)) replaces the two random
        `point_sample` draws (tests, reproducible runs)."""
        require_cuda()
        preds1, preds2 = outputs["preds1"], outputs["preds2"]
        kp_map1, kp_map2 = preds1["local_point"], preds2["local_point"]
        xf1, xf2 = preds1["local_map"], preds2["local_map"]
        b = xf1.shape[0]
        _, _, h, w = kp_map1.shape
        T = min(self.t_base + outputs["epoch"], self.t_max)
        (coord1, logp1, acc1), (coord2, logp2, acc2) = samples if samples is not None else (
            self.point_sample(kp_map1), self.point_sample(kp_map2))
        coord1 = coord1.reshape(b, -1, 2).float()
        coord2 = coord2.reshape(b, -1, 2).float()
        lp1, lp2 = logp1.reshape(b, -1), logp2.reshape(b, -1)
        a1, a2 = acc1.reshape(b, -1), acc2.reshape(b, -1)
        cos = self.config["loss_distance"] == "cos"
        if self.dense:
            return self._forward_dense(inputs, xf1, xf2, coord1, coord2, lp1, lp2, a1, a2, h, w, T, cos)
        with torch.no_grad():
            f1 = sample_feat_by_coord(xf1, normalize_coords(coord1, h, w), cos).contiguous()
            f2 = sample_feat_by_coord(xf2, normalize_coords(coord2, h, w), cos).contiguous()
            m, n = f1.shape[1], f2.shape[1]
            ones1 = torch.ones((n, 1), dtype=torch.float32, device=f1.device)
            ones2 = torch.ones((m, 1), dtype=torch.float32, device=f1.device)
            # logsumexp of T*<f1,f2> over columns and over rows (the constant -T of the affinity is added below)
            lse_r = corr_expect(f1, f2, ones1, float(T), want_lse=True)[1] - T
            lse_c = corr_expect(f2, f1, ones2, float(T), want_lse=True)[1] - T
            l1 = _epipolar_lines(inputs["F1"], coord1)          # lines of image-1 points in image 2
            l2 = _epipolar_lines(inputs["F2"], coord2)
            tab1 = torch.cat([lse_r[..., None], l1, coord1, lp1.detach()[..., None].float(), a1[..., None].float()], -1).contiguous()
            tab2 = torch.cat([lse_c[..., None], l2, coord2, lp2.detach()[..., None].float(), a2[..., None].float()], -1).contiguous()
            thr = self.config["reward_config"]["reward_thr"]
            r1 = _rows(f1, f2, tab1, tab2, T, thr, thr, self.good_reward, self.bad_reward, self.dynamic)
            r2 = _rows(f2, f1, tab2, tab1, T, thr, thr, self.good_reward, self.bad_reward, self.dynamic)
        reinforce = _Reinforce.apply(lp1, lp2, r1[..., 0].sum(), r1[..., 1], r2[..., 1])
        kp_penalty = self.kp_penalty * (lp1[a1].sum() + lp2[a2].sum())
        loss = -reinforce - kp_penalty
        with torch.no_grad():
            rowsum, colsum = r1[..., 2], r2[..., 2]                    # sample_p.sum(2), sample_p.sum(1)
            pmax_b = r1[..., 3].max(-1)[0]
            psum_b = rowsum.sum(-1)
            one = rowsum.new_tensor(1.)
            components = {"reinforce": reinforce.detach(), "kp_penalty": kp_penalty.detach(), "scale1": one, "scale2": one,
                          "cor minmax": pmax_b.min(), "cor minmean": (psum_b / (m * n)).min(), "cor max": pmax_b.max(),
                          "cor mean": psum_b.sum() / (b * m * n),
                          "cor summin": torch.min(colsum.min(), rowsum.min()), "cor summax": torch.max(colsum.max(), rowsum.max()),
                          "n_kps": (a1.sum(-1, keepdim=True) + a2.sum(-1, keepdim=True)).float().mean(),
                          "n_pairs": psum_b.mean(), "temperature": rowsum.new_tensor(float(T))}
        return loss, components

    # ---- the settings the fused kernels do not serve: [b,m,n] tensors with autograd ---------------------------
    def _reward_dense(self, inputs, coord1, coord2):
        """constant_reward / dynamic_reward incl. rescale_thr (kploss.py:52-129) as [b,m,n] tensors, no gradient."""
        with torch.no_grad():
            l1 = _epipolar_lines(inputs["F1"], coord1)                       # [b,m,3]: lines of image-1 points in image 2
            l2 = _epipolar_lines(inputs["F2"], coord2)                       # [b,n,3]
            d12 = torch.abs(l1 @ homogenize(coord2).transpose(1, 2))         # [b,m,n] distance of point j to the line of i
            d21 = torch.abs(l2 @ homogenize(coord1).transpose(1, 2)).transpose(1, 2)
            thr = self.config["reward_config"]["reward_thr"]
            if self.rescale_thr:
                b = d12.shape[0]
                m1 = d12.reshape(b, -1).mean(1, True)
                m2 = d21.reshape(b, -1).mean(1, True)
                low = torch.minimum(m1, m2).clamp(1e-6)
                scale1, scale2 = m1 / low, m2 / low
                thr1, thr2 = (thr * scale1).reshape(b, 1, 1), (thr * scale2).reshape(b, 1, 1)
            else:
                thr1 = thr2 = thr
                scale1 = scale2 = d12.new_tensor(1.)
            if self.dynamic:
                reward = (torch.exp(-d12 / thr1) + torch.exp(-d21 / thr2) - 2 / torch.exp(torch.ones_like(d12)))
                reward = reward.clamp(min=self.bad_reward)
            else:
                good = (d12 < thr1) & (d21 < thr2)
                reward = self.good_reward * good + self.bad_reward * (~good)
        return reward, scale1, scale2

    def _forward_dense(self, inputs, xf1, xf2, coord1, coord2, lp1, lp2, a1, a2, h, w, T, cos):
        b = xf1.shape[0]
        with torch.set_grad_enabled(self.match_grad and torch.is_grad_enabled()):
            f1 = sample_feat_by_coord(xf1, normalize_coords(coord1, h, w), cos)
            f2 = sample_feat_by_coord(xf2, normalize_coords(coord2, h, w), cos)
            affinity = -T * (1 - f1 @ f2.transpose(1, 2))                    # [b,m,n]
        m, n = f1.shape[1], f2.shape[1]
        logr = torch.log_softmax(affinity, dim=2)                            # Categorical(logits=affinity).logits
        logc = torch.log_softmax(affinity, dim=1)                            # ... of the transposed affinity, transposed back
        dense_logp = logr + logc
        dense_p = torch.exp(dense_logp)                                      # = probs_I * probs_T
        sample_p = dense_p.detach() if self.cor_detach else dense_p
        reward, scale1, scale2 = self._reward_dense(inputs, coord1, coord2)
        accept = (a1[:, :, None] & a2[:, None, :]).to(dense_p.dtype)
        plogp = sample_p * (dense_logp + lp1[:, :, None] + lp2[:, None, :])
        reinforce = (reward * plogp * accept).sum()                          # = (reward[mask] * plogp[mask]).sum()
        kp_penalty = self.kp_penalty * (lp1[a1].sum() + lp2[a2].sum())
        loss = -reinforce - kp_penalty
        with torch.no_grad():
            sp = sample_p.detach()
            components = {"reinforce": reinforce.detach(), "kp_penalty": kp_penalty.detach(), "scale1": scale1, "scale2": scale2,
                          "cor minmax": sp.reshape(b, -1).max(-1)[0].min(), "cor minmean": sp.reshape(b, -1).mean(-1).min(),
                          "cor max": sp.max(), "cor mean": sp.mean(),
                          "cor summin": torch.min(sp.sum(1).min(), sp.sum(2).min()),
                          "cor summax": torch.max(sp.sum(1).max(), sp.sum(2).max()),
                          "n_kps": (a1.sum(-1, keepdim=True) + a2.sum(-1, keepdim=True)).float().mean(),
                          "n_pairs": sp.sum(-1).sum(-1).mean(), "temperature": sp.new_tensor(float(T))}
        return loss, components
