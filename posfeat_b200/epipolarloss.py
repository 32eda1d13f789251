"""Drop-in for ``losses/epipolarloss.py`` (EpipolarLoss_full, :8-101): the consumer of the training-side
expectations.  SURVEY.md 8a row a-19 keeps it in plain tensor arithmetic -- a few [B,n] element-wise
operations on what the correlation kernels produced -- so there is no kernel here; the class exists so that
``getattr(losses, 'EpipolarLoss_full')(cfg, device)`` (managers/trainer.py:96-107) finds the same name, config
keys (``weight_grid``, ``weight_window``, ``grid_cost_thr``, ``win_cost_thr``, ``use_std_as_weight``), call
signature and return value ``(loss, components)`` in this package."""
from __future__ import annotations

import torch
import torch.nn as nn


def _homogeneous(c):
    return torch.cat((c, torch.ones_like(c[..., :1])), -1)


def point_to_line_distance(coord1, coord2, fmatrix):
    """|l . x2| with l = F x1 scaled to a unit normal (epipolarloss.py:16-22): [B,n,2] x2 -> [B,n]."""
    line = fmatrix.bmm(_homogeneous(coord1).transpose(1, 2))                      # B,3,n
    line = line / torch.clamp(torch.norm(line[:, :2, :], dim=1, keepdim=True), min=1e-8)
    return torch.abs(torch.sum(_homogeneous(coord2).transpose(1, 2) * line, dim=1))


class EpipolarLoss_full(nn.Module):
    def __init__(self, configs, device=None):
        super().__init__()
        self.__lossname__ = "EpipolarLoss_fullinfo"
        self.config = configs
        self.w_g = configs["weight_grid"]
        self.w_w = configs["weight_window"]

    def _weights(self, std, mask):
        """epipolarloss.py:25-36 with inverse_std = 1 / clamp(std, 1e-10): detached, mean 1, then masked and
        renormalised.  ``use_std_as_weight: False`` reads an undefined name in the reference (``std``)."""
        if not self.config["use_std_as_weight"]:
            raise NameError("name 'std' is not defined (use_std_as_weight=False takes this branch in the reference, "
                            "losses/epipolarloss.py:31)")
        inv = 1 / std.clamp(min=1e-10)
        w = (inv / inv.mean()).detach() * mask.float()
        return w / (w.mean() + 1e-8)

    def forward(self, inputs, outputs, processed):
        short = min(inputs["im1"].size()[2:])
        has_valid = "valid_epi1" in processed
        terms = {}
        for side in ("1", "2"):
            src, fm = processed["coord" + side], inputs["F" + side]
            for kind, thr in (("g", self.config["grid_cost_thr"]), ("w", self.config["win_cost_thr"])):
                cost = point_to_line_distance(src, processed[f"feat{side}{kind}_corloc"], fm)
                mask = cost < short * thr                                            # large costs dropped, as in CAPS
                if has_valid:
                    mask = mask & processed["valid_epi" + side]
                w = self._weights(processed[f"feat{side}{kind}_std"], mask)
                terms[kind + side] = ((w * cost).mean(), mask)
        loss = self.w_g * (terms["g1"][0] + terms["g2"][0]) + self.w_w * (terms["w1"][0] + terms["w2"][0])

        def kept(a, b):
            return (a.sum() / a.numel() + b.sum() / b.numel()) / 2
        components = {"loss_g1": terms["g1"][0], "loss_w1": terms["w1"][0], "loss_g2": terms["g2"][0],
                      "loss_w2": terms["w2"][0], "percent_g": kept(terms["g1"][1], terms["g2"][1]),
                      "percent_w": kept(terms["w1"][1], terms["w2"][1])}
        return loss, components
