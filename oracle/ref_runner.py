"""Runs the staged reference (oracle/_ref, see build_ref.py) through its own functions
(TEST / BASELINE INFRASTRUCTURE -- the product never imports this).

The reference path per image pair is exactly what its scripts execute:
  managers/extractor.py:338-346   detector(local_point, **detector_config) -> denormalize -> sample_feat_by_coord
  evaluations/hpatches/evaluation.py:64-67   mnn_matcher(desc_a, desc_b)
on whatever device the tensors live on (CPU: the cpu_baseline / --impl reference arm; cuda: the
"PyTorch-on-B200" comparison leg, with the reference's host-built grid and host syncs included because
that IS the reference path, losses/preprocess_utils.py:217-221).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_mods = None


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "losses", "preprocess_utils.py"))


def modules():
    """(losses.preprocess_utils, losses.preprocess, losses.epipolarloss, aachen matchers) of the staged reference."""
    global _mods
    if _mods is None:
        if not available():
            raise ImportError("oracle/_ref is empty: run `python oracle/build_ref.py` where the reference checkout exists")
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import importlib
        pu = importlib.import_module("losses.preprocess_utils")
        pp = importlib.import_module("losses.preprocess")
        el = importlib.import_module("losses.epipolarloss")
        am = importlib.import_module("ref_eval.aachen_matchers")
        if os.path.dirname(os.path.abspath(pu.__file__)) != os.path.join(REF_DIR, "losses"):
            raise ImportError(f"`losses` resolved to {pu.__file__}, not to the staged reference")
        _mods = (pu, pp, el, am)
    return _mods


def run_pair(score2, fmap2, det_cfg, norm=True):
    """One image pair through the reference's own functions.  score2 [2,1,H,W], fmap2 [2,D,h,w] (any device).
    Returns (kps_n [2,n,2], desc [2,n,D], matches (K,2) int64 ndarray)."""
    import torch
    pu = modules()[0]
    with torch.no_grad():
        kps, _ = pu.generate_kpts_single(score2, **det_cfg)
        desc = pu.sample_feat_by_coord(fmap2, kps, norm)
        m = pu.mnn_matcher(desc[0], desc[1])
    return kps, desc, m


def keypoint_idx(score2, kps):
    """Linear interior-grid index of the reference's keypoints is not returned by generate_kpts_single;
    recover the winners' pixel from the centroid (always within half a pixel of its interior pixel)."""
    import torch
    h, w = score2.shape[2:]
    x = torch.round((kps[..., 0] + 1) * (w - 1) / 2).long() - 1
    y = torch.round((kps[..., 1] + 1) * (h - 1) / 2).long() - 1
    return y * (w - 2) + x
