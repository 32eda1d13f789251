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


def run_pair(score2, fmap2, det_cfg, norm=True, want_idx=False):
    """One image pair through the reference's own functions.  score2 [2,1,H,W], fmap2 [2,D,h,w] (any device).
    Returns (kps_n [2,n,2], desc [2,n,D], matches (K,2) int64 ndarray) and, with ``want_idx``, the keypoint
    indices the reference's own ``topk`` call returned (losses/preprocess_utils.py:264; the function does not
    return them, so the call is observed -- not altered -- through a wrapper around Tensor.topk)."""
    import torch
    pu = modules()[0]
    seen = []
    orig = torch.Tensor.topk

    def spy(self, *a, **k):
        r = orig(self, *a, **k)
        seen.append(r[1])
        return r
    with torch.no_grad():
        if want_idx:
            torch.Tensor.topk = spy
        try:
            kps, _ = pu.generate_kpts_single(score2, **det_cfg)
        finally:
            torch.Tensor.topk = orig
        desc = pu.sample_feat_by_coord(fmap2, kps, norm)
        m = pu.mnn_matcher(desc[0], desc[1])
    if want_idx:
        assert len(seen) == 1, "generate_kpts_single is expected to call topk exactly once"
        return kps, desc, m, seen[0]
    return kps, desc, m
