"""posfeat_b200 -- B200-native (sm_100a) post-backbone feature pipeline of PoSFeat.

Public surface mirrors the reference's plugin points:
  generate_kpts_single, sample_feat_by_coord, mnn_matcher          (losses/preprocess_utils.py)
  mutual_nn_matcher                                                 (evaluations/*/matchers)
  Preprocess_Line2Window and the correlation expectations           (losses/preprocess.py)
  process / save_desc                                               (managers/extractor.py)
"""
from ._lib import LIB_PATH, PosfeatError  # noqa: F401

__all__ = ["generate_kpts_single", "generate_kpts_single_noavg", "sample_feat_by_coord", "mnn_matcher",
           "mutual_nn_matcher", "ratio_matcher", "mutual_nn_ratio_matcher", "AsyncDescWriter",
           "normalize_coords", "denormalize_coords", "compute_prob", "install", "LIB_PATH", "PosfeatError"]

_LAZY = {
    "generate_kpts_single": "preprocess_utils", "sample_feat_by_coord": "preprocess_utils",
    "mnn_matcher": "preprocess_utils", "normalize_coords": "preprocess_utils",
    "denormalize_coords": "preprocess_utils", "detect_topk": "preprocess_utils",
    "mnn_match": "preprocess_utils", "sample_l2norm": "preprocess_utils",
    "mutual_nn_matcher": "matchers",
    "get_expected_correspondence_locs": "preprocess", "compute_prob": "preprocess",
    "get_expected_correspondence_within_window": "preprocess", "normalize_scale_channels_last": "preprocess",
    "Preprocess_Line2Window": "preprocess",
    "DiskLoss": "kploss", "EpipolarLoss_full": "epipolarloss", "GradAllReducer": "dist",
    "process": "extractor", "save_desc": "extractor", "FeatureExtractor": "extractor",
    "AsyncDescWriter": "extractor", "generate_kpts_single_noavg": "preprocess_utils",
    "ratio_matcher": "matchers", "mutual_nn_ratio_matcher": "matchers", "GraphedMatcher": "matchers",
    "PairPipeline": "pairs", "GraphedPairPipeline": "pairs",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod = importlib.import_module("." + _LAZY[name], __name__)
        return getattr(mod, name)
    raise AttributeError(name)


def install(putils_module, matchers_module=None):
    """Monkey-patch a reference checkout's modules so its own managers and
    evaluation scripts call the B200 kernels."""
    from . import matchers, preprocess_utils as pu
    for name in ("generate_kpts_single", "generate_kpts_single_noavg", "sample_feat_by_coord", "mnn_matcher"):
        setattr(putils_module, name, getattr(pu, name))
    from . import preprocess as pp
    for name in ("compute_prob", "get_expected_correspondence_locs", "get_expected_correspondence_within_window"):
        setattr(putils_module, name, getattr(pp, name))
    if matchers_module is not None:
        for name in ("mutual_nn_matcher", "ratio_matcher", "mutual_nn_ratio_matcher"):
            setattr(matchers_module, name, getattr(matchers, name))
    return putils_module
