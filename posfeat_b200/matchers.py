"""Matchers with the names the evaluation scripts import
(evaluations/aachen/matchers.py, evaluations/ETH_local_feature/custom_matcher.py,
evaluations/hpatches/evaluation.py:27)."""
from .preprocess_utils import mnn_matcher


def mutual_nn_matcher(descriptors1, descriptors2, **kw):
    """evaluations/aachen/matchers.py:5-13 -- identical maths to mnn_matcher."""
    return mnn_matcher(descriptors1, descriptors2)


def ratio_matcher(descriptors1, descriptors2, ratio=0.95):
    raise NotImplementedError("ratio_matcher is a 'next' row (SURVEY.md section 8f-1)")


def mutual_nn_ratio_matcher(descriptors1, descriptors2, ratio=0.95):
    raise NotImplementedError("mutual_nn_ratio_matcher is a 'next' row (SURVEY.md section 8f-1)")
