"""Matchers with the names the evaluation scripts import
(evaluations/aachen/matchers.py, evaluations/ETH_local_feature/custom_matcher.py,
evaluations/hpatches/evaluation.py:27)."""
import torch

from ._runtime import check, lib, stream_ptr, to_device, workspace
from .preprocess_utils import mnn_matcher


def mutual_nn_matcher(descriptors1, descriptors2, **kw):
    """evaluations/aachen/matchers.py:5-13 -- identical maths to mnn_matcher."""
    return mnn_matcher(descriptors1, descriptors2)


def _ratio(descriptors1, descriptors2, ratio, mutual):
    a, _ = to_device(descriptors1)
    b, _ = to_device(descriptors2)
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise ValueError(f"descriptors must be [N,D] and [M,D], got {tuple(a.shape)} {tuple(b.shape)}")
    N, D = a.shape
    M = b.shape[0]
    if N < 2 or M < 2:
        raise RuntimeError("selected index k out of range (torch.topk(sim, 2) on the reference path)")
    a = a if a.stride(1) == 1 else a.contiguous()
    b = b if b.stride(1) == 1 else b.contiguous()
    dev = a.device
    L = lib()
    nn12 = torch.empty(N, dtype=torch.int32, device=dev)
    matches = torch.empty((N, 2), dtype=torch.int64, device=dev)
    nm = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = workspace("ratio", L.posfeat_ratio_match_workspace_bytes(N, M, D), dev)
        check(L.posfeat_ratio_match_f32(a.data_ptr(), N, a.stride(0), b.data_ptr(), M, b.stride(0), D, float(ratio),
                                        int(mutual), nn12.data_ptr(), matches.data_ptr(), nm.data_ptr(),
                                        ws.data_ptr(), ws.numel(), stream_ptr(dev)))
    return matches[:int(nm.item())].cpu().numpy()


def ratio_matcher(descriptors1, descriptors2, ratio=0.95):
    """Symmetric Lowe ratio test, evaluations/aachen/matchers.py:17-43."""
    return _ratio(descriptors1, descriptors2, ratio, False)


def mutual_nn_ratio_matcher(descriptors1, descriptors2, ratio=0.95):
    """Mutual NN + symmetric ratio test, evaluations/aachen/matchers.py:47-75
    (the matcher configs/extract_ETH.yaml:40-42 selects, ratio 0.75)."""
    return _ratio(descriptors1, descriptors2, ratio, True)
