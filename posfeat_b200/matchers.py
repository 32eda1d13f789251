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


class GraphedMatcher:
    """mnn_matcher for a FIXED pair of shapes, captured once as a CUDA graph: the five launches of a matches-only
    call (operand preparation, tensor-core contraction, rescoring, verification, compaction) replay as one
    submission.  Below ~8k x 8k descriptors a call is bound by launch latency, not by the kernels; this is the form
    for loops over many small pairs (SURVEY 8(a) C1 / the HPatches loop with top_k).

        gm = GraphedMatcher(4096, 4096)
        matches = gm(desc_a, desc_b)                 # (K,2) int64 ndarray, same as mnn_matcher(desc_a, desc_b)
        m, nm, nn12 = gm.match_device(desc_a, desc_b)  # device tensors of gm: valid until the next call

    Inputs are copied into the graph's static buffers (device-to-device, or host-to-device for CPU tensors)."""

    def __init__(self, n_a: int, n_b: int, dim: int = 128, algo=None, device=None):
        from . import _lib
        from ._runtime import require_cuda, workspace_scope
        from .preprocess_utils import mnn_match
        require_cuda()
        if n_a <= 0 or n_b <= 0:
            raise IndexError("GraphedMatcher: empty descriptor set")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.algo = _lib.MNN_AUTO if algo is None else algo
        self.a = torch.zeros((n_a, dim), dtype=torch.float32, device=dev)
        self.b = torch.zeros((n_b, dim), dtype=torch.float32, device=dev)
        self.a[:, 0] = 1.0
        self.b[:, 0] = 1.0
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        self._scratch = workspace_scope()         # the graph's scratch memory is its own, for as long as it lives
        with self._scratch:
            with torch.cuda.stream(side):
                for _ in range(2):                # warm-up on the capture stream: workspace, function attributes
                    mnn_match(self.a, self.b, self.algo, want_nn21=False)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self.matches, self.nm, self.nn12, _ = mnn_match(self.a, self.b, self.algo, want_nn21=False)

    def match_device(self, descriptors1, descriptors2):
        if tuple(descriptors1.shape) != tuple(self.a.shape) or tuple(descriptors2.shape) != tuple(self.b.shape):
            raise ValueError(f"GraphedMatcher was built for {tuple(self.a.shape)} x {tuple(self.b.shape)}, "
                             f"got {tuple(descriptors1.shape)} x {tuple(descriptors2.shape)}")
        self.a.copy_(descriptors1, non_blocking=True)
        self.b.copy_(descriptors2, non_blocking=True)
        self.graph.replay()
        return self.matches, self.nm, self.nn12

    def __call__(self, descriptors1, descriptors2):
        m, nm, _ = self.match_device(descriptors1, descriptors2)
        return m[:int(nm.item())].cpu().numpy()
