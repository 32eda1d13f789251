"""Fused post-backbone pair pipeline: detect -> sample -> mutual-NN for batches
of image pairs, everything resident on one GPU (SURVEY.md section 7 step 4 and
section 8e: pairs are independent, so ranks shard the pair list with no
collective).

The stages are exactly the calls the reference makes per image and per pair
(managers/extractor.py:342-346 then evaluations/hpatches/evaluation.py:64-67),
minus the .npz round trip between them.
"""
from __future__ import annotations

import torch

from . import _lib
from ._runtime import check, lib, stream_ptr, workspace
from .preprocess_utils import MIN_PTS, denormalize_coords, detect_topk, sample_l2norm


def shard(items, rank: int, world: int):
    """Static round-robin sharding of an image / pair list (no communication)."""
    return list(items)[rank::world]


class PairPipeline:
    def __init__(self, detector_config: dict, normalize: bool = True, mnn_algo: int = _lib.MNN_AUTO):
        cfg = dict(detector_config)
        if not cfg.get("stable", True):
            raise NotImplementedError("stable=False is out of scope")
        self.cfg = dict(nms_radius=cfg["nms_radius"], num_pts=cfg.get("num_pts", False),
                        use_nms=cfg.get("use_nms", True), thr=cfg.get("thr", False),
                        thr_mod=cfg.get("thr_mod", "mean"))
        self.normalize = normalize
        self.mnn_algo = mnn_algo
        self._host = None

    # -- per-image stage -------------------------------------------------
    def extract(self, score: torch.Tensor, fmap: torch.Tensor):
        """score [b,1,H,W], fmap [b,D,H/4,W/4] (device) -> dict with kps_n [b,n,2],
        kpt [b,n,2] pixel xy, kp_score [b,n], desc [b,n,D], idx [b,n]."""
        r = detect_topk(score, sync=True, **self.cfg)
        n = r["n"]
        kps = r["kps"][:, :n]
        desc = sample_l2norm(fmap, kps.contiguous(), self.normalize)
        h, w = score.shape[2:]
        return {"kps_n": kps, "kpt": denormalize_coords(kps, h, w), "kp_score": r["score"][:, :n],
                "desc": desc, "idx": r["idx"][:, :n], "n": n}

    # -- per-pair stage --------------------------------------------------
    def match(self, desc: torch.Tensor):
        """desc [2P,n,D]: images (2i, 2i+1) form pair i.  Returns matches
        [P,n,2] int64 and n_matches [P] int32 (device)."""
        L = lib()
        b, n, D = desc.shape
        P = b // 2
        dev = desc.device
        matches = torch.empty((P, n, 2), dtype=torch.int64, device=dev)
        nm = torch.empty(P, dtype=torch.int32, device=dev)
        nn12 = torch.empty((P, n), dtype=torch.int32, device=dev)
        # nn21 is not requested: the tensor-core matcher then computes one direction and verifies
        # mutuality by a column scan; the exact SIMT matcher (small sizes / algo=1) needs the buffer
        use_simt = self.mnn_algo == _lib.MNN_SIMT or D != 128 or (self.mnn_algo == _lib.MNN_AUTO and n * n < 1024 * 1024)
        nn21 = torch.empty((P, n), dtype=torch.int32, device=dev) if use_simt else None
        da, db = desc[0::2], desc[1::2]          # strided views: pair stride = 2 images
        with torch.cuda.device(dev):
            ws_bytes = L.posfeat_mnn_batched_workspace_bytes(P, n, n, D, self.mnn_algo)
            ws = workspace("mnn", ws_bytes, dev)
            check(L.posfeat_mnn_batched_f32(da.data_ptr(), da.stride(0), n, da.stride(1), db.data_ptr(),
                                            db.stride(0), n, db.stride(1), D, P, self.mnn_algo,
                                            nn12.data_ptr(), 0 if nn21 is None else nn21.data_ptr(), matches.data_ptr(), nm.data_ptr(),
                                            ws.data_ptr(), ws.numel(), stream_ptr(dev)))
        return matches, nm

    def run(self, score: torch.Tensor, fmap: torch.Tensor):
        """Whole path for 2P images -> (features dict, matches, n_matches), on device."""
        feats = self.extract(score, fmap)
        matches, nm = self.match(feats["desc"])
        return feats, matches, nm

    # -- host-buffer entry (what a caller holding CPU tensors uses) --------
    def run_host(self, score_host: torch.Tensor, fmap_host: torch.Tensor):
        """Inputs in (pinned) host memory; returns host tensors: kpt [2P,n,2],
        matches [P,n,2], n_matches [P].  Copies are part of the call."""
        dev = torch.device("cuda", torch.cuda.current_device())
        key = (tuple(score_host.shape), tuple(fmap_host.shape), tuple(fmap_host.stride()))
        if self._host is None or self._host[0] != key:
            self._host = (key, torch.empty_like(score_host, device=dev), torch.empty_like(fmap_host, device=dev))
        _, s_dev, f_dev = self._host
        s_dev.copy_(score_host, non_blocking=True)
        f_dev.copy_(fmap_host, non_blocking=True)
        feats, matches, nm = self.run(s_dev, f_dev)
        out = (feats["kpt"].to("cpu", non_blocking=True), matches.to("cpu", non_blocking=True),
               nm.to("cpu", non_blocking=True))
        torch.cuda.current_stream().synchronize()
        return out

    @staticmethod
    def h2d_bytes(score, fmap):
        return score.numel() * 4 + fmap.numel() * 4

    @staticmethod
    def d2h_bytes(n_images, n, P):
        return n_images * n * 2 * 4 + P * n * 2 * 8 + P * 4
