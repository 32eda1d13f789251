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
from ._runtime import check, lib, map_ptr, require_cuda, stream_ptr, workspace, workspace_scope
from .preprocess_utils import MIN_PTS, denormalize_coords, detect_finish, detect_topk, sample_l2norm


from .dist import shard  # noqa: E402,F401  (kept importable from here: static round-robin sharding, no communication)


class PairPipeline:
    def __init__(self, detector_config: dict, normalize: bool = True, mnn_algo: int = _lib.MNN_AUTO, streams: int = 1):
        cfg = dict(detector_config)
        if not cfg.get("stable", True):
            raise NotImplementedError("stable=False is out of scope")
        self.cfg = dict(nms_radius=cfg["nms_radius"], num_pts=cfg.get("num_pts", False),
                        use_nms=cfg.get("use_nms", True), thr=cfg.get("thr", False),
                        thr_mod=cfg.get("thr_mod", "mean"))
        self.normalize = normalize
        self.mnn_algo = mnn_algo
        self.streams = int(streams)          # >1: batches are split over CUDA streams (see run)
        self._side = None
        self._host = None
        self._hstreams = None
        self._fetch_ws = []

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
    def _tc_applies(self, n, D):
        return D == 128 and self.mnn_algo in (_lib.MNN_AUTO, _lib.MNN_TC) and (self.mnn_algo == _lib.MNN_TC or n * n >= 1024 * 1024)

    def sample_for_pairs(self, fmap, kps, ws_key: str = "mnn", out=None):
        """Sampler for the pair path: when the tensor-core matcher will run, one kernel writes the
        descriptors AND the matcher's bf16 operands / norms into its workspace (posfeat_sample_pairs_f32),
        so the matcher skips its own pass over the descriptors.  Returns (desc, prepared).
        ``fmap`` may live in pinned host memory (see run_host): the kernel then gathers over the host link."""
        b, D, h, w = fmap.shape
        n = kps.shape[1]
        fused = (b % 2 == 0 and n >= 1 and self._tc_applies(n, D) and fmap.dtype == torch.float32 and
                 fmap.is_contiguous(memory_format=torch.channels_last) and fmap.data_ptr() % 16 == 0)
        if not fused:
            return sample_l2norm(fmap, kps, self.normalize, out=out), False
        L = lib()
        dev = kps.device
        kps = kps.contiguous()
        if out is None:
            out = torch.empty((b, n, D), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = workspace(ws_key, L.posfeat_mnn_batched_workspace_bytes(b // 2, n, n, D, _lib.MNN_TC), dev)
            check(L.posfeat_sample_pairs_f32(map_ptr(fmap), b, D, h, w, fmap.stride(0), fmap.stride(1), fmap.stride(2),
                                             fmap.stride(3), kps.data_ptr(), n, int(bool(self.normalize)), out.data_ptr(),
                                             ws.data_ptr(), ws.numel(), stream_ptr(dev)))
        return out, True

    def match(self, desc: torch.Tensor, ws_key: str = "mnn", prepared: bool = False):
        """desc [2P,n,D]: images (2i, 2i+1) form pair i.  Returns matches
        [P,n,2] int64 and n_matches [P] int32 (device).  ``prepared``: the workspace ``ws_key`` already
        holds the operands written by sample_for_pairs."""
        return self.match_views(desc[0::2], desc[1::2], ws_key, prepared)      # strided views: pair stride = 2 images

    def match_views(self, da: torch.Tensor, db: torch.Tensor, ws_key: str = "mnn", prepared: bool = False):
        """Batched matcher over P pairs given as two [P,n,D] views with uniform pair strides (any stride,
        including 0: ``desc[q].expand(K, n, D)`` matches one query image against K database images, the shape
        of evaluations/aachen/reconstruct_pipeline.py:182-221).  Returns (matches [P,n,2] int64, n_matches [P])."""
        L = lib()
        P, n, D = da.shape
        if tuple(db.shape) != (P, n, D) or da.stride(2) != 1 or db.stride(2) != 1:
            raise ValueError("match_views: two [P,n,D] views with unit channel stride and equal shapes are required")
        dev = da.device
        matches = torch.empty((P, n, 2), dtype=torch.int64, device=dev)
        nm = torch.empty(P, dtype=torch.int32, device=dev)
        nn12 = torch.empty((P, n), dtype=torch.int32, device=dev)
        # nn21 is not requested: the tensor-core matcher then computes one direction and verifies
        # mutuality per column chunk; the exact SIMT matcher (small sizes / algo=1) needs the buffer
        use_simt = not self._tc_applies(n, D)
        algo = (_lib.MNN_TC | _lib.MNN_PREPARED) if prepared else self.mnn_algo
        nn21 = torch.empty((P, n), dtype=torch.int32, device=dev) if use_simt else None
        with torch.cuda.device(dev):
            ws_bytes = L.posfeat_mnn_batched_workspace_bytes(P, n, n, D, _lib.MNN_TC if prepared else self.mnn_algo)
            ws = workspace(ws_key, ws_bytes, dev)
            check(L.posfeat_mnn_batched_f32(da.data_ptr(), da.stride(0), n, da.stride(1), db.data_ptr(),
                                            db.stride(0), n, db.stride(1), D, P, algo,
                                            nn12.data_ptr(), 0 if nn21 is None else nn21.data_ptr(), matches.data_ptr(), nm.data_ptr(),
                                            ws.data_ptr(), ws.numel(), stream_ptr(dev)))
        return matches, nm

    def stage_taps(self, fmap_host: torch.Tensor, fmap_dev: torch.Tensor, kps: torch.Tensor):
        """Sparse host->device copy of exactly the descriptor-map pixels the sampler will read for
        ``kps`` (posfeat_fetch_taps_f32): every pixel under some keypoint's 2x2 tap block crosses the host
        link once, into its place in ``fmap_dev`` (same shape and strides as the pinned ``fmap_host``)."""
        b, D, h, w = fmap_host.shape
        if tuple(fmap_dev.shape) != (b, D, h, w) or fmap_dev.stride() != fmap_host.stride():
            raise ValueError("fmap_dev must have the shape and strides of fmap_host")
        L = lib()
        dev = fmap_dev.device
        kps = kps.contiguous()
        with torch.cuda.device(dev):
            ws = workspace("fetch", L.posfeat_fetch_taps_workspace_bytes(b, h, w), dev)
            check(L.posfeat_fetch_taps_f32(map_ptr(fmap_host), fmap_dev.data_ptr(), b, D, h, w, fmap_host.stride(0),
                                           fmap_host.stride(1), fmap_host.stride(2), fmap_host.stride(3), kps.data_ptr(),
                                           kps.shape[1], ws.data_ptr(), ws.numel(), stream_ptr(dev)))
        self._fetch_ws.append((ws, b, h, w))
        del self._fetch_ws[:-64]                  # bounded: direct callers of run(stage_from=...) never reset it
        return ws

    def staged_pixels(self, fmap_host: torch.Tensor = None, dev=None) -> int:
        """Pixels moved by the stage_taps calls of the last run_host / run (synchronises)."""
        import ctypes as C
        total = 0
        for ws, b, h, w in self._fetch_ws:
            out = C.c_uint64(0)
            with torch.cuda.device(ws.device):
                torch.cuda.synchronize(ws.device)
                check(lib().posfeat_fetch_taps_count(ws.data_ptr(), b, h, w, C.byref(out), stream_ptr(ws.device)))
            total += int(out.value)
        return total

    def run(self, score: torch.Tensor, fmap: torch.Tensor, stage_from: torch.Tensor = None):
        """Whole path for 2P images -> (features dict, matches, n_matches), on device.

        ``stage_from``: pinned host descriptor map; ``fmap`` is then an uninitialised device buffer of the
        same shape/strides that receives only the pixels the sampler needs (stage_taps) once the keypoints
        are known.

        With ``streams > 1`` and a fixed ``num_pts`` the batch is cut into that many groups of pairs,
        each queued on its own stream with no host round trip in between: the keypoint count n stays on
        the device and the group is processed at n = num_pts, which is what the detector returns whenever
        every image has that many survivors.  The latency-bound selection kernel of one group then runs
        under the sampler / matcher of another.  n is read once after everything is queued; if an image
        came up short (n < num_pts) the batch is redone on the plain path.  Measured gain on a B200 at
        64 pairs per step: +1.7 % with two streams, negative with more -- the default stays 1."""
        P = score.shape[0] // 2
        num_pts = self.cfg["num_pts"]
        if stage_from is None and self.streams > 1 and num_pts and num_pts >= MIN_PTS and P >= 2 * self.streams:
            out = self._run_streams(score, fmap)
            if out is not None:
                return out
        r = detect_topk(score, sync=False, **self.cfg)
        # while the selection kernel runs: everything that only needs the capacity, not n itself
        cap, D = r["cap"], fmap.shape[1]
        desc_buf = torch.empty((score.shape[0], cap, D), dtype=torch.float32, device=score.device)
        if P >= 1 and self._tc_applies(cap, D):
            with torch.cuda.device(score.device):
                workspace("mnn", lib().posfeat_mnn_batched_workspace_bytes(P, cap, cap, D, _lib.MNN_TC), score.device)
        n = detect_finish(r)                      # the one host round trip of the step
        kps = r["kps"][:, :n]
        out = desc_buf if n == cap else desc_buf.view(-1)[:score.shape[0] * n * D].view(score.shape[0], n, D)
        if stage_from is not None:
            self.stage_taps(stage_from, fmap, kps)
        desc, prepared = self.sample_for_pairs(fmap, kps, out=out)
        h, w = score.shape[2:]
        feats = {"kps_n": kps, "kpt": denormalize_coords(kps, h, w), "kp_score": r["score"][:, :n],
                 "desc": desc, "idx": r["idx"][:, :n], "n": n}
        matches, nm = self.match(desc, prepared=prepared)
        return feats, matches, nm

    def run_nosync(self, score: torch.Tensor, fmap: torch.Tensor):
        """The whole path queued WITHOUT a host round trip: everything runs at n = num_pts (what the detector
        returns whenever every image has that many survivors); the true n stays on the device.  Returns
        (result, feats, matches, n_matches) where ``result`` is the detect_topk record to pass to
        ``detect_finish`` once the caller wants n and the device status -- if that n differs from num_pts the
        outputs are void and the call has to be redone with ``run``.  Needs a fixed ``num_pts`` >= 128.  This is
        the form that can be captured in a CUDA graph (see GraphedPairPipeline)."""
        num_pts = self.cfg["num_pts"]
        if not num_pts or num_pts < MIN_PTS:
            raise ValueError("run_nosync needs a fixed num_pts >= 128 (with num_pts=False the count decides the shapes)")
        r = detect_topk(score, sync=False, **self.cfg)
        kps = r["kps"]                                              # [b, cap, 2]; all rows valid iff n == cap
        desc, prepared = self.sample_for_pairs(fmap, kps)
        matches, nm = self.match(desc, prepared=prepared)
        h, w = score.shape[2:]
        feats = {"kps_n": kps, "kpt": denormalize_coords(kps, h, w), "kp_score": r["score"], "desc": desc,
                 "idx": r["idx"], "n": r["cap"]}
        return r, feats, matches, nm

    def _run_streams(self, score, fmap):
        dev = score.device
        G = self.streams
        if self._side is None or len(self._side) != G:
            self._side = [torch.cuda.Stream(device=dev) for _ in range(G)]
        main = torch.cuda.current_stream(dev)
        P = score.shape[0] // 2
        cap = int(self.cfg["num_pts"])
        h, w = score.shape[2:]
        bounds = [2 * ((P * g) // G) for g in range(G + 1)]          # image index ranges, whole pairs
        desc_all = torch.empty((2 * P, cap, fmap.shape[1]), dtype=torch.float32, device=dev)
        parts = []
        for g, st in enumerate(self._side):
            lo, hi = bounds[g], bounds[g + 1]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                r = detect_topk(score[lo:hi], sync=False, ws_key=f"detect{g}", **self.cfg)
                kps = r["kps"]                                          # [b, cap, 2]; all rows valid iff n == cap
                desc, prepared = self.sample_for_pairs(fmap[lo:hi], kps, ws_key=f"mnn{g}", out=desc_all[lo:hi])
                matches, nm = self.match(desc, ws_key=f"mnn{g}", prepared=prepared)
                parts.append((r, kps, desc, matches, nm))
        for st in self._side:
            main.wait_stream(st)
        n_all = torch.stack([p[0]["n"].reshape(()) for p in parts]).cpu()      # the one host read of the step
        if int(n_all.min()) != cap or int(n_all.max()) != cap:
            return None
        kps = torch.cat([p[1] for p in parts])
        feats = {"kps_n": kps, "kpt": denormalize_coords(kps, h, w), "kp_score": torch.cat([p[0]["score"] for p in parts]),
                 "desc": desc_all, "idx": torch.cat([p[0]["idx"] for p in parts]), "n": cap}
        return feats, torch.cat([p[3] for p in parts]), torch.cat([p[4] for p in parts])

    # -- host-buffer entry (what a caller holding CPU tensors uses) --------
    def run_host(self, score_host: torch.Tensor, fmap_host: torch.Tensor, gather=None, chunks=None):
        """Inputs in (pinned) host memory; returns host tensors (pinned): kpt [2P,n,2],
        matches [P,n,2], n_matches [P].  Copies are part of the call, and the host link bounds it.

        ``gather`` selects how the dense descriptor map (34.4 MB per 896x1200 image) reaches the GPU:
        * ``"stage"`` (default whenever the map is pinned and channels-last): only the score maps are
          copied whole; once the keypoints are known, exactly the map pixels under their 2x2 tap blocks
          are fetched from the pinned map, each once (stage_taps: 13.3 MB per image at 8192 keypoints);
        * ``"direct"`` / ``True``: the sampler reads its taps straight from the pinned map (16.8 MB per
          image requested: pixels shared by two keypoints cross the link twice);
        * ``False``: the whole map is copied first (the plain path).

        ``chunks``: the batch is cut into that many groups of whole pairs, each on its own stream (default 2 for
        batches of 8 pairs or more with a fixed ``num_pts``): one group's selection / sampling / matching and its
        device->host result copy run under the next group's host->device traffic, and the host's wait for a
        group's keypoint count no longer idles the link.  The detector couples the images of one call through
        ``n = min`` over the batch (losses/preprocess_utils.py:251-259); if a group comes back with fewer than
        ``num_pts`` keypoints the call is redone unsplit, so the result never depends on ``chunks``."""
        if gather is None:
            gather = "stage" if self.host_gather_applies(fmap_host) else False
        elif gather is True:
            gather = "direct"
        if gather not in ("stage", "direct", False):
            raise ValueError(f"gather must be 'stage', 'direct', True, False or None, got {gather!r}")
        if gather and not self.host_gather_applies(fmap_host):
            raise ValueError("host gather needs a pinned, channels-last float32 descriptor map")
        require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        key = (tuple(score_host.shape), tuple(fmap_host.shape), tuple(fmap_host.stride()), gather)
        if self._host is None or self._host[0] != key:
            self._host = None                            # release the old buffers first
            self._host = (key, torch.empty_like(score_host, device=dev),
                          None if gather == "direct" else torch.empty_like(fmap_host, device=dev))
        _, s_dev, f_dev = self._host
        P = score_host.shape[0] // 2
        num_pts = self.cfg["num_pts"]
        if chunks is None:
            chunks = 2 if (P >= 8 and num_pts and num_pts >= MIN_PTS and score_host.is_pinned()) else 1
        chunks = max(1, min(int(chunks), P))
        self._fetch_ws = []
        if chunks > 1 and num_pts and num_pts >= MIN_PTS:
            out = self._run_host_chunks(score_host, fmap_host, s_dev, f_dev, gather, chunks)
            if out is not None:
                return out
            self._fetch_ws = []
        s_dev.copy_(score_host, non_blocking=True)
        feats, matches, nm = self._run_one(s_dev, f_dev, fmap_host, gather)
        pin = score_host.is_pinned()
        n = feats["n"]
        out = (torch.empty((2 * P, n, 2), dtype=torch.float32, pin_memory=pin),
               torch.empty((P, n, 2), dtype=torch.int64, pin_memory=pin), torch.empty(P, dtype=torch.int32, pin_memory=pin))
        for dst, src in zip(out, (feats["kpt"], matches, nm)):
            dst.copy_(src, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out

    def _run_one(self, s_dev, f_dev, fmap_host, gather):
        if gather == "stage":
            return self.run(s_dev, f_dev, stage_from=fmap_host)
        if gather == "direct":
            return self.run(s_dev, fmap_host)
        f_dev.copy_(fmap_host, non_blocking=True)
        return self.run(s_dev, f_dev)

    def _run_host_chunks(self, score_host, fmap_host, s_dev, f_dev, gather, chunks):
        dev = s_dev.device
        P = score_host.shape[0] // 2
        cap = int(self.cfg["num_pts"])
        if self._hstreams is None or len(self._hstreams) != chunks:
            self._hstreams = [torch.cuda.Stream(device=dev) for _ in range(chunks)]
        main = torch.cuda.current_stream(dev)
        bounds = [2 * ((P * g) // chunks) for g in range(chunks + 1)]
        pin = score_host.is_pinned()
        kpt_h = torch.empty((2 * P, cap, 2), dtype=torch.float32, pin_memory=pin)
        m_h = torch.empty((P, cap, 2), dtype=torch.int64, pin_memory=pin)
        nm_h = torch.empty(P, dtype=torch.int32, pin_memory=pin)
        # every group's score maps are queued first: the link works on group g+1 while the host waits for group g's count
        for g, st in enumerate(self._hstreams):
            lo, hi = bounds[g], bounds[g + 1]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                s_dev[lo:hi].copy_(score_host[lo:hi], non_blocking=True)
        ok = True
        for g, st in enumerate(self._hstreams):
            lo, hi = bounds[g], bounds[g + 1]
            with torch.cuda.stream(st):
                feats, matches, nm = self._run_one(s_dev[lo:hi], None if f_dev is None else f_dev[lo:hi], fmap_host[lo:hi], gather)
                if feats["n"] != cap:          # an image ran short: n couples the whole batch -> redo unsplit
                    ok = False
                    break
                kpt_h[lo:hi].copy_(feats["kpt"], non_blocking=True)
                m_h[lo // 2:hi // 2].copy_(matches, non_blocking=True)
                nm_h[lo // 2:hi // 2].copy_(nm, non_blocking=True)
        for st in self._hstreams:
            st.synchronize()
            main.wait_stream(st)
        return (kpt_h, m_h, nm_h) if ok else None

    @staticmethod
    def host_gather_applies(fmap_host):
        return (fmap_host.device.type == "cpu" and fmap_host.is_pinned() and fmap_host.dtype == torch.float32 and
                fmap_host.dim() == 4 and fmap_host.is_contiguous(memory_format=torch.channels_last))

    @staticmethod
    def h2d_bytes(score, fmap, n=None):
        """Bytes crossing the host link per call; with ``n`` (host gather) the descriptor part is the
        taps the sampler requests, 4 x D x 4 bytes per keypoint (duplicates included)."""
        if n is None:
            return score.numel() * 4 + fmap.numel() * 4
        return score.numel() * 4 + fmap.shape[0] * n * 4 * fmap.shape[1] * 4

    @staticmethod
    def d2h_bytes(n_images, n, P):
        return n_images * n * 2 * 4 + P * n * 2 * 8 + P * 4


class GraphedPairPipeline:
    """A PairPipeline call of fixed shape captured in a CUDA graph: the eight launches, two memsets and the small
    tensor operations of detect -> sample -> match replay as ONE submission.  For a single small pair (BASELINE
    config 1: 480x640, 4096 keypoints) the call is bound by launch latency and the host's wait for the keypoint
    count, not by the kernels; the graph removes both (the count is read once, after the replay).

        g = GraphedPairPipeline(PairPipeline(cfg), score.shape, fmap.shape)
        feats, matches, nm = g(score, fmap)          # tensors of g: valid until the next call

    Inputs are copied into the graph's static buffers (device-to-device).  When an image has fewer than num_pts
    survivors the replayed result is void and the call is redone on the plain path, so the result never depends
    on the graph."""

    def __init__(self, pipe: PairPipeline, score_shape, fmap_shape, channels_last: bool = True, device=None):
        require_cuda()
        self.pipe = pipe
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.score = torch.zeros(tuple(score_shape), dtype=torch.float32, device=dev)
        self.fmap = torch.zeros(tuple(fmap_shape), dtype=torch.float32, device=dev)
        if channels_last:
            self.fmap = self.fmap.contiguous(memory_format=torch.channels_last)
        self.score.fill_(1.0)
        self.score[..., ::2, ::2] = 2.0                      # a map with plenty of local maxima for the warm-up runs
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        self._scratch = workspace_scope()                     # the graph's scratch memory is its own while it lives
        with self._scratch:
            with torch.cuda.stream(side):
                for _ in range(2):                            # warm-up on the capture stream: workspaces, caches, attributes
                    pipe.run_nosync(self.score, self.fmap)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self.rec, self.feats, self.matches, self.nm = pipe.run_nosync(self.score, self.fmap)
        self.cap = self.rec["cap"]

    def __call__(self, score: torch.Tensor, fmap: torch.Tensor):
        self.score.copy_(score, non_blocking=True)
        self.fmap.copy_(fmap, non_blocking=True)
        self.graph.replay()
        n = detect_finish(self.rec)                           # the one host round trip: device status and n
        if n != self.cap:
            return self.pipe.run(score, fmap)
        return self.feats, self.matches, self.nm
