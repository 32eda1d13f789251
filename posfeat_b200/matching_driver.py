"""Multi-GPU matching over real feature lists: the loops of the reference's evaluation scripts, sharded.

  evaluations/hpatches/evaluation.py:40-90      per sequence: image 1 against images 2..6, mutual NN, homography error
  evaluations/aachen/reconstruct_pipeline.py:182-221 / evaluations/ETH_local_feature/reconstruction_pipeline.py:98-131
                                                  per line "name1 name2" of a match list: load two .npz, mutual NN

Pairs are independent: every rank takes whole GROUPS of the pair list -- an HPatches sequence, an
Aachen query with its retrieved database images -- so that the image the group shares is read and uploaded once;
there is no collective on the data path, only the per-rank summaries travel to rank 0 over the host at the end.
Descriptor files are the ones ``save_desc`` / ``AsyncDescWriter`` write (np.savez with keypoints / scores /
descriptors, ``<root>/<image name>.<method>``).  A small device-side cache keeps the descriptors of recently used
images resident, so an image that appears in many pairs crosses the host link once.
"""
from __future__ import annotations

import collections
import os
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .dist import gather_objects, shard_by_group

Pair = Tuple[str, str]


def read_features(root: str, name: str, method: str, top_k: Optional[int] = None):
    """(keypoints [n,2+], descriptors [n,D]) of ``<root>/<name>.<method>``; ``top_k`` keeps the best-scoring
    keypoints like generate_read_function (evaluations/hpatches/evaluation.py:106-115)."""
    with np.load(os.path.join(root, f"{name}.{method}")) as z:
        kp, desc = z["keypoints"], z["descriptors"]
        if top_k is not None:
            if "scores" not in z:
                raise KeyError("top_k needs a 'scores' array in the feature file")
            ids = np.argsort(z["scores"].reshape(-1))[-top_k:]
            kp, desc = kp[ids], desc[ids]
    return kp, desc


def hpatches_pairs(seq_names: Iterable[str], extension: str = "ppm") -> List[Pair]:
    """The reference -> target pairs of the HPatches protocol: image 1 against images 2..6 of every sequence."""
    return [(f"{s}/1.{extension}", f"{s}/{k}.{extension}") for s in seq_names for k in range(2, 7)]


def read_pair_list(path: str) -> List[Pair]:
    """``name1 name2`` per line (aachen/reconstruct_pipeline.py:195-197)."""
    out = []
    with open(path) as f:
        for line in f:
            parts = line.strip("\n").split(" ")
            if len(parts) >= 2 and parts[0]:
                out.append((parts[0], parts[1]))
    return out


class DescriptorCache:
    """Descriptors of the most recently used images, resident on the device (LRU over a byte budget)."""

    def __init__(self, root: str, method: str, device, budget_bytes: int = 8 << 30, top_k: Optional[int] = None,
                 max_kpts: Optional[int] = None):
        self.root, self.method, self.device = root, method, device
        self.budget, self.used = int(budget_bytes), 0
        self.top_k, self.max_kpts = top_k, max_kpts
        self.items: "collections.OrderedDict[str, tuple]" = collections.OrderedDict()
        self.loads = 0

    def get(self, name: str):
        hit = self.items.get(name)
        if hit is not None:
            self.items.move_to_end(name)
            return hit
        kp, desc = read_features(self.root, name, self.method, self.top_k)
        if self.max_kpts is not None and kp.shape[0] > self.max_kpts:          # evaluation.py:52-54: at most 60000
            kp, desc = kp[:self.max_kpts], desc[:self.max_kpts]
        t = torch.from_numpy(np.ascontiguousarray(desc, dtype=np.float32))
        if torch.device(self.device).type == "cuda":
            t = t.pin_memory().to(self.device, non_blocking=True)
        self.loads += 1
        nbytes = t.numel() * 4
        while self.items and self.used + nbytes > self.budget:
            _, (_, old) = self.items.popitem(last=False)
            self.used -= old.numel() * 4
        self.items[name] = (kp, t)
        self.used += nbytes
        return kp, t


def _default_matcher(a: torch.Tensor, b: torch.Tensor) -> np.ndarray:
    from .preprocess_utils import mnn_matcher
    return mnn_matcher(a, b)


def match_pairs(pairs: Sequence[Pair], features_root: str, method: str, rank: int = 0, world: int = 1,
                group_key: Optional[Callable[[Pair], str]] = None, device=None, matcher: Optional[Callable] = None,
                top_k: Optional[int] = None, max_kpts: Optional[int] = 60000, cache_bytes: int = 8 << 30,
                on_result: Optional[Callable[[Pair, np.ndarray, np.ndarray, np.ndarray], None]] = None) -> Dict[Pair, np.ndarray]:
    """Mutual-NN matches of this rank's share of ``pairs`` (whole groups; default group = first image of the pair,
    i.e. the HPatches reference image / the Aachen query).  Returns {pair: (K,2) int64 matches}; ``on_result`` is
    called per pair with (pair, matches, keypoints1, keypoints2) -- e.g. to score or to write a database row."""
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("posfeat_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
    matcher = matcher or _default_matcher
    mine = shard_by_group(pairs, group_key or (lambda p: p[0]), rank, world)
    cache = DescriptorCache(features_root, method, device, cache_bytes, top_k, max_kpts)
    out: Dict[Pair, np.ndarray] = {}
    for pair in mine:
        kp1, d1 = cache.get(pair[0])
        kp2, d2 = cache.get(pair[1])
        m = matcher(d1, d2)
        out[pair] = m
        if on_result is not None:
            on_result(pair, m, kp1, kp2)
    match_pairs.last_cache_loads = cache.loads
    return out


def homography_errors(kp_a: np.ndarray, kp_b: np.ndarray, matches: np.ndarray, homography: np.ndarray) -> np.ndarray:
    """Reprojection distance of every match under the ground-truth homography (evaluation.py:73-82), float64."""
    if matches.shape[0] == 0:
        return np.array([float("inf")])
    pa = np.concatenate([kp_a[matches[:, 0], :2], np.ones([matches.shape[0], 1])], axis=1)
    proj = (homography @ pa.T).T
    proj = proj[:, :2] / proj[:, 2:]
    return np.sqrt(np.sum((kp_b[matches[:, 1], :2] - proj) ** 2, axis=1))


def homography_hits(kp_a, kp_b, matches, n_matches, homography, thresholds) -> "torch.Tensor":
    """The scoring of one pair (evaluation.py:73-90) on tensors, in float64, on whatever device they live on:
    ``mean(dist <= t)`` for every threshold t, where dist is the reprojection distance of the first ``n_matches`` rows of
    ``matches`` under the ground-truth homography.  With the matches already on the device (``mnn_match``) only
    ``len(thresholds)`` numbers per pair cross the host link instead of the (K, 2) int64 match list.  No matches: zeros
    (the reference's ``dist = [inf]``)."""
    dev = matches.device
    k = int(n_matches)
    thr = torch.as_tensor(list(thresholds), dtype=torch.float64, device=dev)
    if k == 0:
        return torch.zeros_like(thr)
    Hm = torch.as_tensor(np.asarray(homography, dtype=np.float64), device=dev)
    a = torch.as_tensor(kp_a, device=dev)[matches[:k, 0], :2].to(torch.float64)
    b = torch.as_tensor(kp_b, device=dev)[matches[:k, 1], :2].to(torch.float64)
    x, y = a[:, 0], a[:, 1]
    px = Hm[0, 0] * x + Hm[0, 1] * y + Hm[0, 2]
    py = Hm[1, 0] * x + Hm[1, 1] * y + Hm[1, 2]
    pw = Hm[2, 0] * x + Hm[2, 1] * y + Hm[2, 2]
    dist = torch.sqrt((b[:, 0] - px / pw) ** 2 + (b[:, 1] - py / pw) ** 2)
    return (dist[:, None] <= thr[None, :]).to(torch.float64).mean(0)


def hpatches_benchmark(seq_names: Sequence[str], features_root: str, method: str, homographies: Callable[[str, int], np.ndarray],
                       rank: int = 0, world: int = 1, thresholds: Sequence[int] = tuple(range(1, 16)), extension: str = "ppm",
                       matcher: Optional[Callable] = None, device=None, top_k: Optional[int] = None,
                       score_on_device: bool = False):
    """benchmark_features (evaluation.py:40-96) sharded by sequence: returns on rank 0 the reference's
    (i_err, v_err, [seq_type, n_feats, n_matches]) accumulated over all ranks (None on the other ranks).
    ``homographies(seq, k)`` supplies H_1_k (the reference reads ``<dataset>/<seq>/H_1_<k>``).
    ``score_on_device``: match with ``mnn_match`` and score with ``homography_hits`` on the device -- the match list
    never leaves it, 15 numbers and a count per pair do."""
    pairs = hpatches_pairs(seq_names, extension)
    acc = {"i": {t: 0.0 for t in thresholds}, "v": {t: 0.0 for t in thresholds}}
    rows = []
    seen_ref = set()

    def score(pair, m, kp1, kp2):
        seq, k = pair[0].split("/")[0], int(pair[1].split("/")[1].split(".")[0])
        dist = homography_errors(kp1, kp2, m, homographies(seq, k))
        kind = "i" if seq[0] == "i" else "v"
        for t in thresholds:
            acc[kind][t] += float(np.mean(dist <= t))
        feats = [kp2.shape[0]]
        if seq not in seen_ref:
            seen_ref.add(seq)
            feats.insert(0, kp1.shape[0])
        rows.append((seq, k, seq[0], feats, int(m.shape[0])))

    if score_on_device:
        from .preprocess_utils import mnn_match

        def device_matcher(a, b):
            m, nm, _, _ = mnn_match(a, b, want_nn21=False)
            return m, int(nm.item())

        def score_dev(pair, res, kp1, kp2):
            m, k = res
            seq, kk = pair[0].split("/")[0], int(pair[1].split("/")[1].split(".")[0])
            hits = homography_hits(kp1, kp2, m, k, homographies(seq, kk), thresholds).cpu().numpy()
            kind = "i" if seq[0] == "i" else "v"
            for t, h in zip(thresholds, hits):
                acc[kind][t] += float(h)
            feats = [kp2.shape[0]]
            if seq not in seen_ref:
                seen_ref.add(seq)
                feats.insert(0, kp1.shape[0])
            rows.append((seq, kk, seq[0], feats, k))

        match_pairs(pairs, features_root, method, rank, world, group_key=lambda p: p[0].split("/")[0], device=device,
                    matcher=matcher or device_matcher, top_k=top_k, on_result=score_dev)
    else:
        match_pairs(pairs, features_root, method, rank, world, group_key=lambda p: p[0].split("/")[0], device=device,
                    matcher=matcher, top_k=top_k, on_result=score)
    parts = gather_objects((acc, rows))
    if parts is None:
        return None
    i_err = {t: sum(p[0]["i"][t] for p in parts) for t in thresholds}
    v_err = {t: sum(p[0]["v"][t] for p in parts) for t in thresholds}
    allrows = sorted(r for p in parts for r in p[1])            # sequence order, as the reference's loop
    seq_type = np.array([r[2] for r in allrows])
    n_feats = np.array([f for r in allrows for f in r[3]])
    n_matches = np.array([r[4] for r in allrows])
    return i_err, v_err, [seq_type, n_feats, n_matches]
