"""Counters and kernel times of the matcher on an Aachen-shaped UNRELATED image pair whose descriptors come from the
real pipeline (bilinear samples of neighbouring keypoints share taps, so descriptors within an image are correlated)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from posfeat_b200 import _lib
from posfeat_b200.pairs import PairPipeline

cfg = dict(bench.AACHEN_CFG, num_pts=int(sys.argv[1]) if len(sys.argv) > 1 else 8192)
score, fmap = bench.synth_pairs(2, 777, "cuda", h=1056, w=1600)
pipe = PairPipeline(cfg)
f = pipe.extract(score, fmap)
desc = f["desc"]
n = desc.shape[1]
for name, (i, j) in (("related (0,1)", (0, 1)), ("unrelated (0,2)", (0, 2)), ("unrelated (1,3)", (1, 3))):
    da, db = desc[i:i + 1].expand(4, n, 128), desc[j:j + 1].expand(4, n, 128)
    os.environ.pop("POSFEAT_MNN_DEBUG", None)
    for _ in range(2):
        pipe.match_views(da, db)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(3):
        m, nm = pipe.match_views(da, db)
    torch.cuda.synchronize()
    pr = _lib.profile_read()
    _lib.profile_enable(False)
    print(name, {k: round(1e3 * v[0] / v[1] / 4, 1) for k, v in pr.items()}, "us per pair; matches", int(nm[0]), flush=True)
    os.environ["POSFEAT_MNN_DEBUG"] = "1"
    pipe.match_views(da, db)
    torch.cuda.synchronize()

# one query against 20 other images (the shape of bench.py's aachen workload)
score, fmap = bench.synth_pairs(11, 777, "cuda", h=1056, w=1600)
f = pipe.extract(score, fmap)
desc = f["desc"]
n = desc.shape[1]
da, db = desc[0].unsqueeze(0).expand(20, n, 128), desc[1:21]
os.environ.pop("POSFEAT_MNN_DEBUG", None)
for _ in range(2):
    pipe.match_views(da, db)
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(3):
    m, nm = pipe.match_views(da, db)
torch.cuda.synchronize()
pr = _lib.profile_read()
_lib.profile_enable(False)
print("query vs 20", {k: round(1e3 * v[0] / v[1] / 20, 1) for k, v in pr.items()}, "us per pair; matches", nm.tolist(), flush=True)
os.environ["POSFEAT_MNN_DEBUG"] = "1"
pipe.match_views(da, db)
torch.cuda.synchronize()
