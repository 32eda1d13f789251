"""Small driver for ncu captures: a few passes of the pair pipeline (P pairs of
896x1200 maps, 8192 keypoints) so every kernel of the hot path launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from posfeat_b200.pairs import PairPipeline

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.cuda.set_device(0)
score, fmap = bench.synth_pairs(P, 1234, "cuda")
pipe = PairPipeline(bench.DET_CFG)
host = len(sys.argv) > 3 and sys.argv[3] == "host"     # host-buffer entry: adds mark_taps / fetch_pixels (pinned-memory staging)
if host:
    score_h, fmap_h = score.cpu().pin_memory(), fmap.cpu().pin_memory()
for _ in range(iters):
    if host:
        kpt, matches, nm = pipe.run_host(score_h, fmap_h)
    else:
        feats, matches, nm = pipe.run(score, fmap)
torch.cuda.synchronize()
print("pairs", P, "host" if host else "device", "matches", nm.tolist())
