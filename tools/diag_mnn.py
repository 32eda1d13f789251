"""Diagnostics of the matches-only matcher (POSFEAT_MNN_DEBUG counters) on the bench shape, the C1 shape and the sweep inputs."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["POSFEAT_MNN_DEBUG"] = "1"
import torch
import bench
from posfeat_b200.pairs import PairPipeline
from posfeat_b200.preprocess_utils import mnn_match

dev = torch.device("cuda", 0)
score, fmap = bench.synth_pairs(4, 1234, dev)
pipe = PairPipeline(bench.DET_CFG)
print("bench shape, 4 pairs", file=sys.stderr)
pipe.run(score, fmap)
score, fmap = bench.synth_pairs(1, 4321, dev, h=480, w=640)
print("c1", file=sys.stderr)
PairPipeline(bench.C1_CFG).run(score, fmap)
for n in (4096, 16384, 65536):
    g = torch.Generator().manual_seed(n)
    a = torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=1)
    b = torch.nn.functional.normalize(a[torch.randperm(n, generator=g)] + 0.06 * torch.randn(n, 128, generator=g), dim=1)
    print("sweep", n, file=sys.stderr)
    mnn_match(a.to(dev), b.to(dev), algo=2, want_nn21=False)
    b = torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=1)
    print("random", n, file=sys.stderr)
    mnn_match(a.to(dev), b.to(dev), algo=2, want_nn21=False)
