"""Times the NMS/candidate kernel alone (library event slots) for a few threshold settings:
thr=0.9 is the bench workload, thr=100 leaves no survivors (pure read + compare)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from posfeat_b200 import _lib
from posfeat_b200.preprocess_utils import detect_topk

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
H, W = 896, 1200
torch.manual_seed(0)
score = torch.nn.functional.softplus(torch.randn(B, 1, H, W, device="cuda"))
for thr in (0.9, 1.5, 2.5, 100.0):
    for _ in range(3):
        r = detect_topk(score, 1, 8192, True, thr, "abs", sync=False)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(10):
        r = detect_topk(score, 1, 8192, True, thr, "abs", sync=False)
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    ms = prof["nms_candidates"][0] / prof["nms_candidates"][1]
    gbs = 4.0 * B * H * W / (ms * 1e-3) / 1e9
    print(f"thr {thr:6.1f}  survivors/img {float(r['counts'].float().mean()):9.0f}  nms {ms * 1e3:7.1f} us  {gbs:7.0f} GB/s  "
          f"select {prof['select_topk'][0] / prof['select_topk'][1] * 1e3:7.1f} us")
