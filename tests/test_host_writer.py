"""Host-side logic that runs without a GPU: the asynchronous descriptor writer must
produce exactly the files save_desc writes (managers/extractor.py:254-271 layout)."""
import os

import numpy as np
import torch


def _fake(i, n=50, d=16):
    g = torch.Generator().manual_seed(i)
    return ({"name1": [f"seq{i % 3}/{i}.ppm"]},
            {"kpt": torch.rand(n, 2, generator=g).numpy().astype(np.float32),
             "desc": torch.randn(1, n, d, generator=g), "kp_score": torch.rand(1, n, 1, generator=g)})


def test_async_writer_matches_save_desc(tmp_path):
    from posfeat_b200.extractor import AsyncDescWriter, save_desc
    a_root, s_root = str(tmp_path / "async"), str(tmp_path / "sync")
    with AsyncDescWriter(a_root, "PoSFeat_t", workers=3, max_pending=4) as wr:
        for i in range(12):
            inputs, pr = _fake(i)
            msg = wr.save(inputs, pr)
            assert "kpts: 50" in msg
            save_desc(inputs, pr, s_root, "PoSFeat_t")
    for i in range(12):
        rel = os.path.join(f"seq{i % 3}", f"{i}.ppm.PoSFeat_t")
        za, zs = np.load(os.path.join(a_root, rel)), np.load(os.path.join(s_root, rel))
        assert set(za.files) == {"keypoints", "scores", "descriptors"}
        for k in za.files:
            np.testing.assert_array_equal(za[k], zs[k])
            assert za[k].dtype == np.float32


def test_async_writer_reports_errors(tmp_path):
    from posfeat_b200.extractor import AsyncDescWriter
    wr = AsyncDescWriter(str(tmp_path / "x"), "p")
    inputs, pr = _fake(0)
    wr.save(inputs, pr)
    pr_bad = dict(pr, desc=torch.zeros(1, 3, 3))
    os.makedirs(str(tmp_path / "x" / "seq0" / "9.ppm.p"))          # a directory where the file should go
    inputs_bad = {"name1": ["seq0/9.ppm"]}
    wr.save(inputs_bad, pr_bad)
    try:
        wr.close()
    except (IsADirectoryError, PermissionError, OSError):
        return
    raise AssertionError("writer error was swallowed")
