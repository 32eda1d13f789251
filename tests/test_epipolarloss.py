"""EpipolarLoss_full drop-in (losses/epipolarloss.py:8-101) against the reference's own run: the fixture
tests/golden/preprocess.npz holds the processed dict, the loss and its components (oracle/make_golden.py).
Plain tensor arithmetic -- runs on the CPU."""
import numpy as np
import pytest
import torch


def test_epipolar_loss_full_golden(golden):
    from posfeat_b200.epipolarloss import EpipolarLoss_full
    g = golden("preprocess")
    H, W = int(g["H"]), int(g["W"])
    B = g["F1"].shape[0]
    inputs = dict(im1=torch.zeros(B, 3, H, W), im2=torch.zeros(B, 3, H, W), F1=torch.from_numpy(g["F1"]),
                  F2=torch.from_numpy(g["F2"]))
    processed = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("p_")}
    processed["temperature"] = 60
    for k in ("feat1w_corloc", "feat2w_corloc"):
        processed[k].requires_grad_()
    cfg = dict(grid_cost_thr=0.5, win_cost_thr=0.1, use_std_as_weight=True, weight_grid=0.3, weight_window=1)
    L = EpipolarLoss_full(cfg)
    assert L.__lossname__ == "EpipolarLoss_fullinfo"
    loss, comp = L(inputs, {}, processed)
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-6)
    assert set(comp) == {"loss_g1", "loss_w1", "loss_g2", "loss_w2", "percent_g", "percent_w"}
    for k, v in comp.items():
        np.testing.assert_allclose(float(v), float(g["c_" + k]), rtol=1e-6, err_msg=k)
    loss.backward()                                                    # weights are detached, costs are not
    assert processed["feat1w_corloc"].grad.abs().sum() > 0
    with pytest.raises(NameError):
        EpipolarLoss_full(dict(cfg, use_std_as_weight=False))(inputs, {}, processed)
