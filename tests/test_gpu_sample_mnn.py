"""GPU parity: descriptor sampling + L2 norm, and mutual-NN matching."""
import numpy as np
import pytest
import torch

from _checks import assert_close_vec, check_argmax_exact, check_mnn_near_tie, check_ratio_near_tie
from oracle import posfeat_oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-5   # north_star: descriptors within 1e-5 relative at fp32 (vector-wise, see _checks)


def test_sample_golden(golden):
    import posfeat_b200 as P
    g = golden("sample")
    x = torch.from_numpy(g["x"]).cuda()
    c = torch.from_numpy(g["coord"]).cuda()
    assert_close_vec(P.sample_feat_by_coord(x, c, False).cpu().numpy(), g["raw"], REL)
    assert_close_vec(P.sample_feat_by_coord(x, c, True).cpu().numpy(), g["normed"], REL)
    x128 = torch.from_numpy(g["x128"]).cuda()
    c128 = torch.from_numpy(g["coord128"]).cuda()
    assert_close_vec(P.sample_feat_by_coord(x128, c128, True).cpu().numpy(), g["normed128"], REL)
    # channels_last (NHWC) input takes the vectorised kernel and must agree
    xcl = x128.contiguous(memory_format=torch.channels_last)
    assert_close_vec(P.sample_feat_by_coord(xcl, c128, True).cpu().numpy(), g["normed128"], REL)


@pytest.mark.parametrize("b,c,h,w,n", [(1, 128, 224, 300, 8192), (2, 128, 120, 160, 4096), (1, 64, 33, 47, 1000),
                                       (1, 20, 9, 11, 77), (1, 256, 16, 16, 300)])
def test_sample_vs_oracle(b, c, h, w, n):
    import posfeat_b200 as P
    g = torch.Generator().manual_seed(c * 1000 + n)
    x = torch.randn(b, c, h, w, generator=g)
    coord = torch.rand(b, n, 2, generator=g) * 2.04 - 1.02
    want = O.sample_feat_by_coord(x.numpy(), coord.numpy(), True)
    got = P.sample_feat_by_coord(x.cuda(), coord.cuda(), True)
    assert_close_vec(got.cpu().numpy(), want, REL)
    got_cl = P.sample_feat_by_coord(x.cuda().contiguous(memory_format=torch.channels_last), coord.cuda(), True)
    assert_close_vec(got_cl.cpu().numpy(), want, REL)
    raw = P.sample_feat_by_coord(x.cuda(), coord.cuda(), False).cpu().numpy()
    assert_close_vec(raw, O.sample_feat_by_coord(x.numpy(), coord.numpy(), False), REL)
    nrm = np.linalg.norm(got.cpu().numpy(), axis=-1)
    assert np.all((np.abs(nrm - 1) < 1e-5) | (nrm == 0))


def test_sample_empty_and_bf16():
    import posfeat_b200 as P
    x = torch.randn(1, 128, 8, 8, device="cuda")
    assert P.sample_feat_by_coord(x, torch.zeros(1, 0, 2, device="cuda"), True).shape == (1, 0, 128)
    c = torch.rand(1, 50, 2, device="cuda") * 2 - 1
    f, fb = P.sample_l2norm(x, c, True, want_bf16=True)
    assert torch.equal(fb, f.to(torch.bfloat16))


ALGOS = [1, 2]


@pytest.mark.parametrize("algo", ALGOS)
def test_mnn_golden(golden, algo):
    import posfeat_b200 as P
    g = golden("mnn")
    a, b = torch.from_numpy(g["a"]).cuda(), torch.from_numpy(g["b"]).cuda()
    m = P.mnn_matcher(a, b, algo=algo)
    assert m.dtype == np.int64 and m.ndim == 2 and m.shape[1] == 2
    check_mnn_near_tie(g["a"], g["b"], m, g["mnn"])
    ad, bd = torch.from_numpy(g["ad"]).cuda(), torch.from_numpy(g["bd"]).cuda()
    np.testing.assert_array_equal(P.mnn_matcher(ad, bd, algo=algo), g["mnn_dup"])   # first-index tie rule
    np.testing.assert_array_equal(P.mnn_matcher(bd, ad, algo=algo), g["mnn_dup_t"])
    from posfeat_b200.matchers import mutual_nn_matcher
    check_mnn_near_tie(g["a"], g["b"], mutual_nn_matcher(a, b), g["mutual_nn"])


def unit_desc(n, d, seed, base=None, noise=0.5):
    g = torch.Generator().manual_seed(seed)
    if base is None:
        x = torch.randn(n, d, generator=g)
    else:
        perm = torch.randperm(base.shape[0], generator=g)[:n]
        x = base[perm] + noise * torch.randn(n, d, generator=g)
    return torch.nn.functional.normalize(x, dim=1)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("N,M,D", [(1500, 1300, 128), (1024, 1024, 128), (4096, 4096, 128), (777, 2049, 128),
                                   (1, 5, 128), (130, 1, 128), (300, 200, 64), (257, 129, 100)])
def test_mnn_vs_oracle(algo, N, M, D):
    import posfeat_b200 as P
    if algo == 2 and D != 128:
        with pytest.raises(Exception):
            P.mnn_matcher(torch.zeros(N, D, device="cuda"), torch.zeros(M, D, device="cuda"), algo=2)
        return
    a = unit_desc(N, D, 1)
    b = unit_desc(M, D, 2, base=a if M <= N else None)
    want, nn12, nn21 = O.mnn_matcher(a.numpy(), b.numpy(), exact=True, return_nn=True)
    matches, nm, g12, g21 = P.mnn_match(a.cuda(), b.cuda(), algo=algo)
    got = matches[:int(nm.item())].cpu().numpy()
    check_mnn_near_tie(a.numpy(), b.numpy(), got, want)
    # the exact (float64) argmax is reproduced index for index: zero mismatches outside proven float64 ties
    check_argmax_exact(a.numpy(), b.numpy(), g12.cpu().numpy(), nn12)
    check_argmax_exact(b.numpy(), a.numpy(), g21.cpu().numpy(), nn21)
    assert np.all(np.diff(got[:, 0]) > 0)


@pytest.mark.parametrize("N,M", [(1500, 1300), (4096, 4096), (777, 2049), (1, 5), (130, 1), (3000, 8192)])
def test_mnn_matches_only_path(N, M):
    """Tensor-core matcher without nn21 (one direction + column verification) must
    return exactly the match list of the two-direction algorithm and of the oracle."""
    import posfeat_b200 as P
    a = unit_desc(N, 128, 31)
    b = unit_desc(M, 128, 32, base=a if M <= N else None)
    want = O.mnn_matcher(a.numpy(), b.numpy(), exact=True)
    m_full, k_full, _, _ = P.mnn_match(a.cuda(), b.cuda(), algo=2, want_nn21=True)
    m_fast, k_fast, nn12, nn21 = P.mnn_match(a.cuda(), b.cuda(), algo=2, want_nn21=False)
    assert nn21 is None
    got = m_fast[:int(k_fast)].cpu().numpy()
    np.testing.assert_array_equal(got, m_full[:int(k_full)].cpu().numpy())
    check_mnn_near_tie(a.numpy(), b.numpy(), got, want)


def test_mnn_matches_only_duplicates(golden):
    import posfeat_b200 as P
    g = golden("mnn")
    ad, bd = torch.from_numpy(g["ad"]).cuda(), torch.from_numpy(g["bd"]).cuda()
    for x, y, key in ((ad, bd, "mnn_dup"), (bd, ad, "mnn_dup_t")):
        m, k, _, _ = P.mnn_match(x, y, algo=2, want_nn21=False)
        np.testing.assert_array_equal(m[:int(k)].cpu().numpy(), g[key])       # first-index tie rule in both directions
    # many identical rows and columns
    z = torch.nn.functional.normalize(torch.randn(4, 128), dim=1).repeat(50, 1).cuda()
    m, k, _, _ = P.mnn_match(z, z, algo=2, want_nn21=False)
    np.testing.assert_array_equal(m[:int(k)].cpu().numpy(), O.mnn_matcher(z.cpu().numpy(), z.cpu().numpy(), exact=True))


def test_mnn_8k_properties():
    """BASELINE size (8192 x 8192 x 128): result is a partial permutation,
    consistent with nn12/nn21, and symmetric under swapping the operands."""
    import posfeat_b200 as P
    a = unit_desc(8192, 128, 11).cuda()
    b = unit_desc(8192, 128, 12, base=a.cpu()).cuda()
    m, nm, nn12, nn21 = P.mnn_match(a, b)
    k = int(nm.item())
    m = m[:k].cpu().numpy()
    assert k > 2000
    np.testing.assert_array_equal(P.mnn_matcher(a, b), m)          # matches-only path (no nn21)
    assert len(np.unique(m[:, 0])) == k and len(np.unique(m[:, 1])) == k
    n12, n21 = nn12.cpu().numpy(), nn21.cpu().numpy()
    assert np.array_equal(n12[m[:, 0]], m[:, 1]) and np.array_equal(n21[m[:, 1]], m[:, 0])
    mt = P.mnn_matcher(b, a)
    assert {tuple(x) for x in m.tolist()} == {(j, i) for i, j in mt.tolist()}
    # both algorithms agree
    m1 = P.mnn_matcher(a, b, algo=1)
    np.testing.assert_array_equal(m1, m)


def test_mnn_host_entry_and_errors():
    import posfeat_b200 as P
    a, b = unit_desc(600, 128, 3), unit_desc(500, 128, 4)
    want = O.mnn_matcher(a.numpy(), b.numpy(), exact=True)
    check_mnn_near_tie(a.numpy(), b.numpy(), P.mnn_matcher(a, b), want)      # CPU tensors -> host entry point
    with pytest.raises(IndexError):
        P.mnn_matcher(torch.zeros(0, 128, device="cuda"), torch.zeros(4, 128, device="cuda"))
    with pytest.raises(ValueError):
        P.mnn_matcher(torch.zeros(3, 128, device="cuda"), torch.zeros(4, 64, device="cuda"))
    # strided rows (a view into a wider matrix)
    v = torch.nn.functional.normalize(torch.randn(300, 128, device="cuda"), dim=1)
    wide = torch.zeros(300, 256, device="cuda")
    wide[:, :128] = v
    assert wide[:, :128].stride(0) == 256
    np.testing.assert_array_equal(P.mnn_matcher(wide[:, :128], v), P.mnn_matcher(v, v))


def test_ratio_matchers_golden(golden):
    """ratio_matcher / mutual_nn_ratio_matcher against outputs of the reference functions
    (evaluations/aachen/matchers.py:17-75) stored by oracle/make_golden.py."""
    from posfeat_b200.matchers import mutual_nn_ratio_matcher, ratio_matcher
    g = golden("mnn")
    a, b = torch.from_numpy(g["a"]).cuda(), torch.from_numpy(g["b"]).cuda()
    got = ratio_matcher(a, b, ratio=0.95)
    assert got.dtype == np.int64 and got.shape[1] == 2
    check_ratio_near_tie(g["a"], g["b"], got, g["ratio"], 0.95, False)
    got = mutual_nn_ratio_matcher(a, b, ratio=0.9)
    check_ratio_near_tie(g["a"], g["b"], got, g["mutual_ratio"], 0.9, True)


def ratio_pair(N, M, D):
    """b = copies of rows of a under per-row noise levels (Lowe ratios spread over
    roughly 0.3..1), scattered among distractors when M > N."""
    g = torch.Generator().manual_seed(42)
    a = unit_desc(N, D, 41)
    k = min(N, M)
    lvl = (0.05 + 0.2 * torch.rand(k, 1, generator=g)) * (128.0 / D) ** 0.5
    b = a[torch.randperm(N, generator=g)[:k]] + lvl * torch.randn(k, D, generator=g)
    if M > k:
        b = torch.cat([b, torch.randn(M - k, D, generator=g)])[torch.randperm(M, generator=g)]
    return a, torch.nn.functional.normalize(b, dim=1)


@pytest.mark.parametrize("mutual", [False, True])
@pytest.mark.parametrize("N,M,D,ratio", [(1500, 1300, 128, 0.9), (777, 2049, 128, 0.75), (2, 2, 16, 0.95),
                                         (300, 5000, 64, 0.8), (4096, 4096, 128, 0.95)])
def test_ratio_matchers_vs_oracle(mutual, N, M, D, ratio):
    from posfeat_b200.matchers import mutual_nn_ratio_matcher, ratio_matcher
    a, b = ratio_pair(N, M, D)
    want = O.ratio_matchers(a.numpy(), b.numpy(), ratio=ratio, mutual=mutual, exact=True)
    fn = mutual_nn_ratio_matcher if mutual else ratio_matcher
    got = fn(a.cuda(), b.cuda(), ratio=ratio)
    amb = check_ratio_near_tie(a.numpy(), b.numpy(), got, want, ratio, mutual)
    assert amb < max(4, N // 50)                      # the check is not vacuous
    if N > 100:
        assert len(want) > 0


def test_ratio_matchers_errors():
    from posfeat_b200.matchers import ratio_matcher
    with pytest.raises(RuntimeError):                  # torch.topk(sim, 2) raises on a single column
        ratio_matcher(torch.zeros(5, 128, device="cuda"), torch.zeros(1, 128, device="cuda"))
    with pytest.raises(ValueError):
        ratio_matcher(torch.zeros(5, 128, device="cuda"), torch.zeros(5, 64, device="cuda"))


def test_ratio_tc_equals_simt(monkeypatch):
    """Large D=128 problems take the tensor-core route (top-2 rescoring); it must return what
    the exact SIMT kernel returns, up to rows sitting on the ratio threshold."""
    from posfeat_b200.matchers import mutual_nn_ratio_matcher, ratio_matcher
    a, b = ratio_pair(3000, 2500, 128)
    ac, bc = a.cuda(), b.cuda()
    for fn, mutual in ((ratio_matcher, False), (mutual_nn_ratio_matcher, True)):
        tc = fn(ac, bc, ratio=0.8)
        monkeypatch.setenv("POSFEAT_RATIO_SIMT", "1")
        simt = fn(ac, bc, ratio=0.8)
        monkeypatch.delenv("POSFEAT_RATIO_SIMT")
        assert len(tc) > 100
        check_ratio_near_tie(a.numpy(), b.numpy(), tc, simt, 0.8, mutual)
