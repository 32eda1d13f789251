"""GPU parity at BASELINE.json's full sizes (the configurations round 1 left without parity evidence):
matcher at 16k and 64k x 64k against a blocked float64 oracle, the pair pipeline at the full C2 shape
(896x1200, k=8192) and at C3 (1056x1600, r=3, thr 0.5, k=20480 and 8192; configs/extract_aachen.yaml:27-40)
against the numpy oracle -- and, where it has been staged (oracle/_ref, oracle/build_ref.py), against the
reference's own functions run on the box's CPU.  Plus the host-side properties SURVEY 8(b) promises:
calls from two host threads on two streams do not share scratch memory."""
import threading

import numpy as np
import pytest
import torch

from _checks import assert_close_vec, check_argmax_exact
from oracle import posfeat_oracle as O

pytestmark = pytest.mark.gpu


def unit_desc(n, d, seed, base=None, noise=0.5):
    g = torch.Generator().manual_seed(seed)
    if base is None:
        x = torch.randn(n, d, generator=g)
    else:
        perm = torch.randperm(base.shape[0], generator=g)[:n]
        x = base[perm] + noise * torch.randn(n, d, generator=g)
    return torch.nn.functional.normalize(x, dim=1)


def assert_matches_exact(a, b, got, want, row_gap, col_gap, tol=1e-12):
    """Identical match lists; a differing pair is accepted only with a proven float64 tie in its row or column."""
    if got.shape == want.shape and np.array_equal(got, want):
        return 0
    gs, ws = {tuple(x) for x in got.tolist()}, {tuple(x) for x in want.tolist()}
    for (i, j) in gs ^ ws:
        assert row_gap[i] <= tol or col_gap[j] <= tol, \
            f"pair {(i, j)} differs without a float64 tie (row gap {row_gap[i]!r}, column gap {col_gap[j]!r})"
    return len(gs ^ ws)


@pytest.mark.parametrize("N,M,noise", [(16384, 16384, 0.5), (65536, 65536, 0.5), (20000, 65536, 1.0), (65536, 3000, 0.3)])
def test_mnn_large_vs_blocked_f64(N, M, noise):
    """BASELINE config 5 sizes.  nn12, nn21 and the match list of the tensor-core matcher (both entry forms:
    with nn21, and matches-only) against the exact float64 similarities evaluated block-wise on the host."""
    import posfeat_b200 as P
    a = unit_desc(N, 128, 101)
    b = unit_desc(M, 128, 102, base=a if M <= N else None, noise=noise)
    if M > N:                                   # the first N columns are noisy copies of the rows, the rest distractors
        b[:N] = unit_desc(N, 128, 103, base=a, noise=noise)
    want, nn12, nn21, row_gap, col_gap = O.mnn_blocked_f64(a.numpy(), b.numpy())
    ac, bc = a.cuda(), b.cuda()
    m2, k2, g12, g21 = P.mnn_match(ac, bc, algo=2, want_nn21=True)
    assert check_argmax_exact(a.numpy(), b.numpy(), g12.cpu().numpy(), nn12) == 0
    assert check_argmax_exact(b.numpy(), a.numpy(), g21.cpu().numpy(), nn21) == 0
    got2 = m2[:int(k2)].cpu().numpy()
    assert assert_matches_exact(a, b, got2, want, row_gap, col_gap) == 0
    m1, k1, h12, h21 = P.mnn_match(ac, bc, algo=2, want_nn21=False)          # matches-only: one direction + verification
    assert h21 is None
    np.testing.assert_array_equal(m1[:int(k1)].cpu().numpy(), got2)
    np.testing.assert_array_equal(h12.cpu().numpy(), g12.cpu().numpy())
    assert len(want) > min(N, M) // 4
    from posfeat_b200 import _runtime
    del ac, bc, m1, m2
    _runtime.release_workspaces()              # 64k x 64k scratch is gigabytes: hand it back
    torch.cuda.empty_cache()


def test_mnn_more_than_65536_columns(monkeypatch):
    """M > 65536 without an nn21 buffer: the list / group-entry form of the matches-only path has no column limit;
    the table form kept for A/B measurements (POSFEAT_MNN_TABLE=1) has one and must refuse BEFORE anything is
    queued (no sticky CUDA error, outputs untouched)."""
    import posfeat_b200 as P
    from posfeat_b200 import _lib
    from posfeat_b200._runtime import stream_ptr, workspace
    N, M = 3000, 65537
    a = unit_desc(N, 128, 7)
    b = unit_desc(M, 128, 8)
    b[:N] = unit_desc(N, 128, 9, base=a, noise=0.4)
    want, nn12, nn21, rg, cg = O.mnn_blocked_f64(a.numpy(), b.numpy())
    got = P.mnn_matcher(a.cuda(), b.cuda(), algo=2)
    assert assert_matches_exact(a, b, got, want, rg, cg) == 0
    m2, k2, _, g21 = P.mnn_match(a.cuda(), b.cuda(), algo=2, want_nn21=True)      # two directions, table form
    assert check_argmax_exact(b.numpy(), a.numpy(), g21.cpu().numpy(), nn21) == 0
    np.testing.assert_array_equal(m2[:int(k2)].cpu().numpy(), got)
    L = _lib.load()
    ac, bc = a.cuda(), b.cuda()
    o12 = torch.empty(N, dtype=torch.int32, device="cuda")
    mt = torch.empty((N, 2), dtype=torch.int64, device="cuda")
    nm = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws = workspace("mnn", L.posfeat_mnn_workspace_bytes(N, M, 128, 2), ac.device)
    torch.cuda.synchronize()
    monkeypatch.setenv("POSFEAT_MNN_TABLE", "1")
    st = L.posfeat_mnn_f32(ac.data_ptr(), N, 128, bc.data_ptr(), M, 128, 128, 2, o12.data_ptr(), None, mt.data_ptr(),
                           nm.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(ac.device))
    monkeypatch.delenv("POSFEAT_MNN_TABLE")
    assert st == 4 and "nn21" in _lib.last_error()
    torch.cuda.synchronize()                    # no sticky error: nothing was launched
    assert int(nm.item()) == 0


@pytest.mark.parametrize("N,M", [(1500, 1300), (4096, 4096), (777, 2049), (1, 5), (130, 1), (3000, 8192), (8192, 8192)])
def test_mnn_list_form_equals_table_form(monkeypatch, N, M):
    """The two formulations of the matches-only path -- candidate lists + group entries (default) and the
    chunk-maximum table + scan kernel (POSFEAT_MNN_TABLE=1) -- return the same nn12 and the same match list, on
    well-matched, noisy and unrelated descriptor sets (the latter stress the column verification: many
    competitors per chunk)."""
    import posfeat_b200 as P
    for noise, seed in ((0.3, 1), (1.0, 2), (None, 3)):
        a = unit_desc(N, 128, 40 + seed)
        b = unit_desc(M, 128, 50 + seed, base=a if (M <= N and noise is not None) else None, noise=noise or 0.0)
        ac, bc = a.cuda(), b.cuda()
        m1, k1, n1, _ = P.mnn_match(ac, bc, algo=2, want_nn21=False)
        monkeypatch.setenv("POSFEAT_MNN_TABLE", "1")
        m0, k0, n0, _ = P.mnn_match(ac, bc, algo=2, want_nn21=False)
        monkeypatch.delenv("POSFEAT_MNN_TABLE")
        np.testing.assert_array_equal(n1.cpu().numpy(), n0.cpu().numpy())
        np.testing.assert_array_equal(m1[:int(k1)].cpu().numpy(), m0[:int(k0)].cpu().numpy())
        want, nn12, nn21, rg, cg = O.mnn_blocked_f64(a.numpy(), b.numpy())
        assert check_argmax_exact(a.numpy(), b.numpy(), n1.cpu().numpy(), nn12) == 0
        assert assert_matches_exact(a, b, m1[:int(k1)].cpu().numpy(), want, rg, cg) == 0


def test_mnn_list_overflow_and_duplicates():
    """Adversarial inputs for the list form: (a) columns ordered so that every row's chunk maxima keep rising
    tile after tile (the 15-entry lists overflow -> exhaustive rescoring of those rows), (b) blocks of identical
    descriptors (every group entry of a chunk ties -> the competitor list overflows -> exhaustive verification),
    (c) all-negative similarities (thresholds <= 0)."""
    import posfeat_b200 as P
    g = torch.Generator().manual_seed(5)
    # (a) b_j = normalize(u + eps_j * noise) with eps decreasing in j: similarity to a = u rises with j
    u = torch.nn.functional.normalize(torch.randn(1, 128, generator=g), dim=1)
    M = 8192
    eps = torch.linspace(2.0, 0.0, M).unsqueeze(1)
    b = torch.nn.functional.normalize(u + eps * torch.randn(M, 128, generator=g) * 0.2, dim=1)
    a = torch.nn.functional.normalize(u + 0.05 * torch.randn(600, 128, generator=g), dim=1)
    for x, y in ((a, b), (b, a)):
        want, nn12, nn21, rg, cg = O.mnn_blocked_f64(x.numpy(), y.numpy())
        m, k, n12, _ = P.mnn_match(x.cuda(), y.cuda(), algo=2, want_nn21=False)
        assert check_argmax_exact(x.numpy(), y.numpy(), n12.cpu().numpy(), nn12) == 0
        assert assert_matches_exact(x, y, m[:int(k)].cpu().numpy(), want, rg, cg) == 0
    # (b) many identical rows and columns
    z = torch.nn.functional.normalize(torch.randn(6, 128, generator=g), dim=1).repeat(400, 1)
    zp = z[torch.randperm(z.shape[0], generator=g)]
    want = O.mnn_blocked_f64(z.numpy(), zp.numpy())[0]
    m, k, _, _ = P.mnn_match(z.cuda(), zp.cuda(), algo=2, want_nn21=False)
    np.testing.assert_array_equal(m[:int(k)].cpu().numpy(), want)      # exact ties: first index, no allowance needed
    # (c) every similarity negative
    v = torch.nn.functional.normalize(torch.rand(1500, 128, generator=g) + 0.1, dim=1)
    want, nn12, _, rg, cg = O.mnn_blocked_f64(v.numpy(), (-v).numpy())
    m, k, n12, _ = P.mnn_match(v.cuda(), (-v).cuda(), algo=2, want_nn21=False)
    assert check_argmax_exact(v.numpy(), (-v).numpy(), n12.cpu().numpy(), nn12) == 0
    assert assert_matches_exact(v, -v, m[:int(k)].cpu().numpy(), want, rg, cg) == 0


def test_mnn_competitor_list_over_capacity(monkeypatch, capfd):
    """Column chunks whose competitor list exceeds the verification kernel's capacity (256 rows): the list is
    narrowed -- members first, then only the non-members that reach a surviving member's threshold -- and only a
    list that is still too long, or more than 32 members, takes the exhaustive path.  Four constructed columns:
      c0    a weak (0.8) and a strong (0.99) member + 400 rows at 0.85 that prefer their own columns: strong member wins
      c2000 as c0 plus one non-member at 0.995 (above the strong member) that prefers its own column:
            the column's nearest row is not a member -> no match for the column
      c4000 one member at 0.8 + 400 non-members at 0.85 (narrowed list still over capacity -> exhaustive): no match
      c6000 40 members at 0.90..0.939 + 300 non-members at 0.92 (over 32 members): the best member wins."""
    import posfeat_b200 as P
    g = torch.Generator().manual_seed(77)
    D, M = 128, 8192
    cols = torch.nn.functional.normalize(torch.randn(M, D, generator=g), dim=1)

    def at(y, p, n):                     # n unit rows with similarity p to the unit vector y
        u = torch.randn(n, D, generator=g)
        u = torch.nn.functional.normalize(u - (u @ y)[:, None] * y[None], dim=1)
        return p * y[None] + (1 - p * p) ** 0.5 * u

    special = {0, 2000 // 8, 4000 // 8, 6000 // 8}
    rows, free = [], iter(j for j in range(100, M, 5) if j // 8 not in special)      # columns for rows that prefer "their own"

    def with_own_columns(r):
        for x in r:
            cols[next(free)] = torch.nn.functional.normalize(x + 0.003 * torch.randn(D, generator=g), dim=0)
        rows.append(r)

    for c in (0, 2000):
        rows.append(at(cols[c], 0.8, 1)); rows.append(at(cols[c], 0.99, 1))
        with_own_columns(at(cols[c], 0.85, 400))
    with_own_columns(at(cols[2000], 0.995, 1))
    rows.append(at(cols[4000], 0.8, 1)); with_own_columns(at(cols[4000], 0.85, 400))
    rows.append(torch.cat([at(cols[6000], 0.90 + 0.001 * k, 1) for k in range(40)]))
    with_own_columns(at(cols[6000], 0.92, 300))
    rows.append(torch.nn.functional.normalize(torch.randn(1000, D, generator=g), dim=1))      # background
    a = torch.cat(rows)
    a = a[torch.randperm(a.shape[0], generator=g)].contiguous()
    want, nn12, nn21, rg, cg = O.mnn_blocked_f64(a.numpy(), cols.numpy())
    monkeypatch.setenv("POSFEAT_MNN_DEBUG", "1")
    capfd.readouterr()
    m, k, n12, _ = P.mnn_match(a.cuda(), cols.cuda(), algo=2, want_nn21=False)
    torch.cuda.synchronize()
    err = capfd.readouterr().err
    monkeypatch.delenv("POSFEAT_MNN_DEBUG")
    assert "overflow-chunks" in err and int(err.split("overflow-chunks")[1].split()[0]) >= 4, err
    assert check_argmax_exact(a.numpy(), cols.numpy(), n12.cpu().numpy(), nn12) == 0
    got = m[:int(k)].cpu().numpy()
    assert assert_matches_exact(a, cols, got, want, rg, cg) == 0
    matched_cols = set(got[:, 1].tolist())
    assert 0 in matched_cols and 6000 in matched_cols and 2000 not in matched_cols and 4000 not in matched_cols
    assert [(nn12 == c).sum() for c in (0, 2000, 4000, 6000)] == [2, 2, 1, 40]          # the construction holds


def test_mnn_batched_chunk_ordered_rescoring(monkeypatch):
    """A batched matches-only call of >= 64k rows takes the chunk-ordered rescoring (enumerate -> chunk buckets ->
    merge); smaller calls and POSFEAT_MNN_RESCORE_WARP=1 take the warp-per-row kernel.  Both against the float64
    oracle on well-matched, noisy and unrelated pairs of one batch, ragged sizes included (N, M not multiples of 8 or
    of the tile), and against each other bit for bit."""
    from posfeat_b200 import _lib
    from posfeat_b200._runtime import check, lib, stream_ptr, workspace
    L = lib()
    for N, M, P in ((8192, 8192, 9), (7001, 9003, 10)):
        A = torch.stack([unit_desc(N, 128, 300 + i) for i in range(P)])
        B = torch.stack([unit_desc(M, 128, 400 + i, base=A[i] if (i % 3 != 2 and M <= N) else None, noise=(0.1, 0.6, 0.0)[i % 3])
                         for i in range(P)])
        if M > N:                                 # noisy copies in the first N columns of two pairs out of three
            for i in range(P):
                if i % 3 != 2:
                    B[i, :N] = unit_desc(N, 128, 500 + i, base=A[i], noise=(0.1, 0.6)[i % 3])
        Ac, Bc = A.cuda(), B.cuda()
        outs = []
        for env in (None, "1"):
            if env:
                monkeypatch.setenv("POSFEAT_MNN_RESCORE_WARP", env)
            nn12 = torch.empty((P, N), dtype=torch.int32, device="cuda")
            matches = torch.empty((P, N, 2), dtype=torch.int64, device="cuda")
            nm = torch.empty(P, dtype=torch.int32, device="cuda")
            ws = workspace("mnn", L.posfeat_mnn_batched_workspace_bytes(P, N, M, 128, 2), Ac.device)
            check(L.posfeat_mnn_batched_f32(Ac.data_ptr(), Ac.stride(0), N, Ac.stride(1), Bc.data_ptr(), Bc.stride(0), M,
                                            Bc.stride(1), 128, P, 2, nn12.data_ptr(), None, matches.data_ptr(), nm.data_ptr(),
                                            ws.data_ptr(), ws.numel(), stream_ptr(Ac.device)))
            torch.cuda.synchronize()
            if env:
                monkeypatch.delenv("POSFEAT_MNN_RESCORE_WARP")
            outs.append((nn12.cpu().numpy(), matches.cpu().numpy(), nm.cpu().numpy()))
        np.testing.assert_array_equal(outs[0][0], outs[1][0])
        np.testing.assert_array_equal(outs[0][2], outs[1][2])
        for i in range(P):
            k = int(outs[0][2][i])
            np.testing.assert_array_equal(outs[0][1][i, :k], outs[1][1][i, :k])
            want, o12, _, rg, cg = O.mnn_blocked_f64(A[i].numpy(), B[i].numpy())
            assert check_argmax_exact(A[i].numpy(), B[i].numpy(), outs[0][0][i], o12) == 0
            assert assert_matches_exact(A[i], B[i], outs[0][1][i, :k], want, rg, cg) == 0


def _pipeline_case(H, W, cfg, seed, P=1):
    g = torch.Generator().manual_seed(seed)
    score = torch.nn.functional.softplus(torch.randn(2 * P, 1, H, W, generator=g))
    score[1::2] = torch.nn.functional.softplus(torch.log(torch.expm1(score[0::2])) + 0.05 * torch.randn(P, 1, H, W, generator=g))
    fmap = torch.randn(2 * P, 128, H // 4, W // 4, generator=g)
    fmap[1::2] = fmap[0::2] + 0.3 * torch.randn(P, 128, H // 4, W // 4, generator=g)
    return score, fmap


@pytest.mark.parametrize("name,H,W,cfg", [
    ("C2", 896, 1200, dict(nms_radius=1, num_pts=8192, thr=0.9, thr_mod="abs", use_nms=True, stable=True)),
    ("C3_k20480", 1056, 1600, dict(nms_radius=3, num_pts=20480, thr=0.5, thr_mod="abs", use_nms=True, stable=True)),
    ("C3_k8192", 1056, 1600, dict(nms_radius=3, num_pts=8192, thr=0.5, thr_mod="abs", use_nms=True, stable=True)),
])
def test_pair_pipeline_full_size_vs_oracle(name, H, W, cfg):
    """PairPipeline.run at the full BASELINE shapes: keypoint indices bit-exact, descriptors 1e-5, match
    lists identical (float64 ties excepted and proven) -- against the numpy oracle, and against the
    reference's own torch functions on the CPU when they are staged on this box."""
    from posfeat_b200.pairs import PairPipeline
    score, fmap = _pipeline_case(H, W, cfg, seed=len(name))
    pipe = PairPipeline(cfg)
    f_cl = fmap.cuda().contiguous(memory_format=torch.channels_last)
    feats, matches, nm = pipe.run(score.cuda(), f_cl)
    n = feats["n"]
    ocfg = {k: v for k, v in cfg.items() if k != "stable"}
    okps, osc, oidx, ocnt = O.generate_kpts_single(score.numpy(), return_idx=True, **ocfg)
    assert n == oidx.shape[1]
    idx = feats["idx"].cpu().numpy()
    sc = score.numpy()[:, 0, 1:-1, 1:-1].reshape(2, -1)
    for b in range(2):
        # equal scores may be ordered either way by an implementation (torch.topk leaves it open); ours is pinned
        # to index-ascending, which is what the oracle does too -> exact equality
        np.testing.assert_array_equal(idx[b], oidx[b])
        assert np.all(np.diff(sc[b][idx[b]]) <= 0)
    np.testing.assert_allclose(feats["kps_n"].cpu().numpy(), okps, rtol=1e-5, atol=2e-6)
    odesc = O.sample_feat_by_coord(fmap.numpy(), okps, True)
    assert_close_vec(feats["desc"].cpu().numpy(), odesc, 1e-5)
    want, nn12, nn21, rg, cg = O.mnn_blocked_f64(odesc[0], odesc[1])
    got = matches[0, :int(nm[0])].cpu().numpy()
    # the GPU matcher works on ITS descriptors (equal to the oracle's to 1e-5): compare on those
    gd = feats["desc"].cpu().numpy()
    want_g, _, _, rg_g, cg_g = O.mnn_blocked_f64(gd[0], gd[1])
    assert assert_matches_exact(gd[0], gd[1], got, want_g, rg_g, cg_g) == 0
    # and the lists from the two descriptor sets differ only where a 1e-5 descriptor difference can flip a near tie
    gs, ws = {tuple(x) for x in got.tolist()}, {tuple(x) for x in want.tolist()}
    for (i, j) in gs ^ ws:
        assert rg[i] < 1e-4 or cg[j] < 1e-4, (i, j, rg[i], cg[j])
    assert len(got) > n // 2
    # ---- the reference's own functions (torch CPU), when staged ----
    from oracle import ref_runner
    if ref_runner.available():
        kps_r, desc_r, m_r, idx_r = ref_runner.run_pair(score, fmap, dict(cfg), want_idx=True)
        assert kps_r.shape[1] == n
        idx_r = idx_r.numpy()
        for b in range(2):
            assert set(idx_r[b].tolist()) == set(idx[b].tolist())
            # same order wherever the selected score is unique
            v = sc[b][idx[b]]
            vals, cnt = np.unique(v, return_counts=True)
            uniq = np.isin(v, vals[cnt == 1])
            np.testing.assert_array_equal(idx_r[b][uniq], idx[b][uniq])
        pix = lambda ix, m: {(int(ix[0][i]), int(ix[1][j])) for i, j in m}
        diff = pix(idx, got) ^ pix(idx_r, m_r)
        # torch's float32 BLAS similarity vs the exact one: only near ties (|gap| < 1e-6) may flip
        pos = {int(p): k for k, p in enumerate(idx[0])}
        posb = {int(p): k for k, p in enumerate(idx[1])}
        for (pa, pb) in diff:
            assert rg_g[pos[pa]] < 1e-6 or cg_g[posb[pb]] < 1e-6, (pa, pb)


def test_workspace_reentrancy_two_threads_two_streams():
    """the entry points are safe from several host threads on different streams.  Two threads
    run different matcher problems concurrently, each on its own stream; results equal the serial ones."""
    import posfeat_b200 as P
    from posfeat_b200 import _runtime
    probs = []
    for t in range(2):
        a = unit_desc(3000 + 500 * t, 128, 50 + t)
        b = unit_desc(2500 + 700 * t, 128, 60 + t, base=a)
        probs.append((a.cuda(), b.cuda()))
    serial = [P.mnn_matcher(a, b, algo=2) for a, b in probs]
    torch.cuda.synchronize()
    out, errs = [None, None], []
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    barrier = threading.Barrier(2)

    def work(t):
        try:
            with torch.cuda.stream(streams[t]):
                barrier.wait()
                for _ in range(20):
                    m, k, _, _ = P.mnn_match(probs[t][0], probs[t][1], algo=2, want_nn21=False)
                streams[t].synchronize()
                out[t] = m[:int(k)].cpu().numpy()
        except Exception as e:      # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs, errs
    for t in range(2):
        np.testing.assert_array_equal(out[t], serial[t])
    keys = [k for k in _runtime._workspaces if k[0] == "mnn"]
    assert len({k[2] for k in keys}) >= 2          # one scratch buffer per stream
