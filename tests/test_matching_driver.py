"""Multi-GPU matching driver over feature files (posfeat_b200/matching_driver.py): host logic on the CPU with the
oracle as the matcher -- sharding by sequence, the descriptor cache, the HPatches scoring loop
(evaluations/hpatches/evaluation.py:40-96) and the gather to rank 0 (gloo, world_size 2)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import posfeat_oracle as O

METHOD = "PoSFeat_t"
SEQS = ["i_ajuntament", "i_bologna", "v_abstract", "v_adam", "v_apprentices", "i_castle", "v_bark"]


def _homography(seq, k):
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"{seq}/{k}".encode()))      # deterministic across processes
    Hm = np.eye(3) + 0.05 * rng.standard_normal((3, 3))
    Hm[2, :2] *= 1e-3
    Hm[2, 2] = 1.0
    return Hm


def _write_features(root):
    """Per sequence: image 1 has n keypoints; image k holds noisy copies of a subset (true matches, projected with
    H_1_k plus a little jitter) and distractors.  Written with np.savez exactly like save_desc."""
    rng = np.random.default_rng(0)
    for s_i, seq in enumerate(SEQS):
        n = 150 + 10 * s_i
        kp1 = rng.uniform(0, 600, size=(n, 2)).astype(np.float32)
        d1 = rng.standard_normal((n, 32)).astype(np.float32)
        d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
        os.makedirs(os.path.join(root, seq), exist_ok=True)
        with open(os.path.join(root, seq, f"1.ppm.{METHOD}"), "wb") as f:
            np.savez(f, keypoints=kp1, scores=rng.random((n, 1)).astype(np.float32), descriptors=d1)
        for k in range(2, 7):
            keep = rng.permutation(n)[: n - 20 * k]
            ph = np.concatenate([kp1[keep], np.ones((len(keep), 1))], 1) @ _homography(seq, k).T
            kpk = (ph[:, :2] / ph[:, 2:] + rng.normal(0, 0.4 * k, size=(len(keep), 2))).astype(np.float32)
            dk = d1[keep] + 0.25 * k * rng.standard_normal((len(keep), 32)).astype(np.float32)
            extra = 30
            kpk = np.concatenate([kpk, rng.uniform(0, 600, size=(extra, 2)).astype(np.float32)])
            dk = np.concatenate([dk, rng.standard_normal((extra, 32)).astype(np.float32)])
            dk /= np.linalg.norm(dk, axis=1, keepdims=True)
            order = rng.permutation(len(kpk))
            with open(os.path.join(root, seq, f"{k}.ppm.{METHOD}"), "wb") as f:
                np.savez(f, keypoints=kpk[order], scores=rng.random((len(kpk), 1)).astype(np.float32),
                         descriptors=dk[order].astype(np.float32))


def _oracle_matcher(a, b):
    return O.mnn_matcher(a.numpy(), b.numpy(), exact=True)


def _reference_loop(root):
    """The reference's benchmark_features loop, sequence by sequence on one process (semantics of
    evaluation.py:40-96 with the oracle as matcher)."""
    rng_thr = range(1, 16)
    i_err, v_err = {t: 0.0 for t in rng_thr}, {t: 0.0 for t in rng_thr}
    n_feats, n_matches, seq_type = [], [], []
    for seq in sorted(SEQS):
        za = np.load(os.path.join(root, seq, f"1.ppm.{METHOD}"))
        ka, da = za["keypoints"], za["descriptors"]
        n_feats.append(ka.shape[0])
        for k in range(2, 7):
            zb = np.load(os.path.join(root, seq, f"{k}.ppm.{METHOD}"))
            kb, db = zb["keypoints"], zb["descriptors"]
            n_feats.append(kb.shape[0])
            m = O.mnn_matcher(da, db, exact=True)
            pa = np.concatenate([ka[m[:, 0], :2], np.ones([m.shape[0], 1])], axis=1)
            pb = np.transpose(np.dot(_homography(seq, k), np.transpose(pa)))
            pb = pb[:, :2] / pb[:, 2:]
            d = np.sqrt(np.sum((kb[m[:, 1], :2] - pb) ** 2, axis=1))
            n_matches.append(m.shape[0])
            seq_type.append(seq[0])
            for t in rng_thr:
                (i_err if seq[0] == "i" else v_err)[t] += np.mean(d <= t)
    return i_err, v_err, np.array(seq_type), np.array(n_feats), np.array(n_matches)


def test_hpatches_benchmark_single_rank_equals_reference_loop(tmp_path):
    from posfeat_b200 import matching_driver as MD
    root = str(tmp_path)
    _write_features(root)
    i_err, v_err, (seq_type, n_feats, n_matches) = MD.hpatches_benchmark(SEQS, root, METHOD, _homography, matcher=_oracle_matcher,
                                                                        device="cpu")
    ri, rv, rtype, rfeats, rmatches = _reference_loop(root)
    for t in ri:
        assert abs(i_err[t] - ri[t]) < 1e-12 and abs(v_err[t] - rv[t]) < 1e-12
    np.testing.assert_array_equal(seq_type, rtype)
    np.testing.assert_array_equal(n_feats, rfeats)
    np.testing.assert_array_equal(n_matches, rmatches)
    assert i_err[15] > i_err[1] and n_matches.min() > 20
    assert MD.match_pairs.last_cache_loads == 6 * len(SEQS)            # every file read once: image 1 is cached


def test_match_pairs_pair_list_and_cache(tmp_path):
    """Aachen-shaped use: a 'name1 name2' list, groups = queries; every rank's share is whole groups."""
    from posfeat_b200 import matching_driver as MD
    root = str(tmp_path)
    _write_features(root)
    lst = os.path.join(root, "pairs.txt")
    pairs = [(f"{s}/1.ppm", f"{t}/{k}.ppm") for s in SEQS[:4] for t in SEQS[4:] for k in (2, 3)]
    with open(lst, "w") as f:
        f.write("\n".join(f"{a} {b}" for a, b in pairs) + "\n")
    assert MD.read_pair_list(lst) == pairs
    got = {}
    for r in range(3):
        part = MD.match_pairs(pairs, root, METHOD, rank=r, world=3, device="cpu", matcher=_oracle_matcher)
        assert len({p[0] for p in part}) * 6 == len(part)               # whole query groups
        assert not (set(part) & set(got))
        got.update(part)
    assert set(got) == set(pairs)
    a, b = pairs[5]
    za, zb = np.load(os.path.join(root, f"{a}.{METHOD}")), np.load(os.path.join(root, f"{b}.{METHOD}"))
    np.testing.assert_array_equal(got[pairs[5]], O.mnn_matcher(za["descriptors"], zb["descriptors"], exact=True))
    kp, d = MD.read_features(root, a, METHOD, top_k=40)
    assert kp.shape == (40, 2) and d.shape == (40, 32)
    # a tiny cache still works (evicts)
    part = MD.match_pairs(pairs[:6], root, METHOD, device="cpu", matcher=_oracle_matcher, cache_bytes=1)
    assert len(part) == 6


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, root, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from posfeat_b200 import matching_driver as MD
    res = MD.hpatches_benchmark(SEQS, root, METHOD, _homography, rank=rank, world=world, matcher=_oracle_matcher, device="cpu")
    if rank == 0:
        out.put(res)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_hpatches_benchmark_two_ranks_equals_one(tmp_path):
    from posfeat_b200 import matching_driver as MD
    root = str(tmp_path)
    _write_features(root)
    one = MD.hpatches_benchmark(SEQS, root, METHOD, _homography, matcher=_oracle_matcher, device="cpu")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, root, q)) for r in range(2)]
    for p in procs:
        p.start()
    two = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for t in one[0]:
        assert abs(one[0][t] - two[0][t]) < 1e-9 and abs(one[1][t] - two[1][t]) < 1e-9
    for a, b in zip(one[2], two[2]):
        np.testing.assert_array_equal(a, b)


def test_homography_hits_equals_host_scoring():
    """homography_hits (tensor form of evaluation.py:73-90, float64) gives the hit fractions of the numpy scoring,
    threshold by threshold, incl. the empty match list."""
    from posfeat_b200 import matching_driver as MD
    rng = np.random.default_rng(4)
    kp1 = rng.uniform(0, 600, size=(300, 2)).astype(np.float32)
    Hm = _homography("v_test", 3)
    ph = np.concatenate([kp1, np.ones((300, 1))], 1) @ Hm.T
    kp2 = (ph[:, :2] / ph[:, 2:] + rng.normal(0, 3.0, size=(300, 2))).astype(np.float32)
    m = np.stack([rng.permutation(300)[:200], rng.permutation(300)[:200]], 1).astype(np.int64)
    m[:120, 1] = m[:120, 0]                                   # 120 true correspondences, 80 random ones
    thr = tuple(range(1, 16))
    dist = MD.homography_errors(kp1, kp2, m, Hm)
    want = np.array([np.mean(dist <= t) for t in thr])
    got = MD.homography_hits(kp1, kp2, torch.from_numpy(m), 200, Hm, thr).numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-15)
    assert got[-1] > 0.5 and got[0] < got[-1]
    pad = torch.cat([torch.from_numpy(m), torch.full((50, 2), -1, dtype=torch.int64)])        # rows past n_matches are ignored
    np.testing.assert_allclose(MD.homography_hits(kp1, kp2, pad, 200, Hm, thr).numpy(), want, rtol=0, atol=1e-15)
    assert float(MD.homography_hits(kp1, kp2, torch.zeros((0, 2), dtype=torch.int64), 0, Hm, thr).sum()) == 0.0


def test_hpatches_benchmark_device_scoring_path_equals_host_path(tmp_path):
    """score_on_device=True (matches kept as tensors, hit fractions computed by homography_hits) accumulates the same
    benchmark as the host path -- here with the oracle as the tensor-returning matcher, on the CPU."""
    from posfeat_b200 import matching_driver as MD
    root = str(tmp_path)
    _write_features(root)

    def tensor_matcher(a, b):
        m = torch.from_numpy(_oracle_matcher(a, b))
        return m, m.shape[0]

    host = MD.hpatches_benchmark(SEQS, root, METHOD, _homography, matcher=_oracle_matcher, device="cpu")
    dev = MD.hpatches_benchmark(SEQS, root, METHOD, _homography, matcher=tensor_matcher, device="cpu", score_on_device=True)
    for t in host[0]:
        assert abs(host[0][t] - dev[0][t]) < 1e-12 and abs(host[1][t] - dev[1][t]) < 1e-12
    np.testing.assert_array_equal(host[2][2], dev[2][2])
    np.testing.assert_array_equal(host[2][1], dev[2][1])
