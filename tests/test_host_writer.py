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


# ---- h5 outputs (managers/extractor.py:273-314, layout per SURVEY.md section 3.4) ------------------
class _RecordingH5:
    """Minimal stand-in for the h5py calls save_desc makes (File(path, 'a') as a context manager,
    item assignment, create_group / create_dataset): records {file: {dataset path: array}} so that the
    layout the writer produces can be checked on a box without h5py."""
    files = {}

    class _Group:
        def __init__(self, store, prefix):
            self.store, self.prefix = store, prefix

        def __setitem__(self, key, value):
            path = self.prefix + key
            if path in self.store:
                raise ValueError(f"unable to create link (name already exists): {path}")   # h5py behaviour
            self.store[path] = np.asarray(value)

        def create_group(self, name):
            path = self.prefix + name + "/"
            if any(k.startswith(path) for k in self.store):
                raise ValueError("unable to create group (name already exists)")
            return _RecordingH5._Group(self.store, path)

        def create_dataset(self, name, data=None):
            self[name] = data

    class File(_Group):
        def __init__(self, path, mode):
            assert mode == "a", "the reference appends (h5py.File(..., 'a'))"
            super().__init__(_RecordingH5.files.setdefault(os.path.normpath(str(path)), {}), "")

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False


def _check_h5_layout(files, root, n_img=6):
    h5_root = os.path.normpath(root.rstrip("/") + "h5")
    feat = files[os.path.join(h5_root, "feat.h5")]
    for i in range(n_img):
        inputs, pr = _fake(i)
        name = inputs["name1"][0]
        seq, stem = f"seq{i % 3}", str(i)
        want = {"keypoints": pr["kpt"], "descriptors": pr["desc"][0].numpy(), "scores": pr["kp_score"][0].numpy(),
                "scales": np.ones_like(pr["kp_score"][0].numpy())}
        for k, v in want.items():                       # image-matching-benchmark files: one dataset per image stem
            got = np.asarray(files[os.path.join(h5_root, seq, k + ".h5")][stem])
            np.testing.assert_array_equal(got, v)
            assert got.dtype == np.float32
        for k in ("keypoints", "scores", "descriptors"):   # hloc file: group <name1>
            np.testing.assert_array_equal(np.asarray(feat[f"{name}/{k}"]), want[k])
        np.testing.assert_array_equal(np.asarray(feat[f"{name}/image_size"]), np.array([640, 480]))


def test_h5_layout_sync_and_async_with_recording_backend(tmp_path, monkeypatch):
    """Executes the h5 branch of save_desc and of AsyncDescWriter against a recording stand-in for h5py and
    checks file names, group / dataset paths, shapes, dtypes and values."""
    import sys
    import types
    from posfeat_b200.extractor import AsyncDescWriter, save_desc
    fake = types.ModuleType("h5py")
    fake.File = _RecordingH5.File
    monkeypatch.setitem(sys.modules, "h5py", fake)
    _RecordingH5.files = {}
    s_root = str(tmp_path / "sync")
    for i in range(6):
        inputs, pr = _fake(i)
        save_desc(inputs, pr, s_root, "PoSFeat_t", save_npz=False, save_h5=True, image_size=(640, 480))
    assert not os.path.exists(os.path.join(s_root, "seq0", "0.ppm.PoSFeat_t"))      # save_npz=False
    _check_h5_layout(_RecordingH5.files, s_root)
    a_root = str(tmp_path / "async")
    with AsyncDescWriter(a_root, "PoSFeat_t", workers=3, save_npz=True, save_h5=True) as wr:
        for i in range(6):
            inputs, pr = _fake(i)
            inputs["im1"] = torch.zeros(1, 3, 480, 640)
            wr.save(inputs, pr)
    _check_h5_layout(_RecordingH5.files, a_root)
    assert os.path.exists(os.path.join(a_root, "seq0", "0.ppm.PoSFeat_t"))
    # writing the same image twice fails like h5py does (the reference opens with 'a' and assigns)
    inputs, pr = _fake(0)
    try:
        save_desc(inputs, pr, s_root, "PoSFeat_t", save_npz=False, save_h5=True)
    except ValueError:
        pass
    else:
        raise AssertionError("duplicate h5 dataset was silently overwritten")


def test_h5_real_files(tmp_path):
    """With h5py installed: the files on disk read back with the documented layout."""
    import pytest
    h5py = pytest.importorskip("h5py")
    from posfeat_b200.extractor import AsyncDescWriter
    root = str(tmp_path / "real")
    with AsyncDescWriter(root, "PoSFeat_t", save_npz=False, save_h5=True) as wr:
        for i in range(6):
            inputs, pr = _fake(i)
            wr.save(inputs, pr, image_size=(640, 480))
    files = {}
    h5_root = root + "h5"
    for dirpath, _, fnames in os.walk(h5_root):
        for fn in fnames:
            store = {}
            with h5py.File(os.path.join(dirpath, fn), "r") as fh:
                fh.visititems(lambda k, o: store.__setitem__(k, o[()]) if hasattr(o, "shape") else None)
            files[os.path.normpath(os.path.join(dirpath, fn))] = store
    _check_h5_layout(files, root)


def test_async_h5_without_h5py_fails_at_construction(tmp_path, monkeypatch):
    import sys
    from posfeat_b200.extractor import AsyncDescWriter
    monkeypatch.setitem(sys.modules, "h5py", None)       # import h5py -> ImportError
    try:
        AsyncDescWriter(str(tmp_path / "x"), "p", save_h5=True)
    except ImportError:
        return
    raise AssertionError("missing h5py must fail loudly")
