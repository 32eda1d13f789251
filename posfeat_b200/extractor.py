"""Drop-in for the boundary functions of the reference's managers/extractor.py:
``process`` (:318-355) and ``save_desc`` (:254-316), backed by the B200 kernels.

The reference's Extractor shell (config/yaml/logging/dataloader) stays host
Python and is out of scope; ``FeatureExtractor`` below only carries the state
``process`` and ``save_desc`` read (config keys are the reference's).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import preprocess_utils as pu


def process(inputs, outputs, config, detector=None, sift_kp=False, remove_pad=False):
    """managers/extractor.py:318-355.  inputs: {'im1', 'name1', ['im1_ori','pad1','coord1','scale']},
    outputs: {'local_map' [1,D,H/4,W/4], 'local_point' [1,1,H,W]}.
    Returns {'kpt': ndarray (n,2) float32 pixel xy, 'desc': Tensor [1,n,D], 'kp_score': Tensor [1,n,1]}."""
    detector = detector or pu.generate_kpts_single
    desc_f = outputs["local_map"]
    name = inputs["name1"][0]
    if remove_pad:
        b, c, h, w = inputs["im1_ori"].shape
        pad = inputs["pad1"]
        desc_f = desc_f[:, :, :-(pad[3] // 4), :-(pad[0] // 4)]
        outputs["local_point"] = outputs["local_point"][:, :, :-(pad[3] // 4), :-(pad[0] // 4)]
    else:
        b, c, h, w = inputs["im1"].shape
    if sift_kp:
        coords = inputs["coord1"]
        coord_n = pu.normalize_coords(coords, h, w)
        kp_score = torch.ones_like(coord_n)[:, :, :1]
    else:
        cfg = config["detector_config"]
        if config.get("data") == "Aachen_Day_Night" and name.split("/")[0] == "query":
            cfg = config["detector_config_query"]
        coord_n, kp_score = detector(outputs["local_point"], **cfg)
        coords = pu.denormalize_coords(coord_n, h, w)
    feat_f = pu.sample_feat_by_coord(desc_f, coord_n, config["loss_distance"] == "cos")
    kpt = coords.cpu().numpy().squeeze(0)
    if "scale" in inputs:
        kpt = kpt * inputs["scale"].cpu().numpy()
    return {"kpt": kpt, "desc": feat_f, "kp_score": kp_score}


def save_desc(inputs, processed, desc_root, postfix, save_npz=True, save_h5=False, image_size=None):
    """managers/extractor.py:254-316.  Writes ``<desc_root>/<name1>.<postfix>`` with
    np.savez(keypoints (n,2) f32, scores (n,1) f32, descriptors (n,D) f32) -- a
    file object is passed, so no '.npz' suffix is appended, exactly like the
    reference.  The h5 layout follows SURVEY.md section 3.4 (the reference's own
    h5 branch references undefined names); it needs h5py."""
    kpt = processed["kpt"]
    name = inputs["name1"][0]
    save_path = os.path.join(desc_root, name)
    os.makedirs(os.path.dirname(save_path), exist_ok=True)
    desc = processed["desc"].squeeze(0).detach().cpu().numpy()
    scores = processed["kp_score"].squeeze(0).detach().cpu().numpy()
    message = "\nkpts: {}".format(kpt.shape[0])
    if save_npz:
        _write_npz(save_path + ".{}".format(postfix), kpt, scores, desc)
    if save_h5:
        _write_h5(desc_root, name, kpt, scores, desc, image_size)
    return message


def _write_npz(path, kpt, scores, desc):
    with open(path, "wb") as f:
        np.savez(f, keypoints=kpt, scores=scores, descriptors=desc)


_h5_lock = None


def _write_h5(desc_root, name, kpt, scores, desc, image_size=None):
    """The h5 outputs of managers/extractor.py:273-314 (layout per SURVEY.md section 3.4; the reference's own
    branch references undefined names w, h, grp, fh5):
      <desc_root>h5/<seq>/{keypoints,descriptors,scores,scales}.h5 : dataset <image stem> = array
          (image-matching-benchmark format; scales = ones_like(scores))
      <desc_root>h5/feat.h5 : group <name1> with datasets keypoints, scores, descriptors, image_size=[w, h]
          (hloc format)
    Files are opened in append mode like the reference; writers are serialised (HDF5 files are not safe to
    append to from several threads)."""
    global _h5_lock
    try:
        import h5py
    except ImportError as e:   # the authoring image has no h5py
        raise RuntimeError("save_h5=True needs h5py, which is not installed") from e
    if _h5_lock is None:
        import threading
        _h5_lock = threading.Lock()
    h5_root = str(desc_root).rstrip("/") + "h5"
    stem = name.split(".")[0]
    seq, key = "/".join(stem.split("/")[:-1]), stem.split("/")[-1]
    with _h5_lock:
        os.makedirs(os.path.join(h5_root, seq), exist_ok=True)
        for fname, data in (("keypoints.h5", kpt), ("descriptors.h5", desc), ("scores.h5", scores),
                            ("scales.h5", np.ones_like(scores))):
            with h5py.File(os.path.join(h5_root, seq, fname), "a") as fh:
                fh[key] = data
        with h5py.File(os.path.join(h5_root, "feat.h5"), "a") as fh:
            grp = fh.create_group(name)
            grp.create_dataset("keypoints", data=kpt)
            grp.create_dataset("scores", data=scores)
            grp.create_dataset("descriptors", data=desc)
            if image_size is not None:
                grp.create_dataset("image_size", data=np.asarray(image_size))


class AsyncDescWriter:
    """Non-blocking ``save_desc`` for the extraction loop.

    The reference writes every image's ``.npz`` synchronously between two backbone
    calls (managers/extractor.py:254-316: D2H copy, ``np.savez``, disk).  Here the
    D2H copies go to pinned staging buffers on a side stream and a small thread pool
    does the ``np.savez``; the GPU keeps working on the next image.  File contents and
    names are identical to ``save_desc`` (``.npz`` and, with ``save_h5=True``, the h5 files: those are
    appended to under a lock, one image at a time).  ``close()`` (or leaving the
    ``with`` block) waits for all pending files and re-raises the first writer error.
    """

    def __init__(self, desc_root, postfix, workers=4, max_pending=64, save_npz=True, save_h5=False):
        from concurrent.futures import ThreadPoolExecutor
        if save_h5:
            import h5py  # noqa: F401  (fail at construction, not in a worker thread)
        self.desc_root, self.postfix = desc_root, postfix
        self.save_npz, self.save_h5 = bool(save_npz), bool(save_h5)
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.pending = []
        self.max_pending = max_pending
        self.copy_stream = torch.cuda.Stream() if torch.cuda.is_available() else None

    def _stage(self, t):
        """Device tensor -> (pinned host tensor, event); host tensors pass through."""
        t = t.detach()
        if not t.is_cuda:
            return t.contiguous(), None
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        self.copy_stream.wait_stream(torch.cuda.current_stream(t.device))
        with torch.cuda.stream(self.copy_stream):
            host.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        t.record_stream(self.copy_stream)
        return host, ev

    def save(self, inputs, processed, image_size=None):
        name = inputs["name1"][0]
        if image_size is None and self.save_h5 and "im1" in inputs:
            h, w = inputs["im1"].shape[2:]
            image_size = (w, h)
        path = os.path.join(self.desc_root, name) + ".{}".format(self.postfix)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        kpt = np.array(processed["kpt"], copy=True)
        desc, e1 = self._stage(processed["desc"].squeeze(0))
        score, e2 = self._stage(processed["kp_score"].squeeze(0))

        def job():
            for e in (e1, e2):
                if e is not None:
                    e.synchronize()
            if self.save_npz:
                _write_npz(path, kpt, score.numpy(), desc.numpy())
            if self.save_h5:
                _write_h5(self.desc_root, name, kpt, score.numpy(), desc.numpy(), image_size)
            return path

        self.pending.append(self.pool.submit(job))
        if len(self.pending) >= self.max_pending:      # bound the pinned memory held by queued files
            self.pending.pop(0).result()
        return "\nkpts: {}".format(kpt.shape[0])

    def flush(self):
        pending, self.pending = self.pending, []
        return [f.result() for f in pending]

    def close(self):
        try:
            return self.flush()
        finally:
            self.pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class FeatureExtractor:
    """Minimal carrier of the state Extractor.process/save_desc use (config keys
    as in configs/extract_hpatches.yaml)."""

    def __init__(self, config: dict, desc_root: str = None):
        self.config = config
        self.sift_kp = bool(config.get("use_sift", False))
        self.detector = getattr(pu, config.get("detector", "generate_kpts_single"))
        self.desc_root = desc_root
        self.save_npz = config.get("save_npz", True)
        self.save_h5 = config.get("save_h5", False)

    def process(self, inputs, outputs, remove_pad=False):
        return process(inputs, outputs, self.config, self.detector, self.sift_kp, remove_pad)

    def save_desc(self, inputs, outputs, processed):
        h, w = inputs["im1"].shape[2:]
        return save_desc(inputs, processed, self.desc_root, self.config["postfix"], self.save_npz,
                         self.save_h5, image_size=(w, h))
