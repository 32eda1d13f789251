"""Stage the UNMODIFIED reference modules of the hot path under oracle/_ref/ (TEST / BASELINE INFRASTRUCTURE).

    python oracle/build_ref.py            # needs the reference checkout (default /root/reference)

The reference is Python, so "building" it means placing the few files of the path where the GPU box can
import them: oracle/_ref/ is git-ignored (nothing of the reference enters the history) but NOT
gpurun-ignored, so it travels with the snapshot exactly like the built .so does.  Staged verbatim:

    losses/{__init__,preprocess_utils,preprocess,epipolarloss,kploss}.py   
    evaluations/aachen/matchers.py                                          (a-13, f-1)
    evaluations/ETH_local_feature/custom_matcher.py                         (f-1)

Users: bench.py --impl reference (the reference's own torch-CPU code on the host cores,
cpu_baseline.kind = "reference"), the `reference_on_b200` comparison leg (same functions on cuda tensors)
and tests that want the real functions on the GPU box.  The product never imports anything from here.
A MANIFEST.json with the sha256 of every staged file is written beside them.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = [
    "losses/__init__.py", "losses/preprocess_utils.py", "losses/preprocess.py", "losses/epipolarloss.py",
    "losses/kploss.py", "evaluations/aachen/matchers.py", "evaluations/ETH_local_feature/custom_matcher.py",
]


def build(ref_root=None, quiet=False):
    ref_root = ref_root or os.environ.get("POSFEAT_REFERENCE", "/root/reference")
    if not os.path.isdir(ref_root):
        raise FileNotFoundError(f"reference checkout not found at {ref_root}")
    manifest = {}
    for rel in FILES:
        src = os.path.join(ref_root, rel)
        # evaluation helpers are staged flat under ref_eval/ (their directories are not packages)
        out_rel = rel if rel.startswith("losses/") else os.path.join(
            "ref_eval", {"evaluations/aachen/matchers.py": "aachen_matchers.py",
                         "evaluations/ETH_local_feature/custom_matcher.py": "eth_custom_matcher.py"}[rel])
        dst = os.path.join(DST, out_rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[out_rel] = {"from": rel, "sha256": hashlib.sha256(f.read()).hexdigest()}
    open(os.path.join(DST, "ref_eval", "__init__.py"), "w").close()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"reference_root": ref_root, "files": manifest}, f, indent=1)
    if not quiet:
        print(f"staged {len(manifest)} reference files under {DST}")
    return DST


if __name__ == "__main__":
    build(sys.argv[1] if len(sys.argv) > 1 else None)
