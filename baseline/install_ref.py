"""Install the UNMODIFIED reference hot-path packages next to the bench so they travel to the GPU box.

    python baseline/install_ref.py            (also run by __graft_entry__.build() when /root/reference exists)

The reference (DIALLab-SKKU/MergeRec) is pure Python with no setup.py / pyproject.toml, so "installing" it means
copying the two packages the hot paths live in -- `rec_retrieval/merger` and `rec_retrieval/evaluator`, which import
nothing but torch -- byte for byte into `baseline/_ref/rec_retrieval/`.  That directory is git-ignored (reference
sources never enter the history) but not gpurun-ignored, so `bench.py --impl reference` and the `cpu_baseline` leg
can time the reference's own code on the GPU box's host cores (`kind: "reference"`).  Nothing under
`mergerec_b200/` ever imports it.  A SHA-256 manifest of the copied files is written beside them.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MR_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
PACKAGES = ("merger", "evaluator")


def install(src_root: str = SRC, dst_root: str = DST) -> bool:
    pkg = os.path.join(src_root, "rec_retrieval")
    if not os.path.isdir(pkg):
        return False
    out = os.path.join(dst_root, "rec_retrieval")
    if os.path.isdir(out):
        shutil.rmtree(out)
    os.makedirs(out)
    shutil.copy2(os.path.join(pkg, "__init__.py"), os.path.join(out, "__init__.py"))
    for name in PACKAGES:
        shutil.copytree(os.path.join(pkg, name), os.path.join(out, name),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {}
    for root, _, files in os.walk(out):
        for f in sorted(files):
            path = os.path.join(root, f)
            with open(path, "rb") as fh:
                manifest[os.path.relpath(path, dst_root)] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(dst_root, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src_root, "files": manifest}, fh, indent=1, sort_keys=True)
    return True


def load():
    """Import the installed reference packages; returns the `rec_retrieval` module or None when absent."""
    if not os.path.isdir(os.path.join(DST, "rec_retrieval", "merger")):
        return None
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import rec_retrieval  # noqa: F401
    import rec_retrieval.evaluator  # noqa: F401
    import rec_retrieval.merger  # noqa: F401
    return rec_retrieval


if __name__ == "__main__":
    ok = install()
    print("installed the reference hot-path packages into", DST if ok else "(nothing: reference not found at %s)" % SRC)
