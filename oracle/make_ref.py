"""Recipe for ``oracle/_ref/`` -- TEST / BENCH INFRASTRUCTURE ONLY, never imported by the product.

``/root/reference`` does not exist on the GPU box, and the reference is pure Python with nothing to compile, so
"building" the real reference for the checker and for ``bench.py --impl reference`` means laying its files out where
they can travel: this script copies the reference's own ``contrastyou/``, ``semi_seg/`` and ``config/`` trees and
extracts the bundled ``deepclustering2-2.0.0`` wheel into ``oracle/_ref/reference/`` (git-ignored, NOT gpurun-ignored:
it ships with the snapshot like the built ``.so``).  Nothing is modified; the import shims the 2020 code needs on
torch 2.11 / Python 3.12 live in ``oracle/ref_loader.py`` and ``oracle/ref_epocher.py`` and are applied at import time.

    python oracle/make_ref.py            # run by __graft_entry__.build() when /root/reference is present

The copy is refreshed when a source file is newer; a MANIFEST.json records the sha256 of every file so that the GPU-box
run can state exactly which reference bytes it timed.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref", "reference")
SRC = os.environ.get("IIC_REFERENCE_ROOT", "/root/reference")
WHEEL = "deepclustering2-2.0.0-py3-none-any.whl"
TREES = ("contrastyou", "semi_seg", "config")


def make(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref/reference is in place (fresh or already there), False when there is no source."""
    if not os.path.isfile(os.path.join(SRC, "contrastyou", "losses", "iic_loss.py")):
        ok = os.path.isfile(os.path.join(DEST, "contrastyou", "losses", "iic_loss.py"))
        if verbose:
            print(f"[make_ref] {SRC} absent; " + ("using the existing oracle/_ref" if ok else "oracle/_ref not built"))
        return ok
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    for t in TREES:
        shutil.copytree(os.path.join(SRC, t), os.path.join(DEST, t),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".data"))
    with zipfile.ZipFile(os.path.join(SRC, WHEEL)) as z:
        for m in z.namelist():
            if m.startswith("deepclustering2/"):
                z.extract(m, DEST)
    manifest = {}
    for root, _dirs, files in os.walk(DEST):
        for f in sorted(files):
            p = os.path.join(root, f)
            with open(p, "rb") as fh:
                manifest[os.path.relpath(p, DEST)] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=0, sort_keys=True)
    if verbose:
        print(f"[make_ref] {len(manifest)} reference files -> {DEST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
