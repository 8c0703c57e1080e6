#!/usr/bin/env python
"""Recipe for ``oracle/_ref``: the UNMODIFIED reference implementation of the hot path, for the CPU arm of bench.py.

The reference (PyXRD v0.8.4) is pure Python; its hot path is the package ``pyxrd.calculations`` plus the settings module
it reads.  This script copies exactly those files, byte for byte, from the reference tree
(``/root/reference``, present in the build container only) into ``oracle/_ref/`` -- a directory that is git-ignored
(no reference source enters the history) but travels to the GPU box with the snapshot, like a built ``.so``.
``__graft_entry__.build()`` runs it when the reference tree is present; ``oracle/reference_arm.py`` loads the result
behind the same four-alias numpy / scipy compatibility shim the golden generator uses (tests/golden/_live_reference.py),
without editing any file.  Nothing under ``pyxrd_b200/`` imports it: it is test / baseline infrastructure only.
"""
import glob
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"
TARGET = os.path.join(HERE, "_ref")
FILES = ["pyxrd/__init__.py", "pyxrd/__version.py", "pyxrd/data/__init__.py", "pyxrd/data/settings.py",
         "pyxrd/data/appdirs.py", "pyxrd/calculations/*.py"]


def make(reference_root=REFERENCE_ROOT, target=TARGET):
    """-> number of files copied (0 when the reference tree is absent: the existing oracle/_ref, if any, is kept)"""
    if not os.path.isdir(os.path.join(reference_root, "pyxrd", "calculations")):
        return 0
    manifest = {}
    for pattern in FILES:
        for src in sorted(glob.glob(os.path.join(reference_root, pattern))):
            rel = os.path.relpath(src, reference_root)
            dst = os.path.join(target, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(target, "MANIFEST.json"), "w") as fh:
        json.dump({"source": reference_root, "sha256": manifest}, fh, indent=1, sort_keys=True)
    return len(manifest)


if __name__ == "__main__":
    n = make()
    print("oracle/_ref: %d files" % n if n else "reference tree not found; nothing copied")
    sys.exit(0)
