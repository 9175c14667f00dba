#!/usr/bin/env python
"""ORACLE — test infrastructure, NOT product code.

Recipe for ``oracle/_ref``: a runnable copy of the UNMODIFIED reference files that make up the
per-ray path, taken from where they lie under ``/root/reference``:

    models.py  epipolar.py  geometry.py  resnet_block_fc.py  encoder.py
    utils/util.py  utils/pixel_util.py

``oracle/_ref/`` is a build output: git-ignored (no reference source enters the history) but not
gpurun-ignored, so it travels to the GPU box, where ``/root/reference`` does not exist.  There the
reference runs ON the B200 (its hard-coded ``.cuda()`` calls, geometry.py:320,398, then do what they
say) as the parity oracle of the ``-m gpu`` tests and as the "reference PyTorch on B200" row of
``bench.py``; on the host cores it is the ``--impl reference`` arm (``cpu_baseline.kind =
"reference"``).  The files are byte-identical copies (``MANIFEST.json`` holds their sha256); the
modules the reference imports at module scope but never touches on this path (timm, matplotlib, midas)
are stubbed at import time by ``oracle/ref_loader.py``, exactly like ``tests/golden/make_golden.py``.

    python oracle/build_ref.py            # no-op with a message when /root/reference is absent
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CAR_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ("models.py", "epipolar.py", "geometry.py", "resnet_block_fc.py", "encoder.py",
         "utils/util.py", "utils/pixel_util.py")


def build(verbose=True):
    """Returns True when oracle/_ref is (now) present."""
    if not os.path.isdir(REF):
        if verbose:
            print(f"oracle/build_ref: {REF} not present - keeping the prebuilt oracle/_ref "
                  f"({'found' if os.path.isdir(OUT) else 'MISSING'})")
        return os.path.isdir(OUT)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    commit = None
    sub = os.path.join(REF, ".SUBMODULES.json")
    if os.path.exists(sub):
        try:
            commit = json.load(open(sub)).get("commit")
        except Exception:
            commit = None
    json.dump({"source": REF, "commit": commit, "sha256": manifest}, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"oracle/build_ref: copied {len(FILES)} reference files to {OUT}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
