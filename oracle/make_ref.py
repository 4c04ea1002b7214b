"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference files of the hot path, staged for the GPU box -- TEST /
BASELINE INFRASTRUCTURE, never imported by the product (``drin_b200/``).

    python oracle/make_ref.py          # run in the build container (needs /root/reference)

The reference is pure Python: "building" it means staging the files of the DRIN path
(``drin/model.py`` -> ``common/args.py``, ``baselines/ghmfc.py``; ``common/utils.py`` for TripletLoss / TopkAccuracy) and
its two callers (``train.py``, ``drin/data.py``: the entry point and the loader, executed unmodified by
``tests/test_reference_train_py.py``) byte for byte under ``oracle/_ref/``.  That directory is git-ignored (reference sources never enter this repository's
history) but not gpurun-ignored, so it travels to the GPU box, where ``/root/reference`` does not exist:
``bench.py --impl reference`` and ``cpu_baseline`` then time the reference's own code (``kind: "reference"``) instead of
the oracle port, and ``tests/test_reference_glue.py`` drives ``drin_b200.Model()`` under the reference's real
``common.args``.  ``oracle/ref_import.py`` loads it (same stubs as for /root/reference).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DRIN_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("drin/model.py", "baselines/ghmfc.py", "common/args.py", "common/utils.py", "drin/data.py", "train.py", "LICENSE")


def stage(verbose: bool = True) -> bool:
    """Copy the files; returns False (and leaves an existing staged copy alone) when the reference tree is absent."""
    if not os.path.isfile(os.path.join(SRC, "drin", "model.py")):
        if verbose:
            print(f"reference tree not found at {SRC}: oracle/_ref left as it is "
                  f"({'present' if staged() else 'absent'})")
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "sha256": manifest}, fh, indent=1)
    if verbose:
        print(f"staged {len(manifest)} reference files under {DST}")
    return True


def staged() -> bool:
    return os.path.isfile(os.path.join(DST, "drin", "model.py"))


if __name__ == "__main__":
    sys.exit(0 if stage() or staged() else 1)
