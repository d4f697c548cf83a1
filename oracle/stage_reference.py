"""Stage the UNMODIFIED reference tree as ``baseline/_ref`` (TEST / BASELINE INFRASTRUCTURE ONLY).

zyang1580/SML is 14 plain-Python files with no build system (nothing to ``pip install``), so the
"install" of the reference arm is a file copy: ``/root/reference`` -> ``baseline/_ref``.  The target is
git-ignored (reference sources never enter the history) but NOT gpurun-ignored, so it travels to the
GPU box, where ``bench.py --impl reference`` and the ``cpu_baseline`` / ``torch_cuda_baseline`` legs
import it through ``oracle/ref_harness.py``.  Nothing under ``sml_b200/`` reads it.

    python oracle/stage_reference.py          # done by __graft_entry__.build() when /root/reference exists
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SML_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def stage(src=SRC, dst=DST) -> bool:
    """Copies the tree (python sources + READMEs only).  Returns False when the source is absent
    (GPU box: the staged copy that came with the snapshot is used as it is)."""
    if not os.path.isdir(os.path.join(src, "model")):
        return os.path.isdir(os.path.join(dst, "model"))
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc"))
    for d, _, files in os.walk(dst):                 # the mount is read-only; the copy need not be
        os.chmod(d, 0o755)
        for f in files:
            os.chmod(os.path.join(d, f), 0o644)
    return True


if __name__ == "__main__":
    ok = stage()
    print("baseline/_ref:", "staged" if ok else "reference tree not available")
    sys.exit(0 if ok else 1)
