"""TEST / BASELINE INFRASTRUCTURE - makes the UNMODIFIED reference importable on the GPU box.

    python oracle/build_ref.py            (build container: needs /root/reference)

``/root/reference`` does not exist on the GPU box, and the reference is a script tree (no setup.py /
pyproject.toml), so ``pip install --target`` does not apply.  This recipe copies, file for file and without
touching a byte, the part of the reference tree its global aligner imports - ``dust3r/`` (minus datasets and caches)
and ``third_party/{__init__.py, raft.py, RAFT/core}`` (``dust3r/cloud_opt_flow/optimizer.py:13`` imports RAFT's
loader at module scope; the weights are absent and never loaded) - into the git-ignored ``oracle/_ref/``.  The directory
is listed in ``.gitignore`` (nothing of the reference enters the history) but not in ``.gpurunignore``: it travels with
the snapshot like the built ``.so`` files.  ``oracle/ref_harness.py`` falls back to it when ``/root/reference`` is absent;
``bench.py --impl reference`` and the ``cpu_baseline`` leg time the reference's own ``global_alignment_iter`` from it.
``__graft_entry__.build()`` runs this recipe whenever ``/root/reference`` is present.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("A3R_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")

TREES = [("dust3r", ("datasets", "__pycache__")),
         (os.path.join("third_party", "RAFT", "core"), ("__pycache__",))]
FILES = [os.path.join("third_party", "__init__.py"), os.path.join("third_party", "raft.py")]


def build(verbose: bool = False) -> str | None:
    """Copies the reference files (no-op without /root/reference).  Returns the destination, or None."""
    if not os.path.isdir(os.path.join(SRC, "dust3r", "cloud_opt_flow")):
        return DST if os.path.isdir(os.path.join(DST, "dust3r", "cloud_opt_flow")) else None
    n = 0
    for tree, skip in TREES:
        for root, dirs, files in os.walk(os.path.join(SRC, tree)):
            dirs[:] = [d for d in dirs if d not in skip]
            rel = os.path.relpath(root, SRC)
            os.makedirs(os.path.join(DST, rel), exist_ok=True)
            for f in files:
                if f.endswith((".pyc", ".pth", ".pt")):
                    continue
                shutil.copyfile(os.path.join(root, f), os.path.join(DST, rel, f))
                n += 1
    for f in FILES:
        os.makedirs(os.path.dirname(os.path.join(DST, f)), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        n += 1
    with open(os.path.join(DST, "ORIGIN.txt"), "w") as fh:
        fh.write(f"verbatim copy of {n} files of {SRC} made by oracle/build_ref.py; git-ignored; not product source\n")
    if verbose:
        print(f"oracle/_ref: {n} files copied from {SRC}")
    return DST


if __name__ == "__main__":
    out = build(verbose=True)
    if out is None:
        print("no reference tree here and no oracle/_ref", file=sys.stderr)
        sys.exit(1)
    print(out)
