"""Recipe for ``oracle/_ref``: an UNMODIFIED, importable copy of the reference's Python packages, so that the GPU box
(which has no /root/reference) can time and check against the reference's OWN ``PatchRefiner.forward`` -- on its host cores
(``bench.py --impl reference``, ``cpu_baseline.kind = "reference"``) and as eager PyTorch on the B200.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is 100 % Python without a setup.py / pyproject.toml, so the base
contract's ``pip install --target baseline/_ref /root/reference`` has nothing to install; this script does what that
install would do -- it copies the two importable trees (``estimator/``, ``external/``) verbatim.  ``oracle/_ref/`` is a build
output: git-ignored (never enters history), not gpurun-ignored (travels to the GPU box like the built .so).  It is imported
only through ``oracle/ref_shim.py`` (which stubs the non-arithmetic packages the reference imports), never by the product.

    python -m oracle.build_ref          # no-op (keeps an existing copy) when /root/reference is absent
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PRV2_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
TREES = ("estimator", "external")


def _tree_digest(root: str) -> str:
    h = hashlib.sha256()
    for t in TREES:
        for d, dirs, files in os.walk(os.path.join(root, t)):
            dirs.sort()
            dirs[:] = [x for x in dirs if x != "__pycache__"]
            for f in sorted(files):
                if f.endswith(".pyc"):
                    continue
                p = os.path.join(d, f)
                h.update(os.path.relpath(p, root).encode())
                with open(p, "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()


def build(verbose: bool = False) -> str | None:
    """Copy the reference packages to oracle/_ref when the source tree is present; return the path (None if neither
    the source nor an earlier copy exists)."""
    if not os.path.isdir(os.path.join(SRC, "estimator")):
        return DST if os.path.isdir(os.path.join(DST, "estimator")) else None
    stamp = os.path.join(DST, "SOURCE_SHA256")
    dig = _tree_digest(SRC)
    if os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for t in TREES:
        shutil.copytree(os.path.join(SRC, t), os.path.join(DST, t), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(stamp, "w") as fh:
        fh.write(dig)
    if verbose:
        print(f"copied {', '.join(TREES)} from {SRC} to {DST} ({dig[:12]})")
    return DST


if __name__ == "__main__":
    print(build(verbose=True))
    sys.exit(0)
