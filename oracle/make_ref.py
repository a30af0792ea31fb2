"""Recipe for oracle/_ref: a verbatim, git-ignored snapshot of the reference's Python sources (TEST INFRASTRUCTURE).

    python oracle/make_ref.py            # only where /root/reference exists (the build container)

The reference is pure Python, so "building" it is a copy: `/root/reference/core` -> `oracle/_ref/core`, byte for
byte (plus LICENSE and a MANIFEST with SHA-256 per file).  `oracle/_ref/` is listed in .gitignore -- reference
sources never enter this repository's history -- but not in .gpurunignore, so the snapshot travels to the GPU box
next to libstreamcorr.so.  There it lets the tests and bench.py run the UNMODIFIED reference:

  * core/corr.py, core/gma.py           the hot-path oracle itself (CPU baseline `kind: "reference"`, parity checks)
  * core/models/streamflow.py, core/update.py, core/encoders/twins_csc.py
                                        the caller: SKFlow_MF8 + SKUpdateBlock_TAM_v3 + Twins_CSC running unchanged
                                        on either the reference operators or streamflow_b200.install()'s shims
                                        (tests/test_reference_model_gpu.py, bench.py `full_model`)

Nothing under streamflow_b200/ imports oracle/_ref.  When the snapshot is absent the dependent tests skip and
bench.py falls back to the restatement in oracle/torch_port.py (`kind: "port"`).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC_DEFAULT = "/root/reference"


def make_ref(src: str = SRC_DEFAULT, force: bool = False) -> str | None:
    """Copy the reference's `core/` tree; returns the destination or None when `src` is absent."""
    core = os.path.join(src, "core")
    if not os.path.isdir(core):
        return None
    dst_core = os.path.join(DEST, "core")
    manifest = os.path.join(DEST, "MANIFEST.sha256")
    if os.path.exists(manifest) and not force:
        return DEST
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    shutil.copytree(core, dst_core, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".DS_Store"))
    for extra in ("LICENSE",):
        p = os.path.join(src, extra)
        if os.path.exists(p):
            shutil.copy2(p, os.path.join(DEST, extra))
    lines = []
    for root, _, files in sorted(os.walk(dst_core)):
        for f in sorted(files):
            p = os.path.join(root, f)
            with open(p, "rb") as fh:
                lines.append(f"{hashlib.sha256(fh.read()).hexdigest()}  {os.path.relpath(p, DEST)}")
    with open(manifest, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return DEST


def ref_core_dir() -> str | None:
    """`oracle/_ref/core` if the snapshot exists, else `/root/reference/core` if that exists, else None."""
    for cand in (os.path.join(DEST, "core"), os.path.join(SRC_DEFAULT, "core")):
        if os.path.isfile(os.path.join(cand, "corr.py")):
            return cand
    return None


if __name__ == "__main__":
    out = make_ref(force="--force" in sys.argv)
    print(out if out else f"{SRC_DEFAULT} not present: nothing to do")
