#!/usr/bin/env python
"""Recipe for oracle/_ref/: the UNMODIFIED reference tree where bench.py's CPU arm can import it on the GPU box.

    python oracle/make_ref.py            # run in the build container (needs /root/reference); idempotent

TEST / BENCH INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, but `gpurun` ships git-ignored files of
/root/repo, so the reference's own Python modules of the hot path (`lib/` and the experiment yamls, byte-for-byte copies,
nothing edited) are placed under oracle/_ref/ -- git-ignored, never part of the history or of the product package.
`bench.py --impl reference` and the `cpu_baseline` leg then time `UVLTrack.forward_test` of these unmodified modules
through oracle/ref_shim.py (`cpu_baseline.kind = "reference"`); without oracle/_ref they fall back to the restatement
(oracle/uvlt_oracle_torch.py, `kind = "port"`).  __graft_entry__.build() calls this when /root/reference is present.
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("UVLT_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")


def main() -> int:
    if not os.path.isdir(os.path.join(SRC, "lib")):
        print(f"make_ref: {SRC}/lib not found (not the build container?) -- nothing done")
        return 1
    os.makedirs(DST, exist_ok=True)
    for sub in ("lib", os.path.join("experiments", "uvltrack")):
        dst = os.path.join(DST, sub)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, sub), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.so", "*.o"))
    n = sum(len(f) for _, _, f in os.walk(DST))
    print(f"make_ref: {n} files of the unmodified reference under {os.path.relpath(DST, ROOT)}/")
    return 0


if __name__ == "__main__":
    sys.exit(main())
