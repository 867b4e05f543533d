"""TEST INFRASTRUCTURE ONLY - recipe that makes the UNMODIFIED reference travel to the GPU box.

The reference is pure Python (no native code to compile), so "building" it means staging its own package where the
bench's reference arm can import it on a machine that has no /root/reference:

    python -m oracle.build_ref            # /root/reference/protnote -> oracle/_ref/protnote   (byte-for-byte copy)

`oracle/_ref/` is git-ignored (the reference's sources never enter this repository's history) but not gpurun-ignored,
so the staged copy ships with the snapshot like a built .so does.  `oracle/ref_import.py` then imports the reference's
real `ProtNote` / `ProteInfer` classes from there (same stub shim for the third-party modules the hot path never
touches), `bench.py --impl reference` and `cpu_baseline` time THOSE classes (`"kind": "reference"`) and fall back to the
port in `oracle/protnote_oracle.py` (`"kind": "port"`) only when the staged copy is absent.
`__graft_entry__.build()` runs this recipe whenever /root/reference is present.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("PROTNOTE_REFERENCE_SRC", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")


def build_ref(verbose: bool = True) -> bool:
    src = os.path.join(REF_SRC, "protnote")
    if not os.path.isdir(src):
        if verbose:
            print(f"oracle/_ref: no reference tree at {REF_SRC}; nothing staged")
        return False
    dst = os.path.join(REF_DST, "protnote")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(REF_DST, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    n = sum(len(files) for _, _, files in os.walk(dst))
    with open(os.path.join(REF_DST, "PROVENANCE.txt"), "w") as f:
        f.write(f"byte-for-byte copy of {src} ({n} files), staged by oracle/build_ref.py; not part of the repository\n")
    if verbose:
        print(f"oracle/_ref: staged {n} files from {src}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build_ref() else 1)
