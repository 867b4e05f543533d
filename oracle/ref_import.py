"""TEST INFRASTRUCTURE ONLY - not part of the product path.

Import shim for the *unmodified* reference classes.

The reference's hot-path modules (`protnote/models/ProtNote.py`,
`protnote/models/protein_encoders.py`, `protnote/data/datasets.py:535-569`)
import cleanly once a handful of third-party modules that the hot path never
touches (`loralib`, `blosum`, `wget`, `Bio.*`) are stubbed in `sys.modules`
(SURVEY.md section 8c).  Where they are imported from:

  * `/root/reference` in the build container - used by `oracle/make_golden*.py`
    and by the CPU tests that pin the travelling oracle
    (`oracle/protnote_oracle.py`) against the real reference;
  * `oracle/_ref/` on the GPU box - a byte-for-byte copy of the reference's own
    `protnote/` package staged by `oracle/build_ref.py` (git-ignored: the
    reference's sources never enter this repository; shipped with the snapshot
    like a built .so).  Only `bench.py`'s CPU arm (`--impl reference`,
    `cpu_baseline`) uses it there, to time the reference's own classes.

Nothing under `-m gpu` or `smoke()` imports this file, and the product package
`protnote_b200` never does.
"""
from __future__ import annotations

import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _default_root() -> str:
    """/root/reference in the build container; the copy staged by oracle/build_ref.py (oracle/_ref, git-ignored, travels
    with the gpurun snapshot) on the GPU box."""
    env = os.environ.get("PROTNOTE_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir(os.path.join("/root/reference", "protnote", "models")):
        return "/root/reference"
    return _STAGED


REFERENCE_ROOT = _default_root()

_STUBS = (
    "loralib",
    "blosum",
    "wget",
    "Bio",
    "Bio.SeqIO",
    "Bio.SeqIO.FastaIO",
    "Bio.ExPASy",
    "Bio.ExPASy.Enzyme",
    "Bio.Seq",
    "Bio.SeqRecord",
    "obonet",
    "torchmetrics",
    "torcheval",
)


class _Anything(types.ModuleType):
    """A module whose every attribute is a harmless placeholder class."""

    def __getattr__(self, name):  # pragma: no cover - trivial
        if name.startswith("__"):
            raise AttributeError(name)
        placeholder = type(name, (), {})
        setattr(self, name, placeholder)
        return placeholder


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "protnote", "models"))


def import_reference():
    """Returns (ProtNote, ProteInfer, set_padding_to_sentinel) from the reference tree."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mod = _Anything(name)
                mod.__path__ = []  # behave like a package
                sys.modules[name] = mod
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from protnote.models.ProtNote import ProtNote  # type: ignore
    from protnote.models.protein_encoders import ProteInfer  # type: ignore
    from protnote.data.datasets import set_padding_to_sentinel  # type: ignore

    return ProtNote, ProteInfer, set_padding_to_sentinel
