"""Import the UNMODIFIED reference model source for oracle pinning.  TEST INFRASTRUCTURE ONLY.

The reference hard-codes ``'cuda:0'`` in six places (GPTST.py:68,112,305,316,389,400), so it
cannot run on a CPU-only host as shipped.  We read the source where it lies, substitute the
device literal in memory and ``exec`` it into a private module.  Nothing is copied to disk.

Search order for the source: ``$GPTST_REFERENCE_ROOT``, ``/root/reference`` (build container),
``<repo>/baseline/_ref`` (driver-provided install that travels to the GPU box).
"""
from __future__ import annotations

import os
import types
from typing import Optional

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REL = os.path.join("model", "Pretrain_model", "GPTST.py")


def reference_root() -> Optional[str]:
    for root in (os.environ.get("GPTST_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if root and os.path.isfile(os.path.join(root, _REL)):
            return root
    return None


def load_reference(device: str = "cpu", root: Optional[str] = None) -> types.ModuleType:
    root = root or reference_root()
    if root is None:
        raise FileNotFoundError("reference GPTST.py not found (no /root/reference, no baseline/_ref)")
    path = os.path.join(root, _REL)
    with open(path, "r") as fh:
        src = fh.read()
    n = src.count("'cuda:0'")
    if n != 6:
        raise RuntimeError(f"expected 6 'cuda:0' literals in {path}, found {n}")
    src = src.replace("'cuda:0'", repr(device))
    mod = types.ModuleType("gptst_reference")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def checkpoint_path(dataset: str = "PEMS08", root: Optional[str] = None) -> Optional[str]:
    for r in (root, os.environ.get("GPTST_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if r:
            p = os.path.join(r, "model", "SAVE", dataset, "GPTST_ada.pth")
            if os.path.isfile(p):
                return p
    return None
