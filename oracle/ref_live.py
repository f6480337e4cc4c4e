"""ORACLE — TEST INFRASTRUCTURE ONLY.

Imports the *live* reference modules from `/root/reference/mvs/mvs_cas/models` (never copies
them).  Only possible in the authoring container: the GPU box has no `/root/reference`, so
nothing that runs there (`-m gpu` tests, `smoke()`, `bench.py`) may call `load()`; tests that use
it are skipped when `available()` is False.  Used to (a) pin `oracle/sweep_torch.py` and
(b) generate the golden vectors under `tests/golden/` (`oracle/make_golden.py`).
"""
from __future__ import annotations

import importlib
import os
import sys
import types
import warnings

REF_ROOT = os.path.join(os.environ.get("DEEP3D_REFERENCE_ROOT", "/root/reference"), "mvs", "mvs_cas")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "module.py"))


def load() -> types.SimpleNamespace:
    """Return a namespace with the reference's hot-path modules imported live."""
    if not available():
        raise RuntimeError("live reference not present at " + REF_ROOT)
    import torch

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if not torch.cuda.is_available():
        # the reference hard-codes .cuda() in its recurrent-state initialisers
        # (adamvs.py:175-176, 451-462; msrednet.py:158-161, 390-398)
        torch.Tensor.cuda = lambda self, *a, **k: self  # type: ignore[assignment]
    warnings.filterwarnings("ignore", message=".*indexing argument.*")
    ns = types.SimpleNamespace()
    for name in ("module", "cas_mvsnet", "adamvs", "msrednet", "ucsnet"):
        setattr(ns, name, importlib.import_module("models." + name))
    return ns
