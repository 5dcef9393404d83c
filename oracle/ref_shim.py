"""Live-reference import shim (TEST INFRASTRUCTURE, container-only).

Imports the UNMODIFIED reference from /root/reference so that the restatement in
``oracle/restatement.py`` / ``oracle/minsum_oracle.c`` can be pinned against it and so that
``tests/golden/make_golden.py`` can generate the committed golden vectors.

/root/reference does not exist on the GPU box; nothing that runs there may import this module
(``available()`` returns False there and callers must skip).

Two value-neutral accommodations, both described in SURVEY.md section 8c:
  * matplotlib is not installed and ``ldpc_decoder.py:18`` imports it at module top, so stub modules
    are injected into ``sys.modules`` (no plotting is ever called on the decode path);
  * ``cache_degrees=True`` replaces ``LDPCCode.check_node_degrees`` / ``variable_node_degrees``
    (``ldpc_decoder.py:38-54``: recomputed with one ``np.sum`` per row/column on *every access*)
    by per-instance cached dicts holding the same values.  Outputs are bit-identical; the N-2D /
    W-RCQ decoders just stop being O(T*E*(m+n)).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LDPC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ldpc_decoder.py"))


_loaded = None


def load(cache_degrees: bool = True):
    """Return a namespace with the reference modules (ldpc_decoder, neural_minsum_decoder,
    neural_2d_decoder, rcq_decoder)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present (expected on the GPU box); skip")
    import logging

    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            sys.modules[name] = mod
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    # The reference modules are flat files that import each other by bare name.  Our own package
    # mirrors those names, so import the reference under a private prefix to avoid any clash.
    import importlib.util

    ns = types.SimpleNamespace()
    saved = {}
    names = ["ldpc_decoder", "neural_minsum_decoder", "neural_2d_decoder", "rcq_decoder"]
    for nm in names:
        saved[nm] = sys.modules.pop(nm, None)
    try:
        for nm in names:
            spec = importlib.util.spec_from_file_location(nm, os.path.join(REFERENCE_ROOT, nm + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[nm] = mod  # later reference files do ``from ldpc_decoder import LDPCCode``
            spec.loader.exec_module(mod)
            setattr(ns, nm, mod)
    finally:
        for nm in names:
            sys.modules.pop(nm, None)
            if saved[nm] is not None:
                sys.modules[nm] = saved[nm]
    logging.getLogger().setLevel(logging.WARNING)
    for nm in names:
        logging.getLogger(nm).setLevel(logging.WARNING)

    if cache_degrees:
        import numpy as np

        LDPCCode = ns.ldpc_decoder.LDPCCode

        def _cnd(self):
            c = self.__dict__.get("_cnd_cache")
            if c is None:
                c = {i: int(d) for i, d in enumerate(np.sum(self.H, axis=1))}
                self.__dict__["_cnd_cache"] = c
            return c

        def _vnd(self):
            c = self.__dict__.get("_vnd_cache")
            if c is None:
                c = {j: int(d) for j, d in enumerate(np.sum(self.H, axis=0))}
                self.__dict__["_vnd_cache"] = c
            return c

        LDPCCode.check_node_degrees = property(_cnd)
        LDPCCode.variable_node_degrees = property(_vnd)
    _loaded = ns
    return ns
