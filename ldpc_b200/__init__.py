"""Short import alias for the package directory whose name is not a Python identifier:
``implementation-of-neural-ldpc-decoders-with-degree-specific-weight-sharing-and-rcq-quantization_b200``.

``import ldpc_b200`` gives that package (and ``ldpc_b200.rcq_decoder`` etc. its submodules)."""
import importlib
import os
import sys

_REAL = "implementation-of-neural-ldpc-decoders-with-degree-specific-weight-sharing-and-rcq-quantization_b200"
_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module(_REAL)
sys.modules[__name__] = _pkg
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_REAL + "."):
        sys.modules[__name__ + _name[len(_REAL):]] = _mod
