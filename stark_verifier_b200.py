"""Import shim: the package directory is named ``stark-verifier_b200`` (hyphen), which the ``import``
statement cannot spell.  ``import stark_verifier_b200 as svb`` gives the same module object."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("stark-verifier_b200")
sys.modules[__name__] = _pkg
