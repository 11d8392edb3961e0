"""Import shim: `import zkdcap_b200` == the package in ./zk-dcap-verifier_b200/ (hyphenated name)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("zk-dcap-verifier_b200")
sys.modules[__name__] = _pkg
