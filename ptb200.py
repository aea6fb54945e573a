"""Import shim: the package directory is named after the reference (`opentk-pathtracer_b200`, not a Python
identifier), so `import ptb200` loads it under this alias."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("opentk-pathtracer_b200")
sys.modules[__name__] = _pkg
