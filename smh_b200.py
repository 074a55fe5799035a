"""Import shim: the package directory is named `squad-mortar-helper_b200` (not a Python
identifier), so it is loaded here under the module name `squad_mortar_helper_b200`.

    import smh_b200
    vision = smh_b200.vision.B200Vision.init()
"""
import importlib.util
import sys
from pathlib import Path

_PKG = Path(__file__).resolve().parent / "squad-mortar-helper_b200"
_NAME = "squad_mortar_helper_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, _PKG / "__init__.py",
                                                   submodule_search_locations=[str(_PKG)])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

pkg = sys.modules[_NAME]
from squad_mortar_helper_b200 import build, vision, capi, synth, shard  # noqa: E402,F401
