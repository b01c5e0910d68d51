"""Import shim: the package directory is named `polar-codes-with-bch-kernel_b200` (not a valid
Python identifier), so load it under the module name `polar_codes_with_bch_kernel_b200`.

    import pkb200; pk = pkb200.pk
"""
import importlib.util
import os
import sys

_NAME = "polar_codes_with_bch_kernel_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "polar-codes-with-bch-kernel_b200")


def _load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


pk = _load()
