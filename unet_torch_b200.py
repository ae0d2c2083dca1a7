"""Import alias: the package directory is `unet-torch_b200/` (not a valid identifier), so `import unet_torch_b200`
lands here and this module replaces itself with that package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "unet-torch_b200")
_spec = importlib.util.spec_from_file_location(
    "unet_torch_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["unet_torch_b200"] = _mod
_spec.loader.exec_module(_mod)
