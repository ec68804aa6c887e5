"""Import shim: `import b200gs` loads the package directory `wgpu-3dgs-viewer-app_b200/`
(whose name is not a valid Python identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wgpu-3dgs-viewer-app_b200")
_spec = importlib.util.spec_from_file_location("wgpu_3dgs_viewer_app_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[_spec.name] = _mod
_spec.loader.exec_module(_mod)
sys.modules[__name__] = _mod
