"""Import alias for the product package.

The package directory is named `object-detection-collection-pytorch_b200/` (the layout the
project prescribes); hyphens are not importable, so this shim registers that directory as the
package `odcp_b200`.  `import odcp_b200` / `from odcp_b200.models.yolov2 import YOLOv2` work
from the repository root.
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "object-detection-collection-pytorch_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
