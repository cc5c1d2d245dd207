"""
Import alias: the product package lives in a directory whose (mandated) name contains hyphens and
therefore cannot appear in an ``import`` statement.  ``import lrc_b200`` loads it and makes this name --
and every sub-module name under it -- refer to the very same module objects.
"""
import importlib
import os
import sys

PACKAGE_DIR_NAME = "indoor-point-cloud-datasets-controllable-generation-method-for-mobile-robots-3d-scene-perception_b200"
_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)

_pkg = importlib.import_module(PACKAGE_DIR_NAME)
for _name, _mod in list(sys.modules.items()):
    if _name == PACKAGE_DIR_NAME or _name.startswith(PACKAGE_DIR_NAME + "."):
        sys.modules[__name__ + _name[len(PACKAGE_DIR_NAME):]] = _mod
