"""Import alias: ``import tinyrec`` loads the package that lives in the
``tiny-newsrec_b200/`` directory (a hyphen is not a legal Python identifier, so the
directory cannot be imported by name).  Sub-modules resolve as ``tinyrec.<name>``."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "tiny-newsrec_b200")
_spec = importlib.util.spec_from_file_location(
    "tinyrec", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["tinyrec"] = _mod
_spec.loader.exec_module(_mod)
