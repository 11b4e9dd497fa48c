"""Import shim.  The package directory is named `genome-minimizer-2_b200/` (the name the
build contract fixes); hyphens are not importable, so this module loads that directory
as the package `genome_minimizer_2_b200` and replaces itself in sys.modules."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "genome-minimizer-2_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
