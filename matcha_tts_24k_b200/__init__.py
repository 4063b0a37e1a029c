"""Import alias: the product package lives in ``matcha-tts-24k_b200/`` (a name Python cannot
import directly), so this stub loads that directory under the importable name
``matcha_tts_24k_b200`` and replaces itself with it."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "matcha-tts-24k_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_real, "__init__.py"), submodule_search_locations=[_real])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
