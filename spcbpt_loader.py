"""Import helper: the package directory is named `spcbpt-optix7_b200` (hyphen, fixed by the build
contract), which Python cannot import by name.  load() registers it as `spcbpt_optix7_b200`."""
import importlib.util
import os
import sys

_NAME = "spcbpt_optix7_b200"
_ROOT = os.path.dirname(os.path.abspath(__file__))


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    pkg_dir = os.path.join(_ROOT, "spcbpt-optix7_b200")
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


def load_oracle():
    """The CPU oracle binding (test infrastructure; only tests/, smoke() and bench.py's CPU legs)."""
    name = "spc_oracle_py"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(_ROOT, "oracle", "orc_py.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
