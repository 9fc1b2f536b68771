"""Import shim: the package directory is named after the reference repo (`mg-cfd-app-plain_b200`, not a valid Python
identifier), so `import mgcfd_b200` loads it from its path and re-exports its public names."""
import importlib.util as _u
import os as _os
import sys as _sys

_path = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "mg-cfd-app-plain_b200", "__init__.py")
_spec = _u.spec_from_file_location("mgcfd_b200_pkg", _path)
_mod = _u.module_from_spec(_spec)
_sys.modules["mgcfd_b200_pkg"] = _mod
_spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
