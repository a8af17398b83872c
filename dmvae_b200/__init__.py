"""Importable alias of the product package, whose directory name (``deep-mixture-vae_b200``) is not a Python
identifier.  ``import dmvae_b200`` resolves every submodule inside that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "deep-mixture-vae_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
