"""Import shim: makes ``import pytorch_simclr_b200`` resolve to the ``pytorch-simclr_b200/`` directory
(a hyphen is not a legal module name)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "pytorch-simclr_b200")]
__package__ = __name__
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
