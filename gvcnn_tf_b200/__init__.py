"""Import shim: the real package lives in ``gvcnn-tf_b200/`` (a directory name
Python cannot import directly).  ``import gvcnn_tf_b200`` resolves its
submodules from there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gvcnn-tf_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
