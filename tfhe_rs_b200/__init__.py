"""Import shim: the product package lives in the directory ``tfhe-rs_b200/``
(the name the project layout prescribes), which is not a valid Python
identifier.  This stub makes it importable as ``tfhe_rs_b200`` by pointing the
package path at that directory and executing its ``__init__``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tfhe-rs_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
