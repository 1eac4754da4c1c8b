"""Importable alias of the product package.

The product lives in ``custom-yolo-implmentation_b200/`` (the name the project layout fixes);
a hyphen cannot appear in a Python import, so this one-file package points its ``__path__`` at
that directory and executes its ``__init__``.  ``import custom_yolo_implmentation_b200.model.losses``
therefore loads ``custom-yolo-implmentation_b200/model/losses.py``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "custom-yolo-implmentation_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
