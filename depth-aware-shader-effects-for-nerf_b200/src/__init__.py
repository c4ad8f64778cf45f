"""Drop-in mirror of the reference's `src` package for the hot path (ray_utils, render, models).

Put this package's parent directory BEFORE the reference checkout on sys.path.  `src.ray_utils`, `src.render` and
`src.models` then resolve to the sm_100a implementations, while every other module of the reference's `src/` (dataset,
train, post_processor, shader_system, ...) keeps resolving to the reference's own files: the reference's `src` directory
(a namespace package, it has no __init__.py) is appended to this package's search path.  `src/train.py`'s relative imports
(`from .models import NeRF`, `from .render import volume_render`) therefore bind to this repo's kernels without a single
edit in the reference tree.
"""
import os as _os
import sys as _sys

_here = _os.path.abspath(_os.path.dirname(__file__))
for _p in list(_sys.path):
    _cand = _os.path.abspath(_os.path.join(_p or ".", "src"))
    if _cand != _here and _os.path.isdir(_cand) and _cand not in __path__:
        __path__.append(_cand)
