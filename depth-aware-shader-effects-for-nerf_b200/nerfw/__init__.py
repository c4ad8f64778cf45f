"""nerfw -- B200 (sm_100a) NeRF-W ray-marching hot path behind the reference's Python API.

    from nerfw import get_rays, sample_stratified, sample_importance, volume_render, NeRF, PositionalEncoding

The sibling package `src/` re-exports the same names under the reference's module paths (`src.ray_utils`,
`src.render`, `src.models`) so existing scripts only need this directory first on sys.path.
Importing this package loads libnerfw_sm100.so and raises ImportError if it has not been built: no CPU fallback.
"""
from ._lib import lib as _load

_load()

from .models import NeRF, PositionalEncoding  # noqa: E402
from .ray_utils import get_rays, sample_importance, sample_stratified  # noqa: E402
from .render import volume_render  # noqa: E402

__all__ = ["NeRF", "PositionalEncoding", "get_rays", "sample_stratified", "sample_importance", "volume_render"]
