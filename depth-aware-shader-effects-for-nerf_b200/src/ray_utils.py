"""Reference module path src/ray_utils.py -> sm_100a implementation."""
from nerfw.ray_utils import get_rays, sample_importance, sample_stratified  # noqa: F401
