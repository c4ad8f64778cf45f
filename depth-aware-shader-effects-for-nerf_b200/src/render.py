"""Reference module path src/render.py -> sm_100a implementation."""
from nerfw.render import volume_render  # noqa: F401
