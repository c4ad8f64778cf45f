"""Reference module path src/models.py -> sm_100a implementation (NeRF, PositionalEncoding)."""
from nerfw.models import NeRF, PositionalEncoding  # noqa: F401
