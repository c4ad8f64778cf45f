"""Drop-in mirror of the reference's `src` package for the hot path (ray_utils, render, models)."""
