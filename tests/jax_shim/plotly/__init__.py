"""Empty stand-in: PyPIC3D/utils.py imports plotly at module level but the hot path never calls it."""
