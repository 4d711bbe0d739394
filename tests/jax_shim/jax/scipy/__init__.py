"""`jax.scipy` placeholder: the hot path does not use it (utils.py's RegularGridInterpolator belongs to the output code)."""
