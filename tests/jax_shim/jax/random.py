"""`jax.random` placeholder (initialisation code only; not on the hot path)."""
import numpy as _np


def key(seed):
    return _np.random.default_rng(int(seed))


PRNGKey = key


def split(k, num=2):
    return [_np.random.default_rng(int(k.integers(0, 2 ** 31))) for _ in range(num)]
