"""Spin / boson matrix builders (mirror of /root/reference/src/models/common/operators.jl:4-29)."""
from __future__ import annotations

import numpy as np


def spin_operators(dimension: int):
    """Returns (S+, Sz) for S=(dimension-1)/2; state index i=1 is m=+S (operators.jl:4-20)."""
    S = (dimension - 1) / 2
    sz = np.zeros((dimension, dimension))
    splus = np.zeros((dimension, dimension))
    for i in range(1, dimension + 1):
        m = S - i + 1
        sz[i - 1, i - 1] = m
        if i < dimension:
            splus[i - 1, i] = np.sqrt((S - m + 1) * (S + m))
    return splus, sz


def bosonic_a_operator(dimension: int):
    """operators.jl:24-27"""
    return np.diag(np.sqrt(np.arange(1, dimension, dtype=np.float64)), k=1)
