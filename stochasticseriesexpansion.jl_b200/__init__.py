"""B200-native sweep backend for stochastic-series-expansion QMC (drop-in for the hot path of
lukas-weber/StochasticSeriesExpansion.jl: diagonal update -> vertex list -> worm update -> op-string
estimators).  Import as `sse_b200` (see /sse_b200.py at the repo root; the directory name contains a
dot and cannot be imported by name).

Host side (this package, Python because Julia is not available in the build image) mirrors the
reference's plugin interfaces; the sweep itself runs in hand-written sm_100a CUDA kernels behind the
C ABI declared in include/sse_b200.h (csrc/).  There is no CPU fallback: every product call goes
through libsse_b200.so and fails loudly if it is missing.
"""
from .util import join_idx, split_idx  # noqa: F401
from .lattice import Lattice, UCBond, UCSite, UnitCell, UnitCells, neel_vector  # noqa: F401
from .operators import spin_operators, bosonic_a_operator  # noqa: F401
from .vertex_data import VertexData, make_vertex_data  # noqa: F401
from .sse_data import SSEBond, SSEData, SSESite  # noqa: F401
from .estimators import (  # noqa: F401
    MagnetizationEstimator,
    all_magnetization_estimators,
    magnetization_estimator_standard_prefix,
)
from .magnet import MagnetModel  # noqa: F401
from .cluster import ClusterBases, ClusterBasis, ClusterModel  # noqa: F401
