"""Mixed-radix index helpers and the bit-packed operator codes of the reference.

Host-side mirror of /root/reference/src/util.jl:3-23 (split_idx / join_idx, first index
fastest, 1-based), /root/reference/src/opercode.jl:2-71 (VertexCode / OperCode, UInt64) and
/root/reference/src/worms.jl:4-7.  These are the encodings used at the checkpoint / C-ABI
boundary (`sse_get_state` / `sse_set_state` exchange reference-format UInt64 op codes).
"""
from __future__ import annotations

import numpy as np

# ---- util.jl ---------------------------------------------------------------------------------


def split_idx(dims, idx: int):
    """1-based compound index -> tuple of 1-based digits, first digit fastest (util.jl:3-13)."""
    idx -= 1
    out = []
    for d in dims:
        out.append(idx % d + 1)
        idx //= d
    return tuple(out)


def join_idx(dims, idxs) -> int:
    """Inverse of split_idx (util.jl:15-23)."""
    r = 0
    for idx, d in zip(reversed(tuple(idxs)), reversed(tuple(dims))):
        r = r * d + (idx - 1)
    return r + 1


# ---- worms.jl --------------------------------------------------------------------------------


def worm_action(worm: int, state: int, basis_size: int) -> int:
    return (state + worm - 1) % basis_size + 1


def worm_inverse(worm: int, basis_size: int) -> int:
    return basis_size - worm


def worm_count(basis_size: int) -> int:
    return basis_size - 1


# ---- opercode.jl -----------------------------------------------------------------------------

VERTEX_CODE_MAXBITS = 8 * 3 + 1
INVALID_VERTEX_CODE = (1 << VERTEX_CODE_MAXBITS) + 1
IDENTITY_OPERCODE = 0


def vertex_code(diagonal: bool, vertex_idx: int) -> int:
    """VertexCode(diagonal, vertex_idx) (opercode.jl:18-21); vertex_idx is 1-based."""
    return int(bool(diagonal)) | (int(vertex_idx) << 1)


def vertex_isdiagonal(v: int) -> bool:
    return bool(v & 1)


def vertex_isinvalid(v: int) -> bool:
    return v >= (1 << VERTEX_CODE_MAXBITS)


def vertex_idx(v: int) -> int:
    return v >> 1


def opercode(bond: int, vcode: int) -> int:
    """OperCode(bond, vertex) (opercode.jl:43-47); bond is 1-based."""
    return 1 | (vcode << 1) | (int(bond) << (1 + VERTEX_CODE_MAXBITS))


def op_bond(op: int) -> int:
    return op >> (1 + VERTEX_CODE_MAXBITS)


def op_vertex(op: int) -> int:
    """get_vertex (opercode.jl:61-62): note only 24 bits of the vertex field survive."""
    return (op & ((1 << VERTEX_CODE_MAXBITS) - 1)) >> 1


def op_isidentity(op: int) -> bool:
    return op == 0


def op_isdiagonal(op: int) -> bool:
    return vertex_isdiagonal(op_vertex(op))


def opercodes_array(bonds, vcodes) -> np.ndarray:
    """Vectorised OperCode(bond, vertex) -> uint64 array."""
    b = np.asarray(bonds, dtype=np.uint64)
    v = np.asarray(vcodes, dtype=np.uint64)
    return np.uint64(1) | (v << np.uint64(1)) | (b << np.uint64(1 + VERTEX_CODE_MAXBITS))
