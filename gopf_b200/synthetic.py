"""Deterministic synthetic inputs for the benchmark and the parity tests
(SURVEY.md 8d 'Synthetic inputs').  Go's math/rand stream cannot be reproduced
here, so initial fields come from a counter-based SplitMix64: u[j] depends only
on (seed, j), hence any slab of a sharded grid can be generated on its own rank.
"""
from __future__ import annotations

import numpy as np

_GOLDEN = 0x9E3779B97F4A7C15


def splitmix64_uniform(seed: int, n: int, offset: int = 0) -> np.ndarray:
    """u[j] in [0, 1) for j = offset .. offset+n-1 (top 53 bits of SplitMix64)."""
    with np.errstate(over="ignore"):
        j = np.arange(offset + 1, offset + n + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + j * np.uint64(_GOLDEN)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def cahn_hilliard_initial(n_nodes: int, seed: int = 0, offset: int = 0, out: np.ndarray | None = None) -> np.ndarray:
    """conc = 2u - 1 (examples/cahnHilliard/main.go:19-23 with the synthetic stream)."""
    if out is None:
        out = np.empty(n_nodes, dtype=np.complex128)
    chunk = 1 << 24
    for s in range(0, n_nodes, chunk):
        m = min(chunk, n_nodes - s)
        out[s:s + m] = 2.0 * splitmix64_uniform(seed, m, offset + s) - 1.0
    return out


CAHN_HILLIARD_EQUATION = "dconc/dt = LAP conc^3 + m1*LAP conc + m1*gamma*LAP^2 conc"  # cahnHilliard/main.go:33
CAHN_HILLIARD_DT = 0.1      # main.go:14
CAHN_HILLIARD_GAMMA = 2.0   # main.go:26
CAHN_HILLIARD_M1 = -1.0     # main.go:27
