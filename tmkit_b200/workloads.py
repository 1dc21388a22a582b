"""Synthetic inputs of the benchmark configurations (BASELINE.json configs).

All joint values are drawn as float32 and promoted to float64, so the float and double
entry points (and the oracle) see bit-identical numbers.
"""
from __future__ import annotations

import numpy as np

from . import scenes as S

OMPL_SEGMENT_FRACTION = 0.01  # longestValidSegmentFraction
OMPL_RANGE_FRACTION = 0.2     # RRT-Connect range


def arm_limit_table(names):
    """limits of right_/left_ arm joints by name (A.6)."""
    base = {suffix: lim for (suffix, _, _, lim) in S._ARM_CHAIN}
    return np.asarray([base[n.split("_", 1)[1]] for n in names], dtype=np.float64)


def max_extent(limits: np.ndarray) -> float:
    return float(np.linalg.norm(limits[:, 1] - limits[:, 0]))


def seg_len(limits: np.ndarray) -> float:
    return OMPL_SEGMENT_FRACTION * max_extent(limits)


def random_configs(limits: np.ndarray, n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    u = rng.random((n, len(limits)), dtype=np.float32)
    lo = limits[:, 0].astype(np.float32)
    hi = limits[:, 1].astype(np.float32)
    q = lo + (hi - lo) * u
    q = np.clip(q, lo, hi)
    return q.astype(np.float64)


def random_edges(limits: np.ndarray, n: int, seed: int, kind: str = "rrt"):
    """kind 'rrt': q1 = clip(q0 + L u), u uniform on the sphere, L ~ U(0, 0.2 maxExtent);
    kind 'uniform': both end points uniform in the limits."""
    rng = np.random.default_rng(seed + 1000)
    q0 = random_configs(limits, n, seed + 17)
    if kind == "uniform":
        q1 = random_configs(limits, n, seed + 31)
    else:
        u = rng.normal(size=(n, len(limits)))
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        L = rng.uniform(0, OMPL_RANGE_FRACTION * max_extent(limits), size=(n, 1))
        q1 = np.clip(q0 + L * u, limits[:, 0], limits[:, 1]).astype(np.float32).astype(np.float64)
    return q0, q1
