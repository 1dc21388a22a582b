"""Sharding of a configuration / edge batch over the GPUs of one box.

The path partitions into independent units, so there is exactly one exchange step: every rank
checks a contiguous slice of the index space [0, n) and the per-slice validity bitmaps are
all-gathered (NCCL over NVLink on the GPU box; gloo in the CPU tests).  Slices are aligned to
``ALIGN`` units so that every rank owns whole 32-bit bitmap words and the gathered buffer is the
global bitmap without any bit shifting.  The scene image (KBs) is replicated on every rank.

This module holds only index arithmetic and the collective; the validity check itself is the
C-ABI call (``aa_rx_cl_check_batch_dev`` / ``aa_rx_cl_check_edges_dev``) the caller passes in.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np

ALIGN = 1024  # units; a multiple of 32 (bitmap word) and of the kernels' 256-configuration tile


def _per_rank(n: int, world: int, align: int) -> int:
    per = -(-n // world)
    return -(-per // align) * align


def slice_bounds(n: int, world: int, align: int = ALIGN) -> List[Tuple[int, int]]:
    """Contiguous slices [lo, hi) of [0, n): equal, align-multiple lengths except the tail."""
    if world < 1 or n < 0 or align % 32:
        raise ValueError("bad sharding arguments")
    per = _per_rank(n, world, align)
    out = []
    for r in range(world):
        lo = min(n, r * per)
        out.append((lo, min(n, lo + per)))
    return out


def words_per_rank(n: int, world: int, align: int = ALIGN) -> int:
    return _per_rank(n, world, align) // 32


def check_sharded(n: int, rank: int, world: int, check_slice: Callable[[int, int, "object"], None], make_words,
                  all_gather) -> "object":
    """Run ``check_slice(lo, hi, local_words)`` on this rank's slice and all-gather the words.

    ``make_words(k)`` allocates k zeroed 32-bit words (a torch tensor on the rank's device);
    ``all_gather(out, local)`` is ``torch.distributed.all_gather_into_tensor``.  Returns the
    global bitmap: ceil(n / 32) words, bit (c & 31) of word c >> 5 = configuration c is valid.
    """
    bounds = slice_bounds(n, world)
    wpr = words_per_rank(n, world)
    local = make_words(wpr)
    lo, hi = bounds[rank]
    if hi > lo:
        check_slice(lo, hi, local)
    if world == 1:
        return local[: (n + 31) // 32]
    out = make_words(wpr * world)
    all_gather(out, local)
    return out[: (n + 31) // 32]


def unpack(words: np.ndarray, n: int) -> np.ndarray:
    return np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")[:n].astype(bool)


def pack(valid: np.ndarray, n_words: int) -> np.ndarray:
    b = np.zeros(n_words * 32, dtype=np.uint8)
    b[: len(valid)] = valid
    return np.packbits(b, bitorder="little").view(np.uint32)
