"""Sort-first partition of the render target across ranks (host-side mirror of `Partition` and the composite
kernels in csrc/). The reference is single-process; this is the new multi-GPU layer.

Rank r of N owns the 8-pixel tile rows ty with (ty // stripe_h) % N == r. The gather buffer of the composite
step holds N equal chunks; chunk r is the row-major colour of rank r's stripes in ascending stripe order, padded to
the largest rank's stripe count so that ncclAllGather can run in place.
"""
from __future__ import annotations

import numpy as np

TILE = 8  # TILE_WIDTH / TILE_HEIGHT, main.c:24-25


def owner_of_tile_row(ty: int, num_ranks: int, stripe_h: int) -> int:
    return (ty // stripe_h) % num_ranks


def owned_tile_rows(height: int, num_ranks: int, rank: int, stripe_h: int):
    return [ty for ty in range(height // TILE) if owner_of_tile_row(ty, num_ranks, stripe_h) == rank]


def chunk_rows(height: int, num_ranks: int, stripe_h: int) -> int:
    """Pixel rows per chunk (same for every rank: padded to the rank with the most stripes)."""
    num_stripes = -(-(height // TILE) // stripe_h)
    local_stripes = -(-num_stripes // num_ranks)
    return local_stripes * stripe_h * TILE


def chunk_row_of(y: int, num_ranks: int, stripe_h: int) -> int:
    """Row inside its owner's chunk where image row y is stored."""
    ty = y // TILE
    local_stripe = (ty // stripe_h) // num_ranks
    return (local_stripe * stripe_h + ty % stripe_h) * TILE + y % TILE


def pack(image: np.ndarray, num_ranks: int, rank: int, stripe_h: int) -> np.ndarray:
    """Row-major image [H, W] -> this rank's chunk [chunk_rows, W] (rows it does not own are left zero)."""
    h, w = image.shape
    out = np.zeros((chunk_rows(h, num_ranks, stripe_h), w), dtype=image.dtype)
    for y in range(h):
        if owner_of_tile_row(y // TILE, num_ranks, stripe_h) == rank:
            out[chunk_row_of(y, num_ranks, stripe_h)] = image[y]
    return out


def unpack(chunks: np.ndarray, height: int, num_ranks: int, stripe_h: int) -> np.ndarray:
    """Gathered chunks [N, chunk_rows, W] -> row-major image [H, W]."""
    out = np.empty((height, chunks.shape[2]), dtype=chunks.dtype)
    for y in range(height):
        out[y] = chunks[owner_of_tile_row(y // TILE, num_ranks, stripe_h), chunk_row_of(y, num_ranks, stripe_h)]
    return out


# ---- sharded uploads of replicated geometry (bench.py e2e at N > 1; include/malevich_b200.h mlv_update_buffer_range) ----
SHARD_GRANULE = 16  # bytes; device buffers are 16-byte aligned and the kernels read them through 128-bit loads


def shard_bytes(nbytes: int, num_ranks: int) -> int:
    """Size of one rank's shard: equal for all ranks (in-place all-gather), a multiple of 16 bytes."""
    per = -(-nbytes // num_ranks)
    return -(-per // SHARD_GRANULE) * SHARD_GRANULE


def shard_range(nbytes: int, num_ranks: int, rank: int):
    """(offset, count) of the bytes of an nbytes-long host buffer that rank uploads itself; count may be 0 for the last
    ranks of a small buffer. The device buffer has capacity shard_bytes * num_ranks; what lies beyond nbytes is padding."""
    sb = shard_bytes(nbytes, num_ranks)
    lo, hi = min(rank * sb, nbytes), min((rank + 1) * sb, nbytes)
    return lo, hi - lo
