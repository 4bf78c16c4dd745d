"""Database sharding for multi-GPU search (host-side index arithmetic only).

The length-sorted database is cut into TILES of 16 consecutive sequences; tile t belongs to shard
t % num_shards (round-robin), so every GPU gets the same residue count and the same length mix.  This is
the B200 counterpart of the reference's chunk -> device distribution (sequences.c:528-557,
MICsearch.c:74-75): there chunks of <= 96 MB are dealt out dynamically and the scores scattered back by
offset (MICsearch.c:332-335); here the deal is static and every hit carries its global index.
"""
from __future__ import annotations

import numpy as np

TILE = 16


def local_to_global(local_index, shard: int, num_shards: int):
    """Position in the whole length-sorted database of a shard-local sequence index."""
    local_index = np.asarray(local_index, dtype=np.int64)
    return (local_index // TILE * num_shards + shard) * TILE + local_index % TILE


def shard_sequences(n_total: int, shard: int, num_shards: int) -> np.ndarray:
    """Global indices of the sequences shard `shard` holds, in local order."""
    tiles = np.arange(shard, (n_total + TILE - 1) // TILE, num_shards, dtype=np.int64)
    idx = (tiles[:, None] * TILE + np.arange(TILE)[None, :]).reshape(-1)
    return idx[idx < n_total]


def extract_shard(lengths: np.ndarray, codes: np.ndarray, shard: int, num_shards: int):
    """(local_lengths, local_codes, global_index) of one shard -- what swg_gpu_load_db_shard takes."""
    lengths = np.asarray(lengths)
    gidx = shard_sequences(len(lengths), shard, num_shards)
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(lengths.astype(np.int64), out=off[1:])
    ll = lengths[gidx]
    starts = off[gidx]
    loff = np.zeros(len(gidx) + 1, dtype=np.int64)
    np.cumsum(ll.astype(np.int64), out=loff[1:])
    take = np.repeat(starts - loff[:-1], ll.astype(np.int64)) + np.arange(loff[-1], dtype=np.int64)
    return ll.astype(np.uint16), np.asarray(codes)[take], gidx


def make_keys(scores: np.ndarray, global_index: np.ndarray) -> np.ndarray:
    """SWG_KEY(score, index) = score << 32 | index (include/swimm_gpu.h)."""
    return (scores.astype(np.uint64) << np.uint64(32)) | global_index.astype(np.uint64)


def top_keys(keys: np.ndarray, top: int) -> np.ndarray:
    """The `top` largest keys, descending, zero padded -- the order of the reference's hit list."""
    k = np.sort(keys.astype(np.uint64))[::-1][:top]
    out = np.zeros(top, dtype=np.uint64)
    out[:len(k)] = k
    return out
