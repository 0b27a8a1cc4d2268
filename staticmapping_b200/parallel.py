"""Multi-GPU plumbing for batches of independent scan-pair alignments.

The reference fans independent alignments out to a thread pool / TBB tasks
(map_builder.cc:655,706-708; loop_detector.cc:224-228).  Here the same units are sharded
across ranks (one process per GPU, no data-path collective: clouds never cross GPUs) and
the resulting poses are exchanged with ONE all-gather: 16 doubles (4x4, column-major) +
score per pair.  Works with the nccl backend on GPUs and with gloo on CPU (tests)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

POSE_RECORD = 17  # 16 doubles of the 4x4 (column-major) + fitness score


def shard_pairs(n_pairs: int, rank: int, world: int) -> List[int]:
    """Contiguous block partition: rank r owns pairs [lo, hi) with sizes differing by <= 1."""
    if world <= 0 or not (0 <= rank < world) or n_pairs < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_pairs, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return list(range(lo, hi))


def pack_poses(results: Sequence[np.ndarray], scores: Sequence[float]) -> np.ndarray:
    """[(4,4) row-major numpy] + scores -> (n, 17) float64 records (pose column-major)."""
    out = np.zeros((len(results), POSE_RECORD), dtype=np.float64)
    for i, (T, s) in enumerate(zip(results, scores)):
        out[i, :16] = np.asarray(T, dtype=np.float64).T.ravel()
        out[i, 16] = s
    return out


def unpack_poses(records: np.ndarray):
    rec = np.asarray(records, dtype=np.float64).reshape(-1, POSE_RECORD)
    return [r[:16].reshape(4, 4).T.copy() for r in rec], [float(r[16]) for r in rec]


def allgather_poses(local_records: np.ndarray, n_pairs: int, device=None) -> np.ndarray:
    """All ranks end with the (n_pairs, 17) table in global pair order.  One all_gather of
    equal-size padded blocks (ceil(n_pairs / world) records per rank)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(local_records, dtype=np.float64).reshape(-1, POSE_RECORD)
    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-n_pairs // world)
    buf = torch.zeros((per, POSE_RECORD), dtype=torch.float64, device=device)
    mine = shard_pairs(n_pairs, rank, world)
    if len(mine):
        buf[: len(mine)] = torch.from_numpy(
            np.asarray(local_records, dtype=np.float64).reshape(len(mine), POSE_RECORD)).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    table = np.zeros((n_pairs, POSE_RECORD), dtype=np.float64)
    for r in range(world):
        idx = shard_pairs(n_pairs, r, world)
        if idx:
            table[idx[0]: idx[-1] + 1] = out[r][: len(idx)].cpu().numpy()
    return table
