"""Multi-GPU proving: one process per GPU, MSMs sharded by index range, one tiny exchange per proof.

The Groth16 prover's five MSMs are sums over independent (scalar, point) pairs, so every rank keeps the
[rank/world) slice of each query vector resident (`zkb_pk_load(rank, world)`) and produces five partial
sums.  Elliptic-curve addition is not an NCCL reduction op, so the partial blobs (a few hundred bytes per
rank) are all-gathered with `torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests) and the
final combination runs once (rank `dst`).  witness_map is replicated: every rank needs all of h for its
h-query slice and the transform is far cheaper than the exchange an NTT sharding would need
(SURVEY.md §8e).
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def gather_partials(partial: np.ndarray, group=None, device=None) -> np.ndarray:
    """all_gather of the per-rank partial-sum blobs; returns the world * partial_bytes concatenation."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.uint8))
    if device is not None:
        mine = mine.to(device)
    out = torch.empty(world * mine.numel(), dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out.cpu().numpy()


def prove_sharded(session, z: Optional[np.ndarray], r: int, s: int, finalize_session=None, dst: int = 0, group=None,
                  device=None) -> Optional[bytes]:
    """One proof across all ranks of `group`.  `session` holds this rank's key shard; `finalize_session`
    (rank `dst` only) may be any session of the same key.  Returns the raw proof on rank `dst`, None elsewhere."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    partial = session.prove_partial(z)
    allp = gather_partials(partial, group, device)
    if rank != dst:
        return None
    return (finalize_session or session).finalize(allp, world, r, s)
