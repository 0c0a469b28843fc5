"""Distribution of a block-sparse batch over ranks (BASELINE configs[2]).

Blocks are independent units: they are sorted by estimated cost (``m * n * min(m, n)``, the
leading term of every factorization on the path) and assigned greedily to the least-loaded rank
(longest-processing-time-first).  Block sizes are heavy-tailed (log-uniform 16..512 in the
judged config), so round-robin would leave ranks idle; LPT keeps the imbalance at the 1e-8 level
for 20 000 blocks.  There is NO collective on the data path: every rank owns the inputs and
outputs of its blocks.  ``gather_block_info`` is the optional end-of-batch exchange of a small
per-block integer (kept rank after truncation, LAPACK-style info) so that every rank knows the
global structure of the result."""
from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def block_cost(m: int, n: int) -> float:
    return float(m) * float(n) * float(min(m, n))


def lpt_partition(shapes: Sequence[Tuple[int, int]], world: int) -> Tuple[np.ndarray, float]:
    """owner[i] = rank of block i; also returns max load / mean load.  Deterministic: ties broken
    by block index (stable sort) and by lowest rank, so every rank computes the same map without
    communication."""
    if world < 1:
        raise ValueError("world size must be positive")
    cost = np.array([block_cost(m, n) for m, n in shapes], dtype=np.float64)
    owner = np.zeros(len(shapes), dtype=np.int64)
    loads = np.zeros(world, dtype=np.float64)
    for i in np.argsort(-cost, kind="stable"):
        r = int(np.argmin(loads))
        owner[i] = r
        loads[r] += cost[i]
    mean = loads.mean() if len(shapes) else 0.0
    return owner, (float(loads.max() / mean) if mean > 0 else 1.0)


def my_blocks(shapes: Sequence[Tuple[int, int]], rank: int = None, world: int = None) -> List[int]:
    """Indices (ascending) of the blocks this rank owns."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    owner, _ = lpt_partition(shapes, world)
    return [int(i) for i in np.nonzero(owner == rank)[0]]


def gather_block_info(shapes: Sequence[Tuple[int, int]], local_values: Sequence[int], device=None) -> np.ndarray:
    """All ranks contribute one integer per OWNED block (in ``my_blocks`` order); returns the
    global per-block vector on every rank.  One all_reduce of ``len(shapes)`` int64 (sum of
    disjoint one-hot contributions) — off the critical path, after the factorizations."""
    nb = len(shapes)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    mine = my_blocks(shapes, rank, world)
    if len(local_values) != len(mine):
        raise ValueError(f"expected {len(mine)} local values, got {len(local_values)}")
    buf = torch.zeros(nb, dtype=torch.int64, device=device)
    if mine:
        buf[torch.as_tensor(mine, dtype=torch.int64, device=device)] = torch.as_tensor(
            [int(v) for v in local_values], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf.cpu().numpy()
