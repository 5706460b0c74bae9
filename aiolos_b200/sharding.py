"""Multi-GPU plumbing: ensemble members are independent, so columns are partitioned over the ranks of one box and
nothing is exchanged during the run.  The only collective is the optional gather of per-column diagnostics."""
from __future__ import annotations

from typing import Dict, List

import numpy as np


def columns_of_rank(n_total: int, rank: int, world: int) -> List[int]:
    """Interleaved partition (column k -> rank k mod world): in a flux sweep the stiff, high-flux columns take more steps,
    interleaving spreads them evenly."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    return list(range(rank, n_total, world))


def gather_diagnostics(local: Dict[str, np.ndarray], n_total: int, rank: int, world: int, group=None) -> Dict[str, np.ndarray] | None:
    """Gather per-column arrays (one entry per local column, in columns_of_rank order) onto rank 0 in global column
    order.  Uses torch.distributed (NCCL or gloo); with world == 1 it is a plain reordering."""
    cols = columns_of_rank(n_total, rank, world)
    for k, v in local.items():
        if len(v) != len(cols):
            raise ValueError(f"{k}: {len(v)} entries for {len(cols)} local columns")
    if world == 1:
        return {k: np.asarray(v).copy() for k, v in local.items()}
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    nmax = (n_total + world - 1) // world
    out = {}
    for k in sorted(local):
        v = np.asarray(local[k], dtype=np.float64)
        buf = torch.zeros(nmax, dtype=torch.float64, device=dev)
        buf[: len(v)] = torch.from_numpy(v).to(dev)
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
        if rank == 0:
            full = np.empty(n_total)
            for r in range(world):
                idx = columns_of_rank(n_total, r, world)
                full[idx] = parts[r][: len(idx)].cpu().numpy()
            out[k] = full
    return out if rank == 0 else None
