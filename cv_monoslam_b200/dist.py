"""Multi-GPU plumbing: one process per GPU, filters sharded contiguously, no collective on the step path.

Filters are independent, so the only exchange in the system is one small all-reduce(sum) of
the Monte-Carlo statistics vector produced by srukf_stats, plus the max-over-ranks of the timed region.
torch.distributed is used for the plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np


@dataclass
class Ctx:
    rank: int = 0
    world: int = 1
    local_rank: int = 0
    backend: str | None = None


def shard_range(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition of B filters over `world` ranks."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init(backend: str | None = None) -> Ctx:
    """Initialise from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return Ctx(0, 1, local, None)
    import torch
    import torch.distributed as td
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend == "nccl":
        torch.cuda.set_device(local)
        td.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    else:
        td.init_process_group(backend=backend)
    return Ctx(rank, world, local, backend)


def _device(ctx: Ctx):
    import torch
    return torch.device("cuda", ctx.local_rank) if ctx.backend == "nccl" else torch.device("cpu")


def allreduce_stats(ctx: Ctx, part: np.ndarray) -> np.ndarray:
    """Sum of the per-GPU partial statistics (the system's only collective)."""
    part = np.asarray(part, dtype=np.float64)
    if ctx.world == 1:
        return part.copy()
    import torch
    import torch.distributed as td
    t = torch.from_numpy(part.copy()).to(_device(ctx))
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return t.cpu().numpy()


def max_over_ranks(ctx: Ctx, value: float) -> float:
    if ctx.world == 1:
        return float(value)
    import torch
    import torch.distributed as td
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device(ctx))
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def barrier(ctx: Ctx):
    if ctx.world > 1:
        import torch.distributed as td
        if ctx.backend == "nccl":
            td.barrier(device_ids=[ctx.local_rank])
        else:
            td.barrier()


def finalize(ctx: Ctx):
    if ctx.world > 1:
        import torch.distributed as td
        td.destroy_process_group()


def summarise(tot: np.ndarray) -> dict:
    """Totals of srukf_stats (sum ex^2, ey^2, etheta^2, NEES, count, #NaN, #GMW-modified) -> RMSE / NEES."""
    cnt = max(float(tot[4]), 1.0)
    return dict(rmse_xy=float(np.sqrt((tot[0] + tot[1]) / cnt)), rmse_theta=float(np.sqrt(tot[2] / cnt)),
                nees=float(tot[3] / cnt), filters=int(tot[4]), flag_nan=int(tot[5]), flag_gmw_modified=int(tot[6]))
