"""Multi-GPU sharding of the tiled inference path (one process per GPU, ``torch.distributed``).

The reference runs the (window, tile) loops of ``FISRnet.py:994,1028`` serially on one GPU; the iterations carry
no dependence, so here they are the unit of sharding.  A *step* covers ``B`` windows x ``T`` tiles.  Units are
ordered tile-major, ``[(t0,w0) .. (t0,wB-1), (t1,w0) ..]``, and rank ``r`` takes the r-th contiguous block: every
window's tiles are spread over the ranks (spatial-tile sharding) while each rank still runs equal-sized tiles as one
batched forward.  One all-gather of the trimmed uint8 tiles (unit-major send buffer) rebuilds the frames everywhere.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def step_units(n_windows: int, tiles_per_window: int) -> List[int]:
    """Unit ids (window * T + tile) of one step in tile-major order."""
    return [w * tiles_per_window + t for t in range(tiles_per_window) for w in range(n_windows)]


def rank_units(rank: int, world: int, n_windows: int, tiles_per_window: int) -> List[int]:
    units = step_units(n_windows, tiles_per_window)
    if len(units) % world:
        raise ValueError(f"{len(units)} units do not divide over {world} ranks")
    k = len(units) // world
    return units[rank * k:(rank + 1) * k]


def gather_units(local: torch.Tensor, world: int, out: torch.Tensor = None, group=None, async_op: bool = False):
    """All-gather of the per-rank unit-major tile buffers [k, sh, sw, 9] -> [world * k, sh, sw, 9] (rank order).
    ``async_op=True`` returns ``(out, work)``: the collective runs on NCCL's stream while the caller's stream goes on (the
    next step's kernels); ``work.wait()`` orders the caller's stream behind it before ``out`` is read or ``local`` reused."""
    if world == 1:
        return (local, None) if async_op else local
    if out is None:
        out = local.new_empty((world * local.shape[0],) + tuple(local.shape[1:]))
    if dist.get_backend(group) == "nccl":
        work = dist.all_gather_into_tensor(out, local, group=group, async_op=async_op)      # NCCL over NVLink 5 / NVSwitch
    else:                                                           # gloo (CPU tests)
        parts = list(out.chunk(world, dim=0))
        work = dist.all_gather(parts, local.contiguous(), group=group, async_op=async_op)
    return (out, work) if async_op else out


def assemble_frames(gathered: torch.Tensor, n_windows: int, grid: Tuple[int, int]) -> torch.Tensor:
    """Unit-major tiles in tile-major step order -> frames [B, pH*sh, pW*sw, 9] (paste of FISRnet.py:1056-1057)."""
    pH, pW = grid
    n, sh, sw, c = gathered.shape
    if n != n_windows * pH * pW:
        raise ValueError(f"{n} tiles for {n_windows} windows of {pH}x{pW}")
    g = gathered.view(pH, pW, n_windows, sh, sw, c).permute(2, 0, 3, 1, 4, 5)
    return g.reshape(n_windows, pH * sh, pW * sw, c)
