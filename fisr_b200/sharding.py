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


# ------------------------------------------------------------------------------------------ frames over peer memory
def unit_rect(unit: int, grid: Tuple[int, int], sh: int, sw: int) -> Tuple[int, int, int]:
    """(window, y0, x0) of a unit's trimmed tile inside its window's frame (paste of FISRnet.py:1056-1057)."""
    T = grid[0] * grid[1]
    w, t = divmod(unit, T)
    return w, (t // grid[1]) * sh, (t % grid[1]) * sw


class _DevPtr:
    """Exposes a raw device allocation to torch through ``__cuda_array_interface__`` (no copy)."""

    def __init__(self, ptr: int, shape, typestr: str = "|u1"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2,
                                         "strides": None}


class PeerFrames:
    """All-gather of the output frames in FRAME layout over NVLink peer memory, driven by the copy engines.

    Every rank owns ``sets`` frame buffers ``[B, OH, OW, 9]`` uint8 allocated by the library and shared through CUDA IPC.  A
    step computes the rank's units straight into its own buffer (``Engine.units(layout="frames", out=self.local(s))``) and
    ``publish`` pushes each finished tile rectangle into the same place of every peer's buffer with ``cudaMemcpy2DAsync`` on
    peer-mapped pointers (copy engines, side streams): no SM-resident collective kernel competes with the persistent conv
    grids and no re-assembly pass runs.  ``drain`` + a barrier make every rank's frames complete."""

    def __init__(self, engine, rank: int, world: int, n_windows: int, oh: int, ow: int, sets: int = 2, group=None, streams: int = 2):
        import ctypes as C
        self.eng, self.rank, self.world, self.group = engine, rank, world, group
        self.B, self.oh, self.ow, self.sets = n_windows, oh, ow, sets
        self.bytes = n_windows * oh * ow * 9
        lib = engine.lib
        self._own, handles = [], []
        for _ in range(sets):
            ptr, h = C.c_void_p(), C.create_string_buffer(64)
            engine._check(lib.fisr_ipc_alloc(engine.h, self.bytes, C.byref(ptr), h), "fisr_ipc_alloc")
            self._own.append(ptr.value)
            handles.append(h.raw)
        all_handles = [None] * world
        dist.all_gather_object(all_handles, handles, group=group)
        self._peer = {}                                     # (rank, set) -> mapped pointer
        for r in range(world):
            for s in range(sets):
                if r == rank:
                    self._peer[(r, s)] = self._own[s]
                    continue
                ptr = C.c_void_p()
                engine._check(lib.fisr_ipc_open(engine.h, all_handles[r][s], C.byref(ptr)), "fisr_ipc_open")
                self._peer[(r, s)] = ptr.value
        dev = torch.device("cuda", engine.device)
        self._local = [torch.as_tensor(_DevPtr(p, (n_windows, oh, ow, 9)), device=dev) for p in self._own]
        self._streams = [torch.cuda.Stream(engine.device) for _ in range(streams)]
        self._events = [torch.cuda.Event() for _ in range(sets)]
        self.copies = 0

    def local(self, s: int) -> torch.Tensor:
        return self._local[s]

    def publish(self, s: int, units, grid: Tuple[int, int]) -> None:
        """Asynchronously copies this rank's tiles of buffer set ``s`` into every peer's set ``s`` (stream-ordered behind the
        work already enqueued on the current stream)."""
        sh, sw = self.oh // grid[0], self.ow // grid[1]
        pitch = self.ow * 9
        ev = self._events[s]
        ev.record(torch.cuda.current_stream(self.eng.device))
        for st in self._streams:
            st.wait_event(ev)
        k = 0
        for shift in range(1, self.world):                  # start with a different peer on every rank: no incast hot spot
            r = (self.rank + shift) % self.world
            for u in units:
                w, y0, x0 = unit_rect(u, grid, sh, sw)
                off = (w * self.oh + y0) * pitch + x0 * 9
                st = self._streams[k % len(self._streams)]
                k += 1
                self.eng._check(self.eng.lib.fisr_copy2d_async(self.eng.h, self._peer[(r, s)] + off, pitch, self._own[s] + off, pitch,
                                                               sw * 9, sh, st.cuda_stream), "fisr_copy2d_async")
        self.copies += k

    def drain(self) -> None:
        for st in self._streams:
            st.synchronize()

    def close(self) -> None:
        self.drain()
        if self.world > 1:
            dist.barrier(group=self.group)
        for (r, s), p in self._peer.items():
            if r != self.rank:
                self.eng.lib.fisr_ipc_close(self.eng.h, p)
        self._local = []
        for p in self._own:
            self.eng.lib.fisr_ipc_free(self.eng.h, p)
        self._own = []
