"""Host-side sharding logic on 2 CPU processes (gloo): unit assignment, all-gather, frame assembly."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fisr_b200 import sharding


def test_unit_assignment_is_a_partition():
    for world, B, T in ((1, 1, 4), (2, 2, 4), (4, 4, 4), (8, 8, 4), (2, 2, 6)):
        seen = []
        for r in range(world):
            u = sharding.rank_units(r, world, B, T)
            assert len(u) == B * T // world
            seen += u
        assert sorted(seen) == list(range(B * T))
    # spatial-tile sharding: with 4 ranks every window's 4 tiles sit on 4 different ranks
    for w in range(4):
        owners = {r for r in range(4) for u in sharding.rank_units(r, 4, 4, 4) if u // 4 == w}
        assert owners == {0, 1, 2, 3}
    with pytest.raises(ValueError):
        sharding.rank_units(0, 3, 1, 4)


def test_unit_rect_tiles_the_frame():
    """PeerFrames pushes a unit's tile to the rectangle this returns: the rectangles of a window's units tile its frame."""
    grid, sh, sw = (2, 3), 4, 5
    cover = torch.zeros(2, grid[0] * sh, grid[1] * sw, dtype=torch.int32)
    for u in range(2 * 6):
        w, y0, x0 = sharding.unit_rect(u, grid, sh, sw)
        assert w == u // 6
        cover[w, y0:y0 + sh, x0:x0 + sw] += 1
    assert bool((cover == 1).all())
    assert sharding.unit_rect(5, (2, 2), 1024, 1920) == (1, 0, 1920)


def _tile_value(unit, sh, sw):
    t = torch.full((sh, sw, 9), unit % 251, dtype=torch.uint8)
    t[0, 0, 0] = unit // 4
    return t


def _worker(rank, world, port, B, grid, sh, sw, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T = grid[0] * grid[1]
    units = sharding.rank_units(rank, world, B, T)
    local = torch.stack([_tile_value(u, sh, sw) for u in units])          # stands in for Engine.units(layout="units")
    gathered = sharding.gather_units(local, world)
    frames = sharding.assemble_frames(gathered, B, grid)
    exp = torch.zeros(B, grid[0] * sh, grid[1] * sw, 9, dtype=torch.uint8)
    for w in range(B):
        for t in range(T):
            ty, tx = t // grid[1], t % grid[1]
            exp[w, ty * sh:(ty + 1) * sh, tx * sw:(tx + 1) * sw] = _tile_value(w * T + t, sh, sw)
    ret[rank] = bool(torch.equal(frames, exp))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B,grid", [(2, (2, 2)), (4, (2, 3))])
def test_gather_and_assemble_world2(B, grid):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, B, grid, 6, 10, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
