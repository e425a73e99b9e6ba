"""torchrun target: tile-sharded windows over WORLD_SIZE GPUs + NCCL all-gather == the same windows on one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import fisr_b200
from fisr_b200 import sharding
from fisr_b200.init import xavier_params

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = fisr_b200.Engine(local)
eng.set_params(xavier_params(0, 0.01))
rng = np.random.default_rng(5)            # same data on every rank
B, H, W, grid = world, 200, 330, (2, 2)
frames = torch.from_numpy(rng.integers(0, 256, (B, H, W, 9), dtype=np.uint8)).cuda()
flow = torch.from_numpy((rng.standard_normal((B, H, W, 8)) * 3).astype(np.float32)).cuda()
warp = torch.from_numpy(rng.random((B, H, W, 12), dtype=np.float32)).cuda()
units = sharding.rank_units(rank, world, B, 4)
local_out = eng.units(frames, flow, warp, units, grid, layout="units")
gathered = sharding.gather_units(local_out, world)
got = sharding.assemble_frames(gathered, B, grid)
ref = eng.units(frames, flow, warp, list(range(B * 4)), grid, layout="frames")
ok = torch.equal(got, ref)
# the same all-gather in frame layout over peer memory (copy engines, sharding.PeerFrames)
oh, ow, _ = eng.canvas_shape(H, W, grid)
peer = sharding.PeerFrames(eng, rank, world, B, oh, ow)
for s in (0, 1, 0):                                          # both buffer sets, and a reused one
    eng.units(frames, flow, warp, units, grid, layout="frames", out=peer.local(s))
    peer.publish(s, units, grid)
    peer.drain()
    dist.barrier()
    torch.cuda.synchronize()
    ok = ok and torch.equal(peer.local(s), ref)
    peer.local(s).zero_()
    torch.cuda.synchronize()
    dist.barrier()
ncopies = peer.copies
peer.close()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"DIST_CHECK world={world} units/rank={len(units)} frames={tuple(got.shape)} nccl-gather and peer-memory exchange "
          f"({ncopies} 2-D copies from rank 0) bit-identical to one GPU: {bool(flag.item())}")
dist.barrier(); eng.close(); dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
