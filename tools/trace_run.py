"""scratch: one traced V-cycle per rank (CUP_TRACE=1), 512^3 split over WORLD_SIZE ranks"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cup3d_b200
from cup3d_b200 import capi, mesh
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
L = int(sys.argv[1]) if len(sys.argv) > 1 else 6
gib, grb = mesh.uniform_blocks(L)
owner = capi.split_owner(len(gib), world)
mine = np.nonzero(owner == rank)[0]
ctx = cup3d_b200.Context(lr, 8)
if world > 1:
    box = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(rank, world, box[0])
ctx.mesh_upload(gib[mine], grb[mine], (1, 1, 1), L + 1)
b = torch.zeros(len(mine) * 512, dtype=torch.float64, device="cuda"); b[0] = 1
z = torch.empty_like(b)
os.environ.pop("CUP_TRACE", None)
for _ in range(3):
    ctx.mg_vcycle_dev(b, z)
torch.cuda.synchronize()
if dist: dist.barrier()
os.environ["CUP_TRACE"] = "1" if rank == 0 else "0"
ctx.mg_vcycle_dev(b, z)
torch.cuda.synchronize()
ctx.close()
