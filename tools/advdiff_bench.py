"""k_advdiff sweep and the advdiff() driver (3 Runge-Kutta stages) on a uniform grid, for A/B runs of the
kernel variants selected by environment (CUP_ADV_SLOTS=1|2, CUP_ADV_FUSE_RK=0|1, CUP_ADV_IMPL=ldg).
    python tools/advdiff_bench.py [level=6] [real_bytes=8]"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cup3d_b200
from cup3d_b200 import capi, mesh

L = int(sys.argv[1]) if len(sys.argv) > 1 else 6
RB = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ib, rb = mesh.uniform_blocks(L)
n = len(ib)
ctx = cup3d_b200.Context(0, RB)
ctx.mesh_upload(ib, rb, (1, 1, 1), L + 1)
ctx.set_params(dt=1e-3, nu=1e-3, uinf=(0.1, 0.0, 0.0), step=5, mean_constraint=2)
cells = n * 512
rt = ctypes.cdll.LoadLibrary("libcudart.so.12")
rng = torch.Generator(device="cuda").manual_seed(1)
dt = torch.float64 if RB == 8 else torch.float32
for f in range(2, 8):
    t = torch.empty(cells, dtype=dt, device="cuda")
    t.uniform_(-1, 1, generator=rng)
    torch.cuda.synchronize()
    rt.cudaMemcpy(ctypes.c_void_p(ctx.state_dev(f)), ctypes.c_void_p(t.data_ptr()), ctypes.c_size_t(cells * RB), 3)
torch.cuda.synchronize()


def timeit(fn, K):
    for _ in range(2):
        fn()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K):
        fn()
    ctx.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


sweep = timeit(lambda: ctx.stencil_apply(capi.ST_ADVDIFF), 8)
full = timeit(ctx.advdiff, 4)
print(json.dumps({"grid": 8 << L, "real_bytes": RB, "env": {k: v for k, v in os.environ.items() if k.startswith("CUP_")},
                  "sweep_ms": round(sweep, 4), "sweep_GBs_at_9_reals": round(cells * RB * 9 / sweep / 1e6, 1),
                  "advdiff_ms": round(full, 4),
                  "advdiff_GBs_at_36_reals_fused": round(cells * RB * 36 / full / 1e6, 1)}))
ctx.close()
