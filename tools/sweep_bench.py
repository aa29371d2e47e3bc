"""Per-sweep rates of the stencil kernels and drivers on a uniform grid (secondary numbers next
to bench.py's V-cycle; the CPU column of BASELINE.md lists the reference's rates for the same
sweeps).  python tools/sweep_bench.py [level]   (5 = 256^3, 6 = 512^3)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cup3d_b200
from cup3d_b200 import capi, mesh

L = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ib, rb = mesh.uniform_blocks(L)
n = len(ib)
ctx = cup3d_b200.Context(0, 8)
ctx.mesh_upload(ib, rb, (1, 1, 1), L + 1)
ctx.set_params(dt=1e-3, nu=1e-3, uinf=(0.1, 0.0, 0.0), step=5, mean_constraint=2, ptol=1e-6, ptol_rel=1e-4)
cells = n * 512
# fill the device state with something smooth (no host round trip of 9 x 1 GB at 512^3)
rng = torch.Generator(device="cuda").manual_seed(1)
for f in range(9):
    t = torch.empty(cells, dtype=torch.float64, device="cuda")
    t.uniform_(-1, 1, generator=rng)
    import ctypes
    ctypes.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(ctypes.c_void_p(ctx.state_dev(f)), ctypes.c_void_p(t.data_ptr()),
                                                        ctypes.c_size_t(cells * 8), 3)
torch.cuda.synchronize()
REALS = {"advdiff": 9, "prhs": 8, "divp": 2, "gradp": 4, "lhs": 2}
IDS = {"advdiff": capi.ST_ADVDIFF, "prhs": capi.ST_PRHS, "divp": capi.ST_DIVP, "gradp": capi.ST_GRADP, "lhs": capi.ST_LHS}
out = {"grid": "%d^3" % (8 << L), "cells": cells}
for name, sid in IDS.items():
    for _ in range(3):
        ctx.stencil_apply(sid)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K):
        ctx.stencil_apply(sid)
    ctx.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    out[name] = {"ms": round(ms, 4), "Gcells_per_s": round(cells / ms / 1e6, 2),
                 "GBs_at_algorithmic": round(cells * 8 * REALS[name] / ms / 1e6, 1)}
print(json.dumps(out))
ctx.close()
