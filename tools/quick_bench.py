"""scratch timing: V-cycle and smoother at a few sizes (not the contract bench)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cup3d_b200
from cup3d_b200 import mesh

for L in [int(a) for a in sys.argv[1:]] or [4, 5, 6]:
    t0 = time.time()
    ib, rb = mesh.uniform_blocks(L)
    ctx = cup3d_b200.Context(0, 8)
    ctx.mesh_upload(ib, rb, (1, 1, 1), L + 1)
    t1 = time.time()
    n = len(ib)
    b = torch.zeros(n * 512, dtype=torch.float64, device="cuda")
    b[0] = 1.0; b[-1] = -1.0
    z = torch.empty_like(b)
    for _ in range(3):
        ctx.mg_vcycle_dev(b, z)
    ctx.synchronize()
    l0 = ctx.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for _ in range(K):
        ctx.mg_vcycle_dev(b, z)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    cells = n * 512
    sm = ctx.time_smooth(L, 20)
    print("L=%d grid=%d^3 nblk=%d setup=%.2fs vcycle=%.3f ms (%.2f Gcell/s, %.0f GB/s @171B) launches/cycle=%d smooth=%.3f ms (%.0f GB/s @24B)" % (
        L, 8 << L, n, t1 - t0, ms, cells / ms / 1e6, cells * 171 / ms / 1e6, (ctx.kernel_launches() - l0 - 20) // K if False else (ctx.kernel_launches() - l0) // K, sm, cells * 24 / sm / 1e6), flush=True)
    ctx.close()
