"""One full time step of the hot path -- advdiff() (3 k_advdiff sweeps + RK3 updates) followed by
projection() (k_prhs, k_divp, GMRES/V-cycle solve, k_gradp, updates) -- on a uniform grid and on a
synthetic 2:1-balanced AMR mesh, fp64 and fp32.  Secondary numbers next to bench.py's V-cycle
(configs 2-3 of BASELINE.json are step timings).   python tools/step_bench.py [uniform_level]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cup3d_b200
from cup3d_b200 import mesh

UL = int(sys.argv[1]) if len(sys.argv) > 1 else 5


def run(tag, ib, rb, bpd, level_max, rbytes):
    n = len(ib)
    ctx = cup3d_b200.Context(0, rbytes)
    t0 = time.time()
    ctx.mesh_upload(ib, rb, bpd, level_max)
    setup = time.time() - t0
    X, Y, Z = mesh.cell_centers(ib, rb)
    st = np.zeros((n, 9, 512))
    st[:, 2] = (np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y)).reshape(n, 512)
    st[:, 3] = (-np.cos(2 * np.pi * X) * np.sin(2 * np.pi * Y)).reshape(n, 512)
    st[:, 4] = (0.1 * np.sin(2 * np.pi * Z)).reshape(n, 512)
    ctx.state_h2d(st)
    hmin = float(rb[:, 0].min())
    ctx.set_params(dt=0.2 * hmin, nu=1e-3, uinf=(0.0, 0.0, 0.0), step=5, mean_constraint=2,
                   ptol=1e-6 if rbytes == 8 else 1e-4, ptol_rel=1e-4 if rbytes == 8 else 1e-3)
    res = {}
    for phase in ("advdiff", "projection"):
        fn = getattr(ctx, phase)
        fn()  # warm-up (graph capture, Krylov allocation)
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 3
        torch.cuda.synchronize()
        e0.record()
        for _ in range(K):
            info = fn()
        ctx.synchronize()
        e1.record()
        torch.cuda.synchronize()
        res[phase + "_ms"] = round(e0.elapsed_time(e1) / K, 3)
        if phase == "projection":
            res["krylov_iterations"] = info.iterations
            res["vcycles"] = info.vcycles
    cells = n * 512
    res.update(tag=tag, blocks=n, cells=cells, levels=np.bincount(ib[:, 0]).tolist(), setup_s=round(setup, 2),
               step_Mcells_per_s=round(cells / (res["advdiff_ms"] + res["projection_ms"]) / 1e3, 1),
               umax=ctx.umax())
    print(json.dumps(res), flush=True)
    ctx.close()


ib, rb = mesh.uniform_blocks(UL)
for rbytes in (8, 4):
    run("uniform %d^3 fp%d" % (8 << UL, 8 * rbytes), ib, rb, (1, 1, 1), UL + 1, rbytes)
# AMR: 64^3-equivalent base (bpd 2, level 2 = 8 blocks/dim), refined twice around a sphere
ib, rb = mesh.amr_blocks(2, 4, mesh.sphere_shell((0.45, 0.5, 0.55), 0.2, band=0.5), bpd=(2, 2, 2))
for rbytes in (8, 4):
    run("amr sphere L2-4 fp%d" % (8 * rbytes), ib, rb, (2, 2, 2), 5, rbytes)
