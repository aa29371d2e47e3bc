#!/usr/bin/env python3
"""oracle/live_ref.py -- TEST/BENCH INFRASTRUCTURE ONLY.

Runs the reference itself (oracle/_ref, the unmodified main.c built by oracle/Makefile; Real =
double, or float with --real 4) LIVE on inputs handed over as .npy files, and writes its outputs
next to them.  Used where no stored golden can exist: the full-size grids of BASELINE.json
(256^3 / 512^3) in tests/test_gpu_live_oracle.py and bench.py's parity check.  One process per
mesh (the reference keeps its mesh in file-statics).

    live_ref.py --level L --dir D --ops vcycle,op,advdiff,proj [--real 8|4] [--threads T]

inputs  D/in_vec.npy   [n,512]    right-hand side / operand of vcycle and op
        D/in_state.npy [n,9,512]  sta.fld for advdiff / proj
outputs D/out_vcycle.npy, D/out_op.npy, D/out_advdiff.npy [n,3,512] (F_VEL),
        D/out_proj.npy [n,4,512] (F_PRES, F_VEL), D/ib.npy, D/meta.json
        gradchi: D/out_gradchi.npy [n,3,512] = F_TMP after stencil_apply(&st_gradchi) (main.c:3649)
        adapt:   vorticity(); gradchi -> D/adapt_mid.npy (the state mesh_adapt's refinement / compression
                 read), then mesh_adapt(--rtol, --ctol) -> D/adapt_ib.npy, D/adapt_rb.npy, D/adapt_state.npy
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=-1, help="uniform grid (8 << level)^3, bpd 1")
    ap.add_argument("--case", default="", help="a mesh of tests/golden/make_golden.py (CASES; adapted ones included)")
    ap.add_argument("--dir", required=True)
    ap.add_argument("--ops", default="vcycle")
    ap.add_argument("--real", type=int, default=8)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--dt", type=float, default=1e-3)
    ap.add_argument("--nu", type=float, default=1e-3)
    ap.add_argument("--uinf", default="0.1,-0.05,0.02")
    ap.add_argument("--ptol", type=float, default=1e-9)
    ap.add_argument("--ptol-rel", type=float, default=1e-14)
    ap.add_argument("--rtol", type=float, default=1e9)
    ap.add_argument("--ctol", type=float, default=-1.0)
    a = ap.parse_args()
    if a.threads > 0:
        os.environ["OMP_NUM_THREADS"] = str(a.threads)
    from oracle import refbind as R
    t0 = time.time()
    if a.case:
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import make_golden as MG
        R.init(real_bytes=a.real, **MG.CASES[a.case])
        MG.adapt_like_golden(R, a.case)  # the reference's own mesh_adapt passes of the adapted cases
    else:
        R.init(real_bytes=a.real, levelStart=a.level, levelMax=a.level + 1)
    meta = {"mesh_init_s": time.time() - t0, "nblk": R.nblk(), "threads": R.threads(), "real_bytes": a.real}
    ib, rb = R.blocks()
    np.save(os.path.join(a.dir, "ib.npy"), ib)
    np.save(os.path.join(a.dir, "rb.npy"), rb)
    ops = [o for o in a.ops.split(",") if o]
    uinf = tuple(float(v) for v in a.uinf.split(","))
    R.set_scalars(dt=a.dt, nu=a.nu, uinf=uinf, step=5, mean_constraint=2, ptol=a.ptol, ptol_rel=a.ptol_rel)
    if "vcycle" in ops or "op" in ops:
        x = np.load(os.path.join(a.dir, "in_vec.npy"))
        if "vcycle" in ops:
            t0 = time.time()
            np.save(os.path.join(a.dir, "out_vcycle.npy"), R.mg_vcycle(x))
            meta["vcycle_s"] = time.time() - t0
        if "op" in ops:
            np.save(os.path.join(a.dir, "out_op.npy"), R.pois_op(x))
        del x
    if any(o in ops for o in ("advdiff", "proj", "gradchi", "adapt")):
        st = np.load(os.path.join(a.dir, "in_state.npy"))
        if "advdiff" in ops:
            R.state_set(st)
            t0 = time.time()
            R.advdiff()
            meta["advdiff_s"] = time.time() - t0
            np.save(os.path.join(a.dir, "out_advdiff.npy"), R.state_get()[:, 2:5])
        if "proj" in ops:
            R.state_set(st)
            t0 = time.time()
            R.projection()
            meta["proj_s"] = time.time() - t0
            np.save(os.path.join(a.dir, "out_proj.npy"), R.state_get()[:, 1:5])
        if "gradchi" in ops:
            R.state_set(st)
            R.stencil("gradchi")
            np.save(os.path.join(a.dir, "out_gradchi.npy"), R.state_get()[:, 5:8])
        if "adapt" in ops:  # last: it changes the mesh
            R.state_set(st)
            R.vorticity()
            R.stencil("gradchi")
            np.save(os.path.join(a.dir, "adapt_mid.npy"), R.state_get())
            R.mesh_adapt(a.rtol, a.ctol)
            ib2, rb2 = R.blocks()
            np.save(os.path.join(a.dir, "adapt_ib.npy"), ib2)
            np.save(os.path.join(a.dir, "adapt_rb.npy"), rb2)
            np.save(os.path.join(a.dir, "adapt_state.npy"), R.state_get())
    with open(os.path.join(a.dir, "meta.json"), "w") as f:
        json.dump(meta, f)


if __name__ == "__main__":
    main()
