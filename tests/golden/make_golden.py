#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE.

The reference (slitvinov/CUP3D) ships no tests and no golden vectors
(SURVEY.md section 4), so parity is pinned on outputs of the reference itself: the
unmodified main.c compiled into oracle/_ref/libcup3d_ref.so (oracle/Makefile).
The reference keeps its mesh in file-statics, hence one subprocess per case.

    python tests/golden/make_golden.py            # all cases
    python tests/golden/make_golden.py CASE       # worker: one case

Inputs are closed-form or seeded (numpy default_rng(seed)), so the fixtures
are reproducible.  Output: tests/golden/<case>.npz
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# name -> reference command-line overrides
CASES = {
    "u16": dict(levelStart=1, levelMax=2),                      # 16^3, 8 blocks, 2 MG levels
    "u32": dict(levelStart=2, levelMax=3),                      # 32^3, 64 blocks
    "u64": dict(levelStart=3, levelMax=4),                      # 64^3, 512 blocks (survey's known-answer grid)
    "b211": dict(bpdx=2, bpdy=1, bpdz=1, levelStart=1, levelMax=2),   # 32x16x16, non-cubic base
    "b222_l0": dict(bpdx=2, bpdy=2, bpdz=2, levelStart=0, levelMax=1),  # single level: V-cycle == mg_bottom
    "b321": dict(bpdx=3, bpdy=2, bpdz=1, levelStart=1, levelMax=2),   # irregular Hilbert base (sfc.is_regular=0)
    # AMR: the reference's own mesh_adapt (main.c:4012) refines around a chi blob (k_gradchi tagging)
    "amr2": dict(bpdx=2, bpdy=2, bpdz=2, levelStart=1, levelMax=4, Rtol=5, Ctol=-1),   # 2 levels, 148 blocks
    "amr3": dict(bpdx=2, bpdy=2, bpdz=2, levelStart=1, levelMax=4, Rtol=5, Ctol=-1),   # 3 levels, 435 blocks
}
ADAPT = {"amr2": 1, "amr3": 2}  # mesh_adapt passes
# what each case stores (fixtures must stay small): "full" = inputs and all outputs;
# "mid" = outputs only (inputs are regenerated from the seed, their checksum is stored);
# "big" = V-cycle / solve outputs only.
TIER = {"u16": "full", "b211": "full", "b222_l0": "full", "b321": "mid", "u32": "big", "u64": "big",
        "amr2": "mid", "amr3": "big"}


def fields(ib, rb, seed):
    """closed-form + seeded test fields on the mesh"""
    from cup3d_b200 import mesh
    X, Y, Z = mesh.cell_centers(ib, rb)
    n = len(ib)
    h = rb[:, 0][:, None, None, None]
    rng = np.random.default_rng(seed)
    out = {}
    out["cosrhs"] = (h ** 3 * np.cos(np.pi * X) * np.cos(2 * np.pi * Y) * np.cos(3 * np.pi * Z)).reshape(n, 512)
    out["rand"] = rng.standard_normal((n, 512))
    # smooth solenoidal-ish velocity + perturbation, chi blob, udef
    vel = np.stack([np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y), -np.cos(2 * np.pi * X) * np.sin(2 * np.pi * Y),
                    0.1 * np.sin(2 * np.pi * Z)], 1).reshape(n, 3, 512)
    out["vel"] = vel + 0.01 * rng.standard_normal(vel.shape)
    r2 = (X - 0.5) ** 2 + (Y - 0.45) ** 2 + (Z - 0.55) ** 2
    out["chi"] = (1.0 / (1.0 + np.exp((np.sqrt(r2) - 0.2) / 0.03))).reshape(n, 512)
    out["udef"] = 0.05 * rng.standard_normal((n, 3, 512))
    out["pres"] = (np.cos(np.pi * X) * np.cos(np.pi * Y) * np.cos(2 * np.pi * Z)).reshape(n, 512) \
        + 0.01 * rng.standard_normal((n, 512))
    return out


def state0(F, n):
    """the full sta.fld [n][9][512] all stencil/driver cases start from"""
    st = np.zeros((n, 9, 512))
    st[:, 0] = F["chi"]
    st[:, 1] = F["pres"]
    st[:, 2:5] = F["vel"]
    st[:, 5:8] = F["udef"]
    st[:, 8] = F["rand"]
    return st


def solve_rhs(F, rb):
    rhs = F["cosrhs"] + 0.1 * rb[:, 0][:, None] ** 3 * F["rand"]
    return rhs - rhs.sum() / rhs.size


def adapt_like_golden(R, case):
    """the reference's own mesh_adapt passes that produce the multi-level meshes of the adapted cases"""
    for _ in range(ADAPT.get(case, 0)):
        # refine where 0 < chi < 0.9 (blob surface); velocity zero so vorticity does not tag
        ib, rb = R.blocks()
        from cup3d_b200 import mesh
        X, Y, Z = mesh.cell_centers(ib, rb)
        r = np.sqrt((X - 0.3) ** 2 + (Y - 0.35) ** 2 + (Z - 0.4) ** 2)
        st = np.zeros((R.nblk(), 9, 512))
        st[:, 0] = (1.0 / (1.0 + np.exp((r - 0.07) / 0.006))).reshape(-1, 512)
        R.state_set(st)
        R.mesh_adapt(5.0, -1.0)


def worker(case):
    from oracle import refbind as R
    R.init(**CASES[case])
    tier = TIER[case]
    adapt_like_golden(R, case)
    ib, rb = R.blocks()
    n = R.nblk()
    F = fields(ib, rb, seed=1234)
    g = {"ib": ib, "rb": rb}
    args = R.default_args(**CASES[case])
    g["bpd"] = np.array([args["bpdx"], args["bpdy"], args["bpdz"]], np.int32)
    g["level_max"] = np.int32(args["levelMax"])
    g["input_checksum"] = np.array([np.abs(F[k]).sum() for k in sorted(F)])
    if tier == "full":
        for k in F:
            g["in_" + k] = F[k]
    # --- mg_vcycle (main.c:4831) on two right-hand sides
    for name in (("cosrhs", "rand") if case not in ADAPT else ("cosrhs",)):
        g["vc_out_" + name] = R.mg_vcycle(F[name])
    # residual history of x += M(b - A x), constraint 2
    R.set_scalars(mean_constraint=2)
    b = F["cosrhs"]
    x = np.zeros_like(b)
    nb = np.sqrt(R.pois_dot(b, b))
    hist = []
    for _ in range(5):
        x = x + R.mg_vcycle(b - R.pois_op(x))
        r = b - R.pois_op(x)
        hist.append(np.sqrt(R.pois_dot(r, r)) / nb)
    g["vc_hist"] = np.array(hist)
    # --- pois_op (main.c:4282) for each mean-constraint mode
    for mc in ((0, 1, 2, 3) if tier != "big" and case not in ADAPT else ((0, 2) if tier != "big" else (2,))):
        R.set_scalars(mean_constraint=mc)
        g["op_out_mc%d" % mc] = R.pois_op(F["pres"])
    g["dot_ab"] = np.float64(R.pois_dot(F["pres"], F["rand"]))
    # --- stencil sweeps + drivers on a full state
    dt, nu, uinf = 1e-3, 1e-3, (0.1, -0.05, 0.02)
    g["scalars"] = np.array([dt, nu, *uinf])
    st = state0(F, n)
    R.set_scalars(dt=dt, nu=nu, uinf=uinf, step=5, mean_constraint=2)
    R.state_set(st)
    g["umax"] = np.float64(R.umax())  # sta_umax (main.c:5918) of the test state
    if tier != "big":
        keep = {"lhs": (R.F_LHS, 1), "advdiff": (R.F_TMP, 3), "prhs": (R.F_LHS, 1), "divp": (R.F_TMP, 1),
                "gradp": (R.F_TMP, 3)}
        keep.update({"vort": (R.F_TMP, 3), "q": (R.F_LHS, 1)})
        for name, (f0, nc) in keep.items():
            R.set_scalars(dt=dt, nu=nu, uinf=uinf, step=5, mean_constraint=2)
            R.state_set(st)
            R.stencil(name)
            g["st_" + name] = R.state_get()[:, f0:f0 + nc]
        # mesh_adapt's tagging input (main.c:4017-4019): vorticity(), then k_gradchi marks blocks at the
        # chi interface (1e10) and zeroes the vorticity inside bodies; mesh_tag_blk takes the block Linf
        R.set_scalars(dt=dt, nu=nu, uinf=uinf, step=5, mean_constraint=2)
        R.state_set(st)
        R.vorticity()
        w = R.state_get()[:, R.F_TMP:R.F_TMP + 3]
        g["tag_linf_all"] = np.abs(np.sqrt((w * w).sum(1))).max(1)
        g["tag_linf_fluid"] = np.where(st[:, R.F_CHI] > 0.9, 0.0, np.abs(np.sqrt((w * w).sum(1)))).max(1)
        R.stencil("gradchi")
        w = R.state_get()[:, R.F_TMP:R.F_TMP + 3]
        g["tag_marked"] = (w[:, 0, (3 * 8 + 3) * 8 + 3] == 1e10)
        g["tag_linf"] = np.abs(np.sqrt((w * w).sum(1))).max(1)
        # where k_gradchi did not mark the block, its zeroing is exactly the chi <= 0.9 mask
        assert np.array_equal(g["tag_linf"][~g["tag_marked"]], g["tag_linf_fluid"][~g["tag_marked"]])
        R.set_scalars(dt=dt, nu=nu, uinf=uinf, step=5, mean_constraint=2)
        R.state_set(st)
        R.advdiff()
        # VEL, TMP (multi-level cases keep only VEL: TMP is beta[2] = 0 times something)
        g["advdiff"] = R.state_get()[:, R.F_VEL:R.F_VEL + (6 if case not in ADAPT else 3)]
    # --- pois_solve (main.c:4875): rhs = F_LHS with zero mean, guess F_PRES = 0
    if case != "b222_l0":
        for mc in ((2, 1) if tier != "big" and case not in ADAPT else (2,)):
            if case == "amr3":
                break
            R.set_scalars(dt=dt, nu=nu, uinf=uinf, step=5, mean_constraint=mc, ptol=1e-10, ptol_rel=1e-12)
            s = st.copy()
            s[:, R.F_LHS] = solve_rhs(F, rb)
            s[:, R.F_PRES] = 0
            R.state_set(s)
            R.pois_solve()
            g["solve_x_mc%d" % mc] = R.state_get()[:, R.F_PRES]
        # --- projection (main.c:5828), both branches of the incremental-pressure switch
        if tier != "big":
            for step in ((1, 5) if case not in ADAPT else (5,)):
                R.set_scalars(dt=dt, nu=nu, uinf=uinf, step=step, mean_constraint=2, ptol=1e-10, ptol_rel=1e-12)
                R.state_set(st)
                R.projection()
                o = R.state_get()
                g["proj_step%d" % step] = o[:, R.F_PRES:R.F_PRES + 4]  # PRES, VEL
    np.savez_compressed(os.path.join(HERE, case + ".npz"), **g)
    print(case, "nblk", n, "hist", " ".join("%.6e" % v for v in hist))


def main():
    if len(sys.argv) > 1:
        worker(sys.argv[1])
        return
    for case in CASES:
        subprocess.run([sys.executable, os.path.abspath(__file__), case], check=True)


if __name__ == "__main__":
    main()
