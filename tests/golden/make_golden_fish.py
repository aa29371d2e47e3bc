#!/usr/bin/env python3
"""Golden fixtures of the obstacle (fish) phases of one time step, made by RUNNING THE REFERENCE
phase by phase (oracle/_ref, ref_harness.c:ref_phase): fish_build -> advdiff -> fish_mom_blk /
fish_vel -> fish_pen -> fish_tmpv -> projection (advance(), main.c:5984-6003).

    python tests/golden/make_golden_fish.py          # all cases
    python tests/golden/make_golden_fish.py CASE     # worker (one reference process per case)

The bodies (ObstacleBlock chi / udef, centre of mass, rigid motion) are the reference's own:
fish_build and fish_vel (fish_solve) stay host code.  Velocity and pressure are the seeded closed
forms of make_golden.fields(); the tests regenerate them.  Stored outputs: full values on the
obstacle blocks plus three checksums per block and field for all blocks (fixtures stay small).
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

_COMMON = ("T=1.0 phi=0 amplitudeFactor=1 ypos=0.5 zpos=0.5 heightProfile=danio widthProfile=stefan "
           "bForcedInSimFrame_x=0 bForcedInSimFrame_y=0 bForcedInSimFrame_z=0 xvel=0 yvel=0 zvel=0 "
           "bFixToPlanar=0 CorrectPosition=0 CorrectPositionZ=0 CorrectRoll=0 wyp=1 wzp=1")


def _fish(L, xpos, angle, fix):
    return ("L=%g xpos=%g planarAngle=%g bFixFrameOfRef_x=%d bFixFrameOfRef_y=%d bFixFrameOfRef_z=%d "
            % (L, xpos, angle, fix, fix, fix)) + _COMMON


CASES = {
    # uniform 64^3, two fish facing each other as in the reference's run.sh (first one fixes the frame)
    "fish64": dict(levelStart=3, levelMax=4, Ctol=-1,
                   **{"factory-content": _fish(0.4, 0.35, 180, 1) + "\n" + _fish(0.4, 0.62, 0, 0)}),
    # the reference's own initial adaptation around one fish (sta_fields, main.c:4201): 2 levels
    "fishamr": dict(levelStart=3, levelMax=5, Ctol=-1, **{"factory-content": _fish(0.5, 0.45, 0, 0)}),
}
WARM = 2   # full reference steps before the captured one (the midline needs time > 0 to deform)
NU = 1e-3


def weights():
    return np.cos(0.37 * np.arange(512) + 0.11)


def checksums(a):
    """[n, nc, 512] -> [n, nc, 3]: sum, sum of |.|, dot with a fixed pattern"""
    return np.stack([a.sum(-1), np.abs(a).sum(-1), (a * weights()).sum(-1)], -1)


def worker(case):
    from oracle import refbind as R
    from oracle import fish_port as P
    import make_golden as MG
    R.init(**CASES[case])
    R.sta_fields()
    nf = R.nfish()
    for _ in range(WARM):
        R.sta_dt()
        R.phase("mesh_adapt")
        for p in ("fish_build", "advdiff", "fish_vel", "fish_pen", "projection"):
            R.phase(p)
        R.step_end()
    # ---- the captured step
    dt = R.sta_dt()
    R.phase("mesh_adapt")
    R.phase("fish_build")
    sc = R.get_scalars()
    ib, rb = R.blocks()
    n = R.nblk()
    F = MG.fields(ib, rb, seed=4321)
    st = R.state_get()
    chi_field = st[:, R.F_CHI].copy()
    st[:, R.F_PRES] = F["pres"]
    st[:, R.F_VEL:R.F_VEL + 3] = F["vel"]
    st[:, R.F_TMP:] = 0
    R.state_set(st)
    step = 5
    R.set_scalars(dt=dt, nu=NU, uinf=sc["uinf"], step=step, mean_constraint=2, ptol=1e-10, ptol_rel=1e-12)
    args = R.default_args(**CASES[case])
    g = {"ib": ib, "rb": rb, "bpd": np.array([args["bpdx"], args["bpdy"], args["bpdz"]], np.int32),
         "level_max": np.int32(args["levelMax"]), "nfish": np.int32(nf),
         "scalars": np.array([dt, NU, *sc["uinf"], step, sc["lam"]]),
         "input_checksum": np.array([np.abs(F[k]).sum() for k in ("pres", "vel")])}
    obs = [R.fish_obstacle(k) for k in range(nf)]
    obu = np.unique(np.concatenate([o[0] for o in obs]))
    assert not np.any(np.delete(chi_field, obu, axis=0)), "F_CHI is non-zero outside the obstacle blocks"
    g["obu"] = obu.astype(np.int32)
    g["chi_field"] = chi_field[obu]
    for k, (blk, chi, udef) in enumerate(obs):
        g["ob%d_blk" % k], g["ob%d_chi" % k], g["ob%d_udef" % k] = blk, chi, udef
    # advdiff
    R.phase("advdiff")
    s1 = R.state_get()
    g["adv_vel"] = s1[obu, R.F_VEL:R.F_VEL + 3]
    g["adv_sums"] = checksums(s1[:, R.F_VEL:R.F_VEL + 3])
    # moments of every body (fish_mom_blk + block sum), then the reference's rigid-body solve
    mot0 = [R.fish_motion(k) for k in range(nf)]
    for k in range(nf):
        g["mom%d" % k] = R.fish_mom(k)
        g["com%d" % k] = mot0[k][0]
        Mp = P.moments(ib, rb, s1[:, R.F_VEL:R.F_VEL + 3], *obs[k], mot0[k][0], dt, sc["lam"])
        assert np.allclose(Mp, g["mom%d" % k], rtol=1e-11, atol=1e-14 * np.abs(Mp).max()), "numpy port: moments"
    R.phase("fish_vel")
    mot = [R.fish_motion(k) for k in range(nf)]
    for k in range(nf):
        assert np.array_equal(mot[k][0], mot0[k][0])
        g["vel%d" % k], g["omega%d" % k] = mot[k][1], mot[k][2]
    # penalisation (fish_pen = fish_hit + the block loop; the numpy port proves fish_hit was a no-op)
    R.phase("fish_pen")
    s2 = R.state_get()
    vp = s1[:, R.F_VEL:R.F_VEL + 3].copy()
    for k in range(nf):
        P.penalize(ib, rb, vp, chi_field, *obs[k], *mot[k], dt, sc["lam"])
    assert np.allclose(vp, s2[:, R.F_VEL:R.F_VEL + 3], rtol=0, atol=1e-14), "fish_hit changed the velocity"
    assert np.array_equal(np.delete(s2, obu, axis=0), np.delete(s1, obu, axis=0))
    g["pen_vel"] = s2[obu, R.F_VEL:R.F_VEL + 3]
    # fish_tmpv on a zeroed F_TMP (what projection() does first)
    s3 = s2.copy()
    s3[:, R.F_TMP:R.F_TMP + 3] = 0
    R.state_set(s3)
    R.phase("fish_tmpv")
    s4 = R.state_get()
    g["tmpv"] = s4[obu, R.F_TMP:R.F_TMP + 3]
    tp = np.zeros((n, 3, 512))
    for k in range(nf):
        P.tmpv(tp, chi_field, *obs[k])
    assert np.array_equal(tp, s4[:, R.F_TMP:R.F_TMP + 3]), "numpy port: fish_tmpv"
    # projection (zeroes F_TMP, fish_tmpv, prhs, divp, solve, gradp update)
    R.state_set(s2)
    R.phase("projection")
    s5 = R.state_get()
    g["proj"] = s5[obu, R.F_PRES:R.F_PRES + 4]
    g["proj_sums"] = checksums(s5[:, R.F_PRES:R.F_PRES + 4])
    # mesh_adapt's tagging input on the projected state (main.c:4017-4019)
    R.vorticity()
    w = R.state_get()[:, R.F_TMP:R.F_TMP + 3]
    mag = np.abs(np.sqrt((w * w).sum(1)))
    g["tag_linf_all"] = mag.max(1)
    g["tag_linf_fluid"] = np.where(chi_field > 0.9, 0.0, mag).max(1)
    R.stencil("gradchi")
    w = R.state_get()[:, R.F_TMP:R.F_TMP + 3]
    g["tag_marked"] = (w[:, 0, (3 * 8 + 3) * 8 + 3] == 1e10)
    g["tag_linf"] = np.abs(np.sqrt((w * w).sum(1))).max(1)
    assert np.array_equal(g["tag_linf"][~g["tag_marked"]], g["tag_linf_fluid"][~g["tag_marked"]])
    np.savez_compressed(os.path.join(HERE, case + ".npz"), **g)
    print(case, "marked", int(g["tag_marked"].sum()), "nblk", n, "levels", np.bincount(ib[:, 0]), "nfish", nf, "nob", [len(o[0]) for o in obs], "dt", dt,
          "uinf", sc["uinf"], "|vel|", [float(np.abs(m[1]).max()) for m in mot],
          "|omega|", [float(np.abs(m[2]).max()) for m in mot])


def main():
    if len(sys.argv) > 1:
        worker(sys.argv[1])
        return
    for case in CASES:
        subprocess.run([sys.executable, os.path.abspath(__file__), case], check=True)


if __name__ == "__main__":
    main()
