#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: the reference's time loop (advance(), main.c:5984-6003) with its data-parallel
hot path served by a pluggable backend, orchestrated exactly as INTEGRATION.md prescribes:

    host (reference)                         backend ("device")
    ---------------------------------------  ----------------------------------------------
    sta_dt, mesh_adapt, fish_build           <- fields come back before them (adaptation needs them)
    -> blocks, F_CHI, obstacle blocks, com   mesh_upload / state_h2d / obstacle_upload / motion
                                             advdiff
    fish_solve  <- moments M[29]             obstacle_moments        (fish_vel, split in two)
    fish_hit; -> vel, omega                  obstacle_motion
                                             obstacle_penalize       (fish_pen's block loop)
                                             projection              (fish_tmpv inside)
    step++, time += dt

Backends: "ref" = the same reference functions called piecewise on the reference's own host state (no
GPU: proves the split of fish_vel / fish_pen and the order of calls reproduce advance() bit for bit);
"port" = an INDEPENDENT implementation with its own memory -- the C restatement (oracle/cup_oracle.c) for
advdiff / projection and the numpy restatement (oracle/fish_port.py) for the obstacle phases -- which
only sees what the orchestration hands over, exactly like a device does (uniform meshes; no GPU: proves
the hand-over is complete); "gpu" = cup3d_b200.Context.  `pure` runs advance()'s phases in the
reference's own order.

    python tests/coupled.py MODE CASE NSTEPS OUT.npz       MODE = pure | ref | port | gpu
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, HERE, os.path.join(HERE, "golden")]

NU = 1e-3


def checks(st):
    """sum and sum of |.| of PRES, VEL over the whole mesh"""
    a = st[:, 1:5]
    return np.array([a.sum(), np.abs(a).sum()])


def record(R, rec, st, umax):
    sc = R.get_scalars()
    rec["dt"].append(sc["dt"])
    rec["nblk"].append(R.nblk())
    rec["umax"].append(umax)
    rec["checks"].append(checks(st))
    rec["motion"].append(np.concatenate([np.concatenate(R.fish_motion(k)) for k in range(R.nfish())]))


def adapt_due(R):
    step = R.get_scalars()["step"]
    return step % 20 == 0 or step < 10      # advance(), main.c:5991


def run_pure(R, nsteps):
    rec = {k: [] for k in ("dt", "nblk", "umax", "checks", "motion")}
    for _ in range(nsteps):
        R.sta_dt()
        if adapt_due(R):
            R.phase("mesh_adapt")
        for p in ("fish_build", "advdiff", "fish_vel", "fish_pen", "projection"):
            R.phase(p)
        R.step_end()
        record(R, rec, R.state_get(), R.umax())
    return rec


class RefBackend:
    """the device API served by the reference's own functions on ITS state (sta.fld)"""

    def __init__(self, R):
        self.R = R

    def mesh(self, ib, rb, bpd, level_max):
        pass

    def h2d(self, st):
        self.R.state_set(st)

    def d2h(self):
        return self.R.state_get()

    def params(self, **kw):
        pass                      # the reference reads its own sta / sim

    def obstacle(self, k, blk, chi, udef, com):
        pass                      # fish_build left them in the reference's Fish structs

    def motion(self, k, vel, omega):
        pass

    def advdiff(self):
        self.R.phase("advdiff")

    def moments(self, k):
        return self.R.fish_mom(k)

    def penalize(self):
        self.R.fish_pen_blocks()

    def projection(self):
        self.R.phase("projection")

    def umax(self):
        return self.R.umax()


class PortBackend:
    """own state + the oracle's C and numpy restatements: everything it knows came through this interface"""

    def __init__(self, ptol, ptol_rel):
        from oracle import fish_port, portbind
        self.FP, self.PB = fish_port, portbind
        self.orc = None
        self.tol = (ptol, ptol_rel)
        self.ob = {}

    def mesh(self, ib, rb, bpd, level_max):
        if self.orc is not None:
            self.orc.close()
        self.ib, self.rb = ib.copy(), rb.copy()
        self.orc = self.PB.Oracle(ib, rb, bpd, level_max)
        self.ob = {}

    def h2d(self, st):
        self.st = np.array(st, dtype=np.float64, order="C", copy=True)

    def d2h(self):
        return self.st.copy()

    def params(self, dt, uinf, step, lam):
        self.dt, self.uinf, self.step, self.lam = dt, tuple(uinf), step, lam

    def obstacle(self, k, blk, chi, udef, com):
        self.ob[k] = dict(blk=blk.copy(), chi=chi.copy(), udef=udef.copy(), com=np.array(com), vel=np.zeros(3),
                          omega=np.zeros(3))

    def motion(self, k, vel, omega):
        self.ob[k]["vel"], self.ob[k]["omega"] = np.array(vel), np.array(omega)

    def advdiff(self):
        self.orc.advdiff(self.st, self.dt, NU, self.uinf)

    def moments(self, k):
        o = self.ob[k]
        return self.FP.moments(self.ib, self.rb, self.st[:, 2:5], o["blk"], o["chi"], o["udef"], o["com"], self.dt,
                               self.lam)

    def penalize(self):
        vel = self.st[:, 2:5]            # a view: penalize() writes through
        for k in sorted(self.ob):
            o = self.ob[k]
            self.FP.penalize(self.ib, self.rb, vel, self.st[:, 0], o["blk"], o["chi"], o["udef"], o["com"], o["vel"],
                             o["omega"], self.dt, self.lam)

    def projection(self):
        self.st[:, 5:8] = 0
        tmp = self.st[:, 5:8]
        for k in sorted(self.ob):
            o = self.ob[k]
            self.FP.tmpv(tmp, self.st[:, 0], o["blk"], o["chi"], o["udef"])
        self.orc.projection(self.st, self.dt, NU, self.uinf, self.step, 2, self.tol[0], self.tol[1], keep_tmp=True)

    def umax(self):
        u = np.asarray(self.uinf)[None, :, None]
        return float(np.max(np.abs(self.st[:, 2:5] + u)))


class GpuBackend:
    def __init__(self, R, ptol, ptol_rel):
        import cup3d_b200
        self.ctx = cup3d_b200.Context(0, 8)
        self.n = 0
        self.tol = (ptol, ptol_rel)

    def mesh(self, ib, rb, bpd, level_max):
        self.ctx.mesh_upload(ib, rb, bpd, level_max)
        self.n = len(ib)

    def h2d(self, st):
        self.ctx.state_h2d(np.ascontiguousarray(st))

    def d2h(self):
        out = np.zeros((self.n, 9, 512))
        self.ctx.state_d2h(out)
        return out

    def params(self, dt, uinf, step, lam):
        self.ctx.set_params(dt=dt, nu=NU, uinf=uinf, step=step, mean_constraint=2, ptol=self.tol[0],
                            ptol_rel=self.tol[1], lam=lam)

    def obstacle(self, k, blk, chi, udef, com):
        self.ctx.obstacle_upload(k, blk, chi, udef)
        self.ctx.obstacle_motion(k, com=com)

    def motion(self, k, vel, omega):
        self.ctx.obstacle_motion(k, vel=vel, omega=omega)

    def advdiff(self):
        self.ctx.advdiff()

    def moments(self, k):
        return self.ctx.obstacle_moments(k)

    def penalize(self):
        self.ctx.obstacle_penalize()

    def projection(self):
        self.ctx.projection()

    def umax(self):
        return self.ctx.umax()


def run_coupled(R, nsteps, dev, bpd, level_max):
    rec = {k: [] for k in ("dt", "nblk", "umax", "checks", "motion")}
    ib, rb = R.blocks()
    dev.mesh(ib, rb, bpd, level_max)
    dev.h2d(R.state_get())
    for _ in range(nsteps):
        # host phases read the fields (sta_umax inside sta_dt, the tagging of mesh_adapt)
        R.state_set(dev.d2h())
        R.sta_dt()
        if adapt_due(R):
            R.phase("mesh_adapt")
        R.phase("fish_build")
        sc = R.get_scalars()
        ib2, rb2 = R.blocks()
        if len(ib2) != len(ib) or not np.array_equal(ib2, ib):
            ib, rb = ib2, rb2
            dev.mesh(ib, rb, bpd, level_max)      # the rebuild hook
        dev.h2d(R.state_get())                    # F_CHI changed (fish_build); all fields if the mesh did
        dev.params(dt=sc["dt"], uinf=sc["uinf"], step=sc["step"], lam=sc["lam"])
        nf = R.nfish()
        for k in range(nf):
            blk, chi, udef = R.fish_obstacle(k)
            dev.obstacle(k, blk, chi, udef, R.fish_motion(k)[0])
        dev.advdiff()
        for k in range(nf):                       # fish_vel = moments (device) + fish_solve (host)
            R.fish_solve_from(k, dev.moments(k))
        R.fish_hit()                              # fish_pen = fish_hit (host) + block loop (device)
        for k in range(nf):
            _, vel, omega = R.fish_motion(k)
            dev.motion(k, vel, omega)
        dev.penalize()
        dev.projection()
        R.step_end()
        record(R, rec, dev.d2h(), dev.umax())
    return rec


def main():
    mode, case, nsteps, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    from oracle import refbind as R
    import make_golden_fish as MGF
    kw = dict(MGF.CASES[case])
    # both runs solve the pressure equation to the same tight tolerance, so they can be compared
    kw.update(poissonTol=1e-10, poissonTolRel=1e-12, nu=NU)
    R.init(**kw)
    R.sta_fields()
    args = R.default_args(**kw)
    bpd = (args["bpdx"], args["bpdy"], args["bpdz"])
    if mode == "pure":
        rec = run_pure(R, nsteps)
    elif mode == "ref":
        rec = run_coupled(R, nsteps, RefBackend(R), bpd, args["levelMax"])
    elif mode == "port":
        rec = run_coupled(R, nsteps, PortBackend(1e-10, 1e-12), bpd, args["levelMax"])
    else:
        rec = run_coupled(R, nsteps, GpuBackend(R, 1e-10, 1e-12), bpd, args["levelMax"])
    np.savez(out, **{k: np.array(v) for k, v in rec.items()})
    print(mode, case, "steps", nsteps, "nblk", rec["nblk"], "umax", rec["umax"])


if __name__ == "__main__":
    main()
