"""GPU, Real = float (the reference's one-line `typedef float Real` build, SURVEY.md section 0;
config 3 of BASELINE.json is fp32): the same kernels instantiated for float.

Pinned twice: against the reference ITSELF built with Real = float (oracle/_ref/libcup3d_ref32.so,
oracle/Makefile target ref32) run live in a subprocess on the same inputs -- single-precision
rounding-level tolerances, the differences are operation order only -- and, as a sanity bound on
the precision itself, against the fp64 goldens with single-precision tolerances."""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from util import case, relerr
from cup3d_b200 import capi

pytestmark = pytest.mark.gpu


def make_ctx(c, **params):
    import cup3d_b200
    ctx = cup3d_b200.Context(0, 4)
    ctx.mesh_upload(c.ib, c.rb, c.bpd, c.level_max)
    ctx.set_params(dt=c.dt, nu=c.nu, uinf=c.uinf, step=5, mean_constraint=2, **params)
    return ctx


@pytest.mark.parametrize("name", ["u16", "u32", "b211", "amr2"])
def test_fp32_vcycle_and_op(built, name):
    c = case(name)
    ctx = make_ctx(c)
    out = ctx.mg_vcycle(np.ascontiguousarray(c.F["cosrhs"]))
    assert relerr(out, c.g["vc_out_cosrhs"]) < 2e-4
    op = ctx.pois_op(np.ascontiguousarray(c.F["pres"]))
    assert relerr(op, c.g["op_out_mc2"]) < 2e-5
    ctx.close()


@pytest.mark.parametrize("name", ["u16", "b211"])
def test_fp32_sweeps(built, name):
    c = case(name)
    for st, (sid, f0, nc) in {"advdiff": (capi.ST_ADVDIFF, 5, 3), "prhs": (capi.ST_PRHS, 8, 1),
                              "gradp": (capi.ST_GRADP, 5, 3)}.items():
        ctx = make_ctx(c)
        s0 = c.state0()
        ctx.state_h2d(s0)
        ctx.stencil_apply(sid)
        out = np.zeros_like(s0)
        ctx.state_d2h(out)
        assert relerr(out[:, f0:f0 + nc], c.g["st_" + st]) < 5e-5, st
        ctx.close()


def test_fp32_solve(built):
    c = case("u32")
    ctx = make_ctx(c, ptol=1e-5, ptol_rel=1e-5)
    st = c.state0()
    st[:, 8] = c.solve_rhs()
    st[:, 1] = 0
    ctx.state_h2d(st)
    info = ctx.pois_solve()
    out = np.zeros_like(st)
    ctx.state_d2h(out, 1, 1)
    assert info.residual < 1e-5 * max(1.0, info.rhs_norm)
    assert relerr(out[:, 1], c.g["solve_x_mc2"]) < 1e-3
    ctx.close()


# ---- against the reference built with Real = float, run live -------------------------------------
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ref32(name, d, ops, ptol=1e-6, ptol_rel=1e-4):
    lib = os.path.join(ROOT, "oracle", "_ref", "libcup3d_ref32.so")
    if not os.path.exists(lib):
        pytest.skip("oracle/_ref/libcup3d_ref32.so not built")
    c = case(name)
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "live_ref.py"), "--case", name, "--dir", d, "--ops", ops,
           "--real", "4", "--dt", repr(c.dt), "--nu", repr(c.nu), "--uinf", ",".join(repr(v) for v in c.uinf),
           "--ptol", repr(ptol), "--ptol-rel", repr(ptol_rel)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    # the Real = float reference must have built the same mesh (its tagging runs in float)
    if not np.array_equal(np.load(os.path.join(d, "ib.npy")), c.ib):
        pytest.skip("the float reference adapted to a different mesh")


@pytest.mark.parametrize("name", ["u32", "b211", "amr2"])
def test_fp32_against_float_reference(built, name):
    """mg_vcycle, pois_op, advdiff(), projection() in float vs the Real = float reference"""
    c = case(name)
    d = tempfile.mkdtemp(prefix="cup_ref32_")
    try:
        x = np.ascontiguousarray(c.F["cosrhs"].astype(np.float32).astype(np.float64))
        s0 = np.ascontiguousarray(c.state0().astype(np.float32).astype(np.float64))
        np.save(os.path.join(d, "in_vec.npy"), x)
        np.save(os.path.join(d, "in_state.npy"), s0)
        ref32(name, d, "vcycle,op,advdiff,proj")
        ctx = make_ctx(c, ptol=1e-6, ptol_rel=1e-4)  # the reference's default tolerances (run.sh): reachable in float
        e = {}
        e["vcycle"] = relerr(ctx.mg_vcycle(x), np.load(os.path.join(d, "out_vcycle.npy")))
        e["op"] = relerr(ctx.pois_op(x), np.load(os.path.join(d, "out_op.npy")))
        out = np.zeros_like(s0)
        ctx.state_h2d(s0)
        ctx.advdiff()
        ctx.state_d2h(out, 2, 3)
        e["advdiff"] = relerr(out[:, 2:5], np.load(os.path.join(d, "out_advdiff.npy")))
        ctx.state_h2d(s0)
        ctx.projection()
        ctx.state_d2h(out, 1, 4)
        ref = np.load(os.path.join(d, "out_proj.npy"))
        e["proj_p"] = relerr(out[:, 1], ref[:, 0])
        e["proj_v"] = relerr(out[:, 2:5], ref[:, 1:4])
        ctx.close()
        print("fp32 vs Real=float reference, %s: %s" % (name, {k: "%.2e" % v for k, v in e.items()}))
        # float epsilon 6e-8; a V-cycle amplifies rounding by the conditioning of ~100 smoothing sweeps
        # measured on the B200 (r02): vcycle 1.0-1.2e-6, op <= 1.8e-7, advdiff 1.2e-7, proj_p 1.4-1.9e-6, proj_v 2.6-5.2e-7
        assert e["op"] < 1e-6 and e["advdiff"] < 1e-6
        assert e["vcycle"] < 1e-5
        assert e["proj_v"] < 1e-5 and e["proj_p"] < 5e-5  # both solves stop at the same relative residual 1e-4
    finally:
        shutil.rmtree(d, ignore_errors=True)
