"""GPU, Real = float (the reference's one-line `typedef float Real` build, SURVEY.md section 0;
config 3 of BASELINE.json is fp32): the same kernels instantiated for float, checked against the
fp64 goldens with single-precision tolerances."""
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
