"""GPU parity on MULTI-LEVEL (AMR) meshes produced by the reference's own
mesh_adapt: ghost faces from coarser leaves (OP_FD interpolation), from finer
leaves (OP_AVG8) and the flux correction at coarse-fine faces (fc_fill), inside
pois_op, k_lhs, the V-cycle (levels with pass-through leaves) and the full
solve.  Goldens: tests/golden/amr2.npz (2 levels), amr3.npz (3 levels)."""
import numpy as np
import pytest

from util import AMR_CASES, case, relerr
from cup3d_b200 import capi

pytestmark = pytest.mark.gpu


def make_ctx(c, **params):
    import cup3d_b200
    ctx = cup3d_b200.Context(0, 8)
    ctx.mesh_upload(c.ib, c.rb, c.bpd, c.level_max)
    ctx.set_params(dt=c.dt, nu=c.nu, uinf=c.uinf, step=5, mean_constraint=2, **params)
    return ctx


@pytest.mark.parametrize("name", AMR_CASES)
def test_amr_levels(built, name):
    c = case(name)
    ctx = make_ctx(c)
    counts = np.bincount(c.ib[:, 0], minlength=c.level_max)
    finest = int(np.max(c.ib[:, 0]))
    assert ctx.mg_nact(finest) == counts[finest]
    # every coarser level = its own leaves + one parent per 8 blocks of the level above
    n_above = counts[finest]
    for L in range(finest - 1, -1, -1):
        assert ctx.mg_nact(L) == counts[L] + n_above // 8
        n_above = ctx.mg_nact(L)
    ctx.close()


@pytest.mark.parametrize("name", AMR_CASES)
def test_amr_pois_op(built, name):
    c = case(name)
    for mc in (0, 2):
        key = "op_out_mc%d" % mc
        if key not in c.g:
            continue
        ctx = make_ctx(c)
        ctx.set_params(mean_constraint=mc)
        out = ctx.pois_op(np.ascontiguousarray(c.F["pres"]))
        e = relerr(out, c.g[key])
        assert e < 1e-12, (mc, e)
        ctx.close()


def test_amr_k_lhs_with_flux_correction(built):
    c = case("amr2")
    ctx = make_ctx(c)
    s0 = c.state0()
    ctx.state_h2d(s0)
    ctx.stencil_apply(capi.ST_LHS)
    out = np.zeros_like(s0)
    ctx.state_d2h(out)
    assert relerr(out[:, 8:9], c.g["st_lhs"]) < 1e-12
    ctx.close()


@pytest.mark.parametrize("name", AMR_CASES)
def test_amr_vcycle(built, name):
    c = case(name)
    ctx = make_ctx(c)
    out = ctx.mg_vcycle(np.ascontiguousarray(c.F["cosrhs"]))
    e = relerr(out, c.g["vc_out_cosrhs"])
    assert e < 1e-11, e
    ctx.close()


def test_amr_pois_solve(built):
    c = case("amr2")
    ctx = make_ctx(c, ptol=1e-10, ptol_rel=1e-12)
    st = c.state0()
    st[:, 8] = c.solve_rhs()
    st[:, 1] = 0
    ctx.state_h2d(st)
    info = ctx.pois_solve()
    out = np.zeros_like(st)
    ctx.state_d2h(out, 1, 1)
    assert info.residual < max(1e-10, 1e-12 * info.rhs_norm)
    assert relerr(out[:, 1], c.g["solve_x_mc2"]) < 1e-8
    ctx.close()


@pytest.mark.parametrize("st,sid,f0,nc", [("prhs", capi.ST_PRHS, 8, 1), ("divp", capi.ST_DIVP, 5, 1),
                                          ("gradp", capi.ST_GRADP, 5, 3), ("advdiff", capi.ST_ADVDIFF, 5, 3),
                                          ("vort", capi.ST_VORT, 5, 3), ("q", capi.ST_Q, 8, 1)])
def test_amr_projection_sweeps(built, st, sid, f0, nc):
    """k_prhs / k_divp / k_gradp / k_vort with their flux correction (and k_q) on a 2-level mesh"""
    c = case("amr2")
    ctx = make_ctx(c)
    s0 = c.state0()
    ctx.state_h2d(s0)
    ctx.stencil_apply(sid)
    out = np.zeros_like(s0)
    ctx.state_d2h(out)
    assert relerr(out[:, f0:f0 + nc], c.g["st_" + st]) < 1e-12, st
    ctx.close()


def test_amr_projection(built):
    c = case("amr2")
    ctx = make_ctx(c, ptol=1e-10, ptol_rel=1e-12)
    s0 = c.state0()
    ctx.state_h2d(s0)
    info = ctx.projection()
    out = np.zeros_like(s0)
    ctx.state_d2h(out)
    ref = c.g["proj_step5"]
    assert info.residual < max(1e-10, 1e-12 * info.rhs_norm)
    assert relerr(out[:, 1], ref[:, 0]) < 1e-7
    assert relerr(out[:, 2:5], ref[:, 1:4]) < 1e-9
    ctx.close()


def test_amr_advdiff_rk3(built):
    """advdiff() (three k_advdiff sweeps with the ss = 3 coarse-fine ghost fill + RK3 updates)"""
    c = case("amr2")
    ctx = make_ctx(c)
    s0 = c.state0()
    ctx.state_h2d(s0)
    ctx.advdiff()
    out = np.zeros_like(s0)
    ctx.state_d2h(out)
    assert relerr(out[:, 2:5], c.g["advdiff"][:, 0:3]) < 1e-12
    ctx.close()


def test_amr_vorticity_block_norms(built):
    """vorticity() + mesh_tag_blk's per-block norm, without and with k_gradchi's zeroing (chi > 0.9)"""
    c = case("amr2")
    ctx = make_ctx(c)
    ctx.state_h2d(c.state0())
    ctx.vorticity()
    la, lf = ctx.block_linf()
    assert np.max(np.abs(la - c.g["tag_linf_all"]) / c.g["tag_linf_all"]) < 1e-12
    assert np.max(np.abs(lf - c.g["tag_linf_fluid"])) < 1e-12 * np.max(c.g["tag_linf_all"])
    # blocks the reference's k_gradchi did not mark: the fluid norm IS mesh_tag_blk's Linf
    um = ~c.g["tag_marked"]
    assert np.max(np.abs(lf[um] - c.g["tag_linf"][um])) < 1e-12 * np.max(c.g["tag_linf_all"])
    ctx.close()
