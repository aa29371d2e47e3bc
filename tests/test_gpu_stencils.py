"""GPU parity of the per-block stencil sweeps and the advdiff / projection
drivers against golden states produced by the reference (stencil_apply(&st_*),
advdiff(), projection(); main.c:3648, :5027, :5828), through the C ABI."""
import numpy as np
import pytest

from util import STENCIL_CASES, SOLVE_CASES, case, relerr
from cup3d_b200 import capi

pytestmark = pytest.mark.gpu

# stencil -> (C ABI id, first output field, components)
ST = {"lhs": (capi.ST_LHS, capi.F_LHS, 1), "advdiff": (capi.ST_ADVDIFF, capi.F_TMP, 3),
      "prhs": (capi.ST_PRHS, capi.F_LHS, 1), "divp": (capi.ST_DIVP, capi.F_TMP, 1),
      "gradp": (capi.ST_GRADP, capi.F_TMP, 3), "vort": (capi.ST_VORT, capi.F_TMP, 3),
      "q": (capi.ST_Q, capi.F_LHS, 1)}


def make_ctx(c, step=5, **params):
    import cup3d_b200
    ctx = cup3d_b200.Context(0, 8)
    ctx.mesh_upload(c.ib, c.rb, c.bpd, c.level_max)
    ctx.set_params(dt=c.dt, nu=c.nu, uinf=c.uinf, step=step, mean_constraint=2, **params)
    return ctx


@pytest.mark.parametrize("name", STENCIL_CASES)
@pytest.mark.parametrize("st", list(ST))
def test_stencil_sweep(built, name, st):
    c = case(name)
    sid, f0, nc = ST[st]
    ctx = make_ctx(c)
    s0 = c.state0()
    ctx.state_h2d(s0)
    ctx.stencil_apply(sid)
    out = np.zeros_like(s0)
    ctx.state_d2h(out)
    ref = c.g["st_" + st]
    e = relerr(out[:, f0:f0 + nc], ref)
    assert e < 1e-12, (st, e)
    # nothing else may change
    mask = np.ones(9, bool)
    mask[f0:f0 + nc] = False
    assert np.array_equal(out[:, mask], s0[:, mask])
    ctx.close()


@pytest.mark.parametrize("name", STENCIL_CASES)
def test_advdiff_rk3(built, name):
    c = case(name)
    ctx = make_ctx(c)
    s0 = c.state0()
    ctx.state_h2d(s0)
    ctx.advdiff()
    out = np.zeros_like(s0)
    ctx.state_d2h(out)
    ref = c.g["advdiff"]  # VEL(3), TMP(3)
    assert relerr(out[:, 2:5], ref[:, 0:3]) < 1e-12
    assert np.max(np.abs(out[:, 5:8] - ref[:, 3:6])) <= 1e-12 * np.max(np.abs(ref[:, 0:3]))
    ctx.close()


@pytest.mark.parametrize("name", [n for n in STENCIL_CASES if n in SOLVE_CASES])
@pytest.mark.parametrize("step", [1, 5])
def test_projection(built, name, step):
    c = case(name)
    ctx = make_ctx(c, step=step, ptol=1e-10, ptol_rel=1e-12)
    s0 = c.state0()
    ctx.state_h2d(s0)
    info = ctx.projection()
    out = np.zeros_like(s0)
    ctx.state_d2h(out)
    ref = c.g["proj_step%d" % step]  # PRES, VEL(3)
    # pois_solve stops on ptol OR ptol_rel*|rhs| (main.c:4915)
    assert info.residual < max(1e-10, 1e-12 * info.rhs_norm)
    ep = relerr(out[:, 1], ref[:, 0])
    ev = relerr(out[:, 2:5], ref[:, 1:4])
    assert ep < 1e-7 and ev < 1e-9, (ep, ev, info.iterations)
    ctx.close()


@pytest.mark.parametrize("name", STENCIL_CASES + ["amr2"])
def test_umax(built, name):
    """sta_umax (main.c:5918) on the device"""
    c = case(name)
    ctx = make_ctx(c)
    ctx.state_h2d(c.state0())
    assert abs(ctx.umax() - float(c.g["umax"])) <= 1e-15 * float(c.g["umax"])
    ctx.close()


@pytest.mark.parametrize("name", STENCIL_CASES)
def test_vorticity_block_norms(built, name):
    """vorticity() (k_vort, 1/h^3) and mesh_tag_blk's per-block norm (main.c:3683), all cells / chi <= 0.9"""
    c = case(name)
    ctx = make_ctx(c)
    ctx.state_h2d(c.state0())
    ctx.vorticity()
    la, lf = ctx.block_linf()
    assert np.max(np.abs(la - c.g["tag_linf_all"]) / c.g["tag_linf_all"]) < 1e-12
    assert np.max(np.abs(lf - c.g["tag_linf_fluid"])) < 1e-12 * np.max(c.g["tag_linf_all"])
    ctx.close()


@pytest.mark.parametrize("name", STENCIL_CASES + ["amr2"])
def test_io_pack(built, name):
    """io_dump's field arrays (main.c:1441-1442, :1525-1535): vorticity(); qcrit(); float32 packing.
    Expected = the reference's k_vort / k_q sweeps, scaled and converted exactly as the reference does."""
    c = case(name)
    ctx = make_ctx(c)
    s0 = c.state0()
    ctx.state_h2d(s0)
    attr, vort, q = ctx.io_pack()
    h = c.rb[:, 0]
    fac = (1.0 / (h * h * h))[:, None, None]
    ev = np.moveaxis(c.g["st_vort"] * fac, 1, 2).astype(np.float32)       # [n,512,3]
    assert np.array_equal(attr, s0[:, 0].astype(np.float32))
    # float32 of two doubles that agree to 1e-12 can differ by one float ulp at a rounding boundary
    assert np.max(np.abs(vort - ev)) <= 2e-7 * np.max(np.abs(ev))
    assert np.mean(vort != ev) < 1e-3
    eq = c.g["st_q"][:, 0].astype(np.float32)
    assert np.max(np.abs(q - eq)) <= 2e-7 * np.max(np.abs(eq)) and np.mean(q != eq) < 1e-3
    ctx.close()


@pytest.mark.parametrize("name", ["b211", "amr2"])
@pytest.mark.parametrize("st", list(ST))
def test_stencil_run_block_list(built, name, st):
    """stencil_run(st, list, n) (main.c:3631-3647): the listed blocks get exactly what the full sweep gives
    them, every other block keeps its output; list == NULL with n < nblk means the first n blocks"""
    c = case(name)
    sid, f0, nc = ST[st]
    ctx = make_ctx(c)
    s0 = c.state0()
    ref = c.g["st_" + st]
    rng = np.random.default_rng(7)
    lst = rng.permutation(c.n)[: max(1, c.n // 3)]
    for blocks, n in ((lst, None), (None, c.n // 2)):
        ctx.state_h2d(s0)
        ctx.stencil_run(sid, blocks, n)
        out = np.zeros_like(s0)
        ctx.state_d2h(out)
        sel = np.zeros(c.n, bool)
        sel[blocks if blocks is not None else np.arange(n)] = True
        scale = np.max(np.abs(ref))
        assert np.max(np.abs(out[sel, f0:f0 + nc] - ref[sel])) <= 1e-12 * scale, st
        assert np.array_equal(out[~sel], s0[~sel]), st
    ctx.close()


def test_stencil_run_rejects_bad_lists(built):
    import cup3d_b200
    c = case("u16")
    ctx = make_ctx(c)
    for bad in ([0, 0], [c.n], [-1]):
        with pytest.raises(cup3d_b200.CupError):
            ctx.stencil_run(capi.ST_DIVP, np.array(bad))
    with pytest.raises(cup3d_b200.CupError):
        ctx.stencil_run(capi.ST_DIVP, None, c.n + 1)
    ctx.close()
