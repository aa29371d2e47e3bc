"""GPU parity of the obstacle (fish) phases of one time step against the reference run phase by
phase (tests/golden/make_golden_fish.py): advdiff -> fish_mom_blk + block sum -> fish_pen_blk ->
fish_tmpv -> projection (advance(), main.c:5993-5997), on a uniform mesh with two bodies and on the
reference's own adapted (2-level) mesh around one body.  The bodies (chi, udef, com, vel, omega) are
the reference's: fish_build / fish_solve are host code on both sides."""
import numpy as np
import pytest

from util import relerr
from fishutil import FISH_CASES, checksums, fish_case

pytestmark = pytest.mark.gpu


def make_ctx(c, real_bytes=8):
    import cup3d_b200
    ctx = cup3d_b200.Context(0, real_bytes)
    ctx.mesh_upload(c.ib, c.rb, c.bpd, c.level_max)
    ctx.set_params(dt=c.dt, nu=c.nu, uinf=c.uinf, step=c.step, mean_constraint=2, ptol=1e-10, ptol_rel=1e-12,
                   lam=c.lam)
    for k in range(c.nfish):
        ctx.obstacle_upload(k, *c.obs[k])
        ctx.obstacle_motion(k, com=c.com[k])
    return ctx


def get(ctx, c):
    out = np.zeros((c.n, 9, 512))
    ctx.state_d2h(out)
    return out


@pytest.mark.parametrize("name", FISH_CASES)
def test_fish_step(built, name):
    c = fish_case(name)
    ctx = make_ctx(c)
    ctx.state_h2d(c.state0())
    # advdiff
    ctx.advdiff()
    s1 = get(ctx, c)
    assert relerr(s1[c.obu, 2:5], c.g["adv_vel"]) < 1e-12
    assert relerr(checksums(s1[:, 2:5]), c.g["adv_sums"]) < 1e-11
    # fish_mom_blk + sum over blocks
    for k in range(c.nfish):
        M = ctx.obstacle_moments(k)
        assert np.max(np.abs(M - c.mom[k])) < 1e-11 * np.max(np.abs(c.mom[k])), (k, M, c.mom[k])
    # fish_solve is the reference's: vel / omega come from the fixture
    for k in range(c.nfish):
        ctx.obstacle_motion(k, vel=c.vel[k], omega=c.omega[k])
    ctx.obstacle_penalize()
    s2 = get(ctx, c)
    assert relerr(s2[c.obu, 2:5], c.g["pen_vel"]) < 1e-12
    assert np.array_equal(np.delete(s2, c.obu, axis=0), np.delete(s1, c.obu, axis=0))
    # projection: zero F_TMP, fish_tmpv, prhs, divp, solve, gradp
    info = ctx.projection()
    s5 = get(ctx, c)
    assert info.residual <= max(1e-10, 1e-12 * info.rhs_norm) * 1.0001
    # same bars as test_projection (both solvers stop at the tolerance, not at round-off)
    ep = relerr(s5[c.obu, 1], c.g["proj"][:, 0])
    ev = relerr(s5[c.obu, 2:5], c.g["proj"][:, 1:4])
    esp = relerr(checksums(s5[:, 1:2]), c.g["proj_sums"][:, 0:1])
    esv = relerr(checksums(s5[:, 2:5]), c.g["proj_sums"][:, 1:4])
    assert ep < 1e-7 and ev < 1e-9 and esp < 1e-7 and esv < 1e-9, (ep, ev, esp, esv, info.iterations)
    # mesh_adapt's tagging input on the projected state: vorticity() and the per-block norms; where the
    # reference's k_gradchi did not mark the block, the fluid norm is mesh_tag_blk's Linf
    ctx.vorticity()
    la, lf = ctx.block_linf()
    scale = np.max(c.g["tag_linf_all"])
    assert np.max(np.abs(la - c.g["tag_linf_all"])) < 1e-7 * scale
    assert np.max(np.abs(lf - c.g["tag_linf_fluid"])) < 1e-7 * scale
    um = ~c.g["tag_marked"]
    assert um.sum() > 0.5 * c.n and np.max(np.abs(lf[um] - c.g["tag_linf"][um])) < 1e-7 * scale
    ctx.close()


@pytest.mark.parametrize("name", FISH_CASES)
def test_fish_tmpv(built, name):
    c = fish_case(name)
    ctx = make_ctx(c)
    ctx.state_h2d(c.state0())       # F_TMP = 0
    ctx.obstacle_tmpv()
    s = get(ctx, c)
    assert np.array_equal(s[c.obu, 5:8], c.g["tmpv"])
    assert not np.any(np.delete(s[:, 5:8], c.obu, axis=0))
    ctx.close()


def test_fish_reupload_and_clear(built):
    """bodies deform every step: a second upload replaces the first; nob = 0 / clear remove them"""
    c = fish_case("fish64")
    ctx = make_ctx(c)
    ctx.state_h2d(c.state0())
    blk, chi, udef = c.obs[0]
    ctx.obstacle_upload(0, blk[:3], chi[:3], udef[:3])       # shrink
    ctx.obstacle_upload(0, blk, chi, udef)                   # grow back
    ctx.obstacle_tmpv()
    s = get(ctx, c)
    assert np.array_equal(s[c.obu, 5:8], c.g["tmpv"])
    ctx.obstacle_clear()
    ctx.state_h2d(c.state0())
    ctx.obstacle_tmpv()
    ctx.obstacle_penalize()
    assert np.array_equal(get(ctx, c), c.state0())
    assert not np.any(ctx.obstacle_moments(0))
    with pytest.raises(Exception):
        ctx.obstacle_upload(0, np.array([c.n], np.int32), chi[:1], udef[:1])   # block out of range
    ctx.close()


def test_fish_fp32(built):
    c = fish_case("fish64")
    ctx = make_ctx(c, real_bytes=4)
    st = c.state0()
    st[c.obu, 2:5] = c.g["adv_vel"]
    ctx.state_h2d(st)
    for k in range(c.nfish):
        M = ctx.obstacle_moments(k)
        assert np.max(np.abs(M - c.mom[k])) < 1e-5 * np.max(np.abs(c.mom[k]))
        ctx.obstacle_motion(k, vel=c.vel[k], omega=c.omega[k])
    ctx.obstacle_penalize()
    s = get(ctx, c)
    assert relerr(s[c.obu, 2:5], c.g["pen_vel"]) < 1e-5
    ctx.close()
