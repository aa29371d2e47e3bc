"""CPU: the C restatement (oracle/cup_oracle.c) against the golden vectors the
REFERENCE produced (tests/golden, generated from oracle/_ref).  This is what
pins the oracle port: every function of the hot path it restates must
reproduce the reference's output on the same inputs."""
import numpy as np
import pytest

from util import ALL_CASES, STENCIL_CASES, case, relerr

ST = {"lhs": (0, 8, 1), "advdiff": (2, 5, 3), "prhs": (3, 8, 1), "divp": (4, 5, 1), "gradp": (5, 5, 3),
      "vort": (6, 5, 3), "q": (7, 8, 1)}


@pytest.fixture(scope="module")
def P(built):
    from oracle import portbind
    return portbind


def orc(P, c):
    return P.Oracle(c.ib, c.rb, c.bpd, c.level_max)


@pytest.mark.parametrize("name", ALL_CASES)
def test_port_vcycle(P, name):
    c = case(name)
    o = orc(P, c)
    for rhs in ("cosrhs", "rand"):
        assert relerr(o.mg_vcycle(c.F[rhs]), c.g["vc_out_" + rhs]) < 1e-13
    o.close()


@pytest.mark.parametrize("name", ALL_CASES)
def test_port_pois_op_and_dot(P, name):
    c = case(name)
    o = orc(P, c)
    for mc in (0, 1, 2, 3):
        key = "op_out_mc%d" % mc
        if key in c.g:
            assert relerr(o.pois_op(c.F["pres"], mc), c.g[key]) < 1e-13
    d = o.pois_dot(c.F["pres"], c.F["rand"])
    assert abs(d - float(c.g["dot_ab"])) <= 1e-12 * abs(float(c.g["dot_ab"]))
    o.close()


@pytest.mark.parametrize("name", STENCIL_CASES)
def test_port_stencils_and_advdiff(P, name):
    c = case(name)
    o = orc(P, c)
    for st, (sid, f0, nc) in ST.items():
        s = c.state0()
        o.stencil(sid, s, c.dt, c.nu, c.uinf)
        assert relerr(s[:, f0:f0 + nc], c.g["st_" + st]) < 1e-14, st
    s = c.state0()
    o.advdiff(s, c.dt, c.nu, c.uinf)
    assert relerr(s[:, 2:5], c.g["advdiff"][:, 0:3]) < 1e-14
    o.close()


@pytest.mark.parametrize("name", ["u16", "b211", "b321", "u32"])
def test_port_solve(P, name):
    c = case(name)
    o = orc(P, c)
    s = c.state0()
    s[:, 8] = c.solve_rhs()
    s[:, 1] = 0
    it, res = o.pois_solve(s, 2, 1e-10, 1e-12)
    assert res < 1e-10
    assert relerr(s[:, 1], c.g["solve_x_mc2"]) < 1e-9
    o.close()


@pytest.mark.parametrize("name", ["u16", "b211"])
@pytest.mark.parametrize("step", [1, 5])
def test_port_projection(P, name, step):
    c = case(name)
    o = orc(P, c)
    s = c.state0()
    o.projection(s, c.dt, c.nu, c.uinf, step, 2, 1e-10, 1e-12)
    ref = c.g["proj_step%d" % step]
    assert relerr(s[:, 1], ref[:, 0]) < 1e-8
    assert relerr(s[:, 2:5], ref[:, 1:4]) < 1e-10
    o.close()


def test_port_fdm_block_inverse_is_exact_inverse(P):
    """pre_blk (main.c:4368) inverts the 7-point Dirichlet Laplacian of one block exactly:
    the identity the CUDA smoother's algebraic form relies on."""
    c = case("b222_l0")
    o = orc(P, c)
    rng = np.random.default_rng(7)
    u = rng.standard_normal((8, 8, 8))
    up = np.pad(u, 1)
    lap = (up[1:-1, 1:-1, :-2] + up[1:-1, 1:-1, 2:] + up[1:-1, :-2, 1:-1] + up[1:-1, 2:, 1:-1] +
           up[:-2, 1:-1, 1:-1] + up[2:, 1:-1, 1:-1] - 6 * u)
    h = 0.37
    back = o.pre_blk((h * lap).reshape(-1), 1 / h)
    assert relerr(back, u.reshape(-1)) < 1e-13
    o.close()


# ---- the numpy restatement of the pointwise obstacle phases (oracle/fish_port.py) ----------
from fishutil import FISH_CASES, fish_case  # noqa: E402


@pytest.mark.parametrize("name", FISH_CASES)
def test_fish_port(name):
    from oracle import fish_port as FP
    c = fish_case(name)
    vel = np.zeros((c.n, 3, 512))
    vel[c.obu] = c.g["adv_vel"]          # velocity after advdiff (obstacle blocks are all that matter)
    for k in range(c.nfish):
        M = FP.moments(c.ib, c.rb, vel, *c.obs[k], c.com[k], c.dt, c.lam)
        assert np.max(np.abs(M - c.mom[k])) < 1e-12 * np.max(np.abs(c.mom[k]))
    for k in range(c.nfish):
        FP.penalize(c.ib, c.rb, vel, c.chi_field, *c.obs[k], c.com[k], c.vel[k], c.omega[k], c.dt, c.lam)
    assert relerr(vel[c.obu], c.g["pen_vel"]) < 1e-14
    tmp = np.zeros((c.n, 3, 512))
    for k in range(c.nfish):
        FP.tmpv(tmp, c.chi_field, *c.obs[k])
    assert np.array_equal(tmp[c.obu], c.g["tmpv"])
    assert not np.any(np.delete(tmp, c.obu, axis=0))
