"""numpy restatement of the reference's pointwise obstacle phases.  TEST INFRASTRUCTURE ONLY:
nothing under cup3d_b200/ may import this (see oracle/__init__.py).

    moments  <- fish_mom_blk + the block sum of fish_vel   main.c:5057-5118, :5315-5325
    penalize <- fish_pen_blk                               main.c:5601-5650
    tmpv     <- fish_tmpv                                  main.c:5799-5827

Pinned against the reference itself in tests/golden/make_golden_fish.py (the generator asserts
that these functions reproduce the reference's output before it writes a fixture) and again in
tests/test_oracle_port.py against the committed fixtures.

Arguments shared by all three: ib/rb = block table (level, ix, iy, iz / h, origin), blk [nob] block
index of each ObstacleBlock, chi [nob,512], udef [nob,512,3] (struct ObstacleBlock, main.c:96-102).
"""
import numpy as np

M_V, M_FX, M_TX, M_J0, M_GFX, M_GPX, M_GJ0, M_GUX, M_GAX, M_N = 0, 1, 4, 7, 13, 14, 17, 23, 26, 29


def _pos(rb, blk):
    """blk_pos (main.c:1405): cell centres of the given blocks -> [nob, 512, 3]"""
    i = np.arange(8) + 0.5
    h = rb[blk, 0][:, None]
    x = rb[blk, 1][:, None] + h * i
    y = rb[blk, 2][:, None] + h * i
    z = rb[blk, 3][:, None] + h * i
    p = np.empty((len(blk), 8, 8, 8, 3))
    p[..., 0] = x[:, None, None, :]
    p[..., 1] = y[:, None, :, None]
    p[..., 2] = z[:, :, None, None]
    return p.reshape(len(blk), 512, 3)


def _inertia(f, p):
    """inertia_add (main.c:1006) for every cell -> [..., 6]"""
    return np.stack([f * (p[..., 1] ** 2 + p[..., 2] ** 2), f * (p[..., 0] ** 2 + p[..., 2] ** 2),
                     f * (p[..., 0] ** 2 + p[..., 1] ** 2), -f * p[..., 0] * p[..., 1], -f * p[..., 0] * p[..., 2],
                     -f * p[..., 1] * p[..., 2]], -1)


def moments(ib, rb, vel, blk, chi, udef, com, dt, lam):
    """vel [n,3,512] -> M[29] of one body"""
    M = np.zeros(M_N)
    if len(blk) == 0:
        return M
    p = _pos(rb, blk) - np.asarray(com)
    u = np.moveaxis(vel[blk], 1, 2)                    # [nob,512,3]
    du = u - udef
    on = chi > 0
    dv = (rb[blk, 0] ** 3)[:, None]
    lambdt = lam * dt
    X1 = (chi > 0.5).astype(float)
    pf = np.where(on, dv * lambdt * X1 / (1 + X1 * lambdt), 0.0)
    xv = np.where(on, chi * dv, 0.0)
    pxu, pxdu = np.cross(p, u), np.cross(p, du)
    M[M_V] = xv.sum()
    M[M_GFX] = pf.sum()
    M[M_J0:M_J0 + 6] = _inertia(xv, p).sum((0, 1))
    M[M_GJ0:M_GJ0 + 6] = _inertia(pf, p).sum((0, 1))
    M[M_FX:M_FX + 3] = (xv[..., None] * u).sum((0, 1))
    M[M_TX:M_TX + 3] = (xv[..., None] * pxu).sum((0, 1))
    M[M_GPX:M_GPX + 3] = (pf[..., None] * p).sum((0, 1))
    M[M_GUX:M_GUX + 3] = (pf[..., None] * du).sum((0, 1))
    M[M_GAX:M_GAX + 3] = (pf[..., None] * pxdu).sum((0, 1))
    return M


def penalize(ib, rb, vel, chi_field, blk, chi, udef, com, v, omega, dt, lam):
    """in place on vel [n,3,512]; chi_field [n,512] = F_CHI"""
    if len(blk) == 0:
        return
    p = _pos(rb, blk) - np.asarray(com)
    act = (chi_field[blk] <= chi) & (chi > 0)
    X = (chi > 0.5).astype(float)
    pen = X * lam / (1 + X * lam * dt)
    for d in range(3):
        e, f = (d + 1) % 3, (d + 2) % 3
        utot = v[d] + omega[e] * p[..., f] - omega[f] * p[..., e] + udef[..., d]
        u = vel[blk, d]
        vel[blk, d] = np.where(act, u + dt * (pen * (utot - u)), u)


def tmpv(tmp, chi_field, blk, chi, udef):
    """in place on tmp [n,3,512]"""
    if len(blk) == 0:
        return
    act = chi_field[blk] <= chi
    for d in range(3):
        tmp[blk, d] += np.where(act, udef[..., d], 0.0)
