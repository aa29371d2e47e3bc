"""GPU vs the reference RUN LIVE at BASELINE.json's full sizes.

The stored goldens stop at 64^3 (fixture size); the 262 144-block path of the 512^3 benchmark
(boundary-first ordering, persistent grids, > 2^31-byte offsets) and the 256^3 time-step pieces are
compared here with the unmodified reference (oracle/_ref, built by oracle/Makefile, travels with the
snapshot) executed in a subprocess on the same inputs -- not with properties.  Tolerances: fp64,
sweeps/operators 1e-12, V-cycle 1e-11 (FMA contraction, even/odd transform factorisation),
projection: pressure 1e-7 / velocity 1e-9 relative (both solvers stop at a residual of 1e-9).
512^3 needs ~20 GB of host RAM for the reference and about two minutes of its mesh_init: it runs
when >= 48 GB are free (CUP_LIVE_512=0 skips it, =1 forces it).
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from util import relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ref_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libcup3d_ref.so"))


def free_ram_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def run_ref(level, d, ops, real=8, timeout=1500):
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "live_ref.py"), "--level", str(level), "--dir", d, "--ops", ops,
           "--real", str(real)]
    env = dict(os.environ)  # OMP_NUM_THREADS: conftest.py's bounded default (the reference does not scale past ~24)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, r.stdout[-3000:]


def scratch():
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    return tempfile.mkdtemp(prefix="cup_live_", dir=base)


def vec_input(ib, rb):
    """cosine right-hand side + a +-1 point-source pair + deterministic roughness"""
    from cup3d_b200 import mesh
    n = len(ib)
    out = np.empty((n, 512))
    step = 4096
    for s in range(0, n, step):  # in slices: the 512^3 coordinate arrays would not fit comfortably
        X, Y, Z = mesh.cell_centers(ib[s:s + step], rb[s:s + step])
        h = rb[s:s + step, 0][:, None, None, None]
        out[s:s + step] = (h ** 3 * (np.cos(np.pi * X) * np.cos(2 * np.pi * Y) * np.cos(3 * np.pi * Z) +
                                     0.05 * np.sin(37 * np.pi * X) * np.sin(53 * np.pi * Y) * np.cos(29 * np.pi * Z))
                           ).reshape(-1, 512)
    lo, hi = rb[:, 1:4], rb[:, 1:4] + 8 * rb[:, 0:1]
    for p, v in ((0.25, 1.0), (0.75, -1.0)):
        i = int(np.nonzero(np.all((lo <= p) & (p < hi), axis=1))[0][0])
        out[i, 0] += v * rb[i, 0] ** 3 * 50
    return out


@pytest.mark.parametrize("level", [5, 6])
def test_vcycle_and_operator_full_size(built, level):
    """mg_vcycle (main.c:4831) and pois_op (main.c:4282) at 256^3 / 512^3 against the live reference"""
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    if level == 6:
        flag = os.environ.get("CUP_LIVE_512", "")
        if flag == "0" or (flag != "1" and free_ram_gb() < 48):
            pytest.skip("512^3 live reference needs >= 48 GB free host RAM")
    import cup3d_b200
    from cup3d_b200 import mesh
    ib, rb = mesh.uniform_blocks(level)
    x = vec_input(ib, rb)
    d = scratch()
    try:
        np.save(os.path.join(d, "in_vec.npy"), x)
        run_ref(level, d, "vcycle,op")
        assert np.array_equal(np.load(os.path.join(d, "ib.npy")), ib)  # same block order
        ctx = cup3d_b200.Context(0, 8)
        ctx.mesh_upload(ib, rb, (1, 1, 1), level + 1)
        ctx.set_params(mean_constraint=2)
        got = ctx.mg_vcycle(x)
        ref = np.load(os.path.join(d, "out_vcycle.npy"))
        e_vc = relerr(got, ref)
        got = ctx.pois_op(x)
        ref = np.load(os.path.join(d, "out_op.npy"))
        e_op = relerr(got, ref)
        ctx.close()
        print("live %d^3: vcycle rel err %.3e, pois_op rel err %.3e" % (8 << level, e_vc, e_op))
        assert e_vc < 1e-11
        assert e_op < 1e-12
    finally:
        shutil.rmtree(d, ignore_errors=True)


def state_input(ib, rb):
    from cup3d_b200 import mesh
    n = len(ib)
    st = np.zeros((n, 9, 512))
    step = 4096
    for s in range(0, n, step):
        X, Y, Z = mesh.cell_centers(ib[s:s + step], rb[s:s + step])
        m = len(X)
        st[s:s + step, 2] = (np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y) + 0.02 * np.sin(41 * np.pi * Z)).reshape(m, 512)
        st[s:s + step, 3] = (-np.cos(2 * np.pi * X) * np.sin(2 * np.pi * Y) + 0.02 * np.cos(33 * np.pi * X)).reshape(m, 512)
        st[s:s + step, 4] = (0.1 * np.sin(2 * np.pi * Z) + 0.02 * np.sin(27 * np.pi * Y)).reshape(m, 512)
        r2 = (X - 0.5) ** 2 + (Y - 0.45) ** 2 + (Z - 0.55) ** 2
        st[s:s + step, 0] = (1.0 / (1.0 + np.exp((np.sqrt(r2) - 0.2) / 0.03))).reshape(m, 512)
        st[s:s + step, 1] = (np.cos(np.pi * X) * np.cos(np.pi * Y) * np.cos(2 * np.pi * Z)).reshape(m, 512)
    return st


def test_time_step_pieces_256(built):
    """advdiff() (main.c:5027) and projection() (main.c:5828) at 256^3 against the live reference"""
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    if free_ram_gb() < 24:
        pytest.skip("needs >= 24 GB free host RAM")
    import cup3d_b200
    from cup3d_b200 import mesh
    level = 5
    ib, rb = mesh.uniform_blocks(level)
    st = state_input(ib, rb)
    d = scratch()
    try:
        np.save(os.path.join(d, "in_state.npy"), st)
        run_ref(level, d, "advdiff,proj")
        ctx = cup3d_b200.Context(0, 8)
        ctx.mesh_upload(ib, rb, (1, 1, 1), level + 1)
        ctx.set_params(dt=1e-3, nu=1e-3, uinf=(0.1, -0.05, 0.02), step=5, mean_constraint=2, ptol=1e-9, ptol_rel=1e-14)
        out = np.zeros_like(st)
        ctx.state_h2d(st)
        ctx.advdiff()
        ctx.state_d2h(out, 2, 3)
        e_adv = relerr(out[:, 2:5], np.load(os.path.join(d, "out_advdiff.npy")))
        ctx.state_h2d(st)
        info = ctx.projection()
        ctx.state_d2h(out, 1, 4)
        ref = np.load(os.path.join(d, "out_proj.npy"))
        e_p, e_v = relerr(out[:, 1], ref[:, 0]), relerr(out[:, 2:5], ref[:, 1:4])
        ctx.close()
        print("live 256^3: advdiff %.3e, projection p %.3e v %.3e (%d iterations)" % (e_adv, e_p, e_v, info.iterations))
        assert e_adv < 1e-12
        assert e_p < 1e-7 and e_v < 1e-9
    finally:
        shutil.rmtree(d, ignore_errors=True)
