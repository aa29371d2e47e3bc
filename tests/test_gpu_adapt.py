"""GPU: the data-parallel half of mesh_adapt on the device -- k_gradchi (main.c:3649), the refinement
interpolation mesh_refine (:3790) and the compression averages (:4129) -- against the reference run LIVE
(oracle/_ref in a subprocess, oracle/live_ref.py) on the reference's own adapted meshes.  The tensorial
labs these kernels read (edges and corners, coarse-fine interpolation) have no stored goldens; the
reference itself is the golden here.  Tolerance 1e-12 relative (same formulas, same operation order;
FMA contraction differs), markers and zeroed cells exactly."""
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
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def live(name, d, ops, *extra):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libcup3d_ref.so")):
        pytest.skip("oracle/_ref not built")
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "live_ref.py"), "--case", name, "--dir", d, "--ops", ops] + \
        [str(e) for e in extra]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]


def blob_state(c, centre, radius, seed=5):
    """chi = a smooth blob whose surface crosses blocks and levels; velocity smooth + noise; TMP random"""
    from cup3d_b200 import mesh
    rng = np.random.default_rng(seed)
    X, Y, Z = mesh.cell_centers(c.ib, c.rb)
    r = np.sqrt((X - centre[0]) ** 2 + (Y - centre[1]) ** 2 + (Z - centre[2]) ** 2)
    st = c.state0()
    st[:, 0] = (1.0 / (1.0 + np.exp((r - radius) / 0.01))).reshape(c.n, 512)
    st[:, 5:8] = rng.standard_normal((c.n, 3, 512))
    return np.ascontiguousarray(st)


@pytest.mark.parametrize("name,centre,radius", [("u32", (0.3, 0.35, 0.4), 0.12), ("amr2", (0.3, 0.35, 0.4), 0.07),
                                                ("amr2", (0.55, 0.5, 0.45), 0.25), ("amr3", (0.3, 0.35, 0.4), 0.1)])
def test_gradchi(built, name, centre, radius):
    import cup3d_b200
    c = case(name)
    st = blob_state(c, centre, radius)
    d = tempfile.mkdtemp(prefix="cup_adapt_")
    try:
        np.save(os.path.join(d, "in_state.npy"), st)
        live(name, d, "gradchi")
        assert np.array_equal(np.load(os.path.join(d, "ib.npy")), c.ib)
        ref = np.load(os.path.join(d, "out_gradchi.npy"))
        ctx = cup3d_b200.Context(0, 8)
        ctx.mesh_upload(c.ib, c.rb, c.bpd, c.level_max)
        ctx.state_h2d(st)
        ctx.stencil_apply(capi.ST_GRADCHI)
        out = np.zeros_like(st)
        ctx.state_d2h(out)
        ctx.close()
        marked_ref = np.any(ref[:, 0] == 1e10, axis=1)
        marked = np.any(out[:, 5] == 1e10, axis=1)
        assert marked_ref.any() and not marked_ref.all()
        assert np.array_equal(marked, marked_ref)
        assert np.array_equal(out[:, 5:8], ref)          # zeroed cells and markers exactly
        assert np.array_equal(out[:, :5], st[:, :5]) and np.array_equal(out[:, 8], st[:, 8])
    finally:
        shutil.rmtree(d, ignore_errors=True)


def plan_from_lists(ib_old, ib_new):
    """kind / src of cup_mesh_adapt from the two block lists (what the host knows after its tree surgery)"""
    old = {tuple(b): i for i, b in enumerate(ib_old.tolist())}
    kind, src = [], []
    for b in ib_new.tolist():
        L, x, y, z = b
        if tuple(b) in old:
            kind.append(0)
            src.append(old[tuple(b)])
        elif (L - 1, x // 2, y // 2, z // 2) in old:
            kind.append(1)
            src.append(old[(L - 1, x // 2, y // 2, z // 2)])
        else:
            kind.append(2)
            src.append(old[(L + 1, 2 * x, 2 * y, 2 * z)])
    return np.array(kind, np.int32), np.array(src, np.int64)


@pytest.mark.parametrize("name,centre,radius,ctol", [("amr2", (0.6, 0.55, 0.5), 0.08, 1e4), ("amr3", (0.62, 0.6, 0.55), 0.1, 1e4),
                                                     ("amr2", (0.3, 0.35, 0.4), 0.12, -1.0)])
def test_mesh_adapt_fields(built, name, centre, radius, ctol):
    """refine where the new blob's surface is, compress what the old one left behind; all nine fields of
    the new mesh against the reference's mesh_adapt"""
    import cup3d_b200
    c = case(name)
    st = blob_state(c, centre, radius)
    d = tempfile.mkdtemp(prefix="cup_adapt_")
    try:
        np.save(os.path.join(d, "in_state.npy"), st)
        live(name, d, "adapt", "--rtol", 1e9, "--ctol", ctol)
        mid = np.load(os.path.join(d, "adapt_mid.npy"))
        ib2, rb2 = np.load(os.path.join(d, "adapt_ib.npy")), np.load(os.path.join(d, "adapt_rb.npy"))
        ref = np.load(os.path.join(d, "adapt_state.npy"))
        kind, src = plan_from_lists(c.ib, ib2)
        assert (kind == 1).any()
        if ctol > 0:
            assert (kind == 2).any()
        ctx = cup3d_b200.Context(0, 8)
        ctx.mesh_upload(c.ib, c.rb, c.bpd, c.level_max)
        # the device's own vorticity + gradchi give the state the transfer reads (== the reference's `mid`)
        ctx.state_h2d(st)
        ctx.vorticity()
        ctx.stencil_apply(capi.ST_GRADCHI)
        got_mid = np.zeros_like(st)
        ctx.state_d2h(got_mid)
        assert relerr(got_mid[:, 5:8], mid[:, 5:8]) < 1e-12
        ctx.state_h2d(np.ascontiguousarray(mid))
        ctx.mesh_adapt(ib2, rb2, kind, src, c.bpd, c.level_max)
        assert ctx.nblk == len(ib2)
        out = np.zeros((len(ib2), 9, 512))
        ctx.state_d2h(out)
        ctx.close()
        for f in range(9):
            scale = max(np.max(np.abs(ref[:, f])), 1e-300)
            for k in (0, 1, 2):
                sel = kind == k
                if not sel.any():
                    continue
                if k == 0:
                    assert np.array_equal(out[sel, f], ref[sel, f]), f  # kept blocks: copied
                else:
                    e = np.max(np.abs(out[sel, f] - ref[sel, f])) / scale
                    assert e < 1e-12, (f, k, e)
    finally:
        shutil.rmtree(d, ignore_errors=True)
