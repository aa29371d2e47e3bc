"""CPU: the integration recipe of INTEGRATION.md, executed.  tests/coupled.py runs the reference's time
loop twice -- once as advance() orders it, once with the hot path behind the backend interface the
library implements (here served by the reference's own functions, called piecewise: fish_vel split into
moments + fish_solve, fish_pen into fish_hit + block loop, fields handed over around the host phases) --
and both must produce the same flow and the same fish motion.  The GPU backend of the same script is
exercised by tests/test_gpu_zzz_coupled.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def run(mode, case, nsteps, tmp_path):
    out = os.path.join(str(tmp_path), "%s_%s.npz" % (mode, case))
    r = subprocess.run([sys.executable, os.path.join(HERE, "coupled.py"), mode, case, str(nsteps), out],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out)


def compare(a, b, rtol):
    assert np.array_equal(a["nblk"], b["nblk"])
    for k in ("dt", "umax", "checks", "motion"):
        scale = np.max(np.abs(a[k]))
        assert np.max(np.abs(a[k] - b[k])) <= rtol * scale, (k, a[k], b[k])


@pytest.mark.parametrize("case,nsteps", [("fish64", 4), ("fishamr", 3)])
def test_piecewise_time_loop_equals_advance(built, tmp_path, case, nsteps):
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    pure = run("pure", case, nsteps, tmp_path)
    coup = run("ref", case, nsteps, tmp_path)
    # same functions, same data; only OpenMP reduction order differs between two runs
    compare(pure, coup, 1e-10)
    assert pure["umax"][-1] > 0 and np.any(np.abs(pure["motion"]) > 0)     # the fish does move the fluid


def test_time_loop_with_an_independent_backend(built, tmp_path):
    """backend "port": separate memory, separate implementation (C + numpy restatements) that only knows what
    the orchestration uploads -- the same position a GPU is in.  A missing hand-over (F_CHI after fish_build,
    lambda, uinf, the bodies' com / vel / omega, udef for the pressure right-hand side ...) shows up here."""
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    pure = run("pure", "fish64", 4, tmp_path)
    port = run("port", "fish64", 4, tmp_path)
    compare(pure, port, 1e-7)      # both solve the pressure equation to 1e-10 / 1e-12
