"""GPU: the reference's time loop with its hot path on the device (tests/coupled.py, backend "gpu") against
the reference alone -- two fish on a uniform 64^3 mesh, four steps from rest: time-step sizes, maximum
velocity, the fish's rigid motion (centre of mass, velocity, angular velocity from the reference's own
fish_solve fed with DEVICE moments) and global sums of pressure and velocity.  Both sides solve the
pressure equation to 1e-10 / 1e-12, so they agree far below 1e-6.  The orchestration itself is checked on
the CPU (tests/test_coupled_ref.py).  Needs oracle/_ref (the host phases ARE the reference)."""
import os

import numpy as np
import pytest

from test_coupled_ref import run

pytestmark = pytest.mark.gpu

CASES = [("fish64", 4)]
if os.environ.get("CUP_COUPLED_AMR"):      # mesh adaptation between steps: tagging could flip on round-off
    CASES.append(("fishamr", 3))


@pytest.mark.parametrize("case,nsteps", CASES)
def test_time_loop_with_device_hot_path(built, tmp_path, case, nsteps):
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built")
    pure = run("pure", case, nsteps, tmp_path)
    dev = run("gpu", case, nsteps, tmp_path)
    assert np.array_equal(pure["nblk"], dev["nblk"])
    for k in ("dt", "umax", "checks", "motion"):
        scale = np.max(np.abs(pure[k]))
        assert np.max(np.abs(pure[k] - dev[k])) <= 1e-6 * scale, (k, pure[k], dev[k])
