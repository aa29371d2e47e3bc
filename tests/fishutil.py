"""Loader of the obstacle (fish) fixtures (tests/golden/make_golden_fish.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)
import make_golden as MG  # noqa: E402
import make_golden_fish as MGF  # noqa: E402

FISH_CASES = list(MGF.CASES)
checksums = MGF.checksums


class FishCase:
    def __init__(self, name):
        g = np.load(os.path.join(GOLD, name + ".npz"))
        self.g = g
        self.ib, self.rb = g["ib"], g["rb"]
        self.bpd = [int(v) for v in g["bpd"]]
        self.level_max = int(g["level_max"])
        self.n = len(self.ib)
        self.nfish = int(g["nfish"])
        s = g["scalars"]
        self.dt, self.nu, self.uinf, self.step, self.lam = float(s[0]), float(s[1]), tuple(s[2:5]), int(s[5]), float(s[6])
        self.F = MG.fields(self.ib, self.rb, seed=4321)
        chk = np.array([np.abs(self.F[k]).sum() for k in ("pres", "vel")])
        assert np.allclose(chk, g["input_checksum"], rtol=1e-13, atol=0), "input regeneration drifted"
        self.obu = g["obu"]
        self.obs = [(g["ob%d_blk" % k], g["ob%d_chi" % k], g["ob%d_udef" % k]) for k in range(self.nfish)]
        self.com = [g["com%d" % k] for k in range(self.nfish)]
        self.vel = [g["vel%d" % k] for k in range(self.nfish)]
        self.omega = [g["omega%d" % k] for k in range(self.nfish)]
        self.mom = [g["mom%d" % k] for k in range(self.nfish)]
        self.chi_field = np.zeros((self.n, 512))
        self.chi_field[self.obu] = g["chi_field"]

    def state0(self):
        """sta.fld right after fish_build of the captured step: CHI from the bodies, seeded PRES / VEL"""
        st = np.zeros((self.n, 9, 512))
        st[:, 0] = self.chi_field
        st[:, 1] = self.F["pres"]
        st[:, 2:5] = self.F["vel"]
        return st


_cache = {}


def fish_case(name):
    if name not in _cache:
        _cache[name] = FishCase(name)
    return _cache[name]
