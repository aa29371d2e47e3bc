"""Shared helpers of the test-suite: golden fixtures and error measures."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)
import make_golden as MG  # noqa: E402  (fields(), state0(), solve_rhs(): the input generators)

AMR_CASES = list(MG.ADAPT)                                  # multi-level meshes (coarse-fine interfaces)
ALL_CASES = [c for c in MG.CASES if c not in AMR_CASES]     # single-level meshes
FULL_CASES = [c for c in ALL_CASES if MG.TIER[c] == "full"]
STENCIL_CASES = [c for c in ALL_CASES if MG.TIER[c] != "big"]
SOLVE_CASES = [c for c in ALL_CASES if c != "b222_l0"]


class Case:
    def __init__(self, name):
        self.name = name
        self.g = np.load(os.path.join(GOLD, name + ".npz"))
        self.ib, self.rb = self.g["ib"], self.g["rb"]
        self.bpd = [int(v) for v in self.g["bpd"]]
        self.level_max = int(self.g["level_max"])
        self.n = len(self.ib)
        self.F = MG.fields(self.ib, self.rb, seed=1234)
        chk = np.array([np.abs(self.F[k]).sum() for k in sorted(self.F)])
        # the regenerated inputs must be the ones the reference was run on
        assert np.allclose(chk, self.g["input_checksum"], rtol=1e-13, atol=0), "input regeneration drifted"
        self.dt, self.nu = float(self.g["scalars"][0]), float(self.g["scalars"][1])
        self.uinf = tuple(float(v) for v in self.g["scalars"][2:5])

    def state0(self):
        return MG.state0(self.F, self.n)

    def solve_rhs(self):
        return MG.solve_rhs(self.F, self.rb)


_cache = {}


def case(name):
    if name not in _cache:
        _cache[name] = Case(name)
    return _cache[name]


def relerr(a, b):
    """max |a-b| / max |b|"""
    a, b = np.asarray(a), np.asarray(b)
    d = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (d if d > 0 else 1.0))
