#!/usr/bin/env python3
"""Run the LOGIC of tests/test_gpu_zz_fullsize.py against the CPU reference (oracle/_ref) at a reduced
size: the context of the test module is replaced by a reference-backed stand-in, so every property and
tolerance of that file is exercised without a GPU.  One test per process (the reference keeps its mesh
in file-statics).  TEST INFRASTRUCTURE; needs oracle/_ref (make -C oracle ref).

    python tests/dryrun_fullsize_on_reference.py            # all tests at 64^3 (128^3 for the known answers)
    python tests/dryrun_fullsize_on_reference.py TEST LEVEL  # e.g. test_vcycle_linearity_512 6 = the true
                                                             # size (about 25 GB of host memory, minutes)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/.. = repo root
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]

PLAN = [("test_vcycle_history_128_known_answers", 4), ("test_vcycle_linearity_512", 3),
        ("test_pois_op_nullspace_and_symmetry_512", 3), ("test_vcycle_contraction_512", 3),
        ("test_divp_is_lhs_and_div_grad_adjoint_256", 3), ("test_advdiff_of_constant_flow_256", 3)]


def worker(test, level):
    import numpy as np
    from oracle import refbind as R
    from cup3d_b200 import capi, mesh
    import test_gpu_zz_fullsize as T

    names = {capi.ST_LHS: "lhs", capi.ST_DIVP: "divp", capi.ST_PRHS: "prhs", capi.ST_GRADP: "gradp",
             capi.ST_ADVDIFF: "advdiff"}

    class RefCtx:
        """the subset of cup3d_b200.Context the tests use, served by the reference"""

        def __init__(self):
            R.init(levelStart=level, levelMax=level + 1)
            self.p = dict(dt=0.0, nu=1e-3, uinf=(0, 0, 0), step=0, mean_constraint=2, ptol=1e-6, ptol_rel=1e-4)

        def set_params(self, **kw):
            self.p.update(kw)
            R.set_scalars(**self.p)

        def mg_vcycle(self, x):
            return R.mg_vcycle(x)

        def pois_op(self, x):
            return R.pois_op(x)

        def state_h2d(self, st):
            R.state_set(st)

        def state_d2h(self, out, f0=0, nc=9):
            out[:, f0:f0 + nc] = R.state_get()[:, f0:f0 + nc]

        def stencil_apply(self, sid):
            R.stencil(names[sid])

        def close(self):
            pass

    def ref_ctx(_level, mc=2):
        c = RefCtx()
        ib, rb = R.blocks()
        ib2, rb2 = mesh.uniform_blocks(level)
        assert np.array_equal(ib, ib2) and np.allclose(rb, rb2)
        c.set_params(dt=1e-3, nu=1e-3, uinf=(0.0, 0.0, 0.0), step=5, mean_constraint=mc, ptol=1e-6, ptol_rel=1e-4)
        return c, ib, rb

    T.uniform_ctx = ref_ctx
    fn = getattr(T, test)
    if test == "test_vcycle_history_128_known_answers":
        fn(None)
    elif test.endswith("_512"):
        fn(ref_ctx(level))
    else:
        fn(ref_ctx(level, 0))
    print(test, "holds on the reference at %d^3" % (8 << level))


def main():
    if len(sys.argv) == 3:
        worker(sys.argv[1], int(sys.argv[2]))
        return
    for test, level in PLAN:
        subprocess.run([sys.executable, os.path.abspath(__file__), test, str(level)], check=True,
                       stdout=None, stderr=subprocess.DEVNULL)


if __name__ == "__main__":
    main()
