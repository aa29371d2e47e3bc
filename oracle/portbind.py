"""ctypes binding of oracle/libcup_oracle.so (the C restatement cup_oracle.c).
TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libcup_oracle.so")
_lib = None
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [ip, dp, C.c_longlong, ip, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_nslot.restype = C.c_longlong
        L.orc_nslot.argtypes = [C.c_void_p]
        L.orc_pois_op.argtypes = [C.c_void_p, dp, dp, C.c_int]
        L.orc_mg_vcycle.argtypes = [C.c_void_p, dp, dp]
        L.orc_pois_dot.restype = C.c_double
        L.orc_pois_dot.argtypes = [C.c_void_p, dp, dp]
        L.orc_pre_blk.argtypes = [C.c_void_p, dp, dp, C.c_double]
        L.orc_pois_solve.argtypes = [C.c_void_p, dp, C.c_int, C.c_double, C.c_double, dp]
        L.orc_stencil.argtypes = [C.c_void_p, C.c_int, dp, C.c_double, C.c_double, dp]
        L.orc_advdiff.argtypes = [C.c_void_p, dp, C.c_double, C.c_double, dp]
        L.orc_projection.argtypes = [C.c_void_p, dp, C.c_double, C.c_double, dp, C.c_int, C.c_int, C.c_double,
                                     C.c_double]
        L.orc_projection_keep.argtypes = [C.c_void_p, dp, C.c_double, C.c_double, dp, C.c_int, C.c_int, C.c_double,
                                          C.c_double, C.c_int]
        L.orc_time_vcycle.restype = C.c_double
        L.orc_time_vcycle.argtypes = [C.c_void_p, dp, dp, C.c_int, C.c_int]
        _lib = L
    return _lib


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(dp)


class Oracle:
    def __init__(self, ib, rb, bpd, level_max):
        L = lib()
        ib = np.ascontiguousarray(ib, np.int32)
        hb = np.ascontiguousarray(rb[:, 0], np.float64)
        b = np.asarray(bpd, np.int32)
        self.h = L.orc_create(ib.ctypes.data_as(ip), _p(hb), len(ib), b.ctypes.data_as(ip), int(level_max))
        if not self.h:
            raise RuntimeError("oracle port: mesh not supported (coarse-fine interfaces)")
        self.n = len(ib)

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def pois_op(self, x, mc):
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros_like(x)
        lib().orc_pois_op(self.h, _p(x), _p(y), mc)
        return y

    def mg_vcycle(self, x):
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros_like(x)
        lib().orc_mg_vcycle(self.h, _p(x), _p(y))
        return y

    def pois_dot(self, a, b):
        return float(lib().orc_pois_dot(self.h, _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b))))

    def pre_blk(self, src, invh):
        src = np.ascontiguousarray(src, np.float64)
        dst = np.zeros_like(src)
        lib().orc_pre_blk(self.h, _p(src), _p(dst), invh)
        return dst

    def pois_solve(self, state, mc, ptol, ptol_rel):
        r = C.c_double()
        it = lib().orc_pois_solve(self.h, _p(state), mc, ptol, ptol_rel, C.byref(r))
        return it, r.value

    def stencil(self, sid, state, dt, nu, uinf):
        u = np.asarray(uinf, np.float64)
        assert lib().orc_stencil(self.h, sid, _p(state), dt, nu, _p(u)) == 0

    def advdiff(self, state, dt, nu, uinf):
        u = np.asarray(uinf, np.float64)
        lib().orc_advdiff(self.h, _p(state), dt, nu, _p(u))

    def projection(self, state, dt, nu, uinf, step, mc, ptol, ptol_rel, keep_tmp=False):
        """keep_tmp: F_TMP already holds fish_tmpv()'s udef (it is not zeroed first)"""
        u = np.asarray(uinf, np.float64)
        return lib().orc_projection_keep(self.h, _p(state), dt, nu, _p(u), step, mc, ptol, ptol_rel, int(keep_tmp))

    def time_vcycle(self, x, warmup, n):
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros_like(x)
        return float(lib().orc_time_vcycle(self.h, _p(x), _p(y), warmup, n))


def time_vcycle_uniform(level, warmup, n):
    """bench helper when oracle/_ref is not available"""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from cup3d_b200 import mesh
    ib, rb = mesh.uniform_blocks(level)
    o = Oracle(ib, rb, (1, 1, 1), level + 1)
    b = np.zeros((len(ib), 512))
    b[0, 0], b[-1, 0] = 1.0, -1.0
    sec = o.time_vcycle(b, warmup, n)
    th = int(lib().orc_threads())
    o.close()
    return sec, len(ib), th
