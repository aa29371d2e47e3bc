"""ctypes binding of oracle/_ref/libcup3d_ref.so (the unmodified reference,
/root/reference/main.c, built by oracle/Makefile).  TEST INFRASTRUCTURE ONLY.

The reference keeps all state in file-statics, so one process can hold exactly
one mesh: call :func:`init` once per process (tests/golden/make_golden.py runs
each case in a subprocess).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")
LIB = os.path.join(REFDIR, "libcup3d_ref.so")
LIB32 = os.path.join(REFDIR, "libcup3d_ref32.so")  # the same source built with Real = float
BS3 = 512
F_N = 9
F_CHI, F_PRES, F_VEL, F_TMP, F_LHS = 0, 1, 2, 5, 8
STENCILS = {"lhs": 0, "mg": 1, "advdiff": 2, "prhs": 3, "divp": 4, "gradp": 5, "vort": 6, "q": 7, "gradchi": 8}

_lib = None
dp = C.POINTER(C.c_double)


def available():
    return os.path.exists(LIB)


def default_args(**kw):
    """The reference's mandatory '-key value' list (main.c:165-191); values as in run.sh."""
    a = {
        "bMeanConstraint": 2, "bpdx": 1, "bpdy": 1, "bpdz": 1, "CFL": 0.4, "Ctol": 0.1, "dt": 0,
        "extent": 1, "factory-content": "", "lambda": 1e6, "levelMax": 4, "levelStart": 3, "nsteps": 0,
        "nu": 0.001, "poissonTol": 1e-6, "poissonTolRel": 1e-4, "rampup": 100, "Rtol": 5,
        "StaticObstacles": 0, "tdump": 0, "tend": 0, "uinfx": 0, "uinfy": 0, "uinfz": 0, "umax": 10,
        "use-dlm": 0,
    }
    a.update(kw)
    return a


def init(real_bytes=8, **kw):
    global _lib
    if _lib is not None:
        raise RuntimeError("reference already initialised in this process")
    lib = C.CDLL(LIB if real_bytes == 8 else LIB32)
    assert lib.ref_real_bytes() == real_bytes
    args = default_args(**kw)
    flat = []
    for k, v in args.items():
        flat += ["-" + k, repr(v) if isinstance(v, float) else str(v)]
    arr = (C.c_char_p * len(flat))(*[s.encode() for s in flat])
    lib.ref_init.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_char_p]
    rc = lib.ref_init(len(flat), arr, REFDIR.encode())
    if rc != 0:
        raise RuntimeError("ref_init failed: %d" % rc)
    lib.ref_nblk.restype = C.c_longlong
    lib.ref_pois_dot.restype = C.c_double
    lib.ref_umax.restype = C.c_double
    lib.ref_time_vcycle.restype = C.c_double
    lib.ref_time_stencil.restype = C.c_double
    lib.ref_set_scalars.argtypes = [C.c_double] * 5 + [C.c_int, C.c_int, C.c_double, C.c_double]
    lib.ref_mesh_adapt.argtypes = [C.c_double, C.c_double]
    lib.ref_pre_blk.argtypes = [dp, dp, C.c_double]
    _lib = lib
    return lib


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(dp)


def nblk():
    return int(_lib.ref_nblk())


def threads():
    return int(_lib.ref_threads())


def blocks():
    """-> (int32 [n,4] level,ix,iy,iz ; float64 [n,4] h,origin)"""
    n = nblk()
    ib = np.zeros((n, 4), np.int32)
    rb = np.zeros((n, 4), np.float64)
    _lib.ref_blocks(ib.ctypes.data_as(C.POINTER(C.c_int)), _p(rb))
    return ib, rb


def state_get():
    s = np.zeros((nblk(), F_N, BS3), np.float64)
    _lib.ref_state_get(_p(s))
    return s


def state_set(s):
    s = np.ascontiguousarray(s, np.float64)
    assert s.shape == (nblk(), F_N, BS3)
    _lib.ref_state_set(_p(s))


def set_scalars(dt=0.0, nu=0.001, uinf=(0.0, 0.0, 0.0), step=0, mean_constraint=2, ptol=1e-6, ptol_rel=1e-4):
    _lib.ref_set_scalars(dt, nu, uinf[0], uinf[1], uinf[2], step, mean_constraint, ptol, ptol_rel)


def mg_vcycle(x):
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros_like(x)
    _lib.ref_mg_vcycle(_p(x), _p(y))
    return y


def pois_op(x):
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros_like(x)
    _lib.ref_pois_op(_p(x), _p(y))
    return y


def pois_dot(a, b):
    return float(_lib.ref_pois_dot(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b))))


def pre_blk(src, invh):
    src = np.ascontiguousarray(src, np.float64)
    dst = np.zeros_like(src)
    _lib.ref_pre_blk(_p(src), _p(dst), invh)
    return dst


def pois_solve():
    _lib.ref_pois_solve()


def advdiff():
    _lib.ref_advdiff()


def projection():
    _lib.ref_projection()


def vorticity():
    _lib.ref_vorticity()


def umax():
    return float(_lib.ref_umax())


def stencil(name):
    assert _lib.ref_stencil(STENCILS[name]) == 0


def mesh_adapt(rtol, ctol):
    _lib.ref_mesh_adapt(rtol, ctol)


def time_vcycle(x, warmup, n):
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros_like(x)
    return float(_lib.ref_time_vcycle(_p(x), _p(y), warmup, n))


def time_stencil(name, warmup, n):
    return float(_lib.ref_time_stencil(STENCILS[name], warmup, n))


# ---- obstacle (fish) phases, SURVEY 8(f) row 1 ---------------------------------------
PHASES = {"fish_build": 0, "advdiff": 1, "fish_vel": 2, "fish_pen": 3, "projection": 4, "fish_tmpv": 5,
          "mesh_adapt": 6}


def sta_fields():
    _lib.ref_sta_fields()


def sta_dt():
    _lib.ref_sta_dt.restype = C.c_double
    return float(_lib.ref_sta_dt())


def phase(name):
    assert _lib.ref_phase(PHASES[name]) == 0


def step_end():
    _lib.ref_step_end()


def get_scalars():
    o = np.zeros(7)
    _lib.ref_get_scalars(_p(o))
    return dict(dt=o[0], lam=o[1], uinf=tuple(o[2:5]), step=int(o[5]), time=o[6])


def set_lambda(lam):
    _lib.ref_set_lambda.argtypes = [C.c_double]
    _lib.ref_set_lambda(lam)


def nfish():
    return int(_lib.ref_nfish())


def fish_obstacle(k):
    """-> (blk int32 [nob], chi [nob,512], udef [nob,512,3]) of fish k (struct ObstacleBlock, main.c:96)"""
    n = int(_lib.ref_fish_nob(k))
    blk = np.zeros(n, np.int32)
    chi = np.zeros((n, BS3))
    udef = np.zeros((n, BS3, 3))
    _lib.ref_fish_ob(k, blk.ctypes.data_as(C.POINTER(C.c_int)), _p(chi), _p(udef))
    return blk, chi, udef


def fish_motion(k):
    o = np.zeros(9)
    _lib.ref_fish_motion(k, _p(o))
    return o[0:3].copy(), o[3:6].copy(), o[6:9].copy()


def fish_mom(k):
    M = np.zeros(int(_lib.ref_m_n()))
    _lib.ref_fish_mom(k, _p(M))
    return M


def fish_solve_from(k, M):
    """the tail of fish_vel for fish k: moments -> fish_solve"""
    M = np.ascontiguousarray(M, np.float64)
    _lib.ref_fish_solve_from(k, _p(M))


def fish_hit():
    _lib.ref_fish_hit()


def fish_pen_blocks():
    _lib.ref_fish_pen_blocks()
