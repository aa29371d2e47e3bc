"""ctypes binding of libcup3d_b200.so -- the C ABI of include/cup3d_b200.h.

This is the reference-side binding a maintainer would write (see
INTEGRATION.md); it contains no numerics.  There is no CPU fallback: if the
shared library is missing or no CUDA device is present every call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcup3d_b200.so")

BS3 = 512
F_CHI, F_PRES, F_VEL, F_TMP, F_LHS, F_N = 0, 1, 2, 5, 8, 9
ST_LHS, ST_MG, ST_ADVDIFF, ST_PRHS, ST_DIVP, ST_GRADP, ST_VORT, ST_Q, ST_GRADCHI = range(9)
M_N = 29  # CUP_M_N: entries of ObstacleBlock.mom (enum M_*, main.c:64-95)


class CupBlk(C.Structure):  # struct Blk, reference main.c:59-63
    _fields_ = [("level", C.c_int), ("ix", C.c_int), ("iy", C.c_int), ("iz", C.c_int), ("Z", C.c_longlong),
                ("h", C.c_double), ("origin", C.c_double * 3)]


class CupParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("nu", C.c_double), ("uinf", C.c_double * 3), ("step", C.c_int),
                ("mean_constraint", C.c_int), ("ptol", C.c_double), ("ptol_rel", C.c_double), ("lam", C.c_double)]


class CupSolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("restarts", C.c_int), ("residual", C.c_double),
                ("rhs_norm", C.c_double), ("vcycles", C.c_int)]


class CupPlan(C.Structure):
    _ip = C.POINTER(C.c_int)
    _fields_ = [("nblk", C.c_longlong), ("nslot", C.c_longlong), ("nact", C.c_int), ("nsend", C.c_int),
                ("nrecv", C.c_int), ("act", _ip), ("ijk", _ip), ("nbr", _ip), ("send_slot", _ip), ("send_plane", _ip),
                ("send_cnt", _ip), ("recv_cnt", _ip), ("pslot", _ip), ("oct", _ip), ("res_send_cnt", _ip),
                ("res_recv_cnt", _ip), ("nres_recv", C.c_int), ("res_recv_slot", _ip), ("res_recv_oct", _ip),
                ("ghosted", C.c_int), ("nghost", C.c_int), ("ext", _ip), ("nbsend", C.c_int), ("nbrecv", C.c_int),
                ("bsend_slot", _ip), ("bsend_kind", _ip), ("bsend_peer", _ip), ("bsend_idx", _ip),
                ("brecv_slot", _ip), ("brecv_kind", _ip)]


# every symbol include/cup3d_b200.h declares: name -> (restype, argtypes)
_vp, _i, _ll, _dp = C.c_void_p, C.c_int, C.c_longlong, C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
SYMBOLS = {
    "cup_last_error": (C.c_char_p, []),
    "cup_version": (_i, []),
    "cup_create": (_i, [C.POINTER(_vp), _i, _i]),
    "cup_destroy": (_i, [_vp]),
    "cup_set_stream": (_i, [_vp, _vp]),
    "cup_set_params": (_i, [_vp, C.POINTER(CupParams)]),
    "cup_synchronize": (_i, [_vp]),
    "cup_mesh_upload": (_i, [_vp, C.POINTER(CupBlk), _ll, C.POINTER(_i), _i]),
    "cup_mesh_adapt": (_i, [_vp, C.POINTER(CupBlk), _ll, C.POINTER(_i), C.POINTER(_ll), C.POINTER(_i), _i]),
    "cup_nblk": (_ll, [_vp]),
    "cup_nslot": (_ll, [_vp]),
    "cup_mg_levels": (_i, [_vp]),
    "cup_mg_nact": (_ll, [_vp, _i]),
    "cup_state_h2d": (_i, [_vp, _vp, _i, _i]),
    "cup_state_d2h": (_i, [_vp, _vp, _i, _i]),
    "cup_state_dev": (_vp, [_vp, _i]),
    "cup_stencil_apply": (_i, [_vp, _i]),
    "cup_stencil_run": (_i, [_vp, _i, C.POINTER(_ll), _ll]),
    "cup_pois_op": (_i, [_vp, _vp, _vp]),
    "cup_mg_vcycle": (_i, [_vp, _vp, _vp]),
    "cup_pois_op_dev": (_i, [_vp, _vp, _vp]),
    "cup_mg_vcycle_dev": (_i, [_vp, _vp, _vp]),
    "cup_pois_dot_dev": (_i, [_vp, _vp, _vp, _dp]),
    "cup_pois_solve": (_i, [_vp, C.POINTER(CupSolveInfo)]),
    "cup_advdiff": (_i, [_vp]),
    "cup_projection": (_i, [_vp, C.POINTER(CupSolveInfo)]),
    "cup_projection_udef_ready": (_i, [_vp, _i]),
    "cup_vorticity": (_i, [_vp]),
    "cup_block_linf": (_i, [_vp, _i, _dp, _dp]),
    "cup_io_pack": (_i, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "cup_obstacle_upload": (_i, [_vp, _i, _i, _ip, _dp, _dp]),
    "cup_obstacle_motion": (_i, [_vp, _i, _dp, _dp, _dp]),
    "cup_obstacle_clear": (_i, [_vp]),
    "cup_obstacle_moments": (_i, [_vp, _i, _dp]),
    "cup_obstacle_penalize": (_i, [_vp]),
    "cup_obstacle_tmpv": (_i, [_vp]),
    "cup_umax": (_i, [_vp, _dp]),
    "cup_comm_init": (_i, [_vp, _i, _i, _vp, C.c_size_t]),
    "cup_nccl_unique_id": (_i, [_vp, C.c_size_t]),
    "cup_comm_init_host": (_i, [_vp, _i, _i, _vp, _vp]),
    "cup_plan_build": (_i, [C.POINTER(CupBlk), _ll, C.POINTER(_i), _i, _i, C.POINTER(_i), _i, _i,
                       C.POINTER(CupPlan)]),
    "cup_plan_free": (None, [C.POINTER(CupPlan)]),
    "cup_kernel_launches": (_ll, [_vp]),
    "cup_trace_report": (_i, [_vp, C.c_char_p, C.c_size_t]),
    "cup_time_smooth": (_i, [_vp, _i, _i, C.POINTER(C.c_float)]),
    "cup_mg_smooth_dev": (_i, [_vp, _i, _i, _vp, _vp]),
    "cup_mg_array": (_vp, [_vp, _i]),
}

ALLGATHER_FN = C.CFUNCTYPE(_i, _vp, _vp, _vp, C.c_size_t)  # CupAllgatherFn

_lib = None


class CupError(RuntimeError):
    pass


def lib():
    """Load libcup3d_b200.so (built by __graft_entry__.build()).  Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CupError("libcup3d_b200.so not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise CupError("cup3d_b200 error %d: %s" % (rc, lib().cup_last_error().decode()))


def blocks_to_struct(ib, rb):
    """(int [n,4] level,ix,iy,iz ; float [n,4] h,origin) -> ctypes array of CupBlk"""
    n = len(ib)
    arr = (CupBlk * n)()
    a = np.frombuffer(arr, dtype=np.dtype([("level", "i4"), ("ix", "i4"), ("iy", "i4"), ("iz", "i4"), ("Z", "i8"),
                                           ("h", "f8"), ("origin", "f8", 3)]))
    a["level"], a["ix"], a["iy"], a["iz"] = ib[:, 0], ib[:, 1], ib[:, 2], ib[:, 3]
    a["Z"] = np.arange(n)
    a["h"] = rb[:, 0]
    a["origin"] = rb[:, 1:4]
    return arr


def plan_build(ib, rb, owner, nranks, rank, bpd, level_max, level):
    """host-only exchange plan of `rank` for one multigrid level -> dict of numpy arrays"""
    arr = blocks_to_struct(np.asarray(ib), np.asarray(rb))
    own = np.ascontiguousarray(owner, np.int32)
    b = (C.c_int * 3)(*bpd)
    p = CupPlan()
    check(lib().cup_plan_build(arr, len(ib), own.ctypes.data_as(C.POINTER(C.c_int)), nranks, rank, b, level_max,
                               level, C.byref(p)))

    def arr_of(ptr, n):
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n > 0 else np.zeros(0, np.int32)

    out = dict(nblk=p.nblk, nslot=p.nslot, nact=p.nact, nsend=p.nsend, nrecv=p.nrecv,
               act=arr_of(p.act, p.nact), ijk=arr_of(p.ijk, 3 * p.nact).reshape(-1, 3), nbr=arr_of(p.nbr, 6 * p.nact).reshape(-1, 6),
               send_slot=arr_of(p.send_slot, p.nsend), send_plane=arr_of(p.send_plane, p.nsend),
               send_cnt=arr_of(p.send_cnt, nranks), recv_cnt=arr_of(p.recv_cnt, nranks),
               pslot=arr_of(p.pslot, p.nact if level > 0 else 0), oct=arr_of(p.oct, p.nact if level > 0 else 0),
               res_send_cnt=arr_of(p.res_send_cnt, nranks), res_recv_cnt=arr_of(p.res_recv_cnt, nranks),
               res_recv_slot=arr_of(p.res_recv_slot, p.nres_recv), res_recv_oct=arr_of(p.res_recv_oct, p.nres_recv),
               ghosted=bool(p.ghosted), nghost=p.nghost, nbsend=p.nbsend, nbrecv=p.nbrecv, ext=arr_of(p.ext, 24 * p.nact).reshape(-1, 6, 4),
               bsend_slot=arr_of(p.bsend_slot, p.nbsend), bsend_kind=arr_of(p.bsend_kind, p.nbsend),
               bsend_peer=arr_of(p.bsend_peer, p.nbsend), bsend_idx=arr_of(p.bsend_idx, p.nbsend),
               brecv_slot=arr_of(p.brecv_slot, p.nbrecv), brecv_kind=arr_of(p.brecv_kind, p.nbrecv))
    lib().cup_plan_free(C.byref(p))
    return out


def split_owner(n, nranks):
    """owner rank of each of n Hilbert-ordered blocks: the reference's contiguous split (mesh_init, main.c:3306-3322)"""
    base, rem = divmod(n, nranks)
    cnt = [base + (1 if r < rem else 0) for r in range(nranks)]
    return np.repeat(np.arange(nranks, dtype=np.int32), cnt)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor (host pinned or device)
        return a.data_ptr()
    return a


def gloo_allgather(group=None):
    """allgather(bytes) -> [bytes] over a torch.distributed (gloo / CPU tensors) process group: a host
    collective for Context.comm_init_host (what MPI_Allgather is for the reference)."""
    import torch
    import torch.distributed as dist

    def ag(payload):
        n = dist.get_world_size(group)
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
        out = [torch.empty_like(t) for _ in range(n)]
        dist.all_gather(out, t, group=group)
        return [bytes(o.numpy().tobytes()) for o in out]

    return ag


def nccl_unique_id():
    buf = (C.c_char * 128)()
    check(lib().cup_nccl_unique_id(buf, 128))
    return bytes(buf)


class Context:
    """One device context == the reference's global `sim`/`sta`/`mg` state for one rank."""

    def __init__(self, device=0, real_bytes=8):
        self.L = lib()
        h = _vp()
        check(self.L.cup_create(C.byref(h), device, real_bytes))
        self.h = h
        self.real_bytes = real_bytes
        self.params = CupParams(dt=0.0, nu=1e-3, uinf=(C.c_double * 3)(0, 0, 0), step=0, mean_constraint=2,
                                ptol=1e-6, ptol_rel=1e-4, lam=1e6)

    def close(self):
        if self.h:
            self.L.cup_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        check(self.L.cup_set_stream(self.h, stream_ptr))

    def set_params(self, **kw):
        for k, v in kw.items():
            if k == "uinf":
                self.params.uinf = (C.c_double * 3)(*v)
            else:
                setattr(self.params, k, v)
        check(self.L.cup_set_params(self.h, C.byref(self.params)))

    def comm_init(self, rank, nranks, id_bytes=None):
        """one rank per GPU; id_bytes = the ncclUniqueId made by nccl_unique_id() on rank 0"""
        buf = (C.c_char * 128).from_buffer_copy(id_bytes) if id_bytes is not None else None
        check(self.L.cup_comm_init(self.h, rank, nranks, buf, 128 if id_bytes is not None else 0))

    def comm_init_host(self, rank, nranks, allgather):
        """cup_comm_init_host: bootstrap by a host collective.  allgather(send: bytes) -> list of nranks
        bytes objects (e.g. built on torch.distributed / mpi4py); the reference passes MPI_Allgather."""
        def cb(user, send, recv, nbytes):
            try:
                parts = allgather(C.string_at(send, nbytes))
                assert len(parts) == nranks and all(len(p) == nbytes for p in parts)
                C.memmove(recv, b"".join(parts), nbytes * nranks)
                return 0
            except Exception as ex:  # reported through cup_last_error as CUP_ERR_COMM
                import sys
                sys.stderr.write("cup3d_b200: allgather callback failed: %r\n" % (ex,))
                return 1
        self._ag_cb = ALLGATHER_FN(cb)  # keep the trampoline alive as long as the context
        check(self.L.cup_comm_init_host(self.h, rank, nranks, C.cast(self._ag_cb, _vp), None))

    def mesh_upload(self, ib, rb, bpd, level_max):
        arr = blocks_to_struct(np.asarray(ib), np.asarray(rb))
        b = (C.c_int * 3)(*bpd)
        check(self.L.cup_mesh_upload(self.h, arr, len(ib), b, level_max))

    def mesh_adapt(self, ib, rb, kind, src, bpd, level_max):
        """cup_mesh_adapt: install the new block list, carrying the fields over on the device (kind 0 keep,
        1 child of the refined old block src, 2 parent of compressed old blocks)"""
        arr = blocks_to_struct(np.asarray(ib), np.asarray(rb))
        k = np.ascontiguousarray(kind, np.int32)
        s = np.ascontiguousarray(src, np.int64)
        b = (C.c_int * 3)(*bpd)
        check(self.L.cup_mesh_adapt(self.h, arr, len(ib), k.ctypes.data_as(C.POINTER(_i)),
                                    s.ctypes.data_as(C.POINTER(_ll)), b, level_max))

    @property
    def nblk(self):
        return int(self.L.cup_nblk(self.h))

    @property
    def nslot(self):
        return int(self.L.cup_nslot(self.h))

    def mg_nact(self, level):
        return int(self.L.cup_mg_nact(self.h, level))

    def mg_levels(self):
        return int(self.L.cup_mg_levels(self.h))

    def state_h2d(self, fld, f0=0, nc=F_N):
        assert fld.dtype == np.float64 and fld.shape == (self.nblk, F_N, BS3) and fld.flags["C_CONTIGUOUS"]
        check(self.L.cup_state_h2d(self.h, _ptr(fld), f0, nc))

    def state_d2h(self, fld, f0=0, nc=F_N):
        assert fld.dtype == np.float64 and fld.shape == (self.nblk, F_N, BS3) and fld.flags["C_CONTIGUOUS"]
        check(self.L.cup_state_d2h(self.h, _ptr(fld), f0, nc))

    def state_dev(self, f):
        return self.L.cup_state_dev(self.h, f)

    def stencil_apply(self, sid):
        check(self.L.cup_stencil_apply(self.h, sid))

    def stencil_run(self, sid, blocks=None, n=None):
        """stencil_run(st, list, n) (main.c:3631): listed local blocks, or the first n when blocks is None"""
        if blocks is None:
            check(self.L.cup_stencil_run(self.h, sid, None, self.nblk if n is None else n))
        else:
            arr = np.ascontiguousarray(blocks, np.int64)
            check(self.L.cup_stencil_run(self.h, sid, arr.ctypes.data_as(C.POINTER(_ll)), len(arr)))

    def pois_op(self, x, out=None):
        out = np.empty_like(x) if out is None else out
        check(self.L.cup_pois_op(self.h, _ptr(x), _ptr(out)))
        return out

    def mg_vcycle(self, x, out=None):
        out = np.empty_like(x) if out is None else out
        check(self.L.cup_mg_vcycle(self.h, _ptr(x), _ptr(out)))
        return out

    def pois_op_dev(self, d_in, d_out):
        check(self.L.cup_pois_op_dev(self.h, _ptr(d_in), _ptr(d_out)))

    def mg_vcycle_dev(self, d_in, d_out):
        check(self.L.cup_mg_vcycle_dev(self.h, _ptr(d_in), _ptr(d_out)))

    def pois_dot_dev(self, a, b):
        r = C.c_double()
        check(self.L.cup_pois_dot_dev(self.h, _ptr(a), _ptr(b), C.byref(r)))
        return r.value

    def pois_solve(self):
        info = CupSolveInfo()
        check(self.L.cup_pois_solve(self.h, C.byref(info)))
        return info

    def advdiff(self):
        check(self.L.cup_advdiff(self.h))

    def projection(self):
        info = CupSolveInfo()
        check(self.L.cup_projection(self.h, C.byref(info)))
        return info

    # ---- obstacle (fish) phases: fish_mom_blk / fish_pen_blk / fish_tmpv on the device
    def obstacle_upload(self, body, blk, chi, udef):
        """blk int32 [nob], chi [nob,512], udef [nob,512,3] (struct ObstacleBlock layout, main.c:96)"""
        blk = np.ascontiguousarray(blk, np.int32)
        chi = np.ascontiguousarray(chi, np.float64)
        udef = np.ascontiguousarray(udef, np.float64)
        n = len(blk)
        assert chi.shape == (n, BS3) and udef.shape == (n, BS3, 3)
        check(self.L.cup_obstacle_upload(self.h, body, n, blk.ctypes.data_as(_ip), chi.ctypes.data_as(_dp),
                                         udef.ctypes.data_as(_dp)))

    def obstacle_motion(self, body, com=None, vel=None, omega=None):
        def p(v):
            return None if v is None else (C.c_double * 3)(*[float(x) for x in v])
        check(self.L.cup_obstacle_motion(self.h, body, p(com), p(vel), p(omega)))

    def obstacle_clear(self):
        check(self.L.cup_obstacle_clear(self.h))

    def obstacle_moments(self, body):
        M = np.zeros(M_N)
        check(self.L.cup_obstacle_moments(self.h, body, M.ctypes.data_as(_dp)))
        return M

    def obstacle_penalize(self):
        check(self.L.cup_obstacle_penalize(self.h))

    def obstacle_tmpv(self):
        check(self.L.cup_obstacle_tmpv(self.h))

    def vorticity(self):
        """vorticity() (main.c:5786): F_VEL -> F_TMP, scaled by 1/h^3"""
        check(self.L.cup_vorticity(self.h))

    def io_pack(self):
        """io_dump's arrays (main.c:1525-1535): float32 chi [n,512], vorticity [n,512,3], Q [n,512]"""
        fp = C.POINTER(C.c_float)
        attr = np.zeros((self.nblk, BS3), np.float32)
        vort = np.zeros((self.nblk, BS3, 3), np.float32)
        q = np.zeros((self.nblk, BS3), np.float32)
        check(self.L.cup_io_pack(self.h, attr.ctypes.data_as(fp), vort.ctypes.data_as(fp), q.ctypes.data_as(fp)))
        return attr, vort, q

    def block_linf(self, f0=F_TMP):
        """-> (linf_all [nblk], linf_fluid [nblk]): mesh_tag_blk's norm, without / with k_gradchi's zeroing"""
        a, f = np.zeros(self.nblk), np.zeros(self.nblk)
        check(self.L.cup_block_linf(self.h, f0, a.ctypes.data_as(_dp), f.ctypes.data_as(_dp)))
        return a, f

    def umax(self):
        r = C.c_double()
        check(self.L.cup_umax(self.h, C.byref(r)))
        return r.value

    def synchronize(self):
        check(self.L.cup_synchronize(self.h))

    def kernel_launches(self):
        return int(self.L.cup_kernel_launches(self.h))

    def trace_report(self):
        """[(phase, ns)] of the last V-cycle (needs CUP_STAMP=1 in the environment)"""
        buf = C.create_string_buffer(1 << 16)
        n = self.L.cup_trace_report(self.h, buf, len(buf))
        return [(a, int(b)) for a, b in (ln.split() for ln in buf.raw[:n].decode().splitlines())]

    def time_smooth(self, level, reps):
        ms = C.c_float()
        check(self.L.cup_time_smooth(self.h, level, reps, C.byref(ms)))
        return ms.value

    def mg_smooth_dev(self, level, n, d_u, d_f):
        check(self.L.cup_mg_smooth_dev(self.h, level, n, _ptr(d_u), _ptr(d_f)))
