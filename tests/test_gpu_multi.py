"""GPU, 2 ranks: the domain-decomposed V-cycle / operator / solve / time-step pieces must reproduce
the reference's single-rank golden result -- decomposition may only change reduction order
(SURVEY.md 8c).

Two bootstraps are covered.  "nccl": cup_comm_init with an ncclUniqueId, one rank per GPU (needs
>= 2 devices, skipped otherwise).  "host": cup_comm_init_host with a host all-gather (gloo here,
MPI_Allgather in the reference); data moves through CUDA-IPC peer windows only, so the ranks may
share ONE device -- that is how the exchange kernels run in a single-GPU CI box (the two processes
are time-sliced by the driver; every flag wait costs a time slice, which is fine for these sizes)."""
import os
import sys

import numpy as np
import pytest

from util import case, relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def connect(rank, world, boot, real_bytes=8):
    """-> Context of this rank on its device, communicator initialised.  boot = ("nccl", id) or ("host", port)"""
    import torch
    import cup3d_b200
    from cup3d_b200 import capi
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    ctx = cup3d_b200.Context(dev, real_bytes)
    if boot[0] == "nccl":
        ctx.comm_init(rank, world, boot[1])
    else:
        import torch.distributed as dist
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % boot[1], rank=rank, world_size=world)
        ctx.comm_init_host(rank, world, capi.gloo_allgather())
    return ctx


def disconnect(ctx, boot):
    ctx.close()
    if boot[0] == "host":
        import torch.distributed as dist
        dist.destroy_process_group()


def boots():
    """the bootstraps this box can run: host always (ranks may share a device), nccl with >= 2 GPUs"""
    import torch
    return ["host"] + (["nccl"] if torch.cuda.device_count() >= 2 else [])


def make_boot(kind):
    if kind == "nccl":
        from cup3d_b200 import capi
        return ("nccl", capi.nccl_unique_id())
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return ("host", port)


def run_ranks(target, world, args, timeout=600):
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    ps = [ctxm.Process(target=target, args=(r, world) + tuple(args) + (q,)) for r in range(world)]
    for p in ps:
        p.start()
    try:
        got = [q.get(timeout=timeout) for _ in ps]
    finally:
        for p in ps:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    for p in ps:
        assert p.exitcode == 0
    return got


def worker(rank, world, boot, name, coarse, q):
    os.environ["CUP_COARSE_BLOCKS"] = str(coarse)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cup3d_b200 import capi
    from util import case
    c = case(name)
    owner = capi.split_owner(c.n, world)
    mine = np.nonzero(owner == rank)[0]
    ctx = connect(rank, world, boot)
    ctx.mesh_upload(c.ib[mine], c.rb[mine], c.bpd, c.level_max)
    ctx.set_params(dt=c.dt, nu=c.nu, uinf=c.uinf, step=5, mean_constraint=2, ptol=1e-10, ptol_rel=1e-12)
    out = {}
    out["vc"] = ctx.mg_vcycle(np.ascontiguousarray(c.F["cosrhs"][mine]))
    out["op"] = ctx.pois_op(np.ascontiguousarray(c.F["pres"][mine]))
    st = c.state0()[mine]
    st[:, 8] = c.solve_rhs()[mine]
    st[:, 1] = 0
    st = np.ascontiguousarray(st)
    ctx.state_h2d(st)
    info = ctx.pois_solve()
    res = np.zeros_like(st)
    ctx.state_d2h(res, 1, 1)
    out["x"] = res[:, 1]
    out["it"] = info.iterations
    out["res"] = info.residual
    q.put((rank, mine, out))
    disconnect(ctx, boot)


@pytest.mark.parametrize("kind", ["host", "nccl"])
@pytest.mark.parametrize("name,coarse,world", [("u32", 0, 2), ("u64", 0, 2), ("u64", 4096, 2), ("u64", 64, 3)])
def test_ranks_match_single_rank_reference(built, name, coarse, world, kind):
    if kind not in boots():
        pytest.skip("the NCCL bootstrap needs one GPU per rank")
    import torch
    if kind == "nccl" and torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    c = case(name)
    got = run_ranks(worker, world, (make_boot(kind), name, coarse))
    vc = np.zeros((c.n, 512))
    op = np.zeros((c.n, 512))
    x = np.zeros((c.n, 512))
    for rank, mine, out in got:
        vc[mine], op[mine], x[mine] = out["vc"], out["op"], out["x"]
        assert out["res"] < 1e-10
    assert relerr(vc, c.g["vc_out_cosrhs"]) < 1e-11
    assert relerr(op, c.g["op_out_mc2"]) < 1e-12
    assert relerr(x, c.g["solve_x_mc2"]) < 1e-8


def worker_step(rank, world, boot, name, q):
    os.environ["CUP_COARSE_BLOCKS"] = "0"
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cup3d_b200 import capi
    from util import case
    c = case(name)
    owner = capi.split_owner(c.n, world)
    mine = np.nonzero(owner == rank)[0]
    ctx = connect(rank, world, boot)
    ctx.mesh_upload(c.ib[mine], c.rb[mine], c.bpd, c.level_max)
    ctx.set_params(dt=c.dt, nu=c.nu, uinf=c.uinf, step=5, mean_constraint=2, ptol=1e-10, ptol_rel=1e-12)
    out = {}
    s0 = np.ascontiguousarray(c.state0()[mine])
    for st, sid in (("advdiff", capi.ST_ADVDIFF), ("prhs", capi.ST_PRHS), ("divp", capi.ST_DIVP),
                    ("gradp", capi.ST_GRADP), ("vort", capi.ST_VORT), ("q", capi.ST_Q)):
        ctx.state_h2d(s0)
        ctx.stencil_apply(sid)
        r = np.zeros_like(s0)
        ctx.state_d2h(r)
        out["st_" + st] = r
    ctx.state_h2d(s0)
    ctx.advdiff()
    r = np.zeros_like(s0)
    ctx.state_d2h(r)
    out["advdiff"] = r
    ctx.state_h2d(s0)
    ctx.projection()
    r = np.zeros_like(s0)
    ctx.state_d2h(r)
    out["proj"] = r
    q.put((rank, mine, out))
    disconnect(ctx, boot)


@pytest.mark.parametrize("kind", ["host", "nccl"])
@pytest.mark.parametrize("name", ["u16", "b211"])
def test_two_rank_time_step_pieces(built, name, kind):
    """advdiff(), projection() and every stencil sweep, domain-decomposed over 2 ranks"""
    if kind not in boots():
        pytest.skip("the NCCL bootstrap needs one GPU per rank")
    c = case(name)
    got = run_ranks(worker_step, 2, (make_boot(kind), name))
    full = {}
    for rank, mine, out in got:
        for k, v in out.items():
            full.setdefault(k, np.zeros((c.n, 9, 512)))[mine] = v
    assert relerr(full["st_advdiff"][:, 5:8], c.g["st_advdiff"]) < 1e-12
    assert relerr(full["st_prhs"][:, 8:9], c.g["st_prhs"]) < 1e-12
    assert relerr(full["st_divp"][:, 5:6], c.g["st_divp"]) < 1e-12
    assert relerr(full["st_gradp"][:, 5:8], c.g["st_gradp"]) < 1e-12
    assert relerr(full["st_vort"][:, 5:8], c.g["st_vort"]) < 1e-12
    assert relerr(full["st_q"][:, 8:9], c.g["st_q"]) < 1e-12
    assert relerr(full["advdiff"][:, 2:5], c.g["advdiff"][:, 0:3]) < 1e-12
    assert relerr(full["proj"][:, 1], c.g["proj_step5"][:, 0]) < 1e-7
    assert relerr(full["proj"][:, 2:5], c.g["proj_step5"][:, 1:4]) < 1e-9


def worker_amr(rank, world, boot, name, q):
    """everything test_gpu_amr.py checks on one rank, on a multi-level mesh split over `world` ranks"""
    os.environ["CUP_COARSE_BLOCKS"] = "0"
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cup3d_b200 import capi
    from util import case
    c = case(name)
    owner = capi.split_owner(c.n, world)
    mine = np.nonzero(owner == rank)[0]
    ctx = connect(rank, world, boot)
    ctx.mesh_upload(c.ib[mine], c.rb[mine], c.bpd, c.level_max)
    ctx.set_params(dt=c.dt, nu=c.nu, uinf=c.uinf, step=5, mean_constraint=2, ptol=1e-10, ptol_rel=1e-12)
    out = {}
    out["vc"] = ctx.mg_vcycle(np.ascontiguousarray(c.F["cosrhs"][mine]))
    out["vc2"] = ctx.mg_vcycle(np.ascontiguousarray(c.F["cosrhs"][mine]))  # CUDA-graph replay
    out["op"] = ctx.pois_op(np.ascontiguousarray(c.F["pres"][mine]))
    s0 = np.ascontiguousarray(c.state0()[mine])
    if name == "amr2":
        for st, sid in (("lhs", capi.ST_LHS), ("advdiff", capi.ST_ADVDIFF), ("prhs", capi.ST_PRHS),
                        ("divp", capi.ST_DIVP), ("gradp", capi.ST_GRADP), ("vort", capi.ST_VORT), ("q", capi.ST_Q)):
            ctx.state_h2d(s0)
            ctx.stencil_apply(sid)
            r = np.zeros_like(s0)
            ctx.state_d2h(r)
            out["st_" + st] = r
        ctx.state_h2d(s0)
        ctx.advdiff()
        r = np.zeros_like(s0)
        ctx.state_d2h(r)
        out["advdiff"] = r
        ctx.state_h2d(s0)
        info = ctx.projection()
        r = np.zeros_like(s0)
        ctx.state_d2h(r)
        out["proj"] = r
        out["proj_res"] = (info.residual, info.rhs_norm)
    q.put((rank, mine, out))
    disconnect(ctx, boot)


@pytest.mark.parametrize("kind", ["host", "nccl"])
@pytest.mark.parametrize("name,world", [("amr2", 2), ("amr2", 3), ("amr3", 2), ("amr3", 4)])
def test_multi_level_mesh_across_ranks(built, name, world, kind):
    """coarse-fine interfaces ACROSS ranks (halo_sync of whole blocks main.c:3101-3112, fc_fill :3228-3293,
    remote parents on AMR levels :4734-4807): same goldens and tolerances as the single-rank AMR tests"""
    if kind not in boots():
        pytest.skip("the NCCL bootstrap needs one GPU per rank")
    import torch
    if kind == "nccl" and torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    c = case(name)
    got = run_ranks(worker_amr, world, (make_boot(kind), name))
    full = {}
    for rank, mine, out in got:
        for k, v in out.items():
            if k == "proj_res":
                assert v[0] < max(1e-10, 1e-12 * v[1])
                continue
            full.setdefault(k, np.zeros((c.n,) + v.shape[1:]))[mine] = v
    assert relerr(full["vc"], c.g["vc_out_cosrhs"]) < 1e-11
    assert relerr(full["vc2"], c.g["vc_out_cosrhs"]) < 1e-11
    assert relerr(full["op"], c.g["op_out_mc2"]) < 1e-12
    if name != "amr2":
        return
    for st, f0, nc in (("lhs", 8, 1), ("advdiff", 5, 3), ("prhs", 8, 1), ("divp", 5, 1), ("gradp", 5, 3),
                       ("vort", 5, 3), ("q", 8, 1)):
        assert relerr(full["st_" + st][:, f0:f0 + nc], c.g["st_" + st]) < 1e-12, st
    assert relerr(full["advdiff"][:, 2:5], c.g["advdiff"][:, 0:3]) < 1e-12
    assert relerr(full["proj"][:, 1], c.g["proj_step5"][:, 0]) < 1e-7
    assert relerr(full["proj"][:, 2:5], c.g["proj_step5"][:, 1:4]) < 1e-9


def worker_sphere(rank, world, boot, rbytes, coarse, q):
    """a synthetic multi-level mesh (spherical-shell refinement, the bench's AMR mesh in small) split over
    `world` ranks (world == 1: the single-rank run the others are compared with)"""
    if coarse is not None:
        os.environ["CUP_COARSE_BLOCKS"] = str(coarse)
    sys.path.insert(0, ROOT)
    import torch
    import cup3d_b200
    from cup3d_b200 import capi, mesh
    gib, grb = mesh.amr_blocks(0, 2, mesh.sphere_shell((0.5, 0.5, 0.5), 0.2, 0.5), bpd=(8, 8, 8))
    owner = capi.split_owner(len(gib), world)
    mine = np.nonzero(owner == rank)[0]
    ctx = connect(rank, world, boot, rbytes) if world > 1 else cup3d_b200.Context(0, rbytes)
    ctx.mesh_upload(gib[mine], grb[mine], (8, 8, 8), 3)
    X, Y, Z = mesh.cell_centers(gib[mine], grb[mine])
    n = len(mine)
    h = grb[mine][:, 0][:, None, None, None]
    rhs = (h ** 3 * (np.cos(np.pi * X) * np.cos(2 * np.pi * Y) * np.cos(3 * np.pi * Z) +
                     0.1 * np.sin(9 * np.pi * X) * np.sin(7 * np.pi * Y))).reshape(n, 512)
    st = np.zeros((n, 9, 512))
    st[:, 2] = (np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y)).reshape(n, 512)
    st[:, 3] = (-np.cos(2 * np.pi * X) * np.sin(2 * np.pi * Y)).reshape(n, 512)
    st[:, 4] = (0.1 * np.sin(2 * np.pi * Z)).reshape(n, 512)
    tol = dict(ptol=1e-9, ptol_rel=1e-10) if rbytes == 8 else dict(ptol=1e-4, ptol_rel=1e-3)
    ctx.set_params(dt=2e-4, nu=1e-3, uinf=(0.0, 0.0, 0.0), step=5, mean_constraint=2, **tol)
    out = {}
    out["vc"] = ctx.mg_vcycle(np.ascontiguousarray(rhs))
    out["vc2"] = ctx.mg_vcycle(np.ascontiguousarray(rhs))
    out["op"] = ctx.pois_op(np.ascontiguousarray(rhs))
    st = np.ascontiguousarray(st)
    its = []
    ctx.state_h2d(st)
    for _ in range(3):
        ctx.advdiff()
        its.append(ctx.projection().iterations)
    r = np.zeros_like(st)
    ctx.state_d2h(r)
    out["vel"] = r[:, 2:5]
    out["its"] = np.array(its)
    q.put((rank, mine, out))
    if world > 1:
        disconnect(ctx, boot)
    else:
        ctx.close()


@pytest.mark.parametrize("rbytes,world,coarse", [(8, 3, None), (8, 4, 0), (4, 4, None)])
def test_sphere_mesh_ranks_agree_with_one_rank(built, rbytes, world, coarse):
    """the multi-rank run (default coarse-level threshold and all-levels-distributed) reproduces the
    single-rank run on a 3-level, ~4000-block mesh: V-cycle (eager and replayed), operator, three full time
    steps and their Krylov iteration counts -- fp64 to rounding, fp32 to single-precision rounding"""
    one = run_ranks(worker_sphere, 1, (("host", 0), rbytes, coarse))
    got = run_ranks(worker_sphere, world, (make_boot("host"), rbytes, coarse))
    ref = one[0][2]
    ntot = len(one[0][1])
    full = {}
    for rank, mine, out in got:
        for k, v in out.items():
            if k == "its":
                assert np.array_equal(v, ref["its"]) or rbytes == 4, (v, ref["its"])
                continue
            full.setdefault(k, np.zeros((ntot,) + v.shape[1:]))[mine] = v
    t_vc, t_op, t_vel = (1e-11, 1e-12, 1e-8) if rbytes == 8 else (2e-4, 2e-5, 2e-3)
    assert relerr(full["vc"], ref["vc"]) < t_vc
    assert relerr(full["vc2"], ref["vc"]) < t_vc
    assert relerr(full["op"], ref["op"]) < t_op
    assert relerr(full["vel"], ref["vel"]) < t_vel
