"""GPU, 2 ranks (needs >= 2 devices; skipped otherwise): the domain-decomposed
V-cycle / operator / solve must reproduce the reference's single-rank golden
result -- decomposition may only change reduction order."""
import os
import sys

import numpy as np
import pytest

from util import case, relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(rank, world, nid, name, coarse, q):
    os.environ["CUP_COARSE_BLOCKS"] = str(coarse)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    torch.cuda.set_device(rank)
    import cup3d_b200
    from cup3d_b200 import capi
    from util import case
    c = case(name)
    owner = capi.split_owner(c.n, world)
    mine = np.nonzero(owner == rank)[0]
    ctx = cup3d_b200.Context(rank, 8)
    ctx.comm_init(rank, world, nid)
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
    ctx.close()


@pytest.mark.parametrize("name,coarse", [("u32", 0), ("u64", 0), ("u64", 4096)])
def test_two_rank_matches_single_rank_reference(built, name, coarse):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from cup3d_b200 import capi
    c = case(name)
    nid = capi.nccl_unique_id()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    ps = [ctxm.Process(target=worker, args=(r, 2, nid, name, coarse, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    vc = np.zeros((c.n, 512))
    op = np.zeros((c.n, 512))
    x = np.zeros((c.n, 512))
    for rank, mine, out in got:
        vc[mine], op[mine], x[mine] = out["vc"], out["op"], out["x"]
        assert out["res"] < 1e-10
    assert relerr(vc, c.g["vc_out_cosrhs"]) < 1e-11
    assert relerr(op, c.g["op_out_mc2"]) < 1e-12
    assert relerr(x, c.g["solve_x_mc2"]) < 1e-8


def worker_step(rank, world, nid, name, q):
    os.environ["CUP_COARSE_BLOCKS"] = "0"
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    torch.cuda.set_device(rank)
    import cup3d_b200
    from cup3d_b200 import capi
    from util import case
    c = case(name)
    owner = capi.split_owner(c.n, world)
    mine = np.nonzero(owner == rank)[0]
    ctx = cup3d_b200.Context(rank, 8)
    ctx.comm_init(rank, world, nid)
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
    ctx.close()


@pytest.mark.parametrize("name", ["u16", "b211"])
def test_two_rank_time_step_pieces(built, name):
    """advdiff(), projection() and every stencil sweep, domain-decomposed over 2 GPUs"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from cup3d_b200 import capi
    c = case(name)
    nid = capi.nccl_unique_id()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    ps = [ctxm.Process(target=worker_step, args=(r, 2, nid, name, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
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
