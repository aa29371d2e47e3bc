"""CPU, world_size 2 and 3 over gloo: the multi-rank exchange PLANS built by the
library's host code (csrc/mesh.cpp via cup_plan_build) are executed with numpy
packing + torch.distributed all_to_all, and every received ghost face /
restricted child is checked against the value the global field has there.
This is the host-side logic of the N>1 path (halo_build / mg_build's transfer
plans in the reference, main.c:3030, :4621-4664); the device kernels that pack
and consume the same buffers are covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def field(L, ijk, comp=0):
    """deterministic value of every cell of block ijk at level L: [z][y][x]"""
    c = np.arange(8)
    X = ijk[0] * 8 + c[None, None, :]
    Y = ijk[1] * 8 + c[None, :, None]
    Z = ijk[2] * 8 + c[:, None, None]
    return np.sin(0.37 * X + 0.11 * L + comp) + np.cos(0.23 * Y) * (1 + 0.01 * Z) + 0.001 * Z * X


def plane(U, p):
    q = 7 if (p & 1) else 0
    if p < 2:
        return U[:, :, q].reshape(-1)   # [z][y]
    if p < 4:
        return U[:, q, :].reshape(-1)   # [z][x]
    return U[q, :, :].reshape(-1)       # [y][x]


def a2a(send, scnt, rcnt, width):
    out = torch.zeros(int(sum(rcnt)) * width, dtype=torch.float64)
    dist.all_to_all_single(out, torch.from_numpy(np.ascontiguousarray(send).reshape(-1)),
                           [int(c) * width for c in rcnt], [int(c) * width for c in scnt])
    return out.numpy().reshape(-1, width)


def worker(rank, world, port, level, bpd, coarse):
    sys.path.insert(0, ROOT)
    os.environ["CUP_COARSE_BLOCKS"] = str(coarse)  # 0: parents follow their children's owner
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cup3d_b200 import capi, mesh
    ib, rb = mesh.uniform_blocks(level, bpd)
    owner = capi.split_owner(len(ib), world)
    level_max = level + 1
    plans = [capi.plan_build(ib, rb, owner, world, rank, bpd, level_max, L) for L in range(level + 1)]
    nchecked = 0
    for L in range(level, -1, -1):
        P = plans[L]
        dim = [b << L for b in bpd]
        slot2k = {int(s): k for k, s in enumerate(P["act"])}
        U = [field(L, P["ijk"][k]) for k in range(P["nact"])]
        # ---- face exchange
        send = np.zeros((P["nsend"], 64))
        for e in range(P["nsend"]):
            send[e] = plane(U[slot2k[int(P["send_slot"][e])]], int(P["send_plane"][e]))
        recv = a2a(send, P["send_cnt"], P["recv_cnt"], 64)
        assert recv.shape[0] == P["nrecv"]
        for k in range(P["nact"]):
            for f in range(6):
                code = int(P["nbr"][k, f])
                nijk = P["ijk"][k].copy()
                nijk[f // 2] += 1 if (f & 1) else -1
                outside = nijk[f // 2] < 0 or nijk[f // 2] >= dim[f // 2]
                if code == -1:
                    assert outside
                elif code >= 0:
                    assert not outside and np.array_equal(P["ijk"][slot2k[code]], nijk)
                else:
                    assert code <= -3 and not outside
                    want = plane(field(L, nijk), f ^ 1)
                    assert np.array_equal(recv[-3 - code], want), (L, k, f)
                    nchecked += 1
        # ---- restriction to remote parents (children send 128 values)
        if L >= 1:
            Q = plans[L - 1]
            nsend = int(P["res_send_cnt"].sum())
            sbuf = np.zeros((nsend, 128))
            qmap = {int(s): i for i, s in enumerate(Q["act"])}
            for k in range(P["nact"]):
                ps = int(P["pslot"][k])
                c = P["ijk"][k]
                assert int(P["oct"][k]) == int((c[0] & 1) + 2 * (c[1] & 1) + 4 * (c[2] & 1))
                if ps <= -3:
                    sbuf[-3 - ps, :64] = field(L, c, 1)[::2, ::2, ::2].reshape(-1)
                    sbuf[-3 - ps, 64:] = field(L, c, 2)[::2, ::2, ::2].reshape(-1)
                else:  # local parent: an active block of level L-1 with the right index
                    assert np.array_equal(Q["ijk"][qmap[ps]], c // 2)
            rbuf = a2a(sbuf, P["res_send_cnt"], P["res_recv_cnt"], 128)
            for e in range(rbuf.shape[0]):
                pijk = Q["ijk"][qmap[int(P["res_recv_slot"][e])]]
                o = int(P["res_recv_oct"][e])
                cijk = 2 * pijk + np.array([o & 1, (o >> 1) & 1, o >> 2])
                assert np.array_equal(rbuf[e, :64], field(L, cijk, 1)[::2, ::2, ::2].reshape(-1))
                assert np.array_equal(rbuf[e, 64:], field(L, cijk, 2)[::2, ::2, ::2].reshape(-1))
                nchecked += 1
    tot = torch.tensor([nchecked])
    dist.all_reduce(tot)
    assert int(tot) > 0
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,level,bpd,coarse", [(2, 2, (1, 1, 1), 0), (3, 2, (1, 1, 1), 0), (2, 1, (2, 1, 3), 0),
                                                    (3, 2, (1, 1, 1), 4096)])
def test_exchange_plans_over_gloo(built, world, level, bpd, coarse):
    mp.spawn(worker, args=(world, free_port(), level, bpd, coarse), nprocs=world, join=True)
