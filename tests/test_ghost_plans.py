"""CPU: the ghost-BLOCK plans of multi-level meshes split over several ranks (csrc/mesh.cpp through
cup_plan_build): on a level with coarse-fine interfaces, and in the leaf context, every block of another
rank that a local block reads gets a local ghost slot, and the owners push whole blocks into them (the
reference's halo_sync ships whole blocks too, main.c:3101-3112; its neighbour rules: lab_load :3579-3602).

The exchange is emulated in-process on block IDENTITIES: after it, every neighbour slot of every local
block must hold exactly the block that the single-rank tables name for that face, every ghost slot must
have been written exactly once, and interface blocks must be able to reach all leaves behind their edges
and corners (the wide advdiff stencil).  Meshes: the reference's own adapted meshes (tests/golden amr2/amr3)."""
import numpy as np
import pytest

from util import AMR_CASES, case

WALL, COARSE, FINE = -1, -2, -2147483648


def build(c, world, rank, level, owner):
    from cup3d_b200 import capi
    return capi.plan_build(c.ib, c.rb, owner, world, rank, c.bpd, c.level_max, level)


@pytest.mark.parametrize("name", AMR_CASES)
@pytest.mark.parametrize("world", [2, 3, 5])
def test_leaf_context_ghost_blocks(built, name, world):
    from cup3d_b200 import capi
    c = case(name)
    owner = capi.split_owner(c.n, world)
    T = build(c, 1, 0, -1, np.zeros(c.n, np.int32))      # single rank: slot == global leaf index
    P = [build(c, world, r, -1, owner) for r in range(world)]
    mine = [np.nonzero(owner == r)[0] for r in range(world)]
    gid = [np.full(len(mine[r]) + P[r]["nghost"], -1, np.int64) for r in range(world)]
    for r in range(world):
        assert P[r]["ghosted"] and P[r]["nbrecv"] == P[r]["nghost"]
        gid[r][: len(mine[r])] = mine[r]
    for s in range(world):  # the exchange, on identities
        for e in range(P[s]["nbsend"]):
            d, idx = int(P[s]["bsend_peer"][e]), int(P[s]["bsend_idx"][e])
            slot = int(P[d]["brecv_slot"][idx])
            assert gid[d][slot] == -1, "ghost slot written twice"
            assert int(P[s]["bsend_slot"][e]) < len(mine[s])
            gid[d][slot] = mine[s][int(P[s]["bsend_slot"][e])]
    where = {tuple(b): i for i, b in enumerate(c.ib.tolist())}
    nchecked = 0
    for r in range(world):
        assert np.all(gid[r] >= 0), "a ghost slot was never filled"
        assert np.all(owner[gid[r][len(mine[r]):]] != r)
        have = set(gid[r].tolist())
        for k, g in enumerate(mine[r]):
            iface = False
            for f in range(6):
                code, tcode = int(P[r]["nbr"][k, f]), int(T["nbr"][g, f])
                if tcode >= 0:
                    assert gid[r][code] == tcode
                elif tcode == COARSE:
                    iface = True
                    assert code == COARSE and gid[r][P[r]["ext"][k, f, 0]] == T["ext"][g, f, 0]
                    assert P[r]["ext"][k, f, 1] == T["ext"][g, f, 1]
                elif tcode == FINE:
                    iface = True
                    assert code == FINE
                    assert [gid[r][q] for q in P[r]["ext"][k, f]] == list(T["ext"][g, f])
                else:
                    assert code == tcode == WALL
                nchecked += 1
            if not iface:
                continue
            # edges and corners of an interface block: the covering leaf at its level, one coarser, or the
            # finer leaves touching it must be readable on this rank
            L, ix, iy, iz = c.ib[g]
            dim = [b << L for b in c.bpd]
            for dz in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        q = (ix + dx, iy + dy, iz + dz)
                        if (dx, dy, dz) == (0, 0, 0) or not all(0 <= q[d] < dim[d] for d in range(3)):
                            continue
                        if (L, *q) in where:
                            assert where[(L, *q)] in have
                        elif (L - 1, q[0] // 2, q[1] // 2, q[2] // 2) in where:
                            assert where[(L - 1, q[0] // 2, q[1] // 2, q[2] // 2)] in have
                        else:
                            for cz in (0, 1):
                                for cy in (0, 1):
                                    for cx in (0, 1):
                                        bits, off = (cx, cy, cz), (dx, dy, dz)
                                        if any((off[d] < 0 and bits[d] == 0) or (off[d] > 0 and bits[d] == 1)
                                               for d in range(3)):
                                            continue
                                        key = (L + 1, 2 * q[0] + cx, 2 * q[1] + cy, 2 * q[2] + cz)
                                        if key in where:
                                            assert where[key] in have
    assert nchecked > 0


@pytest.mark.parametrize("name", AMR_CASES)
@pytest.mark.parametrize("world,coarse", [(2, 0), (3, 0), (3, 4096), (5, 0)])
def test_multigrid_levels_ghost_blocks(built, name, world, coarse, monkeypatch):
    from cup3d_b200 import capi
    monkeypatch.setenv("CUP_COARSE_BLOCKS", str(coarse))
    c = case(name)
    owner = capi.split_owner(c.n, world)
    nlev = c.level_max
    P = [[build(c, world, r, L, owner) for L in range(nlev)] for r in range(world)]
    mine = [np.nonzero(owner == r)[0] for r in range(world)]
    # identity (level, ix, iy, iz) of every local slot of every rank
    ident = [dict() for _ in range(world)]
    for r in range(world):
        for j, g in enumerate(mine[r]):
            ident[r][j] = tuple(int(v) for v in c.ib[g])
        for L in range(nlev):
            for k, s in enumerate(P[r][L]["act"]):
                t = (L, *[int(v) for v in P[r][L]["ijk"][k]])
                assert ident[r].setdefault(int(s), t) == t
    own = [dict(d) for d in ident]
    leaves = set(map(tuple, c.ib.tolist()))
    any_ghosted = False
    for L in range(nlev):
        for s in range(world):
            for e in range(P[s][L]["nbsend"]):
                d, idx = int(P[s][L]["bsend_peer"][e]), int(P[s][L]["bsend_idx"][e])
                slot, kind = int(P[d][L]["brecv_slot"][idx]), int(P[s][L]["bsend_kind"][e])
                t = own[s][int(P[s][L]["bsend_slot"][e])]
                assert kind == int(P[d][L]["brecv_kind"][idx]) == (0 if t[0] == L else 1)
                if kind == 1:
                    assert t[0] == L - 1 and t in leaves  # a coarser LEAF behind an interface
                assert ident[d].setdefault(slot, t) == t and slot not in own[d]
        for r in range(world):
            Q = P[r][L]
            dim = [b << L for b in c.bpd]
            any_ghosted |= Q["ghosted"]
            got = set()
            for idx in range(Q["nbrecv"]):
                assert int(Q["brecv_slot"][idx]) in ident[r], "ghost slot never filled"
                got.add(int(Q["brecv_slot"][idx]))
            for k in range(Q["nact"]):
                for f in range(6):
                    code = int(Q["nbr"][k, f])
                    nijk = [int(v) for v in Q["ijk"][k]]
                    nijk[f // 2] += 1 if (f & 1) else -1
                    if code == WALL:
                        assert not 0 <= nijk[f // 2] < dim[f // 2]
                    elif code == COARSE:
                        cs = int(Q["ext"][k, f, 0])
                        assert ident[r][cs] == (L - 1, nijk[0] // 2, nijk[1] // 2, nijk[2] // 2)
                        assert cs in own[r] or cs in got
                    elif code >= 0:
                        assert ident[r][code] == (L, *nijk)
                        assert code in own[r] or code in got
                    else:
                        assert code <= -3 and not Q["ghosted"]  # uniform level: an 8x8 face travels
    assert any_ghosted
