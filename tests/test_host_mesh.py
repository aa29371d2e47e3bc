"""CPU: the host-side topology builder (csrc/mesh.cpp:build_tables, reached through the host-only
cup_plan_build) on the meshes the reference produced -- uniform, non-cubic, multi-level -- and its error
behaviour on malformed input.  What it replaces: tree_sync / halo_build / mg_build (main.c:2928, :3030,
:4435-4681); the reference fatal()s on an inconsistent tree, the library returns a status + message."""
import numpy as np
import pytest

from util import ALL_CASES, AMR_CASES, case
from cup3d_b200 import capi

WALL, COARSE, FINE = -1, -2, -2147483648


def plans(c, nranks=1, rank=0, owner=None):
    own = np.zeros(c.n, np.int32) if owner is None else owner
    return [capi.plan_build(c.ib, c.rb, own, nranks, rank, c.bpd, c.level_max, L) for L in range(int(c.ib[:, 0].max()) + 1)]


@pytest.mark.parametrize("name", ALL_CASES + AMR_CASES)
def test_level_population_and_slots(built, name):
    """level L holds its own leaves plus one synthesised parent per 8 blocks of level L+1 (main.c:4540, :4590)"""
    c = case(name)
    P = plans(c)
    counts = np.bincount(c.ib[:, 0], minlength=len(P))
    above = 0
    for L in range(len(P) - 1, -1, -1):
        assert P[L]["nact"] == counts[L] + above // 8, (name, L)
        above = P[L]["nact"]
    nslot = P[0]["nslot"]
    assert all(p["nslot"] == nslot and p["nblk"] == c.n for p in P)
    # slots: leaves keep their block index, parents follow; every slot is active on exactly one level
    seen = np.concatenate([p["act"] for p in P])
    assert len(seen) == nslot and np.array_equal(np.sort(seen), np.arange(nslot))
    for L, p in enumerate(P):
        leaves = p["act"][p["act"] < c.n]
        assert np.all(c.ib[leaves, 0] == L)


@pytest.mark.parametrize("name", ALL_CASES + AMR_CASES)
def test_neighbour_tables(built, name):
    c = case(name)
    P = plans(c)
    for L, p in enumerate(P):
        nb, ijk, act = p["nbr"], p["ijk"], p["act"]
        pos = {int(s): k for k, s in enumerate(act)}
        dims = [b << L for b in c.bpd]
        for k in range(p["nact"]):
            for f in range(6):
                d, hi = f >> 1, f & 1
                at_wall = ijk[k, d] == (dims[d] - 1 if hi else 0)
                code = int(nb[k, f])
                assert (code == WALL) == bool(at_wall), (name, L, k, f, code)
                if code >= 0:                       # same-level neighbour: adjacent index, reciprocal entry
                    j = pos[code]
                    exp = ijk[k].copy()
                    exp[d] += 1 if hi else -1
                    assert np.array_equal(ijk[j], exp)
                    assert int(nb[j, f ^ 1]) == int(act[k])
                elif not at_wall:
                    # multigrid contexts only ever see coarser neighbours (finer ones have a parent here)
                    assert code == COARSE and name in AMR_CASES
    # parents: each has 8 distinct octants, children one level up
    for L in range(1, len(P)):
        p = P[L]
        ps, oc = p["pslot"], p["oct"]
        assert np.all((oc >= 0) & (oc < 8))
        assert np.array_equal(oc, (p["ijk"][:, 0] & 1) + 2 * (p["ijk"][:, 1] & 1) + 4 * (p["ijk"][:, 2] & 1))
        parents, cnt = np.unique(ps, return_counts=True)
        assert np.all(cnt == 8) and set(parents) <= set(P[L - 1]["act"])


def test_uniform_mesh_has_no_coarse_codes(built):
    for name in ALL_CASES:
        c = case(name)
        for p in plans(c):
            assert not np.any(p["nbr"] == COARSE) and not np.any(p["nbr"] == FINE)


def _expect_error(fn, text):
    with pytest.raises(capi.CupError) as e:
        fn()
    assert text in str(e.value), str(e.value)


def test_malformed_meshes_are_rejected(built):
    c = case("amr2")
    own = np.zeros(c.n, np.int32)
    fine = int(np.argmax(c.ib[:, 0]))
    # duplicate block
    ib = np.vstack([c.ib, c.ib[:1]])
    rb = np.vstack([c.rb, c.rb[:1]])
    _expect_error(lambda: capi.plan_build(ib, rb, np.zeros(len(ib), np.int32), 1, 0, c.bpd, c.level_max, 0), "duplicate")
    # a missing sibling: drop one finest block
    keep = np.arange(c.n) != fine
    _expect_error(lambda: capi.plan_build(c.ib[keep], c.rb[keep], own[keep], 1, 0, c.bpd, c.level_max, 0), "")
    # level outside [0, level_max)
    ib = c.ib.copy()
    ib[0, 0] = c.level_max
    _expect_error(lambda: capi.plan_build(ib, c.rb, own, 1, 0, c.bpd, c.level_max, 0), "level")
    # index outside the level
    ib = c.ib.copy()
    ib[fine, 1] = (c.bpd[0] << int(ib[fine, 0]))
    _expect_error(lambda: capi.plan_build(ib, c.rb, own, 1, 0, c.bpd, c.level_max, 0), "outside")
    # owner outside the communicator
    bad = own.copy()
    bad[3] = 5
    _expect_error(lambda: capi.plan_build(c.ib, c.rb, bad, 2, 0, c.bpd, c.level_max, 0), "owner")
    # empty mesh / bad arguments
    _expect_error(lambda: capi.plan_build(c.ib[:0], c.rb[:0], own[:0], 1, 0, c.bpd, c.level_max, 0), "bad arguments")
    # a level that does not exist
    _expect_error(lambda: capi.plan_build(c.ib, c.rb, own, 1, 0, c.bpd, c.level_max, 99), "level")


def test_multi_level_meshes_split_over_ranks(built):
    """multi-level meshes across ranks: levels with interfaces use ghost blocks, uniform levels 8x8 faces
    (the plans themselves are checked in tests/test_ghost_plans.py)"""
    c = case("amr2")
    own = capi.split_owner(c.n, 2)
    finest = int(c.ib[:, 0].max())
    for r in range(2):
        leafp = capi.plan_build(c.ib, c.rb, own, 2, r, c.bpd, c.level_max, -1)
        assert leafp["ghosted"] and leafp["nghost"] > 0 and leafp["nact"] == int((own == r).sum())
        top = capi.plan_build(c.ib, c.rb, own, 2, r, c.bpd, c.level_max, finest)
        assert top["ghosted"] and top["nsend"] == 0 and top["nrecv"] == 0
        base = capi.plan_build(c.ib, c.rb, own, 2, r, c.bpd, c.level_max, 0)
        assert not base["ghosted"] and base["nghost"] == 0


def test_broken_two_to_one_balance(built):
    """a leaf two levels finer than its face neighbour is rejected (the reference's mesh_fix forbids it)"""
    c = case("u16")                       # 2x2x2 blocks at level 1
    h1 = c.rb[0, 0]

    def children(level, ix, iy, iz, h):
        out = []
        for q in range(8):
            x, y, z = 2 * ix + (q & 1), 2 * iy + ((q >> 1) & 1), 2 * iz + (q >> 2)
            out.append(((level + 1, x, y, z), (h / 2, x * 8 * h / 2, y * 8 * h / 2, z * 8 * h / 2)))
        return out

    def build(refine_lvl2):
        """level-1 mesh with block (0,0,0) refined, and the listed level-2 children refined once more"""
        ib, rb = [], []
        for i in range(c.n):
            if tuple(c.ib[i, 1:4]) != (0, 0, 0):
                ib.append(tuple(c.ib[i]))
                rb.append(tuple(c.rb[i]))
        for (i4, r4) in children(1, 0, 0, 0, h1):
            if i4[1:] in refine_lvl2:
                for (j4, s4) in children(2, *i4[1:], h1 / 2):
                    ib.append(j4)
                    rb.append(s4)
            else:
                ib.append(i4)
                rb.append(r4)
        return np.array(ib, np.int32), np.array(rb, np.float64)

    # the domain-corner child touches only walls and its level-2 siblings: refining it keeps the balance
    ib, rb = build({(0, 0, 0)})
    p = capi.plan_build(ib, rb, np.zeros(len(ib), np.int32), 1, 0, c.bpd, 4, 3)
    assert p["nact"] == 8 and p["nblk"] == len(ib) == 7 + 7 + 8
    # the opposite child touches the level-1 blocks: refining it puts level 3 next to level 1
    ib, rb = build({(0, 0, 0), (1, 1, 1)})
    _expect_error(lambda: capi.plan_build(ib, rb, np.zeros(len(ib), np.int32), 1, 0, c.bpd, 4, 0), "balance")
