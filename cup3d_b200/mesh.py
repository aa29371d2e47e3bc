"""Host-side block lists for synthetic meshes (bench/tests).

The reference orders sta.blk[] along a Hilbert curve (sfc_forward,
main.c:1355; Skilling's transpose algorithm).  The device tables do not depend
on the order, but a locality-preserving order keeps neighbouring blocks close
in HBM (better L2 reuse of ghost faces) and makes the flat vectors directly
comparable with the reference's.  `hilbert_index` is an independent,
vectorised numpy implementation of the same curve.
"""
import numpy as np


def hilbert_index(ix, iy, iz, bits):
    """Hilbert index of integer points (arrays) in a 2^bits cube; matches the
    reference's axes_to_transpose (main.c:1213) for its regular (power-of-two) case."""
    X = [np.asarray(ix, np.int64).copy(), np.asarray(iy, np.int64).copy(), np.asarray(iz, np.int64).copy()]
    if bits == 0:
        return np.zeros_like(X[0])
    M = 1 << (bits - 1)
    # inverse undo excess work
    Q = M
    while Q > 1:
        P = Q - 1
        for i in range(3):
            hit = (X[i] & Q) != 0
            X[0] = np.where(hit, X[0] ^ P, X[0])
            t = np.where(hit, 0, (X[0] ^ X[i]) & P)
            X[0] ^= t
            X[i] ^= t
        Q >>= 1
    # Gray encode
    for i in range(1, 3):
        X[i] ^= X[i - 1]
    t = np.zeros_like(X[0])
    Q = M
    while Q > 1:
        t = np.where((X[2] & Q) != 0, t ^ (Q - 1), t)
        Q >>= 1
    for i in range(3):
        X[i] ^= t
    # interleave: bit `l` of X[0] is the most significant of its triple
    out = np.zeros_like(X[0])
    for l in range(bits):
        out |= ((X[2] >> l) & 1) << (3 * l)
        out |= ((X[1] >> l) & 1) << (3 * l + 1)
        out |= ((X[0] >> l) & 1) << (3 * l + 2)
    return out


def uniform_blocks(level, bpd=(1, 1, 1), extent=1.0):
    """All blocks of one level of a bpd[0] x bpd[1] x bpd[2] base grid, Hilbert ordered
    (power-of-two cubes; other shapes fall back to per-base-block ordering).
    Returns (ib int32 [n,4] level,ix,iy,iz ; rb float64 [n,4] h,origin)."""
    n = [b << level for b in bpd]
    h0 = extent / max(bpd) / 8.0  # sim.h0, main.c:1589
    h = h0 / (1 << level)
    iz, iy, ix = np.meshgrid(np.arange(n[2]), np.arange(n[1]), np.arange(n[0]), indexing="ij")
    ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()
    nmax = max(n)
    bits = int(np.ceil(np.log2(nmax))) if nmax > 1 else 0
    order = np.argsort(hilbert_index(ix, iy, iz, bits), kind="stable")
    ix, iy, iz = ix[order], iy[order], iz[order]
    ib = np.stack([np.full_like(ix, level), ix, iy, iz], 1).astype(np.int32)
    rb = np.stack([np.full(len(ix), h), ix * 8 * h, iy * 8 * h, iz * 8 * h], 1).astype(np.float64)
    return ib, rb


def cell_centers(ib, rb):
    """x,y,z of every cell: three arrays [n,8,8,8] indexed [blk, z, y, x]"""
    h = rb[:, 0]
    c = np.arange(8) + 0.5
    X = rb[:, 1, None, None, None] + h[:, None, None, None] * c[None, None, None, :] + 0 * c[None, :, None, None]
    Y = rb[:, 2, None, None, None] + h[:, None, None, None] * c[None, None, :, None] + 0 * c[None, None, None, :]
    Z = rb[:, 3, None, None, None] + h[:, None, None, None] * c[None, :, None, None] + 0 * c[None, None, None, :]
    X = np.broadcast_to(X, (len(h), 8, 8, 8))
    Y = np.broadcast_to(Y, (len(h), 8, 8, 8))
    Z = np.broadcast_to(Z, (len(h), 8, 8, 8))
    return X, Y, Z


def amr_blocks(base_level, max_level, refine, bpd=(1, 1, 1), extent=1.0):
    """A 2:1-balanced multi-level block list for synthetic tests/benchmarks.

    Starts from the uniform `base_level` grid and refines every leaf (below `max_level`) for which
    refine(level, x0, y0, z0, size) is true, then enforces the reference's balance rule (no leaf
    touches -- over its 26 neighbours -- a leaf more than one level coarser; mesh_fix, main.c:3717).
    Leaves are ordered along the Hilbert curve of the finest level (children stay contiguous), like
    the reference's blk_sort.  Returns (ib, rb) as uniform_blocks."""
    h0 = extent / max(bpd) / 8.0
    leaves = set()
    n0 = [b << base_level for b in bpd]
    for k in range(n0[2]):
        for j in range(n0[1]):
            for i in range(n0[0]):
                leaves.add((base_level, i, j, k))

    def split(leaf):
        l, i, j, k = leaf
        leaves.discard(leaf)
        for dk in range(2):
            for dj in range(2):
                for di in range(2):
                    leaves.add((l + 1, 2 * i + di, 2 * j + dj, 2 * k + dk))

    for _ in range(max_level - base_level):
        for leaf in list(leaves):
            l, i, j, k = leaf
            s = 8 * h0 / (1 << l)
            if l < max_level and refine(l, i * s, j * s, k * s, s):
                split(leaf)

    def covering(l, i, j, k):
        """leaf covering block (l,i,j,k) at its own or a coarser level, else None"""
        while l >= 0:
            if (l, i, j, k) in leaves:
                return (l, i, j, k)
            l, i, j, k = l - 1, i >> 1, j >> 1, k >> 1
        return None

    changed = True
    while changed:
        changed = False
        for leaf in sorted(leaves, key=lambda t: -t[0]):
            if leaf not in leaves:
                continue
            l, i, j, k = leaf
            nl = [b << l for b in bpd]
            for dk in (-1, 0, 1):
                for dj in (-1, 0, 1):
                    for di in (-1, 0, 1):
                        a, b, c = i + di, j + dj, k + dk
                        if (di, dj, dk) == (0, 0, 0) or not (0 <= a < nl[0] and 0 <= b < nl[1] and 0 <= c < nl[2]):
                            continue
                        cov = covering(l, a, b, c)
                        if cov is not None and cov[0] < l - 1:
                            split(cov)
                            changed = True
    arr = np.array(sorted(leaves), dtype=np.int64)
    lv, ix, iy, iz = arr[:, 0], arr[:, 1], arr[:, 2], arr[:, 3]
    sh = max_level - lv
    nmax = max(bpd) << max_level
    bits = int(np.ceil(np.log2(nmax))) if nmax > 1 else 0
    key = hilbert_index(ix << sh, iy << sh, iz << sh, bits)
    order = np.lexsort((lv, key))
    lv, ix, iy, iz = lv[order], ix[order], iy[order], iz[order]
    h = h0 / (1 << lv).astype(np.float64)
    ib = np.stack([lv, ix, iy, iz], 1).astype(np.int32)
    rb = np.stack([h, ix * 8 * h, iy * 8 * h, iz * 8 * h], 1).astype(np.float64)
    return ib, rb


def sphere_shell(center, radius, band=1.0):
    """refine() predicate: blocks intersecting the shell | |x - c| - radius | < band * block size"""
    c = np.asarray(center, float)

    def f(level, x0, y0, z0, s):
        lo = np.array([x0, y0, z0])
        nearest = np.clip(c, lo, lo + s)
        dmin = np.linalg.norm(nearest - c)
        far = np.where(np.abs(lo - c) > np.abs(lo + s - c), lo, lo + s)
        dmax = np.linalg.norm(far - c)
        return dmin - band * s < radius < dmax + band * s

    return f
