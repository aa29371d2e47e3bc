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
