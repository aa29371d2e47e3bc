// mg_device.cuh -- device-side building blocks shared by the multigrid and
// stencil kernels (sm_100a).
//
// Work decomposition used by every kernel here: ONE 8x8x8 BLOCK PER 64-THREAD
// GROUP.  Thread t = x + 8*y owns the z-LINE (x, y, 0..7) in registers.  A
// z-line has stride 64 Reals in the block-major layout, so for every k the 64
// threads touch 64 consecutive Reals: global loads/stores are fully coalesced
// without staging, and the z-neighbours of the 7-point stencil are already in
// registers.  x/y neighbours and the 8-point sine transforms along x and y go
// through shared memory.
#pragma once
#include <cuda_runtime.h>

namespace cup {

enum { TPB = 64 };  // threads per 8^3 block
// neighbour codes (cup_internal.h): >= 0 local slot, -1 wall, <= kRemote0 received face number
enum { kWall = -1, kCoarse = -2, kRemote0 = -3, kFine = -2147483647 - 1 };

// slot -> pointer.  Slots below nleaf are leaves and live in the caller's flat
// vector (block-index order); the rest are synthesised multigrid parents and
// live in library scratch.  This is what lets mg_vcycle run directly on the
// caller's in/out vectors without the two vec_copy passes of the reference
// (main.c:4835, :4852).
template <typename Real>
struct SlotVec {
  Real *leaf;
  Real *extra;
  int nleaf;
  __device__ __forceinline__ Real *at(int slot) const {
    return slot < nleaf ? leaf + (size_t)slot * 512 : extra + (size_t)(slot - nleaf) * 512;
  }
  Real *at_host(int slot) const {  // same mapping, evaluated on the host
    return slot < nleaf ? leaf + (size_t)slot * 512 : extra + (size_t)(slot - nleaf) * 512;
  }
};

struct LevelView {
  const int *act;  // [nact]
  const int *nbr;  // [nact][6]
  int nact;
  const void *rface;  // faces received from other ranks: [nface][64] Reals, plane order (a, c)
  // one-sided transport: the face area is double buffered by the parity of the level's
  // exchange counter (device memory, so a captured CUDA graph stays valid on replay)
  const unsigned long long *seq;
  long long rface_stride;  // Reals between the two copies
  // coarse-fine interfaces (AMR): ext[k][6][4] -- kCoarse: {coarse slot, quadrant of its face};
  // kFine: the four finer blocks across the face, quadrant-major (a/4 + 2*(c/4))
  const int *ext;
  // ghost slabs received from other ranks (stencil sweeps): [face][SLAB planes][64], double buffered
  const void *rslab;
  long long rslab_stride;
  // optional work list (stencil_run(st, list, n), main.c:3631): the sweep visits act[] indices
  // sub[0..nsub) (or 0..nsub-1 when sub is null); nsub < 0: all nact blocks
  const int *sub = nullptr;
  int nsub = -1;
};

__device__ __forceinline__ int lv_count(const LevelView &lv) { return lv.nsub >= 0 ? lv.nsub : lv.nact; }
__device__ __forceinline__ int lv_item(const LevelView &lv, int i) { return lv.sub ? lv.sub[i] : i; }

enum { kSlabPlanes = 9 };
template <typename Real>
__device__ __forceinline__ const Real *rslab_of(const LevelView &lv) {
  const Real *p = (const Real *)lv.rslab;
  if (lv.seq && (*lv.seq & 1))
    p += lv.rslab_stride;
  return p;
}

template <typename Real>
__device__ __forceinline__ const Real *rface_of(const LevelView &lv) {
  const Real *p = (const Real *)lv.rface;
  if (lv.seq && (*lv.seq & 1))
    p += lv.rface_stride;
  return p;
}

// 8-point DST-I matrix S[j][k] = sqrt(2/9) sin(pi (j+1)(k+1)/9), k < 4 only:
// S[j][7-k] = (-1)^j S[j][k]  (pois_init, main.c:4322-4329).
extern __constant__ double cS64[8][4];
extern __constant__ float cS32[8][4];

template <typename Real>
__device__ __forceinline__ Real Sjk(int j, int k);
template <>
__device__ __forceinline__ double Sjk<double>(int j, int k) {
  return cS64[j][k];
}
template <>
__device__ __forceinline__ float Sjk<float>(int j, int k) {
  return cS32[j][k];
}

// In-register 8-point sine transform (pre_x/pre_y/pre_z, main.c:4335-4367, are
// the dense 8x8 form).  The even/odd symmetry of S halves the multiplies:
// 8 add + 32 fma instead of 64 fma.
template <typename Real>
__device__ __forceinline__ void dst8(Real (&v)[8]) {
  const Real p0 = v[0] + v[7], p1 = v[1] + v[6], p2 = v[2] + v[5], p3 = v[3] + v[4];
  const Real m0 = v[0] - v[7], m1 = v[1] - v[6], m2 = v[2] - v[5], m3 = v[3] - v[4];
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    v[j] = ((Sjk<Real>(j, 0) * p0 + Sjk<Real>(j, 1) * p1) + Sjk<Real>(j, 2) * p2) + Sjk<Real>(j, 3) * p3;
    v[j + 1] = ((Sjk<Real>(j + 1, 0) * m0 + Sjk<Real>(j + 1, 1) * m1) + Sjk<Real>(j + 1, 2) * m2) +
               Sjk<Real>(j + 1, 3) * m3;
  }
}

// Bank-conflict-free shared-memory layout of one 8^3 block for the transposes
// between the x, y and z passes.  With 64-bit words a half-warp must hit 16
// distinct 8-byte bank pairs; bank pair = index mod 16 = 8*(row&1) + column.
// Storing (x,y,z) at row (y ^ (z&1)), column x ^ y makes all three access
// patterns (fixed x / fixed y / fixed z across a half-warp) conflict free
// with no padding, and keeps the index a pure XOR of bit fields.
__device__ __forceinline__ int sw(int x, int y, int z) { return (z << 6) + (((y ^ (z & 1))) << 3) + (x ^ y); }

// The block-local exact inverse  v <- Lint^-1 v  (pre_blk, main.c:4368) on the CTA's 8^3 block:
// every thread passes its z-line in and gets its z-line back.  Transform order z,y,x,(scale by
// w = 1/(lam_i+lam_j+lam_k)),x,y,z through the shared buffer ex[512] in layout sw(); four
// __syncthreads.  All shared addresses are one XOR with an immediate away from a per-thread base:
//   z pass  idx = (zb ^ 8(k&1)) + 64k      zb = 8y + (x^y)
//   y pass  idx = yb ^ 9k                  yb = 64 z2 + 8(z2&1) + x2
//   x pass  idx = xb ^ k                   xb = 64 z3 + 8(y3 ^ (z3&1)) + y3
template <typename Real>
__device__ __forceinline__ void fdm_solve(Real (&v)[8], Real *ex, const Real (&w)[8], int t) {
  const int a = t & 7, c = t >> 3;  // (x,y) in the z pass, (x2,z2) in the y pass, (y3,z3) in the x pass
  const int zb0 = (c << 3) + (a ^ c), zb1 = zb0 ^ 8;
  const int yb = (c << 6) + ((c & 1) << 3) + a;
  const int xb = (c << 6) + ((a ^ (c & 1)) << 3) + a;
  dst8<Real>(v);
#pragma unroll
  for (int k = 0; k < 8; k++)
    ex[((k & 1) ? zb1 : zb0) + 64 * k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; k++)
    v[k] = ex[yb ^ (9 * k)];
  dst8<Real>(v);
#pragma unroll
  for (int k = 0; k < 8; k++)
    ex[yb ^ (9 * k)] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; k++)
    v[k] = ex[xb ^ k];
  dst8<Real>(v);
#pragma unroll
  for (int k = 0; k < 8; k++)
    v[k] *= w[k];
  dst8<Real>(v);
#pragma unroll
  for (int k = 0; k < 8; k++)
    ex[xb ^ k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; k++)
    v[k] = ex[yb ^ (9 * k)];
  dst8<Real>(v);
#pragma unroll
  for (int k = 0; k < 8; k++)
    ex[yb ^ (9 * k)] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; k++)
    v[k] = ex[((k & 1) ? zb1 : zb0) + 64 * k];
  dst8<Real>(v);
}

// Load the six ghost faces of block `b` into halo[6][64] (shared).  Face f of a
// regular neighbour is that neighbour's opposite boundary plane; at a domain
// wall the reference's BC op copies the nearest interior cell (zero-gradient
// scalar BC: OP_BC main.c:3524, gen_table.py:195), i.e. the block's OWN plane.
// Element (a, c) = (t & 7, t >> 3) of each face:
//   x faces: a = y, c = z;   y faces: a = x, c = z;   z faces: a = x, c = y.
template <typename Real>
__device__ __forceinline__ void load_halo(const SlotVec<Real> &u, const Real *own, const int *nbr6, int t,
                                          Real (*halo)[64], const Real *rface = nullptr) {
  const int a = t & 7, c = t >> 3;
#pragma unroll
  for (int f = 0; f < 6; f++) {
    const int nb = nbr6[f];
    if (nb <= kRemote0) {  // packed by the owner in exactly this (a, c) order
      halo[f][t] = __ldcg(rface + (size_t)(kRemote0 - nb) * 64 + t);  // stored by another GPU: not through L1
      continue;
    }
    const Real *src = nb >= 0 ? u.at(nb) : own;
    // plane coordinate inside the source block
    const int p = (nb >= 0) ? ((f & 1) ? 0 : 7) : ((f & 1) ? 7 : 0);
    int idx;
    if (f < 2)
      idx = (c << 6) + (a << 3) + p;
    else if (f < 4)
      idx = (c << 6) + (p << 3) + a;
    else
      idx = (p << 6) + t;
    halo[f][t] = src[idx];
  }
}

// ---------------------------------------------------------------------------
// Coarse-fine ghost faces (reference: lab_load :3544 + lab_exec :3401 with the
// tables of gen_table.py for ss = 1, non-tensorial stencils, where only the six
// faces matter).
//
//  * finer neighbours: ghost = 0.125 * sum of the 2x2x2 fine cells (OP_AVG8, fine())
//  * coarser neighbour: OP_FD -- one-sided / centred quadratic interpolation in the two
//    tangential directions on the 4x4 coarse cells facing this block, plus a mixed term,
//    blended with the block's own first two interior cells: (8 v + 10 b - 3 c) / 15.
// Plane element (a, c) and plane order as in load_halo.
// ---------------------------------------------------------------------------
extern __constant__ double cFDp[9], cFDm[9];  // d_coef_plus / d_coef_minus, main.c:3371-3374

// index inside a block of the cell with normal coordinate n (direction d = f/2) and tangential (a, c)
__device__ __forceinline__ int face_idx(int f, int n, int a, int c) {
  return f < 2 ? (c << 6) + (a << 3) + n : (f < 4 ? (c << 6) + (n << 3) + a : (n << 6) + (c << 3) + a);
}

// phase 1 (before a __syncthreads): threads 0..15 fetch the 4x4 coarse cells facing this block
template <typename Real>
__device__ __forceinline__ void coarse_patch_load(const Real *cblk, int f, int quad, int t, Real *patch16) {
  if (t < 16) {
    const int q1 = (t & 3) + 4 * (quad & 1), q2 = (t >> 2) + 4 * (quad >> 1);
    patch16[t] = cblk[face_idx(f, (f & 1) ? 0 : 7, q1, q2)];
  }
}

// OP_FD (main.c:3465-3521) for fine plane element (a, c); bb / cq = the fine block's first and
// second interior cell behind that element; patch = 4x4 coarse cells, index C1 + 4*C2
template <typename Real>
__device__ __forceinline__ Real fd_ghost(const Real *patch, int a, int c, Real bb, Real cq) {
  const int C1 = a >> 1, C2 = c >> 1;
  const double d1 = 0.25 * (2 * (a & 1) - 1), d2 = 0.25 * (2 * (c & 1) - 1);
  const double *c1 = d1 > 0 ? cFDp : cFDm, *c2 = d2 > 0 ? cFDp : cFDm;
  const Real *p0 = patch + C1 + 4 * C2;
  double mixed_coef = 1.0;
  int P1, M1, P2, M2;
  Real x1, x2;
  if (C1 != 0 && C1 != 3) {
    x1 = (c1[6] * p0[-1] + c1[8] * p0[1]) + c1[7] * p0[0];
    P1 = 1;
    M1 = -1;
    mixed_coef *= 0.5;
  } else if (C1 == 0) {
    x1 = (c1[0] * p0[2] + c1[1] * p0[1]) + c1[2] * p0[0];
    P1 = 1;
    M1 = 0;
  } else {
    x1 = (c1[3] * p0[-2] + c1[4] * p0[-1]) + c1[5] * p0[0];
    P1 = 0;
    M1 = -1;
  }
  if (C2 != 0 && C2 != 3) {
    x2 = (c2[6] * p0[-4] + c2[8] * p0[4]) + c2[7] * p0[0];
    P2 = 4;
    M2 = -4;
    mixed_coef *= 0.5;
  } else if (C2 == 0) {
    x2 = (c2[0] * p0[8] + c2[1] * p0[4]) + c2[2] * p0[0];
    P2 = 4;
    M2 = 0;
  } else {
    x2 = (c2[3] * p0[-8] + c2[4] * p0[-4]) + c2[5] * p0[0];
    P2 = 0;
    M2 = -4;
  }
  const Real mixed = mixed_coef * d1 * d2 * ((p0[M1 + M2] + p0[P1 + P2]) - (p0[P1 + M2] + p0[M1 + P2]));
  const Real v = (x1 + x2) + mixed;
  return (Real)(1.0 / 15.0) * ((Real)8.0 * v + ((Real)10.0 * bb - (Real)3.0 * cq));
}

// OP_AVG8 ghost of plane element (a, c) from the four finer blocks across face f.  The 2x2x2
// cluster is summed in the table's order (x outermost, z innermost; gen_table.py fine()).
// Also returns the cluster's two layers: lay[0] = cells touching the face, lay[1] = one behind,
// each [j2][j1] over the 2x2 fine tangential cells (needed by the flux correction).
template <typename Real>
__device__ __forceinline__ Real fine_avg(const SlotVec<Real> &u, const int *ext4, int f, int a, int c,
                                         Real (&lay)[2][2][2]) {
  const Real *fb = u.at(ext4[(a >> 2) + 2 * (c >> 2)]);
  const int g1 = 2 * (a & 3), g2 = 2 * (c & 3);
  const int n0 = (f & 1) ? 0 : 7, n1 = (f & 1) ? 1 : 6;
#pragma unroll
  for (int j2 = 0; j2 < 2; j2++)
#pragma unroll
    for (int j1 = 0; j1 < 2; j1++) {
      lay[0][j2][j1] = fb[face_idx(f, n0, g1 + j1, g2 + j2)];
      lay[1][j2][j1] = fb[face_idx(f, n1, g1 + j1, g2 + j2)];
    }
  // cluster cell (dx,dy,dz) relative to its lowest corner, in block coordinates
  Real s = 0;
  bool first = true;
#pragma unroll
  for (int dx = 0; dx < 2; dx++)
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
      for (int dz = 0; dz < 2; dz++) {
        // normal offset dn (0 = lower coordinate) and tangential offsets of this cluster cell
        int dn, j1, j2;
        if (f < 2) {
          dn = dx; j1 = dy; j2 = dz;
        } else if (f < 4) {
          dn = dy; j1 = dx; j2 = dz;
        } else {
          dn = dz; j1 = dx; j2 = dy;
        }
        // lower normal coordinate is the second layer for a -dir neighbour (cells 6,7), the first for +dir (0,1)
        const int layer = (f & 1) ? dn : 1 - dn;
        const Real val = lay[layer][j2][j1];
        s = first ? val : s + val;
        first = false;
      }
  return (Real)0.125 * s;
}

// Sum of the ghost values adjacent to cell (x, y, k) (zero for interior cells).
template <typename Real>
__device__ __forceinline__ Real ghost_sum(const Real (*halo)[64], int x, int y, int k) {
  Real g = 0;
  if (x == 0)
    g += halo[0][y + 8 * k];
  if (x == 7)
    g += halo[1][y + 8 * k];
  if (y == 0)
    g += halo[2][x + 8 * k];
  if (y == 7)
    g += halo[3][x + 8 * k];
  if (k == 0)
    g += halo[4][x + 8 * y];
  if (k == 7)
    g += halo[5][x + 8 * y];
  return g;
}

// 7-point operator on a z-line: out[k] = h*(xm+xp+ym+yp+zm+zp-6u), summation
// order as in k_lhs/k_mg (main.c:4263, :4277).  `tile` holds the block's own
// values, tile[k*64+t]; halo as above; uu = the thread's own line.
template <typename Real>
__device__ __forceinline__ void lap_line(const Real *tile, const Real (*halo)[64], const Real (&uu)[8], int x, int y,
                                         int t, Real h, Real (&out)[8]) {
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const Real xm = x > 0 ? tile[k * 64 + t - 1] : halo[0][y + 8 * k];
    const Real xp = x < 7 ? tile[k * 64 + t + 1] : halo[1][y + 8 * k];
    const Real ym = y > 0 ? tile[k * 64 + t - 8] : halo[2][x + 8 * k];
    const Real yp = y < 7 ? tile[k * 64 + t + 8] : halo[3][x + 8 * k];
    const Real zm = k > 0 ? uu[k > 0 ? k - 1 : 0] : halo[4][t];
    const Real zp = k < 7 ? uu[k < 7 ? k + 1 : 7] : halo[5][t];
    out[k] = h * ((((((xm + xp) + ym) + yp) + zm) + zp) - (Real)6.0 * uu[k]);
  }
}

}  // namespace cup
