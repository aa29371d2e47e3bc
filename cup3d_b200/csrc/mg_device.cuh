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
enum { kWall = -1, kCoarse = -2, kRemote0 = -3 };

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
};

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
      halo[f][t] = rface[(size_t)(kRemote0 - nb) * 64 + t];
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
