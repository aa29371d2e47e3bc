// amr_sample.cuh -- the level-(L-1) VIEW of a field around a block, as a function (device).
//
// The reference builds, per block, a "coarse scratch" (struct Lab.coarse, main.c:3357): the level L-1
// picture of the neighbourhood -- coarser leaves copied, same-level blocks averaged 2x2x2 (same_cfill /
// own_avg in gen_table.py), the boundary cell repeated beyond walls -- and interpolates ghost cells of
// coarser neighbours from it (OP_INTERP, main.c:3439; OP_FD, :3465).  Here the scratch is never stored:
// cs_sample() evaluates it cell by cell through a hash of the leaves this rank can read (own + ghost
// blocks).  Used by the wide advdiff ghost fill (amr_advdiff.cu) and by the tensorial labs of the
// adaptation kernels (adapt_kernels.cu).
#pragma once
#include "mg_device.cuh"

namespace cup {

struct LeafGeom {
  const int *bijk;                  // [nleaf][4] level, ix, iy, iz
  const unsigned long long *hkeys;  // key + 1, 0 = empty
  const int *hvals;
  unsigned long long hmask;
  int bpd[3];
};

__device__ __forceinline__ int leaf_find(const LeafGeom &g, int level, int ix, int iy, int iz) {
  const unsigned long long k =
      (((unsigned long long)level << 57) | ((unsigned long long)iz << 38) | ((unsigned long long)iy << 19) |
       (unsigned long long)ix) + 1ULL;
  unsigned long long h = ((k * 0x9E3779B97F4A7C15ULL) >> 20) & g.hmask;
  for (;;) {
    const unsigned long long kk = g.hkeys[h];
    if (kk == k)
      return g.hvals[h];
    if (kk == 0)
      return -1;
    h = (h + 1) & g.hmask;
  }
}

// value of component `cidx` (flat vector cp) in the level-Lc view at global cell (gx, gy, gz)
template <typename Real>
__device__ __noinline__ Real cs_sample(const LeafGeom &g, const Real *__restrict__ cp, int cidx, int Lc, int gx, int gy,
                                       int gz) {
  Real sign = 1;
  int q[3] = {gx, gy, gz};
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int n = (g.bpd[d] << Lc) * 8;
    if (q[d] < 0) {
      q[d] = 0;
      if (cidx == d)
        sign = -sign;
    } else if (q[d] >= n) {
      q[d] = n - 1;
      if (cidx == d)
        sign = -sign;
    }
  }
  int slot = leaf_find(g, Lc, q[0] >> 3, q[1] >> 3, q[2] >> 3);
  if (slot >= 0)
    return sign * cp[(size_t)slot * 512 + ((q[2] & 7) << 6) + ((q[1] & 7) << 3) + (q[0] & 7)];
  const int fx = 2 * q[0], fy = 2 * q[1], fz = 2 * q[2];
  slot = leaf_find(g, Lc + 1, fx >> 3, fy >> 3, fz >> 3);
  if (slot < 0)
    return 0;  // not reachable on a 2:1 balanced mesh
  const Real *b = cp + (size_t)slot * 512 + ((fz & 7) << 6) + ((fy & 7) << 3) + (fx & 7);
  // same_cfill order: x outermost, z innermost
  const Real s = ((((((b[0] + b[64]) + b[8]) + b[72]) + b[1]) + b[65]) + b[9]) + b[73];
  return sign * (Real)0.125 * s;
}

// OP_INTERP (main.c:3439-3463): Taylor expansion around coarse cell (cx,cy,cz) towards the
// child with offsets (sx,sy,sz) = +-1
template <typename Real>
__device__ __noinline__ Real interp_ghost(const LeafGeom &g, const Real *__restrict__ cp, int cidx, int Lc, int cx,
                                          int cy, int cz, Real sx, Real sy, Real sz) {
#define C3(I, J, K) cs_sample<Real>(g, cp, cidx, Lc, cx + (I)-1, cy + (J)-1, cz + (K)-1)
  const Real c111 = C3(1, 1, 1);
  const Real c011 = C3(0, 1, 1), c211 = C3(2, 1, 1), c101 = C3(1, 0, 1), c121 = C3(1, 2, 1), c110 = C3(1, 1, 0),
             c112 = C3(1, 1, 2);
  const Real dudx = (Real)0.125 * (c211 - c011);
  const Real dudy = (Real)0.125 * (c121 - c101);
  const Real dudz = (Real)0.125 * (c112 - c110);
  const Real dudxdy = (Real)0.015625 * (((C3(0, 0, 1) + C3(2, 2, 1)) - C3(2, 0, 1)) - C3(0, 2, 1));
  const Real dudxdz = (Real)0.015625 * (((C3(0, 1, 0) + C3(2, 1, 2)) - C3(2, 1, 0)) - C3(0, 1, 2));
  const Real dudydz = (Real)0.015625 * (((C3(1, 0, 0) + C3(1, 2, 2)) - C3(1, 2, 0)) - C3(1, 0, 2));
  const Real lap =
      c111 + (Real)0.03125 * ((((((c011 + c211) + c101) + c121) + c110) + c112) + (Real)(-6.0) * c111);
#undef C3
  return (((((lap + sx * dudx) + sy * dudy) + sz * dudz) + sx * sy * dudxdy) + sx * sz * dudxdz) + sy * sz * dudydz;
}

// tangential part of OP_FD (the value v before the blend); see fd_ghost in mg_device.cuh
template <typename Real>
__device__ __forceinline__ Real fd_tangential(const Real *patch, int a, int c) {
  const int C1 = a >> 1, C2 = c >> 1;
  const double d1 = 0.25 * (2 * (a & 1) - 1), d2 = 0.25 * (2 * (c & 1) - 1);
  const double *c1 = d1 > 0 ? cFDp : cFDm, *c2 = d2 > 0 ? cFDp : cFDm;
  const Real *p0 = patch + C1 + 4 * C2;
  double mixed_coef = 1.0;
  int P1, M1, P2, M2;
  Real x1, x2;
  if (C1 != 0 && C1 != 3) {
    x1 = (c1[6] * p0[-1] + c1[8] * p0[1]) + c1[7] * p0[0];
    P1 = 1; M1 = -1; mixed_coef *= 0.5;
  } else if (C1 == 0) {
    x1 = (c1[0] * p0[2] + c1[1] * p0[1]) + c1[2] * p0[0];
    P1 = 1; M1 = 0;
  } else {
    x1 = (c1[3] * p0[-2] + c1[4] * p0[-1]) + c1[5] * p0[0];
    P1 = 0; M1 = -1;
  }
  if (C2 != 0 && C2 != 3) {
    x2 = (c2[6] * p0[-4] + c2[8] * p0[4]) + c2[7] * p0[0];
    P2 = 4; M2 = -4; mixed_coef *= 0.5;
  } else if (C2 == 0) {
    x2 = (c2[0] * p0[8] + c2[1] * p0[4]) + c2[2] * p0[0];
    P2 = 4; M2 = 0;
  } else {
    x2 = (c2[3] * p0[-8] + c2[4] * p0[-4]) + c2[5] * p0[0];
    P2 = 0; M2 = -4;
  }
  const Real mixed = mixed_coef * d1 * d2 * ((p0[M1 + M2] + p0[P1 + P2]) - (p0[P1 + M2] + p0[M1 + P2]));
  return (x1 + x2) + mixed;
}

}  // namespace cup
