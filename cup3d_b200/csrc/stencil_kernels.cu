// stencil_kernels.cu -- the per-block stencil sweeps and the time-step drivers
// built on them (uniform-level meshes: every face neighbour is a same-level
// block or a domain wall).
//
// Reference (main.c): k_advdiff :4986 / advdiff :5027, k_prhs :5663, k_divp
// :5700, k_gradp :5715, k_lhs :4254, projection :5828, wall BC OP_BC :3524.
//
// Decomposition as in mg_device.cuh: one 8^3 block per 64 threads, thread
// (x, y) owns a z-line in registers, x/y neighbours through shared memory.
// Wall ghosts copy the nearest interior cell, with the wall-normal component of
// vector fields negated (free slip; gen_table.py:195-231).
#include <cmath>
#include <cstdlib>
#include <utility>

#include "blas_kernels.cuh"
#include "cup_internal.h"
#include "comm.cuh"
#include "advdiff_tma.cuh"
#include "amr_kernels.cuh"
#include "mg_device.cuh"

namespace cup {

// ---------------------------------------------------------------------------
// k_advdiff: TMP_c += fac_a * (u . grad) u_c + fac_d * lap u_c
// ---------------------------------------------------------------------------
// x / 60 correctly rounded without the division sequence: q0 = x*c, one FMA residual, one FMA
// correction (Markstein; c = RN(1/60) and 60's significand is not all ones, so the result equals
// IEEE x/60 -- checked exhaustively-at-random on 2e9 doubles against the division).
template <typename Real>
__device__ __forceinline__ Real div60(Real x) {
  const Real c = (Real)1 / (Real)60;
  const Real q0 = x * c;
  const Real r = fma(-q0, (Real)60, x);
  return fma(r, c, q0);
}

template <typename Real>
__device__ __forceinline__ Real upwind(Real U, Real um3, Real um2, Real um1, Real u, Real up1, Real up2, Real up3) {
  // derivative(), main.c:4980-4985: 5th-order upwind-biased, /60.  The U <= 0 branch is the exact
  // negation of the U > 0 expression on the mirrored operands (same operation order), so one
  // polynomial on selected operands gives both, bit for bit.
  const bool pos = U > 0;
  const Real a3 = pos ? um3 : up3, a2 = pos ? um2 : up2, a1 = pos ? um1 : up1;
  const Real b1 = pos ? up1 : um1, b2 = pos ? up2 : um2;
  const Real r = (((((Real)-2 * a3 + (Real)15 * a2) - (Real)60 * a1) + (Real)20 * u) + (Real)30 * b1 - (Real)3 * b2);
  const Real q = div60<Real>(r);
  return pos ? q : -q;
}

// padded tile [8 z][14 y][14 x], halo offset 3.  Row stride 24 and slab stride = 8 (mod 16) Reals keep
// the per-half-warp accesses (two y rows, or two z planes of a y halo) on disjoint banks.
enum { AD_ROW = 24, AD_SLAB = 14 * 24 + 8 };

template <typename Real, int MINB>
__global__ void __launch_bounds__(TPB, MINB) k_advdiff(LevelView lv, const int *__restrict__ sub, int nsub,
                                                       const Real *__restrict__ hblk, Real dtnu_dt, Real dtnu_nu,
                                                       const Real *__restrict__ v0, const Real *__restrict__ v1,
                                                       const Real *__restrict__ v2, Real *__restrict__ t0,
                                                       Real *__restrict__ t1, Real *__restrict__ t2, Real fac_a0,
                                                       Real fac_d0, Real ux, Real uy, Real uz) {
  __shared__ Real tile[8 * AD_SLAB];
  const int t = threadIdx.x, x = t & 7, y = t >> 3;
  const int a = t & 7, c2 = t >> 3;  // halo element coordinates
  const Real *vel[3] = {v0, v1, v2};
  Real *tmp[3] = {t0, t1, t2};
  const Real uinf[3] = {ux, uy, uz};
  const Real *rsl = lv.rslab ? rslab_of<Real>(lv) : nullptr;
  // ghost value of component c, layer l (counted from the face) of the slab received for code nbc
  // (element e of the 8x8 face: the thread's own for z and y faces, any for x faces)
  auto rem = [&](int nbc, int c, int l, int e) -> Real {
    return __ldcg(rsl + (size_t)(kRemote0 - nbc) * (64 * kSlabPlanes) + (c * 3 + l) * 64 + e);
  };
  for (int wi = blockIdx.x; wi < nsub; wi += gridDim.x) {
    const int b = sub ? sub[wi] : wi;
    const size_t own = (size_t)lv.act[b] * 512;
    const int *nbr6 = lv.nbr + (size_t)b * 6;
    int nb[6];
#pragma unroll
    for (int f = 0; f < 6; f++)
      nb[f] = nbr6[f];
    // per-block factors on multi-level meshes (main.c:4993-4995), else the level's
    Real fac_a = fac_a0, fac_d = fac_d0;
    if (hblk) {
      const Real hb = hblk[b], h3b = hb * hb * hb;
      fac_a = -dtnu_dt / hb * h3b;
      fac_d = (dtnu_nu / hb) * (dtnu_dt / hb) * h3b;
    }
    Real vv[3][8];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int k = 0; k < 8; k++)
        vv[c][k] = vel[c][own + k * 64 + t];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const Real *vc = vel[c];
      // z-extended line: ghost layers -3..-1 and 8..10
      Real line[14];
#pragma unroll
      for (int k = 0; k < 8; k++)
        line[3 + k] = vv[c][k];
      {
        const Real sg = (c == 2) ? (Real)-1 : (Real)1;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          line[i] = nb[4] >= 0 ? vc[(size_t)nb[4] * 512 + (5 + i) * 64 + t]
                               : (nb[4] == kWall ? sg * vv[c][0] : rem(nb[4], c, 2 - i, t));
          line[11 + i] = nb[5] >= 0 ? vc[(size_t)nb[5] * 512 + i * 64 + t]
                                    : (nb[5] == kWall ? sg * vv[c][7] : rem(nb[5], c, i, t));
        }
      }
      __syncthreads();  // previous component's tile fully consumed
#pragma unroll
      for (int k = 0; k < 8; k++)
        tile[k * AD_SLAB + (y + 3) * AD_ROW + (x + 3)] = vv[c][k];
      {
        const Real sx = (c == 0) ? (Real)-1 : (Real)1, sy = (c == 1) ? (Real)-1 : (Real)1;
        // -x / +x: 2 sides x 3 layers x 64 (y,z) = 6 elements per thread, layer fastest so that three
        // lanes share a sector and neighbouring lanes spread over the banks
#pragma unroll
        for (int j = 0; j < 6; j++) {
          const int r = (j % 3) * 64 + t, p = r % 3, yz = r / 3, yo = (yz & 7) * 8 + (yz >> 3) * 64;
          Real g;
          if (j < 3)
            g = nb[0] >= 0 ? vc[(size_t)nb[0] * 512 + yo + (5 + p)]
                           : (nb[0] == kWall ? sx * vc[own + yo] : rem(nb[0], c, 2 - p, yz));
          else
            g = nb[1] >= 0 ? vc[(size_t)nb[1] * 512 + yo + p]
                           : (nb[1] == kWall ? sx * vc[own + yo + 7] : rem(nb[1], c, p, yz));
          tile[(yz >> 3) * AD_SLAB + ((yz & 7) + 3) * AD_ROW + (j < 3 ? 0 : 11) + p] = g;
        }
#pragma unroll
        for (int p = 0; p < 3; p++) {
          // -y / +y: element (x = a, z = c2)
          const Real ym = nb[2] >= 0 ? vc[(size_t)nb[2] * 512 + c2 * 64 + (5 + p) * 8 + a]
                                     : (nb[2] == kWall ? sy * vc[own + c2 * 64 + a] : rem(nb[2], c, 2 - p, t));
          const Real yp = nb[3] >= 0 ? vc[(size_t)nb[3] * 512 + c2 * 64 + p * 8 + a]
                                     : (nb[3] == kWall ? sy * vc[own + c2 * 64 + 56 + a] : rem(nb[3], c, p, t));
          tile[c2 * AD_SLAB + p * AD_ROW + (a + 3)] = ym;
          tile[c2 * AD_SLAB + (11 + p) * AD_ROW + (a + 3)] = yp;
        }
      }
      __syncthreads();
      Real *oc = tmp[c];
      const int a1 = (c + 1) % 3, a2 = (c + 2) % 3;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const Real *row = tile + k * AD_SLAB + (y + 3) * AD_ROW + (x + 3);
        const Real u = line[3 + k];
        Real dd[3], pr[3];
        const Real U0 = vv[0][k] + uinf[0], U1 = vv[1][k] + uinf[1], U2 = vv[2][k] + uinf[2];
        dd[0] = upwind<Real>(U0, row[-3], row[-2], row[-1], u, row[1], row[2], row[3]);
        pr[0] = row[1] + row[-1];
        dd[1] = upwind<Real>(U1, row[-3 * AD_ROW], row[-2 * AD_ROW], row[-AD_ROW], u, row[AD_ROW], row[2 * AD_ROW],
                             row[3 * AD_ROW]);
        pr[1] = row[AD_ROW] + row[-AD_ROW];
        dd[2] = upwind<Real>(U2, line[k], line[k + 1], line[k + 2], u, line[k + 4], line[k + 5], line[k + 6]);
        pr[2] = line[k + 4] + line[k + 2];
        const Real Uabs[3] = {U0, U1, U2};
        const Real adv = Uabs[c] * dd[c] + (Uabs[a1] * dd[a1] + Uabs[a2] * dd[a2]);
        const Real lap = (pr[c] + (pr[a1] + pr[a2])) - (Real)6 * u;
        oc[own + k * 64 + t] += fac_a * adv + fac_d * lap;
      }
    }
    __syncthreads();
  }
}

// RK3 stage update (advdiff, main.c:5039-5054): V += TMP*alpha/h^3 ; TMP *= beta
template <typename Real>
__global__ void __launch_bounds__(256) k_rk_update(Real *__restrict__ v, Real *__restrict__ tm, long long n, Real ih3,
                                                   Real beta) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const Real tv = tm[i];
    v[i] += tv * ih3;
    tm[i] = tv * beta;
  }
}

// ---------------------------------------------------------------------------
// k_prhs: LHS = fac*div(u) - chi*fac*div(udef), fac = h^2/(2 dt)   (main.c:5663)
// Only the face-normal component is needed on each face: u on x faces, v on y
// faces, w on z faces (and the same of udef), so the halo is 6 components x 2
// faces instead of 6 x 6.
// ---------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(TPB, 8) k_prhs(LevelView lv, const Real *__restrict__ v0, const Real *__restrict__ v1,
                                              const Real *__restrict__ v2, const Real *__restrict__ d0,
                                              const Real *__restrict__ d1, const Real *__restrict__ d2,
                                              const Real *__restrict__ chi, Real *__restrict__ lhs, Real fac) {
  __shared__ Real tl[4][512];      // u, v, udef_x, udef_y cores
  __shared__ Real hl[4][2][64];    // their -/+ faces (x faces for u/udef_x, y faces for v/udef_y)
  const int t = threadIdx.x, x = t & 7, y = t >> 3;
  const int a = t & 7, c2 = t >> 3;
  const Real *rsl = lv.rslab ? rslab_of<Real>(lv) : nullptr;
  // component q (0..5 = u v w udef_x udef_y udef_z), single layer, of the slab received for code nbc
  auto rem = [&](int nbc, int q) -> Real {
    return __ldcg(rsl + (size_t)(kRemote0 - nbc) * (64 * kSlabPlanes) + q * 64 + t);
  };
  // the block index and neighbour list of the NEXT iteration are fetched one iteration ahead, so the
  // field loads of a block do not wait behind a dependent index load (ncu r01: 17.6 warps per issue
  // stalled on the long scoreboard, three dependent DRAM round trips per block)
  const int nwork = lv_count(lv);
  int wi = blockIdx.x, slot_n = 0, nn[6] = {0, 0, 0, 0, 0, 0};
  if (wi < nwork) {
    const int b = lv_item(lv, wi);
    slot_n = lv.act[b];
#pragma unroll
    for (int f = 0; f < 6; f++)
      nn[f] = lv.nbr[(size_t)b * 6 + f];
  }
  for (; wi < nwork; wi += gridDim.x) {
    const size_t own = (size_t)slot_n * 512;
    const int n0 = nn[0], n1 = nn[1], n2 = nn[2], n3 = nn[3], n4 = nn[4], n5 = nn[5];
    {
      const int wn = wi + gridDim.x;
      if (wn < nwork) {
        const int bn = lv_item(lv, wn);
        slot_n = lv.act[bn];
#pragma unroll
        for (int f = 0; f < 6; f++)
          nn[f] = lv.nbr[(size_t)bn * 6 + f];
      }
    }
    Real w[8], dz[8], ch[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      tl[0][k * 64 + t] = v0[own + k * 64 + t];
      tl[1][k * 64 + t] = v1[own + k * 64 + t];
      tl[2][k * 64 + t] = d0[own + k * 64 + t];
      tl[3][k * 64 + t] = d1[own + k * 64 + t];
      w[k] = v2[own + k * 64 + t];
      dz[k] = d2[own + k * 64 + t];
      ch[k] = chi[own + k * 64 + t];
    }
    // x faces of u and udef_x: element (y = a, z = c2); wall: -own (normal component flips)
    hl[0][0][t] = n0 >= 0 ? v0[(size_t)n0 * 512 + c2 * 64 + a * 8 + 7] : (n0 == kWall ? -v0[own + c2 * 64 + a * 8] : rem(n0, 0));
    hl[0][1][t] = n1 >= 0 ? v0[(size_t)n1 * 512 + c2 * 64 + a * 8] : (n1 == kWall ? -v0[own + c2 * 64 + a * 8 + 7] : rem(n1, 0));
    hl[2][0][t] = n0 >= 0 ? d0[(size_t)n0 * 512 + c2 * 64 + a * 8 + 7] : (n0 == kWall ? -d0[own + c2 * 64 + a * 8] : rem(n0, 3));
    hl[2][1][t] = n1 >= 0 ? d0[(size_t)n1 * 512 + c2 * 64 + a * 8] : (n1 == kWall ? -d0[own + c2 * 64 + a * 8 + 7] : rem(n1, 3));
    // y faces of v and udef_y: element (x = a, z = c2)
    hl[1][0][t] = n2 >= 0 ? v1[(size_t)n2 * 512 + c2 * 64 + 56 + a] : (n2 == kWall ? -v1[own + c2 * 64 + a] : rem(n2, 1));
    hl[1][1][t] = n3 >= 0 ? v1[(size_t)n3 * 512 + c2 * 64 + a] : (n3 == kWall ? -v1[own + c2 * 64 + 56 + a] : rem(n3, 1));
    hl[3][0][t] = n2 >= 0 ? d1[(size_t)n2 * 512 + c2 * 64 + 56 + a] : (n2 == kWall ? -d1[own + c2 * 64 + a] : rem(n2, 4));
    hl[3][1][t] = n3 >= 0 ? d1[(size_t)n3 * 512 + c2 * 64 + a] : (n3 == kWall ? -d1[own + c2 * 64 + 56 + a] : rem(n3, 4));
    // z faces of w and udef_z straight into registers
    const Real wzm = n4 >= 0 ? v2[(size_t)n4 * 512 + 448 + t] : (n4 == kWall ? -w[0] : rem(n4, 2));
    const Real wzp = n5 >= 0 ? v2[(size_t)n5 * 512 + t] : (n5 == kWall ? -w[7] : rem(n5, 2));
    const Real dzm = n4 >= 0 ? d2[(size_t)n4 * 512 + 448 + t] : (n4 == kWall ? -dz[0] : rem(n4, 5));
    const Real dzp = n5 >= 0 ? d2[(size_t)n5 * 512 + t] : (n5 == kWall ? -dz[7] : rem(n5, 5));
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int i = k * 64 + t;
      const Real uxp = x < 7 ? tl[0][i + 1] : hl[0][1][y + 8 * k];
      const Real uxm = x > 0 ? tl[0][i - 1] : hl[0][0][y + 8 * k];
      const Real vyp = y < 7 ? tl[1][i + 8] : hl[1][1][x + 8 * k];
      const Real vym = y > 0 ? tl[1][i - 8] : hl[1][0][x + 8 * k];
      const Real wzp_ = k < 7 ? w[k < 7 ? k + 1 : 7] : wzp;
      const Real wzm_ = k > 0 ? w[k > 0 ? k - 1 : 0] : wzm;
      const Real dxp = x < 7 ? tl[2][i + 1] : hl[2][1][y + 8 * k];
      const Real dxm = x > 0 ? tl[2][i - 1] : hl[2][0][y + 8 * k];
      const Real dyp = y < 7 ? tl[3][i + 8] : hl[3][1][x + 8 * k];
      const Real dym = y > 0 ? tl[3][i - 8] : hl[3][0][x + 8 * k];
      const Real dzp_ = k < 7 ? dz[k < 7 ? k + 1 : 7] : dzp;
      const Real dzm_ = k > 0 ? dz[k > 0 ? k - 1 : 0] : dzm;
      Real p = fac * (((((uxp - uxm) + vyp) - vym) + wzp_) - wzm_);
      const Real div_us = ((((dxp - dxm) + dyp) - dym) + dzp_) - dzm_;
      p += -ch[k] * fac * div_us;
      lhs[own + i] = p;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// k_divp (main.c:5700): TMP0 = h * (sum of 6 neighbours - 6 p); note the
// summation order differs from k_lhs.   k_gradp (:5715): TMP_a = fac*(p+ - p-).
// ---------------------------------------------------------------------------
template <typename Real, int WHAT>  // 0 divp, 1 gradp
__global__ void __launch_bounds__(TPB) k_pres(LevelView lv, const Real *__restrict__ p, Real *__restrict__ o0,
                                              Real *__restrict__ o1, Real *__restrict__ o2, Real fac) {
  __shared__ Real tu[512];
  __shared__ Real halo[6][64];
  const int t = threadIdx.x, x = t & 7, y = t >> 3;
  SlotVec<Real> pv{const_cast<Real *>(p), nullptr, 0x7fffffff};
  for (int wi = blockIdx.x; wi < lv_count(lv); wi += gridDim.x) {
    const int b = lv_item(lv, wi);
    const size_t own = (size_t)lv.act[b] * 512;
    Real uu[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      uu[k] = p[own + k * 64 + t];
      tu[k * 64 + t] = uu[k];
    }
    load_halo<Real>(pv, p + own, lv.nbr + (size_t)b * 6, t, halo, rface_of<Real>(lv));
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int i = k * 64 + t;
      const Real xm = x > 0 ? tu[i - 1] : halo[0][y + 8 * k];
      const Real xp = x < 7 ? tu[i + 1] : halo[1][y + 8 * k];
      const Real ym = y > 0 ? tu[i - 8] : halo[2][x + 8 * k];
      const Real yp = y < 7 ? tu[i + 8] : halo[3][x + 8 * k];
      const Real zm = k > 0 ? uu[k > 0 ? k - 1 : 0] : halo[4][t];
      const Real zp = k < 7 ? uu[k < 7 ? k + 1 : 7] : halo[5][t];
      if (WHAT == 0) {
        o0[own + i] = fac * ((((((xp + xm) + yp) + ym) + zp) + zm) - (Real)6.0 * uu[k]);
      } else {
        o0[own + i] = fac * (xp - xm);
        o1[own + i] = fac * (yp - ym);
        o2[own + i] = fac * (zp - zm);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// k_vort (main.c:5736): TMP_a = h^2/2 * ((d_b u_c) - (d_c u_b)) central differences, b=(a+1)%3,
// c=(a+2)%3 (vorticity() then scales by 1/h^3, :5786);  k_q (:5762): Q-criterion
// q = -1/2 sum_ab g_ab g_ba, g_ab = (u_a(+b) - u_a(-b)) / 2h  -> F_LHS.
// All three velocity components are needed on all six faces.
// ---------------------------------------------------------------------------
template <typename Real, int WHAT>  // 0 vorticity, 1 Q
__global__ void __launch_bounds__(TPB) k_velgrad(LevelView lv, const Real *__restrict__ v0,
                                                 const Real *__restrict__ v1, const Real *__restrict__ v2,
                                                 Real *__restrict__ o0, Real *__restrict__ o1, Real *__restrict__ o2,
                                                 Real h) {
  __shared__ Real tl[3][512];
  __shared__ Real hl[3][6][64];
  const int t = threadIdx.x, x = t & 7, y = t >> 3, a = t & 7, c2 = t >> 3;
  const Real *vel[3] = {v0, v1, v2};
  const Real *rsl = lv.rslab ? rslab_of<Real>(lv) : nullptr;
  for (int wi = blockIdx.x; wi < lv_count(lv); wi += gridDim.x) {
    const int b = lv_item(lv, wi);
    const size_t own = (size_t)lv.act[b] * 512;
    const int *nbr6 = lv.nbr + (size_t)b * 6;
#pragma unroll
    for (int q = 0; q < 3; q++) {
#pragma unroll
      for (int k = 0; k < 8; k++)
        tl[q][k * 64 + t] = vel[q][own + k * 64 + t];
#pragma unroll
      for (int f = 0; f < 6; f++) {
        const int nb = nbr6[f];
        const int p = (nb >= 0) ? ((f & 1) ? 0 : 7) : ((f & 1) ? 7 : 0);
        const int idx = f < 2 ? (c2 << 6) + (a << 3) + p : (f < 4 ? (c2 << 6) + (p << 3) + a : (p << 6) + t);
        // wall: nearest interior cell, wall-normal component negated (OP_BC, vflip = 0)
        // face owned by another rank: component q, single layer, of the received slab
        hl[q][f][t] = nb >= 0 ? vel[q][(size_t)nb * 512 + idx]
                              : (nb == kWall ? ((f >> 1) == q ? -vel[q][own + idx] : vel[q][own + idx])
                                             : __ldcg(rsl + (size_t)(kRemote0 - nb) * (64 * kSlabPlanes) + q * 64 + t));
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int i = k * 64 + t;
      // g[q][d] = u_q(+d) - u_q(-d)
      Real g[3][3];
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const Real xp = x < 7 ? tl[q][i + 1] : hl[q][1][y + 8 * k], xm = x > 0 ? tl[q][i - 1] : hl[q][0][y + 8 * k];
        const Real yp = y < 7 ? tl[q][i + 8] : hl[q][3][x + 8 * k], ym = y > 0 ? tl[q][i - 8] : hl[q][2][x + 8 * k];
        const Real zp = k < 7 ? tl[q][i + 64] : hl[q][5][t], zm = k > 0 ? tl[q][i - 64] : hl[q][4][t];
        g[q][0] = xp - xm;
        g[q][1] = yp - ym;
        g[q][2] = zp - zm;
      }
      if (WHAT == 0) {
        const Real inv2h = (Real).5 * h * h;
        // o_a = inv2h * ((LS(b,1,c) - LS(b,-1,c)) - (LS(c,1,b) - LS(c,-1,b))): LS(axis, k, comp)
        o0[own + i] = inv2h * (g[2][1] - g[1][2]);
        o1[own + i] = inv2h * (g[0][2] - g[2][0]);
        o2[own + i] = inv2h * (g[1][0] - g[0][1]);
      } else {
        const Real inv2h = (Real).5 / h;
        Real gg[3][3];
#pragma unroll
        for (int q = 0; q < 3; q++)
#pragma unroll
          for (int d = 0; d < 3; d++)
            gg[q][d] = inv2h * g[q][d];
        Real qq = 0;
#pragma unroll
        for (int q = 0; q < 3; q++)
#pragma unroll
          for (int d = 0; d < 3; d++)
            qq -= (Real)0.5 * gg[q][d] * gg[d][q];
        o0[own + i] = qq;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// pointwise pieces of projection() (main.c:5841-5916)
// ---------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(256) k_lhs_minus(Real *__restrict__ lhs, const Real *__restrict__ b,
                                                   Real *__restrict__ p, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    lhs[i] -= b[i];
    p[i] = 0;
  }
}

// p -= avg ; p += p_old (optional), avg = scal[0]/scal[1]
template <typename Real>
__global__ void __launch_bounds__(256) k_pres_fix(Real *__restrict__ p, const Real *__restrict__ pold, long long n,
                                                  const double *__restrict__ scal, double vol) {
  const Real avg = (Real)(scal[0] / vol);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    Real v = p[i] - avg;
    if (pold)
      v += pold[i];
    p[i] = v;
  }
}

template <typename Real>
__global__ void __launch_bounds__(256) k_vel_add(Real *__restrict__ v, const Real *__restrict__ g, long long n,
                                                 Real fac) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    v[i] += fac * g[i];
}

// sum_i p_i * h_i^3 over leaves -> out (projection's avg, :5873-5882)
template <typename Real>
__global__ void __launch_bounds__(256) k_wsum_blk(const Real *__restrict__ a, const Real *__restrict__ h3,
                                                  long long nblk, double *out) {
  __shared__ double red[8];
  double s = 0;
  for (long long b = blockIdx.x; b < nblk; b += gridDim.x) {
    const Real *p = a + b * 512;
    s += ((double)p[threadIdx.x] + (double)p[threadIdx.x + 256]) * (double)h3[b];
  }
  for (int o = 16; o > 0; o >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0;
    for (int i = 0; i < 8; i++)
      tot += red[i];
    atomicAdd(out, tot);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
namespace {

inline int sgrid(const CupCtx *c, long long n) {
  long long g = (n + 255) / 256, cap = (long long)c->num_sms * 8;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

inline int bgrid(const CupCtx *c, long long nb, int per_sm) {
  long long g = (long long)c->num_sms * per_sm;
  return (int)(g < nb ? g : (nb < 1 ? 1 : nb));
}

// CUP_ADV_IMPL=ldg selects the plain-load k_advdiff (diagnostics); default: TMA-staged ghosts
bool adv_tma() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("CUP_ADV_IMPL");
    v = !(e && !strcmp(e, "ldg"));
  }
  return v != 0;
}

// k_prhs: TMA-staged (prhs_tma.cu) by default; CUP_PRHS_IMPL=ldg selects the plain-load kernel
bool prhs_tma() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("CUP_PRHS_IMPL");
    v = !(e && !strcmp(e, "ldg"));
  }
  return v != 0;
}

const Level *leaf_level(CupCtx *c) {
  int top = c->top;
  while (top > 0 && c->lv[top].gnact == 0)
    top--;
  return &c->lv[top];
}


// RK3 stage update with per-block h (multi-level meshes): V += TMP*alpha/h^3 ; TMP *= beta
template <typename Real>
__global__ void __launch_bounds__(256) k_rk_update_blk(Real *__restrict__ v, Real *__restrict__ tm,
                                                       const Real *__restrict__ h3, long long nblk, Real alpha,
                                                       Real beta) {
  for (long long b = blockIdx.x; b < nblk; b += gridDim.x) {
    const Real ih3 = alpha / h3[b];
    for (int j = threadIdx.x; j < 512; j += blockDim.x) {
      const Real tv = tm[b * 512 + j];
      v[b * 512 + j] += tv * ih3;
      tm[b * 512 + j] = tv * beta;
    }
  }
}

// v += g / h^3 per block (projection's velocity update on multi-level meshes, :5907-5916)
template <typename Real>
__global__ void __launch_bounds__(256) k_vel_add_blk(Real *__restrict__ v, const Real *__restrict__ g,
                                                     const Real *__restrict__ h3, long long nblk) {
  for (long long b = blockIdx.x; b < nblk; b += gridDim.x) {
    const Real fac = (Real)1.0 / h3[b];
    for (int j = threadIdx.x; j < 512; j += blockDim.x)
      v[b * 512 + j] += fac * g[b * 512 + j];
  }
}

template <typename Real>
int stencil_amr_t(CupCtx *c, CupStencilId id) {
  const Level &v = c->leafv;
  LevelView lv{v.d_act, v.d_nbr, (int)v.act.size(), nullptr, nullptr, 0, v.d_ext};
  lv.sub = c->run_sub;
  lv.nsub = c->run_nsub;
  Real **S = (Real **)c->state;
  const double dt = c->prm.dt;
  if (c->nranks > 1) {
    // halo_sync (main.c:3632): the input fields of the leaves other ranks own, as ghost blocks behind the
    // own ones (each state component holds nstate = nblk + ghost blocks)
    int fl[BLK_COMPS], nf = 0;
    switch (id) {
    case CUP_ST_ADVDIFF:
    case CUP_ST_VORT:
    case CUP_ST_Q:
      for (int q = 0; q < 3; q++)
        fl[nf++] = CUP_F_VEL + q;
      break;
    case CUP_ST_PRHS:
      for (int q = 0; q < 3; q++)
        fl[nf++] = CUP_F_VEL + q;
      for (int q = 0; q < 3; q++)
        fl[nf++] = CUP_F_TMP + q;
      fl[nf++] = CUP_F_CHI;
      break;
    case CUP_ST_DIVP:
    case CUP_ST_GRADP:
      fl[nf++] = CUP_F_PRES;
      break;
    default:
      break;
    }
    const Real *src[BLK_COMPS];
    Real *dst[BLK_COMPS];
    for (int q = 0; q < nf; q++)
      src[q] = dst[q] = S[fl[q]];
    if (nf > 0)
      CUP_TRY(block_exchange_leaf<Real>(c, src, dst, nf, 0));
  }
  switch (id) {
  case CUP_ST_ADVDIFF: {
    // blocks whose neighbours are all same-level / wall: the uniform kernel with per-block factors;
    // interface blocks: the ss = 3 coarse-fine ghost fill.  With a caller's list: its two halves.
    const int *d_reg = v.d_reg, *d_irr = v.d_irr;
    int nreg = (int)v.reg.size(), nirr = (int)v.irr.size();
    if (c->run_nsub >= 0) {
      std::vector<char> is_irr(v.act.size(), 0);
      for (int k : v.irr)
        is_irr[(size_t)k] = 1;
      std::vector<int> lr, li;
      for (int b : c->run_list)
        (is_irr[(size_t)b] ? li : lr).push_back(b);
      int *d2 = c->d_list + c->nblk;  // second half of the list scratch
      if (!lr.empty())
        CUP_CUDA(cudaMemcpyAsync(d2, lr.data(), lr.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
      if (!li.empty())
        CUP_CUDA(cudaMemcpyAsync(d2 + lr.size(), li.data(), li.size() * sizeof(int), cudaMemcpyHostToDevice,
                                 c->stream));
      CUP_CUDA(cudaStreamSynchronize(c->stream));  // lr / li are stack vectors
      d_reg = d2;
      d_irr = d2 + lr.size();
      nreg = (int)lr.size();
      nirr = (int)li.size();
    }
    LevelView lv0 = lv;  // these kernels take their list as arguments
    lv0.sub = nullptr;
    lv0.nsub = -1;
    if (nreg > 0 && adv_tma()) {
      CUP_TRY(advdiff_tma_launch<Real>(c, lv0, d_reg, nreg, v.d_hblk, dt, c->prm.nu, 0.0, 0.0));
    } else if (nreg > 0) {
      k_advdiff<Real, 8><<<bgrid(c, (long long)nreg, 8), TPB, 0, c->stream>>>(
          lv0, d_reg, nreg, (const Real *)v.d_hblk, (Real)dt, (Real)c->prm.nu, S[CUP_F_VEL],
          S[CUP_F_VEL + 1], S[CUP_F_VEL + 2], S[CUP_F_TMP], S[CUP_F_TMP + 1], S[CUP_F_TMP + 2], (Real)0, (Real)0,
          (Real)c->prm.uinf[0], (Real)c->prm.uinf[1], (Real)c->prm.uinf[2]);
      c->launches++;
    }
    if (nirr > 0)
      CUP_TRY(advdiff_amr_launch<Real>(c, v, S, d_irr, nirr));
    break;
  }
  case CUP_ST_PRHS:
    CUP_TRY(prhs_amr_launch<Real>(c, lv, v.d_hblk, S, (Real)(1.0 / dt)));
    break;
  case CUP_ST_DIVP:
    CUP_TRY(pres_amr_launch<Real>(c, lv, v.d_hblk, S[CUP_F_PRES], S[CUP_F_TMP], nullptr, nullptr, (Real)0, 0));
    break;
  case CUP_ST_GRADP:
    CUP_TRY(pres_amr_launch<Real>(c, lv, v.d_hblk, S[CUP_F_PRES], S[CUP_F_TMP], S[CUP_F_TMP + 1], S[CUP_F_TMP + 2],
                                  (Real)(-0.5 * dt), 1));
    break;
  case CUP_ST_VORT:
    CUP_TRY(velgrad_amr_launch<Real>(c, lv, v.d_hblk, S, 0));
    break;
  case CUP_ST_Q:
    CUP_TRY(velgrad_amr_launch<Real>(c, lv, v.d_hblk, S, 1));
    break;
  default:
    set_error("stencil %d on a multi-level mesh is not available in this build", (int)id);
    return CUP_ERR_UNSUPPORTED;
  }
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

template <typename Real>
int stencil_t(CupCtx *c, CupStencilId id, const int *d_sub, long long nsub) {
  if (!c->leaf_uniform && id != CUP_ST_LHS && id != CUP_ST_MG)
    return stencil_amr_t<Real>(c, id);
  const Level &v = *leaf_level(c);
  LevelView lv{v.d_act, v.d_nbr, (int)v.act.size(), v.d_frecv, (const unsigned long long *)v.d_seq,
               v.rface_stride, v.d_ext, v.d_srecv, v.slab_stride};
  if (c->nranks > 1) {
    // ghost data owned by other ranks (halo_sync, main.c:3632)
    Level &vm = const_cast<Level &>(v);
    if (id == CUP_ST_ADVDIFF) {
      SlabSrc<Real> src{{(const Real *)c->state[CUP_F_VEL], (const Real *)c->state[CUP_F_VEL + 1],
                         (const Real *)c->state[CUP_F_VEL + 2], nullptr, nullptr, nullptr}};
      CUP_TRY(slab_exchange<Real>(c, vm, src, 3, 3));
    } else if (id == CUP_ST_PRHS) {
      SlabSrc<Real> src{{(const Real *)c->state[CUP_F_VEL], (const Real *)c->state[CUP_F_VEL + 1],
                         (const Real *)c->state[CUP_F_VEL + 2], (const Real *)c->state[CUP_F_TMP],
                         (const Real *)c->state[CUP_F_TMP + 1], (const Real *)c->state[CUP_F_TMP + 2]}};
      CUP_TRY(slab_exchange<Real>(c, vm, src, 6, 1));
    } else if (id == CUP_ST_VORT || id == CUP_ST_Q) {
      SlabSrc<Real> src{{(const Real *)c->state[CUP_F_VEL], (const Real *)c->state[CUP_F_VEL + 1],
                         (const Real *)c->state[CUP_F_VEL + 2], nullptr, nullptr, nullptr}};
      CUP_TRY(slab_exchange<Real>(c, vm, src, 3, 1));
    } else if (id == CUP_ST_DIVP || id == CUP_ST_GRADP) {
      SlotVec<Real> pv{(Real *)c->state[CUP_F_PRES], nullptr, (int)c->nblk};
      CUP_TRY(halo_exchange<Real>(c, vm, pv));
    }
  }
  (void)d_sub;
  (void)nsub;
  const bool listed = c->run_nsub >= 0;
  lv.sub = c->run_sub;
  lv.nsub = c->run_nsub;
  Real **S = (Real **)c->state;
  const Real h = (Real)v.h;
  const double dt = c->prm.dt, hd = v.h;
  switch (id) {
  case CUP_ST_ADVDIFF: {
    // fac_a = -dt/h*h^3 ; fac_d = (nu/h)*(dt/h)*h^3   (main.c:4993-4995, coef = 1)
    const double h3 = hd * hd * hd;
    const double fa = -dt / hd * h3 * 1.0, fd = (c->prm.nu / hd) * (dt / hd) * h3 * 1.0;
    if (adv_tma()) {
      LevelView lv0 = lv;  // takes its list as arguments
      lv0.sub = nullptr;
      lv0.nsub = -1;
      return advdiff_tma_launch<Real>(c, lv0, c->run_sub, listed ? c->run_nsub : lv.nact, nullptr, 0.0, 0.0, fa, fd,
                                      (const AdvRk *)c->adv_rk);
    }
    static int minb = getenv("CUP_ADV_MINB") ? atoi(getenv("CUP_ADV_MINB")) : 8;
#define ADV_LAUNCH(M)                                                                                              \
  k_advdiff<Real, M><<<bgrid(c, c->nblk, M), TPB, 0, c->stream>>>(                                                 \
      lv, c->run_sub, listed ? c->run_nsub : lv.nact, nullptr, (Real)0, (Real)0, S[CUP_F_VEL], S[CUP_F_VEL + 1],   \
      S[CUP_F_VEL + 2],                                                                                            \
      S[CUP_F_TMP], S[CUP_F_TMP + 1], S[CUP_F_TMP + 2], (Real)fa, (Real)fd, (Real)c->prm.uinf[0],                 \
      (Real)c->prm.uinf[1], (Real)c->prm.uinf[2])
    if (minb == 5)
      ADV_LAUNCH(5);
    else if (minb == 6)
      ADV_LAUNCH(6);
    else if (minb == 10)
      ADV_LAUNCH(10);
    else if (minb == 12)
      ADV_LAUNCH(12);
    else
      ADV_LAUNCH(8);
#undef ADV_LAUNCH
    break;
  }
  case CUP_ST_PRHS: {
    const double fac = 0.5 * hd * hd / dt;
    if (prhs_tma() && !listed)
      return prhs_tma_launch<Real>(c, lv, fac);
    static bool carve = false;  // 8 CTAs x 21.5 KB need the large shared-memory configuration
    if (!carve) {
      carve = true;
      cudaFuncSetAttribute(k_prhs<Real>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
    }
    k_prhs<Real><<<bgrid(c, c->nblk, 8), TPB, 0, c->stream>>>(lv, S[CUP_F_VEL], S[CUP_F_VEL + 1], S[CUP_F_VEL + 2],
                                                              S[CUP_F_TMP], S[CUP_F_TMP + 1], S[CUP_F_TMP + 2],
                                                              S[CUP_F_CHI], S[CUP_F_LHS], (Real)fac);
    break;
  }
  case CUP_ST_DIVP:
    k_pres<Real, 0><<<bgrid(c, c->nblk, 12), TPB, 0, c->stream>>>(lv, S[CUP_F_PRES], S[CUP_F_TMP], nullptr, nullptr,
                                                                  h);
    break;
  case CUP_ST_GRADP: {
    const double fac = -0.5 * dt * hd * hd;
    k_pres<Real, 1><<<bgrid(c, c->nblk, 12), TPB, 0, c->stream>>>(lv, S[CUP_F_PRES], S[CUP_F_TMP], S[CUP_F_TMP + 1],
                                                                  S[CUP_F_TMP + 2], (Real)fac);
    break;
  }
  case CUP_ST_VORT:
    k_velgrad<Real, 0><<<bgrid(c, c->nblk, 8), TPB, 0, c->stream>>>(lv, S[CUP_F_VEL], S[CUP_F_VEL + 1],
                                                                    S[CUP_F_VEL + 2], S[CUP_F_TMP], S[CUP_F_TMP + 1],
                                                                    S[CUP_F_TMP + 2], h);
    break;
  case CUP_ST_Q:
    k_velgrad<Real, 1><<<bgrid(c, c->nblk, 8), TPB, 0, c->stream>>>(lv, S[CUP_F_VEL], S[CUP_F_VEL + 1],
                                                                    S[CUP_F_VEL + 2], S[CUP_F_LHS], nullptr, nullptr,
                                                                    h);
    break;
  case CUP_ST_LHS:
  case CUP_ST_MG: {
    // k_lhs / k_mg on the state: F_PRES -> F_LHS, no mean term (that is pois_op's); st_mg has no
    // flux faces (outc = 0, main.c:4280)
    const int mc = c->prm.mean_constraint;
    c->prm.mean_constraint = 0;
    c->no_flux_correction = (id == CUP_ST_MG);
    int rc = pois_op_dev(c, S[CUP_F_PRES], S[CUP_F_LHS]);
    c->no_flux_correction = false;
    c->prm.mean_constraint = mc;
    return rc;
  }
  default:
    set_error("stencil_run: unknown stencil id %d", (int)id);
    return CUP_ERR_ARG;
  }
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

template <typename Real>
int advdiff_t(CupCtx *c) {
  // 3-stage low-storage RK (Williamson), main.c:5027-5056
  const double alpha[3] = {1.0 / 3.0, 15.0 / 16.0, 8.0 / 15.0};
  const double beta[3] = {-5.0 / 9.0, -153.0 / 128.0, 0.0};
  const Level &v = *leaf_level(c);
  const long long N = c->nblk * 512;
  const size_t rb = sizeof(Real);
  Real **S = (Real **)c->state;
  for (int q = 0; q < 3; q++)
    CUP_CUDA(cudaMemsetAsync(S[CUP_F_TMP + q], 0, N * rb, c->stream));
  static const bool fuse = !(getenv("CUP_ADV_FUSE_RK") && atoi(getenv("CUP_ADV_FUSE_RK")) == 0);
  if (c->leaf_uniform && adv_tma() && fuse) {
    // Uniform meshes: the stage update V += T alpha/h^3, T *= beta rides in the sweep that produced T
    // (12 Reals per cell and stage instead of 9 + 12, three launches instead of twelve).  The sweep reads
    // the neighbours' OLD velocity, so the new one goes to a second buffer and the two swap roles.
    const size_t bytes = (size_t)c->nstate * 512 * rb;
    for (int q = 0; q < 3; q++)
      if (!c->vel_spare[q])
        CUP_CUDA(cudaMalloc(&c->vel_spare[q], bytes));
    for (int s = 0; s < 3; s++) {
      AdvRk rk;
      for (int q = 0; q < 3; q++)
        rk.vout[q] = c->vel_spare[q];
      rk.ih3 = alpha[s] / (v.h * v.h * v.h);
      rk.beta = beta[s];
      c->adv_rk = &rk;
      const int rc = stencil_t<Real>(c, CUP_ST_ADVDIFF, nullptr, 0);
      c->adv_rk = nullptr;
      CUP_TRY(rc);
      for (int q = 0; q < 3; q++)
        std::swap(c->state[CUP_F_VEL + q], c->vel_spare[q]);  // cup_state_dev(F_VEL..) follows
    }
    CUP_CUDA(cudaGetLastError());
    return CUP_OK;
  }
  for (int s = 0; s < 3; s++) {
    CUP_TRY(stencil_t<Real>(c, CUP_ST_ADVDIFF, nullptr, 0));
    const double ih3 = alpha[s] / (v.h * v.h * v.h);
    for (int q = 0; q < 3; q++) {
      if (c->leaf_uniform)
        k_rk_update<Real><<<sgrid(c, N), 256, 0, c->stream>>>(S[CUP_F_VEL + q], S[CUP_F_TMP + q], N, (Real)ih3,
                                                               (Real)beta[s]);
      else
        k_rk_update_blk<Real><<<bgrid(c, c->nblk, 8), 256, 0, c->stream>>>(
            S[CUP_F_VEL + q], S[CUP_F_TMP + q], (const Real *)c->d_hw, c->nblk, (Real)alpha[s], (Real)beta[s]);
      c->launches++;
    }
  }
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

template <typename Real>
int projection_t(CupCtx *c, CupSolveInfo *info) {
  const Level &v = *leaf_level(c);
  const long long N = c->nblk * 512;
  const size_t rb = sizeof(Real);
  Real **S = (Real **)c->state;
  enum { STEP_2ND = 2 };  // main.c:130
  // p_old = p ; TMP = 0 ; fish_tmpv adds udef into TMP (main.c:5842-5847).  keep_tmp_udef: the
  // caller uploaded F_TMP already containing it.
  CUP_CUDA(cudaMemcpyAsync(c->p_old, S[CUP_F_PRES], N * rb, cudaMemcpyDeviceToDevice, c->stream));
  if (!c->keep_tmp_udef) {
    for (int q = 0; q < 3; q++)
      CUP_CUDA(cudaMemsetAsync(S[CUP_F_TMP + q], 0, N * rb, c->stream));
    if (obstacle_count(c) > 0)
      CUP_TRY(obstacle_tmpv(c));
  }
  CUP_TRY(stencil_t<Real>(c, CUP_ST_PRHS, nullptr, 0));
  if (c->prm.step > STEP_2ND) {
    CUP_TRY(stencil_t<Real>(c, CUP_ST_DIVP, nullptr, 0));
    k_lhs_minus<Real><<<sgrid(c, N), 256, 0, c->stream>>>(S[CUP_F_LHS], S[CUP_F_TMP], S[CUP_F_PRES], N);
    c->launches++;
  } else {
    CUP_CUDA(cudaMemsetAsync(S[CUP_F_PRES], 0, N * rb, c->stream));
  }
  CUP_TRY(pois_solve(c, info));
  // subtract the volume-weighted mean (main.c:5871-5895)
  const double vol = c->gvol;
  double *q = c->d_scal + 5;
  CUP_CUDA(cudaMemsetAsync(q, 0, sizeof(double), c->stream));
  k_wsum_blk<Real><<<bgrid(c, c->nblk, 8), 256, 0, c->stream>>>(S[CUP_F_PRES], (const Real *)c->d_hw, c->nblk, q);
  c->launches++;
  CUP_TRY(comm_allreduce(c, 5, 1));
  k_pres_fix<Real><<<sgrid(c, N), 256, 0, c->stream>>>(S[CUP_F_PRES],
                                                        c->prm.step > STEP_2ND ? (const Real *)c->p_old : nullptr, N, q,
                                                        vol);
  c->launches++;
  CUP_TRY(stencil_t<Real>(c, CUP_ST_GRADP, nullptr, 0));
  const double fac = 1.0 / (v.h * v.h * v.h);
  for (int a = 0; a < 3; a++) {
    if (c->leaf_uniform)
      k_vel_add<Real><<<sgrid(c, N), 256, 0, c->stream>>>(S[CUP_F_VEL + a], S[CUP_F_TMP + a], N, (Real)fac);
    else
      k_vel_add_blk<Real><<<bgrid(c, c->nblk, 8), 256, 0, c->stream>>>(S[CUP_F_VEL + a], S[CUP_F_TMP + a],
                                                                      (const Real *)c->d_hw, c->nblk);
    c->launches++;
  }
  CUP_CUDA(cudaGetLastError());
  CUP_CUDA(cudaStreamSynchronize(c->stream));
  return CUP_OK;
}

}  // namespace

// stencil_run(st, list, n), main.c:3631-3647: the kernel runs on blocks list[0..n) -- or on the
// first n blocks when list is NULL -- after the ghost exchange of the whole field.  Blocks not
// listed keep their output untouched.  The flux correction is applied inside the pass of the
// listed coarse blocks, i.e. for exactly the blocks whose output is recomputed.
int stencil_run(CupCtx *c, CupStencilId id, const long long *list, long long n) {
  if (c->nblk == 0) {
    set_error("stencil_run: no mesh uploaded");
    return CUP_ERR_STATE;
  }
  if (n < 0 || n > c->nblk) {
    set_error("stencil_run: n = %lld outside [0, %lld]", n, c->nblk);
    return CUP_ERR_ARG;
  }
  if ((int)id < 0 || (int)id > (int)CUP_ST_GRADCHI) {
    set_error("stencil_run: unknown stencil id %d", (int)id);
    return CUP_ERR_ARG;
  }
  c->run_sub = nullptr;
  c->run_nsub = -1;
  if (list != nullptr || n != c->nblk) {
    c->run_list.resize((size_t)n);
    std::vector<char> seen((size_t)c->nblk, 0);
    for (long long k = 0; k < n; k++) {
      const long long b = list ? list[k] : k;
      if (b < 0 || b >= c->nblk || seen[(size_t)b]) {
        set_error("stencil_run: list[%lld] = %lld is %s", k, b, (b < 0 || b >= c->nblk) ? "out of range" : "a duplicate");
        return CUP_ERR_ARG;
      }
      seen[(size_t)b] = 1;
      c->run_list[(size_t)k] = (int)b;
    }
    if (n == 0)
      return CUP_OK;  // the reference would still exchange ghosts; no output changes
    if (!c->d_list)
      CUP_CUDA(cudaMalloc((void **)&c->d_list, (size_t)c->nblk * 2 * sizeof(int)));
    CUP_CUDA(cudaMemcpyAsync(c->d_list, c->run_list.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice,
                             c->stream));
    CUP_CUDA(cudaStreamSynchronize(c->stream));
    c->run_sub = c->d_list;
    c->run_nsub = (int)n;
  }
  const int rc = id == CUP_ST_GRADCHI
                     ? gradchi(c)
                     : (c->real_bytes == 8 ? stencil_t<double>(c, id, nullptr, n) : stencil_t<float>(c, id, nullptr, n));
  c->run_sub = nullptr;
  c->run_nsub = -1;
  return rc;
}

int vorticity(CupCtx *c) {
  CUP_TRY(stencil_run(c, CUP_ST_VORT, nullptr, c->nblk));
  return scale_blk3(c, c->state[CUP_F_TMP], c->state[CUP_F_TMP + 1], c->state[CUP_F_TMP + 2]);
}

int advdiff(CupCtx *c) {
  if (c->nblk == 0) {
    set_error("advdiff: no mesh uploaded");
    return CUP_ERR_STATE;
  }
  return c->real_bytes == 8 ? advdiff_t<double>(c) : advdiff_t<float>(c);
}

int projection(CupCtx *c, CupSolveInfo *info) {
  if (c->nblk == 0) {
    set_error("projection: no mesh uploaded");
    return CUP_ERR_STATE;
  }
  return c->real_bytes == 8 ? projection_t<double>(c, info) : projection_t<float>(c, info);
}

}  // namespace cup
