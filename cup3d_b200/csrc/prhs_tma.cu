// prhs_tma.cu -- k_prhs (main.c:5663) with all seven fields and their ghost faces staged by the
// TMA engine.  The plain-load kernel (stencil_kernels.cu:k_prhs) stays latency bound even with its
// index prefetch (ncu r01: 10 warps per issue on the long scoreboard, DRAM 51 %): every thread has
// 56 field loads to wait for before it can do 30 flops per cell.  Here one mbarrier completes when
// the copy engine has delivered the block's u, v, w, udef_x/y/z, chi (1-D bulk copies) and the one
// ghost layer each face-normal component needs: x faces of u / udef_x as tensor boxes {16 B,8,8},
// y faces of v / udef_y as boxes {8,1,8}, z faces of w / udef_z as 64-element bulk copies.  The CTA
// computes straight from the stage, stores LHS, and the producer thread requests the next block;
// six resident CTAs per SM cover each other's fetch latency (the smoother's scheme).
// Walls (ghost = -own, the normal component flips) and faces received from other ranks are
// written into the stage by the threads.  Arithmetic identical to k_prhs.
#include "advdiff_tma.cuh"
#include "cup_internal.h"
#include "stencil7_tma.cuh"
#include "tma.cuh"

namespace cup {

namespace {

template <typename Real>
struct alignas(128) PrhsStage {
  static constexpr int NCOL = 16 / (int)sizeof(Real);
  alignas(128) Real f[7][512];             // u v w udef_x udef_y udef_z chi
  alignas(128) Real xf[2][2][64 * NCOL];   // [u, udef_x][-x, +x]
  alignas(128) Real yf[2][2][64];          // [v, udef_y][-y, +y]
  alignas(128) Real zf[2][2][64];          // [w, udef_z][-z, +z]
  alignas(8) uint64_t bar;
};

struct PrhsMaps {
  CUtensorMap xu, xd, yv, yd;  // x-face maps of u, udef_x; y-face maps of v, udef_y
};

template <typename Real>
struct PrhsArgs {
  const Real *fld[7];
  Real *lhs;
  Real fac;
};

template <typename Real>
__device__ __forceinline__ void prhs_issue(PrhsStage<Real> &s, const PrhsArgs<Real> &A, const PrhsMaps &M, int slot,
                                           const int (&nb)[6]) {
  constexpr int NCOL = PrhsStage<Real>::NCOL;
  constexpr uint32_t RB = sizeof(Real);
  uint32_t bytes = 7 * 512 * RB;
  bytes += ((nb[0] >= 0) + (nb[1] >= 0)) * 2 * 64 * NCOL * RB;
  bytes += ((nb[2] >= 0) + (nb[3] >= 0) + (nb[4] >= 0) + (nb[5] >= 0)) * 2 * 64 * RB;
  mbar_arrive_expect_tx(&s.bar, bytes);
#pragma unroll
  for (int q = 0; q < 7; q++)
    tma_load_1d(s.f[q], A.fld[q] + (size_t)slot * 512, 512 * RB, &s.bar);
  if (nb[4] >= 0) {
    tma_load_1d(s.zf[0][0], A.fld[2] + (size_t)nb[4] * 512 + 448, 64 * RB, &s.bar);
    tma_load_1d(s.zf[1][0], A.fld[5] + (size_t)nb[4] * 512 + 448, 64 * RB, &s.bar);
  }
  if (nb[5] >= 0) {
    tma_load_1d(s.zf[0][1], A.fld[2] + (size_t)nb[5] * 512, 64 * RB, &s.bar);
    tma_load_1d(s.zf[1][1], A.fld[5] + (size_t)nb[5] * 512, 64 * RB, &s.bar);
  }
  if (nb[2] >= 0) {
    tma_load_3d(s.yf[0][0], &M.yv, 0, 7, nb[2] * 8, &s.bar);
    tma_load_3d(s.yf[1][0], &M.yd, 0, 7, nb[2] * 8, &s.bar);
  }
  if (nb[3] >= 0) {
    tma_load_3d(s.yf[0][1], &M.yv, 0, 0, nb[3] * 8, &s.bar);
    tma_load_3d(s.yf[1][1], &M.yd, 0, 0, nb[3] * 8, &s.bar);
  }
  if (nb[0] >= 0) {
    tma_load_3d(s.xf[0][0], &M.xu, 8 - NCOL, 0, nb[0] * 8, &s.bar);
    tma_load_3d(s.xf[1][0], &M.xd, 8 - NCOL, 0, nb[0] * 8, &s.bar);
  }
  if (nb[1] >= 0) {
    tma_load_3d(s.xf[0][1], &M.xu, 0, 0, nb[1] * 8, &s.bar);
    tma_load_3d(s.xf[1][1], &M.xd, 0, 0, nb[1] * 8, &s.bar);
  }
}

template <typename Real>
__global__ void __launch_bounds__(TPB, 6) k_prhs_tma(LevelView lv, PrhsArgs<Real> A,
                                                     const __grid_constant__ PrhsMaps M) {
  constexpr int NCOL = PrhsStage<Real>::NCOL;
  __shared__ PrhsStage<Real> s;
  const int t = threadIdx.x, x = t & 7, y = t >> 3, a = t & 7, c2 = t >> 3;
  const Real *rsl = lv.rslab ? rslab_of<Real>(lv) : nullptr;
  auto rem = [&](int nbc, int q) -> Real {  // stored by another GPU: read through L2
    return __ldcg(rsl + (size_t)(kRemote0 - nbc) * (64 * kSlabPlanes) + q * 64 + t);
  };
  const int nmine = lv.nact > (int)blockIdx.x ? (lv.nact - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (nmine == 0)
    return;
  auto load_nb = [&](int b, int (&nb)[6]) {
#pragma unroll
    for (int f = 0; f < 6; f++)
      nb[f] = lv.nbr[(size_t)b * 6 + f];
  };
  if (t == 0) {
    mbar_init(&s.bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  int bcur = blockIdx.x, slot = lv.act[bcur], nbc[6], bnext = -1, slotn = 0, nbn[6] = {0, 0, 0, 0, 0, 0};
  load_nb(bcur, nbc);
  if (nmine > 1) {
    bnext = bcur + gridDim.x;
    slotn = lv.act[bnext];
    load_nb(bnext, nbn);
  }
  if (t == 0)
    prhs_issue<Real>(s, A, M, slot, nbc);
  for (int j = 0; j < nmine; j++) {
    const size_t own = (size_t)slot * 512;
    bool odd = false;
#pragma unroll
    for (int f = 0; f < 6; f++)
      odd |= nbc[f] < 0;
    mbar_wait(&s.bar, j & 1);
    if (odd) {
      if (nbc[0] < 0) {
        s.xf[0][0][(c2 * 8 + a) * NCOL + NCOL - 1] = nbc[0] == kWall ? -s.f[0][c2 * 64 + a * 8] : rem(nbc[0], 0);
        s.xf[1][0][(c2 * 8 + a) * NCOL + NCOL - 1] = nbc[0] == kWall ? -s.f[3][c2 * 64 + a * 8] : rem(nbc[0], 3);
      }
      if (nbc[1] < 0) {
        s.xf[0][1][(c2 * 8 + a) * NCOL] = nbc[1] == kWall ? -s.f[0][c2 * 64 + a * 8 + 7] : rem(nbc[1], 0);
        s.xf[1][1][(c2 * 8 + a) * NCOL] = nbc[1] == kWall ? -s.f[3][c2 * 64 + a * 8 + 7] : rem(nbc[1], 3);
      }
      if (nbc[2] < 0) {
        s.yf[0][0][t] = nbc[2] == kWall ? -s.f[1][c2 * 64 + a] : rem(nbc[2], 1);
        s.yf[1][0][t] = nbc[2] == kWall ? -s.f[4][c2 * 64 + a] : rem(nbc[2], 4);
      }
      if (nbc[3] < 0) {
        s.yf[0][1][t] = nbc[3] == kWall ? -s.f[1][c2 * 64 + 56 + a] : rem(nbc[3], 1);
        s.yf[1][1][t] = nbc[3] == kWall ? -s.f[4][c2 * 64 + 56 + a] : rem(nbc[3], 4);
      }
      if (nbc[4] < 0) {
        s.zf[0][0][t] = nbc[4] == kWall ? -s.f[2][t] : rem(nbc[4], 2);
        s.zf[1][0][t] = nbc[4] == kWall ? -s.f[5][t] : rem(nbc[4], 5);
      }
      if (nbc[5] < 0) {
        s.zf[0][1][t] = nbc[5] == kWall ? -s.f[2][448 + t] : rem(nbc[5], 2);
        s.zf[1][1][t] = nbc[5] == kWall ? -s.f[5][448 + t] : rem(nbc[5], 5);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }
    Real w[8], dz[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      w[k] = s.f[2][k * 64 + t];
      dz[k] = s.f[5][k * 64 + t];
    }
    const Real wzm = s.zf[0][0][t], wzp = s.zf[0][1][t], dzm = s.zf[1][0][t], dzp = s.zf[1][1][t];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int i = k * 64 + t;
      const Real uxp = x < 7 ? s.f[0][i + 1] : s.xf[0][1][(k * 8 + y) * NCOL];
      const Real uxm = x > 0 ? s.f[0][i - 1] : s.xf[0][0][(k * 8 + y) * NCOL + NCOL - 1];
      const Real vyp = y < 7 ? s.f[1][i + 8] : s.yf[0][1][k * 8 + x];
      const Real vym = y > 0 ? s.f[1][i - 8] : s.yf[0][0][k * 8 + x];
      const Real wzp_ = k < 7 ? w[k < 7 ? k + 1 : 7] : wzp;
      const Real wzm_ = k > 0 ? w[k > 0 ? k - 1 : 0] : wzm;
      const Real dxp = x < 7 ? s.f[3][i + 1] : s.xf[1][1][(k * 8 + y) * NCOL];
      const Real dxm = x > 0 ? s.f[3][i - 1] : s.xf[1][0][(k * 8 + y) * NCOL + NCOL - 1];
      const Real dyp = y < 7 ? s.f[4][i + 8] : s.yf[1][1][k * 8 + x];
      const Real dym = y > 0 ? s.f[4][i - 8] : s.yf[1][0][k * 8 + x];
      const Real dzp_ = k < 7 ? dz[k < 7 ? k + 1 : 7] : dzp;
      const Real dzm_ = k > 0 ? dz[k > 0 ? k - 1 : 0] : dzm;
      Real p = A.fac * (((((uxp - uxm) + vyp) - vym) + wzp_) - wzm_);
      const Real div_us = ((((dxp - dxm) + dyp) - dym) + dzp_) - dzm_;
      p += -s.f[6][i] * A.fac * div_us;
      A.lhs[own + i] = p;
    }
    __syncthreads();  // the stage is free
    if (t == 0 && bnext >= 0)
      prhs_issue<Real>(s, A, M, slotn, nbn);
    bcur = bnext;
    slot = slotn;
#pragma unroll
    for (int f = 0; f < 6; f++)
      nbc[f] = nbn[f];
    if (j + 2 < nmine) {
      bnext = bcur + gridDim.x;
      slotn = lv.act[bnext];
      load_nb(bnext, nbn);
    } else {
      bnext = -1;
    }
  }
}

}  // namespace

template <typename Real>
int prhs_tma_launch(CupCtx *c, LevelView lv, double fac) {
  if (lv.nact <= 0)
    return CUP_OK;
  PrhsArgs<Real> A;
  const int order[7] = {CUP_F_VEL, CUP_F_VEL + 1, CUP_F_VEL + 2, CUP_F_TMP, CUP_F_TMP + 1, CUP_F_TMP + 2, CUP_F_CHI};
  for (int q = 0; q < 7; q++)
    A.fld[q] = (const Real *)c->state[order[q]];
  A.lhs = (Real *)c->state[CUP_F_LHS];
  A.fac = (Real)fac;
  PrhsMaps M;
  CUtensorMap m[4];
  CUP_TRY(tma_face_maps(c, c->state[CUP_F_VEL], nullptr, m));      // u: x faces
  M.xu = m[0];
  CUP_TRY(tma_face_maps(c, c->state[CUP_F_TMP], nullptr, m));      // udef_x: x faces
  M.xd = m[0];
  CUP_TRY(tma_face_maps(c, c->state[CUP_F_VEL + 1], nullptr, m));  // v: y faces
  M.yv = m[1];
  CUP_TRY(tma_face_maps(c, c->state[CUP_F_TMP + 1], nullptr, m));  // udef_y: y faces
  M.yd = m[1];
  static int per_sm = getenv("CUP_PRHS_PER_SM") ? atoi(getenv("CUP_PRHS_PER_SM")) : 6;
  long long g = (long long)c->num_sms * per_sm;
  if (g > lv.nact)
    g = lv.nact;
  k_prhs_tma<Real><<<(int)g, TPB, 0, c->stream>>>(lv, A, M);
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

template int prhs_tma_launch<double>(CupCtx *, LevelView, double);
template int prhs_tma_launch<float>(CupCtx *, LevelView, double);

}  // namespace cup
