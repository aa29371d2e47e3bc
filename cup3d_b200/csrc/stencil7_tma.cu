// stencil7_tma.cu -- TMA-staged versions of the residual+restriction kernel
// (mg_down, main.c:4734) and of the 7-point operator (k_lhs/k_mg/mg_tau,
// main.c:4254-4280, :4758).  Same staging as the smoother (smooth_tma.cu): the
// block and its six ghost faces arrive through bulk / tensor-map copies on one
// mbarrier, the next block's loads are issued as soon as the stage has been
// read, one __syncthreads per block.
#include <cuda.h>

#include "comm_dev.cuh"
#include "cup_internal.h"
#include "stencil7_tma.cuh"
#include "tma_stage.cuh"

namespace cup {

// r = f - A u, then parent octant <- (sum_8 r, mean_8 u).  The 2x2x2 sums are
// formed in registers (z pairs) and with two shuffles (x and y pairs); the
// order differs from mg_sum's left-to-right sum by rounding only.
template <typename Real>
__global__ void __launch_bounds__(TPB, 12)
    k_down_tma(LevelView lv, const int *__restrict__ sub, int nsub, const int *__restrict__ pslot,
               const int *__restrict__ oct, SlotVec<Real> u,
               SlotVec<Real> f, Real h, Real *const *__restrict__ rptr, WaitDesc wait, PostDesc post,
               const __grid_constant__ CUtensorMap mxl,
               const __grid_constant__ CUtensorMap myl, const __grid_constant__ CUtensorMap mxe,
               const __grid_constant__ CUtensorMap mye) {
  __shared__ Stage<Real> st;
  __shared__ Real s_res[128];
  const int t = threadIdx.x, x = t & 7, y = t >> 3, G = gridDim.x;
  const Real *rf = rface_of<Real>(lv);
  if (t == 0) {
    mbar_init(&st.mbar, 1);
    mbar_fence_init();
    // the ghost faces of u posted by the last pre-smoothing sweep (one-sided transport)
    if (wait.seq)
      comm_wait_for(wait, *(const volatile unsigned long long *)wait.seq);
  }
  __syncthreads();
  int i = blockIdx.x;
  if (t == 0 && i < nsub) {
    const int b0 = sub ? sub[i] : i;
    int nb[6];
#pragma unroll
    for (int q = 0; q < 6; q++)
      nb[q] = lv.nbr[(size_t)b0 * 6 + q];
    stage_issue<Real, true>(st, u, f, lv.act[b0], nb, rf, &mxl, &myl, &mxe, &mye);
  }
  uint32_t phase = 0;
  for (; i < nsub; i += G) {
    const int b = sub ? sub[i] : i;
    int nslot = 0, nnb[6] = {0, 0, 0, 0, 0, 0};
    const bool more = (i + G) < nsub;
    if (t == 0 && more) {
      const int bn = sub ? sub[i + G] : i + G;
      nslot = lv.act[bn];
#pragma unroll
      for (int q = 0; q < 6; q++)
        nnb[q] = lv.nbr[(size_t)bn * 6 + q];
    }
    const int ps = pslot[b], o = oct[b];
    mbar_wait(&st.mbar, phase);
    phase ^= 1;
    const XFace<Real> xf(st.flags);
    Real uu[8], r[8];
#pragma unroll
    for (int k = 0; k < 8; k++)
      uu[k] = st.u[k * 64 + t];
    lap_line_stage<Real>(st, xf, uu, x, y, t, h, r);
#pragma unroll
    for (int k = 0; k < 8; k++)
      r[k] = st.f[k * 64 + t] - r[k];
    __syncthreads();  // stage consumed
    if (t == 0 && more)
      stage_issue<Real, true>(st, u, f, nslot, nnb, rf, &mxl, &myl, &mxe, &mye);
    Real sr[4], su[4];
#pragma unroll
    for (int k2 = 0; k2 < 4; k2++) {
      sr[k2] = r[2 * k2] + r[2 * k2 + 1];
      su[k2] = uu[2 * k2] + uu[2 * k2 + 1];
      sr[k2] += __shfl_xor_sync(0xffffffffu, sr[k2], 1);
      su[k2] += __shfl_xor_sync(0xffffffffu, su[k2], 1);
      sr[k2] += __shfl_xor_sync(0xffffffffu, sr[k2], 8);
      su[k2] += __shfl_xor_sync(0xffffffffu, su[k2], 8);
    }
    if (ps >= 0) {
      if (((x | y) & 1) == 0) {
        const int cx = x >> 1, cy = y >> 1;
        Real *pf = f.at(ps), *pu = u.at(ps);
#pragma unroll
        for (int cz = 0; cz < 4; cz++) {
          const int pidx = ((4 * (o >> 2) + cz) << 6) + ((4 * ((o >> 1) & 1) + cy) << 3) + 4 * (o & 1) + cx;
          pf[pidx] = sr[cz];
          pu[pidx] = (Real)0.125 * su[cz];
        }
      }
    } else {
      // parent on another rank: MG_M layout (64 r, 64 u), main.c:4750.  The 128 values are gathered in shared
      // memory first so that they cross NVLink as two contiguous 64-element stores of the whole CTA instead
      // of sixteen threads' scattered words (ps is the same for the whole block: no divergence)
      if (((x | y) & 1) == 0) {
        const int cx = x >> 1, cy = y >> 1;
#pragma unroll
        for (int cz = 0; cz < 4; cz++) {
          s_res[(cz * 4 + cy) * 4 + cx] = sr[cz];
          s_res[64 + (cz * 4 + cy) * 4 + cx] = (Real)0.125 * su[cz];
        }
      }
      __syncthreads();
      Real *q = rptr[kRemote0 - ps];
      q[t] = s_res[t];
      q[64 + t] = s_res[64 + t];
      __syncthreads();
    }
  }
  comm_post_at_exit(post);  // children of remote parents are in their owners' windows
}

// out = A u on the blocks sub[0..nsub) (or all).  TAU: out += A u, us = u (mg_tau).
template <typename Real, bool TAU>
__global__ void __launch_bounds__(TPB, 12)
    k_apply_tma(LevelView lv, const int *__restrict__ sub, int nsub, SlotVec<Real> u, SlotVec<Real> out,
                SlotVec<Real> us, Real h, const double *__restrict__ shift, Real h3, WaitDesc wait,
                const __grid_constant__ CUtensorMap mxl, const __grid_constant__ CUtensorMap myl,
                const __grid_constant__ CUtensorMap mxe, const __grid_constant__ CUtensorMap mye) {
  __shared__ Stage<Real> st;
  const int t = threadIdx.x, x = t & 7, y = t >> 3, G = gridDim.x;
  const Real *rf = rface_of<Real>(lv);
  const Real add = shift ? (Real)(*shift) * h3 : (Real)0;
  if (t == 0) {
    mbar_init(&st.mbar, 1);
    mbar_fence_init();
    if (wait.seq)
      comm_wait_for(wait, *(const volatile unsigned long long *)wait.seq);
  }
  __syncthreads();
  int i = blockIdx.x;
  if (t == 0 && i < nsub) {
    const int b0 = sub ? sub[i] : i;
    int nb[6];
#pragma unroll
    for (int q = 0; q < 6; q++)
      nb[q] = lv.nbr[(size_t)b0 * 6 + q];
    stage_issue<Real, false>(st, u, u, lv.act[b0], nb, rf, &mxl, &myl, &mxe, &mye);
  }
  uint32_t phase = 0;
  for (; i < nsub; i += G) {
    const int b = sub ? sub[i] : i;
    const int slot = lv.act[b];
    int nslot = 0, nnb[6] = {0, 0, 0, 0, 0, 0};
    const bool more = (i + G) < nsub;
    if (t == 0 && more) {
      const int bn = sub ? sub[i + G] : i + G;
      nslot = lv.act[bn];
#pragma unroll
      for (int q = 0; q < 6; q++)
        nnb[q] = lv.nbr[(size_t)bn * 6 + q];
    }
    Real *ob = out.at(slot);
    Real ov[8];
    if (TAU) {
#pragma unroll
      for (int k = 0; k < 8; k++)
        ov[k] = ob[k * 64 + t];  // in flight while the stage lands
    }
    mbar_wait(&st.mbar, phase);
    phase ^= 1;
    const XFace<Real> xf(st.flags);
    Real uu[8], tt[8];
#pragma unroll
    for (int k = 0; k < 8; k++)
      uu[k] = st.u[k * 64 + t];
    lap_line_stage<Real>(st, xf, uu, x, y, t, h, tt);
    __syncthreads();
    if (t == 0 && more)
      stage_issue<Real, false>(st, u, u, nslot, nnb, rf, &mxl, &myl, &mxe, &mye);
    if (TAU) {
      Real *sb = us.at(slot);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        ob[k * 64 + t] = ov[k] + tt[k];
        sb[k * 64 + t] = uu[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++)
        ob[k * 64 + t] = tt[k] + add;
    }
  }
}

static inline int pgrid(const CupCtx *c, long long n) {
  long long g = (long long)c->num_sms * 12;
  return (int)(g < n ? g : (n < 1 ? 1 : n));
}

template <typename Real>
int down_tma_launch(CupCtx *c, LevelView lv, const int *pslot, const int *oct, SlotVec<Real> u, SlotVec<Real> f,
                    Real h, void *const *rptr, const int *sub, int nsub, const WaitDesc *wait, const PostDesc *post) {
  CUtensorMap m[4];
  CUP_TRY(tma_face_maps(c, u.leaf, u.extra, m));
  if (nsub < 0)
    nsub = lv.nact;
  k_down_tma<Real><<<pgrid(c, nsub), TPB, 0, c->stream>>>(lv, sub, nsub, pslot, oct, u, f, h, (Real *const *)rptr,
                                                         wait ? *wait : WaitDesc{}, post ? *post : PostDesc{}, m[0],
                                                         m[1], m[2], m[3]);
  return CUP_OK;
}

template <typename Real>
int apply_tma_launch(CupCtx *c, LevelView lv, const int *sub, int nsub, SlotVec<Real> u, SlotVec<Real> out,
                     SlotVec<Real> us, Real h, const double *shift, Real h3, bool tau, const WaitDesc *wait) {
  CUtensorMap m[4];
  CUP_TRY(tma_face_maps(c, u.leaf, u.extra, m));
  const WaitDesc w = wait ? *wait : WaitDesc{};
  if (tau)
    k_apply_tma<Real, true><<<pgrid(c, nsub), TPB, 0, c->stream>>>(lv, sub, nsub, u, out, us, h, shift, h3, w, m[0],
                                                                    m[1], m[2], m[3]);
  else
    k_apply_tma<Real, false><<<pgrid(c, nsub), TPB, 0, c->stream>>>(lv, sub, nsub, u, out, us, h, shift, h3, w, m[0],
                                                                     m[1], m[2], m[3]);
  return CUP_OK;
}

template int down_tma_launch<double>(CupCtx *, LevelView, const int *, const int *, SlotVec<double>, SlotVec<double>,
                                     double, void *const *, const int *, int, const WaitDesc *, const PostDesc *);
template int down_tma_launch<float>(CupCtx *, LevelView, const int *, const int *, SlotVec<float>, SlotVec<float>,
                                    float, void *const *, const int *, int, const WaitDesc *, const PostDesc *);
template int apply_tma_launch<double>(CupCtx *, LevelView, const int *, int, SlotVec<double>, SlotVec<double>,
                                      SlotVec<double>, double, const double *, double, bool, const WaitDesc *);
template int apply_tma_launch<float>(CupCtx *, LevelView, const int *, int, SlotVec<float>, SlotVec<float>,
                                     SlotVec<float>, float, const double *, float, bool, const WaitDesc *);

}  // namespace cup
