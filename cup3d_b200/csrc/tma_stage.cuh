// tma_stage.cuh -- one shared-memory stage holding an 8^3 block, optionally a
// second vector's block, and its six ghost faces, filled entirely by the TMA
// engine (see smooth_tma.cu for the rationale).  Used by the residual /
// restriction and operator kernels.
#pragma once
#include "mg_device.cuh"
#include "tma.cuh"

namespace cup {

template <typename Real>
struct alignas(128) Stage {
  static constexpr int NCOL = 16 / (int)sizeof(Real);
  alignas(128) Real u[512];
  alignas(128) Real f[512];
  alignas(128) Real z[2][64];
  alignas(128) Real y[2][64];
  alignas(128) Real x[2][64 * NCOL];
  alignas(8) uint64_t mbar;
  int flags;  // bit f: wall; bit 8+f: received (compact) face
};

struct FaceMaps {
  CUtensorMap x_leaf, y_leaf, x_extra, y_extra;
};

// producer (one thread): stage block `slot` of u (and of fv when WITH_F) plus its ghost faces
template <typename Real, bool WITH_F>
__device__ __forceinline__ void stage_issue(Stage<Real> &s, const SlotVec<Real> &u, const SlotVec<Real> &fv, int slot,
                                            const int (&nb)[6], const Real *rf, const CUtensorMap *mxl,
                                            const CUtensorMap *myl, const CUtensorMap *mxe, const CUtensorMap *mye) {
  constexpr int NCOL = Stage<Real>::NCOL;
  int fl = 0;
#pragma unroll
  for (int f = 0; f < 6; f++)
    fl |= ((nb[f] == kWall) << f) | ((nb[f] <= kRemote0) << (8 + f));
  s.flags = fl;
  uint32_t bytes = (512 * (WITH_F ? 2 : 1) + 64 * 4 + 64 * NCOL * 2) * (uint32_t)sizeof(Real);
  if (nb[0] <= kRemote0) bytes -= 64 * (NCOL - 1) * (uint32_t)sizeof(Real);
  if (nb[1] <= kRemote0) bytes -= 64 * (NCOL - 1) * (uint32_t)sizeof(Real);
  mbar_arrive_expect_tx(&s.mbar, bytes);
  const Real *own = u.at(slot);
  tma_load_1d(s.u, own, 512 * sizeof(Real), &s.mbar);
  if (WITH_F)
    tma_load_1d(s.f, fv.at(slot), 512 * sizeof(Real), &s.mbar);
  tma_load_1d(s.z[0], nb[4] >= 0 ? u.at(nb[4]) + 7 * 64 : (nb[4] == kWall ? own : rf + (size_t)(kRemote0 - nb[4]) * 64),
              64 * sizeof(Real), &s.mbar);
  tma_load_1d(s.z[1], nb[5] >= 0 ? u.at(nb[5]) : (nb[5] == kWall ? own + 7 * 64 : rf + (size_t)(kRemote0 - nb[5]) * 64),
              64 * sizeof(Real), &s.mbar);
#pragma unroll
  for (int f = 0; f < 4; f++) {
    if (nb[f] <= kRemote0) {
      tma_load_1d(f < 2 ? (Real *)s.x[f] : (Real *)s.y[f - 2], rf + (size_t)(kRemote0 - nb[f]) * 64, 64 * sizeof(Real),
                  &s.mbar);
      continue;
    }
    const int ts = nb[f] >= 0 ? nb[f] : slot;
    const bool leaf = ts < u.nleaf;
    const int row0 = (leaf ? ts : ts - u.nleaf) * 8;
    const bool high = (nb[f] >= 0) ? !(f & 1) : (f & 1);
    if (f < 2)
      tma_load_3d(s.x[f], leaf ? mxl : mxe, high ? 8 - NCOL : 0, 0, row0, &s.mbar);
    else
      tma_load_3d(s.y[f - 2], leaf ? myl : mye, 0, high ? 7 : 0, row0, &s.mbar);
  }
}

// consumer-side view of the x faces (stride / column depend on wall-ness and origin)
template <typename Real>
struct XFace {
  int stm, om, stp, op;
  __device__ __forceinline__ explicit XFace(int fl) {
    constexpr int NCOL = Stage<Real>::NCOL;
    const int cxm = (fl & 1) ? 0 : NCOL - 1, cxp = (fl & 2) ? NCOL - 1 : 0;
    stm = (fl & 0x100) ? 1 : NCOL;
    stp = (fl & 0x200) ? 1 : NCOL;
    om = (fl & 0x100) ? 0 : cxm;
    op = (fl & 0x200) ? 0 : cxp;
  }
};

// 7-point operator on the thread's z-line from a stage: k_lhs/k_mg summation order
template <typename Real>
__device__ __forceinline__ void lap_line_stage(const Stage<Real> &s, const XFace<Real> &xf, const Real (&uu)[8], int x,
                                               int y, int t, Real h, Real (&out)[8]) {
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const Real xm = x > 0 ? s.u[k * 64 + t - 1] : s.x[0][(k * 8 + y) * xf.stm + xf.om];
    const Real xp = x < 7 ? s.u[k * 64 + t + 1] : s.x[1][(k * 8 + y) * xf.stp + xf.op];
    const Real ym = y > 0 ? s.u[k * 64 + t - 8] : s.y[0][k * 8 + x];
    const Real yp = y < 7 ? s.u[k * 64 + t + 8] : s.y[1][k * 8 + x];
    const Real zm = k > 0 ? uu[k > 0 ? k - 1 : 0] : s.z[0][t];
    const Real zp = k < 7 ? uu[k < 7 ? k + 1 : 7] : s.z[1][t];
    out[k] = h * ((((((xm + xp) + ym) + yp) + zm) + zp) - (Real)6.0 * uu[k]);
  }
}

}  // namespace cup
