// advdiff_tma.cuh -- launcher of the TMA-staged k_advdiff (advdiff_tma.cu)
#pragma once
#include <cuda.h>

#include "mg_device.cuh"
struct CupCtx;
namespace cup {
// x slab {4,8,8} and y slab {8,3,8} tensor maps over a leaf vector (smooth_tma.cu)
int tma_slab_maps(CupCtx *c, const void *leaf, CUtensorMap out[2]);
// TMP_c += fac_a (U . grad) u_c + fac_d lap u_c over the blocks sub[0..nsub) (or 0..nsub) whose six
// neighbours are same-level blocks, walls or faces received from other ranks.  d_hblk != null: per-block
// factors from the block's own h (multi-level meshes).
// rk != null: the Runge-Kutta stage update is fused into the sweep (uniform meshes): with T = TMP + rhs,
// vout_c = u_c + T * ih3 and TMP_c = T * beta (advdiff(), main.c:5039-5054); vout must not alias F_VEL
struct AdvRk {
  void *vout[3];
  double ih3, beta;
};
template <typename Real>
int advdiff_tma_launch(CupCtx *c, LevelView lv, const int *d_sub, int nsub, const void *d_hblk, double dtnu_dt,
                       double dtnu_nu, double fac_a, double fac_d, const AdvRk *rk = nullptr);
// k_prhs with TMA-staged fields and ghost faces (prhs_tma.cu); uniform leaf level, fac = h^2/(2 dt)
template <typename Real>
int prhs_tma_launch(CupCtx *c, LevelView lv, double fac);
}  // namespace cup
