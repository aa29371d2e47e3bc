// smooth_tma.cuh -- launcher of the TMA-staged smoother (smooth_tma.cu)
#pragma once
#include "mg_device.cuh"
struct CupCtx;
namespace cup {
template <typename Real>
int smooth_tma_launch(CupCtx *c, cudaStream_t stream, int grid, LevelView lv, const int *sub, int nsub,
                      SlotVec<Real> src, SlotVec<Real> dst, SlotVec<Real> f, Real h,
                      Real invh, Real om, const double *fmean);
void free_tma_cache(CupCtx *c);
}  // namespace cup
