// stencil7_tma.cuh -- launchers of the TMA-staged residual/restriction and operator kernels
#pragma once
#include <cuda.h>

#include "comm_dev.cuh"
#include "mg_device.cuh"
struct CupCtx;
namespace cup {
// the four face tensor maps (x/y boxes over the leaf and the extra part) of a slot vector
int tma_face_maps(CupCtx *c, const void *leaf, const void *extra, CUtensorMap out[4]);
template <typename Real>
int down_tma_launch(CupCtx *c, LevelView lv, const int *pslot, const int *oct, SlotVec<Real> u, SlotVec<Real> f,
                    Real h, void *const *rptr, const int *sub = nullptr, int nsub = -1,
                    const WaitDesc *wait = nullptr, const PostDesc *post = nullptr);
template <typename Real>
int apply_tma_launch(CupCtx *c, LevelView lv, const int *sub, int nsub, SlotVec<Real> u, SlotVec<Real> out,
                     SlotVec<Real> us, Real h, const double *shift, Real h3, bool tau,
                     const WaitDesc *wait = nullptr);
}  // namespace cup
