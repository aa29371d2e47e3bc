// amr_kernels.cuh -- launchers of the kernels for levels with coarse-fine interfaces
#pragma once
#include "cup_internal.h"
#include "mg_device.cuh"
namespace cup {
int amr_setup_constants();
// src: ping-pong source (same-level neighbours); can: canonical vector (coarser leaves)
template <typename Real>
int smooth_amr_launch(CupCtx *c, LevelView lv, SlotVec<Real> src, SlotVec<Real> can, SlotVec<Real> dst,
                      SlotVec<Real> f, Real h, const double *fmean, bool zero_src, const int *sub = nullptr,
                      int nsub = -1);
template <typename Real>
int down_amr_launch(CupCtx *c, LevelView lv, const int *pslot, const int *oct, SlotVec<Real> u, SlotVec<Real> f,
                    Real h, const int *sub = nullptr, int nsub = -1, void *const *rptr = nullptr);
// mode 0: out = A u (+shift h^3); 1: tau (out += A u, us = u); 2: out = A u with flux correction
template <typename Real>
int apply_amr_launch(CupCtx *c, LevelView lv, const int *sub, int nsub, SlotVec<Real> u, SlotVec<Real> out,
                     SlotVec<Real> us, Real h, const void *hblk, const double *shift, int mode);
// leaf sweeps with flux correction: what 0 = k_divp (facs unused), 1 = k_gradp (facs = -dt/2)
template <typename Real>
int pres_amr_launch(CupCtx *c, LevelView lv, const void *hblk, const Real *p, Real *o0, Real *o1, Real *o2, Real facs,
                    int what);
// k_prhs; S = the nine state components, idt2 = 1/dt
template <typename Real>
int prhs_amr_launch(CupCtx *c, LevelView lv, const void *hblk, Real *const *S, Real idt2);
// what 0 = k_vort (F_VEL -> F_TMP, with its flux correction), 1 = k_q (F_VEL -> F_LHS)
template <typename Real>
int velgrad_amr_launch(CupCtx *c, LevelView lv, const void *hblk, Real *const *S, int what);
// k_advdiff on the leaves of a multi-level mesh (amr_advdiff.cu)
template <typename Real>
int advdiff_amr_launch(CupCtx *c, const Level &v, Real *const *S, const int *sub = nullptr, int nsub = -1);
}  // namespace cup
