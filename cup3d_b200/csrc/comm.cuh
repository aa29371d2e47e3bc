// comm.cuh -- multi-rank exchanges (comm.cu)
#pragma once
#include <vector>

#include "comm_dev.cuh"
#include "cup_internal.h"
#include "mg_device.cuh"
namespace cup {
int comm_gather_blocks(CupCtx *c, const CupBlk *blk, long long n, std::vector<CupBlk> &gblk, std::vector<int> &owner);
int comm_allreduce(CupCtx *c, int first, int n);  // in-place sum of d_scal[first..first+n) over ranks
int comm_allreduce_max(CupCtx *c, int first, int n);
int comm_alloc_level_buffers(CupCtx *c);
struct FusedComm;
// descriptor of the fused sweep+exchange for level v; false when the level cannot use it
bool comm_fused_desc(CupCtx *c, Level &v, FusedComm *out);
void comm_free_level_buffers(CupCtx *c);
void comm_free(CupCtx *c);
// publish the ghost faces of u to the neighbours' owners / make sure the last posted ones arrived
template <typename Real>
int halo_post(CupCtx *c, Level &v, SlotVec<Real> u);
int halo_wait(CupCtx *c, Level &v);
template <typename Real>
inline int halo_exchange(CupCtx *c, Level &v, SlotVec<Real> u) {
  int rc = halo_post<Real>(c, v, u);
  return rc != CUP_OK ? rc : halo_wait(c, v);
}
// up to six component vectors (flat [slot][512]) whose ghost slabs travel together
template <typename Real>
struct SlabSrc {
  const Real *c[6];
};
template <typename Real>
int slab_exchange(CupCtx *c, Level &v, const SlabSrc<Real> &src, int ncomp, int nlayer);
// `posted`: k_down published the exchange itself (it ran with comm_post_desc(c, v, COMM_RES))
template <typename Real>
int restrict_exchange(CupCtx *c, Level &v, SlotVec<Real> f, SlotVec<Real> u, bool posted = false);
// one-sided transport: descriptors for kernels that wait for / publish an exchange themselves
// (empty descriptors -- null seq -- on one rank, in NCCL mode, or when there is nothing to do)
enum { COMM_FACE = 0, COMM_RES = 1, COMM_PRO = 2 };
// ghost blocks (levels with coarse-fine interfaces / the leaf context of a multi-level mesh across ranks)
template <typename Real>
int block_exchange_mg(CupCtx *c, Level &v, SlotVec<Real> same, SlotVec<Real> can);
template <typename Real>
int block_exchange_leaf(CupCtx *c, const Real *const *src, Real *const *dst, int ncomp, long long dst_off);
WaitDesc comm_wait_desc(CupCtx *c, Level &v, int kind);
PostDesc comm_post_desc(CupCtx *c, Level &v, int kind);
template <typename Real>
int prolong_exchange(CupCtx *c, Level &v, SlotVec<Real> u, SlotVec<Real> us);
}  // namespace cup
