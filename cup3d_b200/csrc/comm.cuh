// comm.cuh -- multi-rank exchanges (comm.cu)
#pragma once
#include <vector>

#include "cup_internal.h"
#include "mg_device.cuh"
namespace cup {
int comm_gather_blocks(CupCtx *c, const CupBlk *blk, long long n, std::vector<CupBlk> &gblk, std::vector<int> &owner);
int comm_allreduce(CupCtx *c, int first, int n);  // in-place sum of d_scal[first..first+n) over ranks
int comm_alloc_level_buffers(CupCtx *c);
void comm_free_level_buffers(CupCtx *c);
void comm_free(CupCtx *c);
template <typename Real>
int halo_exchange(CupCtx *c, Level &v, SlotVec<Real> u);  // faces of u -> v.d_frecv on the neighbours' owners
template <typename Real>
int restrict_exchange(CupCtx *c, Level &v, SlotVec<Real> f, SlotVec<Real> u);
template <typename Real>
int prolong_exchange(CupCtx *c, Level &v, SlotVec<Real> u, SlotVec<Real> us);
}  // namespace cup
