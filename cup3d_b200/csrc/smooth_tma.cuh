// smooth_tma.cuh -- launcher of the TMA-staged smoother (smooth_tma.cu)
#pragma once
#include <cstddef>

#include "mg_device.cuh"
struct CupCtx;
namespace cup {
// Fused halo exchange (one-sided transport, see comm.cu): the first `nbnd` work items are the
// blocks with a neighbour on another rank.  Their CTAs (1) wait for the peers' flags, (2) sweep,
// (3) store the new boundary planes straight into the neighbours' receive windows over NVLink,
// and the CTA that retires the last of them publishes the new sequence number -- all while the
// remaining CTAs / iterations sweep the interior.  One kernel per sweep, no pack/NCCL kernels.
struct FusedComm {
  const int *bsend;                    // [nact][6] face-send entry of (block, plane) or -1
  void *const *fptr0, *const *fptr1;   // per entry: destination for even / odd sequence numbers
  unsigned long long *seq;             // this level's face sequence number (device)
  const unsigned long long *my_flags;  // [nranks] what each peer has published to this rank
  const int *rpeers;                   // ranks this one receives faces from
  int nrp;
  char *const *peer_win;               // base of every rank's window
  const int *speers;                   // ranks this one sends faces to
  int nsp;
  size_t flag_index;                   // index of this rank's flag word in a peer's window
  unsigned int *counters;              // [0] boundary blocks retired, [1] CTAs retired
  int nbnd;
  int *err;                            // host-visible error word (comm_dev.cuh)
  int code;
  int push_mode;                       // 0: plain stores, 1: staged coalesced stores (default), 2: TMA bulk stores
};

// Fused prolongation (mg_up/mg_up2, main.c:4787-4807): the first post-smoothing sweep of a level
// reads u + P(d), d = u_c - us of the parent level (precomputed in place of us), instead of a
// separate pass that adds the correction to u.  Per block the producer additionally stages the
// 4^3 octant of the parent's d and the 4x4 patches of the six neighbours' parents.
struct UpFuse {
  const int *info;  // [nact][7]: own + six face neighbours: (parent index in the extra part << 3) | octant;
                    // -1 for faces received from another rank (their owner adds the correction)
};

template <typename Real>
int smooth_tma_launch(CupCtx *c, cudaStream_t stream, int grid, LevelView lv, const int *sub, int nsub,
                      SlotVec<Real> src, SlotVec<Real> dst, SlotVec<Real> f, Real h,
                      Real invh, Real om, const double *fmean, const FusedComm *fused = nullptr,
                      const UpFuse *upf = nullptr, const void *d_extra = nullptr);
void free_tma_cache(CupCtx *c);
}  // namespace cup
