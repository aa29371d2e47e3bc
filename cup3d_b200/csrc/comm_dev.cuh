// comm_dev.cuh -- device side of the one-sided (peer window) transport (comm.cu).
//
// Every exchange kind of every multigrid level has a SEQUENCE NUMBER in this rank's device memory
// (how many times this rank has posted it) and one FLAG WORD per peer in this rank's receive
// window (how many times that peer has posted it).  Both sides of an exchange post equally often,
// so "the data of my n-th post's counterpart has arrived" == "flag[peer] >= n".
//
//   producer kernel: all CTAs store into the peers' windows, then comm_post_at_exit():
//                    the LAST CTA to retire bumps the sequence number and writes it into the
//                    peers' flag words -- no separate signal kernel;
//   consumer kernel: one thread runs comm_wait() before the first read of received data --
//                    no separate wait kernel.
//
// A wait that does not complete within COMM_SPIN_MAX polls (several seconds: a peer died or the
// protocol is broken) records a code in a host-visible error word and RETURNS; the host reports
// it as CUP_ERR_COMM at the next synchronisation instead of the context being poisoned by a trap.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cup {

enum { COMM_SPIN_MAX = 1 << 25 };

struct WaitDesc {
  const unsigned long long *seq = nullptr;    // this rank's sequence number; null: nothing to wait for
  const unsigned long long *flags = nullptr;  // [nranks] flag words of this (level, kind) in the own window
  const int *peers = nullptr;                 // ranks to wait for
  int np = 0;
  int *err = nullptr;                         // mapped host word
  int code = 0;                               // what to record on a timeout
};

struct PostDesc {
  unsigned long long *seq = nullptr;  // null: nothing to post
  char *const *peer_win = nullptr;    // base of every rank's window
  const int *peers = nullptr;         // ranks to notify
  int np = 0;
  size_t flag_index = 0;              // index of this rank's flag word inside a peer's window
  unsigned int *counter = nullptr;    // CTAs retired (re-armed by the last one)
};

// order generic-proxy reads/writes (the flag acquire) before async-proxy (TMA) reads of the data
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ONE thread: wait until every listed peer has posted at least `want` times
__device__ __forceinline__ void comm_wait_for(const WaitDesc &w, unsigned long long want) {
  for (int i = 0; i < w.np; i++) {
    const volatile unsigned long long *f = w.flags + w.peers[i];
    int spins = 0;
    while (*f < want) {
      __nanosleep(100);
      if (++spins > COMM_SPIN_MAX) {
        if (w.err)
          *(volatile int *)w.err = w.code ? w.code : 1;  // mapped host memory: a plain store, no PCIe atomic
        break;
      }
    }
  }
  __threadfence_system();
  fence_proxy_async();
}

// whole CTA, at the top of a consumer kernel (before any read of received data)
__device__ __forceinline__ void comm_wait_cta(const WaitDesc &w) {
  if (w.seq == nullptr)
    return;
  if (threadIdx.x == 0)
    comm_wait_for(w, *(const volatile unsigned long long *)w.seq);
  __syncthreads();
}

// whole CTA, after its last store of posted data (every CTA of the grid must call it)
__device__ __forceinline__ void comm_post_at_exit(const PostDesc &p) {
  if (p.seq == nullptr)
    return;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int old = atomicAdd(p.counter, 1u);
    if (old == gridDim.x - 1) {
      *p.counter = 0;
      const unsigned long long s = *(volatile unsigned long long *)p.seq + 1;
      __threadfence_system();
      for (int i = 0; i < p.np; i++) {
        volatile unsigned long long *f = (volatile unsigned long long *)p.peer_win[p.peers[i]] + p.flag_index;
        *f = s;
      }
      *(volatile unsigned long long *)p.seq = s;
      __threadfence();
    }
  }
}

// Store the boundary planes of a block's NEW values (thread t = x + 8y holds the z-line v[0..7])
// into the neighbours' owners' windows: bs[f] = send entry of plane f or -1, fp = per-entry
// destination of the exchange being posted.  Plane element order as load_halo expects.
template <typename Real>
__device__ __forceinline__ void push_faces(const int *__restrict__ bs, void *const *__restrict__ fp,
                                           const Real (&v)[8], int t, int x, int y) {
  const int e0 = bs[0], e1 = bs[1], e2 = bs[2], e3 = bs[3], e4 = bs[4], e5 = bs[5];
  if (e4 >= 0)
    ((Real *)fp[e4])[t] = v[0];
  if (e5 >= 0)
    ((Real *)fp[e5])[t] = v[7];
  if (e2 >= 0 && y == 0) {
    Real *d = (Real *)fp[e2];
#pragma unroll
    for (int k = 0; k < 8; k++)
      d[k * 8 + x] = v[k];
  }
  if (e3 >= 0 && y == 7) {
    Real *d = (Real *)fp[e3];
#pragma unroll
    for (int k = 0; k < 8; k++)
      d[k * 8 + x] = v[k];
  }
  if (e0 >= 0 && x == 0) {
    Real *d = (Real *)fp[e0];
#pragma unroll
    for (int k = 0; k < 8; k++)
      d[k * 8 + y] = v[k];
  }
  if (e1 >= 0 && x == 7) {
    Real *d = (Real *)fp[e1];
#pragma unroll
    for (int k = 0; k < 8; k++)
      d[k * 8 + y] = v[k];
  }
}

// The same through shared memory: x planes would leave the SM as 64 separate 8-byte stores (one per
// thread, 64-byte apart in the thread's view), which crosses NVLink at a fraction of its rate (r02 trace at
// 8 GPUs: ~30 us per finest-level sweep above the 1/8 share of the single-GPU time).  The CTA first gathers
// the planes it has to send in `stage` (>= 384 Reals, free at the call), then every plane leaves as ONE
// contiguous 64-element store of the whole CTA.  All 64 threads must call it (it synchronises).
template <typename Real>
__device__ __forceinline__ void push_faces_staged(const int *__restrict__ bs, void *const *__restrict__ fp,
                                                  const Real (&v)[8], int t, int x, int y, Real *stage) {
  const int e0 = bs[0], e1 = bs[1], e2 = bs[2], e3 = bs[3], e4 = bs[4], e5 = bs[5];
  __syncthreads();  // whoever used `stage` before is done with it
  if (e0 >= 0 && x == 0) {
#pragma unroll
    for (int k = 0; k < 8; k++)
      stage[0 * 64 + k * 8 + y] = v[k];
  }
  if (e1 >= 0 && x == 7) {
#pragma unroll
    for (int k = 0; k < 8; k++)
      stage[1 * 64 + k * 8 + y] = v[k];
  }
  if (e2 >= 0 && y == 0) {
#pragma unroll
    for (int k = 0; k < 8; k++)
      stage[2 * 64 + k * 8 + x] = v[k];
  }
  if (e3 >= 0 && y == 7) {
#pragma unroll
    for (int k = 0; k < 8; k++)
      stage[3 * 64 + k * 8 + x] = v[k];
  }
  __syncthreads();
  if (e0 >= 0)
    ((Real *)fp[e0])[t] = stage[t];
  if (e1 >= 0)
    ((Real *)fp[e1])[t] = stage[64 + t];
  if (e2 >= 0)
    ((Real *)fp[e2])[t] = stage[128 + t];
  if (e3 >= 0)
    ((Real *)fp[e3])[t] = stage[192 + t];
  if (e4 >= 0)
    ((Real *)fp[e4])[t] = v[0];
  if (e5 >= 0)
    ((Real *)fp[e5])[t] = v[7];
}

// The same with the copy engine: the planes are gathered in `stage` (>= 384 Reals) and ONE thread hands
// each of them to the TMA unit as a bulk store (cp.async.bulk.global.shared::cta, 64 Reals = 256 / 512 B)
// into the neighbour's window.  Stores issued from the SM to peer memory stall the warp on NVLink credits
// (r02 traces: ~50 us per finest-level sweep at 2 and 8 GPUs whether the planes left as scattered words or
// as coalesced rows); the bulk stores are asynchronous, the CTA goes on with its next block.
//   push_faces_tma  : all 64 threads; returns after the stores were ISSUED
//   push_tma_drain  : thread 0, before the CTA reports its boundary blocks retired: all bulk stores complete
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void push_tma_drain() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <typename Real>
__device__ __forceinline__ void push_faces_tma(const int *__restrict__ bs, void *const *__restrict__ fp,
                                               const Real (&v)[8], int t, int x, int y, Real *stage) {
  const int e0 = bs[0], e1 = bs[1], e2 = bs[2], e3 = bs[3], e4 = bs[4], e5 = bs[5];
  if (t == 0)  // the previous block's bulk stores have finished READING the staging area
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncthreads();
  if (e0 >= 0 && x == 0) {
#pragma unroll
    for (int k = 0; k < 8; k++)
      stage[0 * 64 + k * 8 + y] = v[k];
  }
  if (e1 >= 0 && x == 7) {
#pragma unroll
    for (int k = 0; k < 8; k++)
      stage[1 * 64 + k * 8 + y] = v[k];
  }
  if (e2 >= 0 && y == 0) {
#pragma unroll
    for (int k = 0; k < 8; k++)
      stage[2 * 64 + k * 8 + x] = v[k];
  }
  if (e3 >= 0 && y == 7) {
#pragma unroll
    for (int k = 0; k < 8; k++)
      stage[3 * 64 + k * 8 + x] = v[k];
  }
  if (e4 >= 0)
    stage[4 * 64 + t] = v[0];
  if (e5 >= 0)
    stage[5 * 64 + t] = v[7];
  fence_proxy_async();  // the generic-proxy writes above before the async-proxy reads of the bulk stores
  __syncthreads();
  if (t == 0) {
    const uint32_t bytes = 64 * (uint32_t)sizeof(Real);
    if (e0 >= 0) bulk_store(fp[e0], stage, bytes);
    if (e1 >= 0) bulk_store(fp[e1], stage + 64, bytes);
    if (e2 >= 0) bulk_store(fp[e2], stage + 128, bytes);
    if (e3 >= 0) bulk_store(fp[e3], stage + 192, bytes);
    if (e4 >= 0) bulk_store(fp[e4], stage + 256, bytes);
    if (e5 >= 0) bulk_store(fp[e5], stage + 320, bytes);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

// read of data another GPU stored into this rank's window: never through a possibly stale L1 line
template <typename Real>
__device__ __forceinline__ Real ld_recv(const Real *p) {
  return __ldcg(p);
}

}  // namespace cup
