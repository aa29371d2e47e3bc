// comm.cu -- one rank per GPU over NCCL / NVLink.
//
// Replaces the reference's MPI layer on the hot path:
//   tree_sync      (Allgatherv of all blocks, main.c:2928)   -> comm_gather_blocks
//   halo_sync      (whole 8^3 blocks by Alltoallv, :3101)    -> halo_exchange: 8x8 FACES stored into
//                                                               the neighbours' receive windows
//                                                               (ghost BLOCKS only where a level has
//                                                               coarse-fine interfaces across ranks)
//   mg_down/mg_up  (Alltoallv of 128 / 64 Reals, :4754,:4791)-> restrict_exchange / prolong_exchange
//   MPI_Allreduce  (:4295, :4820, :4860, ...)                -> ncclAllReduce on device scalars
// Everything is stream-ordered on the context's stream; the host never waits.
#include <dlfcn.h>
#include <nccl.h>

#include "comm.cuh"
#include "comm_dev.cuh"
#include "cup_internal.h"
#include "mg_device.cuh"
#include "smooth_tma.cuh"
#include <map>

namespace cup {

// NCCL is bound at run time (dlopen), never at link time: a process that has
// already loaded a libnccl.so.2 (PyTorch ships its own, newer one) must keep
// using that copy -- linking the system library first would shadow it.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  const char *(*GetErrorString)(ncclResult_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
};
static NcclApi g_nccl;
static bool g_nccl_ok = false;

static int load_nccl() {
  if (g_nccl_ok)
    return CUP_OK;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy already in the process
  if (!h)
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h)
    h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("cannot load libnccl.so.2: %s", dlerror());
    return CUP_ERR_NCCL;
  }
#define BIND(field, name)                                   \
  *(void **)(&g_nccl.field) = dlsym(h, name);               \
  if (!g_nccl.field) {                                      \
    set_error("libnccl: missing symbol %s", name);          \
    return CUP_ERR_NCCL;                                    \
  }
  BIND(GetUniqueId, "ncclGetUniqueId")
  BIND(CommInitRank, "ncclCommInitRank")
  BIND(CommDestroy, "ncclCommDestroy")
  BIND(GetErrorString, "ncclGetErrorString")
  BIND(AllReduce, "ncclAllReduce")
  BIND(AllGather, "ncclAllGather")
  BIND(Send, "ncclSend")
  BIND(Recv, "ncclRecv")
  BIND(GroupStart, "ncclGroupStart")
  BIND(GroupEnd, "ncclGroupEnd")
#undef BIND
  g_nccl_ok = true;
  return CUP_OK;
}

enum { K_FACE = 0, K_RES = 1, K_PRO = 2, K_BLK = 3, NKIND = 4 };  // exchange kinds of one context
enum { RED_MAX = 32 };  // values per window all-reduce (GMRES: j + 2 <= 32; fish moments: 29)

struct Comm {
  ncclComm_t nccl = nullptr;           // NCCL bootstrap (cup_comm_init); null with a host bootstrap
  CupAllgatherFn host_ag = nullptr;    // host bootstrap (cup_comm_init_host)
  void *host_user = nullptr;
  // ---- one-sided mode: every rank exposes one receive window through CUDA IPC ----
  bool p2p = false;
  char *win = nullptr;                 // [flags | all-reduce area | per-level receive areas]
  size_t flag_bytes = 0;               // bytes in front of the per-level areas
  std::vector<char *> peer_win;        // mapped base of every rank's window (own = win)
  char **d_peer_win = nullptr;
  unsigned long long *d_seq = nullptr; // [contexts][NKIND] exchange counters of THIS rank (face, restrict, prolong, blocks) + [1] all-reduce
  size_t red_flag_index = 0;           // index (in 8-byte words) of the [nranks] all-reduce flag words
  size_t red_data_off = 0;             // byte offset of the all-reduce data [2][nranks][RED_MAX] doubles
};

#define CUP_NCCL(call)                                                                      \
  do {                                                                                      \
    ncclResult_t r_ = (call);                                                               \
    if (r_ != ncclSuccess) {                                                                \
      cup::set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_));   \
      return CUP_ERR_NCCL;                                                                  \
    }                                                                                       \
  } while (0)

int comm_unique_id(void *out, size_t bytes) {
  if (bytes < sizeof(ncclUniqueId)) {
    set_error("cup_nccl_unique_id: need %zu bytes", sizeof(ncclUniqueId));
    return CUP_ERR_ARG;
  }
  CUP_TRY(load_nccl());
  ncclUniqueId id;
  CUP_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out, &id, sizeof id);
  return CUP_OK;
}

int comm_init(CupCtx *c, int rank, int nranks, const void *idp, size_t id_bytes) {
  if (nranks < 1 || rank < 0 || rank >= nranks) {
    set_error("cup_comm_init: rank %d of %d", rank, nranks);
    return CUP_ERR_ARG;
  }
  if (c->nblk != 0) {
    set_error("cup_comm_init must precede cup_mesh_upload");
    return CUP_ERR_STATE;
  }
  c->rank = rank;
  c->nranks = nranks;
  if (nranks == 1)
    return CUP_OK;
  if (!idp || id_bytes < sizeof(ncclUniqueId)) {
    set_error("cup_comm_init: ncclUniqueId (%zu bytes) required", sizeof(ncclUniqueId));
    return CUP_ERR_ARG;
  }
  CUP_CUDA(cudaSetDevice(c->device));
  CUP_TRY(load_nccl());
  Comm *cm = new Comm;
  ncclUniqueId id;
  memcpy(&id, idp, sizeof id);
  ncclResult_t r = g_nccl.CommInitRank(&cm->nccl, nranks, id, rank);
  if (r != ncclSuccess) {
    delete cm;
    set_error("ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    return CUP_ERR_NCCL;
  }
  c->comm = cm;
  return CUP_OK;
}

int comm_init_host(CupCtx *c, int rank, int nranks, CupAllgatherFn fn, void *user) {
  if (nranks < 1 || rank < 0 || rank >= nranks) {
    set_error("cup_comm_init_host: rank %d of %d", rank, nranks);
    return CUP_ERR_ARG;
  }
  if (c->nblk != 0) {
    set_error("cup_comm_init_host must precede cup_mesh_upload");
    return CUP_ERR_STATE;
  }
  if (nranks > 1 && !fn) {
    set_error("cup_comm_init_host: allgather callback required");
    return CUP_ERR_ARG;
  }
  c->rank = rank;
  c->nranks = nranks;
  if (nranks == 1)
    return CUP_OK;
  Comm *cm = new Comm;
  cm->host_ag = fn;
  cm->host_user = user;
  c->comm = cm;
  return CUP_OK;
}

void comm_free(CupCtx *c) {
  Comm *cm = (Comm *)c->comm;
  if (!cm)
    return;
  if (cm->nccl)
    g_nccl.CommDestroy(cm->nccl);
  delete cm;
  c->comm = nullptr;
}

// every rank contributes `bytes` bytes; recv = nranks * bytes in rank order (host memory)
static int allgather_bytes(CupCtx *c, const void *send, void *recv, size_t bytes) {
  Comm *cm = (Comm *)c->comm;
  const int R = c->nranks;
  if (cm->host_ag) {
    const int rc = cm->host_ag(cm->host_user, send, recv, bytes);
    if (rc != 0) {
      set_error("host allgather callback failed: %d", rc);
      return CUP_ERR_COMM;
    }
    return CUP_OK;
  }
  char *d;
  CUP_CUDA(cudaMalloc((void **)&d, bytes * (size_t)(R + 1)));
  auto run = [&]() -> int {
    CUP_CUDA(cudaMemcpyAsync(d + bytes * R, send, bytes, cudaMemcpyHostToDevice, c->stream));
    CUP_NCCL(g_nccl.AllGather(d + bytes * R, d, bytes, ncclChar, cm->nccl, c->stream));
    CUP_CUDA(cudaMemcpyAsync(recv, d, bytes * (size_t)R, cudaMemcpyDeviceToHost, c->stream));
    CUP_CUDA(cudaStreamSynchronize(c->stream));
    return CUP_OK;
  };
  const int rc = run();
  cudaFree(d);
  return rc;
}

int comm_check_error(CupCtx *c) {
  if (c->h_err && *(volatile int *)c->h_err != 0) {
    const int code = *(volatile int *)c->h_err;
    *(volatile int *)c->h_err = 0;
    set_error("rank %d: a peer did not post in time (exchange code %d: level %d, kind %d) -- peer lost or ranks "
              "out of step", c->rank, code, (code - 1) / 4, (code - 1) % 4);
    return CUP_ERR_COMM;
  }
  return CUP_OK;
}

// tree_sync: every rank contributes its blocks; result = global list in rank order + owner per block
int comm_gather_blocks(CupCtx *c, const CupBlk *blk, long long n, std::vector<CupBlk> &gblk,
                       std::vector<int> &owner) {
  if (c->nranks == 1) {
    gblk.assign(blk, blk + n);
    owner.assign((size_t)n, 0);
    return CUP_OK;
  }
  const int R = c->nranks;
  std::vector<long long> cnt((size_t)R);
  CUP_TRY(allgather_bytes(c, &n, cnt.data(), sizeof(long long)));
  long long mx = 0, tot = 0;
  for (int r = 0; r < R; r++) {
    mx = cnt[r] > mx ? cnt[r] : mx;
    tot += cnt[r];
  }
  const size_t chunk = (size_t)mx * sizeof(CupBlk);
  std::vector<char> mine(chunk, 0), h(chunk * (size_t)R);
  memcpy(mine.data(), blk, (size_t)n * sizeof(CupBlk));
  CUP_TRY(allgather_bytes(c, mine.data(), h.data(), chunk));
  gblk.clear();
  owner.clear();
  gblk.reserve((size_t)tot);
  for (int r = 0; r < R; r++) {
    const CupBlk *p = (const CupBlk *)(h.data() + chunk * r);
    for (long long i = 0; i < cnt[r]; i++) {
      gblk.push_back(p[i]);
      owner.push_back(r);
    }
  }
  return CUP_OK;
}

// All-reduce of a few device scalars through the peer windows: every rank stores its values into
// every rank's window (parity-double-buffered), publishes a flag, waits for the others' flags and
// reduces in RANK ORDER -- the result is bitwise identical on all ranks and independent of timing.
// One small kernel, no NCCL kernel, replayable from a CUDA graph.
__global__ void __launch_bounds__(256) k_allreduce_win(double *vals, int n, int is_max, unsigned long long *seq,
                                                       char *const *peer_win, int me, int R, size_t flag_index,
                                                       size_t data_off, int *err) {
  __shared__ unsigned long long s_seq;
  const int t = threadIdx.x;
  if (t == 0)
    s_seq = *(volatile unsigned long long *)seq + 1;
  __syncthreads();
  const unsigned long long s = s_seq;
  const size_t par = (size_t)(s & 1);
  for (int idx = t; idx < n * R; idx += blockDim.x) {
    const int p = idx / n, j = idx - p * n;
    volatile double *dst = (volatile double *)(peer_win[p] + data_off) + ((par * R + me) * RED_MAX + j);
    *dst = vals[j];
  }
  __threadfence_system();
  __syncthreads();
  if (t < R && t != me) {
    volatile unsigned long long *f = (volatile unsigned long long *)peer_win[t] + flag_index + me;
    *f = s;
    const volatile unsigned long long *mine = (const volatile unsigned long long *)peer_win[me] + flag_index + t;
    int spins = 0;
    while (*mine < s) {
      __nanosleep(100);
      if (++spins > COMM_SPIN_MAX) {
        *(volatile int *)err = 4000;
        break;
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (t < n) {
    const double *src = (const double *)(peer_win[me] + data_off) + (par * R) * RED_MAX + t;
    double acc = __ldcg(src);
    for (int r = 1; r < R; r++) {
      const double v = __ldcg(src + (size_t)r * RED_MAX);
      acc = is_max ? (acc < v ? v : acc) : acc + v;
    }
    vals[t] = acc;
  }
  if (t == 0)
    *(volatile unsigned long long *)seq = s;
}

static int allreduce_impl(CupCtx *c, int first, int n, bool is_max) {
  if (c->nranks == 1)
    return CUP_OK;
  Comm *cm = (Comm *)c->comm;
  if (cm->p2p) {
    for (int o = 0; o < n; o += RED_MAX) {
      const int m = n - o < RED_MAX ? n - o : RED_MAX;
      k_allreduce_win<<<1, 256, 0, c->stream>>>(c->d_scal + first + o, m, is_max ? 1 : 0,
                                                cm->d_seq + (size_t)(c->top + 2) * NKIND, cm->d_peer_win, c->rank,
                                                c->nranks, cm->red_flag_index, cm->red_data_off, c->h_err);
      c->launches++;
    }
    CUP_CUDA(cudaGetLastError());
    return CUP_OK;
  }
  if (!cm->nccl) {
    set_error("all-reduce without peer windows needs the NCCL bootstrap (cup_comm_init)");
    return CUP_ERR_STATE;
  }
  CUP_NCCL(g_nccl.AllReduce(c->d_scal + first, c->d_scal + first, (size_t)n, ncclDouble, is_max ? ncclMax : ncclSum,
                            cm->nccl, c->stream));
  return CUP_OK;
}

int comm_allreduce(CupCtx *c, int first, int n) { return allreduce_impl(c, first, n, false); }

// ===========================================================================
// Exchanges.  Two transports behind one post/wait interface:
//
//  * one-sided (default): the producer kernel stores straight into the
//    consumer's receive window over NVLink (CUDA-IPC mapped peer memory); its
//    LAST CTA to retire bumps the (context, kind) sequence number and writes it
//    into the peers' flag words (comm_dev.cuh: comm_post_at_exit), and the
//    consumer kernel waits for those flags itself before its first read of
//    received data (comm_wait_cta) -- no signal or wait kernels in between
//    (k_signal / k_wait below remain for posts and waits that have no kernel
//    to ride on).  No NCCL kernel has to find room on SMs that the persistent
//    sweep kernels occupy, so the exchange really overlaps the interior sweep.
//    Areas are double buffered by sequence parity.  Everything is replayable
//    from a CUDA graph: epochs live in device memory.
//  * NCCL (CUP_P2P=0 or IPC unavailable): pack to a staging buffer, grouped
//    ncclSend/ncclRecv per peer.
// ===========================================================================

int comm_allreduce_max(CupCtx *c, int first, int n) { return allreduce_impl(c, first, n, true); }

// grouped point-to-point exchange: entries of `entry_bytes` bytes, peer-major on both sides
static int exchange(CupCtx *c, const void *sbuf, const std::vector<int> &scnt, void *rbuf,
                    const std::vector<int> &rcnt, size_t entry_bytes) {
  Comm *cm = (Comm *)c->comm;
  size_t so = 0, ro = 0;
  CUP_NCCL(g_nccl.GroupStart());
  for (int p = 0; p < c->nranks; p++) {
    if (scnt[p]) {
      CUP_NCCL(g_nccl.Send((const char *)sbuf + so, (size_t)scnt[p] * entry_bytes, ncclChar, p, cm->nccl, c->stream));
      so += (size_t)scnt[p] * entry_bytes;
    }
    if (rcnt[p]) {
      CUP_NCCL(g_nccl.Recv((char *)rbuf + ro, (size_t)rcnt[p] * entry_bytes, ncclChar, p, cm->nccl, c->stream));
      ro += (size_t)rcnt[p] * entry_bytes;
    }
  }
  CUP_NCCL(g_nccl.GroupEnd());
  return CUP_OK;
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
// one 64-thread CTA per face: plane p of block `slot`, element (a, c) = (t&7, t>>3) in the
// receiver's convention (mg_device.cuh load_halo): x planes (y,z), y planes (x,z), z planes (x,y).
// dst0/dst1: per-entry destination for even / odd sequence numbers (may be peer memory).
template <typename Real>
__global__ void __launch_bounds__(64) k_pack_faces(const int *__restrict__ sslot, const int *__restrict__ splane,
                                                   int n, SlotVec<Real> u, Real *const *__restrict__ dst0,
                                                   Real *const *__restrict__ dst1,
                                                   const unsigned long long *__restrict__ seq, PostDesc post) {
  const int t = threadIdx.x, a = t & 7, cc = t >> 3;
  // the exchange being posted is number *seq + 1 (the last CTA to retire bumps it, so every CTA reads the old value)
  const bool odd = seq ? ((*(const volatile unsigned long long *)seq + 1) & 1) : false;
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const Real *b = u.at(sslot[e]);
    const int p = splane[e], q = (p & 1) ? 7 : 0;
    int idx;
    if (p < 2)
      idx = (cc << 6) + (a << 3) + q;
    else if (p < 4)
      idx = (cc << 6) + (q << 3) + a;
    else
      idx = (q << 6) + t;
    (odd ? dst1 : dst0)[e][t] = b[idx];
  }
  comm_post_at_exit(post);
}

// ghost slabs of the stencil sweeps: for each face entry, `ncomp` components x `nlayer` planes
// counted inward from the face; entry layout [(comp*nlayer + layer)*64 + t], SLAB_PLANES*64 apart
template <typename Real>
__global__ void __launch_bounds__(64) k_pack_slabs(const int *__restrict__ sslot, const int *__restrict__ splane,
                                                   int n, SlabSrc<Real> src, int ncomp, int nlayer,
                                                   Real *const *__restrict__ dst0, Real *const *__restrict__ dst1,
                                                   const unsigned long long *__restrict__ seq, PostDesc post) {
  const int t = threadIdx.x, a = t & 7, cc = t >> 3;
  const bool odd = seq ? ((*(const volatile unsigned long long *)seq + 1) & 1) : false;
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const size_t base = (size_t)sslot[e] * 512;
    const int p = splane[e];
    Real *d = (odd ? dst1 : dst0)[e];
    for (int q = 0; q < ncomp; q++)
      for (int l = 0; l < nlayer; l++) {
        const int n0 = (p & 1) ? 7 - l : l;
        d[(q * nlayer + l) * 64 + t] = src.c[q][base + face_idx(p, n0, a, cc)];
      }
  }
  comm_post_at_exit(post);
}

// mg_put (main.c:4722): received 64 r + 64 u of a remote child -> the parent's octant
template <typename Real>
__global__ void __launch_bounds__(64) k_put(const int *__restrict__ rslot, const int *__restrict__ roct, int n,
                                            const Real *__restrict__ in, SlotVec<Real> f, SlotVec<Real> u,
                                            WaitDesc wait) {
  const int t = threadIdx.x, cx = t & 3, cy = (t >> 2) & 3, cz = t >> 4;
  comm_wait_cta(wait);  // the children's owners posted them (k_down's exit)
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const int o = roct[e], ps = rslot[e];
    const int pidx = ((4 * (o >> 2) + cz) << 6) + ((4 * ((o >> 1) & 1) + cy) << 3) + 4 * (o & 1) + cx;
    f.at(ps)[pidx] = ld_recv(in + (size_t)e * 128 + t);
    u.at(ps)[pidx] = ld_recv(in + (size_t)e * 128 + 64 + t);
  }
}

// mg_get (main.c:4771): u_c - us of the parent's octant, for a remote child
template <typename Real>
__global__ void __launch_bounds__(64) k_get(const int *__restrict__ rslot, const int *__restrict__ roct, int n,
                                            SlotVec<Real> u, SlotVec<Real> us, Real *const *__restrict__ dst,
                                            PostDesc post) {
  const int t = threadIdx.x, cx = t & 3, cy = (t >> 2) & 3, cz = t >> 4;
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const int o = roct[e], ps = rslot[e];
    const int pidx = ((4 * (o >> 2) + cz) << 6) + ((4 * ((o >> 1) & 1) + cy) << 3) + 4 * (o & 1) + cx;
    dst[e][t] = u.at(ps)[pidx] - us.at(ps)[pidx];
  }
  comm_post_at_exit(post);
}

// ---------------------------------------------------------------------------
// ghost BLOCKS (levels with coarse-fine interfaces and the leaf context of multi-level meshes):
// whole 8^3 blocks, as the reference's halo_sync ships them (main.c:3101-3112)
// ---------------------------------------------------------------------------
// multigrid context: entry e = block sslot[e] of the swept vector (kind 0) or of the canonical U0
// (kind 1: a coarser leaf behind an interface)
template <typename Real>
__global__ void __launch_bounds__(128) k_pack_blocks_mg(const int *__restrict__ sslot, const int *__restrict__ skind,
                                                        int n, SlotVec<Real> same, SlotVec<Real> can,
                                                        Real *const *__restrict__ dst0, Real *const *__restrict__ dst1,
                                                        const unsigned long long *__restrict__ seq, PostDesc post) {
  const bool odd = ((*(const volatile unsigned long long *)seq + 1) & 1) != 0;
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const Real *src = (skind[e] ? can : same).at(sslot[e]);
    Real *d = (odd ? dst1 : dst0)[e];
    for (int j = threadIdx.x; j < 512; j += blockDim.x)
      d[j] = src[j];
  }
  comm_post_at_exit(post);
}

template <typename Real>
__global__ void __launch_bounds__(128) k_unpack_blocks_mg(const int *__restrict__ rslot, const int *__restrict__ rkind,
                                                          int n, const Real *__restrict__ area, long long stride,
                                                          const unsigned long long *__restrict__ seq,
                                                          SlotVec<Real> same, SlotVec<Real> can, WaitDesc wait) {
  comm_wait_cta(wait);
  const Real *in = area + ((*(const volatile unsigned long long *)seq & 1) ? stride : 0);
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    Real *d = (rkind[e] ? can : same).at(rslot[e]);
    const Real *q = in + (size_t)e * 512;
    for (int j = threadIdx.x; j < 512; j += blockDim.x)
      d[j] = ld_recv(q + j);
  }
}

// leaf context: ncomp flat component arrays, entry layout [comp][512]
template <typename Real>
struct BlkComps {
  const Real *src[BLK_COMPS];
  Real *dst[BLK_COMPS];
};

template <typename Real>
__global__ void __launch_bounds__(128) k_pack_blocks_leaf(const int *__restrict__ sslot, int n, int ncomp, int stride_c,
                                                          BlkComps<Real> cp, Real *const *__restrict__ dst0,
                                                          Real *const *__restrict__ dst1,
                                                          const unsigned long long *__restrict__ seq, PostDesc post) {
  const bool odd = ((*(const volatile unsigned long long *)seq + 1) & 1) != 0;
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    Real *d = (odd ? dst1 : dst0)[e];
    const size_t off = (size_t)sslot[e] * 512;
    for (int q = 0; q < ncomp; q++)
      for (int j = threadIdx.x; j < 512; j += blockDim.x)
        d[(size_t)q * stride_c + j] = cp.src[q][off + j];
  }
  comm_post_at_exit(post);
}

template <typename Real>
__global__ void __launch_bounds__(128) k_unpack_blocks_leaf(const int *__restrict__ rslot, int n, int ncomp,
                                                            int stride_c, int entry_reals,
                                                            const Real *__restrict__ area, long long stride,
                                                            const unsigned long long *__restrict__ seq,
                                                            BlkComps<Real> cp, long long dst_off, WaitDesc wait) {
  comm_wait_cta(wait);
  const Real *in = area + ((*(const volatile unsigned long long *)seq & 1) ? stride : 0);
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const size_t off = (size_t)((long long)rslot[e] - dst_off) * 512;
    const Real *q0 = in + (size_t)e * entry_reals;
    for (int q = 0; q < ncomp; q++)
      for (int j = threadIdx.x; j < 512; j += blockDim.x)
        cp.dst[q][off + j] = ld_recv(q0 + (size_t)q * stride_c + j);
  }
}

// bump this rank's sequence number for (level, kind) and publish it to the peers' flag words
__global__ void __launch_bounds__(64) k_signal(unsigned long long *seq, char *const *peer_win, const int *peers, int np,
                                               size_t flag_index) {
  __shared__ unsigned long long s;
  if (threadIdx.x == 0) {
    s = *seq + 1;
    *seq = s;
  }
  __syncthreads();
  __threadfence_system();  // the pack kernel's stores (previous kernel) are visible before the flag
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    volatile unsigned long long *f = (volatile unsigned long long *)peer_win[peers[i]] + flag_index;
    *f = s;
  }
}

// wait until every listed peer has published at least this rank's current sequence number
__global__ void __launch_bounds__(64) k_wait(WaitDesc w) {
  // several seconds without an answer: a lost peer is an ERROR reported to the host, not a hang or a trap
  comm_wait_cta(w);
}

static inline int cgrid(const CupCtx *c, int n) {
  int g = c->num_sms * 16;
  return n < g ? (n < 1 ? 1 : n) : g;
}

static int sum(const std::vector<int> &v) {
  int s = 0;
  for (int x : v)
    s += x;
  return s;
}

template <typename T>
static int up(T **d, const std::vector<T> &h) {
  *d = nullptr;
  if (h.empty())
    return CUP_OK;
  CUP_CUDA(cudaMalloc((void **)d, h.size() * sizeof(T)));
  CUP_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return CUP_OK;
}

static std::vector<int> peers_of(const std::vector<int> &cnt) {
  std::vector<int> p;
  for (size_t i = 0; i < cnt.size(); i++)
    if (cnt[i])
      p.push_back((int)i);
  return p;
}

void comm_free_level_buffers(CupCtx *c) {
  Comm *cm = (Comm *)c->comm;
  if (cm && cm->p2p) {
    // nobody may still be writing into a window that is about to disappear: local work done, then a barrier
    cudaStreamSynchronize(c->stream);
    char one = 0;
    std::vector<char> all((size_t)c->nranks);
    allgather_bytes(c, &one, all.data(), 1);
    for (int p = 0; p < c->nranks; p++)
      if (p != c->rank && cm->peer_win.size() > (size_t)p && cm->peer_win[p])
        cudaIpcCloseMemHandle(cm->peer_win[p]);
    cm->peer_win.clear();
    // ... and nobody may still have this rank's window mapped when it is freed
    allgather_bytes(c, &one, all.data(), 1);
    cudaFree(cm->win);
    cudaFree(cm->d_peer_win);
    cudaFree(cm->d_seq);
    cm->win = nullptr;
    cm->d_peer_win = nullptr;
    cm->d_seq = nullptr;
    cm->p2p = false;
    for (auto &v : c->lv)
      v.d_frecv = v.d_rrecv = v.d_precv = v.d_srecv = nullptr;  // lived inside the window
  }
  for (auto &v : c->lv) {
    cudaFree(v.d_fsend);
    cudaFree(v.d_frecv);
    cudaFree(v.d_rsend);
    cudaFree(v.d_rrecv);
    cudaFree(v.d_ssend);
    cudaFree(v.d_srecv);
    cudaFree(v.d_sptr0);
    cudaFree(v.d_sptr1);
    v.d_ssend = v.d_srecv = nullptr;
    v.d_sptr0 = v.d_sptr1 = nullptr;
    cudaFree(v.d_order);
    cudaFree(v.d_bsend);
    cudaFree(v.d_counters);
    v.d_order = v.d_bsend = nullptr;
    v.d_counters = nullptr;
    cudaFree(v.d_fptr0);
    cudaFree(v.d_fptr1);
    cudaFree(v.d_rptr);
    cudaFree(v.d_pptr);
    cudaFree(v.d_bptr0);
    cudaFree(v.d_bptr1);
    cudaFree(v.d_blk_speers);
    cudaFree(v.d_blk_rpeers);
    v.d_bptr0 = v.d_bptr1 = nullptr;
    v.d_blk_speers = v.d_blk_rpeers = nullptr;
    v.d_brecv = nullptr;
    for (int k = 0; k < 3; k++) {
      cudaFree(v.d_speers[k]);
      cudaFree(v.d_rpeers[k]);
      v.d_speers[k] = v.d_rpeers[k] = nullptr;
    }
    v.d_fsend = v.d_frecv = v.d_rsend = v.d_rrecv = v.d_precv = nullptr;
    v.d_fptr0 = v.d_fptr1 = v.d_rptr = v.d_pptr = nullptr;
  }
  {
    Level &v = c->leafv;  // the leaf context only has ghost blocks
    cudaFree(v.d_bptr0);
    cudaFree(v.d_bptr1);
    cudaFree(v.d_counters);
    cudaFree(v.d_blk_speers);
    cudaFree(v.d_blk_rpeers);
    v.d_bptr0 = v.d_bptr1 = nullptr;
    v.d_counters = nullptr;
    v.d_blk_speers = v.d_blk_rpeers = nullptr;
    v.d_brecv = nullptr;
    v.p2p = false;
  }
}

static bool want_p2p() {
  const char *e = getenv("CUP_P2P");
  return !(e && atoi(e) == 0);
}

// open every rank's receive window; *ok = false if CUDA IPC is not usable here
static int open_windows(CupCtx *c, bool *ok) {
  Comm *cm = (Comm *)c->comm;
  const int R = c->nranks;
  const size_t rb = (size_t)c->real_bytes;
  *ok = false;
  // [context][kind][rank] exchange flags | [rank] all-reduce flags | all-reduce data [2][R][RED_MAX];
  // contexts: the multigrid levels 0..top and the leaf context (top + 1)
  const size_t nflag = (size_t)(c->top + 2) * NKIND * R;
  cm->red_flag_index = nflag;
  cm->red_data_off = ((nflag + (size_t)R) * sizeof(unsigned long long) + 255) / 256 * 256;
  cm->flag_bytes = (cm->red_data_off + (size_t)2 * R * RED_MAX * sizeof(double) + 255) / 256 * 256;
  const size_t bytes = cm->flag_bytes + (size_t)c->win_reals[c->rank] * rb + 256;
  CUP_CUDA(cudaMalloc((void **)&cm->win, bytes));
  CUP_CUDA(cudaMemset(cm->win, 0, bytes));
  CUP_CUDA(cudaDeviceSynchronize());
  // all-gather the IPC handles (and a usable flag)
  struct Msg {
    cudaIpcMemHandle_t h;
    int ok, pad[15];
  };
  Msg mine;
  memset(&mine, 0, sizeof mine);
  mine.ok = cudaIpcGetMemHandle(&mine.h, cm->win) == cudaSuccess;
  cudaGetLastError();
  std::vector<Msg> all((size_t)R);
  CUP_TRY(allgather_bytes(c, &mine, all.data(), sizeof(Msg)));
  bool good = true;
  for (int p = 0; p < R; p++)
    good = good && all[p].ok;
  cm->peer_win.assign((size_t)R, nullptr);
  cm->peer_win[c->rank] = cm->win;
  if (good)
    for (int p = 0; p < R; p++) {
      if (p == c->rank)
        continue;
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[p].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        good = false;
        break;
      }
      cm->peer_win[p] = (char *)ptr;
    }
  // everyone must agree on the transport
  char flag = good ? 0 : 1;
  std::vector<char> flags((size_t)R);
  CUP_TRY(allgather_bytes(c, &flag, flags.data(), 1));
  bool all_good = true;
  for (int p = 0; p < R; p++)
    all_good = all_good && flags[p] == 0;
  if (!all_good) {
    for (int p = 0; p < R; p++)
      if (p != c->rank && cm->peer_win[p])
        cudaIpcCloseMemHandle(cm->peer_win[p]);
    cm->peer_win.clear();
    cudaFree(cm->win);
    cm->win = nullptr;
    return CUP_OK;
  }
  CUP_TRY(up(&cm->d_peer_win, cm->peer_win));
  const size_t nseq = (size_t)(c->top + 2) * NKIND + 1;
  CUP_CUDA(cudaMalloc((void **)&cm->d_seq, nseq * sizeof(unsigned long long)));
  CUP_CUDA(cudaMemset(cm->d_seq, 0, nseq * sizeof(unsigned long long)));
  *ok = true;
  return CUP_OK;
}

// ghost blocks of one context: own area inside the window, destination of every send entry, peers
static int setup_ghost_blocks(CupCtx *c, Level &v, bool p2p, bool is_leaf) {
  Comm *cm = (Comm *)c->comm;
  if (!v.ghosted)
    return CUP_OK;
  if (!p2p) {
    set_error("multi-level meshes across ranks need the peer-window transport (CUDA IPC); CUP_P2P=0 / NCCL staging "
              "is implemented for uniform meshes only");
    return CUP_ERR_UNSUPPORTED;
  }
  const size_t rb = (size_t)c->real_bytes;
  const int me = c->rank;
  const size_t ent = (size_t)512 * v.blk_ncomp;
  const size_t ns = v.blk_sslot.size();
  std::vector<char *> b0(ns), b1(ns);
  v.d_brecv = cm->win + cm->flag_bytes + (size_t)v.win_blk[me] * rb;
  v.blk_stride = (long long)v.win_nblk[me] * (long long)ent;
  for (size_t e = 0; e < ns; e++) {
    const int p = v.blk_speer[e];
    b0[e] = cm->peer_win[p] + cm->flag_bytes + ((size_t)v.win_blk[p] + (size_t)v.blk_sidx[e] * ent) * rb;
    b1[e] = b0[e] + (size_t)v.win_nblk[p] * ent * rb;
  }
  CUP_TRY(up((char ***)&v.d_bptr0, b0));
  CUP_TRY(up((char ***)&v.d_bptr1, b1));
  // Who reads whose blocks is NOT symmetric (a fine block reads the coarser leaf behind an interface, the
  // coarse leaf's owner may read nothing back), but the double-buffered areas rely on "nobody is more than
  // one exchange ahead of a peer": every rank therefore notifies AND awaits every rank it exchanges blocks
  // with in either direction.
  {
    std::vector<int> both((size_t)c->nranks, 0);
    for (int p = 0; p < c->nranks; p++)
      both[(size_t)p] = (v.blk_scnt[(size_t)p] > 0 || v.blk_rcnt[(size_t)p] > 0) ? 1 : 0;
    v.blk_speers = peers_of(both);
    v.blk_rpeers = v.blk_speers;
  }
  CUP_TRY(up(&v.d_blk_speers, v.blk_speers));
  CUP_TRY(up(&v.d_blk_rpeers, v.blk_rpeers));
  if (is_leaf) {
    CUP_CUDA(cudaMalloc((void **)&v.d_counters, 8 * sizeof(unsigned int)));
    CUP_CUDA(cudaMemset(v.d_counters, 0, 8 * sizeof(unsigned int)));
    v.p2p = true;
    v.d_seq = cm->d_seq + (size_t)v.xid * NKIND;
  }
  return CUP_OK;
}

// scratch of one level + destination pointers of every entry
int comm_alloc_level_buffers(CupCtx *c) {
  const size_t rb = (size_t)c->real_bytes;
  comm_free_level_buffers(c);
  if (c->nranks == 1)
    return CUP_OK;
  Comm *cm = (Comm *)c->comm;
  if (!cm) {
    set_error("rank %d of %d: cup_comm_init / cup_comm_init_host was not called", c->rank, c->nranks);
    return CUP_ERR_STATE;
  }
  if (!c->h_err) {
    CUP_CUDA(cudaHostAlloc((void **)&c->h_err, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
    *c->h_err = 0;
  }
  bool p2p = false;
  if (want_p2p())
    CUP_TRY(open_windows(c, &p2p));
  if (!p2p && !cm->nccl) {
    set_error("the host-bootstrapped transport needs CUDA IPC peer windows between all ranks (not available here%s)",
              want_p2p() ? "" : ": CUP_P2P=0");
    return CUP_ERR_UNSUPPORTED;
  }
  cm->p2p = p2p;
  const int me = c->rank;
  c->leafv.xid = c->top + 1;
  CUP_TRY(setup_ghost_blocks(c, c->leafv, p2p, true));
  for (auto &v : c->lv) {
    v.xid = v.L;
    const size_t ns = v.face_sslot.size(), nr = (size_t)v.nface_recv;
    const size_t cs = (size_t)sum(v.res_scnt), cr = (size_t)sum(v.res_rcnt);
    std::vector<char *> f0(ns), f1(ns), rp(cs), pp(cr);
    if (p2p) {
      char *base = cm->win + cm->flag_bytes;
      v.d_frecv = base + (size_t)v.win_face[me] * rb;
      v.d_rrecv = base + (size_t)v.win_res[me] * rb;
      v.d_precv = base + (size_t)v.win_pro[me] * rb;
      for (size_t e = 0; e < ns; e++) {
        const int p = v.face_speer[e];
        char *pb = cm->peer_win[p] + cm->flag_bytes + (size_t)v.win_face[p] * rb;
        f0[e] = pb + (size_t)v.face_sidx[e] * 64 * rb;
        f1[e] = f0[e] + (size_t)v.win_nrecv[p] * 64 * rb;
      }
      for (size_t q = 0; q < cs; q++) {
        const int p = v.res_speer[q];
        rp[q] = cm->peer_win[p] + cm->flag_bytes + ((size_t)v.win_res[p] + (size_t)v.res_sidx[q] * 128) * rb;
      }
      for (size_t e = 0; e < cr; e++) {
        const int p = v.pro_speer[e];
        pp[e] = cm->peer_win[p] + cm->flag_bytes + ((size_t)v.win_pro[p] + (size_t)v.pro_sidx[e] * 64) * rb;
      }
    } else {
      if (ns)
        CUP_CUDA(cudaMalloc(&v.d_fsend, ns * 64 * rb));
      if (nr)
        CUP_CUDA(cudaMalloc(&v.d_frecv, nr * 64 * rb));
      if (cs)
        CUP_CUDA(cudaMalloc(&v.d_rsend, cs * 128 * rb));
      if (cr)
        CUP_CUDA(cudaMalloc(&v.d_rrecv, cr * 128 * rb));
      v.d_precv = v.d_rsend;  // prolongation arrives where the restriction was staged
      for (size_t e = 0; e < ns; e++)
        f0[e] = f1[e] = (char *)v.d_fsend + e * 64 * rb;
      for (size_t q = 0; q < cs; q++)
        rp[q] = (char *)v.d_rsend + q * 128 * rb;
      for (size_t e = 0; e < cr; e++)
        pp[e] = (char *)v.d_rrecv + e * 64 * rb;
    }
    // ghost slabs (stencil sweeps): same entries, SLAB_PLANES planes each
    if (v.win_slab.size() == (size_t)c->nranks && v.win_slab[me] >= 0 && (ns || nr)) {
      std::vector<char *> s0(ns), s1(ns);
      const size_t eb = 64 * SLAB_PLANES * rb;
      if (p2p) {
        v.d_srecv = cm->win + cm->flag_bytes + (size_t)v.win_slab[me] * rb;
        v.slab_stride = (long long)nr * 64 * SLAB_PLANES;
        for (size_t e = 0; e < ns; e++) {
          const int p = v.face_speer[e];
          s0[e] = cm->peer_win[p] + cm->flag_bytes + (size_t)v.win_slab[p] * rb + (size_t)v.face_sidx[e] * eb;
          s1[e] = s0[e] + (size_t)v.win_nrecv[p] * eb;
        }
      } else {
        if (ns)
          CUP_CUDA(cudaMalloc(&v.d_ssend, ns * eb));
        if (nr)
          CUP_CUDA(cudaMalloc(&v.d_srecv, nr * eb));
        v.slab_stride = 0;
        for (size_t e = 0; e < ns; e++)
          s0[e] = s1[e] = (char *)v.d_ssend + e * eb;
      }
      CUP_TRY(up((char ***)&v.d_sptr0, s0));
      CUP_TRY(up((char ***)&v.d_sptr1, s1));
    }
    CUP_TRY(up((char ***)&v.d_fptr0, f0));
    CUP_TRY(up((char ***)&v.d_fptr1, f1));
    CUP_TRY(up((char ***)&v.d_rptr, rp));
    CUP_TRY(up((char ***)&v.d_pptr, pp));
    v.speers[K_FACE] = peers_of(v.face_scnt);
    v.rpeers[K_FACE] = peers_of(v.face_rcnt);
    v.speers[K_RES] = peers_of(v.res_scnt);
    v.rpeers[K_RES] = peers_of(v.res_rcnt);
    v.speers[K_PRO] = v.rpeers[K_RES];  // corrections flow back along the same edges
    v.rpeers[K_PRO] = v.speers[K_RES];
    for (int k = 0; k < 3; k++) {
      CUP_TRY(up(&v.d_speers[k], v.speers[k]));
      CUP_TRY(up(&v.d_rpeers[k], v.rpeers[k]));
    }
    if (p2p) {
      CUP_CUDA(cudaMalloc((void **)&v.d_counters, 8 * sizeof(unsigned int)));
      CUP_CUDA(cudaMemset(v.d_counters, 0, 8 * sizeof(unsigned int)));
    }
    if (p2p && !v.bnd.empty()) {
      std::map<std::pair<int, int>, int> ent;
      for (size_t e = 0; e < ns; e++)
        ent[{v.face_sslot[e], v.face_splane[e]}] = (int)e;
      v.bsend.assign(v.act.size() * 6, -1);
      for (size_t k = 0; k < v.act.size(); k++)
        for (int f = 0; f < 6; f++)
          if (v.nbr[k * 6 + f] <= NBR_REMOTE0)
            v.bsend[k * 6 + f] = ent.at({v.act[k], f});
      v.order = v.bnd;
      v.order.insert(v.order.end(), v.inner.begin(), v.inner.end());
      CUP_TRY(up(&v.d_bsend, v.bsend));
      CUP_TRY(up(&v.d_order, v.order));
    }
    v.p2p = p2p;
    v.rface_stride = p2p ? (long long)nr * 64 : 0;
    v.d_seq = p2p ? cm->d_seq + (size_t)v.xid * NKIND : nullptr;
    CUP_TRY(setup_ghost_blocks(c, v, p2p, false));
  }
  return CUP_OK;
}

// ---- descriptors for kernels that wait / post themselves (one-sided transport only) ----
enum { CNT_PACK = 2, CNT_DOWN = 3, CNT_GET = 4, CNT_UP = 5, CNT_SLAB = 6 };

static WaitDesc make_wait(CupCtx *c, const Level &v, int kind) {
  WaitDesc w;
  Comm *cm = (Comm *)c->comm;
  if (!cm || !v.p2p || v.rpeers[kind].empty())
    return w;
  w.seq = cm->d_seq + (size_t)v.xid * NKIND + kind;
  w.flags = (const unsigned long long *)cm->win + ((size_t)v.xid * NKIND + kind) * c->nranks;
  w.peers = v.d_rpeers[kind];
  w.np = (int)v.rpeers[kind].size();
  w.err = c->h_err;
  w.code = 1 + v.xid * 4 + kind;
  return w;
}

static PostDesc make_post(CupCtx *c, const Level &v, int kind, int counter) {
  PostDesc p;
  Comm *cm = (Comm *)c->comm;
  if (!cm || !v.p2p)
    return p;
  p.seq = cm->d_seq + (size_t)v.xid * NKIND + kind;
  p.peer_win = cm->d_peer_win;
  p.peers = v.d_speers[kind];
  p.np = (int)v.speers[kind].size();
  p.flag_index = ((size_t)v.xid * NKIND + kind) * c->nranks + c->rank;
  p.counter = v.d_counters + counter;
  return p;
}

// bump the sequence number of an exchange this rank takes part in without sending anything
static int post_empty(CupCtx *c, Level &v, int kind) {
  Comm *cm = (Comm *)c->comm;
  const size_t fidx = ((size_t)v.xid * NKIND + kind) * c->nranks + c->rank;
  k_signal<<<1, 64, 0, c->stream>>>(cm->d_seq + (size_t)v.xid * NKIND + kind, cm->d_peer_win, v.d_speers[kind],
                                    (int)v.speers[kind].size(), fidx);
  c->launches++;
  return CUP_OK;
}

static int wait(CupCtx *c, Level &v, int kind) {
  const WaitDesc w = make_wait(c, v, kind);
  if (!w.seq)
    return CUP_OK;
  k_wait<<<1, 64, 0, c->stream>>>(w);
  c->launches++;
  return CUP_OK;
}

static bool level_has_faces(const CupCtx *c, const Level &v) {
  return c->nranks > 1 && (!v.face_sslot.empty() || v.nface_recv > 0);
}

bool comm_fused_desc(CupCtx *c, Level &v, FusedComm *out) {
  Comm *cm = (Comm *)c->comm;
  if (!cm || !v.p2p || v.bnd.empty() || !v.d_order)
    return false;
  out->bsend = v.d_bsend;
  out->fptr0 = v.d_fptr0;
  out->fptr1 = v.d_fptr1;
  out->seq = cm->d_seq + (size_t)v.xid * NKIND + K_FACE;
  out->my_flags = (const unsigned long long *)cm->win + ((size_t)v.xid * NKIND + K_FACE) * c->nranks;
  out->rpeers = v.d_rpeers[K_FACE];
  out->nrp = (int)v.rpeers[K_FACE].size();
  out->peer_win = cm->d_peer_win;
  out->speers = v.d_speers[K_FACE];
  out->nsp = (int)v.speers[K_FACE].size();
  out->flag_index = ((size_t)v.xid * NKIND + K_FACE) * c->nranks + c->rank;
  out->counters = v.d_counters;
  out->nbnd = (int)v.bnd.size();
  out->err = c->h_err;
  out->code = 1 + v.xid * 4 + K_FACE;
  {
    // how boundary planes leave the SM (CUP_PUSH): 1 = gathered in shared memory, one coalesced 64-element
    // store of the CTA per plane (default); 2 = TMA bulk stores from that staging area; 0 = each thread's own
    // words.  Measured at 2 GPUs (r02, 512^3 cycle): 2.48 / 2.57 ms for 1 / 2; at 8 GPUs 1.05 ms with 0.
    static int mode = getenv("CUP_PUSH") ? atoi(getenv("CUP_PUSH")) : 1;
    out->push_mode = mode;
  }
  return true;
}

// what a kernel that consumes / produces an exchange itself needs (empty descriptors elsewhere)
WaitDesc comm_wait_desc(CupCtx *c, Level &v, int kind) { return make_wait(c, v, kind); }
PostDesc comm_post_desc(CupCtx *c, Level &v, int kind) {
  // only kinds whose producer is a single kernel: restriction (k_down) and the faces after k_up
  if (kind == K_RES && sum(v.res_scnt) == 0)
    return PostDesc{};
  if (kind == K_FACE && !level_has_faces(c, v))
    return PostDesc{};
  return make_post(c, v, kind, kind == K_RES ? CNT_DOWN : CNT_UP);
}

// publish the ghost faces of `u` (this rank's boundary blocks of level v) to their consumers
template <typename Real>
int halo_post(CupCtx *c, Level &v, SlotVec<Real> u) {
  if (!level_has_faces(c, v))
    return CUP_OK;
  const int ns = (int)v.face_sslot.size();
  if (ns) {
    // one-sided: the pack kernel stores into the peers' windows and its last CTA publishes
    k_pack_faces<Real><<<cgrid(c, ns), 64, 0, c->stream>>>(v.d_face_sslot, v.d_face_splane, ns, u,
                                                           (Real *const *)v.d_fptr0, (Real *const *)v.d_fptr1,
                                                           (const unsigned long long *)v.d_seq,
                                                           make_post(c, v, K_FACE, CNT_PACK));
    c->launches++;
  } else if (v.p2p) {
    CUP_TRY(post_empty(c, v, K_FACE));
  }
  if (!v.p2p)
    CUP_TRY(exchange(c, v.d_fsend, v.face_scnt, v.d_frecv, v.face_rcnt, 64 * sizeof(Real)));
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

// publish `ncomp` x `nlayer` ghost planes per face of the leaf level (stencil sweeps) and wait for
// the peers' ones; afterwards they sit in v.d_srecv (+ parity stride)
template <typename Real>
int slab_exchange(CupCtx *c, Level &v, const SlabSrc<Real> &src, int ncomp, int nlayer) {
  if (!level_has_faces(c, v))
    return CUP_OK;
  if (ncomp * nlayer > SLAB_PLANES || !v.d_sptr0) {
    set_error("slab_exchange: %d x %d planes not provisioned", ncomp, nlayer);
    return CUP_ERR_STATE;
  }
  const int ns = (int)v.face_sslot.size();
  if (ns) {
    k_pack_slabs<Real><<<cgrid(c, ns), 64, 0, c->stream>>>(v.d_face_sslot, v.d_face_splane, ns, src, ncomp, nlayer,
                                                           (Real *const *)v.d_sptr0, (Real *const *)v.d_sptr1,
                                                           (const unsigned long long *)v.d_seq,
                                                           make_post(c, v, K_FACE, CNT_SLAB));
    c->launches++;
  } else if (v.p2p) {
    CUP_TRY(post_empty(c, v, K_FACE));
  }
  if (v.p2p)
    CUP_TRY(wait(c, v, K_FACE));
  else
    CUP_TRY(exchange(c, v.d_ssend, v.face_scnt, v.d_srecv, v.face_rcnt, 64 * SLAB_PLANES * sizeof(Real)));
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

// the faces posted last are complete in this rank's receive area (stand-alone wait kernel: for
// consumers that do not wait themselves)
int halo_wait(CupCtx *c, Level &v) {
  if (!level_has_faces(c, v) || !v.p2p)
    return CUP_OK;
  return wait(c, v, K_FACE);
}

// after k_down stored the children with remote parents through v.d_rptr.  `posted`: the kernel
// published the exchange itself (comm_post_desc(K_RES) was non-empty)
template <typename Real>
int restrict_exchange(CupCtx *c, Level &v, SlotVec<Real> f, SlotVec<Real> u, bool posted) {
  if (c->nranks == 1)
    return CUP_OK;
  const int cr = sum(v.res_rcnt), cs = sum(v.res_scnt);
  if (cr == 0 && cs == 0)
    return CUP_OK;
  if (v.p2p) {
    if (!posted)
      CUP_TRY(post_empty(c, v, K_RES));  // also the plain signal when k_down ran without a descriptor
  } else {
    CUP_TRY(exchange(c, v.d_rsend, v.res_scnt, v.d_rrecv, v.res_rcnt, 128 * sizeof(Real)));
  }
  if (cr) {
    k_put<Real><<<cgrid(c, cr), 64, 0, c->stream>>>(v.d_res_rslot, v.d_res_roct, cr, (const Real *)v.d_rrecv, f, u,
                                                    make_wait(c, v, K_RES));
    c->launches++;
  }
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

// before k_up: parents' owners send u_c - us to remote children; lands in v.d_precv.  In one-sided
// mode the consumer (k_up) waits itself with comm_wait_desc(K_PRO).
template <typename Real>
int prolong_exchange(CupCtx *c, Level &v, SlotVec<Real> u, SlotVec<Real> us) {
  if (c->nranks == 1)
    return CUP_OK;
  const int cr = sum(v.res_rcnt), cs = sum(v.res_scnt);
  if (cr == 0 && cs == 0)
    return CUP_OK;
  if (cr) {
    k_get<Real><<<cgrid(c, cr), 64, 0, c->stream>>>(v.d_res_rslot, v.d_res_roct, cr, u, us, (Real *const *)v.d_pptr,
                                                    make_post(c, v, K_PRO, CNT_GET));
    c->launches++;
  } else if (v.p2p) {
    CUP_TRY(post_empty(c, v, K_PRO));
  }
  if (!v.p2p)
    CUP_TRY(exchange(c, v.d_rrecv, v.res_rcnt, v.d_precv, v.res_scnt, 64 * sizeof(Real)));
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

// ---- ghost blocks --------------------------------------------------------------------------
static void blk_descs(CupCtx *c, Level &v, PostDesc *post, WaitDesc *wait) {
  Comm *cm = (Comm *)c->comm;
  post->peers = v.d_blk_speers;
  post->np = (int)v.blk_speers.size();
  wait->peers = v.d_blk_rpeers;
  wait->np = (int)v.blk_rpeers.size();
  post->seq = cm->d_seq + (size_t)v.xid * NKIND + K_BLK;
  post->peer_win = cm->d_peer_win;
  post->flag_index = ((size_t)v.xid * NKIND + K_BLK) * c->nranks + c->rank;
  post->counter = v.d_counters + 7;
  wait->seq = post->seq;
  wait->flags = (const unsigned long long *)cm->win + ((size_t)v.xid * NKIND + K_BLK) * c->nranks;
  wait->err = c->h_err;
  wait->code = 1 + v.xid * 4 + K_BLK;
}

// multigrid level with coarse-fine interfaces: refresh the ghost copies of `same` (blocks of this level
// owned by other ranks) and of `can` (coarser leaves behind interfaces, canonical U0) before a sweep
template <typename Real>
int block_exchange_mg(CupCtx *c, Level &v, SlotVec<Real> same, SlotVec<Real> can) {
  if (c->nranks == 1 || !v.ghosted)
    return CUP_OK;
  Comm *cm = (Comm *)c->comm;
  PostDesc post;
  WaitDesc wait;
  blk_descs(c, v, &post, &wait);
  const unsigned long long *seq = cm->d_seq + (size_t)v.xid * NKIND + K_BLK;
  const int ns = (int)v.blk_sslot.size(), nr = (int)v.blk_rslot.size();
  if (ns) {
    k_pack_blocks_mg<Real><<<cgrid(c, ns), 128, 0, c->stream>>>(v.d_blk_sslot, v.d_blk_skind, ns, same, can,
                                                               (Real *const *)v.d_bptr0, (Real *const *)v.d_bptr1, seq,
                                                               post);
    c->launches++;
  } else {
    k_signal<<<1, 64, 0, c->stream>>>(post.seq, cm->d_peer_win, post.peers, post.np, post.flag_index);
    c->launches++;
  }
  if (nr) {
    k_unpack_blocks_mg<Real><<<cgrid(c, nr), 128, 0, c->stream>>>(v.d_blk_rslot, v.d_blk_rkind, nr,
                                                                 (const Real *)v.d_brecv, v.blk_stride, seq, same, can,
                                                                 wait);
    c->launches++;
  } else if (wait.np > 0) {
    k_wait<<<1, 64, 0, c->stream>>>(wait);  // a pure sender still stays in step with its receivers
    c->launches++;
  }
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

// leaf context of a multi-level mesh: ghost blocks of ncomp component arrays.  src[q]: [nblk][512] (own
// leaves); dst[q]: where slot s >= nblk of component q lives is dst[q] + (s - dst_off) * 512 (dst_off = 0 for
// the state arrays, which hold nstate blocks; = nblk for the ghost scratch of a flat vector)
template <typename Real>
int block_exchange_leaf(CupCtx *c, const Real *const *src, Real *const *dst, int ncomp, long long dst_off) {
  Level &v = c->leafv;
  if (c->nranks == 1 || !v.ghosted)
    return CUP_OK;
  if (ncomp > BLK_COMPS) {
    set_error("block_exchange_leaf: %d components > %d", ncomp, (int)BLK_COMPS);
    return CUP_ERR_ARG;
  }
  Comm *cm = (Comm *)c->comm;
  PostDesc post;
  WaitDesc wait;
  blk_descs(c, v, &post, &wait);
  const unsigned long long *seq = cm->d_seq + (size_t)v.xid * NKIND + K_BLK;
  BlkComps<Real> cp;
  for (int q = 0; q < BLK_COMPS; q++) {
    cp.src[q] = q < ncomp ? src[q] : nullptr;
    cp.dst[q] = q < ncomp ? dst[q] : nullptr;
  }
  const int ns = (int)v.blk_sslot.size(), nr = (int)v.blk_rslot.size();
  if (ns) {
    k_pack_blocks_leaf<Real><<<cgrid(c, ns), 128, 0, c->stream>>>(v.d_blk_sslot, ns, ncomp, 512, cp,
                                                                 (Real *const *)v.d_bptr0, (Real *const *)v.d_bptr1, seq,
                                                                 post);
    c->launches++;
  } else {
    k_signal<<<1, 64, 0, c->stream>>>(post.seq, cm->d_peer_win, post.peers, post.np, post.flag_index);
    c->launches++;
  }
  if (nr) {
    k_unpack_blocks_leaf<Real><<<cgrid(c, nr), 128, 0, c->stream>>>(v.d_blk_rslot, nr, ncomp, 512, 512 * v.blk_ncomp,
                                                                   (const Real *)v.d_brecv, v.blk_stride, seq, cp,
                                                                   dst_off, wait);
    c->launches++;
  } else if (wait.np > 0) {
    k_wait<<<1, 64, 0, c->stream>>>(wait);
    c->launches++;
  }
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

template int block_exchange_mg<double>(CupCtx *, Level &, SlotVec<double>, SlotVec<double>);
template int block_exchange_mg<float>(CupCtx *, Level &, SlotVec<float>, SlotVec<float>);
template int block_exchange_leaf<double>(CupCtx *, const double *const *, double *const *, int, long long);
template int block_exchange_leaf<float>(CupCtx *, const float *const *, float *const *, int, long long);
template int slab_exchange<double>(CupCtx *, Level &, const SlabSrc<double> &, int, int);
template int slab_exchange<float>(CupCtx *, Level &, const SlabSrc<float> &, int, int);
template int halo_post<double>(CupCtx *, Level &, SlotVec<double>);
template int halo_post<float>(CupCtx *, Level &, SlotVec<float>);
template int restrict_exchange<double>(CupCtx *, Level &, SlotVec<double>, SlotVec<double>, bool);
template int restrict_exchange<float>(CupCtx *, Level &, SlotVec<float>, SlotVec<float>, bool);
template int prolong_exchange<double>(CupCtx *, Level &, SlotVec<double>, SlotVec<double>);
template int prolong_exchange<float>(CupCtx *, Level &, SlotVec<float>, SlotVec<float>);

}  // namespace cup
