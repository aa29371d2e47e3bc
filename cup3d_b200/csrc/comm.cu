// comm.cu -- one rank per GPU over NCCL / NVLink.
//
// Replaces the reference's MPI layer on the hot path:
//   tree_sync      (Allgatherv of all blocks, main.c:2928)   -> comm_gather_blocks
//   halo_sync      (whole 8^3 blocks by Alltoallv, :3101)    -> halo_exchange: device-side
//                                                               packing of 8x8 FACES + grouped
//                                                               ncclSend/ncclRecv per peer
//   mg_down/mg_up  (Alltoallv of 128 / 64 Reals, :4754,:4791)-> restrict_exchange / prolong_exchange
//   MPI_Allreduce  (:4295, :4820, :4860, ...)                -> ncclAllReduce on device scalars
// Everything is stream-ordered on the context's stream; the host never waits.
#include <dlfcn.h>
#include <nccl.h>

#include "comm.cuh"
#include "cup_internal.h"
#include "mg_device.cuh"

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

struct Comm {
  ncclComm_t nccl = nullptr;
  void *gather_dev = nullptr;
  size_t gather_bytes = 0;
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

void comm_free(CupCtx *c) {
  Comm *cm = (Comm *)c->comm;
  if (!cm)
    return;
  cudaFree(cm->gather_dev);
  if (cm->nccl)
    g_nccl.CommDestroy(cm->nccl);
  delete cm;
  c->comm = nullptr;
}

// tree_sync: every rank contributes its blocks; result = global list in rank order + owner per block
int comm_gather_blocks(CupCtx *c, const CupBlk *blk, long long n, std::vector<CupBlk> &gblk,
                       std::vector<int> &owner) {
  if (c->nranks == 1) {
    gblk.assign(blk, blk + n);
    owner.assign((size_t)n, 0);
    return CUP_OK;
  }
  Comm *cm = (Comm *)c->comm;
  const int R = c->nranks;
  long long *d_cnt;
  CUP_CUDA(cudaMalloc((void **)&d_cnt, (size_t)(R + 1) * sizeof(long long)));
  CUP_CUDA(cudaMemcpyAsync(d_cnt + R, &n, sizeof n, cudaMemcpyHostToDevice, c->stream));
  CUP_NCCL(g_nccl.AllGather(d_cnt + R, d_cnt, sizeof(long long), ncclChar, cm->nccl, c->stream));
  std::vector<long long> cnt((size_t)R);
  CUP_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)R * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  CUP_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_cnt);
  long long mx = 0, tot = 0;
  for (int r = 0; r < R; r++) {
    mx = cnt[r] > mx ? cnt[r] : mx;
    tot += cnt[r];
  }
  const size_t chunk = (size_t)mx * sizeof(CupBlk);
  char *d;
  CUP_CUDA(cudaMalloc((void **)&d, chunk * (size_t)(R + 1)));
  CUP_CUDA(cudaMemcpyAsync(d + chunk * R, blk, (size_t)n * sizeof(CupBlk), cudaMemcpyHostToDevice, c->stream));
  CUP_NCCL(g_nccl.AllGather(d + chunk * R, d, chunk, ncclChar, cm->nccl, c->stream));
  std::vector<char> h(chunk * (size_t)R);
  CUP_CUDA(cudaMemcpyAsync(h.data(), d, h.size(), cudaMemcpyDeviceToHost, c->stream));
  CUP_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d);
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

int comm_allreduce(CupCtx *c, int first, int n) {
  if (c->nranks == 1)
    return CUP_OK;
  Comm *cm = (Comm *)c->comm;
  CUP_NCCL(g_nccl.AllReduce(c->d_scal + first, c->d_scal + first, (size_t)n, ncclDouble, ncclSum, cm->nccl, c->stream));
  return CUP_OK;
}

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
// pack kernels
// ---------------------------------------------------------------------------
// one 64-thread CTA per face: plane p of block `slot`, element (a, c) = (t&7, t>>3) in the
// receiver's convention (mg_device.cuh load_halo): x planes (y,z), y planes (x,z), z planes (x,y)
template <typename Real>
__global__ void __launch_bounds__(64) k_pack_faces(const int *__restrict__ sslot, const int *__restrict__ splane,
                                                   int n, SlotVec<Real> u, Real *__restrict__ out) {
  const int t = threadIdx.x, a = t & 7, cc = t >> 3;
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
    out[(size_t)e * 64 + t] = b[idx];
  }
}

// mg_put (main.c:4722): received 64 r + 64 u of a remote child -> the parent's octant
template <typename Real>
__global__ void __launch_bounds__(64) k_put(const int *__restrict__ rslot, const int *__restrict__ roct, int n,
                                            const Real *__restrict__ in, SlotVec<Real> f, SlotVec<Real> u) {
  const int t = threadIdx.x, cx = t & 3, cy = (t >> 2) & 3, cz = t >> 4;
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const int o = roct[e], ps = rslot[e];
    const int pidx = ((4 * (o >> 2) + cz) << 6) + ((4 * ((o >> 1) & 1) + cy) << 3) + 4 * (o & 1) + cx;
    f.at(ps)[pidx] = in[(size_t)e * 128 + t];
    u.at(ps)[pidx] = in[(size_t)e * 128 + 64 + t];
  }
}

// mg_get (main.c:4771): u_c - us of the parent's octant, for a remote child
template <typename Real>
__global__ void __launch_bounds__(64) k_get(const int *__restrict__ rslot, const int *__restrict__ roct, int n,
                                            SlotVec<Real> u, SlotVec<Real> us, Real *__restrict__ out) {
  const int t = threadIdx.x, cx = t & 3, cy = (t >> 2) & 3, cz = t >> 4;
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const int o = roct[e], ps = rslot[e];
    const int pidx = ((4 * (o >> 2) + cz) << 6) + ((4 * ((o >> 1) & 1) + cy) << 3) + 4 * (o & 1) + cx;
    out[(size_t)e * 64 + t] = u.at(ps)[pidx] - us.at(ps)[pidx];
  }
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

// scratch of one level: faces out/in, restriction out/in (prolongation reuses them reversed)
int comm_alloc_level_buffers(CupCtx *c) {
  const size_t rb = (size_t)c->real_bytes;
  for (auto &v : c->lv) {
    cudaFree(v.d_fsend);
    cudaFree(v.d_frecv);
    cudaFree(v.d_rsend);
    cudaFree(v.d_rrecv);
    v.d_fsend = v.d_frecv = v.d_rsend = v.d_rrecv = nullptr;
    if (c->nranks == 1)
      continue;
    const size_t ns = v.face_sslot.size(), nr = (size_t)v.nface_recv;
    const size_t cs = (size_t)sum(v.res_scnt), cr = (size_t)sum(v.res_rcnt);
    if (ns)
      CUP_CUDA(cudaMalloc(&v.d_fsend, ns * 64 * rb));
    if (nr)
      CUP_CUDA(cudaMalloc(&v.d_frecv, nr * 64 * rb));
    if (cs)
      CUP_CUDA(cudaMalloc(&v.d_rsend, cs * 128 * rb));
    if (cr)
      CUP_CUDA(cudaMalloc(&v.d_rrecv, cr * 128 * rb));
  }
  return CUP_OK;
}

void comm_free_level_buffers(CupCtx *c) {
  for (auto &v : c->lv) {
    cudaFree(v.d_fsend);
    cudaFree(v.d_frecv);
    cudaFree(v.d_rsend);
    cudaFree(v.d_rrecv);
    v.d_fsend = v.d_frecv = v.d_rsend = v.d_rrecv = nullptr;
  }
}

template <typename Real>
int halo_exchange(CupCtx *c, Level &v, SlotVec<Real> u) {
  if (c->nranks == 1)
    return CUP_OK;
  const int ns = (int)v.face_sslot.size();
  if (ns == 0 && v.nface_recv == 0)
    return CUP_OK;
  if (ns) {
    k_pack_faces<Real><<<cgrid(c, ns), 64, 0, c->stream>>>(v.d_face_sslot, v.d_face_splane, ns, u, (Real *)v.d_fsend);
    c->launches++;
  }
  CUP_TRY(exchange(c, v.d_fsend, v.face_scnt, v.d_frecv, v.face_rcnt, 64 * sizeof(Real)));
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

// after k_down wrote the children with remote parents into v.d_rsend
template <typename Real>
int restrict_exchange(CupCtx *c, Level &v, SlotVec<Real> f, SlotVec<Real> u) {
  if (c->nranks == 1)
    return CUP_OK;
  const int cr = sum(v.res_rcnt);
  if (cr == 0 && sum(v.res_scnt) == 0)
    return CUP_OK;
  CUP_TRY(exchange(c, v.d_rsend, v.res_scnt, v.d_rrecv, v.res_rcnt, 128 * sizeof(Real)));
  if (cr) {
    k_put<Real><<<cgrid(c, cr), 64, 0, c->stream>>>(v.d_res_rslot, v.d_res_roct, cr, (const Real *)v.d_rrecv, f, u);
    c->launches++;
  }
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

// before k_up: parents' owners send u_c - us to remote children; lands in v.d_rsend (as receive buffer)
template <typename Real>
int prolong_exchange(CupCtx *c, Level &v, SlotVec<Real> u, SlotVec<Real> us) {
  if (c->nranks == 1)
    return CUP_OK;
  const int cr = sum(v.res_rcnt);
  if (cr == 0 && sum(v.res_scnt) == 0)
    return CUP_OK;
  if (cr) {
    k_get<Real><<<cgrid(c, cr), 64, 0, c->stream>>>(v.d_res_rslot, v.d_res_roct, cr, u, us, (Real *)v.d_rrecv);
    c->launches++;
  }
  CUP_TRY(exchange(c, v.d_rrecv, v.res_rcnt, v.d_rsend, v.res_scnt, 64 * sizeof(Real)));
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

template int halo_exchange<double>(CupCtx *, Level &, SlotVec<double>);
template int halo_exchange<float>(CupCtx *, Level &, SlotVec<float>);
template int restrict_exchange<double>(CupCtx *, Level &, SlotVec<double>, SlotVec<double>);
template int restrict_exchange<float>(CupCtx *, Level &, SlotVec<float>, SlotVec<float>);
template int prolong_exchange<double>(CupCtx *, Level &, SlotVec<double>, SlotVec<double>);
template int prolong_exchange<float>(CupCtx *, Level &, SlotVec<float>, SlotVec<float>);

}  // namespace cup
