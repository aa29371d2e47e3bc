// capi.cu -- the extern "C" boundary declared in include/cup3d_b200.h.
#include <algorithm>
#include <thread>

#include "blas_kernels.cuh"
#include "cup_internal.h"
#include "smooth_tma.cuh"
#include "comm.cuh"

namespace cup {

static thread_local std::string g_err;

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
}

template <typename Real>
__global__ void k_cvt_in(Real *__restrict__ dst, const double *__restrict__ src, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (Real)src[i];
}
template <typename Real>
__global__ void k_cvt_out(double *__restrict__ dst, const Real *__restrict__ src, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (double)src[i];
}

static int alloc_state(CupCtx *c) {
  // nstate = own leaves + the leaf context's ghost blocks (multi-level meshes across ranks)
  const size_t bytes = (size_t)c->nstate * 512 * (size_t)c->real_bytes;
  cudaFree(c->leaf_ghost);
  c->leaf_ghost = nullptr;
  for (int q = 0; q < 3; q++) {
    cudaFree(c->vel_spare[q]);
    c->vel_spare[q] = nullptr;
  }
  if (c->leafv.nghost > 0) {
    const size_t rows = (size_t)std::max<long long>(c->leafv.nghost, c->nslot - c->nblk + 1);
    CUP_CUDA(cudaMalloc(&c->leaf_ghost, rows * 512 * (size_t)c->real_bytes));
    CUP_CUDA(cudaMemset(c->leaf_ghost, 0, rows * 512 * (size_t)c->real_bytes));
  }
  for (int f = 0; f < CUP_F_N; f++) {
    cudaFree(c->state[f]);
    c->state[f] = nullptr;
    CUP_CUDA(cudaMalloc(&c->state[f], bytes));
    CUP_CUDA(cudaMemset(c->state[f], 0, bytes));
  }
  cudaFree(c->tmp_in);
  cudaFree(c->tmp_out);
  cudaFree(c->p_old);
  cudaFree(c->tmp_stage);
  cudaFree(c->io_buf);
  c->io_buf = nullptr;
  c->tmp_in = c->tmp_out = c->p_old = c->tmp_stage = nullptr;
  // staging for host flat vectors: always room for doubles
  CUP_CUDA(cudaMalloc(&c->tmp_in, (size_t)c->nblk * 512 * 8));
  CUP_CUDA(cudaMalloc(&c->tmp_out, (size_t)c->nblk * 512 * 8));
  CUP_CUDA(cudaMalloc(&c->p_old, bytes));
  if (c->real_bytes == 4)
    CUP_CUDA(cudaMalloc(&c->tmp_stage, (size_t)c->nblk * 512 * 8));
  return CUP_OK;
}

// host flat vector of doubles -> device vector of Real (dst may alias stage for fp64)
static int vec_h2d(CupCtx *c, void *d_dst, const double *h_src, long long n) {
  if (c->real_bytes == 8) {
    CUP_CUDA(cudaMemcpyAsync(d_dst, h_src, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  } else {
    CUP_CUDA(cudaMemcpyAsync(c->tmp_stage, h_src, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    k_cvt_in<float><<<c->num_sms * 8, 256, 0, c->stream>>>((float *)d_dst, (const double *)c->tmp_stage, n);
    c->launches++;
  }
  return CUP_OK;
}

// Every entry point that touches a context runs on the context's device whatever the caller's
// current device is, and restores the caller's device on return (one process may hold contexts on
// several devices).
struct DevGuard {
  int prev = -1;
  bool switched = false;
  explicit DevGuard(const CupCtx *c) {
    if (!c)
      return;
    if (cudaGetDevice(&prev) == cudaSuccess && prev != c->device)
      switched = cudaSetDevice(c->device) == cudaSuccess;
  }
  ~DevGuard() {
    if (switched)
      cudaSetDevice(prev);
  }
};

#define CUP_ENTER(c)                         \
  if (!(c)) {                                \
    cup::set_error("%s: null context", __func__); \
    return CUP_ERR_ARG;                      \
  }                                          \
  cup::DevGuard guard_(c)

// ---- sta.fld <-> device.  The host layout is [blk][9][512] doubles, the device keeps one flat vector per
// field.  A strided (2-D) DMA copy pays ~10 us per 4 KB row -- 7 s for the three velocity fields of a
// 241k-block mesh -- so the transfer goes through CONTIGUOUS chunks instead: a pinned bounce buffer
// [blocks of the chunk][nc][512] (filled / emptied by a few host threads when nc < 9; the user's memory
// itself when all nine fields move), one 1-D copy per chunk, and a device kernel that (un)interleaves
// and converts to / from Real.
namespace {
enum { XFER_CHUNK_BYTES = 64 << 20 };

struct Xfer {
  double *h_bounce[2] = {nullptr, nullptr};
  double *d_stage[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
};
Xfer *xfer_of(CupCtx *c) {
  if (!c->xfer) {
    Xfer *x = new Xfer;
    for (int i = 0; i < 2; i++) {
      if (cudaMallocHost((void **)&x->h_bounce[i], XFER_CHUNK_BYTES) != cudaSuccess ||
          cudaMalloc((void **)&x->d_stage[i], XFER_CHUNK_BYTES) != cudaSuccess ||
          cudaEventCreateWithFlags(&x->done[i], cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        for (int j = 0; j < 2; j++) {
          if (x->h_bounce[j])
            cudaFreeHost(x->h_bounce[j]);
          cudaFree(x->d_stage[j]);
          if (x->done[j])
            cudaEventDestroy(x->done[j]);
        }
        delete x;
        return nullptr;
      }
    }
    c->xfer = x;
  }
  return (Xfer *)c->xfer;
}

struct FieldPtrs {
  void *p[CUP_F_N];
};

// stage [nb][nc][512] doubles <-> fields f0..f0+nc-1, blocks b0..b0+nb-1
template <typename Real, bool IN>
__global__ void __launch_bounds__(256) k_xfer(double *__restrict__ stage, FieldPtrs F, int f0, int nc, long long b0,
                                              int nb) {
  const long long n = (long long)nb * nc * 512;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int cell = (int)(i & 511);
    const long long r = i >> 9;
    const int q = (int)(r % nc);
    const long long blk = b0 + r / nc;
    Real *fp = (Real *)F.p[f0 + q] + blk * 512 + cell;
    if (IN)
      *fp = (Real)stage[i];
    else
      stage[i] = (double)*fp;
  }
}

// copy rows of `rowd` doubles between the user's strided layout and a dense buffer with a few threads
void host_rows(double *dense, double *strided, long long nrows, size_t rowd, size_t pitchd, bool gather) {
  const int nt = nrows > 4096 ? 8 : 1;
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++)
    th.emplace_back([=]() {
      for (long long r = nrows * t / nt; r < nrows * (t + 1) / nt; r++) {
        if (gather)
          memcpy(dense + (size_t)r * rowd, strided + (size_t)r * pitchd, rowd * sizeof(double));
        else
          memcpy(strided + (size_t)r * pitchd, dense + (size_t)r * rowd, rowd * sizeof(double));
      }
    });
  for (auto &x : th)
    x.join();
}

int state_xfer(CupCtx *c, double *h_fld, int f0, int nc, bool in) {
  Xfer *x = xfer_of(c);
  if (!x) {
    set_error("cup_state_%s: cannot allocate the transfer buffers", in ? "h2d" : "d2h");
    return CUP_ERR_CUDA;
  }
  FieldPtrs F;
  for (int f = 0; f < CUP_F_N; f++)
    F.p[f] = c->state[f];
  const size_t rowd = (size_t)nc * 512, pitchd = (size_t)CUP_F_N * 512;
  const long long per = (long long)(XFER_CHUNK_BYTES / (rowd * sizeof(double)));
  const bool dense = nc == CUP_F_N;  // the user's memory is contiguous for this range
  int k = 0;
  // d2h with a bounce buffer: the scatter of chunk i runs while chunk i+1 is in flight
  long long pend_b0 = -1, pend_nb = 0;
  int pend_k = 0;
  auto scatter_pending = [&]() -> int {
    if (pend_b0 < 0)
      return CUP_OK;
    CUP_CUDA(cudaEventSynchronize(x->done[pend_k]));
    host_rows(x->h_bounce[pend_k], h_fld + (size_t)pend_b0 * pitchd + (size_t)f0 * 512, pend_nb, rowd, pitchd, false);
    pend_b0 = -1;
    return CUP_OK;
  };
  for (long long b0 = 0; b0 < c->nblk; b0 += per, k ^= 1) {
    const int nb = (int)std::min<long long>(per, c->nblk - b0);
    const size_t bytes = (size_t)nb * rowd * sizeof(double);
    const int grid = (int)std::min<long long>(((long long)nb * rowd + 255) / 256, (long long)c->num_sms * 8);
    double *hsrc = dense ? h_fld + (size_t)b0 * pitchd : x->h_bounce[k];
    if (in) {
      if (!dense) {
        CUP_CUDA(cudaEventSynchronize(x->done[k]));  // the copy that last used this bounce buffer has left it
        host_rows(x->h_bounce[k], h_fld + (size_t)b0 * pitchd + (size_t)f0 * 512, nb, rowd, pitchd, true);
      }
      CUP_CUDA(cudaMemcpyAsync(x->d_stage[k], hsrc, bytes, cudaMemcpyHostToDevice, c->stream));
      CUP_CUDA(cudaEventRecord(x->done[k], c->stream));
      if (c->real_bytes == 8)
        k_xfer<double, true><<<grid, 256, 0, c->stream>>>(x->d_stage[k], F, f0, nc, b0, nb);
      else
        k_xfer<float, true><<<grid, 256, 0, c->stream>>>(x->d_stage[k], F, f0, nc, b0, nb);
    } else {
      if (c->real_bytes == 8)
        k_xfer<double, false><<<grid, 256, 0, c->stream>>>(x->d_stage[k], F, f0, nc, b0, nb);
      else
        k_xfer<float, false><<<grid, 256, 0, c->stream>>>(x->d_stage[k], F, f0, nc, b0, nb);
      if (!dense && pend_b0 >= 0 && pend_k == k)
        CUP_TRY(scatter_pending());  // this bounce buffer still holds an unscattered chunk
      CUP_CUDA(cudaMemcpyAsync(hsrc, x->d_stage[k], bytes, cudaMemcpyDeviceToHost, c->stream));
      CUP_CUDA(cudaEventRecord(x->done[k], c->stream));
      if (!dense) {
        CUP_TRY(scatter_pending());
        pend_b0 = b0;
        pend_nb = nb;
        pend_k = k;
      }
    }
    c->launches++;
  }
  CUP_CUDA(cudaGetLastError());
  CUP_TRY(scatter_pending());
  CUP_CUDA(cudaStreamSynchronize(c->stream));
  return comm_check_error(c);
}
}  // namespace


}  // namespace cup

using namespace cup;

extern "C" {

const char *cup_last_error(void) { return g_err.c_str(); }
int cup_version(void) { return 100; }

int cup_create(CupCtx **out, int device, int real_bytes) {
  if (!out || (real_bytes != 8 && real_bytes != 4)) {
    set_error("cup_create: real_bytes must be 8 or 4");
    return CUP_ERR_ARG;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    // no CPU fallback: the product path is CUDA only
    set_error("cup_create: no CUDA device (%s)", cudaGetErrorString(e));
    return CUP_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    set_error("cup_create: device %d of %d", device, ndev);
    return CUP_ERR_ARG;
  }
  CUP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUP_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("cup_create: device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major,
              prop.minor);
    return CUP_ERR_CUDA;
  }
  CupCtx *c = new CupCtx;
  c->device = device;
  c->real_bytes = real_bytes;
  c->num_sms = prop.multiProcessorCount;
  c->prm.mean_constraint = 2;
  c->prm.ptol = 1e-6;
  c->prm.ptol_rel = 1e-4;
  c->prm.nu = 1e-3;
  // the NULL stream cannot be captured into a CUDA graph: work goes to an own BLOCKING stream
  // (implicitly ordered against legacy-default-stream work of the caller) unless one is set
  cudaError_t e1 = cudaMalloc((void **)&c->d_scal, SCAL_N * sizeof(double));
  if (e1 == cudaSuccess)
    e1 = cudaMemset(c->d_scal, 0, SCAL_N * sizeof(double));
  if (e1 == cudaSuccess)
    e1 = cudaMallocHost((void **)&c->h_scal, SCAL_N * sizeof(double));
  if (e1 == cudaSuccess)
    e1 = cudaStreamCreate(&c->own_stream);
  if (e1 != cudaSuccess) {
    set_error("cup_create: %s", cudaGetErrorString(e1));
    cudaFree(c->d_scal);
    if (c->h_scal)
      cudaFreeHost(c->h_scal);
    delete c;
    return CUP_ERR_CUDA;
  }
  c->stream = c->own_stream;
  *out = c;
  return CUP_OK;
}

int cup_destroy(CupCtx *c) {
  if (!c)
    return CUP_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  free_krylov(c);
  free_obstacles(c);
  free_graph_cache(c);
  comm_free_level_buffers(c);
  free_mesh(c);
  free_tma_cache(c);
  comm_free(c);
  for (int f = 0; f < CUP_F_N; f++)
    cudaFree(c->state[f]);
  cudaFree(c->u1_leaf);
  cudaFree(c->u0_x);
  cudaFree(c->u1_x);
  cudaFree(c->f_x);
  cudaFree(c->us_x);
  cudaFree(c->tmp_in);
  cudaFree(c->tmp_out);
  cudaFree(c->p_old);
  cudaFree(c->tmp_stage);
  cudaFree(c->io_buf);
  cudaFree(c->leaf_ghost);
  for (int q = 0; q < 3; q++)
    cudaFree(c->vel_spare[q]);
  cudaFree(c->d_W);
  cudaFree(c->d_hw);
  if (c->xfer) {
    Xfer *x = (Xfer *)c->xfer;
    for (int i = 0; i < 2; i++) {
      cudaFreeHost(x->h_bounce[i]);
      cudaFree(x->d_stage[i]);
      cudaEventDestroy(x->done[i]);
    }
    delete x;
    c->xfer = nullptr;
  }
  cudaFree(c->d_scal);
  cudaFreeHost(c->h_scal);
  if (c->h_err)
    cudaFreeHost(c->h_err);
  cudaStreamDestroy(c->own_stream);
  delete c;
  return CUP_OK;
}

int cup_set_stream(CupCtx *c, void *stream) {
  CUP_ENTER(c);
  c->stream = stream ? (cudaStream_t)stream : c->own_stream;
  return CUP_OK;
}

int cup_set_params(CupCtx *c, const CupParams *p) {
  CUP_ENTER(c);
  if (!p) {
    set_error("cup_set_params: null");
    return CUP_ERR_ARG;
  }
  c->prm = *p;
  return CUP_OK;
}

int cup_synchronize(CupCtx *c) {
  CUP_ENTER(c);
  CUP_CUDA(cudaStreamSynchronize(c->stream));
  return comm_check_error(c);
}

static int mesh_upload_impl(CupCtx *c, const CupBlk *blk, long long n, const int bpd[3], int level_max) {
  CUP_CUDA(cudaStreamSynchronize(c->stream));
  free_krylov(c);
  free_obstacles(c);
  free_graph_cache(c);
  free_tma_cache(c);  // descriptors are keyed by pointers that are about to be freed
  comm_free_level_buffers(c);
  // tree_sync (main.c:2928): all ranks learn all blocks; owner = contributing rank
  std::vector<CupBlk> gblk;
  std::vector<int> owner;
  CUP_TRY(comm_gather_blocks(c, blk, n, gblk, owner));
  CUP_TRY(build_mesh(c, gblk.data(), (long long)gblk.size(), owner.data(), bpd, level_max));
  CUP_TRY(alloc_state(c));
  CUP_TRY(mg_setup(c));
  CUP_TRY(comm_alloc_level_buffers(c));
  return CUP_OK;
}

int cup_mesh_upload(CupCtx *c, const CupBlk *blk, long long n, const int bpd[3], int level_max) {
  CUP_ENTER(c);
  if (!blk || !bpd || n <= 0) {
    set_error("cup_mesh_upload: blk/bpd null or n = %lld", n);
    return CUP_ERR_ARG;
  }
  const int rc = mesh_upload_impl(c, blk, n, bpd, level_max);
  if (rc != CUP_OK) {
    // never leave a half-built mesh behind: later calls must fail with CUP_ERR_STATE, not launch
    // kernels on null tables
    const std::string keep = cup_last_error();
    comm_free_level_buffers(c);
    free_mesh(c);
    set_error("%s", keep.c_str());
  }
  return rc;
}

int cup_mesh_adapt(CupCtx *c, const CupBlk *nb, long long n, const int *kind, const long long *src, const int bpd[3],
                   int level_max) {
  CUP_ENTER(c);
  if (!nb || !kind || !src || !bpd || n <= 0) {
    set_error("cup_mesh_adapt: null argument or n = %lld", n);
    return CUP_ERR_ARG;
  }
  if (c->nblk == 0) {
    set_error("cup_mesh_adapt: no mesh uploaded");
    return CUP_ERR_STATE;
  }
  void *out[CUP_F_N] = {nullptr};
  int rc = adapt_fields(c, nb, n, kind, src, out);
  if (rc == CUP_OK) {
    rc = mesh_upload_impl(c, nb, n, bpd, level_max);
    if (rc != CUP_OK) {
      const std::string keep = cup_last_error();
      comm_free_level_buffers(c);
      free_mesh(c);
      set_error("%s", keep.c_str());
    }
  }
  if (rc == CUP_OK) {
    const size_t bytes = (size_t)n * 512 * (size_t)c->real_bytes;
    for (int f = 0; f < CUP_F_N && rc == CUP_OK; f++)
      if (cudaMemcpyAsync(c->state[f], out[f], bytes, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) {
        set_error("cup_mesh_adapt: copy of field %d failed", f);
        rc = CUP_ERR_CUDA;
      }
    cudaStreamSynchronize(c->stream);
  }
  for (int f = 0; f < CUP_F_N; f++)
    cudaFree(out[f]);
  return rc;
}

long long cup_nblk(const CupCtx *c) { return c->nblk; }
long long cup_nslot(const CupCtx *c) { return c->nslot; }
int cup_mg_levels(const CupCtx *c) { return c->top + 1; }
long long cup_mg_nact(const CupCtx *c, int level) {
  return (level < 0 || level > c->top) ? -1 : (long long)c->lv[level].act.size();
}

int cup_state_h2d(CupCtx *c, const double *h_fld, int f0, int nc) {
  CUP_ENTER(c);
  if (!h_fld || f0 < 0 || nc < 1 || f0 + nc > CUP_F_N || c->nblk == 0) {
    set_error("cup_state_h2d: bad field range %d+%d (or no mesh / null pointer)", f0, nc);
    return CUP_ERR_ARG;
  }
  return state_xfer(c, const_cast<double *>(h_fld), f0, nc, true);
}

int cup_state_d2h(CupCtx *c, double *h_fld, int f0, int nc) {
  CUP_ENTER(c);
  if (!h_fld || f0 < 0 || nc < 1 || f0 + nc > CUP_F_N || c->nblk == 0) {
    set_error("cup_state_d2h: bad field range %d+%d (or no mesh / null pointer)", f0, nc);
    return CUP_ERR_ARG;
  }
  return state_xfer(c, h_fld, f0, nc, false);
}

void *cup_state_dev(CupCtx *c, int f) { return (f < 0 || f >= CUP_F_N) ? nullptr : c->state[f]; }

int cup_pois_op_dev(CupCtx *c, const void *d_in, void *d_out) {
  CUP_ENTER(c);
  return pois_op_dev(c, d_in, d_out);
}
int cup_mg_vcycle_dev(CupCtx *c, const void *d_in, void *d_out) {
  CUP_ENTER(c);
  return mg_vcycle_dev(c, d_in, d_out);
}

static int host_op(CupCtx *c, const double *h_in, double *h_out, int which) {
  CUP_ENTER(c);
  if (c->nblk == 0) {
    set_error("no mesh uploaded");
    return CUP_ERR_STATE;
  }
  const long long N = c->nblk * 512;
  void *din = c->tmp_in, *dout = c->tmp_out;
  CUP_TRY(vec_h2d(c, din, h_in, N));
  CUP_TRY(which == 0 ? pois_op_dev(c, din, dout) : mg_vcycle_dev(c, din, dout));
  if (c->real_bytes == 4) {
    k_cvt_out<float><<<c->num_sms * 8, 256, 0, c->stream>>>((double *)c->tmp_stage, (const float *)dout, N);
    c->launches++;
    dout = c->tmp_stage;
  }
  CUP_CUDA(cudaMemcpyAsync(h_out, dout, (size_t)N * 8, cudaMemcpyDeviceToHost, c->stream));
  CUP_CUDA(cudaStreamSynchronize(c->stream));
  return comm_check_error(c);
}

int cup_pois_op(CupCtx *c, const double *h_in, double *h_out) { return host_op(c, h_in, h_out, 0); }
int cup_mg_vcycle(CupCtx *c, const double *h_in, double *h_out) { return host_op(c, h_in, h_out, 1); }

int cup_pois_dot_dev(CupCtx *c, const void *a, const void *b, double *result) {
  CUP_ENTER(c);
  CUP_TRY(wdot(c, a, b, 4));
  CUP_TRY(fetch_scalars(c, 4, 1));
  *result = c->h_scal[4];
  return CUP_OK;
}

int cup_pois_solve(CupCtx *c, CupSolveInfo *info) {
  CUP_ENTER(c);
  return pois_solve(c, info);
}
int cup_umax(CupCtx *c, double *out) {
  CUP_ENTER(c);
  return umax(c, out);
}
int cup_advdiff(CupCtx *c) {
  CUP_ENTER(c);
  return advdiff(c);
}
int cup_vorticity(CupCtx *c) {
  CUP_ENTER(c);
  return vorticity(c);
}
int cup_io_pack(CupCtx *c, float *attr, float *vort, float *q) {
  CUP_ENTER(c);
  if (c->nblk == 0) {
    set_error("cup_io_pack: no mesh uploaded");
    return CUP_ERR_STATE;
  }
  // io_dump (main.c:1441-1442): vorticity(); qcrit();
  CUP_TRY(vorticity(c));
  CUP_TRY(stencil_run(c, CUP_ST_Q, nullptr, c->nblk));
  return io_pack(c, attr, vort, q);
}
int cup_block_linf(CupCtx *c, int f0, double *linf_all, double *linf_fluid) {
  CUP_ENTER(c);
  return block_linf(c, f0, linf_all, linf_fluid);
}
int cup_projection(CupCtx *c, CupSolveInfo *info) {
  CUP_ENTER(c);
  return projection(c, info);
}
int cup_projection_udef_ready(CupCtx *c, int flag) {
  CUP_ENTER(c);
  c->keep_tmp_udef = flag != 0;
  return CUP_OK;
}
int cup_obstacle_upload(CupCtx *c, int body, int nob, const int *blk, const double *chi, const double *udef) {
  CUP_ENTER(c);
  return obstacle_upload(c, body, nob, blk, chi, udef);
}
int cup_obstacle_motion(CupCtx *c, int body, const double com[3], const double vel[3], const double omega[3]) {
  CUP_ENTER(c);
  return obstacle_motion(c, body, com, vel, omega);
}
int cup_obstacle_clear(CupCtx *c) {
  CUP_ENTER(c);
  return obstacle_clear(c);
}
int cup_obstacle_moments(CupCtx *c, int body, double *M) {
  CUP_ENTER(c);
  if (!M) {
    set_error("cup_obstacle_moments: M is NULL");
    return CUP_ERR_ARG;
  }
  return obstacle_moments(c, body, M);
}
int cup_obstacle_penalize(CupCtx *c) {
  CUP_ENTER(c);
  return obstacle_penalize(c);
}
int cup_obstacle_tmpv(CupCtx *c) {
  CUP_ENTER(c);
  return obstacle_tmpv(c);
}
int cup_stencil_apply(CupCtx *c, CupStencilId id) {
  CUP_ENTER(c);
  return stencil_run(c, id, nullptr, c->nblk);
}
int cup_stencil_run(CupCtx *c, CupStencilId id, const long long *list, long long n) {
  CUP_ENTER(c);
  return stencil_run(c, id, list, n);
}

int cup_comm_init(CupCtx *c, int rank, int nranks, const void *id, size_t id_bytes) {
  CUP_ENTER(c);
  return comm_init(c, rank, nranks, id, id_bytes);
}
int cup_nccl_unique_id(void *out, size_t bytes) { return comm_unique_id(out, bytes); }
int cup_comm_init_host(CupCtx *c, int rank, int nranks, CupAllgatherFn allgather, void *user) {
  CUP_ENTER(c);
  return comm_init_host(c, rank, nranks, allgather, user);
}

static int *dup_ints(const std::vector<int> &v) {
  int *p = (int *)malloc((v.size() + 1) * sizeof(int));
  if (!v.empty())
    memcpy(p, v.data(), v.size() * sizeof(int));
  return p;
}

int cup_plan_build(const CupBlk *gblk, long long G, const int *owner, int nranks, int rank, const int bpd[3],
                   int level_max, int level, CupPlan *out) {
  HostMesh m;
  CUP_TRY(build_tables(&m, gblk, G, owner, nranks, rank, bpd, level_max));
  if (level < -1 || level > m.top || !out) {
    set_error("cup_plan_build: level %d of %d", level, m.top + 1);
    return CUP_ERR_ARG;
  }
  const Level &v = level < 0 ? m.leafv : m.lv[level];
  memset(out, 0, sizeof *out);
  out->nblk = m.nblk;
  out->nslot = m.nslot;
  out->nact = (int)v.act.size();
  out->nsend = (int)v.face_sslot.size();
  out->nrecv = v.nface_recv;
  out->act = dup_ints(v.act);
  std::vector<int> ijk = v.ijk;
  if (level < 0) {  // leaf context: block indices of the local leaves, each at its own level
    ijk.clear();
    for (int sl : v.act) {
      ijk.push_back(m.blk[(size_t)sl].ix);
      ijk.push_back(m.blk[(size_t)sl].iy);
      ijk.push_back(m.blk[(size_t)sl].iz);
    }
  }
  ijk.resize(v.act.size() * 3, 0);
  out->ijk = dup_ints(ijk);
  out->nbr = dup_ints(v.nbr);
  out->send_slot = dup_ints(v.face_sslot);
  out->send_plane = dup_ints(v.face_splane);
  out->send_cnt = dup_ints(v.face_scnt);
  out->recv_cnt = dup_ints(v.face_rcnt);
  out->pslot = dup_ints(v.pslot);
  out->oct = dup_ints(v.oct);
  out->res_send_cnt = dup_ints(v.res_scnt);
  out->res_recv_cnt = dup_ints(v.res_rcnt);
  out->nres_recv = (int)v.res_rslot.size();
  out->res_recv_slot = dup_ints(v.res_rslot);
  out->res_recv_oct = dup_ints(v.res_roct);
  out->ghosted = v.ghosted ? 1 : 0;
  out->nghost = v.nghost;
  std::vector<int> ext = v.ext;
  ext.resize(v.act.size() * 24, -1);
  out->ext = dup_ints(ext);
  out->nbsend = (int)v.blk_sslot.size();
  out->nbrecv = (int)v.blk_rslot.size();
  out->bsend_slot = dup_ints(v.blk_sslot);
  out->bsend_kind = dup_ints(v.blk_skind);
  out->bsend_peer = dup_ints(v.blk_speer);
  out->bsend_idx = dup_ints(v.blk_sidx);
  out->brecv_slot = dup_ints(v.blk_rslot);
  out->brecv_kind = dup_ints(v.blk_rkind);
  return CUP_OK;
}

void cup_plan_free(CupPlan *p) {
  if (!p)
    return;
  free(p->act);
  free(p->ijk);
  free(p->nbr);
  free(p->send_slot);
  free(p->send_plane);
  free(p->send_cnt);
  free(p->recv_cnt);
  free(p->pslot);
  free(p->oct);
  free(p->res_send_cnt);
  free(p->res_recv_cnt);
  free(p->res_recv_slot);
  free(p->res_recv_oct);
  free(p->ext);
  free(p->bsend_slot);
  free(p->bsend_kind);
  free(p->bsend_peer);
  free(p->bsend_idx);
  free(p->brecv_slot);
  free(p->brecv_kind);
  memset(p, 0, sizeof *p);
}

long long cup_kernel_launches(const CupCtx *c) { return c->launches; }
int cup_trace_report(CupCtx *c, char *out, size_t cap) {
  CUP_ENTER(c);
  return trace_report(c, out, cap);
}
int cup_time_smooth(CupCtx *c, int level, int reps, float *ms) {
  CUP_ENTER(c);
  return time_smooth(c, level, reps, ms);
}
int cup_mg_smooth_dev(CupCtx *c, int level, int n, void *d_u, const void *d_f) {
  CUP_ENTER(c);
  return mg_smooth_slots(c, level, n, d_u, d_f);
}
void *cup_mg_array(CupCtx *c, int which) {
  switch (which) {
  case 0: return c->u0_x;
  case 1: return c->f_x;
  case 2: return c->us_x;
  case 3: return c->u1_x;
  }
  return nullptr;
}

}  // extern "C"
