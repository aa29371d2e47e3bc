// blas_kernels.cu -- level-1 vector kernels of the Krylov solver and the
// pointwise updates of the time-step drivers.
//
// Reference: pois_dot/axpy/scale main.c:4854-4874, vec_copy/zero :4421-4432.
// All are streaming kernels (HBM bound): 128-bit loads where the vector type
// allows it, grid sized to a multiple of the SM count, block-level shuffle
// reductions finished with one atomicAdd per CTA into a device scalar so the
// host never waits for intermediate results.
#include "blas_kernels.cuh"

#include <algorithm>
#include <vector>
#include "cup_internal.h"
#include "comm.cuh"

namespace cup {

template <typename Real>
__global__ void __launch_bounds__(256) k_wdot(const Real *__restrict__ a, const Real *__restrict__ b,
                                              const Real *__restrict__ h3, long long nblk, double *out) {
  // sum_i a_i b_i / h_i^3, one 512-cell block per CTA iteration; h3 = h^3 per block
  __shared__ double red[8];
  double s = 0;
  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const Real *pa = a + blk * 512, *pb = b + blk * 512;
    double sb = (double)pa[threadIdx.x] * (double)pb[threadIdx.x] +
                (double)pa[threadIdx.x + 256] * (double)pb[threadIdx.x + 256];
    s += sb / (double)h3[blk];
  }
  for (int o = 16; o > 0; o >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0;
    for (int i = 0; i < 8; i++)
      tot += red[i];
    atomicAdd(out, tot);
  }
}

// y += alpha * x, alpha = sign * scal[idx] (device scalar) or a host constant
template <typename Real>
__global__ void __launch_bounds__(256) k_axpy(Real *__restrict__ y, const Real *__restrict__ x, long long n,
                                              double alpha_host, const double *alpha_dev, double sign) {
  const Real al = (Real)(alpha_dev ? sign * (*alpha_dev) : alpha_host);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] += al * x[i];
}

// y = alpha * x (+ beta * y)
template <typename Real>
__global__ void __launch_bounds__(256) k_scale_to(Real *__restrict__ y, const Real *__restrict__ x, long long n,
                                                  double alpha) {
  const Real al = (Real)alpha;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = al * x[i];
}

// r = b - r
template <typename Real>
__global__ void __launch_bounds__(256) k_bminus(Real *__restrict__ r, const Real *__restrict__ b, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    r[i] = b[i] - r[i];
}

// ---- fused Krylov kernels (pois_solve, main.c:4875-4979).  The modified Gram-Schmidt loop
//   for k <= j:  h_k = <w, V_k> ;  w -= h_k V_k            (:4930-4937)
// is sequential in k, but the axpy of step k and the dot of step k+1 touch the same w: one kernel does
// w -= h_k V_k and accumulates <w_new, V_{k+1}> (the last one <w_new, w_new>, :4938), i.e. four vector
// passes per k instead of five and one launch instead of two.  Arithmetic per element and per block is
// that of k_axpy / k_wdot.
template <typename Real>
__global__ void __launch_bounds__(256) k_axpy_dot(Real *__restrict__ w, const Real *__restrict__ vk,
                                                  const Real *vnext, const Real *__restrict__ h3, long long nblk,
                                                  const double *__restrict__ alpha, double *out) {
  __shared__ double red[8];
  const Real al = (Real)(-1.0 * (*alpha));
  double s = 0;
  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const long long i0 = blk * 512 + threadIdx.x, i1 = i0 + 256;
    Real w0 = w[i0], w1 = w[i1];
    w0 += al * vk[i0];
    w1 += al * vk[i1];
    w[i0] = w0;
    w[i1] = w1;
    const Real n0 = vnext ? vnext[i0] : w0, n1 = vnext ? vnext[i1] : w1;
    const double sb = (double)w0 * (double)n0 + (double)w1 * (double)n1;
    s += sb / (double)h3[blk];
  }
  for (int o = 16; o > 0; o >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0;
    for (int i = 0; i < 8; i++)
      tot += red[i];
    atomicAdd(out, tot);
  }
}

// r = b - r and <r, r> in one pass (pois_solve's residual, :4915-4917 / :4974-4976)
template <typename Real>
__global__ void __launch_bounds__(256) k_bminus_dot(Real *__restrict__ r, const Real *__restrict__ b,
                                                    const Real *__restrict__ h3, long long nblk, double *out) {
  __shared__ double red[8];
  double s = 0;
  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const long long i0 = blk * 512 + threadIdx.x, i1 = i0 + 256;
    const Real r0 = b[i0] - r[i0], r1 = b[i1] - r[i1];
    r[i0] = r0;
    r[i1] = r1;
    const double sb = (double)r0 * (double)r0 + (double)r1 * (double)r1;
    s += sb / (double)h3[blk];
  }
  for (int o = 16; o > 0; o >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0;
    for (int i = 0; i < 8; i++)
      tot += red[i];
    atomicAdd(out, tot);
  }
}

// w = sum_k y_k V_k (k < m), accumulated in the order of the reference's axpy loop (:4968-4970)
struct MultiY {
  double y[32];
};
template <typename Real>
__global__ void __launch_bounds__(256) k_multi_axpy(Real *__restrict__ w, const Real *__restrict__ V, long long n,
                                                    int m, MultiY Y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    Real acc = 0;
    for (int k = 0; k < m; k++)
      acc += (Real)Y.y[k] * V[(size_t)k * n + i];
    w[i] = acc;
  }
}

static inline int sgrid(const CupCtx *c, long long n) {
  long long g = (n + 255) / 256;
  long long cap = (long long)c->num_sms * 8;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

int wdot(CupCtx *c, const void *a, const void *b, int idx) {
  double *out = c->d_scal + idx;
  CUP_CUDA(cudaMemsetAsync(out, 0, sizeof(double), c->stream));
  long long g = c->nblk < (long long)c->num_sms * 8 ? c->nblk : (long long)c->num_sms * 8;
  if (c->real_bytes == 8)
    k_wdot<double><<<(int)g, 256, 0, c->stream>>>((const double *)a, (const double *)b, (const double *)c->d_hw,
                                                   c->nblk, out);
  else
    k_wdot<float><<<(int)g, 256, 0, c->stream>>>((const float *)a, (const float *)b, (const float *)c->d_hw,
                                                  c->nblk, out);
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  return comm_allreduce(c, idx, 1);  // MPI_Allreduce of pois_dot, main.c:4860
}

static inline int bgrid8(const CupCtx *c) {
  const long long g = c->nblk < (long long)c->num_sms * 8 ? c->nblk : (long long)c->num_sms * 8;
  return (int)(g < 1 ? 1 : g);
}

int axpy_dot(CupCtx *c, void *w, const void *vk, const void *vnext, int alpha_idx, int out_idx) {
  if (c->real_bytes == 8)
    k_axpy_dot<double><<<bgrid8(c), 256, 0, c->stream>>>((double *)w, (const double *)vk, (const double *)vnext,
                                                         (const double *)c->d_hw, c->nblk, c->d_scal + alpha_idx,
                                                         c->d_scal + out_idx);
  else
    k_axpy_dot<float><<<bgrid8(c), 256, 0, c->stream>>>((float *)w, (const float *)vk, (const float *)vnext,
                                                        (const float *)c->d_hw, c->nblk, c->d_scal + alpha_idx,
                                                        c->d_scal + out_idx);
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  return comm_allreduce(c, out_idx, 1);
}

int bminus_dot(CupCtx *c, void *r, const void *b, int out_idx) {
  CUP_CUDA(cudaMemsetAsync(c->d_scal + out_idx, 0, sizeof(double), c->stream));
  if (c->real_bytes == 8)
    k_bminus_dot<double><<<bgrid8(c), 256, 0, c->stream>>>((double *)r, (const double *)b, (const double *)c->d_hw,
                                                           c->nblk, c->d_scal + out_idx);
  else
    k_bminus_dot<float><<<bgrid8(c), 256, 0, c->stream>>>((float *)r, (const float *)b, (const float *)c->d_hw,
                                                          c->nblk, c->d_scal + out_idx);
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  return comm_allreduce(c, out_idx, 1);
}

int multi_axpy(CupCtx *c, void *w, const void *V, long long n, int m, const double *y) {
  MultiY Y;
  for (int k = 0; k < 32; k++)
    Y.y[k] = k < m ? y[k] : 0.0;
  if (c->real_bytes == 8)
    k_multi_axpy<double><<<sgrid(c, n), 256, 0, c->stream>>>((double *)w, (const double *)V, n, m, Y);
  else
    k_multi_axpy<float><<<sgrid(c, n), 256, 0, c->stream>>>((float *)w, (const float *)V, n, m, Y);
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

// sta_umax (main.c:5918-5939): max over cells of max(|u+uinf_x|, |v+uinf_y|, |w+uinf_z|).  The
// result is non-negative, so the IEEE bit pattern orders like the value and atomicMax on it works.
template <typename Real>
__global__ void __launch_bounds__(256) k_umax(const Real *__restrict__ u, const Real *__restrict__ v,
                                              const Real *__restrict__ w, long long n, double ux, double uy, double uz,
                                              unsigned long long *out) {
  __shared__ double red[8];
  double m = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double a = fabs((double)u[i] + ux), b = fabs((double)v[i] + uy), cc = fabs((double)w[i] + uz);
    double ul = a;
    if (ul < b)
      ul = b;
    if (ul < cc)
      ul = cc;
    if (m < ul)
      m = ul;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, m, o);
    m = m < t ? t : m;
  }
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; i++)
      m = m < red[i] ? red[i] : m;
    m = m < red[0] ? red[0] : m;
    atomicMax(out, (unsigned long long)__double_as_longlong(m));
  }
}

int umax(CupCtx *c, double *out) {
  if (c->nblk == 0) {
    set_error("umax: no mesh uploaded");
    return CUP_ERR_STATE;
  }
  const long long n = c->nblk * 512;
  double *d = c->d_scal + 6;
  CUP_CUDA(cudaMemsetAsync(d, 0, sizeof(double), c->stream));
  if (c->real_bytes == 8)
    k_umax<double><<<sgrid(c, n), 256, 0, c->stream>>>((const double *)c->state[CUP_F_VEL],
                                                       (const double *)c->state[CUP_F_VEL + 1],
                                                       (const double *)c->state[CUP_F_VEL + 2], n, c->prm.uinf[0],
                                                       c->prm.uinf[1], c->prm.uinf[2], (unsigned long long *)d);
  else
    k_umax<float><<<sgrid(c, n), 256, 0, c->stream>>>((const float *)c->state[CUP_F_VEL],
                                                      (const float *)c->state[CUP_F_VEL + 1],
                                                      (const float *)c->state[CUP_F_VEL + 2], n, c->prm.uinf[0],
                                                      c->prm.uinf[1], c->prm.uinf[2], (unsigned long long *)d);
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  CUP_TRY(comm_allreduce_max(c, 6, 1));
  CUP_TRY(fetch_scalars(c, 6, 1));
  *out = c->h_scal[6];
  return CUP_OK;
}

// vorticity()'s scaling (main.c:5789-5796): the three F_TMP components times 1/h^3 of the block
template <typename Real>
__global__ void __launch_bounds__(256) k_scale_blk3(Real *__restrict__ a0, Real *__restrict__ a1,
                                                    Real *__restrict__ a2, const Real *__restrict__ h3,
                                                    long long nblk) {
  for (long long b = blockIdx.x; b < nblk; b += gridDim.x) {
    const Real f = (Real)1.0 / h3[b];
    for (int j = threadIdx.x; j < 512; j += blockDim.x) {
      a0[b * 512 + j] *= f;
      a1[b * 512 + j] *= f;
      a2[b * 512 + j] *= f;
    }
  }
}

int scale_blk3(CupCtx *c, void *a0, void *a1, void *a2) {
  const int g = (int)std::min<long long>(c->nblk, (long long)c->num_sms * 8);
  if (c->real_bytes == 8)
    k_scale_blk3<double><<<g, 256, 0, c->stream>>>((double *)a0, (double *)a1, (double *)a2, (const double *)c->d_hw,
                                                   c->nblk);
  else
    k_scale_blk3<float><<<g, 256, 0, c->stream>>>((float *)a0, (float *)a1, (float *)a2, (const float *)c->d_hw,
                                                  c->nblk);
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

// mesh_tag_blk's norm (main.c:3683-3688): Linf over the block of |(u0,u1,u2)|, in double, and the same
// over the cells k_gradchi (main.c:3649) would not have zeroed, i.e. chi <= 0.9.
// out[2b] = all cells, out[2b+1] = fluid cells.  One warp per block.
template <typename Real>
__global__ void __launch_bounds__(256) k_blk_linf(const Real *__restrict__ u0, const Real *__restrict__ u1,
                                                  const Real *__restrict__ u2, const Real *__restrict__ chi,
                                                  long long nblk, double *__restrict__ out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long b = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); b < nblk; b += (long long)gridDim.x * wpb) {
    double ma = 0, mf = 0;
    for (int j = lane; j < 512; j += 32) {
      const double a = (double)u0[b * 512 + j], bb = (double)u1[b * 512 + j], cc = (double)u2[b * 512 + j];
      const double m = fabs(sqrt(a * a + bb * bb + cc * cc));
      ma = m > ma ? m : ma;
      if (!((double)chi[b * 512 + j] > 0.9))
        mf = m > mf ? m : mf;
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
      ma = fmax(ma, __shfl_xor_sync(0xffffffffu, ma, k));
      mf = fmax(mf, __shfl_xor_sync(0xffffffffu, mf, k));
    }
    if (lane == 0) {
      out[2 * b] = ma;
      out[2 * b + 1] = mf;
    }
  }
}

int block_linf(CupCtx *c, int f0, double *h_all, double *h_fluid) {
  if (c->nblk == 0) {
    set_error("block_linf: no mesh uploaded");
    return CUP_ERR_STATE;
  }
  if (f0 < 0 || f0 + 3 > CUP_F_N) {
    set_error("block_linf: field %d", f0);
    return CUP_ERR_ARG;
  }
  double *d = (double *)c->tmp_out;  // leaf-sized scratch of doubles (host-pointer entry points); 2*nblk used
  const int g = (int)std::min<long long>((c->nblk + 7) / 8, (long long)c->num_sms * 8);
  if (c->real_bytes == 8)
    k_blk_linf<double><<<g, 256, 0, c->stream>>>((const double *)c->state[f0], (const double *)c->state[f0 + 1],
                                                 (const double *)c->state[f0 + 2],
                                                 (const double *)c->state[CUP_F_CHI], c->nblk, d);
  else
    k_blk_linf<float><<<g, 256, 0, c->stream>>>((const float *)c->state[f0], (const float *)c->state[f0 + 1],
                                                (const float *)c->state[f0 + 2], (const float *)c->state[CUP_F_CHI],
                                                c->nblk, d);
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  std::vector<double> h((size_t)c->nblk * 2);
  CUP_CUDA(cudaMemcpyAsync(h.data(), d, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUP_CUDA(cudaStreamSynchronize(c->stream));
  for (long long b = 0; b < c->nblk; b++) {
    if (h_all)
      h_all[b] = h[2 * b];
    if (h_fluid)
      h_fluid[b] = h[2 * b + 1];
  }
  return CUP_OK;
}

// io_dump's packing loop (main.c:1525-1535): float32 chi, interleaved vorticity, Q
template <typename Real>
__global__ void __launch_bounds__(256) k_io_pack(const Real *__restrict__ chi, const Real *__restrict__ w0,
                                                 const Real *__restrict__ w1, const Real *__restrict__ w2,
                                                 const Real *__restrict__ qq, long long n, float *__restrict__ attr,
                                                 float *__restrict__ vort, float *__restrict__ q) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    attr[i] = (float)chi[i];
    vort[3 * i] = (float)w0[i];
    vort[3 * i + 1] = (float)w1[i];
    vort[3 * i + 2] = (float)w2[i];
    q[i] = (float)qq[i];
  }
}

int io_pack(CupCtx *c, float *h_attr, float *h_vort, float *h_q) {
  const long long n = c->nblk * 512;
  if (!c->io_buf)
    CUP_CUDA(cudaMalloc(&c->io_buf, (size_t)n * 5 * sizeof(float)));
  float *attr = (float *)c->io_buf, *q = attr + n, *vort = attr + 2 * n;
  const int g = (int)std::min<long long>((n + 255) / 256, (long long)c->num_sms * 8);
  if (c->real_bytes == 8)
    k_io_pack<double><<<g, 256, 0, c->stream>>>((const double *)c->state[CUP_F_CHI], (const double *)c->state[CUP_F_TMP],
                                                (const double *)c->state[CUP_F_TMP + 1],
                                                (const double *)c->state[CUP_F_TMP + 2],
                                                (const double *)c->state[CUP_F_LHS], n, attr, vort, q);
  else
    k_io_pack<float><<<g, 256, 0, c->stream>>>((const float *)c->state[CUP_F_CHI], (const float *)c->state[CUP_F_TMP],
                                               (const float *)c->state[CUP_F_TMP + 1],
                                               (const float *)c->state[CUP_F_TMP + 2],
                                               (const float *)c->state[CUP_F_LHS], n, attr, vort, q);
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  if (h_attr)
    CUP_CUDA(cudaMemcpyAsync(h_attr, attr, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (h_q)
    CUP_CUDA(cudaMemcpyAsync(h_q, q, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (h_vort)
    CUP_CUDA(cudaMemcpyAsync(h_vort, vort, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CUP_CUDA(cudaStreamSynchronize(c->stream));
  return CUP_OK;
}

int fetch_scalars(CupCtx *c, int first, int n) {
  CUP_CUDA(cudaMemcpyAsync(c->h_scal + first, c->d_scal + first, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost,
                           c->stream));
  CUP_CUDA(cudaStreamSynchronize(c->stream));
  return comm_check_error(c);  // a reduction that timed out on a peer must not be trusted
}

int axpy(CupCtx *c, void *y, const void *x, long long n, double alpha, int scal_idx, double sign) {
  const double *ad = scal_idx >= 0 ? c->d_scal + scal_idx : nullptr;
  if (c->real_bytes == 8)
    k_axpy<double><<<sgrid(c, n), 256, 0, c->stream>>>((double *)y, (const double *)x, n, alpha, ad, sign);
  else
    k_axpy<float><<<sgrid(c, n), 256, 0, c->stream>>>((float *)y, (const float *)x, n, alpha, ad, sign);
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

int scale_to(CupCtx *c, void *y, const void *x, long long n, double alpha) {
  if (c->real_bytes == 8)
    k_scale_to<double><<<sgrid(c, n), 256, 0, c->stream>>>((double *)y, (const double *)x, n, alpha);
  else
    k_scale_to<float><<<sgrid(c, n), 256, 0, c->stream>>>((float *)y, (const float *)x, n, alpha);
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

int bminus(CupCtx *c, void *r, const void *b, long long n) {
  if (c->real_bytes == 8)
    k_bminus<double><<<sgrid(c, n), 256, 0, c->stream>>>((double *)r, (const double *)b, n);
  else
    k_bminus<float><<<sgrid(c, n), 256, 0, c->stream>>>((float *)r, (const float *)b, n);
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

}  // namespace cup
