// blas_kernels.cuh -- host launchers of the streaming vector kernels.
#pragma once
struct CupCtx;
namespace cup {
// y += alpha*x ; alpha = sign * d_scal[scal_idx] when scal_idx >= 0, else the host value
int axpy(CupCtx *c, void *y, const void *x, long long n, double alpha, int scal_idx, double sign);
int scale_to(CupCtx *c, void *y, const void *x, long long n, double alpha);  // y = alpha*x
int bminus(CupCtx *c, void *r, const void *b, long long n);                  // r = b - r
// fused Krylov kernels: w -= d_scal[alpha_idx] * vk and d_scal[out_idx] += <w, vnext> (vnext null: <w, w>),
// all-reduced; the caller zeroes d_scal[out_idx] first
int axpy_dot(CupCtx *c, void *w, const void *vk, const void *vnext, int alpha_idx, int out_idx);
int bminus_dot(CupCtx *c, void *r, const void *b, int out_idx);  // r = b - r ; d_scal[out_idx] = <r, r> (all ranks)
int multi_axpy(CupCtx *c, void *w, const void *V, long long n, int m, const double *y);  // w = sum_k y_k V[k], V = [m][n]
}  // namespace cup
