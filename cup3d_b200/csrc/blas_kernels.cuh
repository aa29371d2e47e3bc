// blas_kernels.cuh -- host launchers of the streaming vector kernels.
#pragma once
struct CupCtx;
namespace cup {
// y += alpha*x ; alpha = sign * d_scal[scal_idx] when scal_idx >= 0, else the host value
int axpy(CupCtx *c, void *y, const void *x, long long n, double alpha, int scal_idx, double sign);
int scale_to(CupCtx *c, void *y, const void *x, long long n, double alpha);  // y = alpha*x
int bminus(CupCtx *c, void *r, const void *b, long long n);                  // r = b - r
}  // namespace cup
