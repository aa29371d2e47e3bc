// solver.cu -- host drivers above the kernels: restarted right-preconditioned
// GMRES(30) with the V-cycle as preconditioner (pois_solve, main.c:4875-4979).
//
// Control flow and arithmetic follow the reference statement by statement; the
// only changes are WHERE things run: vectors never leave HBM, every dot
// product reduces into a device scalar, and the modified Gram-Schmidt axpy
// reads its coefficient straight from that scalar, so the host synchronises
// once per Krylov iteration (to run the 30x30 Givens update and the
// convergence test) instead of once per dot product.
#include <cmath>

#include "blas_kernels.cuh"
#include "cup_internal.h"

namespace cup {

enum { KR_M = 30, KR_MAXIT = 1000 };  // main.c:4400

struct Krylov {
  long long cap = 0;
  // x and b are the state's F_PRES / F_LHS themselves (the reference solves in place, :4900-4905)
  void *r = nullptr, *w = nullptr, *z = nullptr, *V = nullptr;
};

void free_krylov(CupCtx *c) {
  if (!c->kr)
    return;
  cudaFree(c->kr->r);
  cudaFree(c->kr->w);
  cudaFree(c->kr->z);
  cudaFree(c->kr->V);
  delete c->kr;
  c->kr = nullptr;
}

static int kr_alloc(CupCtx *c, long long N) {
  if (c->kr && c->kr->cap >= N)
    return CUP_OK;
  free_krylov(c);
  c->kr = new Krylov;
  const size_t rb = (size_t)c->real_bytes;
  CUP_CUDA(cudaMalloc(&c->kr->r, N * rb));
  CUP_CUDA(cudaMalloc(&c->kr->w, N * rb));
  CUP_CUDA(cudaMalloc(&c->kr->z, N * rb));
  CUP_CUDA(cudaMalloc(&c->kr->V, (size_t)(KR_M + 1) * N * rb));
  c->kr->cap = N;
  return CUP_OK;
}

template <typename Real>
__global__ void k_setcell(Real *p, long long off, Real v) {
  p[off] = v;
}

int pois_solve(CupCtx *c, CupSolveInfo *info) {
  const long long N = c->nblk * 512;
  const size_t rb = (size_t)c->real_bytes;
  if (c->nblk == 0) {
    set_error("pois_solve: no mesh uploaded");
    return CUP_ERR_STATE;
  }
  CUP_TRY(kr_alloc(c, N));
  Krylov &K = *c->kr;
  const int mc = c->prm.mean_constraint;
  const double ptol = c->prm.ptol, ptol_rel = c->prm.ptol_rel;
  double H[KR_M + 1][KR_M], cs[KR_M], sn[KR_M], g[KR_M + 1], y[KR_M];
  const double vol = c->gvol;          // sum over ALL ranks (MPI_Allreduce, main.c:4891)
  const long long pin = c->pin_local;  // block (0,0,0) on its owner, else -1
  auto Vj = [&](int j) { return (void *)((char *)K.V + (size_t)j * N * rb); };
  // b = F_LHS (with the pinned cell zeroed for constraint 1 / >2), x = F_PRES: used in place
  void *const Kb = c->state[CUP_F_LHS], *const Kx = c->state[CUP_F_PRES];
  if ((mc == 1 || mc > 2) && pin >= 0) {
    if (c->real_bytes == 8)
      k_setcell<double><<<1, 1, 0, c->stream>>>((double *)Kb, pin * 512, 0.0);
    else
      k_setcell<float><<<1, 1, 0, c->stream>>>((float *)Kb, pin * 512, 0.f);
    c->launches++;
  }
  CUP_TRY(wdot(c, Kb, Kb, 2));
  CUP_TRY(pois_op_dev(c, Kx, K.r));
  CUP_TRY(bminus_dot(c, K.r, Kb, 3));  // r = b - A x and <r, r> in one pass
  CUP_TRY(fetch_scalars(c, 2, 2));
  const double bnorm = std::sqrt(c->h_scal[2] / vol);
  double beta = std::sqrt(c->h_scal[3]);
  double norm = beta / std::sqrt(vol);
  int it = 0, restarts = 0, vcycles = 0;
  for (;;) {
    if (norm < ptol || norm < ptol_rel * bnorm)
      break;
    if (it >= KR_MAXIT) {
      fprintf(stderr, "cup3d_b200: poisson did not converge in %d iterations: residual %.3e, rhs %.3e\n", it, norm,
              bnorm);
      break;
    }
    CUP_TRY(scale_to(c, Vj(0), K.r, N, 1 / beta));
    memset(g, 0, sizeof g);
    g[0] = beta;
    int j;
    for (j = 0; j < KR_M; j++) {
      CUP_TRY(mg_vcycle_dev(c, Vj(j), K.z));
      vcycles++;
      CUP_TRY(pois_op_dev(c, K.z, K.w));
      // modified Gram-Schmidt: H[k][j] stays on the device (scalar 8+k).  The axpy of step k and the dot
      // of step k+1 run in one kernel (the last one accumulates <w, w>): same sequential semantics
      CUP_CUDA(cudaMemsetAsync(c->d_scal + 8, 0, (size_t)(j + 2) * sizeof(double), c->stream));
      CUP_TRY(wdot(c, K.w, Vj(0), 8));
      for (int k = 0; k <= j; k++)
        CUP_TRY(axpy_dot(c, K.w, Vj(k), k < j ? Vj(k + 1) : nullptr, 8 + k, 8 + k + 1));
      CUP_TRY(fetch_scalars(c, 8, j + 2));
      for (int k = 0; k <= j; k++)
        H[k][j] = c->h_scal[8 + k];
      H[j + 1][j] = std::sqrt(c->h_scal[8 + j + 1]);
      const int brk = H[j + 1][j] <= 1e-13 * std::fabs(H[j][j]);
      if (H[j + 1][j] > 0)
        CUP_TRY(scale_to(c, Vj(j + 1), K.w, N, 1 / H[j + 1][j]));
      else
        CUP_CUDA(cudaMemcpyAsync(Vj(j + 1), K.w, N * rb, cudaMemcpyDeviceToDevice, c->stream));
      for (int k = 0; k < j; k++) {
        const double t = cs[k] * H[k][j] + sn[k] * H[k + 1][j];
        H[k + 1][j] = -sn[k] * H[k][j] + cs[k] * H[k + 1][j];
        H[k][j] = t;
      }
      {
        const double d = std::sqrt(H[j][j] * H[j][j] + H[j + 1][j] * H[j + 1][j]);
        cs[j] = d > 0 ? H[j][j] / d : 1;
        sn[j] = d > 0 ? H[j + 1][j] / d : 0;
        H[j][j] = d;
        H[j + 1][j] = 0;
        g[j + 1] = -sn[j] * g[j];
        g[j] = cs[j] * g[j];
      }
      it++;
      norm = std::fabs(g[j + 1]) / std::sqrt(vol);
      if (norm < ptol || norm < ptol_rel * bnorm || it >= KR_MAXIT || brk) {
        j++;
        break;
      }
    }
    for (int k = j - 1; k >= 0; k--) {
      y[k] = g[k];
      for (int m = k + 1; m < j; m++)
        y[k] -= H[k][m] * y[m];
      y[k] = H[k][k] != 0 ? y[k] / H[k][k] : 0;
    }
    // x += M(sum_k y_k V_k): one more V-cycle on the combination (main.c:4968-4972)
    CUP_TRY(multi_axpy(c, K.w, K.V, N, j, y));  // w = sum_k y_k V_k in one pass over the basis
    CUP_TRY(mg_vcycle_dev(c, K.w, K.z));
    vcycles++;
    CUP_TRY(axpy(c, Kx, K.z, N, 1.0, -1, 1.0));
    CUP_TRY(pois_op_dev(c, Kx, K.r));
    CUP_TRY(bminus_dot(c, K.r, Kb, 3));
    CUP_TRY(fetch_scalars(c, 3, 1));
    beta = std::sqrt(c->h_scal[3]);
    norm = beta / std::sqrt(vol);
    restarts++;
  }
  CUP_CUDA(cudaStreamSynchronize(c->stream));
  if (info) {
    info->iterations = it;
    info->restarts = restarts;
    info->residual = norm;
    info->rhs_norm = bnorm;
    info->vcycles = vcycles;
  }
  return CUP_OK;
}

}  // namespace cup
