// obstacle.cu -- the pointwise obstacle (fish) phases of the time step on the device
// (SURVEY.md 8(f) row 1): fish_mom_blk (main.c:5057), fish_pen_blk (main.c:5601) and
// fish_tmpv (main.c:5799).  They sit between advdiff() and projection() in advance()
// (main.c:5993-5997); running them here keeps F_VEL / F_TMP resident in HBM for the whole step.
//
// What stays on the host (reference code, out of scope): the midline kinematics and the
// signed-distance rasterisation that PRODUCE ObstacleBlock.chi / .udef (fish_build), the 6x6
// rigid-body solve (fish_solve) and the collision model (fish_hit).  The boundary is therefore
//   host: fish_build            -> cup_obstacle_upload (chi, udef of the fish's blocks)
//   dev : advdiff
//   dev : cup_obstacle_moments  -> host: fish_solve -> cup_obstacle_motion (com, vel, omega)
//   dev : cup_obstacle_penalize
//   dev : projection (adds udef to F_TMP itself when bodies are present, main.c:5846)
//
// Layout: per body, chi [nob][512] and udef [nob][3][512] (SoA: the reference's [z][y][x][3]
// is transposed on the device after the upload) + geometry [nob][4] doubles (h, origin).
// One 512-thread CTA per obstacle block; all three kernels are a few hundred KB of traffic.
#include <vector>

#include "comm.cuh"
#include "cup_internal.h"

namespace cup {

struct Body {
  int nob = 0, cap = 0;
  int *d_blk = nullptr;      // [nob] local block index
  double *d_geo = nullptr;   // [nob][4] h, origin
  void *d_chi = nullptr;     // Real [nob][512]
  void *d_udef = nullptr;    // Real [nob][3][512]
  double com[3] = {0, 0, 0}, vel[3] = {0, 0, 0}, omega[3] = {0, 0, 0};
};

struct Obstacles {
  std::vector<Body> body;
  double *d_stage = nullptr;  // upload staging: chi then udef, doubles
  size_t stage_cap = 0;
  double *d_part = nullptr;   // [nob][CUP_M_N] per-block moments
  size_t part_cap = 0;
};

static Obstacles *obst(CupCtx *c) {
  if (!c->obst)
    c->obst = new Obstacles;
  return (Obstacles *)c->obst;
}

void free_obstacles(CupCtx *c) {
  Obstacles *o = (Obstacles *)c->obst;
  if (!o)
    return;
  for (Body &b : o->body) {
    cudaFree(b.d_blk);
    cudaFree(b.d_geo);
    cudaFree(b.d_chi);
    cudaFree(b.d_udef);
  }
  cudaFree(o->d_stage);
  cudaFree(o->d_part);
  delete o;
  c->obst = nullptr;
}

int obstacle_count(const CupCtx *c) {
  const Obstacles *o = (const Obstacles *)c->obst;
  int n = 0;
  if (o)
    for (const Body &b : o->body)
      n += b.nob > 0;
  return n;
}

// staging (doubles, reference layout) -> Real SoA
template <typename Real>
__global__ void __launch_bounds__(512) k_ob_unpack(const double *__restrict__ schi, const double *__restrict__ sudef,
                                                   Real *__restrict__ chi, Real *__restrict__ udef) {
  const int o = blockIdx.x, t = threadIdx.x;
  __shared__ double s[1536];
  chi[(size_t)o * 512 + t] = (Real)schi[(size_t)o * 512 + t];
  for (int k = t; k < 1536; k += 512)
    s[k] = sudef[(size_t)o * 1536 + k];
  __syncthreads();
#pragma unroll
  for (int d = 0; d < 3; d++)
    udef[((size_t)o * 3 + d) * 512 + t] = (Real)s[3 * t + d];
}

// fish_tmpv (main.c:5799-5827): TMP += udef where the field chi does not exceed the body's chi
template <typename Real>
__global__ void __launch_bounds__(512) k_tmpv(const int *__restrict__ blk, const Real *__restrict__ ochi,
                                              const Real *__restrict__ udef, const Real *__restrict__ chi,
                                              Real *__restrict__ t0, Real *__restrict__ t1, Real *__restrict__ t2) {
  const int o = blockIdx.x, t = threadIdx.x;
  const size_t g = (size_t)blk[o] * 512 + t;
  if (chi[g] > ochi[(size_t)o * 512 + t])
    return;
  const Real *u = udef + (size_t)o * 1536 + t;
  t0[g] += u[0];
  t1[g] += u[512];
  t2[g] += u[1024];
}

struct Motion {
  double com[3], vel[3], omega[3];
};

// fish_pen_blk (main.c:5601-5650): implicit penalisation towards the body velocity
template <typename Real>
__global__ void __launch_bounds__(512) k_pen(const int *__restrict__ blk, const double *__restrict__ geo,
                                             const Real *__restrict__ ochi, const Real *__restrict__ udef,
                                             const Real *__restrict__ chi, Real *__restrict__ v0, Real *__restrict__ v1,
                                             Real *__restrict__ v2, Motion m, Real dt, Real lambda) {
  const int o = blockIdx.x, t = threadIdx.x;
  const size_t g = (size_t)blk[o] * 512 + t;
  const Real oc = ochi[(size_t)o * 512 + t];
  if (chi[g] > oc || oc <= 0)
    return;
  const Real h = (Real)geo[4 * o];
  Real p[3];
  p[0] = ((Real)geo[4 * o + 1] + h * ((t & 7) + (Real)0.5)) - (Real)m.com[0];
  p[1] = ((Real)geo[4 * o + 2] + h * (((t >> 3) & 7) + (Real)0.5)) - (Real)m.com[1];
  p[2] = ((Real)geo[4 * o + 3] + h * ((t >> 6) + (Real)0.5)) - (Real)m.com[2];
  const Real X = oc > (Real)0.5 ? (Real)1 : (Real)0;
  const Real pen = X * lambda / (1 + X * lambda * dt);
  Real *v[3] = {v0, v1, v2};
  const Real *U = udef + (size_t)o * 1536 + t;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int e = (d + 1) % 3, f = (d + 2) % 3;
    const Real utot = (Real)m.vel[d] + (Real)m.omega[e] * p[f] - (Real)m.omega[f] * p[e] + U[512 * d];
    const Real u = v[d][g];
    v[d][g] = u + dt * (pen * (utot - u));
  }
}

// fish_mom_blk (main.c:5057-5118): the 29 moments of one obstacle block.  Accumulated in double
// for both precisions; cells are reduced by a fixed shuffle tree, blocks summed in block order
// by k_mom_sum (the reference sums cells, then blocks, sequentially: same value up to rounding).
template <typename Real>
__global__ void __launch_bounds__(512) k_mom(const int *__restrict__ blk, const double *__restrict__ geo,
                                             const Real *__restrict__ ochi, const Real *__restrict__ udef,
                                             const Real *__restrict__ v0, const Real *__restrict__ v1,
                                             const Real *__restrict__ v2, Motion m, double lambdt,
                                             double *__restrict__ part) {
  const int o = blockIdx.x, t = threadIdx.x;
  const size_t g = (size_t)blk[o] * 512 + t;
  __shared__ double red[16][CUP_M_N];
  double M[CUP_M_N];
#pragma unroll
  for (int q = 0; q < CUP_M_N; q++)
    M[q] = 0;
  const double X = (double)ochi[(size_t)o * 512 + t];
  if (X > 0) {
    const double h = geo[4 * o], dv = h * h * h;
    double p[3], u[3], du[3], pxu[3], pxdu[3];
    p[0] = (geo[4 * o + 1] + h * ((t & 7) + 0.5)) - m.com[0];
    p[1] = (geo[4 * o + 2] + h * (((t >> 3) & 7) + 0.5)) - m.com[1];
    p[2] = (geo[4 * o + 3] + h * ((t >> 6) + 0.5)) - m.com[2];
    u[0] = (double)v0[g];
    u[1] = (double)v1[g];
    u[2] = (double)v2[g];
#pragma unroll
    for (int d = 0; d < 3; d++)
      du[d] = u[d] - (double)udef[(size_t)o * 1536 + 512 * d + t];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      const int e = (d + 1) % 3, f = (d + 2) % 3;
      pxu[d] = p[e] * u[f] - p[f] * u[e];
      pxdu[d] = p[e] * du[f] - p[f] * du[e];
    }
    const double X1 = X > 0.5 ? 1.0 : 0.0;
    const double pf = dv * lambdt * X1 / (1 + X1 * lambdt), xv = X * dv;
    M[CUP_M_V] = xv;
    M[CUP_M_GFX] = pf;
    const double j0 = p[1] * p[1] + p[2] * p[2], j1 = p[0] * p[0] + p[2] * p[2], j2 = p[0] * p[0] + p[1] * p[1];
    M[CUP_M_J0 + 0] = xv * j0;
    M[CUP_M_J0 + 1] = xv * j1;
    M[CUP_M_J0 + 2] = xv * j2;
    M[CUP_M_J0 + 3] = -(xv * p[0] * p[1]);
    M[CUP_M_J0 + 4] = -(xv * p[0] * p[2]);
    M[CUP_M_J0 + 5] = -(xv * p[1] * p[2]);
    M[CUP_M_GJ0 + 0] = pf * j0;
    M[CUP_M_GJ0 + 1] = pf * j1;
    M[CUP_M_GJ0 + 2] = pf * j2;
    M[CUP_M_GJ0 + 3] = -(pf * p[0] * p[1]);
    M[CUP_M_GJ0 + 4] = -(pf * p[0] * p[2]);
    M[CUP_M_GJ0 + 5] = -(pf * p[1] * p[2]);
#pragma unroll
    for (int d = 0; d < 3; d++) {
      M[CUP_M_FX + d] = xv * u[d];
      M[CUP_M_TX + d] = xv * pxu[d];
      M[CUP_M_GPX + d] = pf * p[d];
      M[CUP_M_GUX + d] = pf * du[d];
      M[CUP_M_GAX + d] = pf * pxdu[d];
    }
  }
#pragma unroll
  for (int q = 0; q < CUP_M_N; q++) {
    double s = M[q];
#pragma unroll
    for (int k = 16; k > 0; k >>= 1)
      s += __shfl_xor_sync(0xffffffffu, s, k);
    if ((t & 31) == 0)
      red[t >> 5][q] = s;
  }
  __syncthreads();
  if (t < CUP_M_N) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < 16; w++)
      s += red[w][t];
    part[(size_t)o * CUP_M_N + t] = s;
  }
}

__global__ void k_mom_sum(const double *__restrict__ part, int nob, double *__restrict__ out) {
  const int q = threadIdx.x;
  if (q >= CUP_M_N)
    return;
  double s = 0;
  for (int o = 0; o < nob; o++)
    s += part[(size_t)o * CUP_M_N + q];
  out[q] = s;
}

static int grow(void **p, size_t bytes) {
  cudaFree(*p);
  *p = nullptr;
  CUP_CUDA(cudaMalloc(p, bytes));
  return CUP_OK;
}

static int check_body(CupCtx *c, int body, const char *who) {
  if (c->nblk == 0) {
    set_error("%s: no mesh uploaded", who);
    return CUP_ERR_STATE;
  }
  if (body < 0 || body >= CUP_MAX_BODIES) {
    set_error("%s: body %d out of range [0,%d)", who, body, CUP_MAX_BODIES);
    return CUP_ERR_ARG;
  }
  return CUP_OK;
}

int obstacle_upload(CupCtx *c, int body, int nob, const int *blk, const double *chi, const double *udef) {
  CUP_TRY(check_body(c, body, "cup_obstacle_upload"));
  if (nob < 0 || (nob > 0 && (!blk || !chi || !udef))) {
    set_error("cup_obstacle_upload: bad arguments");
    return CUP_ERR_ARG;
  }
  Obstacles *ob = obst(c);
  if ((int)ob->body.size() <= body)
    ob->body.resize(body + 1);
  Body &b = ob->body[body];
  b.nob = nob;
  if (nob == 0)
    return CUP_OK;
  std::vector<double> geo((size_t)nob * 4);
  for (int o = 0; o < nob; o++) {
    if (blk[o] < 0 || blk[o] >= c->nblk) {
      b.nob = 0;
      set_error("cup_obstacle_upload: block %d of %lld", blk[o], c->nblk);
      return CUP_ERR_ARG;
    }
    const CupBlk &k = c->blk[blk[o]];
    geo[4 * o] = k.h;
    geo[4 * o + 1] = k.origin[0];
    geo[4 * o + 2] = k.origin[1];
    geo[4 * o + 3] = k.origin[2];
  }
  if (nob > b.cap) {
    b.cap = 0;  // a failed allocation below must not leave a stale capacity behind
    b.nob = 0;
    CUP_TRY(grow((void **)&b.d_blk, (size_t)nob * sizeof(int)));
    CUP_TRY(grow((void **)&b.d_geo, (size_t)nob * 4 * sizeof(double)));
    CUP_TRY(grow(&b.d_chi, (size_t)nob * 512 * c->real_bytes));
    CUP_TRY(grow(&b.d_udef, (size_t)nob * 1536 * c->real_bytes));
    b.cap = nob;
    b.nob = nob;
  }
  const size_t need = (size_t)nob * 2048;
  if (need > ob->stage_cap) {
    ob->stage_cap = 0;
    CUP_TRY(grow((void **)&ob->d_stage, need * sizeof(double)));
    ob->stage_cap = need;
  }
  if ((size_t)nob * CUP_M_N > ob->part_cap) {
    ob->part_cap = 0;
    CUP_TRY(grow((void **)&ob->d_part, (size_t)nob * CUP_M_N * sizeof(double)));
    ob->part_cap = (size_t)nob * CUP_M_N;
  }
  double *schi = ob->d_stage, *sudef = ob->d_stage + (size_t)nob * 512;
  CUP_CUDA(cudaMemcpyAsync(b.d_blk, blk, (size_t)nob * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CUP_CUDA(cudaMemcpyAsync(b.d_geo, geo.data(), geo.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUP_CUDA(cudaMemcpyAsync(schi, chi, (size_t)nob * 512 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUP_CUDA(cudaMemcpyAsync(sudef, udef, (size_t)nob * 1536 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (c->real_bytes == 8)
    k_ob_unpack<double><<<nob, 512, 0, c->stream>>>(schi, sudef, (double *)b.d_chi, (double *)b.d_udef);
  else
    k_ob_unpack<float><<<nob, 512, 0, c->stream>>>(schi, sudef, (float *)b.d_chi, (float *)b.d_udef);
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  // geo lives on this stack frame and the staging buffer is shared by all bodies
  CUP_CUDA(cudaStreamSynchronize(c->stream));
  return CUP_OK;
}

int obstacle_motion(CupCtx *c, int body, const double com[3], const double vel[3], const double omega[3]) {
  CUP_TRY(check_body(c, body, "cup_obstacle_motion"));
  Obstacles *ob = obst(c);
  if ((int)ob->body.size() <= body)
    ob->body.resize(body + 1);
  Body &b = ob->body[body];
  for (int d = 0; d < 3; d++) {
    if (com)
      b.com[d] = com[d];
    if (vel)
      b.vel[d] = vel[d];
    if (omega)
      b.omega[d] = omega[d];
  }
  return CUP_OK;
}

int obstacle_clear(CupCtx *c) {
  Obstacles *ob = (Obstacles *)c->obst;
  if (ob)
    for (Body &b : ob->body)
      b.nob = 0;
  return CUP_OK;
}

static Motion motion_of(const Body &b) {
  Motion m;
  for (int d = 0; d < 3; d++) {
    m.com[d] = b.com[d];
    m.vel[d] = b.vel[d];
    m.omega[d] = b.omega[d];
  }
  return m;
}

template <typename Real>
static int tmpv_t(CupCtx *c) {
  Obstacles *ob = (Obstacles *)c->obst;
  if (!ob)
    return CUP_OK;
  for (Body &b : ob->body) {  // bodies in order, as the reference's k loop (main.c:5801)
    if (b.nob == 0)
      continue;
    k_tmpv<Real><<<b.nob, 512, 0, c->stream>>>(b.d_blk, (const Real *)b.d_chi, (const Real *)b.d_udef,
                                               (const Real *)c->state[CUP_F_CHI], (Real *)c->state[CUP_F_TMP],
                                               (Real *)c->state[CUP_F_TMP + 1], (Real *)c->state[CUP_F_TMP + 2]);
    c->launches++;
  }
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

int obstacle_tmpv(CupCtx *c) { return c->real_bytes == 8 ? tmpv_t<double>(c) : tmpv_t<float>(c); }

template <typename Real>
static int pen_t(CupCtx *c) {
  Obstacles *ob = (Obstacles *)c->obst;
  if (!ob)
    return CUP_OK;
  for (Body &b : ob->body) {  // per block the bodies are applied in order k (main.c:5659)
    if (b.nob == 0)
      continue;
    k_pen<Real><<<b.nob, 512, 0, c->stream>>>(b.d_blk, b.d_geo, (const Real *)b.d_chi, (const Real *)b.d_udef,
                                              (const Real *)c->state[CUP_F_CHI], (Real *)c->state[CUP_F_VEL],
                                              (Real *)c->state[CUP_F_VEL + 1], (Real *)c->state[CUP_F_VEL + 2],
                                              motion_of(b), (Real)c->prm.dt, (Real)c->prm.lambda);
    c->launches++;
  }
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

int obstacle_penalize(CupCtx *c) { return c->real_bytes == 8 ? pen_t<double>(c) : pen_t<float>(c); }

enum { SCAL_MOM = 64 };  // d_scal[64..64+CUP_M_N)

int obstacle_moments(CupCtx *c, int body, double *M) {
  CUP_TRY(check_body(c, body, "cup_obstacle_moments"));
  Obstacles *ob = obst(c);
  double *out = c->d_scal + SCAL_MOM;
  const bool have = body < (int)ob->body.size() && ob->body[body].nob > 0;
  if (!have) {
    CUP_CUDA(cudaMemsetAsync(out, 0, CUP_M_N * sizeof(double), c->stream));
  } else {
    Body &b = ob->body[body];
    const double lambdt = c->prm.lambda * c->prm.dt;
    if (c->real_bytes == 8)
      k_mom<double><<<b.nob, 512, 0, c->stream>>>(b.d_blk, b.d_geo, (const double *)b.d_chi, (const double *)b.d_udef,
                                                  (const double *)c->state[CUP_F_VEL],
                                                  (const double *)c->state[CUP_F_VEL + 1],
                                                  (const double *)c->state[CUP_F_VEL + 2], motion_of(b), lambdt,
                                                  ob->d_part);
    else
      k_mom<float><<<b.nob, 512, 0, c->stream>>>(b.d_blk, b.d_geo, (const float *)b.d_chi, (const float *)b.d_udef,
                                                 (const float *)c->state[CUP_F_VEL],
                                                 (const float *)c->state[CUP_F_VEL + 1],
                                                 (const float *)c->state[CUP_F_VEL + 2], motion_of(b), lambdt,
                                                 ob->d_part);
    k_mom_sum<<<1, 32, 0, c->stream>>>(ob->d_part, b.nob, out);
    c->launches += 2;
    CUP_CUDA(cudaGetLastError());
  }
  CUP_TRY(comm_allreduce(c, SCAL_MOM, CUP_M_N));  // MPI_Allreduce of M, main.c:5326
  CUP_TRY(fetch_scalars(c, SCAL_MOM, CUP_M_N));
  for (int q = 0; q < CUP_M_N; q++)
    M[q] = c->h_scal[SCAL_MOM + q];
  return CUP_OK;
}

}  // namespace cup
