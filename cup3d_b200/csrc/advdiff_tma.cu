// advdiff_tma.cu -- k_advdiff (main.c:4986) with every ghost layer staged by the TMA engine.
//
// Why: the plain-load kernel (stencil_kernels.cu:k_advdiff) is latency bound -- ncu (r01,
// profiles/r01_advdiff_ldg_ncu_summary.json): issue slots 27 % busy, 8.2 warps per issue stalled
// on the long scoreboard (the ghost gathers), FP64 pipe 18 %, LSU 20 %, 16 warps per SM.  The ghost
// layers are 3 deep in every direction, so the register-staged path cannot prefetch them without
// spilling.  Here the copy engine does the fetching, asynchronously, into dense shared arrays:
//
//   work item = (block, component c).  Per item the kernel needs
//     core_c  [8][8][8]          the block's own values of u_c        (1-D bulk copy, 4 KB)
//     zlo/zhi [3][8][8]          planes 5..7 / 0..2 of the z neighbours (1-D bulk copies)
//     ylo/yhi [8][3][8]          rows   5..7 / 0..2 of the y neighbours (tensor box {8,3,8})
//     xlo/xhi [8][8][4]          columns 4..7 / 0..3 of the x neighbours (tensor box {4,8,8})
//   The three cores of a block stay resident for the whole block (every component needs
//   U = u + uinf of all three at every cell; they are read into registers once per block) and are
//   refilled for the NEXT block as soon as their own item is done; the ghost arrays are double
//   buffered across items.  So while component c is computed, the ghosts of c+1 and c+2 and the
//   cores of the next block are in flight, and no register is tied up by a pending load.
//
// Faces without a same-level neighbour on this rank -- walls (ghost = +-boundary cell,
// main.c:3480 OP_BC) and faces received from another rank -- are rare; their ghost arrays are
// written by the threads themselves after the copies of the other faces have landed.
//
// The arithmetic (operation order, the exact x/60) is the same as the plain-load kernel's.
#include "advdiff_tma.cuh"
#include "cup_internal.h"
#include "tma.cuh"

namespace cup {

namespace {

enum { CORE = 512, ZH = 192, YH = 192, XH = 256, HALO = 2 * ZH + 2 * YH + 2 * XH };  // Reals
enum { O_ZLO = 0, O_ZHI = ZH, O_YLO = 2 * ZH, O_YHI = 2 * ZH + YH, O_XLO = 2 * ZH + 2 * YH, O_XHI = 2 * ZH + 2 * YH + XH };

template <typename Real>
__device__ __forceinline__ Real div60(Real x) {  // IEEE x/60 without the division sequence (see stencil_kernels.cu)
  const Real c = (Real)1 / (Real)60;
  const Real q0 = x * c;
  const Real r = fma(-q0, (Real)60, x);
  return fma(r, c, q0);
}

template <typename Real>
__device__ __forceinline__ Real upwind(Real U, Real um3, Real um2, Real um1, Real u, Real up1, Real up2, Real up3) {
  // derivative(), main.c:4980-4985; one polynomial on sign-selected operands (bitwise the same)
  const bool pos = U > 0;
  const Real a3 = pos ? um3 : up3, a2 = pos ? um2 : up2, a1 = pos ? um1 : up1;
  const Real b1 = pos ? up1 : um1, b2 = pos ? up2 : um2;
  const Real r = (((((Real)-2 * a3 + (Real)15 * a2) - (Real)60 * a1) + (Real)20 * u) + (Real)30 * b1 - (Real)3 * b2);
  const Real q = div60<Real>(r);
  return pos ? q : -q;
}

struct AdvMaps {
  CUtensorMap x[3], y[3];
};

template <typename Real>
struct AdvArgs {
  const Real *vel[3];
  Real *tmp[3];
  const int *sub;
  int nsub;
  const Real *hblk;
  Real dtnu_dt, dtnu_nu, fac_a0, fac_d0, uinf[3];
  // fused Runge-Kutta stage (advdiff(), main.c:5039-5054): with T = TMP + rhs the kernel stores
  // vout = u + T * rk_ih3 and TMP = T * rk_beta instead of TMP = T (vout != vel: neighbours still read u)
  Real *vout[3];
  Real rk_ih3, rk_beta;
};

// producer: ghost layers of component c of the block with neighbours nb[] -> ghost slot h
template <typename Real>
__device__ __forceinline__ void issue_halo(Real *h, uint64_t *bar, const Real *vc, const CUtensorMap *mx,
                                           const CUtensorMap *my, const int (&nb)[6]) {
  constexpr uint32_t RB = sizeof(Real);
  uint32_t bytes = 0;
  bytes += (nb[0] >= 0 ? XH : 0) + (nb[1] >= 0 ? XH : 0) + (nb[2] >= 0 ? YH : 0) + (nb[3] >= 0 ? YH : 0) +
           (nb[4] >= 0 ? ZH : 0) + (nb[5] >= 0 ? ZH : 0);
  mbar_arrive_expect_tx(bar, bytes * RB);
  if (nb[4] >= 0)
    tma_load_1d(h + O_ZLO, vc + (size_t)nb[4] * 512 + 5 * 64, ZH * RB, bar);
  if (nb[5] >= 0)
    tma_load_1d(h + O_ZHI, vc + (size_t)nb[5] * 512, ZH * RB, bar);
  if (nb[2] >= 0)
    tma_load_3d(h + O_YLO, my, 0, 5, nb[2] * 8, bar);
  if (nb[3] >= 0)
    tma_load_3d(h + O_YHI, my, 0, 0, nb[3] * 8, bar);
  if (nb[0] >= 0)
    tma_load_3d(h + O_XLO, mx, 4, 0, nb[0] * 8, bar);
  if (nb[1] >= 0)
    tma_load_3d(h + O_XHI, mx, 0, 0, nb[1] * 8, bar);
}

// NS = 2: ghost arrays double buffered (32 KB per CTA in fp64, 6 CTAs per SM); NS = 1: single
// buffered (22 KB, 8 CTAs per SM), the next item's ghosts are requested when the current item is done
// The component loop is ROLLED (one copy of the 8-cell body instead of three: 64 KB of SASS thrashed
// the instruction cache with only three warps per scheduler -- ncu r01: 1.4 warps per issue stalled on
// instruction fetch).  What depends on the component is selected at run time: the line of u_c comes
// from the shared core, the summation order of main.c:5016-5017 from three selects.
template <typename Real, int NS, int MINB, bool RK>
__global__ void __launch_bounds__(TPB, MINB) k_advdiff_tma(LevelView lv, AdvArgs<Real> A,
                                                           const __grid_constant__ AdvMaps M) {
  __shared__ __align__(128) Real core[3][CORE];
  __shared__ __align__(128) Real halo[NS][HALO];
  __shared__ __align__(8) uint64_t bar_core[3], bar_halo[2];
  const int t = threadIdx.x, x = t & 7, y = t >> 3;
  const Real *rsl = lv.rslab ? rslab_of<Real>(lv) : nullptr;
  auto rem = [&](int nbc, int c, int l, int e) -> Real {
    return __ldcg(rsl + (size_t)(kRemote0 - nbc) * (64 * kSlabPlanes) + (c * 3 + l) * 64 + e);
  };
  // blocks of this CTA: wi = blockIdx.x + j * gridDim.x
  const int nmine = A.nsub > (int)blockIdx.x ? (A.nsub - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (nmine == 0)
    return;
  auto block_of = [&](int j) -> int {
    const int wi = blockIdx.x + j * gridDim.x;
    return A.sub ? A.sub[wi] : wi;
  };
  auto load_nb = [&](int b, int (&nb)[6]) {
    const int *p = lv.nbr + (size_t)b * 6;
#pragma unroll
    for (int f = 0; f < 6; f++)
      nb[f] = p[f];
  };
  if (t == 0) {
#pragma unroll
    for (int c = 0; c < 3; c++)
      mbar_init(&bar_core[c], 1);
    mbar_init(&bar_halo[0], 1);
    mbar_init(&bar_halo[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  int bcur = block_of(0), nbc[6], bnext = -1, nbn[6];
  load_nb(bcur, nbc);
  if (nmine > 1) {
    bnext = block_of(1);
    load_nb(bnext, nbn);
  }
  if (t == 0) {
    const size_t own = (size_t)lv.act[bcur] * 512;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      mbar_arrive_expect_tx(&bar_core[c], CORE * sizeof(Real));
      tma_load_1d(core[c], A.vel[c] + own, CORE * sizeof(Real), &bar_core[c]);
    }
    issue_halo<Real>(halo[0], &bar_halo[0], A.vel[0], &M.x[0], &M.y[0], nbc);
    if (NS == 2)
      issue_halo<Real>(halo[1], &bar_halo[1], A.vel[1], &M.x[1], &M.y[1], nbc);
  }
  int item = 0;
  for (int j = 0; j < nmine; j++) {
    const size_t own = (size_t)lv.act[bcur] * 512;
    Real fac_a = A.fac_a0, fac_d = A.fac_d0;
    if (A.hblk) {  // per-block factors on multi-level meshes (main.c:4993-4995)
      const Real hb = A.hblk[bcur], h3b = hb * hb * hb;
      fac_a = -A.dtnu_dt / hb * h3b;
      fac_d = (A.dtnu_nu / hb) * (A.dtnu_dt / hb) * h3b;
    }
    bool odd = false;  // any face the copy engine did not fill
#pragma unroll
    for (int f = 0; f < 6; f++)
      odd |= nbc[f] < 0;
    // U of all three components: once per block into registers when the double-buffered variant is
    // shared-memory limited anyway (NS == 2); read cell by cell from the resident cores in the
    // single-buffered variant, whose eight CTAs per SM leave 128 registers per thread
    Real vv[3][NS == 2 ? 8 : 1];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      mbar_wait(&bar_core[c], j & 1);
      if (NS == 2) {
#pragma unroll
        for (int k = 0; k < 8; k++)
          vv[c][NS == 2 ? k : 0] = core[c][k * 64 + t];
      }
    }
#pragma unroll 1
    for (int c = 0; c < 3; c++, item++) {
      const int s = NS == 2 ? (item & 1) : 0;
      Real *H = halo[s];
      const Real *C = core[c];
      // the accumulator is read-modify-write: get the reads going before anything else
      // (only where the register budget allows it: NS == 2)
      Real acc[8];
      if (NS == 2) {
#pragma unroll
        for (int k = 0; k < 8; k++)
          acc[k] = A.tmp[c][own + k * 64 + t];
      }
      mbar_wait(&bar_halo[s], NS == 2 ? (item >> 1) & 1 : item & 1);
      if (odd) {
        const Real sx = (c == 0) ? (Real)-1 : (Real)1, sy = (c == 1) ? (Real)-1 : (Real)1,
                   sz = (c == 2) ? (Real)-1 : (Real)1;
        // thread t <-> face element (a = t & 7, c2 = t >> 3): (y, z) on x faces, (x, z) on y faces
        const int a = x, c2 = y;
#pragma unroll
        for (int p = 0; p < 3; p++) {
          if (nbc[0] < 0)
            H[O_XLO + (c2 * 8 + a) * 4 + 1 + p] = nbc[0] == kWall ? sx * C[c2 * 64 + a * 8] : rem(nbc[0], c, 2 - p, t);
          if (nbc[1] < 0)
            H[O_XHI + (c2 * 8 + a) * 4 + p] = nbc[1] == kWall ? sx * C[c2 * 64 + a * 8 + 7] : rem(nbc[1], c, p, t);
          if (nbc[2] < 0)
            H[O_YLO + c2 * 24 + p * 8 + a] = nbc[2] == kWall ? sy * C[c2 * 64 + a] : rem(nbc[2], c, 2 - p, t);
          if (nbc[3] < 0)
            H[O_YHI + c2 * 24 + p * 8 + a] = nbc[3] == kWall ? sy * C[c2 * 64 + 56 + a] : rem(nbc[3], c, p, t);
          if (nbc[4] < 0)
            H[O_ZLO + p * 64 + t] = nbc[4] == kWall ? sz * C[t] : rem(nbc[4], c, 2 - p, t);
          if (nbc[5] < 0)
            H[O_ZHI + p * 64 + t] = nbc[5] == kWall ? sz * C[448 + t] : rem(nbc[5], c, p, t);
        }
        // these generic-proxy writes are followed (next use of the slot) by async-proxy writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
      }
      // thread-constant addressing of the six x and six y neighbours: inside the core (plane
      // stride 64) or in a ghost array (plane stride 32 / 24).  Offsets are relative to C.
      const int hoff = (int)(H - C);
      int ox[6], sxk[6], oy[6], syk[6];
#pragma unroll
      for (int m = 0; m < 6; m++) {
        const int d = m < 3 ? m - 3 : m - 2;  // -3 -2 -1 +1 +2 +3
        const int xi = x + d, yi = y + d;
        if (xi < 0) {
          ox[m] = hoff + O_XLO + y * 4 + xi + 4;
          sxk[m] = 32;
        } else if (xi > 7) {
          ox[m] = hoff + O_XHI + y * 4 + xi - 8;
          sxk[m] = 32;
        } else {
          ox[m] = y * 8 + xi;
          sxk[m] = 64;
        }
        if (yi < 0) {
          oy[m] = hoff + O_YLO + (yi + 3) * 8 + x;
          syk[m] = 24;
        } else if (yi > 7) {
          oy[m] = hoff + O_YHI + (yi - 8) * 8 + x;
          syk[m] = 24;
        } else {
          oy[m] = yi * 8 + x;
          syk[m] = 64;
        }
      }
      Real line[14];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        line[i] = H[O_ZLO + i * 64 + t];
        line[11 + i] = H[O_ZHI + i * 64 + t];
      }
#pragma unroll
      for (int k = 0; k < 8; k++)
        line[3 + k] = C[k * 64 + t];  // == vv[c][k]; from shared so that c can stay a run-time value
      Real *oc = A.tmp[c];
      Real *vo = RK ? A.vout[c] : nullptr;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const Real u = line[3 + k];
        Real xv[6], yv[6];
#pragma unroll
        for (int m = 0; m < 6; m++) {
          xv[m] = C[ox[m] + k * sxk[m]];
          yv[m] = C[oy[m] + k * syk[m]];
        }
        Real dd[3], pr[3];
        const Real U0 = (NS == 2 ? vv[0][NS == 2 ? k : 0] : core[0][k * 64 + t]) + A.uinf[0],
                   U1 = (NS == 2 ? vv[1][NS == 2 ? k : 0] : core[1][k * 64 + t]) + A.uinf[1],
                   U2 = (NS == 2 ? vv[2][NS == 2 ? k : 0] : core[2][k * 64 + t]) + A.uinf[2];
        dd[0] = upwind<Real>(U0, xv[0], xv[1], xv[2], u, xv[3], xv[4], xv[5]);
        pr[0] = xv[3] + xv[2];
        dd[1] = upwind<Real>(U1, yv[0], yv[1], yv[2], u, yv[3], yv[4], yv[5]);
        pr[1] = yv[3] + yv[2];
        dd[2] = upwind<Real>(U2, line[k], line[k + 1], line[k + 2], u, line[k + 4], line[k + 5], line[k + 6]);
        pr[2] = line[k + 4] + line[k + 2];
        // adv = U_c d_c + (U_a1 d_a1 + U_a2 d_a2), a1 = (c+1)%3, a2 = (c+2)%3 (main.c:5016): same order, by selects
        const Real p0 = U0 * dd[0], p1 = U1 * dd[1], p2 = U2 * dd[2];
        const Real pc = c == 0 ? p0 : (c == 1 ? p1 : p2);
        const Real pa = c == 0 ? p1 : (c == 1 ? p2 : p0);
        const Real pb = c == 0 ? p2 : (c == 1 ? p0 : p1);
        const Real adv = pc + (pa + pb);
        const Real qc = c == 0 ? pr[0] : (c == 1 ? pr[1] : pr[2]);
        const Real qa = c == 0 ? pr[1] : (c == 1 ? pr[2] : pr[0]);
        const Real qb = c == 0 ? pr[2] : (c == 1 ? pr[0] : pr[1]);
        const Real lap = (qc + (qa + qb)) - (Real)6 * u;
        Real T;
        if (NS == 2)
          T = acc[k] + (fac_a * adv + fac_d * lap);
        else
          T = oc[own + k * 64 + t] + (fac_a * adv + fac_d * lap);
        if (RK) {
          // the Runge-Kutta stage update of this cell (main.c:5047-5052): V += T ih3 ; TMP = T beta
          vo[own + k * 64 + t] = u + T * A.rk_ih3;
          oc[own + k * 64 + t] = T * A.rk_beta;
        } else {
          oc[own + k * 64 + t] = T;
        }
      }
      __syncthreads();  // core c and ghost slot s are free
      if (t == 0) {
        if (bnext >= 0) {
          mbar_arrive_expect_tx(&bar_core[c], CORE * sizeof(Real));
          tma_load_1d(core[c], A.vel[c] + (size_t)lv.act[bnext] * 512, CORE * sizeof(Real), &bar_core[c]);
        }
        if (NS == 2) {
          // item + 2: component (c + 2) % 3 of this block (c == 0) or of the next one
          if (c == 0)
            issue_halo<Real>(H, &bar_halo[s], A.vel[2], &M.x[2], &M.y[2], nbc);
          else if (bnext >= 0)
            issue_halo<Real>(H, &bar_halo[s], A.vel[c - 1], &M.x[c - 1], &M.y[c - 1], nbn);
        } else {
          // item + 1
          if (c < 2)
            issue_halo<Real>(H, &bar_halo[s], A.vel[c + 1], &M.x[c + 1], &M.y[c + 1], nbc);
          else if (bnext >= 0)
            issue_halo<Real>(H, &bar_halo[s], A.vel[0], &M.x[0], &M.y[0], nbn);
        }
      }
    }
    // advance to the next block; fetch the neighbour list of the one after
    bcur = bnext;
#pragma unroll
    for (int f = 0; f < 6; f++)
      nbc[f] = nbn[f];
    if (j + 2 < nmine) {
      bnext = block_of(j + 2);
      load_nb(bnext, nbn);
    } else {
      bnext = -1;
    }
  }
}

}  // namespace

template <typename Real>
int advdiff_tma_launch(CupCtx *c, LevelView lv, const int *d_sub, int nsub, const void *d_hblk, double dtnu_dt,
                       double dtnu_nu, double fac_a, double fac_d, const AdvRk *rk) {
  if (nsub <= 0)
    return CUP_OK;
  AdvMaps M;
  AdvArgs<Real> A;
  for (int q = 0; q < 3; q++) {
    CUtensorMap m[2];
    CUP_TRY(tma_slab_maps(c, c->state[CUP_F_VEL + q], m));
    M.x[q] = m[0];
    M.y[q] = m[1];
    A.vel[q] = (const Real *)c->state[CUP_F_VEL + q];
    A.tmp[q] = (Real *)c->state[CUP_F_TMP + q];
    A.uinf[q] = (Real)c->prm.uinf[q];
  }
  A.sub = d_sub;
  A.nsub = nsub;
  A.hblk = (const Real *)d_hblk;
  A.dtnu_dt = (Real)dtnu_dt;
  A.dtnu_nu = (Real)dtnu_nu;
  A.fac_a0 = (Real)fac_a;
  A.fac_d0 = (Real)fac_d;
  for (int q = 0; q < 3; q++)
    A.vout[q] = rk ? (Real *)rk->vout[q] : nullptr;
  A.rk_ih3 = rk ? (Real)rk->ih3 : (Real)0;
  A.rk_beta = rk ? (Real)rk->beta : (Real)0;
  static int ns = getenv("CUP_ADV_SLOTS") ? atoi(getenv("CUP_ADV_SLOTS")) : 2;
  static int per_sm = getenv("CUP_ADV_PER_SM") ? atoi(getenv("CUP_ADV_PER_SM")) : (ns == 2 ? 6 : 8);
  long long g = (long long)c->num_sms * per_sm;
  if (g > nsub)
    g = nsub;
  if (rk) {
    if (ns == 2)
      k_advdiff_tma<Real, 2, 6, true><<<(int)g, TPB, 0, c->stream>>>(lv, A, M);
    else
      k_advdiff_tma<Real, 1, 8, true><<<(int)g, TPB, 0, c->stream>>>(lv, A, M);
  } else if (ns == 2)
    k_advdiff_tma<Real, 2, 6, false><<<(int)g, TPB, 0, c->stream>>>(lv, A, M);
  else if (per_sm <= 7)
    k_advdiff_tma<Real, 1, 7, false><<<(int)g, TPB, 0, c->stream>>>(lv, A, M);
  else
    k_advdiff_tma<Real, 1, 8, false><<<(int)g, TPB, 0, c->stream>>>(lv, A, M);
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

template int advdiff_tma_launch<double>(CupCtx *, LevelView, const int *, int, const void *, double, double, double,
                                        double, const AdvRk *);
template int advdiff_tma_launch<float>(CupCtx *, LevelView, const int *, int, const void *, double, double, double,
                                       double, const AdvRk *);

}  // namespace cup
