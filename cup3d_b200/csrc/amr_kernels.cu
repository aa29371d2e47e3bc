// amr_kernels.cu -- the multigrid / Poisson kernels on levels with coarse-fine
// interfaces (block-structured AMR).  Same decomposition as the uniform
// kernels (one 8^3 block per 64 threads, z-line per thread); the ghost faces
// additionally come from one coarser leaf (OP_FD interpolation) or from four
// finer leaves (OP_AVG8), and the leaf-level operator applies the reference's
// flux correction (fc_fill, main.c:3228) in the same pass.
//
// Reference: lab_load :3544, lab_exec :3401, gen_table.py (ss = 1, te = 0),
// face_grad :4227, fc_prepare/fc_fill :3155-3293, mg_smooth/down/tau :4689-4770.
#include "amr_kernels.cuh"
#include "cup_internal.h"
#include "mg_device.cuh"

namespace cup {

__constant__ double cFDp[9];
__constant__ double cFDm[9];

int amr_setup_constants() {
  // d_coef_plus / d_coef_minus, main.c:3371-3374
  const double p[9] = {-0.09375, 0.4375, 0.15625, 0.15625, -0.5625, 0.90625, -0.09375, 0.4375, 0.15625};
  const double m[9] = {0.15625, -0.5625, 0.90625, -0.09375, 0.4375, 0.15625, 0.15625, 0.4375, -0.09375};
  CUP_CUDA(cudaMemcpyToSymbol(cFDp, p, sizeof p));
  CUP_CUDA(cudaMemcpyToSymbol(cFDm, m, sizeof m));
  return CUP_OK;
}

// Ghost faces of block b into halo[6][64] for a level that may have coarser neighbours
// (multigrid contexts) or coarser and finer ones (leaf context).  usame: vector holding the
// same-level neighbours (the ping-pong source); ucan: canonical vector holding coarser / finer
// leaves.  Two phases around one __syncthreads (the coarse patches are shared).
template <typename Real>
__device__ __forceinline__ void halo_amr_phase1(const SlotVec<Real> &usame, const SlotVec<Real> &ucan, const Real *own,
                                                const int *nbr6, const int *ext24, int t, Real (*halo)[64],
                                                Real (*patch)[16]) {
  const int a = t & 7, c = t >> 3;
#pragma unroll
  for (int f = 0; f < 6; f++) {
    const int nb = nbr6[f];
    if (nb == kCoarse) {
      coarse_patch_load<Real>(ucan.at(ext24[f * 4]), f, ext24[f * 4 + 1], t, patch[f]);
    } else if (nb == kFine) {
      Real lay[2][2][2];
      halo[f][t] = fine_avg<Real>(ucan, ext24 + f * 4, f, a, c, lay);
    } else {
      const Real *src = nb >= 0 ? usame.at(nb) : own;
      const int p = (nb >= 0) ? ((f & 1) ? 0 : 7) : ((f & 1) ? 7 : 0);
      halo[f][t] = src[face_idx(f, p, a, c)];
    }
  }
}

template <typename Real>
__device__ __forceinline__ void halo_amr_phase2(const Real *own, const int *nbr6, int t, Real (*halo)[64],
                                                const Real (*patch)[16]) {
  const int a = t & 7, c = t >> 3;
#pragma unroll
  for (int f = 0; f < 6; f++)
    if (nbr6[f] == kCoarse) {
      const Real bb = own[face_idx(f, (f & 1) ? 7 : 0, a, c)], cq = own[face_idx(f, (f & 1) ? 6 : 1, a, c)];
      halo[f][t] = fd_ghost<Real>(patch[f], a, c, bb, cq);
    }
}

// mg_smooth on a level with coarser neighbours.  Same algebraic form as k_smooth (the ghost of
// a coarse face depends on the block's own OLD values through the blend, which is exactly what
// A u_old contains).
template <typename Real>
__global__ void __launch_bounds__(TPB) k_smooth_amr(LevelView lv, const int *__restrict__ sub, int nsub,
                                                    SlotVec<Real> usrc, SlotVec<Real> ucan,
                                                    SlotVec<Real> udst, SlotVec<Real> fvec,
                                                    const Real *__restrict__ Wl, Real h, Real invh, Real omega,
                                                    const double *__restrict__ fmean, int zero_src) {
  __shared__ Real ex[512];
  __shared__ Real halo[6][64];
  __shared__ Real patch[6][16];
  const int t = threadIdx.x, x = t & 7, y = t >> 3;
  Real w[8];
#pragma unroll
  for (int k = 0; k < 8; k++)
    w[k] = Wl[k * 64 + t];
  const Real q0 = fmean ? (Real)(*fmean) : (Real)0;
  for (int i = blockIdx.x; i < nsub; i += gridDim.x) {
    const int b = sub ? sub[i] : i;
    const int slot = lv.act[b];
    const Real *fb = fvec.at(slot);
    Real uu[8], v[8];
#pragma unroll
    for (int k = 0; k < 8; k++)
      v[k] = fb[k * 64 + t];
    if (!zero_src) {
      const Real *ub = usrc.at(slot);
      const int *nbr6 = lv.nbr + (size_t)b * 6;
#pragma unroll
      for (int k = 0; k < 8; k++)
        uu[k] = ub[k * 64 + t];
      halo_amr_phase1<Real>(usrc, ucan, ub, nbr6, lv.ext + (size_t)b * 24, t, halo, patch);
      __syncthreads();
      halo_amr_phase2<Real>(ub, nbr6, t, halo, patch);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; k++)
        v[k] = invh * ((v[k] - q0) - h * ghost_sum<Real>(halo, x, y, k));
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        uu[k] = 0;
        v[k] = invh * (v[k] - q0);
      }
    }
    fdm_solve<Real>(v, ex, w, t);
    Real *ob = udst.at(slot);
#pragma unroll
    for (int k = 0; k < 8; k++)
      ob[k * 64 + t] = uu[k] + omega * (v[k] - uu[k]);
    __syncthreads();
  }
}

// mg_down on a level with coarser neighbours
template <typename Real>
__global__ void __launch_bounds__(TPB) k_down_amr(LevelView lv, const int *__restrict__ sub, int nsub,
                                                  const int *__restrict__ pslot, const int *__restrict__ oct,
                                                  SlotVec<Real> u, SlotVec<Real> f, Real h,
                                                  Real *const *__restrict__ rptr) {
  __shared__ Real tu[512];
  __shared__ Real tr[512];
  __shared__ Real halo[6][64];
  __shared__ Real patch[6][16];
  const int t = threadIdx.x, x = t & 7, y = t >> 3;
  for (int i = blockIdx.x; i < nsub; i += gridDim.x) {
    const int b = sub ? sub[i] : i;
    const int slot = lv.act[b];
    const Real *ub = u.at(slot);
    const Real *fb = f.at(slot);
    const int *nbr6 = lv.nbr + (size_t)b * 6;
    Real uu[8], ff[8], tt[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      uu[k] = ub[k * 64 + t];
      ff[k] = fb[k * 64 + t];
      tu[k * 64 + t] = uu[k];
    }
    halo_amr_phase1<Real>(u, u, ub, nbr6, lv.ext + (size_t)b * 24, t, halo, patch);
    __syncthreads();
    halo_amr_phase2<Real>(ub, nbr6, t, halo, patch);
    __syncthreads();
    lap_line<Real>(tu, halo, uu, x, y, t, h, tt);
#pragma unroll
    for (int k = 0; k < 8; k++)
      tr[k * 64 + t] = ff[k] - tt[k];
    __syncthreads();
    {
      const int cx = t & 3, cy = (t >> 2) & 3, cz = t >> 4;
      const int base = ((2 * cz) << 6) + ((2 * cy) << 3) + 2 * cx;
      const Real sr = ((((((tr[base] + tr[base + 1]) + tr[base + 8]) + tr[base + 9]) + tr[base + 64]) +
                         tr[base + 65]) + tr[base + 72]) + tr[base + 73];
      const Real su = ((((((tu[base] + tu[base + 1]) + tu[base + 8]) + tu[base + 9]) + tu[base + 64]) +
                         tu[base + 65]) + tu[base + 72]) + tu[base + 73];
      const int o = oct[b], ps = pslot[b];
      if (ps >= 0) {
        const int pidx = ((4 * (o >> 2) + cz) << 6) + ((4 * ((o >> 1) & 1) + cy) << 3) + 4 * (o & 1) + cx;
        f.at(ps)[pidx] = sr;
        u.at(ps)[pidx] = (Real)0.125 * su;
      } else {  // parent on another rank: 64 r + 64 u into its owner's window (MG_M layout, main.c:4750)
        Real *q = rptr[kRemote0 - ps];
        q[t] = sr;
        q[64 + t] = (Real)0.125 * su;
      }
    }
    __syncthreads();
  }
}

// out = A u on blocks sub[] of a level / of the leaf context.
//   TAU : out += A u, us = u                      (mg_tau)
//   FC  : leaf operator with flux correction      (k_lhs + fc_fill: coarse cells facing finer
//         blocks get  + h_c (u_c - ghost) + sum of the 2x2 fine fluxes h_f (u_f - ghost_f) )
//   per-block h from lv-level hblk when given (leaf context), else the level's h
template <typename Real, bool TAU, bool FC>
__global__ void __launch_bounds__(TPB) k_apply_amr(LevelView lv, const int *__restrict__ sub, int nsub,
                                                   SlotVec<Real> u, SlotVec<Real> out, SlotVec<Real> us, Real hlev,
                                                   const Real *__restrict__ hblk, const double *__restrict__ shift) {
  __shared__ Real tu[512];
  __shared__ Real halo[6][64];
  __shared__ Real patch[6][16];
  const int t = threadIdx.x, x = t & 7, y = t >> 3;
  const int a = t & 7, c = t >> 3;
  for (int i = blockIdx.x; i < nsub; i += gridDim.x) {
    const int b = sub ? sub[i] : i;
    const int slot = lv.act[b];
    const Real *ub = u.at(slot);
    const int *nbr6 = lv.nbr + (size_t)b * 6;
    const int *ext24 = lv.ext + (size_t)b * 24;
    const Real h = hblk ? hblk[b] : hlev;
    Real uu[8], tt[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      uu[k] = ub[k * 64 + t];
      tu[k * 64 + t] = uu[k];
    }
    halo_amr_phase1<Real>(u, u, ub, nbr6, ext24, t, halo, patch);
    __syncthreads();
    halo_amr_phase2<Real>(ub, nbr6, t, halo, patch);
    __syncthreads();
    lap_line<Real>(tu, halo, uu, x, y, t, h, tt);
    Real *ob = out.at(slot);
    if (TAU) {
      Real *sb = us.at(slot);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        ob[k * 64 + t] += tt[k];
        sb[k * 64 + t] = uu[k];
      }
    } else {
      const Real add = shift ? (Real)(*shift) * (h * h * h) : (Real)0;
#pragma unroll
      for (int k = 0; k < 8; k++)
        ob[k * 64 + t] = tt[k] + add;
    }
    if (FC) {
      // flux correction at faces with finer neighbours; face order 0..5 == the reference's
      // d = 0,1,2 loop, so a cell on an edge receives its additions in the same order
      __syncthreads();  // ob of this block is complete (other threads' cells are updated below)
#pragma unroll
      for (int f = 0; f < 6; f++) {
        if (nbr6[f] != kFine)
          continue;
        // this thread's face cell (a, c): own flux h (u_c - ghost)
        const int nI = (f & 1) ? 7 : 0;
        const int cell = face_idx(f, nI, a, c);
        const Real Fown = h * (tu[cell] - halo[f][t]);
        // the finer blocks' fluxes on the 2x2 fine cells facing this cell
        Real lay[2][2][2];
        (void)fine_avg<Real>(u, ext24 + f * 4, f, a, c, lay);
        // their ghosts: OP_FD from MY 4x4 cells of the quadrant (a>>2, c>>2)
        Real mine[16];
#pragma unroll
        for (int q = 0; q < 16; q++)
          mine[q] = tu[face_idx(f, nI, 4 * (a >> 2) + (q & 3), 4 * (c >> 2) + (q >> 2))];
        const Real hf = (Real)0.5 * h;
        Real Ff[2][2];
#pragma unroll
        for (int j2 = 0; j2 < 2; j2++)
#pragma unroll
          for (int j1 = 0; j1 < 2; j1++) {
            const int g1 = 2 * (a & 3) + j1, g2 = 2 * (c & 3) + j2;
            const Real gh = fd_ghost<Real>(mine, g1, g2, lay[0][j2][j1], lay[1][j2][j1]);
            Ff[j2][j1] = hf * (lay[0][j2][j1] - gh);
          }
        const Real fsum = (Ff[0][0] + Ff[0][1]) + (Ff[1][0] + Ff[1][1]);  // fc_fill :3245-3246
        ob[cell] += Fown + fsum;
        __syncthreads();
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// Pressure sweeps on the leaves of a multi-level mesh: k_divp (:5700) and k_gradp (:5715)
// with their flux correction (face_grad :4227 / face_sum :4241 + fc_fill).
//   WHAT 0: o0 = h (sum of 6 neighbours - 6 p)             [k_divp summation order]
//   WHAT 1: o_a = fac (p(+a) - p(-a)), fac = facs * h^2     [facs = -dt/2]
template <typename Real, int WHAT>
__global__ void __launch_bounds__(TPB) k_pres_amr(LevelView lv, const Real *__restrict__ hblk,
                                                  const Real *__restrict__ p, Real *__restrict__ o0,
                                                  Real *__restrict__ o1, Real *__restrict__ o2, Real facs) {
  __shared__ Real tu[512];
  __shared__ Real halo[6][64];
  __shared__ Real patch[6][16];
  const int t = threadIdx.x, x = t & 7, y = t >> 3, a = t & 7, c = t >> 3;
  SlotVec<Real> pv{const_cast<Real *>(p), nullptr, 0x7fffffff};
  Real *outs[3] = {o0, o1, o2};
  for (int wi = blockIdx.x; wi < lv_count(lv); wi += gridDim.x) {
    const int b = lv_item(lv, wi);
    const size_t own = (size_t)lv.act[b] * 512;
    const int *nbr6 = lv.nbr + (size_t)b * 6;
    const int *ext24 = lv.ext + (size_t)b * 24;
    const Real h = hblk[b];
    const Real fac = WHAT == 0 ? h : facs * h * h;
    Real uu[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      uu[k] = p[own + k * 64 + t];
      tu[k * 64 + t] = uu[k];
    }
    halo_amr_phase1<Real>(pv, pv, p + own, nbr6, ext24, t, halo, patch);
    __syncthreads();
    halo_amr_phase2<Real>(p + own, nbr6, t, halo, patch);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int i = k * 64 + t;
      const Real xm = x > 0 ? tu[i - 1] : halo[0][y + 8 * k];
      const Real xp = x < 7 ? tu[i + 1] : halo[1][y + 8 * k];
      const Real ym = y > 0 ? tu[i - 8] : halo[2][x + 8 * k];
      const Real yp = y < 7 ? tu[i + 8] : halo[3][x + 8 * k];
      const Real zm = k > 0 ? uu[k > 0 ? k - 1 : 0] : halo[4][t];
      const Real zp = k < 7 ? uu[k < 7 ? k + 1 : 7] : halo[5][t];
      if (WHAT == 0) {
        o0[own + i] = fac * ((((((xp + xm) + yp) + ym) + zp) + zm) - (Real)6.0 * uu[k]);
      } else {
        o0[own + i] = fac * (xp - xm);
        o1[own + i] = fac * (yp - ym);
        o2[own + i] = fac * (zp - zm);
      }
    }
    __syncthreads();
#pragma unroll
    for (int f = 0; f < 6; f++) {
      if (nbr6[f] != kFine)
        continue;
      const int nI = (f & 1) ? 7 : 0;
      const int cell = face_idx(f, nI, a, c);
      Real lay[2][2][2], mine[16];
      (void)fine_avg<Real>(pv, ext24 + f * 4, f, a, c, lay);
#pragma unroll
      for (int q = 0; q < 16; q++)
        mine[q] = tu[face_idx(f, nI, 4 * (a >> 2) + (q & 3), 4 * (c >> 2) + (q >> 2))];
      const Real hf = (Real)0.5 * h;
      Real Fown, Ff[2][2];
      if (WHAT == 0) {
        Fown = fac * (tu[cell] - halo[f][t]);  // face_grad: coef (c - n)
      } else {
        const Real sgn = (f & 1) ? -fac : fac;  // face_sum: s (n + c)
        Fown = sgn * (halo[f][t] + tu[cell]);
      }
#pragma unroll
      for (int j2 = 0; j2 < 2; j2++)
#pragma unroll
        for (int j1 = 0; j1 < 2; j1++) {
          const Real gh = fd_ghost<Real>(mine, 2 * (a & 3) + j1, 2 * (c & 3) + j2, lay[0][j2][j1], lay[1][j2][j1]);
          if (WHAT == 0) {
            Ff[j2][j1] = hf * (lay[0][j2][j1] - gh);
          } else {
            const Real facf = facs * hf * hf;
            const Real sf = ((f ^ 1) & 1) ? -facf : facf;  // the fine block's face is f^1
            Ff[j2][j1] = sf * (gh + lay[0][j2][j1]);
          }
        }
      const Real fsum = (Ff[0][0] + Ff[0][1]) + (Ff[1][0] + Ff[1][1]);
      Real *o = WHAT == 0 ? o0 : outs[f >> 1];
      o[own + cell] += Fown + fsum;
      __syncthreads();
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// k_prhs (:5663) on the leaves of a multi-level mesh, with flux correction.
// Only the face-normal components are needed on each face (u, udef_x on x faces, ...).
template <typename Real>
__global__ void __launch_bounds__(TPB) k_prhs_amr(LevelView lv, const Real *__restrict__ hblk,
                                                  const Real *__restrict__ v0, const Real *__restrict__ v1,
                                                  const Real *__restrict__ v2, const Real *__restrict__ d0,
                                                  const Real *__restrict__ d1, const Real *__restrict__ d2,
                                                  const Real *__restrict__ chi, Real *__restrict__ lhs, Real idt2) {
  __shared__ Real tl[6][512];     // u v w udef_x udef_y udef_z cores
  __shared__ Real hl[2][6][64];   // [vel / udef][face]: ghost plane of the face-normal component
  __shared__ Real patch[2][6][16];
  const int t = threadIdx.x, x = t & 7, y = t >> 3, a = t & 7, c = t >> 3;
  const Real *comp[6] = {v0, v1, v2, d0, d1, d2};
  for (int wi = blockIdx.x; wi < lv_count(lv); wi += gridDim.x) {
    const int b = lv_item(lv, wi);
    const size_t own = (size_t)lv.act[b] * 512;
    const int *nbr6 = lv.nbr + (size_t)b * 6;
    const int *ext24 = lv.ext + (size_t)b * 24;
    const Real h = hblk[b];
    const Real fac = (Real)0.5 * h * h * idt2;  // 0.5 h^2 / dt
#pragma unroll
    for (int q = 0; q < 6; q++)
#pragma unroll
      for (int k = 0; k < 8; k++)
        tl[q][k * 64 + t] = comp[q][own + k * 64 + t];
    // phase 1: ghost planes of component (f/2) and (3 + f/2) on face f
#pragma unroll
    for (int g = 0; g < 2; g++)
#pragma unroll
      for (int f = 0; f < 6; f++) {
        const Real *cp = comp[3 * g + (f >> 1)];
        SlotVec<Real> cv{const_cast<Real *>(cp), nullptr, 0x7fffffff};
        const int nb = nbr6[f];
        if (nb == kCoarse) {
          coarse_patch_load<Real>(cv.at(ext24[f * 4]), f, ext24[f * 4 + 1], t, patch[g][f]);
        } else if (nb == kFine) {
          Real lay[2][2][2];
          hl[g][f][t] = fine_avg<Real>(cv, ext24 + f * 4, f, a, c, lay);
        } else if (nb >= 0) {
          hl[g][f][t] = cp[(size_t)nb * 512 + face_idx(f, (f & 1) ? 0 : 7, a, c)];
        } else {  // wall: nearest interior cell, normal component negated (OP_BC, vflip)
          hl[g][f][t] = -cp[own + face_idx(f, (f & 1) ? 7 : 0, a, c)];
        }
      }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < 2; g++)
#pragma unroll
      for (int f = 0; f < 6; f++)
        if (nbr6[f] == kCoarse) {
          const Real *tc = tl[3 * g + (f >> 1)];
          hl[g][f][t] = fd_ghost<Real>(patch[g][f], a, c, tc[face_idx(f, (f & 1) ? 7 : 0, a, c)],
                                       tc[face_idx(f, (f & 1) ? 6 : 1, a, c)]);
        }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int i = k * 64 + t;
      Real dv[2];
#pragma unroll
      for (int g = 0; g < 2; g++) {
        const Real *U = tl[3 * g], *V = tl[3 * g + 1], *W = tl[3 * g + 2];
        const Real uxp = x < 7 ? U[i + 1] : hl[g][1][y + 8 * k];
        const Real uxm = x > 0 ? U[i - 1] : hl[g][0][y + 8 * k];
        const Real vyp = y < 7 ? V[i + 8] : hl[g][3][x + 8 * k];
        const Real vym = y > 0 ? V[i - 8] : hl[g][2][x + 8 * k];
        const Real wzp = k < 7 ? W[i + 64] : hl[g][5][t];
        const Real wzm = k > 0 ? W[i - 64] : hl[g][4][t];
        dv[g] = ((((uxp - uxm) + vyp) - vym) + wzp) - wzm;
      }
      Real pp = fac * dv[0];
      pp += -chi[own + i] * fac * dv[1];
      lhs[own + i] = pp;
    }
    __syncthreads();
#pragma unroll
    for (int f = 0; f < 6; f++) {
      if (nbr6[f] != kFine)
        continue;
      const int d = f >> 1, nI = (f & 1) ? 7 : 0;
      const int cell = face_idx(f, nI, a, c);
      const Real s = (f & 1) ? (Real)-1.0 : (Real)1.0;
      // own flux (k_prhs :5694): s (fac (n_d + c_d) - chi fac (n_{3+d} + c_{3+d}))
      const Real Fown = s * (fac * (hl[0][f][t] + tl[d][cell]) - chi[own + cell] * fac * (hl[1][f][t] + tl[3 + d][cell]));
      // the four fine cells facing this cell
      const Real hf = (Real)0.5 * h, facf = (Real)0.5 * hf * hf * idt2;
      const Real sf = ((f ^ 1) & 1) ? (Real)-1.0 : (Real)1.0;
      Real gh[2][2][2], bnd[2][2][2];  // [vel/udef][j2][j1]
#pragma unroll
      for (int g = 0; g < 2; g++) {
        const Real *cp = comp[3 * g + d];
        SlotVec<Real> cv{const_cast<Real *>(cp), nullptr, 0x7fffffff};
        Real lay[2][2][2], mine[16];
        (void)fine_avg<Real>(cv, ext24 + f * 4, f, a, c, lay);
#pragma unroll
        for (int q = 0; q < 16; q++)
          mine[q] = tl[3 * g + d][face_idx(f, nI, 4 * (a >> 2) + (q & 3), 4 * (c >> 2) + (q >> 2))];
#pragma unroll
        for (int j2 = 0; j2 < 2; j2++)
#pragma unroll
          for (int j1 = 0; j1 < 2; j1++) {
            gh[g][j2][j1] = fd_ghost<Real>(mine, 2 * (a & 3) + j1, 2 * (c & 3) + j2, lay[0][j2][j1], lay[1][j2][j1]);
            bnd[g][j2][j1] = lay[0][j2][j1];
          }
      }
      // chi of the fine boundary cells
      const Real *fchi = chi + (size_t)ext24[f * 4 + (a >> 2) + 2 * (c >> 2)] * 512;
      const int nF = (f & 1) ? 0 : 7;
      Real Ff[2][2];
#pragma unroll
      for (int j2 = 0; j2 < 2; j2++)
#pragma unroll
        for (int j1 = 0; j1 < 2; j1++) {
          const Real cf = fchi[face_idx(f, nF, 2 * (a & 3) + j1, 2 * (c & 3) + j2)];
          Ff[j2][j1] = sf * (facf * (gh[0][j2][j1] + bnd[0][j2][j1]) - cf * facf * (gh[1][j2][j1] + bnd[1][j2][j1]));
        }
      const Real fsum = (Ff[0][0] + Ff[0][1]) + (Ff[1][0] + Ff[1][1]);
      lhs[own + cell] += Fown + fsum;
      __syncthreads();
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------

// ---------------------------------------------------------------------------
// k_vort (main.c:5736) / k_q (:5762) on the leaves of a multi-level mesh.  All three velocity
// components are needed on all six faces (same ghost fill as the pressure sweeps, per component).
// k_vort carries a flux correction: face_sum(f, in = 3-a-d, out = a, +-inv2h) for every face
// direction d and output component a != d (:5752-5758); k_q has none (outc = 0).
// ---------------------------------------------------------------------------
template <typename Real, int WHAT>  // 0 vorticity -> o0..o2, 1 Q -> o0
__global__ void __launch_bounds__(TPB) k_velgrad_amr(LevelView lv, const Real *__restrict__ hblk,
                                                     const Real *__restrict__ v0, const Real *__restrict__ v1,
                                                     const Real *__restrict__ v2, Real *__restrict__ o0,
                                                     Real *__restrict__ o1, Real *__restrict__ o2) {
  __shared__ Real tl[3][512];
  __shared__ Real hl[3][6][64];
  __shared__ Real patch[3][6][16];
  const int t = threadIdx.x, x = t & 7, y = t >> 3, a = t & 7, c = t >> 3;
  const Real *vel[3] = {v0, v1, v2};
  Real *outs[3] = {o0, o1, o2};
  for (int wi = blockIdx.x; wi < lv_count(lv); wi += gridDim.x) {
    const int b = lv_item(lv, wi);
    const size_t own = (size_t)lv.act[b] * 512;
    const int *nbr6 = lv.nbr + (size_t)b * 6;
    const int *ext24 = lv.ext + (size_t)b * 24;
    const Real h = hblk[b];
#pragma unroll
    for (int q = 0; q < 3; q++) {
      SlotVec<Real> pv{const_cast<Real *>(vel[q]), nullptr, 0x7fffffff};
#pragma unroll
      for (int k = 0; k < 8; k++)
        tl[q][k * 64 + t] = vel[q][own + k * 64 + t];
      halo_amr_phase1<Real>(pv, pv, vel[q] + own, nbr6, ext24, t, hl[q], patch[q]);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 3; q++) {
      halo_amr_phase2<Real>(vel[q] + own, nbr6, t, hl[q], patch[q]);
      // walls: the wall-normal component flips sign (OP_BC); phase 1 copied the boundary cell
#pragma unroll
      for (int f = 0; f < 6; f++)
        if (nbr6[f] == kWall && (f >> 1) == q)
          hl[q][f][t] = -hl[q][f][t];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int i = k * 64 + t;
      Real g[3][3];  // g[q][d] = u_q(+d) - u_q(-d)
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const Real xp = x < 7 ? tl[q][i + 1] : hl[q][1][y + 8 * k], xm = x > 0 ? tl[q][i - 1] : hl[q][0][y + 8 * k];
        const Real yp = y < 7 ? tl[q][i + 8] : hl[q][3][x + 8 * k], ym = y > 0 ? tl[q][i - 8] : hl[q][2][x + 8 * k];
        const Real zp = k < 7 ? tl[q][i + 64] : hl[q][5][t], zm = k > 0 ? tl[q][i - 64] : hl[q][4][t];
        g[q][0] = xp - xm;
        g[q][1] = yp - ym;
        g[q][2] = zp - zm;
      }
      if (WHAT == 0) {
        const Real inv2h = (Real).5 * h * h;
        o0[own + i] = inv2h * (g[2][1] - g[1][2]);
        o1[own + i] = inv2h * (g[0][2] - g[2][0]);
        o2[own + i] = inv2h * (g[1][0] - g[0][1]);
      } else {
        const Real inv2h = (Real).5 / h;
        Real gg[3][3];
#pragma unroll
        for (int q = 0; q < 3; q++)
#pragma unroll
          for (int d = 0; d < 3; d++)
            gg[q][d] = inv2h * g[q][d];
        Real qq = 0;
#pragma unroll
        for (int q = 0; q < 3; q++)
#pragma unroll
          for (int d = 0; d < 3; d++)
            qq -= (Real)0.5 * gg[q][d] * gg[d][q];
        o0[own + i] = qq;
      }
    }
    __syncthreads();
    if (WHAT == 0) {
#pragma unroll 1
      for (int f = 0; f < 6; f++) {
        if (nbr6[f] != kFine)
          continue;
        const int d = f >> 1, nI = (f & 1) ? 7 : 0;
        const int cell = face_idx(f, nI, a, c);
        const Real hf = (Real)0.5 * h;
        const Real i2h = (Real).5 * h * h, i2hf = (Real).5 * hf * hf;
#pragma unroll 1
        for (int oa = 0; oa < 3; oa++) {
          if (oa == d)
            continue;
          const int q = 3 - oa - d;                            // input component
          const Real sg = (d == (oa + 1) % 3) ? (Real)1 : (Real)-1;
          SlotVec<Real> pv{const_cast<Real *>(vel[q]), nullptr, 0x7fffffff};
          Real lay[2][2][2], mine[16];
          (void)fine_avg<Real>(pv, ext24 + f * 4, f, a, c, lay);
#pragma unroll
          for (int m = 0; m < 16; m++)
            mine[m] = tl[q][face_idx(f, nI, 4 * (a >> 2) + (m & 3), 4 * (c >> 2) + (m >> 2))];
          const Real so = (f & 1) ? -(sg * i2h) : sg * i2h;          // face_sum on my face f
          const Real sf = ((f ^ 1) & 1) ? -(sg * i2hf) : sg * i2hf;  // and on the fine blocks' face f^1
          const Real Fown = so * (hl[q][f][t] + tl[q][cell]);
          Real Ff[2][2];
#pragma unroll
          for (int j2 = 0; j2 < 2; j2++)
#pragma unroll
            for (int j1 = 0; j1 < 2; j1++) {
              const Real gh =
                  fd_ghost<Real>(mine, 2 * (a & 3) + j1, 2 * (c & 3) + j2, lay[0][j2][j1], lay[1][j2][j1]);
              Ff[j2][j1] = sf * (gh + lay[0][j2][j1]);
            }
          const Real fsum = (Ff[0][0] + Ff[0][1]) + (Ff[1][0] + Ff[1][1]);
          outs[oa][own + cell] += Fown + fsum;
        }
        __syncthreads();
      }
    }
    __syncthreads();
  }
}

static inline int agrid(const CupCtx *c, long long n) {
  long long g = (long long)c->num_sms * 8;
  return (int)(g < n ? g : (n < 1 ? 1 : n));
}

template <typename Real>
int smooth_amr_launch(CupCtx *c, LevelView lv, SlotVec<Real> src, SlotVec<Real> can, SlotVec<Real> dst,
                      SlotVec<Real> f, Real h, const double *fmean, bool zero_src, const int *sub, int nsub) {
  if (nsub < 0)
    nsub = lv.nact;
  k_smooth_amr<Real><<<agrid(c, nsub), TPB, 0, c->stream>>>(lv, sub, nsub, src, can, dst, f, (const Real *)c->d_W, h,
                                                               (Real)(1.0 / (double)h), (Real)0.8, fmean,
                                                               zero_src ? 1 : 0);
  return CUP_OK;
}

template <typename Real>
int down_amr_launch(CupCtx *c, LevelView lv, const int *pslot, const int *oct, SlotVec<Real> u, SlotVec<Real> f,
                    Real h, const int *sub, int nsub, void *const *rptr) {
  if (nsub < 0)
    nsub = lv.nact;
  k_down_amr<Real><<<agrid(c, nsub), TPB, 0, c->stream>>>(lv, sub, nsub, pslot, oct, u, f, h, (Real *const *)rptr);
  return CUP_OK;
}

template <typename Real>
int apply_amr_launch(CupCtx *c, LevelView lv, const int *sub, int nsub, SlotVec<Real> u, SlotVec<Real> out,
                     SlotVec<Real> us, Real h, const void *hblk, const double *shift, int mode) {
  if (mode == 1)
    k_apply_amr<Real, true, false><<<agrid(c, nsub), TPB, 0, c->stream>>>(lv, sub, nsub, u, out, us, h,
                                                                         (const Real *)hblk, shift);
  else if (mode == 2)
    k_apply_amr<Real, false, true><<<agrid(c, nsub), TPB, 0, c->stream>>>(lv, sub, nsub, u, out, us, h,
                                                                         (const Real *)hblk, shift);
  else
    k_apply_amr<Real, false, false><<<agrid(c, nsub), TPB, 0, c->stream>>>(lv, sub, nsub, u, out, us, h,
                                                                          (const Real *)hblk, shift);
  return CUP_OK;
}

template <typename Real>
int pres_amr_launch(CupCtx *c, LevelView lv, const void *hblk, const Real *p, Real *o0, Real *o1, Real *o2, Real facs,
                    int what) {
  if (what == 0)
    k_pres_amr<Real, 0><<<agrid(c, lv.nact), TPB, 0, c->stream>>>(lv, (const Real *)hblk, p, o0, o1, o2, facs);
  else
    k_pres_amr<Real, 1><<<agrid(c, lv.nact), TPB, 0, c->stream>>>(lv, (const Real *)hblk, p, o0, o1, o2, facs);
  return CUP_OK;
}

template <typename Real>
int prhs_amr_launch(CupCtx *c, LevelView lv, const void *hblk, Real *const *S, Real idt2) {
  k_prhs_amr<Real><<<agrid(c, lv.nact), TPB, 0, c->stream>>>(lv, (const Real *)hblk, S[CUP_F_VEL], S[CUP_F_VEL + 1],
                                                             S[CUP_F_VEL + 2], S[CUP_F_TMP], S[CUP_F_TMP + 1],
                                                             S[CUP_F_TMP + 2], S[CUP_F_CHI], S[CUP_F_LHS], idt2);
  return CUP_OK;
}

template <typename Real>
int velgrad_amr_launch(CupCtx *c, LevelView lv, const void *hblk, Real *const *S, int what) {
  if (what == 0)
    k_velgrad_amr<Real, 0><<<agrid(c, lv.nact), TPB, 0, c->stream>>>(lv, (const Real *)hblk, S[CUP_F_VEL],
                                                                     S[CUP_F_VEL + 1], S[CUP_F_VEL + 2], S[CUP_F_TMP],
                                                                     S[CUP_F_TMP + 1], S[CUP_F_TMP + 2]);
  else
    k_velgrad_amr<Real, 1><<<agrid(c, lv.nact), TPB, 0, c->stream>>>(lv, (const Real *)hblk, S[CUP_F_VEL],
                                                                     S[CUP_F_VEL + 1], S[CUP_F_VEL + 2], S[CUP_F_LHS],
                                                                     nullptr, nullptr);
  return CUP_OK;
}
template int velgrad_amr_launch<double>(CupCtx *, LevelView, const void *, double *const *, int);
template int velgrad_amr_launch<float>(CupCtx *, LevelView, const void *, float *const *, int);

template int pres_amr_launch<double>(CupCtx *, LevelView, const void *, const double *, double *, double *, double *,
                                     double, int);
template int pres_amr_launch<float>(CupCtx *, LevelView, const void *, const float *, float *, float *, float *, float,
                                    int);
template int prhs_amr_launch<double>(CupCtx *, LevelView, const void *, double *const *, double);
template int prhs_amr_launch<float>(CupCtx *, LevelView, const void *, float *const *, float);

template int smooth_amr_launch<double>(CupCtx *, LevelView, SlotVec<double>, SlotVec<double>, SlotVec<double>,
                                       SlotVec<double>, double, const double *, bool, const int *, int);
template int smooth_amr_launch<float>(CupCtx *, LevelView, SlotVec<float>, SlotVec<float>, SlotVec<float>,
                                      SlotVec<float>, float, const double *, bool, const int *, int);
template int down_amr_launch<double>(CupCtx *, LevelView, const int *, const int *, SlotVec<double>, SlotVec<double>,
                                     double, const int *, int, void *const *);
template int down_amr_launch<float>(CupCtx *, LevelView, const int *, const int *, SlotVec<float>, SlotVec<float>,
                                    float, const int *, int, void *const *);
template int apply_amr_launch<double>(CupCtx *, LevelView, const int *, int, SlotVec<double>, SlotVec<double>,
                                      SlotVec<double>, double, const void *, const double *, int);
template int apply_amr_launch<float>(CupCtx *, LevelView, const int *, int, SlotVec<float>, SlotVec<float>,
                                     SlotVec<float>, float, const void *, const double *, int);

}  // namespace cup
