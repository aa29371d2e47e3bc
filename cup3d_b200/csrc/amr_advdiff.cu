// amr_advdiff.cu -- k_advdiff (main.c:4986) on the leaves of a multi-level mesh.
//
// The 5th-order upwind stencil needs three ghost layers along each axis.  The
// reference fills them with its ss = 3 tables (gen_table.py, lab_load :3544):
//   same level   copy                                              (OP_COPY)
//   wall         nearest interior cell, normal component negated   (OP_BC)
//   finer        each ghost = mean of 2x2x2 fine cells             (OP_AVG8, fine())
//   coarser      layers 1,2: tangential quadratic interpolation on the first coarse
//                layer, blended with the block's own cells         (OP_FD, both blends)
//                layer 3   : second-order Taylor expansion around the coarse cell from
//                its 19-point neighbourhood                        (OP_INTERP)
// OP_INTERP samples the "coarse scratch": the level L-1 view of the solution around
// the block -- coarser leaves directly, same-level blocks averaged 2x2x2
// (same_cfill), beyond walls the boundary cell (bc on the coarse buffer).  Here
// that view is a function, cs_sample(), backed by a device hash of the leaves.
// Only axis-aligned ghosts are produced (k_advdiff reads nothing else).
// The diffusive flux correction (face_grad :4227 + fc_fill) is applied in the
// same pass, as in amr_kernels.cu.  Correctness first: interface blocks do a
// lot of scattered lookups; blocks with same-level/wall neighbours only cost
// what the uniform kernel costs.
#include "amr_kernels.cuh"
#include "amr_sample.cuh"
#include "cup_internal.h"
#include "mg_device.cuh"

namespace cup {

// OP_AVG8 ghost at distance g = 1..3 from face f (fine(), gen_table.py:93): fine layers
// 2g-2, 2g-1 counted from the face, 2x2 in the tangential plane; x outermost, z innermost
template <typename Real>
__device__ __forceinline__ Real fine_avg_layer(const Real *__restrict__ cp, const int *ext4, int f, int g, int a, int c) {
  const Real *fb = cp + (size_t)ext4[(a >> 2) + 2 * (c >> 2)] * 512;
  const int g1 = 2 * (a & 3), g2 = 2 * (c & 3);
  const int nlo = (f & 1) ? 2 * g - 2 : 8 - 2 * g;  // lower normal coordinate of the pair
  Real s = 0;
  bool first = true;
#pragma unroll
  for (int dx = 0; dx < 2; dx++)
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
      for (int dz = 0; dz < 2; dz++) {
        int dn, j1, j2;
        if (f < 2) {
          dn = dx; j1 = dy; j2 = dz;
        } else if (f < 4) {
          dn = dy; j1 = dx; j2 = dz;
        } else {
          dn = dz; j1 = dx; j2 = dy;
        }
        const Real val = fb[face_idx(f, nlo + dn, g1 + j1, g2 + j2)];
        s = first ? val : s + val;
        first = false;
      }
  return (Real)0.125 * s;
}

enum { AD_ROW = 24, AD_SLAB = 14 * 24 + 8 };  // bank-friendly padding, see stencil_kernels.cu

template <typename Real>
__global__ void __launch_bounds__(TPB) k_advdiff_amr(LevelView lv, const int *__restrict__ sub, int nsub, LeafGeom geo,
                                                     const Real *__restrict__ hblk,
                                                     const Real *__restrict__ v0, const Real *__restrict__ v1,
                                                     const Real *__restrict__ v2, Real *__restrict__ t0,
                                                     Real *__restrict__ t1, Real *__restrict__ t2, Real dt, Real nu,
                                                     Real ux, Real uy, Real uz) {
  __shared__ Real tile[8 * AD_SLAB];
  __shared__ Real patch[6][16];
  // level L-1 view behind a face towards a coarser leaf: 3 coarse layers x 6 x 6 (the 3x3x3 neighbourhoods of
  // the coarse cells under the third ghost layer), sampled ONCE per (block, component, face) -- the Taylor
  // expansion of every ghost cell then reads shared memory instead of probing the leaf hash 19 times
  __shared__ Real cscr[6][108];
  const int t = threadIdx.x, x = t & 7, y = t >> 3, a = t & 7, c2 = t >> 3;
  const Real *vel[3] = {v0, v1, v2};
  Real *tmp[3] = {t0, t1, t2};
  const Real uinf[3] = {ux, uy, uz};
  for (int wi = blockIdx.x; wi < nsub; wi += gridDim.x) {
    const int b = sub ? sub[wi] : wi;
    const int slot = lv.act[b];
    const size_t own = (size_t)slot * 512;
    const int *ext24 = lv.ext + (size_t)b * 24;
    int nb[6];
#pragma unroll
    for (int f = 0; f < 6; f++)
      nb[f] = lv.nbr[(size_t)b * 6 + f];
    const int L = geo.bijk[4 * b], bx = geo.bijk[4 * b + 1], by = geo.bijk[4 * b + 2], bz = geo.bijk[4 * b + 3];
    const Real h = hblk[b];
    // fac_a = -dt/h*h^3 ; fac_d = (nu/h)*(dt/h)*h^3   (main.c:4993-4995)
    const Real h3 = h * h * h;
    const Real fac_a = -dt / h * h3, fac_d = (nu / h) * (dt / h) * h3;
    Real vv[3][8];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int k = 0; k < 8; k++)
        vv[c][k] = vel[c][own + k * 64 + t];
#pragma unroll 1
    for (int c = 0; c < 3; c++) {
      const Real *vc = vel[c];
      __syncthreads();  // previous component's tile / patches consumed
      // coarse patches (first coarse layer, 4x4) of this component
#pragma unroll
      for (int f = 0; f < 6; f++)
        if (nb[f] == kCoarse)
          coarse_patch_load<Real>(vc + (size_t)ext24[f * 4] * 512, f, ext24[f * 4 + 1], t, patch[f]);
#pragma unroll 1
      for (int f = 0; f < 6; f++) {
        if (nb[f] != kCoarse)
          continue;
        const int d = f >> 1, t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
        const int corg[3] = {bx * 4, by * 4, bz * 4};  // the block's origin in level L-1 cells
        for (int e = t; e < 108; e += TPB) {
          const int n = e / 36, j = (e / 6) % 6, i = e % 6;
          int cc[3];
          cc[d] = (f & 1) ? corg[d] + 4 + n : corg[d] - 3 + n;
          cc[t1] = corg[t1] - 1 + i;
          cc[t2] = corg[t2] - 1 + j;
          cscr[f][e] = cs_sample<Real>(geo, vc, c, L - 1, cc[0], cc[1], cc[2]);
        }
      }
      __syncthreads();
      // ghost of face f at distance g for plane element (pa, pc)
      auto ghost = [&](int f, int g, int pa, int pc) -> Real {
        const int code = nb[f];
        const int d = f >> 1;
        if (code >= 0)
          return vc[(size_t)code * 512 + face_idx(f, (f & 1) ? g - 1 : 8 - g, pa, pc)];
        if (code == kWall) {
          const Real v = vc[own + face_idx(f, (f & 1) ? 7 : 0, pa, pc)];
          return c == d ? -v : v;
        }
        if (code == kFine)
          return fine_avg_layer<Real>(vc, ext24 + f * 4, f, g, pa, pc);
        // coarser neighbour
        if (g <= 2) {
          const Real vt = fd_tangential<Real>(patch[f], pa, pc);
          const Real bb = vc[own + face_idx(f, (f & 1) ? 7 : 0, pa, pc)];
          const Real cq = vc[own + face_idx(f, (f & 1) ? 6 : 1, pa, pc)];
          return g == 1 ? (Real)(1.0 / 15.0) * ((Real)8.0 * vt + ((Real)10.0 * bb - (Real)3.0 * cq))
                        : (Real)(1.0 / 15.0) * ((Real)24.0 * vt + ((Real)-15.0 * bb + (Real)6 * cq));
        }
        // third layer: fine cell n = -3 (or 10) along d, tangential (pa, pc): OP_INTERP (main.c:3439) around the
        // coarse cell under it -- the middle one of the three sampled layers
        const int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
        const int i1 = (pa >> 1) + 1, i2 = (pc >> 1) + 1;
        const Real *S0 = cscr[f];
        auto C3 = [&](int I, int J, int K) -> Real {
          const int o[3] = {I - 1, J - 1, K - 1};
          return S0[((1 + o[d]) * 6 + (i2 + o[t2])) * 6 + (i1 + o[t1])];
        };
        Real sg[3];
        {
          // child offsets: parity of the global fine cell (the block origin is even in every direction)
          const int fn = (f & 1) ? 10 : -3;
          sg[d] = (fn & 1) ? (Real)1 : (Real)-1;
          sg[t1] = (pa & 1) ? (Real)1 : (Real)-1;
          sg[t2] = (pc & 1) ? (Real)1 : (Real)-1;
        }
        const Real c111 = C3(1, 1, 1);
        const Real c011 = C3(0, 1, 1), c211 = C3(2, 1, 1), c101 = C3(1, 0, 1), c121 = C3(1, 2, 1), c110 = C3(1, 1, 0),
                   c112 = C3(1, 1, 2);
        const Real dudx = (Real)0.125 * (c211 - c011);
        const Real dudy = (Real)0.125 * (c121 - c101);
        const Real dudz = (Real)0.125 * (c112 - c110);
        const Real dudxdy = (Real)0.015625 * (((C3(0, 0, 1) + C3(2, 2, 1)) - C3(2, 0, 1)) - C3(0, 2, 1));
        const Real dudxdz = (Real)0.015625 * (((C3(0, 1, 0) + C3(2, 1, 2)) - C3(2, 1, 0)) - C3(0, 1, 2));
        const Real dudydz = (Real)0.015625 * (((C3(1, 0, 0) + C3(1, 2, 2)) - C3(1, 2, 0)) - C3(1, 0, 2));
        const Real lap =
            c111 + (Real)0.03125 * ((((((c011 + c211) + c101) + c121) + c110) + c112) + (Real)(-6.0) * c111);
        const Real sx = sg[0], sy = sg[1], sz = sg[2];
        return (((((lap + sx * dudx) + sy * dudy) + sz * dudz) + sx * sy * dudxdy) + sx * sz * dudxdz) + sy * sz * dudydz;
      };
      // z-extended line
      Real line[14];
#pragma unroll
      for (int k = 0; k < 8; k++)
        line[3 + k] = vv[c][k];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        line[i] = ghost(4, 3 - i, x, y);
        line[11 + i] = ghost(5, i + 1, x, y);
      }
#pragma unroll
      for (int k = 0; k < 8; k++)
        tile[k * AD_SLAB + (y + 3) * AD_ROW + (x + 3)] = vv[c][k];
#pragma unroll
      for (int p = 0; p < 3; p++) {
        tile[c2 * AD_SLAB + (a + 3) * AD_ROW + p] = ghost(0, 3 - p, a, c2);
        tile[c2 * AD_SLAB + (a + 3) * AD_ROW + 11 + p] = ghost(1, p + 1, a, c2);
        tile[c2 * AD_SLAB + p * AD_ROW + (a + 3)] = ghost(2, 3 - p, a, c2);
        tile[c2 * AD_SLAB + (11 + p) * AD_ROW + (a + 3)] = ghost(3, p + 1, a, c2);
      }
      __syncthreads();
      Real *oc = tmp[c];
      const int a1 = (c + 1) % 3, a2 = (c + 2) % 3;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const Real *row = tile + k * AD_SLAB + (y + 3) * AD_ROW + (x + 3);
        const Real u = line[3 + k];
        Real dd[3], pr[3];
        const Real U[3] = {vv[0][k] + uinf[0], vv[1][k] + uinf[1], vv[2][k] + uinf[2]};
#define UPW(UU, m3, m2, m1, p1, p2, p3)                                                                              \
  ((UU) > 0 ? ((((((Real)-2 * (m3) + (Real)15 * (m2)) - (Real)60 * (m1)) + (Real)20 * u) + (Real)30 * (p1) -       \
                (Real)3 * (p2)) / (Real)60.)                                                                         \
            : (((((((Real)2 * (p3) - (Real)15 * (p2)) + (Real)60 * (p1)) - (Real)20 * u) - (Real)30 * (m1)) +      \
                (Real)3 * (m2)) / (Real)60.))
        dd[0] = UPW(U[0], row[-3], row[-2], row[-1], row[1], row[2], row[3]);
        pr[0] = row[1] + row[-1];
        dd[1] = UPW(U[1], row[-3 * AD_ROW], row[-2 * AD_ROW], row[-AD_ROW], row[AD_ROW], row[2 * AD_ROW],
                    row[3 * AD_ROW]);
        pr[1] = row[AD_ROW] + row[-AD_ROW];
        dd[2] = UPW(U[2], line[k], line[k + 1], line[k + 2], line[k + 4], line[k + 5], line[k + 6]);
        pr[2] = line[k + 4] + line[k + 2];
#undef UPW
        const Real adv = U[c] * dd[c] + (U[a1] * dd[a1] + U[a2] * dd[a2]);
        const Real lap = (pr[c] + (pr[a1] + pr[a2])) - (Real)6 * u;
        oc[own + k * 64 + t] += fac_a * adv + fac_d * lap;
      }
      // diffusive flux correction at faces with finer neighbours (face_grad coef fac_d + fc_fill)
      __syncthreads();
#pragma unroll 1
      for (int f = 0; f < 6; f++) {
        if (nb[f] != kFine)
          continue;
        const int nI = (f & 1) ? 7 : 0;
        const int cell = face_idx(f, nI, a, c2);
        const Real uc = vc[own + cell];
        const Real Fown = fac_d * (uc - ghost(f, 1, a, c2));
        Real mine[16];
#pragma unroll
        for (int q = 0; q < 16; q++)
          mine[q] = vc[own + face_idx(f, nI, 4 * (a >> 2) + (q & 3), 4 * (c2 >> 2) + (q >> 2))];
        const Real *fb = vc + (size_t)ext24[f * 4 + (a >> 2) + 2 * (c2 >> 2)] * 512;
        const int nF0 = (f & 1) ? 0 : 7, nF1 = (f & 1) ? 1 : 6;
        const Real hf = (Real)0.5 * h;
        const Real facf = (nu / hf) * (dt / hf) * (hf * hf * hf);
        Real Ff[2][2];
#pragma unroll
        for (int j2 = 0; j2 < 2; j2++)
#pragma unroll
          for (int j1 = 0; j1 < 2; j1++) {
            const int g1 = 2 * (a & 3) + j1, g2 = 2 * (c2 & 3) + j2;
            const Real bbf = fb[face_idx(f, nF0, g1, g2)], cqf = fb[face_idx(f, nF1, g1, g2)];
            const Real gh = fd_ghost<Real>(mine, g1, g2, bbf, cqf);
            Ff[j2][j1] = facf * (bbf - gh);
          }
        const Real fsum = (Ff[0][0] + Ff[0][1]) + (Ff[1][0] + Ff[1][1]);
        oc[own + cell] += Fown + fsum;
        __syncthreads();
      }
    }
    __syncthreads();
  }
}

template <typename Real>
int advdiff_amr_launch(CupCtx *c, const Level &v, Real *const *S, const int *sub, int nsub) {
  LevelView lv{v.d_act, v.d_nbr, (int)v.act.size(), nullptr, nullptr, 0, v.d_ext};
  LeafGeom g;
  g.bijk = v.d_bijk;
  g.hkeys = v.d_hkeys;
  g.hvals = v.d_hvals;
  g.hmask = (unsigned long long)v.hkeys.size() - 1;
  for (int d = 0; d < 3; d++)
    g.bpd[d] = c->bpd[d];
  if (nsub < 0)
    nsub = (int)v.act.size();
  long long grid = (long long)c->num_sms * 4;
  if (grid > (long long)nsub)
    grid = nsub < 1 ? 1 : nsub;
  k_advdiff_amr<Real><<<(int)grid, TPB, 0, c->stream>>>(lv, sub, nsub, g, (const Real *)v.d_hblk, S[CUP_F_VEL], S[CUP_F_VEL + 1],
                                                       S[CUP_F_VEL + 2], S[CUP_F_TMP], S[CUP_F_TMP + 1],
                                                       S[CUP_F_TMP + 2], (Real)c->prm.dt, (Real)c->prm.nu,
                                                       (Real)c->prm.uinf[0], (Real)c->prm.uinf[1],
                                                       (Real)c->prm.uinf[2]);
  return CUP_OK;
}

template int advdiff_amr_launch<double>(CupCtx *, const Level &, double *const *, const int *, int);
template int advdiff_amr_launch<float>(CupCtx *, const Level &, float *const *, const int *, int);

}  // namespace cup
