// adapt_kernels.cu -- the data-parallel half of mesh_adapt (main.c:4012-4190) on the device:
//
//   k_gradchi   (:3649)  tagging marker: scan the extended chi neighbourhood of every block
//   mesh_refine (:3790)  second-order interpolation of F_PRES / F_VEL of a refined block into its 8 children
//   compression (:4129)  2x2x2 averages of the nine fields of 8 siblings into their parent
//
// so that an adapting time step moves only block LISTS across PCIe; the tree surgery (mesh_tag's
// decisions, mesh_fix's 2:1 balance, load balance) stays host code, as in the reference.
//
// k_gradchi and mesh_refine read TENSORIAL labs (te = 1: edges and corners filled, ss = 2 and 1).  The
// reference interprets gen_table.py's op lists; here the lab is a FUNCTION, lab_value(): for a cell of
// the extended cube it finds who covers it -- the block itself, a same-level leaf (OP_COPY), finer
// leaves (OP_AVG8), a coarser leaf (OP_FD on the faces, OP_INTERP behind edges and corners), or the
// wall (OP_BC: nearest interior cell, wall-normal vector component negated) -- through the hash of the
// leaves this rank can read (amr_sample.cuh).
#include <algorithm>
#include <unordered_map>
#include <vector>

#include "amr_kernels.cuh"
#include "amr_sample.cuh"
#include "comm.cuh"
#include "cup_internal.h"
#include "mg_device.cuh"

namespace cup {

namespace {

// lab_load (main.c:3544) for te = 1, ss <= 2, one cell: local coordinates (x, y, z) in [-ss, 8+ss) of the
// block (L, bx, by, bz) stored at `slot`; cidx >= 0: the field is component cidx of a vector (vflip)
template <typename Real>
__device__ Real lab_value(const LeafGeom &g, const Real *__restrict__ cp, int cidx, int L, int bx, int by, int bz,
                          int slot, int x, int y, int z) {
  Real sign = 1;
  int q[3] = {bx * 8 + x, by * 8 + y, bz * 8 + z};
  const int org[3] = {bx * 8, by * 8, bz * 8};
#pragma unroll
  for (int d = 0; d < 3; d++) {  // OP_BC: beyond a wall the nearest interior cell counts
    const int n = (g.bpd[d] << L) * 8;
    if (q[d] < 0) {
      q[d] = 0;
      if (cidx == d)
        sign = -sign;
    } else if (q[d] >= n) {
      q[d] = n - 1;
      if (cidx == d)
        sign = -sign;
    }
  }
  const int l[3] = {q[0] - org[0], q[1] - org[1], q[2] - org[2]};
  const int code[3] = {l[0] < 0 ? -1 : (l[0] > 7 ? 1 : 0), l[1] < 0 ? -1 : (l[1] > 7 ? 1 : 0),
                       l[2] < 0 ? -1 : (l[2] > 7 ? 1 : 0)};
  if (code[0] == 0 && code[1] == 0 && code[2] == 0)
    return sign * cp[(size_t)slot * 512 + (l[2] << 6) + (l[1] << 3) + l[0]];
  int s = leaf_find(g, L, q[0] >> 3, q[1] >> 3, q[2] >> 3);
  if (s >= 0)  // same level: OP_COPY
    return sign * cp[(size_t)s * 512 + ((q[2] & 7) << 6) + ((q[1] & 7) << 3) + (q[0] & 7)];
  if (L > 0 && (s = leaf_find(g, L - 1, q[0] >> 4, q[1] >> 4, q[2] >> 4)) >= 0) {
    const int nz = (code[0] != 0) + (code[1] != 0) + (code[2] != 0);
    if (nz == 1) {
      // a face towards ONE coarser leaf: OP_FD (main.c:3465) -- tangential quadratic interpolation on the 4x4
      // coarse cells facing the block, blended with the block's own first two cells behind the face
      const int d = code[0] != 0 ? 0 : (code[1] != 0 ? 1 : 2);
      const int f = 2 * d + (code[d] > 0 ? 1 : 0);
      const int layer = code[d] < 0 ? -l[d] : l[d] - 7;  // 1 or 2
      const int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
      const int bidx[3] = {bx, by, bz};
      const int quad = (bidx[t1] & 1) + 2 * (bidx[t2] & 1);
      const Real *cb = cp + (size_t)s * 512;
      Real patch[16];
#pragma unroll
      for (int t = 0; t < 16; t++)
        patch[t] = cb[face_idx(f, (f & 1) ? 0 : 7, (t & 3) + 4 * (quad & 1), (t >> 2) + 4 * (quad >> 1))];
      const int pa = l[t1], pc = l[t2];
      const Real vt = fd_tangential<Real>(patch, pa, pc);
      const Real *ob = cp + (size_t)slot * 512;
      const Real bb = ob[face_idx(f, (f & 1) ? 7 : 0, pa, pc)], cq = ob[face_idx(f, (f & 1) ? 6 : 1, pa, pc)];
      const Real v = layer == 1 ? (Real)(1.0 / 15.0) * ((Real)8.0 * vt + ((Real)10.0 * bb - (Real)3.0 * cq))
                                : (Real)(1.0 / 15.0) * ((Real)24.0 * vt + ((Real)-15.0 * bb + (Real)6 * cq));
      return sign * v;
    }
    // behind an edge or a corner: OP_INTERP (main.c:3439), Taylor expansion around the coarse cell
    return sign * interp_ghost<Real>(g, cp, cidx, L - 1, q[0] >> 1, q[1] >> 1, q[2] >> 1,
                                     (q[0] & 1) ? (Real)1 : (Real)-1, (q[1] & 1) ? (Real)1 : (Real)-1,
                                     (q[2] & 1) ? (Real)1 : (Real)-1);
  }
  // finer leaves: OP_AVG8 of the 2x2x2 cells of level L + 1 (x outermost, z innermost: gen_table.py fine())
  const int fx = 2 * q[0], fy = 2 * q[1], fz = 2 * q[2];
  s = leaf_find(g, L + 1, fx >> 3, fy >> 3, fz >> 3);
  if (s < 0)
    return 0;  // not reachable on a 2:1 balanced mesh
  const Real *b = cp + (size_t)s * 512 + ((fz & 7) << 6) + ((fy & 7) << 3) + (fx & 7);
  const Real sum = ((((((b[0] + b[64]) + b[8]) + b[72]) + b[1]) + b[65]) + b[9]) + b[73];
  return sign * (Real)0.125 * sum;
}

// k_gradchi (main.c:3649-3680).  The reference scans the extended cube (offset 2 on the finest level, 1
// elsewhere) in z, y, x order, zeroes F_TMP at every inside-block cell with chi > 0.9 it passes, and at the
// FIRST cell with 1e-5 < chi < 0.9 writes 1e10 into the eight central cells of F_TMP[0] and stops.  Same
// result here: the index of that first cell is a block-wide minimum, the zeroing applies to the cells
// scanned before it.
template <typename Real>
__global__ void __launch_bounds__(128) k_gradchi(LeafGeom g, const int *__restrict__ sub, int nwork,
                                                 const Real *__restrict__ chi, Real *__restrict__ t0,
                                                 Real *__restrict__ t1, Real *__restrict__ t2, int level_max) {
  __shared__ unsigned int first;
  for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
    const int b = sub ? sub[wi] : wi;
    const int L = g.bijk[4 * b], bx = g.bijk[4 * b + 1], by = g.bijk[4 * b + 2], bz = g.bijk[4 * b + 3];
    const int o = (L == level_max - 1) ? 2 : 1, n = 8 + 2 * o, total = n * n * n;
    if (threadIdx.x == 0)
      first = 0xffffffffu;
    __syncthreads();
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int x = idx % n - o, y = (idx / n) % n - o, z = idx / (n * n) - o;
      Real v = lab_value<Real>(g, chi, -1, L, bx, by, bz, b, x, y, z);
      v = (Real)1.0 < v ? (Real)1.0 : v;
      v = v < (Real)0.0 ? (Real)0.0 : v;
      if (v > 0.00001 && v < 0.9)
        atomicMin(&first, (unsigned int)idx);
    }
    __syncthreads();
    const unsigned int stop = first;
    for (int j = threadIdx.x; j < 512; j += blockDim.x) {
      const int x = j & 7, y = (j >> 3) & 7, z = j >> 6;
      const unsigned int idx = (unsigned int)(((z + o) * n + (y + o)) * n + (x + o));
      Real v = chi[(size_t)b * 512 + j];
      v = (Real)1.0 < v ? (Real)1.0 : v;
      if (idx < stop && v > 0.9) {
        t0[(size_t)b * 512 + j] = 0;
        t1[(size_t)b * 512 + j] = 0;
        t2[(size_t)b * 512 + j] = 0;
      }
    }
    __syncthreads();
    if (stop != 0xffffffffu && threadIdx.x < 8) {
      const int qq = threadIdx.x;
      t0[(size_t)b * 512 + ((3 + (qq >> 2)) << 6) + ((3 + ((qq >> 1) & 1)) << 3) + 3 + (qq & 1)] = (Real)1e10;
    }
    __syncthreads();
  }
}

// mesh_refine (main.c:3790-3837): old block `par[r]` -> the eight new blocks kid[r][K*4 + J*2 + I], fields
// F_PRES, F_VEL (lab: ss = 1, tensorial, vflip), every other field of the children stays zero
template <typename Real>
struct RefineArgs {
  const Real *src[4];  // old F_PRES, F_VEL x3
  Real *dst[4];        // new arrays
};

template <typename Real>
__global__ void __launch_bounds__(64) k_refine(LeafGeom g, const int *__restrict__ par, const int *__restrict__ kid,
                                               int nref, RefineArgs<Real> A) {
  __shared__ Real lab[1000];
  const int t = threadIdx.x;
  for (int r = blockIdx.x; r < nref; r += gridDim.x) {
    const int b = par[r];
    const int L = g.bijk[4 * b], bx = g.bijk[4 * b + 1], by = g.bijk[4 * b + 2], bz = g.bijk[4 * b + 3];
    for (int c = 0; c < 4; c++) {
      __syncthreads();
      for (int i = t; i < 1000; i += 64)
        lab[i] = lab_value<Real>(g, A.src[c], c - 1, L, bx, by, bz, b, i % 10 - 1, (i / 10) % 10 - 1, i / 100 - 1);
      __syncthreads();
#define LB(X, Y, Z) lab[(((Z) + 1) * 10 + ((Y) + 1)) * 10 + (X) + 1]
      for (int cell = t; cell < 512; cell += 64) {
        const int x = cell & 7, y = (cell >> 3) & 7, z = cell >> 6;
        const Real dudx = (Real)0.5 * (LB(x + 1, y, z) - LB(x - 1, y, z));
        const Real dudy = (Real)0.5 * (LB(x, y + 1, z) - LB(x, y - 1, z));
        const Real dudz = (Real)0.5 * (LB(x, y, z + 1) - LB(x, y, z - 1));
        const Real dudx2 = (LB(x + 1, y, z) + LB(x - 1, y, z)) - (Real)2.0 * LB(x, y, z);
        const Real dudy2 = (LB(x, y + 1, z) + LB(x, y - 1, z)) - (Real)2.0 * LB(x, y, z);
        const Real dudz2 = (LB(x, y, z + 1) + LB(x, y, z - 1)) - (Real)2.0 * LB(x, y, z);
        const Real dudxdy = (Real)0.25 * ((LB(x + 1, y + 1, z) + LB(x - 1, y - 1, z)) -
                                          (LB(x + 1, y - 1, z) + LB(x - 1, y + 1, z)));
        const Real dudxdz = (Real)0.25 * ((LB(x + 1, y, z + 1) + LB(x - 1, y, z - 1)) -
                                          (LB(x + 1, y, z - 1) + LB(x - 1, y, z + 1)));
        const Real dudydz = (Real)0.25 * ((LB(x, y + 1, z + 1) + LB(x, y - 1, z - 1)) -
                                          (LB(x, y + 1, z - 1) + LB(x, y - 1, z + 1)));
        const Real u = LB(x, y, z);
        const Real lap = (Real)0.03125 * ((dudx2 + dudy2) + dudz2);
        const int I = x >> 2, J = y >> 2, K = z >> 2;
        Real *out = A.dst[c] + (size_t)kid[r * 8 + K * 4 + J * 2 + I] * 512;
        const int i0 = 2 * (x & 3), j0 = 2 * (y & 3), k0 = 2 * (z & 3);
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const Real sx = (q & 1) ? (Real)1.0 : (Real)-1.0, sy = (q & 2) ? (Real)1.0 : (Real)-1.0,
                     sz = (q & 4) ? (Real)1.0 : (Real)-1.0;
          out[((k0 + (q >> 2)) << 6) + ((j0 + ((q >> 1) & 1)) << 3) + i0 + (q & 1)] =
              ((u + (Real)0.25 * ((sx * dudx + sy * dudy) + sz * dudz)) + lap) +
              (Real)0.0625 * ((sx * sy * dudxdy + sx * sz * dudxdz) + sy * sz * dudydz);
        }
      }
#undef LB
    }
    __syncthreads();
  }
}

// compression (main.c:4129-4146): new block dst[r] = 2x2x2 averages of the old siblings kid[r][K*4+J*2+I]
template <typename Real>
__global__ void __launch_bounds__(256) k_compress(const int *__restrict__ dst, const int *__restrict__ kid, int ncom,
                                                  const Real *__restrict__ src, Real *__restrict__ out) {
  for (int r = blockIdx.x; r < ncom; r += gridDim.x) {
    Real *o = out + (size_t)dst[r] * 512;
    for (int cell = threadIdx.x; cell < 512; cell += blockDim.x) {
      const int X = cell & 7, Y = (cell >> 3) & 7, Z = cell >> 6;
      const int I = X >> 2, J = Y >> 2, K = Z >> 2;
      const Real *s = src + (size_t)kid[r * 8 + K * 4 + J * 2 + I] * 512 + ((2 * (Z & 3)) << 6) + ((2 * (Y & 3)) << 3) +
                      2 * (X & 3);
      // CELL(i,j,k) + CELL(i+1,j+1,k+1), ... in the reference's pairing
      o[cell] = (Real)0.125 * ((((s[0] + s[73]) + (s[1] + s[72])) + (s[8] + s[65])) + (s[9] + s[64]));
    }
  }
}

template <typename Real>
__global__ void __launch_bounds__(256) k_copy_blocks(const int *__restrict__ dst, const int *__restrict__ srcb, int n,
                                                     const Real *__restrict__ src, Real *__restrict__ out) {
  for (int r = blockIdx.x; r < n; r += gridDim.x)
    for (int j = threadIdx.x; j < 512; j += blockDim.x)
      out[(size_t)dst[r] * 512 + j] = src[(size_t)srcb[r] * 512 + j];
}

LeafGeom leaf_geom(const CupCtx *c) {
  const Level &v = c->leafv;
  LeafGeom g;
  g.bijk = v.d_bijk;
  g.hkeys = v.d_hkeys;
  g.hvals = v.d_hvals;
  g.hmask = (unsigned long long)v.hkeys.size() - 1;
  for (int d = 0; d < 3; d++)
    g.bpd[d] = c->bpd[d];
  return g;
}

template <typename T>
int upl(T **d, const std::vector<T> &h) {
  *d = nullptr;
  if (h.empty())
    return CUP_OK;
  CUP_CUDA(cudaMalloc((void **)d, h.size() * sizeof(T)));
  CUP_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return CUP_OK;
}

template <typename Real>
int gradchi_t(CupCtx *c) {
  if (c->nranks > 1 && c->leaf_uniform) {
    set_error("k_gradchi on a uniform mesh split over ranks is not available (edge / corner neighbours of other "
              "ranks travel as ghost blocks on multi-level meshes only)");
    return CUP_ERR_UNSUPPORTED;
  }
  Real **S = (Real **)c->state;
  if (c->nranks > 1) {  // chi of the leaves other ranks own
    const Real *src[1] = {S[CUP_F_CHI]};
    Real *dst[1] = {S[CUP_F_CHI]};
    CUP_TRY(block_exchange_leaf<Real>(c, src, dst, 1, 0));
  }
  const int nwork = c->run_nsub >= 0 ? c->run_nsub : (int)c->nblk;
  const int grid = (int)std::min<long long>(nwork, (long long)c->num_sms * 8);
  k_gradchi<Real><<<grid < 1 ? 1 : grid, 128, 0, c->stream>>>(leaf_geom(c), c->run_nsub >= 0 ? c->run_sub : nullptr,
                                                            nwork, S[CUP_F_CHI], S[CUP_F_TMP], S[CUP_F_TMP + 1],
                                                            S[CUP_F_TMP + 2], c->level_max);
  c->launches++;
  CUP_CUDA(cudaGetLastError());
  return CUP_OK;
}

}  // namespace

int gradchi(CupCtx *c) { return c->real_bytes == 8 ? gradchi_t<double>(c) : gradchi_t<float>(c); }

namespace {

template <typename Real>
int adapt_fields_t(CupCtx *c, long long n_new, const std::vector<int> &keep_dst, const std::vector<int> &keep_src,
                   const std::vector<int> &ref_par, const std::vector<int> &ref_kid, const std::vector<int> &com_dst,
                   const std::vector<int> &com_kid, void *out[CUP_F_N]) {
  const size_t bytes = (size_t)n_new * 512 * sizeof(Real);
  for (int f = 0; f < CUP_F_N; f++) {
    out[f] = nullptr;
    CUP_CUDA(cudaMalloc(&out[f], bytes));
    CUP_CUDA(cudaMemsetAsync(out[f], 0, bytes, c->stream));  // children: every field but F_PRES / F_VEL is zero
  }
  int *d_kd = nullptr, *d_ks = nullptr, *d_rp = nullptr, *d_rk = nullptr, *d_cd = nullptr, *d_ck = nullptr;
  CUP_TRY(upl(&d_kd, keep_dst));
  CUP_TRY(upl(&d_ks, keep_src));
  CUP_TRY(upl(&d_rp, ref_par));
  CUP_TRY(upl(&d_rk, ref_kid));
  CUP_TRY(upl(&d_cd, com_dst));
  CUP_TRY(upl(&d_ck, com_kid));
  Real **S = (Real **)c->state;
  const int cap = c->num_sms * 8;
  if (!keep_dst.empty())
    for (int f = 0; f < CUP_F_N; f++) {
      k_copy_blocks<Real><<<std::min<int>((int)keep_dst.size(), cap), 256, 0, c->stream>>>(
          d_kd, d_ks, (int)keep_dst.size(), S[f], (Real *)out[f]);
      c->launches++;
    }
  if (!ref_par.empty()) {
    RefineArgs<Real> A;
    const int fl[4] = {CUP_F_PRES, CUP_F_VEL, CUP_F_VEL + 1, CUP_F_VEL + 2};
    for (int q = 0; q < 4; q++) {
      A.src[q] = S[fl[q]];
      A.dst[q] = (Real *)out[fl[q]];
    }
    k_refine<Real><<<std::min<int>((int)ref_par.size(), cap), 64, 0, c->stream>>>(leaf_geom(c), d_rp, d_rk,
                                                                                (int)ref_par.size(), A);
    c->launches++;
  }
  if (!com_dst.empty())
    for (int f = 0; f < CUP_F_N; f++) {
      k_compress<Real><<<std::min<int>((int)com_dst.size(), cap), 256, 0, c->stream>>>(
          d_cd, d_ck, (int)com_dst.size(), S[f], (Real *)out[f]);
      c->launches++;
    }
  CUP_CUDA(cudaGetLastError());
  CUP_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_kd);
  cudaFree(d_ks);
  cudaFree(d_rp);
  cudaFree(d_rk);
  cudaFree(d_cd);
  cudaFree(d_ck);
  return CUP_OK;
}

}  // namespace

// Build the nine fields of the NEW block list from the old mesh's fields (still loaded): out[f] = device
// arrays [n_new][512].  kind / src: see cup_mesh_adapt in the header.
int adapt_fields(CupCtx *c, const CupBlk *nb, long long n_new, const int *kind, const long long *src,
                 void *out[CUP_F_N]) {
  if (c->nranks > 1) {
    set_error("cup_mesh_adapt: one rank only (the reference rebalances blocks between ranks while it adapts, "
              "mesh_bal main.c:3900ff; that data movement is not built)");
    return CUP_ERR_UNSUPPORTED;
  }
  const long long n_old = c->nblk;
  std::unordered_map<unsigned long long, int> old;
  auto key = [](int L, int x, int y, int z) {
    return ((unsigned long long)L << 57) | ((unsigned long long)z << 38) | ((unsigned long long)y << 19) |
           (unsigned long long)x;
  };
  old.reserve((size_t)n_old * 2);
  for (long long i = 0; i < n_old; i++)
    old.emplace(key(c->blk[(size_t)i].level, c->blk[(size_t)i].ix, c->blk[(size_t)i].iy, c->blk[(size_t)i].iz), (int)i);
  std::vector<int> keep_dst, keep_src, ref_par, ref_kid, com_dst, com_kid;
  std::unordered_map<long long, int> ref_of;  // old parent -> entry
  for (long long i = 0; i < n_new; i++) {
    const CupBlk &b = nb[i];
    const long long s = src[i];
    if (s < 0 || s >= n_old) {
      set_error("cup_mesh_adapt: src[%lld] = %lld outside the old mesh", i, s);
      return CUP_ERR_ARG;
    }
    const CupBlk &o = c->blk[(size_t)s];
    if (kind[i] == 0) {
      if (o.level != b.level || o.ix != b.ix || o.iy != b.iy || o.iz != b.iz) {
        set_error("cup_mesh_adapt: new block %lld is not old block %lld", i, s);
        return CUP_ERR_ARG;
      }
      keep_dst.push_back((int)i);
      keep_src.push_back((int)s);
    } else if (kind[i] == 1) {
      if (b.level != o.level + 1 || b.ix / 2 != o.ix || b.iy / 2 != o.iy || b.iz / 2 != o.iz) {
        set_error("cup_mesh_adapt: new block %lld is not a child of old block %lld", i, s);
        return CUP_ERR_ARG;
      }
      auto it = ref_of.find(s);
      int e;
      if (it == ref_of.end()) {
        e = (int)ref_par.size();
        ref_of.emplace(s, e);
        ref_par.push_back((int)s);
        ref_kid.resize(ref_kid.size() + 8, -1);
      } else {
        e = it->second;
      }
      ref_kid[(size_t)e * 8 + (b.iz & 1) * 4 + (b.iy & 1) * 2 + (b.ix & 1)] = (int)i;
    } else if (kind[i] == 2) {
      if (o.level != b.level + 1 || o.ix / 2 != b.ix || o.iy / 2 != b.iy || o.iz / 2 != b.iz) {
        set_error("cup_mesh_adapt: old block %lld is not a child of new block %lld", s, i);
        return CUP_ERR_ARG;
      }
      com_dst.push_back((int)i);
      for (int q = 0; q < 8; q++) {
        auto it = old.find(key(o.level, 2 * b.ix + (q & 1), 2 * b.iy + ((q >> 1) & 1), 2 * b.iz + (q >> 2)));
        if (it == old.end()) {
          set_error("cup_mesh_adapt: a sibling of old block %lld is missing", s);
          return CUP_ERR_MESH;
        }
        com_kid.push_back(it->second);
      }
    } else {
      set_error("cup_mesh_adapt: kind[%lld] = %d", i, kind[i]);
      return CUP_ERR_ARG;
    }
  }
  for (int k : ref_kid)
    if (k < 0) {
      set_error("cup_mesh_adapt: a refined block has fewer than 8 children in the new list");
      return CUP_ERR_MESH;
    }
  return c->real_bytes == 8
             ? adapt_fields_t<double>(c, n_new, keep_dst, keep_src, ref_par, ref_kid, com_dst, com_kid, out)
             : adapt_fields_t<float>(c, n_new, keep_dst, keep_src, ref_par, ref_kid, com_dst, com_kid, out);
}

}  // namespace cup
