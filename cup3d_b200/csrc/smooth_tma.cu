// smooth_tma.cu -- the production smoother: block-Jacobi sweep with TMA staging.
//
// Same arithmetic as k_smooth<Real,0> (mg_kernels.cu; reference mg_smooth,
// main.c:4689) but every global read goes through the TMA engine:
//
//   u block, f block      1-D bulk copies (4 KB each in fp64)
//   z faces               1-D bulk copies (one 8x8 plane is contiguous)
//   y faces               3-D tensor-map boxes {8 x, 1 y, 8 z}
//   x faces               3-D tensor-map boxes {16 bytes of x, 8 y, 8 z}
//
// all completing on one mbarrier.  The stage is consumed into registers at the
// top of an iteration, so the loads of the NEXT block are issued before the
// transforms of the current one start: HBM latency overlaps the FDM work
// without holding registers, and the strided x-face gather costs no LSU
// wavefronts.  One 8^3 block per 64-thread CTA iteration, persistent grid.
#include <cuda.h>

#include <map>
#include <tuple>

#include "cup_internal.h"
#include "mg_device.cuh"
#include "smooth_tma.cuh"
#include "stencil7_tma.cuh"
#include "comm_dev.cuh"
#include "tma.cuh"

namespace cup {

template <typename Real>
struct TmaCfg {
  static constexpr int NCOL = 16 / (int)sizeof(Real);  // x columns per 16-byte TMA row
  static constexpr uint32_t BYTES = (512 * 2 + 64 * 2 + 64 * 2 + 64 * NCOL * 2) * (uint32_t)sizeof(Real);
};

template <typename Real, bool COMM, bool UPF>
__global__ void __launch_bounds__(TPB, UPF ? 10 : 12)
    k_smooth_tma(LevelView lv, const int *__restrict__ sub, int nsub, FusedComm fc, UpFuse up, SlotVec<Real> usrc, SlotVec<Real> udst, SlotVec<Real> fvec, const Real *__restrict__ Wl,
                 Real h, Real invh, Real omega, const double *__restrict__ fmean,
                 const __grid_constant__ CUtensorMap mx_leaf, const __grid_constant__ CUtensorMap my_leaf,
                 const __grid_constant__ CUtensorMap mx_extra, const __grid_constant__ CUtensorMap my_extra,
                 const __grid_constant__ CUtensorMap md_oct, const __grid_constant__ CUtensorMap md_x,
                 const __grid_constant__ CUtensorMap md_y, const __grid_constant__ CUtensorMap md_z) {
  constexpr int NCOL = TmaCfg<Real>::NCOL;
  __shared__ __align__(128) Real s_d[UPF ? 64 : 1];                // parent's correction, own octant [z][y][x]
  __shared__ __align__(128) Real s_dz[2][UPF ? 16 : 1];            // neighbours' parents: 4x4 patches
  __shared__ __align__(128) Real s_dy[2][UPF ? 16 : 1];
  __shared__ __align__(128) Real s_dx[2][UPF ? 16 * NCOL : 1];
  __shared__ __align__(128) Real s_u[512];
  __shared__ __align__(128) Real s_f[512];
  __shared__ __align__(128) Real s_z[2][64];
  __shared__ __align__(128) Real s_y[2][64];
  __shared__ __align__(128) Real s_x[2][64 * NCOL];
  __shared__ __align__(128) Real ex[512];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ int s_wall;  // bit f: face f is a domain wall (ghost = own boundary plane); bit 8+f: face f
                          // came from another rank (compact 8x8 plane out of lv.rface)

  const int t = threadIdx.x, x = t & 7, y = t >> 3;
  const int G = gridDim.x;
  Real w[8];
#pragma unroll
  for (int k = 0; k < 8; k++)
    w[k] = Wl[k * 64 + t];
  const Real q0 = fmean ? (Real)(*fmean) : (Real)0;
  // sequence number of the faces this sweep CONSUMES (its own faces go out as seq0 + 1)
  unsigned long long seq0 = 0;
  const Real *rf = (const Real *)lv.rface;
  if (COMM) {
    seq0 = *(volatile const unsigned long long *)fc.seq;
    if (seq0 & 1)
      rf += lv.rface_stride;
  } else {
    rf = rface_of<Real>(lv);
  }

  // thread 0 is the TMA producer: stage one block (its u, f and six ghost faces)
  auto issue = [&](int slot, const int (&nb)[6], const int (&ui)[7]) {
    int wall = 0;
#pragma unroll
    for (int f = 0; f < 6; f++)
      wall |= ((nb[f] == kWall) << f) | ((nb[f] <= kRemote0) << (8 + f));
    s_wall = wall;
    // x faces from a local block arrive as 16-byte rows (NCOL values each), received ones compact
    uint32_t bytes = TmaCfg<Real>::BYTES;
    if (nb[0] <= kRemote0) bytes -= 64 * (NCOL - 1) * (uint32_t)sizeof(Real);
    if (nb[1] <= kRemote0) bytes -= 64 * (NCOL - 1) * (uint32_t)sizeof(Real);
    if (UPF) {
      bytes += 64 * (uint32_t)sizeof(Real);
#pragma unroll
      for (int f = 0; f < 6; f++)
        if (nb[f] > kRemote0)
          bytes += (f < 2 ? 16 * NCOL : 16) * (uint32_t)sizeof(Real);
    }
    mbar_arrive_expect_tx(&mbar, bytes);
    if (UPF) {
      // own octant of the parent's d, and for every face the 4x4 coarse cells behind it: the
      // neighbour's parent on the neighbour's facing side, the own parent on the own side at a wall
      tma_load_3d(s_d, &md_oct, 4 * (ui[0] & 1), 4 * ((ui[0] >> 1) & 1), (ui[0] >> 3) * 8 + 4 * ((ui[0] >> 2) & 1), &mbar);
#pragma unroll
      for (int f = 0; f < 6; f++) {
        if (nb[f] <= kRemote0)
          continue;
        const int w = nb[f] >= 0 ? ui[1 + f] : ui[0];
        const int ox = 4 * (w & 1), oy = 4 * ((w >> 1) & 1), oz = (w >> 3) * 8 + 4 * ((w >> 2) & 1);
        const bool high = (nb[f] >= 0) ? !(f & 1) : (f & 1);
        if (f < 2)
          tma_load_3d(s_dx[f], &md_x, ox + (high ? 4 - NCOL : 0), oy, oz, &mbar);
        else if (f < 4)
          tma_load_3d(s_dy[f - 2], &md_y, ox, oy + (high ? 3 : 0), oz, &mbar);
        else
          tma_load_3d(s_dz[f - 4], &md_z, ox, oy, oz + (high ? 3 : 0), &mbar);
      }
    }
    const Real *own = usrc.at(slot);
    tma_load_1d(s_u, own, 512 * sizeof(Real), &mbar);
    tma_load_1d(s_f, fvec.at(slot), 512 * sizeof(Real), &mbar);
    // z faces: neighbour's opposite plane, or own plane at a wall
    tma_load_1d(s_z[0],
                nb[4] >= 0 ? usrc.at(nb[4]) + 7 * 64 : (nb[4] == kWall ? own : rf + (size_t)(kRemote0 - nb[4]) * 64),
                64 * sizeof(Real), &mbar);
    tma_load_1d(s_z[1],
                nb[5] >= 0 ? usrc.at(nb[5]) : (nb[5] == kWall ? own + 7 * 64 : rf + (size_t)(kRemote0 - nb[5]) * 64),
                64 * sizeof(Real), &mbar);
#pragma unroll
    for (int f = 0; f < 4; f++) {
      if (nb[f] <= kRemote0) {  // received face: already a compact plane in the consumer's order
        tma_load_1d(f < 2 ? (Real *)s_x[f] : (Real *)s_y[f - 2], rf + (size_t)(kRemote0 - nb[f]) * 64,
                    64 * sizeof(Real), &mbar);
        continue;
      }
      const int ts = nb[f] >= 0 ? nb[f] : slot;
      const bool leaf = ts < usrc.nleaf;
      const int row0 = (leaf ? ts : ts - usrc.nleaf) * 8;
      // plane inside the source block: opposite side for a neighbour, same side at a wall
      const bool high = (nb[f] >= 0) ? !(f & 1) : (f & 1);
      if (f < 2)
        tma_load_3d(s_x[f], leaf ? &mx_leaf : &mx_extra, high ? 8 - NCOL : 0, 0, row0, &mbar);
      else
        tma_load_3d(s_y[f - 2], leaf ? &my_leaf : &my_extra, 0, high ? 7 : 0, row0, &mbar);
    }
  };

  if (t == 0) {
    mbar_init(&mbar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  // work items: sub[i] (indices into act[]) or 0..nsub-1
  int i = blockIdx.x;
  if (COMM && t == 0 && i < fc.nbnd) {
    // this CTA starts with blocks that read received faces: wait until every peer has published seq0
    // (acquire at system scope, then a proxy fence: the faces are read by the TMA engine)
    WaitDesc w;
    w.seq = fc.seq;
    w.flags = fc.my_flags;
    w.peers = fc.rpeers;
    w.np = fc.nrp;
    w.err = fc.err;
    w.code = fc.code;
    comm_wait_for(w, seq0);
  }
  if (t == 0 && i < nsub) {
    const int b0 = sub ? sub[i] : i;
    int nb[6], ui[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int f = 0; f < 6; f++)
      nb[f] = lv.nbr[(size_t)b0 * 6 + f];
    if (UPF) {
#pragma unroll
      for (int j = 0; j < 7; j++)
        ui[j] = up.info[(size_t)b0 * 7 + j];
    }
    issue(lv.act[b0], nb, ui);
  }
  uint32_t phase = 0;
  for (; i < nsub; i += G) {
    const int b = sub ? sub[i] : i;
    const int slot = lv.act[b];
    // producer: fetch the NEXT block's indices now so they are in registers when needed
    int nslot = 0, nnb[6] = {0, 0, 0, 0, 0, 0}, nui[7] = {0, 0, 0, 0, 0, 0, 0};
    const bool more = (i + G) < nsub;
    if (t == 0 && more) {
      const int bn = sub ? sub[i + G] : i + G;
      nslot = lv.act[bn];
#pragma unroll
      for (int f = 0; f < 6; f++)
        nnb[f] = lv.nbr[(size_t)bn * 6 + f];
      if (UPF) {
#pragma unroll
        for (int j = 0; j < 7; j++)
          nui[j] = up.info[(size_t)bn * 7 + j];
      }
    }
    mbar_wait(&mbar, phase);
    phase ^= 1;
    const int wall = s_wall;
    Real uu[8], v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      uu[k] = s_u[k * 64 + t];
      v[k] = s_f[k * 64 + t];
    }
    {
      // ghost sums (see ghost_sum in mg_device.cuh); x-face column depends on wall-ness
      const int cxm = (wall & 1) ? 0 : NCOL - 1, cxp = (wall & 2) ? NCOL - 1 : 0;
      const int stm = (wall & 0x100) ? 1 : NCOL, stp = (wall & 0x200) ? 1 : NCOL;
      const int om = (wall & 0x100) ? 0 : cxm, op = (wall & 0x200) ? 0 : cxp;
      // ghost sums, branch per face class (uniform per thread, so the k loops stay straight-line)
      Real g[8];
#pragma unroll
      for (int k = 0; k < 8; k++)
        g[k] = 0;
      g[0] = s_z[0][t];
      g[7] = s_z[1][t];
      if (x == 0) {
#pragma unroll
        for (int k = 0; k < 8; k++)
          g[k] += s_x[0][(k * 8 + y) * stm + om];
      }
      if (x == 7) {
#pragma unroll
        for (int k = 0; k < 8; k++)
          g[k] += s_x[1][(k * 8 + y) * stp + op];
      }
      if (y == 0) {
#pragma unroll
        for (int k = 0; k < 8; k++)
          g[k] += s_y[0][k * 8 + x];
      }
      if (y == 7) {
#pragma unroll
        for (int k = 0; k < 8; k++)
          g[k] += s_y[1][k * 8 + x];
      }
      if (UPF) {
        // prolongated correction: own cells ...
#pragma unroll
        for (int k = 0; k < 8; k++)
          uu[k] += s_d[(k >> 1) * 16 + (y >> 1) * 4 + (x >> 1)];
        // ... and the ghosts (faces received from other ranks already carry it)
        if (!(wall & 0x1000))
          g[0] += s_dz[0][(y >> 1) * 4 + (x >> 1)];
        if (!(wall & 0x2000))
          g[7] += s_dz[1][(y >> 1) * 4 + (x >> 1)];
        if (x == 0 && !(wall & 0x100)) {
          const int col = (wall & 1) ? 0 : NCOL - 1;
#pragma unroll
          for (int k = 0; k < 8; k++)
            g[k] += s_dx[0][((k >> 1) * 4 + (y >> 1)) * NCOL + col];
        }
        if (x == 7 && !(wall & 0x200)) {
          const int col = (wall & 2) ? NCOL - 1 : 0;
#pragma unroll
          for (int k = 0; k < 8; k++)
            g[k] += s_dx[1][((k >> 1) * 4 + (y >> 1)) * NCOL + col];
        }
        if (y == 0 && !(wall & 0x400)) {
#pragma unroll
          for (int k = 0; k < 8; k++)
            g[k] += s_dy[0][(k >> 1) * 4 + (x >> 1)];
        }
        if (y == 7 && !(wall & 0x800)) {
#pragma unroll
          for (int k = 0; k < 8; k++)
            g[k] += s_dy[1][(k >> 1) * 4 + (x >> 1)];
        }
      }
#pragma unroll
      for (int k = 0; k < 8; k++)
        v[k] = invh * ((v[k] - q0) - h * g[k]);
    }
    if (COMM && fc.push_mode == 2 && t == 0)  // ex doubles as the staging area of the bulk-store push:
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // its last reads are done before fdm_solve rewrites it
    __syncthreads();  // stage fully consumed
    if (t == 0 && more)
      issue(nslot, nnb, nui);
    fdm_solve<Real>(v, ex, w, t);
    Real *ob = udst.at(slot);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      v[k] = uu[k] + omega * (v[k] - uu[k]);
      ob[k * 64 + t] = v[k];
    }
    if (COMM && i < fc.nbnd) {
      // push the new boundary planes to the neighbours' owners (plane order as load_halo expects)
      void *const *fp = ((seq0 + 1) & 1) ? fc.fptr1 : fc.fptr0;
      // ex is free: the transposes are done
      if (fc.push_mode == 2)
        push_faces_tma<Real>(fc.bsend + (size_t)b * 6, fp, v, t, x, y, ex);
      else if (fc.push_mode == 1)
        push_faces_staged<Real>(fc.bsend + (size_t)b * 6, fp, v, t, x, y, ex);
      else
        push_faces<Real>(fc.bsend + (size_t)b * 6, fp, v, t, x, y);
      if (i + G >= fc.nbnd) {
        // that was this CTA's last boundary block: retire them; whoever retires the last one of
        // the whole grid publishes the new sequence number to the peers
        if (fc.push_mode == 2 && t == 0)
          push_tma_drain();  // the bulk stores of this CTA's boundary blocks have completed
        __threadfence_system();
        __syncthreads();
        if (t == 0) {
          const unsigned int mine = (unsigned int)((fc.nbnd - 1 - (int)blockIdx.x) / G + 1);
          const unsigned int old = atomicAdd(&fc.counters[0], mine);
          if (old + mine == (unsigned int)fc.nbnd) {
            __threadfence_system();
            for (int p = 0; p < fc.nsp; p++) {
              volatile unsigned long long *f = (volatile unsigned long long *)fc.peer_win[fc.speers[p]] + fc.flag_index;
              *f = seq0 + 1;
            }
          }
        }
      }
    }
    // next iteration's first ex write happens after its own __syncthreads: no extra barrier
  }
  if (COMM && t == 0) {
    // the last CTA to leave advances this rank's own sequence number and re-arms the counters
    __threadfence();
    const unsigned int old = atomicAdd(&fc.counters[1], 1u);
    if (old == (unsigned int)G - 1) {
      fc.counters[0] = 0;
      fc.counters[1] = 0;
      *fc.seq = seq0 + 1;
      __threadfence();
    }
  }
}

// ---------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)p;
  }
  return fn;
}

struct MapCache {
  std::map<std::tuple<const void *, long long, int, int>, CUtensorMap> m;
};

// kind 0: x faces, box {16 B, 8, 8}; kind 1: y faces, box {8, 1, 8}
int get_map(CupCtx *c, const void *base, long long nblocks, int kind, CUtensorMap *out) {
  if (!c->tma_cache)
    c->tma_cache = new MapCache;
  MapCache *mc = (MapCache *)c->tma_cache;
  auto key = std::make_tuple(base, nblocks, kind, c->real_bytes);
  auto it = mc->m.find(key);
  if (it != mc->m.end()) {
    *out = it->second;
    return CUP_OK;
  }
  EncodeFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return CUP_ERR_CUDA;
  }
  const cuuint64_t rb = (cuuint64_t)c->real_bytes;
  cuuint64_t dims[3] = {8, 8, (cuuint64_t)nblocks * 8};
  cuuint64_t strides[2] = {8 * rb, 64 * rb};
  cuuint32_t box[3] = {kind == 0 ? (cuuint32_t)(16 / rb) : 8u, kind == 0 ? 8u : 1u, 8u};
  // kinds 2..5: parent-level correction d: 4^3 octant, x patch {16 B,4,4}, y patch {4,1,4}, z patch {4,4,1}
  if (kind == 2) { box[0] = 4; box[1] = 4; box[2] = 4; }
  if (kind == 3) { box[0] = (cuuint32_t)(16 / rb); box[1] = 4; box[2] = 4; }
  if (kind == 4) { box[0] = 4; box[1] = 1; box[2] = 4; }
  if (kind == 5) { box[0] = 4; box[1] = 4; box[2] = 1; }
  // kinds 6, 7: three ghost layers of the 5th-order upwind stencil (k_advdiff_tma): x slab {4,8,8}
  // (the layers are columns 1..3 or 0..2 of it), y slab {8,3,8}
  if (kind == 6) { box[0] = 4; box[1] = 8; box[2] = 8; }
  if (kind == 7) { box[0] = 8; box[1] = 3; box[2] = 8; }
  cuuint32_t es[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, c->real_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: %d (base %p, blocks %lld, kind %d)", (int)r, base, nblocks, kind);
    return CUP_ERR_CUDA;
  }
  if (mc->m.size() > 4096)
    mc->m.clear();
  mc->m[key] = m;
  *out = m;
  return CUP_OK;
}

}  // namespace

int tma_slab_maps(CupCtx *c, const void *leaf, CUtensorMap out[2]) {
  // state components hold nstate = nblk + leaf-context ghost blocks
  CUP_TRY(get_map(c, leaf, c->nstate, 6, &out[0]));
  CUP_TRY(get_map(c, leaf, c->nstate, 7, &out[1]));
  return CUP_OK;
}

int tma_face_maps(CupCtx *c, const void *leaf, const void *extra, CUtensorMap out[4]) {
  // the extra part normally holds the multigrid parents and ghosts (nslot - nblk + 1 blocks); the leaf
  // context's ghost scratch of a flat vector can be larger (c->tma_extra_rows, set by its caller)
  const long long nleaf = c->nblk, nx = c->tma_extra_rows > 0 ? c->tma_extra_rows : c->nslot - c->nblk + 1;
  // a part that holds no blocks of this vector still needs a valid (unused) descriptor
  const void *lb = leaf ? leaf : extra;
  const void *eb = extra ? extra : leaf;
  CUP_TRY(get_map(c, lb, leaf ? nleaf : 1, 0, &out[0]));
  CUP_TRY(get_map(c, lb, leaf ? nleaf : 1, 1, &out[1]));
  CUP_TRY(get_map(c, eb, extra ? nx : 1, 0, &out[2]));
  CUP_TRY(get_map(c, eb, extra ? nx : 1, 1, &out[3]));
  return CUP_OK;
}

void free_tma_cache(CupCtx *c) {
  delete (MapCache *)c->tma_cache;
  c->tma_cache = nullptr;
}

template <typename Real>
int smooth_tma_launch(CupCtx *c, cudaStream_t stream, int grid, LevelView lv, const int *sub, int nsub,
                      SlotVec<Real> src, SlotVec<Real> dst, SlotVec<Real> f, Real h,
                      Real invh, Real om, const double *fmean, const FusedComm *fused, const UpFuse *upf,
                      const void *d_extra) {
  CUtensorMap mxl, myl, mxe, mye;
  CUtensorMap md[4];
  if (upf) {
    for (int k = 0; k < 4; k++)
      CUP_TRY(get_map(c, d_extra, c->nslot - c->nblk + 1, 2 + k, &md[k]));
  } else {
    memset(md, 0, sizeof md);
  }
  const long long nleaf = c->nblk, nx = c->nslot - c->nblk + 1;
  // a part that holds no blocks of this vector still needs a valid (unused) descriptor
  const void *lb = src.leaf ? (const void *)src.leaf : (const void *)src.extra;
  const void *eb = src.extra ? (const void *)src.extra : (const void *)src.leaf;
  CUP_TRY(get_map(c, lb, src.leaf ? nleaf : 1, 0, &mxl));
  CUP_TRY(get_map(c, lb, src.leaf ? nleaf : 1, 1, &myl));
  CUP_TRY(get_map(c, eb, src.extra ? nx : 1, 0, &mxe));
  CUP_TRY(get_map(c, eb, src.extra ? nx : 1, 1, &mye));
  const Real *W = (const Real *)c->d_W;
  if (upf && fused)
    k_smooth_tma<Real, true, true><<<grid, TPB, 0, stream>>>(lv, sub, nsub, *fused, *upf, src, dst, f, W, h, invh, om,
                                                              fmean, mxl, myl, mxe, mye, md[0], md[1], md[2], md[3]);
  else if (upf)
    k_smooth_tma<Real, false, true><<<grid, TPB, 0, stream>>>(lv, sub, nsub, FusedComm{}, *upf, src, dst, f, W, h,
                                                               invh, om, fmean, mxl, myl, mxe, mye, md[0], md[1],
                                                               md[2], md[3]);
  else if (fused)
    k_smooth_tma<Real, true, false><<<grid, TPB, 0, stream>>>(lv, sub, nsub, *fused, UpFuse{}, src, dst, f, W, h, invh,
                                                               om, fmean, mxl, myl, mxe, mye, md[0], md[1], md[2],
                                                               md[3]);
  else
    k_smooth_tma<Real, false, false><<<grid, TPB, 0, stream>>>(lv, sub, nsub, FusedComm{}, UpFuse{}, src, dst, f, W, h,
                                                                invh, om, fmean, mxl, myl, mxe, mye, md[0], md[1],
                                                                md[2], md[3]);
  return CUP_OK;
}

template int smooth_tma_launch<double>(CupCtx *, cudaStream_t, int, LevelView, const int *, int, SlotVec<double>,
                                       SlotVec<double>, SlotVec<double>, double, double, double, const double *,
                                       const FusedComm *, const UpFuse *, const void *);
template int smooth_tma_launch<float>(CupCtx *, cudaStream_t, int, LevelView, const int *, int, SlotVec<float>,
                                      SlotVec<float>, SlotVec<float>, float, float, float, const double *,
                                      const FusedComm *, const UpFuse *, const void *);

}  // namespace cup
