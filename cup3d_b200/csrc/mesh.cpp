// mesh.cpp -- host-side topology flattening for the device kernels.
//
// The reference recomputes Hilbert indices and probes a node hash for every
// neighbour of every block on every sweep (lab_load, main.c:3579-3602, about
// 8 % of its run time).  Here the same information is resolved ONCE per mesh
// change into flat int tables that the kernels index directly:
//
//   per multigrid level L (== AMR level L, reference mg_build main.c:4522):
//     act[k]      slot of the k-th active block            (v->act,  :4538)
//     nbr[k][6]   slot of each face neighbour / wall code  (znei+node_get)
//     pslot[k]    slot of the parent at level L-1          (v->pslot,:4628)
//     oct[k]      octant inside the parent                 (v->oct,  :4559)
//     par[]       which level-(L) blocks are synthesised parents (w->par, :4616)
//
// Slots follow the reference: leaves keep their block index, synthesised
// parents are appended (mg.nslot++, main.c:4592).  Block ORDER is whatever the
// caller's sta.blk[] has (the reference sorts by a Hilbert key; nothing here
// depends on it, the tables are built from (level, ix, iy, iz) alone).
#include <algorithm>
#include <cstdint>
#include <unordered_map>

#include "cup_internal.h"

namespace cup {

namespace {

struct Ent {
  int level, ix, iy, iz, slot;
};

inline uint64_t key_of(int level, int ix, int iy, int iz) {
  return ((uint64_t)level << 57) | ((uint64_t)iz << 38) | ((uint64_t)iy << 19) | (uint64_t)ix;
}

template <typename T>
int upload(T **d, const std::vector<T> &h) {
  *d = nullptr;
  if (h.empty())
    return CUP_OK;
  CUP_CUDA(cudaMalloc((void **)d, h.size() * sizeof(T)));
  CUP_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return CUP_OK;
}

}  // namespace

void free_mesh(CupCtx *c) {
  for (auto &v : c->lv) {
    cudaFree(v.d_act);
    cudaFree(v.d_nbr);
    cudaFree(v.d_pslot);
    cudaFree(v.d_oct);
    cudaFree(v.d_par);
  }
  c->lv.clear();
  c->blk.clear();
  c->nblk = c->nslot = 0;
  c->top = -1;
}

int build_mesh(CupCtx *c, const CupBlk *blk, long long n, const int bpd[3], int level_max) {
  if (n <= 0 || n > (1LL << 30) || level_max < 1 || level_max > 20 || bpd[0] < 1 || bpd[1] < 1 || bpd[2] < 1) {
    set_error("cup_mesh_upload: bad arguments (n=%lld level_max=%d)", n, level_max);
    return CUP_ERR_ARG;
  }
  free_mesh(c);
  c->blk.assign(blk, blk + n);
  c->nblk = n;
  c->level_max = level_max;
  for (int d = 0; d < 3; d++)
    c->bpd[d] = bpd[d];
  // mg.top = sim.level_max - 1 (main.c:4526): levels above the finest leaf are
  // simply empty, exactly as in the reference.
  c->top = level_max - 1;
  c->lv.resize(c->top + 1);

  int lmin = 1 << 30, lmax = -1;
  std::vector<Ent> cur((size_t)n);
  for (long long i = 0; i < n; i++) {
    const CupBlk &b = blk[i];
    if (b.level < 0 || b.level > c->top) {
      set_error("block %lld: level %d outside [0,%d)", i, b.level, level_max);
      return CUP_ERR_MESH;
    }
    int nx = bpd[0] << b.level, ny = bpd[1] << b.level, nz = bpd[2] << b.level;
    if (b.ix < 0 || b.iy < 0 || b.iz < 0 || b.ix >= nx || b.iy >= ny || b.iz >= nz) {
      set_error("block %lld: index (%d,%d,%d) outside level %d", i, b.ix, b.iy, b.iz, b.level);
      return CUP_ERR_MESH;
    }
    cur[i] = {b.level, b.ix, b.iy, b.iz, (int)i};
    lmin = std::min(lmin, b.level);
    lmax = std::max(lmax, b.level);
  }
  c->leaf_uniform = (lmin == lmax);
  // h of level L: the reference stores h per block = h0 / 2^L (main.c:1400)
  double h0 = blk[0].h * (double)(1 << blk[0].level);
  long long nslot = n;

  for (int L = c->top; L >= 0; L--) {
    Level &v = c->lv[L];
    v.L = L;
    v.h = h0 / (double)(1 << L);
    std::unordered_map<uint64_t, int> map;
    map.reserve(cur.size() * 2);
    for (const Ent &e : cur) {
      if (!map.emplace(key_of(e.level, e.ix, e.iy, e.iz), e.slot).second) {
        set_error("duplicate block level %d (%d,%d,%d)", e.level, e.ix, e.iy, e.iz);
        return CUP_ERR_MESH;
      }
    }
    std::vector<const Ent *> active;
    for (const Ent &e : cur)
      if (e.level == L)
        active.push_back(&e);
    size_t na = active.size();
    v.act.resize(na);
    v.nbr.assign(na * 6, NBR_WALL);
    const int dim[3] = {bpd[0] << L, bpd[1] << L, bpd[2] << L};
    for (size_t k = 0; k < na; k++) {
      const Ent &e = *active[k];
      v.act[k] = e.slot;
      const int idx[3] = {e.ix, e.iy, e.iz};
      for (int f = 0; f < 6; f++) {
        int d = f / 2, s = (f & 1) ? 1 : -1;
        int q[3] = {idx[0], idx[1], idx[2]};
        q[d] += s;
        if (q[d] < 0 || q[d] >= dim[d])
          continue;  // domain wall (nei_outside, main.c:2837)
        auto it = map.find(key_of(L, q[0], q[1], q[2]));
        if (it != map.end()) {
          v.nbr[k * 6 + f] = it->second;
        } else {
          // must be covered by a coarser leaf (2:1 balance): check its presence
          if (L == 0 || map.find(key_of(L - 1, q[0] / 2, q[1] / 2, q[2] / 2)) == map.end()) {
            set_error("level %d block (%d,%d,%d): face %d neighbour missing", L, e.ix, e.iy, e.iz, f);
            return CUP_ERR_MESH;
          }
          v.nbr[k * 6 + f] = NBR_COARSE;
          v.uniform = false;
        }
      }
    }
    if (L == 0)
      break;
    // parents: one new slot per distinct parent, first-appearance order
    std::unordered_map<uint64_t, int> pmap;
    std::vector<Ent> next;
    next.reserve(cur.size() - na + na / 8 + 1);
    for (const Ent &e : cur)
      if (e.level < L)
        next.push_back(e);
    size_t first_parent = next.size();
    v.pslot.resize(na);
    v.oct.resize(na);
    std::vector<int> nchild;
    for (size_t k = 0; k < na; k++) {
      const Ent &e = *active[k];
      uint64_t pk = key_of(L - 1, e.ix / 2, e.iy / 2, e.iz / 2);
      auto it = pmap.find(pk);
      int ps;
      if (it == pmap.end()) {
        ps = (int)nslot++;
        pmap.emplace(pk, ps);
        next.push_back({L - 1, e.ix / 2, e.iy / 2, e.iz / 2, ps});
        nchild.push_back(0);
      } else {
        ps = it->second;
      }
      nchild[(size_t)(ps - (nslot - (long long)nchild.size()))]++;
      v.pslot[k] = ps;
      v.oct[k] = (e.ix & 1) + 2 * (e.iy & 1) + 4 * (e.iz & 1);
    }
    for (size_t p = 0; p < nchild.size(); p++)
      if (nchild[p] != 8) {
        // mg_build: "sibling of level %d missing" (main.c:4566)
        set_error("level %d: a parent has %d of 8 children (siblings missing)", L, nchild[p]);
        return CUP_ERR_MESH;
      }
    // which active blocks of level L-1 are those parents is resolved below,
    // once level L-1's active list exists
    cur.swap(next);
    (void)first_parent;
  }
  // par[]: indices (into act[]) of synthesised parents, i.e. slot >= nblk
  for (int L = 0; L <= c->top; L++) {
    Level &v = c->lv[L];
    for (size_t k = 0; k < v.act.size(); k++)
      if (v.act[k] >= n)
        v.par.push_back((int)k);
  }
  c->nslot = nslot;
  for (auto &v : c->lv) {
    CUP_TRY(upload(&v.d_act, v.act));
    CUP_TRY(upload(&v.d_nbr, v.nbr));
    CUP_TRY(upload(&v.d_pslot, v.pslot));
    CUP_TRY(upload(&v.d_oct, v.oct));
    CUP_TRY(upload(&v.d_par, v.par));
  }
  return CUP_OK;
}

}  // namespace cup
