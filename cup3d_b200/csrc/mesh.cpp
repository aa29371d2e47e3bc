// mesh.cpp -- host-side topology flattening for the device kernels.
//
// The reference recomputes Hilbert indices and probes a node hash for every
// neighbour of every block on every sweep (lab_load, main.c:3579-3602, about
// 8 % of its run time).  Here the same information is resolved ONCE per mesh
// change into flat int tables that the kernels index directly:
//
//   per multigrid level L (== AMR level L, reference mg_build main.c:4522):
//     act[k]      slot of the k-th active block            (v->act,  :4538)
//     nbr[k][6]   face neighbour: slot / wall / remote face (znei+node_get)
//     pslot[k]    slot of the parent at level L-1          (v->pslot,:4628)
//     oct[k]      octant inside the parent                 (v->oct,  :4559)
//     par[]       which level-L blocks are synthesised parents (w->par, :4616)
//
// Multi-rank (one rank per GPU): every rank holds the GLOBAL leaf list (the
// reference all-gathers it in tree_sync, main.c:2928) with an owner per block,
// builds the same global hierarchy deterministically, and then localises it:
// owned blocks get local slots, neighbours owned by another rank become
// entries of a face-exchange plan (the analogue of halo_build, main.c:3030,
// but exchanging 8x8 FACES instead of whole blocks), remote parents become
// entries of restrict/prolong plans (v->x / v->xr, main.c:4621-4664; parent
// owner = smallest rank among the 8 children, :4561-4569).  Both sides of
// every exchange order their entries by (peer, position of the SENDER's block
// in the level's global order, plane), so no plan negotiation is needed.
//
// Block ORDER is whatever the caller's sta.blk[] has (the reference sorts by
// a Hilbert key; nothing here depends on it).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <unordered_map>

#include "cup_internal.h"

namespace cup {

namespace {

struct Ent {
  int level, ix, iy, iz, gslot;
};

inline uint64_t key_of(int level, int ix, int iy, int iz) {
  return ((uint64_t)level << 57) | ((uint64_t)iz << 38) | ((uint64_t)iy << 19) | (uint64_t)ix;
}

struct Xent {  // one exchange entry before sorting
  int peer, pos, plane, ref;
  bool operator<(const Xent &o) const {
    if (peer != o.peer) return peer < o.peer;
    if (pos != o.pos) return pos < o.pos;
    return plane < o.plane;
  }
};

long long coarse_threshold() {
  const char *e = getenv("CUP_COARSE_BLOCKS");
  return e ? atoll(e) : 4096;
}

void count_by_peer(const std::vector<Xent> &v, int nranks, std::vector<int> &cnt) {
  cnt.assign(nranks, 0);
  for (const Xent &e : v)
    cnt[e.peer]++;
}

}  // namespace

int build_tables(HostMesh *m, const CupBlk *gblk, long long G, const int *owner_in, int nranks, int rank,
                 const int bpd[3], int level_max) {
  if (G <= 0 || G > (1LL << 30) || level_max < 1 || level_max > 20 || bpd[0] < 1 || bpd[1] < 1 || bpd[2] < 1 ||
      nranks < 1 || rank < 0 || rank >= nranks) {
    set_error("mesh: bad arguments (blocks=%lld level_max=%d ranks=%d)", G, level_max, nranks);
    return CUP_ERR_ARG;
  }
  m->lv.clear();
  m->nranks = nranks;
  m->rank = rank;
  m->level_max = level_max;
  for (int d = 0; d < 3; d++)
    m->bpd[d] = bpd[d];
  // mg.top = sim.level_max - 1 (main.c:4526): levels above the finest leaf are
  // simply empty, exactly as in the reference.
  m->top = level_max - 1;
  m->lv.resize(m->top + 1);
  m->gblocks = G;

  std::vector<int> owner((size_t)G);  // per GLOBAL slot, grows with parents
  int lmin = 1 << 30, lmax = -1;
  std::vector<Ent> cur((size_t)G);
  m->gvol = 0;
  m->pin_local = -1;
  for (long long i = 0; i < G; i++) {
    const CupBlk &b = gblk[i];
    if (b.level < 0 || b.level > m->top) {
      set_error("block %lld: level %d outside [0,%d)", i, b.level, level_max);
      return CUP_ERR_MESH;
    }
    int nx = bpd[0] << b.level, ny = bpd[1] << b.level, nz = bpd[2] << b.level;
    if (b.ix < 0 || b.iy < 0 || b.iz < 0 || b.ix >= nx || b.iy >= ny || b.iz >= nz) {
      set_error("block %lld: index (%d,%d,%d) outside level %d", i, b.ix, b.iy, b.iz, b.level);
      return CUP_ERR_MESH;
    }
    owner[i] = owner_in ? owner_in[i] : 0;
    if (owner[i] < 0 || owner[i] >= nranks) {
      set_error("block %lld: owner %d outside [0,%d)", i, owner[i], nranks);
      return CUP_ERR_ARG;
    }
    cur[i] = {b.level, b.ix, b.iy, b.iz, (int)i};
    lmin = std::min(lmin, b.level);
    lmax = std::max(lmax, b.level);
    m->gvol += 512.0 * (b.h * b.h * b.h);
  }
  m->leaf_uniform = (lmin == lmax);
  // h of level L: the reference stores h per block = h0 / 2^L (main.c:1400)
  const double h0 = gblk[0].h * (double)(1 << gblk[0].level);
  long long gnslot = G;

  // local slots: owned leaves in global order, then owned parents in creation order
  std::vector<int> g2l((size_t)G, -1);
  m->blk.clear();
  for (long long i = 0; i < G; i++)
    if (owner[i] == rank) {
      g2l[i] = (int)m->blk.size();
      if (gblk[i].ix == 0 && gblk[i].iy == 0 && gblk[i].iz == 0)
        m->pin_local = (long long)m->blk.size();  // pois_pin, main.c:4888
      m->blk.push_back(gblk[i]);
    }
  m->nblk = (long long)m->blk.size();
  long long lnslot = m->nblk;

  // ---- leaf context: every local leaf with its six face neighbours (same level, wall,
  // one coarser leaf, or four finer leaves).  Context of pois_op and the stencil sweeps on
  // multi-level meshes (the reference's top-level mesh: lab_load :3579-3602).  Leaves of other
  // ranks that a local leaf reads become ghost blocks (slots nblk, nblk+1, ...).
  {
    Level &lf = m->leafv;
    lf = Level();
    lf.L = lmax;
    lf.h = 0;
    lf.gnact = G;
    if (!m->leaf_uniform) {
      std::unordered_map<uint64_t, int> map;
      map.reserve((size_t)G * 2);
      for (long long i = 0; i < G; i++)
        map.emplace(key_of(gblk[i].level, gblk[i].ix, gblk[i].iy, gblk[i].iz), (int)i);
      auto find = [&](int L, int x, int y, int z) -> int {
        auto it = map.find(key_of(L, x, y, z));
        return it == map.end() ? -1 : it->second;
      };
      // face f of leaf i: code (global id / NBR_WALL / NBR_COARSE / NBR_FINE) and e4 (global ids / quadrant)
      auto face_of = [&](long long i, int f, int &code, int (&e4)[4]) -> bool {
        const CupBlk &b = gblk[i];
        const int L = b.level;
        const int dim[3] = {bpd[0] << L, bpd[1] << L, bpd[2] << L};
        const int idx[3] = {b.ix, b.iy, b.iz};
        const int d = f / 2, sgn = (f & 1) ? 1 : -1, t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
        int q[3] = {idx[0], idx[1], idx[2]};
        q[d] += sgn;
        code = NBR_WALL;
        e4[0] = e4[1] = e4[2] = e4[3] = -1;
        if (q[d] < 0 || q[d] >= dim[d])
          return true;
        int id = find(L, q[0], q[1], q[2]);
        if (id >= 0) {
          code = id;
          return true;
        }
        if (L > 0 && (id = find(L - 1, q[0] / 2, q[1] / 2, q[2] / 2)) >= 0) {
          code = NBR_COARSE;
          e4[0] = id;
          e4[1] = (idx[t1] & 1) + 2 * (idx[t2] & 1);
          return true;
        }
        code = NBR_FINE;
        for (int qd = 0; qd < 4; qd++) {
          int cc[3];
          cc[d] = 2 * q[d] + (sgn > 0 ? 0 : 1);
          cc[t1] = 2 * q[t1] + (qd & 1);
          cc[t2] = 2 * q[t2] + (qd >> 1);
          e4[qd] = find(L + 1, cc[0], cc[1], cc[2]);
          if (e4[qd] < 0) {
            set_error("leaf level %d (%d,%d,%d): face %d has no neighbour at levels %d..%d (2:1 balance broken)", L,
                      b.ix, b.iy, b.iz, f, L - 1, L + 1);
            return false;
          }
        }
        return true;
      };
      // every other leaf that leaf i reads: face neighbours; for blocks at a coarse-fine interface also
      // the leaves behind the edges and corners (the wide advdiff stencil samples the coarse-level view
      // around the block, cs_sample in amr_advdiff.cu)
      std::vector<int> tmp;
      auto reads = [&](long long i, std::vector<int> &out) -> bool {
        out.clear();
        bool iface = false;
        for (int f = 0; f < 6; f++) {
          int code, e4[4];
          if (!face_of(i, f, code, e4))
            return false;
          if (code >= 0)
            out.push_back(code);
          else if (code == NBR_COARSE) {
            out.push_back(e4[0]);
            iface = true;
          } else if (code == NBR_FINE) {
            for (int z = 0; z < 4; z++)
              out.push_back(e4[z]);
            iface = true;
          }
        }
        if (!iface)
          return true;
        const CupBlk &b = gblk[i];
        const int L = b.level;
        const int dim[3] = {bpd[0] << L, bpd[1] << L, bpd[2] << L};
        for (int dz = -1; dz <= 1; dz++)
          for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
              if ((dx != 0) + (dy != 0) + (dz != 0) < 2)
                continue;  // faces are done, the block itself is not a neighbour
              const int o[3] = {dx, dy, dz}, q[3] = {b.ix + dx, b.iy + dy, b.iz + dz};
              if (q[0] < 0 || q[1] < 0 || q[2] < 0 || q[0] >= dim[0] || q[1] >= dim[1] || q[2] >= dim[2])
                continue;
              int id = find(L, q[0], q[1], q[2]);
              if (id < 0 && L > 0)
                id = find(L - 1, q[0] / 2, q[1] / 2, q[2] / 2);
              if (id >= 0) {
                out.push_back(id);
                continue;
              }
              // finer leaves that touch this block: the children of q on the side facing it
              for (int c = 0; c < 8; c++) {
                int cc[3];
                bool touch = true;
                for (int d = 0; d < 3; d++) {
                  const int bit = (c >> d) & 1;
                  if ((o[d] < 0 && bit == 0) || (o[d] > 0 && bit == 1))
                    touch = false;
                  cc[d] = 2 * q[d] + bit;
                }
                if (touch && (id = find(L + 1, cc[0], cc[1], cc[2])) >= 0)
                  out.push_back(id);
              }
            }
        return true;
      };
      // ghost blocks: need[r] = leaves of other ranks that rank r reads, ordered by (owner, global index).
      // Every rank derives every rank's list, so both sides of the exchange and the window layouts agree
      // without negotiation.
      std::vector<std::vector<int>> need((size_t)nranks);
      if (nranks > 1) {
        for (long long i = 0; i < G; i++) {
          if (!reads(i, tmp))
            return CUP_ERR_MESH;
          for (int g : tmp)
            if (owner[g] != owner[i])
              need[(size_t)owner[i]].push_back(g);
        }
        for (auto &v : need) {
          std::sort(v.begin(), v.end(), [&](int a, int b2) {
            return owner[a] != owner[b2] ? owner[a] < owner[b2] : a < b2;
          });
          v.erase(std::unique(v.begin(), v.end()), v.end());
        }
      }
      std::unordered_map<int, int> ghost;  // global leaf -> local ghost slot
      {
        const std::vector<int> &mine = need[(size_t)rank];
        ghost.reserve(mine.size() * 2);
        lf.blk_rcnt.assign(nranks, 0);
        lf.blk_scnt.assign(nranks, 0);
        for (size_t e = 0; e < mine.size(); e++) {
          ghost[mine[e]] = (int)(m->nblk + (long long)e);
          lf.blk_rslot.push_back((int)(m->nblk + (long long)e));
          lf.blk_rkind.push_back(0);
          lf.blk_rcnt[owner[mine[e]]]++;
        }
        lf.nghost = (int)mine.size();
        for (int r = 0; r < nranks; r++) {
          if (r == rank)
            continue;
          for (size_t e = 0; e < need[(size_t)r].size(); e++)
            if (owner[need[(size_t)r][e]] == rank) {
              lf.blk_sslot.push_back(g2l[need[(size_t)r][e]]);
              lf.blk_skind.push_back(0);
              lf.blk_speer.push_back(r);
              lf.blk_sidx.push_back((int)e);
              lf.blk_scnt[r]++;
            }
        }
        lf.win_nblk.assign(nranks, 0);
        for (int r = 0; r < nranks; r++)
          lf.win_nblk[r] = (int)need[(size_t)r].size();
        lf.ghosted = nranks > 1;
        lf.blk_ncomp = BLK_COMPS;
      }
      auto local_of = [&](int g) -> int { return owner[g] == rank ? g2l[g] : ghost.at(g); };
      for (long long i = 0; i < G; i++) {
        if (owner[i] != rank)
          continue;
        lf.act.push_back(g2l[i]);
        lf.hblk.push_back(gblk[i].h);
        for (int f = 0; f < 6; f++) {
          int code, e4[4];
          if (!face_of(i, f, code, e4))
            return CUP_ERR_MESH;
          if (code >= 0)
            code = local_of(code);
          else if (code == NBR_COARSE)
            e4[0] = local_of(e4[0]);
          else if (code == NBR_FINE)
            for (int z = 0; z < 4; z++)
              e4[z] = local_of(e4[z]);
          lf.nbr.push_back(code);
          for (int z = 0; z < 4; z++)
            lf.ext.push_back(e4[z]);
        }
      }
      lf.uniform = false;
      // device hash of the leaves this rank can read (own + ghosts): key = key_of(...) + 1 (0 = empty bucket)
      size_t cap = 64;
      while (cap < (size_t)(m->nblk + lf.nghost) * 2)
        cap *= 2;
      lf.hkeys.assign(cap, 0ULL);
      lf.hvals.assign(cap, -1);
      auto hput = [&](int g, int slot) {
        const unsigned long long k = key_of(gblk[g].level, gblk[g].ix, gblk[g].iy, gblk[g].iz) + 1;
        size_t h = (size_t)((k * 0x9E3779B97F4A7C15ULL) >> 20) & (cap - 1);
        while (lf.hkeys[h] != 0)
          h = (h + 1) & (cap - 1);
        lf.hkeys[h] = k;
        lf.hvals[h] = slot;
      };
      for (long long i = 0; i < G; i++) {
        if (owner[i] != rank)
          continue;
        hput((int)i, g2l[i]);
        lf.bijk.push_back(gblk[i].level);
        lf.bijk.push_back(gblk[i].ix);
        lf.bijk.push_back(gblk[i].iy);
        lf.bijk.push_back(gblk[i].iz);
      }
      for (int g : need[(size_t)rank])
        hput(g, ghost.at(g));
    } else {
      // single-level mesh: only the hash and the block coordinates (the tensorial labs of the adaptation
      // kernels look their 26 neighbours up; sweeps use the level's own tables)
      size_t cap = 64;
      while (cap < (size_t)m->nblk * 2)
        cap *= 2;
      lf.hkeys.assign(cap, 0ULL);
      lf.hvals.assign(cap, -1);
      for (long long i = 0; i < G; i++) {
        if (owner[i] != rank)
          continue;
        const unsigned long long k = key_of(gblk[i].level, gblk[i].ix, gblk[i].iy, gblk[i].iz) + 1;
        size_t h = (size_t)((k * 0x9E3779B97F4A7C15ULL) >> 20) & (cap - 1);
        while (lf.hkeys[h] != 0)
          h = (h + 1) & (cap - 1);
        lf.hkeys[h] = k;
        lf.hvals[h] = g2l[i];
        lf.bijk.push_back(gblk[i].level);
        lf.bijk.push_back(gblk[i].ix);
        lf.bijk.push_back(gblk[i].iy);
        lf.bijk.push_back(gblk[i].iz);
      }
    }
  }

  // ---- global hierarchy, finest to coarsest ------------------------------
  struct GLevel {
    std::vector<Ent> act;          // global active list (global order)
    std::vector<int> nbr;          // [n][6] global slot / NBR_WALL / NBR_COARSE / NBR_FINE
    std::vector<int> ext;          // [n][6][4] global slots of the coarse / 4 fine neighbours
    std::vector<int> pg;           // parent global slot
  };
  std::vector<GLevel> gl(m->top + 1);
  for (int L = m->top; L >= 0; L--) {
    GLevel &g = gl[L];
    std::unordered_map<uint64_t, int> map;
    map.reserve(cur.size() * 2);
    for (const Ent &e : cur)
      if (!map.emplace(key_of(e.level, e.ix, e.iy, e.iz), e.gslot).second) {
        set_error("duplicate block level %d (%d,%d,%d)", e.level, e.ix, e.iy, e.iz);
        return CUP_ERR_MESH;
      }
    for (const Ent &e : cur)
      if (e.level == L)
        g.act.push_back(e);
    const size_t na = g.act.size();
    g.nbr.assign(na * 6, NBR_WALL);
    g.ext.assign(na * 24, -1);
    const int dim[3] = {bpd[0] << L, bpd[1] << L, bpd[2] << L};
    for (size_t k = 0; k < na; k++) {
      const Ent &e = g.act[k];
      const int idx[3] = {e.ix, e.iy, e.iz};
      for (int f = 0; f < 6; f++) {
        int d = f / 2, s = (f & 1) ? 1 : -1;
        int q[3] = {idx[0], idx[1], idx[2]};
        q[d] += s;
        if (q[d] < 0 || q[d] >= dim[d])
          continue;  // domain wall (nei_outside, main.c:2837)
        auto it = map.find(key_of(L, q[0], q[1], q[2]));
        if (it != map.end()) {
          g.nbr[k * 6 + f] = it->second;
        } else {
          // must be covered by a coarser leaf (2:1 balance): check its presence
          auto ic = L == 0 ? map.end() : map.find(key_of(L - 1, q[0] / 2, q[1] / 2, q[2] / 2));
          if (ic == map.end()) {
            set_error("level %d block (%d,%d,%d): face %d neighbour missing", L, e.ix, e.iy, e.iz, f);
            return CUP_ERR_MESH;
          }
          g.nbr[k * 6 + f] = NBR_COARSE;
          // which quadrant of the coarse block's face this block touches: parity of its own
          // index in the two tangential directions (lower dimension first)
          const int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
          g.ext[(k * 6 + f) * 4 + 0] = ic->second;
          g.ext[(k * 6 + f) * 4 + 1] = (idx[t1] & 1) + 2 * (idx[t2] & 1);
        }
      }
    }
    if (L == 0)
      break;
    // parents: one new global slot per distinct parent, first-appearance order;
    // owner = smallest rank among the children (mg_build, main.c:4561-4569)
    // (their level is L - 1: glevel below)
    std::unordered_map<uint64_t, int> pmap;
    std::vector<Ent> next;
    next.reserve(cur.size() - na + na / 8 + 1);
    for (const Ent &e : cur)
      if (e.level < L)
        next.push_back(e);
    g.pg.resize(na);
    std::vector<int> nchild;
    const long long first = gnslot;
    for (size_t k = 0; k < na; k++) {
      const Ent &e = g.act[k];
      const uint64_t pk = key_of(L - 1, e.ix / 2, e.iy / 2, e.iz / 2);
      auto it = pmap.find(pk);
      int ps;
      if (it == pmap.end()) {
        ps = (int)gnslot++;
        pmap.emplace(pk, ps);
        next.push_back({L - 1, e.ix / 2, e.iy / 2, e.iz / 2, ps});
        nchild.push_back(0);
        owner.push_back(owner[e.gslot]);
      } else {
        ps = it->second;
        owner[ps] = std::min(owner[ps], owner[e.gslot]);
      }
      nchild[(size_t)(ps - first)]++;
      g.pg[k] = ps;
    }
    // Small coarse levels are latency-bound and would cost one exchange per sweep
    // for a handful of blocks per rank: below a threshold the whole level lives on
    // rank 0 (one restriction/prolongation hop instead of ~7 halo exchanges per level).
    {
      long long nnext = 0;
      for (const Ent &e : next)
        nnext += (e.level == L - 1);
      if (nranks > 1 && nnext <= coarse_threshold())
        for (long long s2 = first; s2 < gnslot; s2++)
          owner[s2] = 0;
    }
    for (size_t p = 0; p < nchild.size(); p++)
      if (nchild[p] != 8) {
        // mg_build: "sibling of level %d missing" (main.c:4566)
        set_error("level %d: a parent has %d of 8 children (siblings missing)", L, nchild[p]);
        return CUP_ERR_MESH;
      }
    cur.swap(next);
  }
  // local slots of owned parents, in global-slot order
  g2l.resize((size_t)gnslot, -1);
  for (long long s = G; s < gnslot; s++)
    if (owner[s] == rank)
      g2l[s] = (int)lnslot++;
  m->nown = lnslot;
  m->gnslot = gnslot;
  // level of every global slot, and which levels have coarse-fine interfaces ANYWHERE: on those, blocks of
  // other ranks are read as ghost blocks (slots >= nown, one per remote block whatever the level that
  // reads it); on uniform levels 8x8 faces travel (plans below)
  std::vector<int> glevel((size_t)gnslot, 0);
  std::vector<char> ghosted_level((size_t)m->top + 1, 0);
  for (int L = 0; L <= m->top; L++) {
    for (const Ent &e : gl[L].act)
      glevel[(size_t)e.gslot] = L;
    if (nranks > 1)
      for (int code : gl[L].nbr)
        if (code == NBR_COARSE) {
          ghosted_level[(size_t)L] = 1;
          break;
        }
  }
  std::unordered_map<int, int> mg_ghost;  // global slot -> local ghost slot of this rank
  auto ghost_slot = [&](int gs) -> int {
    auto it = mg_ghost.find(gs);
    if (it != mg_ghost.end())
      return it->second;
    const int ls = (int)lnslot++;
    mg_ghost.emplace(gs, ls);
    return ls;
  };

  // ---- localise ----------------------------------------------------------
  for (int L = 0; L <= m->top; L++) {
    Level &v = m->lv[L];
    const GLevel &g = gl[L];
    v = Level();
    v.L = L;
    v.h = h0 / (double)(1 << L);
    v.gnact = (long long)g.act.size();
    std::vector<Xent> rx, sx;  // face recv / send
    // traffic matrices of this level over ALL ranks: fmat[s][d] faces, rmat[s][d] restricted children
    std::vector<long long> fmat((size_t)nranks * nranks, 0), rmat((size_t)nranks * nranks, 0);
    const bool ghosted = nranks > 1 && ghosted_level[(size_t)L];
    std::vector<std::vector<int>> need((size_t)nranks);  // ghosted level: global slots rank r reads from others
    if (nranks > 1)
      for (size_t k = 0; k < g.act.size(); k++) {
        const int os = owner[g.act[k].gslot];
        for (int f = 0; f < 6; f++) {
          const int gn = g.nbr[k * 6 + f];
          if (ghosted) {
            const int src = gn >= 0 ? gn : (gn == NBR_COARSE ? g.ext[(k * 6 + f) * 4] : -1);
            if (src >= 0 && owner[src] != os)
              need[(size_t)os].push_back(src);
          } else if (gn >= 0 && owner[gn] != os)
            fmat[(size_t)os * nranks + owner[gn]]++;  // os sends its plane f to the neighbour's owner
        }
        if (L >= 1 && owner[g.pg[k]] != os)
          rmat[(size_t)os * nranks + owner[g.pg[k]]]++;
      }
    if (ghosted) {
      for (auto &nv : need) {
        std::sort(nv.begin(), nv.end(), [&](int a, int b2) {
          return owner[a] != owner[b2] ? owner[a] < owner[b2] : a < b2;
        });
        nv.erase(std::unique(nv.begin(), nv.end()), nv.end());
      }
      v.ghosted = true;
      v.uniform = false;  // the sweeps of this level go through the interface path, which refreshes the ghost blocks
      v.blk_ncomp = 1;
      v.blk_rcnt.assign(nranks, 0);
      v.blk_scnt.assign(nranks, 0);
      // kind 0: a block of this level (read from the vector being swept); kind 1: a coarser leaf
      // behind an interface (read from the canonical U0)
      for (int gs : need[(size_t)rank]) {
        v.blk_rslot.push_back(ghost_slot(gs));
        v.blk_rkind.push_back(glevel[(size_t)gs] == L ? 0 : 1);
        v.blk_rcnt[owner[gs]]++;
      }
      v.nghost = (int)need[(size_t)rank].size();
      for (int r = 0; r < nranks; r++) {
        if (r == rank)
          continue;
        for (size_t e = 0; e < need[(size_t)r].size(); e++) {
          const int gs = need[(size_t)r][e];
          if (owner[gs] != rank)
            continue;
          v.blk_sslot.push_back(g2l[gs]);
          v.blk_skind.push_back(glevel[(size_t)gs] == L ? 0 : 1);
          v.blk_speer.push_back(r);
          v.blk_sidx.push_back((int)e);
          v.blk_scnt[r]++;
        }
      }
      v.win_nblk.assign(nranks, 0);
      for (int r = 0; r < nranks; r++)
        v.win_nblk[r] = (int)need[(size_t)r].size();
    }
    for (size_t k = 0; k < g.act.size(); k++) {
      const int gs = g.act[k].gslot;
      if (owner[gs] != rank)
        continue;
      const int lk = (int)v.act.size();
      v.act.push_back(g2l[gs]);
      v.ijk.push_back(g.act[k].ix);
      v.ijk.push_back(g.act[k].iy);
      v.ijk.push_back(g.act[k].iz);
      for (int f = 0; f < 6; f++) {
        const int gn = g.nbr[k * 6 + f];
        int code;
        if (gn == NBR_WALL || gn == NBR_COARSE) {
          code = gn;
          if (gn == NBR_COARSE) {
            v.uniform = false;
            const int gc = g.ext[(k * 6 + f) * 4];
            if (v.ext.size() < v.nbr.size() * 4 + 4)
              v.ext.resize((size_t)(v.nbr.size() + 1) * 4, -1);
            v.ext[v.nbr.size() * 4 + 0] = owner[gc] == rank ? g2l[gc] : mg_ghost.at(gc);
            v.ext[v.nbr.size() * 4 + 1] = g.ext[(k * 6 + f) * 4 + 1];
          }
        } else if (owner[gn] == rank) {
          code = g2l[gn];
        } else if (ghosted) {
          code = mg_ghost.at(gn);  // a ghost block: addressed like a local one
        } else {
          code = 0;  // patched after sorting
          rx.push_back({owner[gn], -1, f ^ 1, lk * 6 + f});
          rx.back().pos = gn;  // provisional: global slot; replaced by level position below
          sx.push_back({owner[gn], (int)k, f, g2l[gs]});
        }
        v.nbr.push_back(code);
      }
    }
    // position of a global slot inside this level's global order
    if (!rx.empty()) {
      std::unordered_map<int, int> pos;
      pos.reserve(g.act.size() * 2);
      for (size_t k = 0; k < g.act.size(); k++)
        pos[g.act[k].gslot] = (int)k;
      for (Xent &e : rx)
        e.pos = pos[e.pos];
    }
    std::sort(rx.begin(), rx.end());
    std::sort(sx.begin(), sx.end());
    count_by_peer(rx, nranks, v.face_rcnt);
    count_by_peer(sx, nranks, v.face_scnt);
    for (size_t i = 0; i < rx.size(); i++)
      v.nbr[(size_t)rx[i].ref] = NBR_REMOTE0 - (int)i;
    {
      int prev = -1, j = 0;
      for (const Xent &e : sx) {
        v.face_sslot.push_back(e.ref);
        v.face_splane.push_back(e.plane);
        if (e.peer != prev) {
          prev = e.peer;
          j = 0;
        }
        long long base = 0;  // entries of lower ranks come first in the destination's face area
        for (int q = 0; q < rank; q++)
          base += fmat[(size_t)q * nranks + e.peer];
        v.face_speer.push_back(e.peer);
        v.face_sidx.push_back((int)(base + j++));
      }
    }
    v.nface_recv = (int)rx.size();
    // restrict / prolong plans (L >= 1)
    if (L >= 1) {
      std::vector<Xent> rs, rr;
      int lk = 0;
      for (size_t k = 0; k < g.act.size(); k++) {
        const int gs = g.act[k].gslot, gp = g.pg[k];
        const Ent &e = g.act[k];
        const int oct = (e.ix & 1) + 2 * (e.iy & 1) + 4 * (e.iz & 1);
        if (owner[gs] == rank) {
          v.oct.push_back(oct);
          if (owner[gp] == rank) {
            v.pslot.push_back(g2l[gp]);
          } else {
            v.pslot.push_back(0);
            rs.push_back({owner[gp], (int)k, 0, lk});
          }
          lk++;
        } else if (owner[gp] == rank) {
          rr.push_back({owner[gs], (int)k, oct, g2l[gp]});
        }
      }
      std::sort(rs.begin(), rs.end());
      std::sort(rr.begin(), rr.end());
      count_by_peer(rs, nranks, v.res_scnt);
      count_by_peer(rr, nranks, v.res_rcnt);
      {
        int prev = -1, j = 0;
        for (size_t i = 0; i < rs.size(); i++) {
          v.pslot[(size_t)rs[i].ref] = NBR_REMOTE0 - (int)i;
          if (rs[i].peer != prev) {
            prev = rs[i].peer;
            j = 0;
          }
          long long base = 0;
          for (int q = 0; q < rank; q++)
            base += rmat[(size_t)q * nranks + rs[i].peer];
          v.res_speer.push_back(rs[i].peer);
          v.res_sidx.push_back((int)(base + j++));
        }
      }
      {
        int prev = -1, j = 0;
        for (const Xent &e : rr) {
          v.res_rslot.push_back(e.ref);
          v.res_roct.push_back(e.plane);
          if (e.peer != prev) {
            prev = e.peer;
            j = 0;
          }
          // the child's owner s keeps its remote-parent children ordered by (parent owner, position):
          // ours start after those it sends to lower ranks
          long long base = 0;
          for (int d = 0; d < rank; d++)
            base += rmat[(size_t)e.peer * nranks + d];
          v.pro_speer.push_back(e.peer);
          v.pro_sidx.push_back((int)(base + j++));
        }
      }
    } else {
      v.res_scnt.assign(nranks, 0);
      v.res_rcnt.assign(nranks, 0);
    }
    if (!v.uniform)
      v.ext.resize(v.nbr.size() * 4, -1);
    // par[]: indices (into act[]) of synthesised parents, i.e. local slot >= nblk
    for (size_t k = 0; k < v.act.size(); k++)
      if (v.act[k] >= m->nblk)
        v.par.push_back((int)k);
    // receive-window bookkeeping of every rank for this level (sizes in entries)
    v.win_nrecv.assign(nranks, 0);
    v.win_face.assign(nranks, 0);  // reused below as: res entries, pro entries (temporarily)
    v.win_res.assign(nranks, 0);
    v.win_pro.assign(nranks, 0);
    for (int p = 0; p < nranks; p++)
      for (int q = 0; q < nranks; q++) {
        v.win_nrecv[p] += (int)fmat[(size_t)q * nranks + p];
        v.win_res[p] += rmat[(size_t)q * nranks + p];   // children received by p
        v.win_pro[p] += rmat[(size_t)p * nranks + q];   // corrections received by p
      }
    // interior / boundary split for comm-compute overlap
    for (size_t k = 0; k < v.act.size(); k++) {
      bool rem = false;
      for (int f = 0; f < 6; f++)
        rem |= v.nbr[k * 6 + f] <= NBR_REMOTE0;
      (rem ? v.bnd : v.inner).push_back((int)k);
    }
  }
  m->nslot = lnslot;  // own slots + ghost blocks
  // lay out every rank's receive window: per level [faces parity 0][faces parity 1][restrict][prolong][ghost blocks x2]
  m->win_reals.assign(nranks, 0);
  int finest = m->top;
  while (finest > 0 && m->lv[finest].gnact == 0)
    finest--;
  for (int L = 0; L <= m->top; L++) {
    m->lv[L].win_slab.assign(nranks, -1);
    m->lv[L].win_blk.assign(nranks, -1);
    if (m->lv[L].win_nblk.empty())
      m->lv[L].win_nblk.assign(nranks, 0);
  }
  m->leafv.win_blk.assign(nranks, -1);
  if (m->leafv.win_nblk.empty())
    m->leafv.win_nblk.assign(nranks, 0);
  for (int p = 0; p < nranks; p++) {
    long long off = 0;
    for (int L = 0; L <= m->top; L++) {
      Level &v = m->lv[L];
      const long long nres = v.win_res[p], npro = v.win_pro[p];
      v.win_face[p] = off;
      off += 2LL * v.win_nrecv[p] * 64;
      v.win_res[p] = off;
      off += nres * 128;
      v.win_pro[p] = off;
      off += npro * 64;
      if (L == finest) {  // ghost slabs of the stencil sweeps live on the leaf level
        v.win_slab[p] = off;
        off += 2LL * v.win_nrecv[p] * 64 * SLAB_PLANES;
      }
      if (v.ghosted) {
        v.win_blk[p] = off;
        off += 2LL * v.win_nblk[p] * 512 * v.blk_ncomp;
      }
    }
    if (m->leafv.ghosted) {
      m->leafv.win_blk[p] = off;
      off += 2LL * m->leafv.win_nblk[p] * 512 * m->leafv.blk_ncomp;
    }
    m->win_reals[p] = off;
  }
  return CUP_OK;
}

namespace {
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
    cudaFree(v.d_ext);
    cudaFree(v.d_hblk);
    cudaFree(v.d_upinfo);
    cudaFree(v.d_reg);
    cudaFree(v.d_irr);
    cudaFree(v.d_par_reg);
    cudaFree(v.d_par_irr);
    cudaFree(v.d_inner);
    cudaFree(v.d_bnd);
    cudaFree(v.d_face_sslot);
    cudaFree(v.d_face_splane);
    cudaFree(v.d_res_rslot);
    cudaFree(v.d_res_roct);
    cudaFree(v.d_blk_sslot);
    cudaFree(v.d_blk_skind);
    cudaFree(v.d_blk_rslot);
    cudaFree(v.d_blk_rkind);
  }
  c->lv.clear();
  {
    Level &v = c->leafv;
    cudaFree(v.d_act);
    cudaFree(v.d_nbr);
    cudaFree(v.d_ext);
    cudaFree(v.d_hblk);
    cudaFree(v.d_reg);
    cudaFree(v.d_irr);
    cudaFree(v.d_bijk);
    cudaFree(v.d_hkeys);
    cudaFree(v.d_hvals);
    for (int *p : v.d_reg_by_level)
      cudaFree(p);
    cudaFree(v.d_blk_sslot);
    cudaFree(v.d_blk_skind);
    cudaFree(v.d_blk_rslot);
    cudaFree(v.d_blk_rkind);
    v = Level();
  }
  cudaFree(c->d_list);
  c->d_list = nullptr;
  c->run_sub = nullptr;
  c->run_nsub = -1;
  c->blk.clear();
  c->nblk = c->nslot = 0;
  c->top = -1;
}

int build_mesh(CupCtx *c, const CupBlk *gblk, long long G, const int *owner, const int bpd[3], int level_max) {
  free_mesh(c);
  HostMesh m;
  CUP_TRY(build_tables(&m, gblk, G, owner, c->nranks, c->rank, bpd, level_max));
  if (m.nblk == 0) {
    set_error("rank %d owns no blocks", c->rank);
    return CUP_ERR_ARG;
  }
  c->blk.swap(m.blk);
  c->lv.swap(m.lv);
  c->leafv = m.leafv;
  c->win_reals.swap(m.win_reals);
  c->nblk = m.nblk;
  c->nslot = m.nslot;
  c->nstate = m.nblk + m.leafv.nghost;
  c->gblocks = m.gblocks;
  c->gvol = m.gvol;
  c->pin_local = m.pin_local;
  c->top = m.top;
  c->level_max = level_max;
  c->leaf_uniform = m.leaf_uniform;
  for (int d = 0; d < 3; d++)
    c->bpd[d] = bpd[d];
  for (auto &v : c->lv) {
    CUP_TRY(upload(&v.d_act, v.act));
    CUP_TRY(upload(&v.d_nbr, v.nbr));
    CUP_TRY(upload(&v.d_pslot, v.pslot));
    CUP_TRY(upload(&v.d_oct, v.oct));
    CUP_TRY(upload(&v.d_par, v.par));
    CUP_TRY(upload(&v.d_ext, v.ext));
    // fused prolongation needs every parent local and every neighbour at the same level
    v.upinfo.clear();
    if (v.L >= 1 && v.uniform && !v.act.empty()) {
      bool ok = true;
      std::unordered_map<int, int> idx_of;
      idx_of.reserve(v.act.size() * 2);
      for (size_t k = 0; k < v.act.size(); k++) {
        idx_of[v.act[k]] = (int)k;
        ok = ok && v.pslot[k] >= (int)c->nblk;
      }
      if (ok) {
        v.upinfo.resize(v.act.size() * 7);
        for (size_t k = 0; k < v.act.size(); k++) {
          auto pack = [&](size_t j) { return ((v.pslot[j] - (int)c->nblk) << 3) | v.oct[j]; };
          v.upinfo[k * 7] = pack(k);
          for (int f = 0; f < 6; f++) {
            const int nb = v.nbr[k * 6 + f];
            v.upinfo[k * 7 + 1 + f] = nb >= 0 ? pack((size_t)idx_of.at(nb)) : (nb == NBR_WALL ? pack(k) : -1);
          }
        }
      }
    }
    CUP_TRY(upload(&v.d_upinfo, v.upinfo));
    v.reg.clear();
    v.irr.clear();
    v.par_reg.clear();
    v.par_irr.clear();
    if (!v.uniform) {
      std::vector<char> is_irr(v.act.size(), 0);
      for (size_t k = 0; k < v.act.size(); k++) {
        for (int f = 0; f < 6; f++)
          is_irr[k] |= (v.nbr[k * 6 + f] == NBR_COARSE || v.nbr[k * 6 + f] == NBR_FINE);
        (is_irr[k] ? v.irr : v.reg).push_back((int)k);
      }
      for (int k : v.par)
        (is_irr[(size_t)k] ? v.par_irr : v.par_reg).push_back(k);
    }
    CUP_TRY(upload(&v.d_reg, v.reg));
    CUP_TRY(upload(&v.d_irr, v.irr));
    CUP_TRY(upload(&v.d_par_reg, v.par_reg));
    CUP_TRY(upload(&v.d_par_irr, v.par_irr));
    CUP_TRY(upload(&v.d_inner, v.inner));
    CUP_TRY(upload(&v.d_bnd, v.bnd));
    CUP_TRY(upload(&v.d_face_sslot, v.face_sslot));
    CUP_TRY(upload(&v.d_face_splane, v.face_splane));
    CUP_TRY(upload(&v.d_res_rslot, v.res_rslot));
    CUP_TRY(upload(&v.d_res_roct, v.res_roct));
    CUP_TRY(upload(&v.d_blk_sslot, v.blk_sslot));
    CUP_TRY(upload(&v.d_blk_skind, v.blk_skind));
    CUP_TRY(upload(&v.d_blk_rslot, v.blk_rslot));
    CUP_TRY(upload(&v.d_blk_rkind, v.blk_rkind));
  }
  {
    Level &v = c->leafv;
    CUP_TRY(upload(&v.d_act, v.act));
    CUP_TRY(upload(&v.d_nbr, v.nbr));
    CUP_TRY(upload(&v.d_ext, v.ext));
    v.reg.clear();
    v.irr.clear();
    for (size_t k = 0; k < v.act.size(); k++) {
      bool ir = false;
      for (int f = 0; f < 6; f++)
        ir |= (v.nbr[k * 6 + f] == NBR_COARSE || v.nbr[k * 6 + f] == NBR_FINE);
      (ir ? v.irr : v.reg).push_back((int)k);
    }
    CUP_TRY(upload(&v.d_reg, v.reg));
    CUP_TRY(upload(&v.d_irr, v.irr));
    v.reg_by_level.assign((size_t)c->top + 1, std::vector<int>());
    for (int k : v.reg)
      v.reg_by_level[(size_t)c->blk[(size_t)v.act[(size_t)k]].level].push_back(k);
    v.d_reg_by_level.assign((size_t)c->top + 1, nullptr);
    for (int L = 0; L <= c->top; L++)
      CUP_TRY(upload(&v.d_reg_by_level[(size_t)L], v.reg_by_level[(size_t)L]));
    CUP_TRY(upload(&v.d_blk_sslot, v.blk_sslot));
    CUP_TRY(upload(&v.d_blk_skind, v.blk_skind));
    CUP_TRY(upload(&v.d_blk_rslot, v.blk_rslot));
    CUP_TRY(upload(&v.d_blk_rkind, v.blk_rkind));
    CUP_TRY(upload(&v.d_bijk, v.bijk));
    CUP_TRY(upload(&v.d_hkeys, v.hkeys));
    CUP_TRY(upload(&v.d_hvals, v.hvals));
    if (!v.hblk.empty()) {
      if (c->real_bytes == 8) {
        CUP_TRY(upload((double **)&v.d_hblk, v.hblk));
      } else {
        std::vector<float> hf(v.hblk.begin(), v.hblk.end());
        CUP_TRY(upload((float **)&v.d_hblk, hf));
      }
    }
  }
  return CUP_OK;
}

}  // namespace cup
