// cup_internal.h -- shared host-side declarations of the B200 CUP3D hot path.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cup3d_b200.h"

namespace cup {

void set_error(const char *fmt, ...);

#define CUP_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) {                                                                        \
      cup::set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));           \
      return CUP_ERR_CUDA;                                                                          \
    }                                                                                               \
  } while (0)

#define CUP_TRY(call)          \
  do {                         \
    int rc_ = (call);          \
    if (rc_ != CUP_OK)         \
      return rc_;              \
  } while (0)

// neighbour codes: >= 0 local slot; NBR_WALL; NBR_COARSE (AMR); <= NBR_REMOTE0: face
// number (NBR_REMOTE0 - code) of the level's received-face buffer (owned by another rank)
enum { NBR_WALL = -1, NBR_COARSE = -2, NBR_REMOTE0 = -3, NBR_FINE = -2147483647 - 1 };

// One multigrid level == one AMR level (reference: struct Lvl, main.c:4443).
// "active" blocks are those whose level == L in that level's context: leaves
// of level L plus the synthesised parents of level L+1 blocks (main.c:4540,
// :4590).  Everything is addressed through SLOTS: slot < nblk is a leaf in
// block-index order, slot >= nblk is a synthesised parent.
struct Level {
  int L = 0;
  double h = 0;
  std::vector<int> act;    // [nact] slot of each active block
  std::vector<int> ijk;    // [nact][3] block index at this level (host only)
  std::vector<int> nbr;    // [nact][6] slot of -x,+x,-y,+y,-z,+z neighbour, NBR_WALL / NBR_COARSE / NBR_FINE
  std::vector<int> ext;    // [nact][6][4] coarse-fine faces: {coarse slot, quadrant} or the 4 finer slots
  std::vector<double> hblk;  // [nact] cell size per block (leaf context only; MG levels share v.h)
  // leaf context only: block coordinates and a (level,ix,iy,iz) -> slot hash for the wide (ss = 3)
  // coarse-fine ghost fill of k_advdiff, which samples the coarse-level view around a block
  std::vector<int> bijk;                   // [nact][4] level, ix, iy, iz
  std::vector<unsigned long long> hkeys;   // open addressing, 0 = empty
  std::vector<int> hvals;
  int *d_bijk = nullptr, *d_hvals = nullptr;
  unsigned long long *d_hkeys = nullptr;
  std::vector<int> pslot;  // [nact] slot of the parent (level L-1), L >= 1
  std::vector<int> oct;    // [nact] octant inside the parent, (ix&1)+2(iy&1)+4(iz&1)
  std::vector<int> par;    // [npar] indices into act[] of blocks that are synthesised parents
  int *d_act = nullptr, *d_nbr = nullptr, *d_pslot = nullptr, *d_oct = nullptr, *d_par = nullptr;
  int *d_ext = nullptr;
  void *d_hblk = nullptr;  // Real [nact]
  bool uniform = true;     // no NBR_COARSE entries
  long long gnact = 0;     // active blocks of this level over all ranks
  // ---- multi-rank plans (empty on one rank) ----
  std::vector<int> inner, bnd;             // act[] indices without / with a remote neighbour
  int *d_inner = nullptr, *d_bnd = nullptr;
  int nface_recv = 0;                      // faces received per exchange (64 Reals each)
  std::vector<int> face_rcnt, face_scnt;   // [nranks] faces received from / sent to each peer
  std::vector<int> face_sslot, face_splane;  // send list: local slot + plane (0..5), peer-major
  int *d_face_sslot = nullptr, *d_face_splane = nullptr;
  // restriction to remote parents: pslot[k] <= NBR_REMOTE0 -> entry of the send buffer
  std::vector<int> res_scnt, res_rcnt;     // [nranks] children sent to / received from each peer
  std::vector<int> res_rslot, res_roct;    // received children: local parent slot, octant
  int *d_res_rslot = nullptr, *d_res_roct = nullptr;
  // ---- one-sided (peer-to-peer) addressing: where each entry lands in the DESTINATION rank's
  // receive window.  Derived from the global mesh, so no plan negotiation is needed.
  std::vector<int> face_speer, face_sidx;  // [nsend] destination rank, entry number in its face area
  std::vector<int> res_speer, res_sidx;    // [children with remote parent] dest rank, entry in its restrict area
  std::vector<int> pro_speer, pro_sidx;    // [received children] dest rank (the child's owner), entry in its prolong area
  std::vector<long long> win_face, win_res, win_pro;  // [nranks] offsets (Reals) of rank p's areas for this level
  std::vector<int> win_nrecv;              // [nranks] faces rank p receives per exchange (parity stride / 64)
  std::vector<long long> win_slab;         // [nranks] offset of rank p's SLAB area (finest level only, else -1)
  // device scratch (comm.cu): faces out/in [n][64], restriction out/in [n][128]
  void *d_fsend = nullptr, *d_frecv = nullptr, *d_rsend = nullptr, *d_rrecv = nullptr;
  void *d_precv = nullptr;                 // prolongation corrections for children with a remote parent
  // per-entry destination pointers (local staging, or peer memory in one-sided mode)
  void **d_fptr0 = nullptr, **d_fptr1 = nullptr, **d_rptr = nullptr, **d_pptr = nullptr;
  std::vector<int> speers[3], rpeers[3];   // ranks signalled / awaited per kind (face, restrict, prolong)
  int *d_speers[3] = {nullptr, nullptr, nullptr}, *d_rpeers[3] = {nullptr, nullptr, nullptr};
  bool p2p = false;
  long long rface_stride = 0;              // Reals between the two parities of the face area
  void *d_seq = nullptr;                   // this level's sequence numbers (one-sided mode)
  // multi-component / multi-layer ghost slabs of the stencil sweeps (SLAB_PLANES planes per face)
  void *d_ssend = nullptr, *d_srecv = nullptr;  // staging (NCCL mode) / receive area
  void **d_sptr0 = nullptr, **d_sptr1 = nullptr;
  long long slab_stride = 0;               // Reals between the two parities of the slab area
  // fused sweep + exchange (smooth_tma.cu): work order = boundary blocks first, then interior
  std::vector<int> order, bsend;           // [nact] act indices; [nact][6] send entry of (block, plane) or -1
  int *d_order = nullptr, *d_bsend = nullptr;
  unsigned int *d_counters = nullptr;      // [8]: [0] retired boundary blocks, [1] retired CTAs (fused sweep);
                                           // [2..7] retired CTAs of the other posting kernels (comm_dev.cuh)
  // fused prolongation (UpFuse, smooth_tma.cuh): [nact][7] parent index/octant of own + face neighbours
  std::vector<int> upinfo;
  int *d_upinfo = nullptr;
  // ---- ghost BLOCKS (multi-level meshes across ranks; the reference ships whole blocks for every
  // remote neighbour, halo_sync main.c:3101-3112).  On a level that has coarse-fine interfaces anywhere
  // (and in the leaf context of a multi-level mesh) every block of another rank that a local block
  // reads -- same-level neighbours, the coarser leaf behind an interface, the finer leaves, and for the
  // wide advdiff stencil the edge / corner neighbours of interface blocks -- has a local GHOST SLOT
  // appended after the rank's own slots; the tables refer to it like to any local block, so the
  // interface kernels need no remote cases.  Before a sweep the owners push those blocks (comm.cu).
  int xid = 0;                             // exchange id: index of this context's sequence numbers / flag words
  bool ghosted = false;
  int nghost = 0;                          // ghost slots of this context (leaf context: after nblk; MG: after own slots)
  std::vector<int> blk_sslot, blk_skind;   // send: local slot, kind (0: the swept vector, 1: the canonical U0)
  std::vector<int> blk_speer, blk_sidx;    // send: destination rank, entry in its block area
  std::vector<int> blk_rslot, blk_rkind;   // recv (entry order): local ghost slot, kind
  std::vector<int> blk_scnt, blk_rcnt;     // [nranks]
  std::vector<long long> win_blk;          // [nranks] offset (Reals) of rank p's block area for this context
  std::vector<int> win_nblk;               // [nranks] entries rank p receives
  int *d_blk_sslot = nullptr, *d_blk_skind = nullptr, *d_blk_rslot = nullptr, *d_blk_rkind = nullptr;
  std::vector<int> blk_speers, blk_rpeers; // ranks notified / awaited by the block exchange
  int *d_blk_speers = nullptr, *d_blk_rpeers = nullptr;
  void **d_bptr0 = nullptr, **d_bptr1 = nullptr;  // per send entry: destination (parity 0 / 1)
  void *d_brecv = nullptr;                 // own block area
  long long blk_stride = 0;                // Reals between the two parities
  int blk_ncomp = 1;                       // components per entry the area is sized for (leaf context: 7)
  // levels with coarse-fine interfaces: blocks whose six neighbours are all same-level / wall run
  // the fast (TMA) kernels, the others the generic ghost fill
  std::vector<int> reg, irr, par_reg, par_irr;
  int *d_reg = nullptr, *d_irr = nullptr, *d_par_reg = nullptr, *d_par_irr = nullptr;
  // leaf context: the regular blocks once more, grouped by level (one h per launch of the fast operator kernel)
  std::vector<std::vector<int>> reg_by_level;
  std::vector<int *> d_reg_by_level;
};

// pure-host result of the topology build (mesh.cpp); also what the CPU tests inspect
struct HostMesh {
  int nranks = 1, rank = 0, top = -1, level_max = 1;
  int bpd[3] = {1, 1, 1};
  long long nblk = 0, nslot = 0, gblocks = 0, gnslot = 0, pin_local = -1;
  long long nown = 0;  // own slots (leaves + own parents); slots >= nown are ghost blocks of other ranks
  std::vector<long long> win_reals;  // [nranks] size of every rank's receive window, in Reals
  double gvol = 0;
  bool leaf_uniform = true;
  std::vector<CupBlk> blk;
  std::vector<Level> lv;
  Level leafv;  // ALL local leaves (every level) with their neighbours: context of pois_op and the stencil sweeps
};

struct Krylov;

}  // namespace cup

struct CupCtx {
  int device = 0;
  int real_bytes = 8;
  int num_sms = 148;
  cudaStream_t stream = nullptr;    // where all work is enqueued (own_stream unless the caller set one)
  cudaStream_t own_stream = nullptr;  // blocking stream: ordered against the legacy default stream
  CupParams prm{};
  long long nblk = 0, nslot = 0;   // LOCAL leaves / local slots (own + ghost blocks of other ranks)
  long long nstate = 0;            // blocks per state component: nblk + leaf-context ghost blocks
  void *leaf_ghost = nullptr;      // ghost blocks of a flat vector in the leaf context (pois_op): [leafv.nghost][512]
  long long gblocks = 0;           // leaves over all ranks
  double gvol = 0;                 // volume over all ranks (pois_solve's vol)
  long long pin_local = -1;        // local index of block (0,0,0) or -1 (pois_pin)
  int rank = 0, nranks = 1;
  std::vector<long long> win_reals;  // [nranks] receive-window sizes (Reals)
  void *comm = nullptr;            // cup::Comm (comm.cu)
  int *h_err = nullptr;            // pinned + mapped: nonzero after a peer-flag wait timed out (comm_dev.cuh)
  int top = -1;
  int bpd[3] = {1, 1, 1};
  int level_max = 1;
  std::vector<CupBlk> blk;
  std::vector<cup::Level> lv;   // index = level
  cup::Level leafv;             // all leaves (multi-level meshes); on single-level meshes == lv[finest]
  bool leaf_uniform = true;     // all leaves on one level (fast stencil path)
  // device state: 9 components, each [nblk][512] Real
  void *state[CUP_F_N] = {nullptr};
  // multigrid scratch (Real): pong for leaves, and u/pong/f/us for parent slots
  void *u1_leaf = nullptr, *u0_x = nullptr, *u1_x = nullptr, *f_x = nullptr, *us_x = nullptr;
  // leaf-sized temporaries used by the host-pointer entry points and drivers
  void *tmp_in = nullptr, *tmp_out = nullptr, *tmp_stage = nullptr;
  void *d_W = nullptr;          // FDM eigenvalue table, lane-major [8][64]
  void *d_hw = nullptr;         // per-leaf h^3 (pois.hw, main.c:4886, is its reciprocal)
  double *d_scal = nullptr;     // device scalars (reductions); always double
  double *h_scal = nullptr;     // pinned mirror
  cup::Krylov *kr = nullptr;
  long long launches = 0;
  void *p_old = nullptr;        // projection(): previous pressure
  void *vel_spare[3] = {nullptr, nullptr, nullptr};  // advdiff(): second velocity buffer of the fused RK stages
  const void *adv_rk = nullptr;  // cup::AdvRk of the stage being swept (advdiff_t -> stencil_t), else null
  void *graph_cache = nullptr;  // captured V-cycles keyed by (in, out) (mg_kernels.cu)
  void *tma_cache = nullptr;    // tensor-map cache (smooth_tma.cu)
  long long tma_extra_rows = 0; // > 0: blocks behind the `extra` pointer of the next face maps (leaf-context ghost scratch)
  bool no_flux_correction = false;  // st_mg on the leaves (stencil_apply(CUP_ST_MG)): k_mg has no flux faces
  void *obst = nullptr;         // cup::Obstacles (obstacle.cu)
  void *io_buf = nullptr;       // io_dump packing: 5 floats per cell (allocated on first use)
  void *xfer = nullptr;         // staging buffers of cup_state_h2d / d2h (capi.cu)
  bool keep_tmp_udef = false;   // projection(): F_TMP already holds fish_tmpv()'s udef
  // stencil_run(st, list, n) (main.c:3631): the caller's block list on the device while a listed
  // sweep runs; d_list holds 2*nblk ints (the list, and its split into regular / interface blocks)
  int *d_list = nullptr;
  const int *run_sub = nullptr;
  int run_nsub = -1;            // < 0: no list, sweep every block
  std::vector<int> run_list;    // host copy of the current list
};

namespace cup {

// mesh.cpp
int build_tables(HostMesh *m, const CupBlk *gblk, long long G, const int *owner, int nranks, int rank,
                 const int bpd[3], int level_max);
int build_mesh(CupCtx *c, const CupBlk *gblk, long long G, const int *owner, const int bpd[3], int level_max);
void free_mesh(CupCtx *c);

// mg_kernels.cu
int mg_setup(CupCtx *c);  // constants + scratch after build_mesh
void free_graph_cache(CupCtx *c);
int mg_vcycle_dev(CupCtx *c, const void *d_in, void *d_out);
int pois_op_dev(CupCtx *c, const void *d_in, void *d_out);
int mg_smooth_slots(CupCtx *c, int level, int n, void *d_u, const void *d_f);
int time_smooth(CupCtx *c, int level, int reps, float *ms);
int trace_report(CupCtx *c, char *out, size_t cap);

// blas_kernels.cu
int wdot(CupCtx *c, const void *a, const void *b, int scal_idx);  // -> d_scal[idx] (accumulates from 0)
int fetch_scalars(CupCtx *c, int first, int n);                   // d_scal -> h_scal, synchronises
int umax(CupCtx *c, double *out);                                 // sta_umax over all ranks
int scale_blk3(CupCtx *c, void *a0, void *a1, void *a2);          // a_q[b] *= 1/h_b^3
int io_pack(CupCtx *c, float *h_attr, float *h_vort, float *h_q);  // F_CHI, F_TMP, F_LHS -> float32 host arrays
int block_linf(CupCtx *c, int f0, double *h_all, double *h_fluid);  // per-block |.|_inf of 3 components

// solver.cu
int pois_solve(CupCtx *c, CupSolveInfo *info);
int advdiff(CupCtx *c);
int projection(CupCtx *c, CupSolveInfo *info);
int vorticity(CupCtx *c);  // vorticity(), main.c:5786: k_vort then 1/h^3
int stencil_run(CupCtx *c, CupStencilId id, const long long *list, long long n);
void free_krylov(CupCtx *c);

// adapt_kernels.cu
int gradchi(CupCtx *c);  // k_gradchi (main.c:3649): F_CHI -> marks in F_TMP
int adapt_fields(CupCtx *c, const CupBlk *nb, long long n_new, const int *kind, const long long *src,
                 void *out[CUP_F_N]);

// obstacle.cu
int obstacle_upload(CupCtx *c, int body, int nob, const int *blk, const double *chi, const double *udef);
int obstacle_motion(CupCtx *c, int body, const double com[3], const double vel[3], const double omega[3]);
int obstacle_clear(CupCtx *c);
int obstacle_moments(CupCtx *c, int body, double *M);
int obstacle_penalize(CupCtx *c);
int obstacle_tmpv(CupCtx *c);
int obstacle_count(const CupCtx *c);
void free_obstacles(CupCtx *c);

// comm.cu
int comm_init(CupCtx *c, int rank, int nranks, const void *id, size_t id_bytes);
int comm_init_host(CupCtx *c, int rank, int nranks, CupAllgatherFn fn, void *user);
int comm_check_error(CupCtx *c);  // after a stream synchronisation: CUP_ERR_COMM if a wait timed out
int comm_unique_id(void *out, size_t bytes);

enum { SCAL_N = 256, SLAB_PLANES = 9, BLK_COMPS = 7 };  // leaf ghost blocks carry up to 7 fields (k_prhs: vel, udef, chi)  // 3 components x 3 layers (k_advdiff) is the largest slab

}  // namespace cup
