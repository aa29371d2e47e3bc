/*
 * cup3d_b200.h -- C ABI of the B200-native CUP3D hot path.
 *
 * The reference (slitvinov/CUP3D, main.c) has no plugin interface: every
 * symbol is file-static in one translation unit.  The boundary reproduced
 * here is therefore the CALL SURFACE its time loop uses:
 *
 *   reference (main.c)                     this library
 *   ------------------------------------   ---------------------------------
 *   struct Blk                  :59-63     CupBlk (same members, same order)
 *   sta.fld [nblk][F_N][512]    :55-58     cup_state_h2d / cup_state_d2h
 *   tree_sync+halo_build+fc_prepare+
 *     mg_build (rebuild hook)   :3325-3328 cup_mesh_upload
 *   struct Stencil st_*         :3627-3630 CupStencilId
 *   stencil_apply / stencil_run :3631-3648 cup_stencil_apply / cup_stencil_run
 *   pois_op(in,out)             :4282      cup_pois_op
 *   mg_vcycle(in,out)           :4831      cup_mg_vcycle
 *   pois_solve()                :4875      cup_pois_solve
 *   advdiff()                   :5027      cup_advdiff
 *   projection()                :5828      cup_projection
 *   halo_sync / xch_exec (MPI)  :2899,:3101 cup_comm_init (NCCL, one rank/GPU)
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every function
 * returns 0 on success and a negative CupStatus otherwise (the reference
 * aborts through fatal(), main.c:134); cup_last_error() gives the text.
 * There is NO CPU fallback: without a CUDA device every call fails.
 * Vectors named h_* are host pointers, d_* are device pointers.  All flat
 * vectors are in block-index order, 512 cells per block, x fastest
 * (IDX(x,y,z) = (z*8+y)*8+x, main.c:58).
 */
#ifndef CUP3D_B200_H
#define CUP3D_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { CUP_BS = 8, CUP_BS3 = 512 };
/* field offsets inside one block of sta.fld (main.c:55) */
enum { CUP_F_CHI = 0, CUP_F_PRES = 1, CUP_F_VEL = 2, CUP_F_TMP = 5, CUP_F_LHS = 8, CUP_F_N = 9 };

typedef enum {
  CUP_OK = 0,
  CUP_ERR_ARG = -1,     /* bad argument / unsupported configuration */
  CUP_ERR_CUDA = -2,    /* CUDA runtime error */
  CUP_ERR_MESH = -3,    /* inconsistent mesh (missing neighbour / sibling) */
  CUP_ERR_STATE = -4,   /* call order (mesh not uploaded, ...) */
  CUP_ERR_NCCL = -5,
  CUP_ERR_UNSUPPORTED = -6,
  CUP_ERR_COMM = -7     /* a peer rank did not answer in time (one-sided exchange) */
} CupStatus;

/* struct Blk, main.c:59-63 (Real = double in the reference build) */
typedef struct CupBlk {
  int level, ix, iy, iz;
  long long Z;
  double h, origin[3];
} CupBlk;

/* the reference's Stencil instances (main.c:4267,4280,5026,5699,5714,5735) */
typedef enum {
  CUP_ST_LHS = 0,     /* st_lhs     F_PRES -> F_LHS   */
  CUP_ST_MG = 1,      /* st_mg      F_PRES -> F_LHS (no flux faces) */
  CUP_ST_ADVDIFF = 2, /* st_advdiff F_VEL  -> F_TMP += */
  CUP_ST_PRHS = 3,    /* st_prhs    F_VEL,F_TMP,F_CHI -> F_LHS */
  CUP_ST_DIVP = 4,    /* st_divp    F_PRES -> F_TMP[0] */
  CUP_ST_GRADP = 5,   /* st_gradp   F_PRES -> F_TMP[0..2] */
  CUP_ST_VORT = 6,    /* st_vort    F_VEL  -> F_TMP[0..2] (h^3-weighted vorticity, with its flux correction) */
  CUP_ST_Q = 7,       /* st_q       F_VEL  -> F_LHS (Q criterion) */
  CUP_ST_GRADCHI = 8  /* st_gradchi F_CHI (ss = 2, tensorial) -> F_TMP: cells inside bodies zeroed, blocks whose
                         extended neighbourhood holds 1e-5 < chi < 0.9 marked with 1e10 (main.c:3649-3681) */
} CupStencilId;

/* run-time scalars the kernels read from the reference's sim/sta globals */
typedef struct CupParams {
  double dt;             /* sta.dt */
  double nu;             /* sim.nu */
  double uinf[3];        /* sta.uinf */
  int step;              /* sta.step (incremental pressure after STEP_2ND) */
  int mean_constraint;   /* sim.mean_constraint (bMeanConstraint) */
  double ptol, ptol_rel; /* sim.ptol, sim.ptol_rel */
  double lambda;         /* sta.lambda (penalisation coefficient, main.c:5969) */
} CupParams;

typedef struct CupSolveInfo {
  int iterations;   /* Krylov iterations (it, main.c:4954) */
  int restarts;
  double residual;  /* final ||r||/sqrt(vol) */
  double rhs_norm;  /* ||b||/sqrt(vol) */
  int vcycles;      /* V-cycles applied */
} CupSolveInfo;

typedef struct CupCtx CupCtx;

const char *cup_last_error(void);
int cup_version(void);

/* real_bytes: 8 (the reference build, typedef double Real) or 4 */
int cup_create(CupCtx **ctx, int device, int real_bytes);
int cup_destroy(CupCtx *ctx);
/* run all later work on this CUDA stream (cudaStream_t as void*).  NULL selects the library's own
 * BLOCKING stream (ordered against the caller's legacy default stream), not the NULL stream: the
 * NULL stream cannot be captured into the CUDA graphs the V-cycle replays from. */
int cup_set_stream(CupCtx *ctx, void *stream);
int cup_set_params(CupCtx *ctx, const CupParams *p);
int cup_synchronize(CupCtx *ctx);

/* Rebuild hook.  blk = this rank's sta.blk[0..n), bpd = sim.bpdx/y/z,
 * level_max = sim.level_max.  Builds neighbour tables, flux-face plans and
 * the multigrid hierarchy on the device (replaces main.c:3325-3328).
 * CONTRACT: the nine state fields are re-allocated for the new block count and ZEROED -- after
 * an adaptation the caller uploads sta.fld again with cup_state_h2d (the reference moves block
 * data inside mesh_adapt on the host, main.c:4012).  Obstacle blocks are dropped as well.  On any
 * failure the context is left WITHOUT a mesh (later calls return CUP_ERR_STATE). */
int cup_mesh_upload(CupCtx *ctx, const CupBlk *blk, long long n, const int bpd[3], int level_max);
long long cup_nblk(const CupCtx *ctx);
long long cup_nslot(const CupCtx *ctx); /* leaves + synthesised MG parents */
int cup_mg_levels(const CupCtx *ctx);
long long cup_mg_nact(const CupCtx *ctx, int level);

/* sta.fld <-> device.  h_fld is the reference's layout [nblk][9][512] doubles;
 * f0/nc select a field range (e.g. CUP_F_VEL,3). */
int cup_state_h2d(CupCtx *ctx, const double *h_fld, int f0, int nc);
int cup_state_d2h(CupCtx *ctx, double *h_fld, int f0, int nc);
/* device pointer of one state component: flat [nblk][512] Reals.  Ask again after cup_advdiff: on uniform
 * meshes its fused Runge-Kutta stages ping-pong F_VEL between two buffers and the pointers swap. */
void *cup_state_dev(CupCtx *ctx, int f);

/* the per-block compute loop.  cup_stencil_run is stencil_run(st, list, n) (main.c:3631-3647):
 * ghost exchange of the whole input field, then the kernel on local blocks list[0..n) -- on the
 * first n blocks when list is NULL.  Blocks that are not listed keep their output.  Entries must
 * be distinct and inside [0, cup_nblk) (CUP_ERR_ARG otherwise).  cup_stencil_apply == (NULL, nblk).
 * Every device is guarded: all entry points switch to the context's device and restore the
 * caller's current device on return. */
int cup_stencil_apply(CupCtx *ctx, CupStencilId id);
int cup_stencil_run(CupCtx *ctx, CupStencilId id, const long long *list, long long n);

/* flat-vector operators, host buffers (H2D + compute + D2H, synchronous) */
int cup_pois_op(CupCtx *ctx, const double *h_in, double *h_out);
int cup_mg_vcycle(CupCtx *ctx, const double *h_in, double *h_out);
/* same on device-resident vectors of Real (asynchronous on the ctx stream).  The V-cycle is captured
 * into a CUDA graph per (d_in, d_out) pair and replayed; a graph only stores the two addresses and the
 * mesh's tables, so it stays valid if the caller frees and re-allocates a vector at the same address, and
 * every graph is dropped when the mesh changes (cup_mesh_upload / cup_mesh_adapt).  At most 128 pairs are
 * kept. */
int cup_pois_op_dev(CupCtx *ctx, const void *d_in, void *d_out);
int cup_mg_vcycle_dev(CupCtx *ctx, const void *d_in, void *d_out);
/* sum_i a_i b_i / h_i^3 (pois_dot, main.c:4854); synchronous */
int cup_pois_dot_dev(CupCtx *ctx, const void *d_a, const void *d_b, double *result);

/* F_LHS (rhs), F_PRES (guess) -> F_PRES; uses params ptol/ptol_rel/mean_constraint */
int cup_pois_solve(CupCtx *ctx, CupSolveInfo *info);
int cup_advdiff(CupCtx *ctx);
int cup_projection(CupCtx *ctx, CupSolveInfo *info);
/* projection() zeroes F_TMP and lets fish_tmpv() (main.c:5799) add the deformation velocity
 * before the divergence sweep: done on the device for bodies given by cup_obstacle_upload.
 * Alternatively, with flag != 0 the caller has uploaded F_TMP = fish_tmpv() result already
 * and both the zeroing and the device fish_tmpv are skipped. */
int cup_projection_udef_ready(CupCtx *ctx, int flag);

/* ---- obstacle (fish) phases of advance() between advdiff and projection (main.c:5993-5997) ----
 * The host keeps producing the obstacle blocks (fish_build) and solving the 6x6 rigid-body
 * system (fish_solve); the per-cell work runs on the device so F_VEL / F_TMP never leave HBM.
 * A "body" is one struct Fish (main.c:37); its blocks are struct ObstacleBlock (main.c:96-102). */
enum { CUP_MAX_BODIES = 64 };
/* index of each entry of ObstacleBlock.mom / the M[] of fish_vel (enum M_*, main.c:64-95) */
enum {
  CUP_M_V = 0, CUP_M_FX = 1, CUP_M_TX = 4, CUP_M_J0 = 7, CUP_M_GFX = 13, CUP_M_GPX = 14, CUP_M_GJ0 = 17,
  CUP_M_GUX = 23, CUP_M_GAX = 26, CUP_M_N = 29
};
/* after fish_build: blk[nob] = local block index of every block with oblock(f,i) != NULL,
 * chi [nob][512] = ObstacleBlock.chi, udef [nob][512][3] = ObstacleBlock.udef (reference layout).
 * nob = 0 removes the body.  Synchronous (the host arrays may be reused on return). */
int cup_obstacle_upload(CupCtx *ctx, int body, int nob, const int *blk, const double *chi, const double *udef);
/* f->com, f->vel, f->omega (NULL = keep) */
int cup_obstacle_motion(CupCtx *ctx, int body, const double com[3], const double vel[3], const double omega[3]);
int cup_obstacle_clear(CupCtx *ctx);
/* fish_mom_blk over the body's blocks + the block sum and MPI_Allreduce of fish_vel
 * (main.c:5057-5118, :5315-5326) -> M[CUP_M_N]; uses params dt, lambda and the body's com.
 * Synchronous.  The caller feeds M into fish_solve and returns vel/omega with cup_obstacle_motion. */
int cup_obstacle_moments(CupCtx *ctx, int body, double *M);
/* the loop of fish_pen (main.c:5656-5661) over all bodies: F_VEL is penalised towards the body
 * velocity; fish_hit (collisions, host) is NOT part of it */
int cup_obstacle_penalize(CupCtx *ctx);
/* fish_tmpv (main.c:5799): F_TMP += udef inside the bodies.  cup_projection calls it itself
 * when bodies are uploaded; exposed for tests and for callers that drive the sweeps one by one */
int cup_obstacle_tmpv(CupCtx *ctx);

/* sta_umax (main.c:5918): max over all cells (all ranks) of max_a |u_a + uinf_a|; the time-step
 * control (sta_dt) needs only this scalar, so F_VEL can stay on the device. */
int cup_umax(CupCtx *ctx, double *umax);

/* ---- mesh_adapt's tagging input (main.c:4012-4019) ----
 * vorticity() (main.c:5786): k_vort sweep F_VEL -> F_TMP (with its coarse-fine flux correction)
 * scaled by 1/h^3 per block. */
int cup_vorticity(CupCtx *ctx);
/* mesh_tag_blk's norm (main.c:3683): per local block, max over cells of |(f0, f0+1, f0+2)|_2.
 * linf_all[b]: all cells; linf_fluid[b]: cells with F_CHI <= 0.9, i.e. what the norm is after
 * k_gradchi (main.c:3649) zeroed the cells inside bodies.  The OTHER effect of k_gradchi, the
 * 1e10 marker of blocks whose extended chi neighbourhood contains 1e-5 < chi < 0.9, depends on F_CHI
 * alone and stays on the host (F_CHI is produced there by fish_build).  NULL = not wanted. */
int cup_block_linf(CupCtx *ctx, int f0, double *linf_all, double *linf_fluid);

/* The data-parallel half of mesh_adapt (main.c:4012-4190).  The host keeps the tree surgery: from
 * cup_vorticity + cup_stencil_apply(CUP_ST_GRADCHI) + cup_block_linf it tags (mesh_tag), balances (mesh_fix)
 * and builds the NEW block list; this call then produces the new state on the device from the old one and
 * installs the new mesh (what cup_mesh_upload does, but the fields are carried over instead of zeroed):
 *   kind[i] = 0  new block i is old block src[i], unchanged                       (all nine fields copied)
 *   kind[i] = 1  new block i is a child of the REFINED old block src[i]; all eight children must be listed;
 *                F_PRES and F_VEL are interpolated (mesh_refine :3790), the other fields are zero (:4068)
 *   kind[i] = 2  new block i is the parent of eight COMPRESSED old blocks, src[i] = any one of them;
 *                all nine fields are 2x2x2 averages (:4129-4146)
 * One rank only (the reference also rebalances blocks between ranks here, mesh_bal; not built).
 * Obstacle blocks are dropped, as by cup_mesh_upload. */
int cup_mesh_adapt(CupCtx *ctx, const CupBlk *new_blk, long long n_new, const int *kind, const long long *src,
                   const int bpd[3], int level_max);

/* io_dump's field part (main.c:1441-1442, :1525-1535): vorticity(); qcrit();
 * then float32 packing on the device -- attr[nblk*512] = chi, vort[nblk*512][3] (interleaved),
 * q[nblk*512] in block-index order, exactly the arrays io_write() receives.  5 floats per cell
 * cross PCIe instead of 5 doubles.  F_TMP / F_LHS hold vorticity / Q afterwards, as in the
 * reference.  NULL = not wanted.  Synchronous. */
int cup_io_pack(CupCtx *ctx, float *attr, float *vort, float *q);

/* One rank per GPU.  nccl_id = the 128 bytes of an ncclUniqueId created on
 * rank 0 and distributed by the caller (torch.distributed / MPI_Bcast). */
int cup_comm_init(CupCtx *ctx, int rank, int nranks, const void *nccl_id, size_t id_bytes);
int cup_nccl_unique_id(void *out, size_t bytes);
/* The same, bootstrapped by the CALLER's host collective instead of NCCL -- the reference has an
 * MPI communicator (sim.comm, main.c:2899ff), so its binding is
 *     static int ag(void *u, const void *s, void *r, size_t n) {
 *       return MPI_Allgather(s, (int)n, MPI_BYTE, r, (int)n, MPI_BYTE, *(MPI_Comm *)u); }
 *     cup_comm_init_host(ctx, sim.rank, sim.size, ag, &sim.comm);
 * allgather(user, send, recv, bytes): every rank contributes `bytes` bytes, recv gets nranks*bytes
 * in rank order; returns 0 on success.  It is called only inside cup_mesh_upload / cup_destroy
 * (block lists, CUDA IPC handles, barriers).  All DATA moves GPU-to-GPU through CUDA-IPC peer
 * windows (NVLink), reductions included, so this mode needs peer access between the ranks'
 * devices (ranks may also share one device: used by the single-GPU tests). */
typedef int (*CupAllgatherFn)(void *user, const void *send, void *recv, size_t bytes);
int cup_comm_init_host(CupCtx *ctx, int rank, int nranks, CupAllgatherFn allgather, void *user);

/* Host-only view of one rank's exchange plan for one multigrid level (no GPU
 * needed; used by the CPU tests that emulate the exchange over gloo).  Built
 * from the GLOBAL block list + owner rank per block, exactly what
 * cup_mesh_upload derives after its all-gather.  Arrays are malloc'ed; release
 * with cup_plan_free. */
typedef struct CupPlan {
  long long nblk, nslot;      /* local leaves / local slots */
  int nact;                   /* local active blocks of the level */
  int nsend, nrecv;           /* faces sent / received per exchange */
  int *act;                   /* [nact] local slot */
  int *ijk;                   /* [nact][3] block index at this level */
  int *nbr;                   /* [nact][6] local slot, -1 wall, <= -3: received face (-3 - code) */
  int *send_slot, *send_plane;/* [nsend] local slot and plane (0..5) to pack, peer-major */
  int *send_cnt, *recv_cnt;   /* [nranks] */
  int *pslot, *oct;           /* [nact] parent slot (<= -3: entry of the restriction send buffer) */
  int *res_send_cnt, *res_recv_cnt; /* [nranks] children sent to / received from each peer */
  int nres_recv;
  int *res_recv_slot, *res_recv_oct; /* [nres_recv] local parent slot + octant of received children */
  /* ghost BLOCKS (levels with coarse-fine interfaces, and the leaf context = level -1, on several ranks):
   * blocks of other ranks get local slots >= the rank's own slots; nbr / ext refer to them like to local
   * blocks.  kind 0: a block of the swept vector, 1: a coarser leaf read from the canonical vector. */
  int ghosted, nghost;
  int *ext;                          /* [nact][6][4] coarse-fine faces: {coarse slot, quadrant} / four finer slots */
  int nbsend, nbrecv;
  int *bsend_slot, *bsend_kind, *bsend_peer, *bsend_idx; /* [nbsend] own slot, kind, destination rank, its entry */
  int *brecv_slot, *brecv_kind;      /* [nbrecv] (entry order) local ghost slot, kind */
} CupPlan;
/* level = -1: the LEAF context (all local leaves of a multi-level mesh; act = local leaf slots) */
int cup_plan_build(const CupBlk *gblk, long long nglobal, const int *owner, int nranks, int rank, const int bpd[3],
                   int level_max, int level, CupPlan *out);
void cup_plan_free(CupPlan *p);

/* instrumentation: number of kernels launched by this context so far */
long long cup_kernel_launches(const CupCtx *ctx);
/* with CUP_STAMP=1 in the environment every phase of a V-cycle is followed by a one-thread kernel
 * that stores the device's %globaltimer (captured into the replayed CUDA graph like any node):
 * writes "phase nanoseconds" lines of the last cycle into out, returns the bytes written */
int cup_trace_report(CupCtx *ctx, char *out, size_t cap);
/* timing of an internal kernel class with CUDA events on the ctx stream:
 * runs `reps` launches of the level-`level` smoother; returns ms per launch */
int cup_time_smooth(CupCtx *ctx, int level, int reps, float *ms_per_launch);

/* unit-test hooks on device vectors in MG slot space (nslot*512 Reals) */
int cup_mg_smooth_dev(CupCtx *ctx, int level, int n, void *d_u, const void *d_f);
void *cup_mg_array(CupCtx *ctx, int which); /* 0 u, 1 f, 2 us, 3 u-pong */

#ifdef __cplusplus
}
#endif
#endif
