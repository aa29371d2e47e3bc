/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the UNMODIFIED reference translation unit (/root/reference/main.c,
 * found through -I at build time; never copied into this repository) together
 * with the single-rank MPI shim (oracle/shim/mpi.h) and exports a handful of
 * thin entry points so that tests and bench.py's cpu_baseline leg can drive
 * the reference's own hot-path functions:
 *
 *   stencil_apply  main.c:3648     mg_vcycle   main.c:4831
 *   pois_op        main.c:4282     pois_solve  main.c:4875
 *   advdiff        main.c:5027     projection  main.c:5828
 *
 * Every symbol of the reference is file-static, hence the #include of the
 * whole TU.  Nothing of the product (cup3d_b200/) may link or load this.
 */
#define main cup3d_reference_main
#include "main.c"
#undef main

#include <omp.h>
#include <time.h>
#include <unistd.h>

#define API __attribute__((visibility("default")))

static int ref_ready;

/* The binding always passes doubles.  In the Real=float build (oracle/Makefile target ref32: the
 * reference's `typedef double Real` / `#define MPI_Real` lines switched to float in a temporary
 * copy at build time, main.c:14-15) vectors are converted at this boundary. */
static Real *to_real(const double *in, long long n) {
  Real *p = malloc((size_t)n * sizeof *p);
  long long i;
  for (i = 0; i < n; i++)
    p[i] = (Real)in[i];
  return p;
}
static void from_real(double *out, const Real *p, long long n) {
  long long i;
  for (i = 0; i < n; i++)
    out[i] = (double)p[i];
}
API int ref_real_bytes(void) { return (int)sizeof(Real); }

/* argv = the reference's own "-key value" command line (without argv[0]);
 * tabdir = directory holding lab_ss*_t*.bin (the reference opens them in cwd,
 * main.c:3350-3353). */
API int ref_init(int argc, char **argv, const char *tabdir) {
  char cwd[4096];
  char **av;
  char *content;
  int i;
  if (ref_ready)
    return 1;
  av = malloc((argc + 2) * sizeof *av);
  av[0] = "ref";
  for (i = 0; i < argc; i++)
    av[i + 1] = argv[i];
  av[argc + 1] = NULL;
  sim.comm = MPI_COMM_WORLD;
  sim.rank = 0;
  sim.size = 1;
  content = param_parse(argc + 1, av);
  sta_init();
  if (getcwd(cwd, sizeof cwd) == NULL)
    return 2;
  if (chdir(tabdir) != 0)
    return 3;
  lab_tables();
  if (chdir(cwd) != 0)
    return 4;
  pois_init();
  fish_parse(content);
  mesh_init();
  free(av);
  ref_ready = 1;
  return 0;
}

API long long ref_nblk(void) { return sta.nblk; }
API int ref_threads(void) { return omp_get_max_threads(); }

/* out: 5 ints per block (level, ix, iy, iz, 0) and h, origin[3] */
API void ref_blocks(int *ib, double *rb) {
  long long i;
  for (i = 0; i < sta.nblk; i++) {
    struct Blk *b = &sta.blk[i];
    ib[4 * i + 0] = b->level;
    ib[4 * i + 1] = b->ix;
    ib[4 * i + 2] = b->iy;
    ib[4 * i + 3] = b->iz;
    rb[4 * i + 0] = b->h;
    rb[4 * i + 1] = b->origin[0];
    rb[4 * i + 2] = b->origin[1];
    rb[4 * i + 3] = b->origin[2];
  }
}

/* raw state: sta.fld is [nblk][F_N][512] (main.c:55-58,131-132) */
API void ref_state_get(double *out) { from_real(out, sta.fld, sta.nblk * BLK_S); }
API void ref_state_set(const double *in) {
  long long i, n = sta.nblk * BLK_S;
  for (i = 0; i < n; i++)
    sta.fld[i] = (Real)in[i];
}

API void ref_set_scalars(double dt, double nu, double uinfx, double uinfy, double uinfz, int step,
                         int mean_constraint, double ptol, double ptol_rel) {
  sta.dt = dt;
  sim.nu = nu;
  sta.uinf[0] = uinfx;
  sta.uinf[1] = uinfy;
  sta.uinf[2] = uinfz;
  sta.step = step;
  sim.mean_constraint = mean_constraint;
  sim.ptol = ptol;
  sim.ptol_rel = ptol_rel;
}

API void ref_mg_vcycle(double *in, double *out) {
  long long n = sta.nblk * BS3;
  Real *a = to_real(in, n), *b = malloc((size_t)n * sizeof *b);
  mg_vcycle(a, b);
  from_real(out, b, n);
  free(a);
  free(b);
}
API void ref_pois_op(double *in, double *out) {
  long long n = sta.nblk * BS3;
  Real *a = to_real(in, n), *b = malloc((size_t)n * sizeof *b);
  pois_op(a, b);
  from_real(out, b, n);
  free(a);
  free(b);
}
API void ref_pois_solve(void) { pois_solve(); }
API void ref_advdiff(void) { advdiff(); }
API void ref_projection(void) { projection(); }

/* 0 lhs, 1 mg, 2 advdiff, 3 prhs, 4 divp, 5 gradp, 6 vort, 7 q, 8 gradchi */
API int ref_stencil(int id) {
  struct Stencil *tab[] = {&st_lhs, &st_mg, &st_advdiff, &st_prhs, &st_divp, &st_gradp, &st_vort, &st_q, &st_gradchi};
  if (id < 0 || id >= (int)(sizeof tab / sizeof *tab))
    return 1;
  stencil_apply(tab[id]);
  return 0;
}

/* Weighted dot used by the Krylov solver (main.c:4854); builds pois.hw first. */
API double ref_pois_dot(double *a, double *b) {
  long long N = sta.nblk * BS3, i;
  Real *ra = to_real(a, N), *rb = to_real(b, N);
  double r;
  pois_alloc(N);
  for (i = 0; i < sta.nblk; i++)
    pois.hw[i] = 1 / (sta.blk[i].h * sta.blk[i].h * sta.blk[i].h);
  r = pois_dot(ra, rb, N);
  free(ra);
  free(rb);
  return r;
}

/* the block-local FDM inverse (main.c:4368) on one 512-vector */
API void ref_pre_blk(double *src, double *dst, double invh) {
  Real a[BS3], b[BS3], s[BS3], d[BS3];
  int i;
  for (i = 0; i < BS3; i++)
    s[i] = (Real)src[i];
  pre_blk(s, d, invh, a, b);
  from_real(dst, d, BS3);
}

static double now(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

/* time n V-cycles (after w warm-ups); returns seconds for the n cycles */
API double ref_time_vcycle(double *in, double *out, int w, int n) {
  int k;
  double t0, t1;
  long long N = sta.nblk * BS3;
  Real *a = to_real(in, N), *b = malloc((size_t)N * sizeof *b);
  for (k = 0; k < w; k++)
    mg_vcycle(a, b);
  t0 = now();
  for (k = 0; k < n; k++)
    mg_vcycle(a, b);
  t1 = now();
  from_real(out, b, N);
  free(a);
  free(b);
  return t1 - t0;
}

API double ref_time_stencil(int id, int w, int n) {
  int k;
  double t0;
  for (k = 0; k < w; k++)
    ref_stencil(id);
  t0 = now();
  for (k = 0; k < n; k++)
    ref_stencil(id);
  return now() - t0;
}

/* vorticity() (main.c:5786): k_vort sweep, then 1/h^3 per block */
API void ref_vorticity(void) { vorticity(); }

/* sta_umax (main.c:5918): max over cells of max_a |u_a + uinf_a| */
API double ref_umax(void) { return sta_umax(); }

/* one mesh adaptation pass with the reference's own tagging (main.c:4012ff) */
API void ref_mesh_adapt(double rtol, double ctol) {
  sim.rtol = rtol;
  sim.ctol = ctol;
  mesh_adapt();
}

/* ---- obstacle (fish) path: SURVEY 8(f) row 1 ------------------------------------------
 * Drives the reference's own per-step phases one at a time so that goldens can be
 * captured between them (advance(), main.c:5984-6003). */
API void ref_sta_fields(void) { sta_fields(); }
API int ref_nfish(void) { return sim.nfish; }
API double ref_sta_dt(void) { return sta_dt(); }
API void ref_step_end(void) {
  sta.step++;
  sta.time += sta.dt;
}
/* 0 fish_build  1 advdiff  2 fish_vel  3 fish_pen  4 projection  5 fish_tmpv  6 mesh_adapt */
API int ref_phase(int p) {
  switch (p) {
  case 0: fish_build(); break;
  case 1: advdiff(); break;
  case 2: fish_vel(); break;
  case 3: fish_pen(); break;
  case 4: projection(); break;
  case 5: fish_tmpv(); break;
  case 6: mesh_adapt(); break;
  default: return 1;
  }
  return 0;
}
/* dt lambda uinf[3] step time */
API void ref_get_scalars(double *o) {
  o[0] = sta.dt;
  o[1] = sta.lambda;
  o[2] = sta.uinf[0];
  o[3] = sta.uinf[1];
  o[4] = sta.uinf[2];
  o[5] = sta.step;
  o[6] = sta.time;
}
API void ref_set_lambda(double l) { sta.lambda = l; }
API int ref_fish_nob(int k) {
  long long i;
  int n = 0;
  for (i = 0; i < sta.nblk; i++)
    n += oblock(&sta.fish[k], i) != NULL;
  return n;
}
/* obstacle blocks of fish k in block order: index, chi[512], udef[512][3] (struct layout) */
API void ref_fish_ob(int k, int *blk, double *chi, double *udef) {
  long long i;
  int n = 0;
  for (i = 0; i < sta.nblk; i++) {
    struct ObstacleBlock *o = oblock(&sta.fish[k], i);
    if (o == NULL)
      continue;
    blk[n] = (int)i;
    memcpy(chi + (size_t)n * BS3, o->chi, BS3 * sizeof(Real));
    memcpy(udef + (size_t)n * 3 * BS3, o->udef, 3 * BS3 * sizeof(Real));
    n++;
  }
}
/* com[3] vel[3] omega[3] */
API void ref_fish_motion(int k, double *o) {
  struct Fish *f = &sta.fish[k];
  int d;
  for (d = 0; d < 3; d++) {
    o[d] = f->com[d];
    o[3 + d] = f->vel[d];
    o[6 + d] = f->omega[d];
  }
}
API int ref_m_n(void) { return M_N; }
/* per-block moments (fish_mom_blk, main.c:5057) summed over blocks in block order, as the
 * first loop of fish_vel does (main.c:5315-5325); fish_solve is NOT called */
API void ref_fish_mom(int k, double *M) {
  struct Fish *f = &sta.fish[k];
  long long i;
  int q;
  for (q = 0; q < M_N; q++)
    M[q] = 0;
  for (i = 0; i < sta.nblk; i++) {
    struct ObstacleBlock *o = oblock(f, i);
    if (o == NULL)
      continue;
    fish_mom_blk(i, f);
    for (q = 0; q < M_N; q++)
      M[q] += o->mom[q];
  }
}

/* ---- the host halves of fish_vel / fish_pen, for a caller whose per-cell halves run elsewhere
 * (tests/coupled.py: the integration recipe of INTEGRATION.md, step by step) ---- */

/* what fish_vel does with the reduced moments M (main.c:5328-5340): hand them to the 6x6
 * rigid-body solve of fish k */
API void ref_fish_solve_from(int k, const double *M) {
  struct Fish *f = &sta.fish[k];
  int d, q;
  f->pen_m = M[M_GfX];
  for (d = 0; d < 3; d++) {
    f->pen_cm[d] = M[M_GpX + d];
    f->pen_lmom[d] = M[M_GuX + d];
    f->pen_amom[d] = M[M_GaX + d];
  }
  for (q = 0; q < 6; q++)
    f->pen_j[q] = M[M_Gj0 + q];
  fish_solve(f);
}
/* fish_pen = fish_hit (collisions, host) + the per-block penalisation loop (main.c:5651-5662) */
API void ref_fish_hit(void) {
  if (sim.nfish > 0)
    fish_hit();
}
API void ref_fish_pen_blocks(void) {
  long long i;
#pragma omp parallel for schedule(dynamic, 1)
  for (i = 0; i < sta.nblk; ++i) {
    int k;
    for (k = 0; k < sim.nfish; k++)
      fish_pen_blk(i, &sta.fish[k]);
  }
}
API int ref_mesh_changed(void) { return sta.mesh_changed; }
