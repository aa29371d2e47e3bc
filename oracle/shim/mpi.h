/*
 * Single-rank MPI shim -- TEST INFRASTRUCTURE ONLY (part of oracle/).
 *
 * The reference (slitvinov/CUP3D, one C99 file) is an MPI program; this image
 * has no MPI.  With exactly one rank every collective degenerates to a memcpy
 * (or to nothing for MPI_IN_PLACE), which is all this header implements.  It
 * exists so that the UNMODIFIED reference source can be compiled where it lies
 * (oracle/Makefile) into oracle/_ref/ and used as the parity oracle and as the
 * CPU baseline.  Only the symbols the reference actually uses are provided.
 *
 * A datatype handle is simply its size in bytes.
 */
#ifndef CUP3D_ORACLE_MPI_SHIM_H
#define CUP3D_ORACLE_MPI_SHIM_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int MPI_Comm;
typedef long MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef FILE *MPI_File;
typedef struct { int unused; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_MAX_ERROR_STRING 64
#define MPI_THREAD_FUNNELED 1
#define MPI_IN_PLACE ((void *)-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_PROC_NULL (-2)
#define MPI_INFO_NULL 0
#define MPI_MODE_CREATE 1
#define MPI_MODE_WRONLY 2

#define MPI_SIGNED_CHAR ((MPI_Datatype)1)
#define MPI_INT ((MPI_Datatype)4)
#define MPI_FLOAT ((MPI_Datatype)4)
#define MPI_LONG ((MPI_Datatype)sizeof(long))
#define MPI_LONG_LONG ((MPI_Datatype)8)
#define MPI_DOUBLE ((MPI_Datatype)8)
#define MPI_DOUBLE_INT ((MPI_Datatype)16)

enum { MPI_SUM = 1, MPI_MAX = 2, MPI_MAXLOC = 3 };

static inline void shim_cp(void *dst, const void *src, long n) {
  if (src != MPI_IN_PLACE && dst != src && n > 0)
    memmove(dst, src, (size_t)n);
}
static inline int MPI_Init_thread(int *c, char ***v, int req, int *prov) {
  (void)c; (void)v; *prov = req; return MPI_SUCCESS;
}
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; fflush(NULL); _Exit(code ? code : 1); }
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return MPI_SUCCESS; }
static inline int MPI_Error_string(int e, char *s, int *n) {
  *n = snprintf(s, MPI_MAX_ERROR_STRING, "shim error %d", e); return MPI_SUCCESS;
}
static inline int MPI_Type_contiguous(int m, MPI_Datatype base, MPI_Datatype *t) { *t = m * base; return MPI_SUCCESS; }
static inline int MPI_Type_commit(MPI_Datatype *t) { (void)t; return MPI_SUCCESS; }
static inline int MPI_Type_free(MPI_Datatype *t) { (void)t; return MPI_SUCCESS; }
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  (void)op; (void)c; shim_cp(r, s, n * t); return MPI_SUCCESS;
}
/* rank 0 of an exclusive scan receives nothing */
static inline int MPI_Exscan(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  (void)s; (void)r; (void)n; (void)t; (void)op; (void)c; return MPI_SUCCESS;
}
static inline int MPI_Allgather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c) {
  (void)rn; (void)rt; (void)c; shim_cp(r, s, sn * st); return MPI_SUCCESS;
}
static inline int MPI_Allgatherv(const void *s, int sn, MPI_Datatype st, void *r, const int *rc, const int *rd,
                                 MPI_Datatype rt, MPI_Comm c) {
  (void)rc; (void)c; shim_cp((char *)r + rd[0] * rt, s, sn * st); return MPI_SUCCESS;
}
static inline int MPI_Gather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root,
                             MPI_Comm c) {
  (void)rn; (void)rt; (void)root; (void)c; shim_cp(r, s, sn * st); return MPI_SUCCESS;
}
static inline int MPI_Gatherv(const void *s, int sn, MPI_Datatype st, void *r, const int *rc, const int *rd,
                              MPI_Datatype rt, int root, MPI_Comm c) {
  (void)rc; (void)root; (void)c; shim_cp((char *)r + rd[0] * rt, s, sn * st); return MPI_SUCCESS;
}
static inline int MPI_Alltoall(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c) {
  (void)rn; (void)rt; (void)c; shim_cp(r, s, sn * st); return MPI_SUCCESS;
}
static inline int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype st, void *r, const int *rc,
                                const int *rd, MPI_Datatype rt, MPI_Comm c) {
  (void)rc; (void)c;
  shim_cp((char *)r + rd[0] * rt, (const char *)s + sd[0] * st, sc[0] * st);
  return MPI_SUCCESS;
}
/* with one rank both partners are MPI_PROC_NULL: nothing moves */
static inline int MPI_Sendrecv(const void *s, int sn, MPI_Datatype st, int dst, int stag, void *r, int rn,
                               MPI_Datatype rt, int src, int rtag, MPI_Comm c, MPI_Status *status) {
  (void)s; (void)sn; (void)st; (void)dst; (void)stag; (void)r; (void)rn; (void)rt; (void)src; (void)rtag; (void)c;
  (void)status; return MPI_SUCCESS;
}
static inline int MPI_File_open(MPI_Comm c, const char *path, int mode, MPI_Info info, MPI_File *fp) {
  (void)c; (void)mode; (void)info; *fp = fopen(path, "wb"); return *fp ? MPI_SUCCESS : 1;
}
static inline int MPI_File_write_at_all(MPI_File fp, MPI_Offset off, const void *buf, int n, MPI_Datatype t,
                                        MPI_Status *st) {
  (void)st;
  if (fseek(fp, (long)off, SEEK_SET) != 0) return 1;
  return fwrite(buf, (size_t)t, (size_t)n, fp) == (size_t)n ? MPI_SUCCESS : 1;
}
static inline int MPI_File_close(MPI_File *fp) { int e = fclose(*fp); *fp = NULL; return e ? 1 : MPI_SUCCESS; }
#endif
