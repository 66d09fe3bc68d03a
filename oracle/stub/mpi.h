/* Stub <mpi.h> used ONLY to compile the reference's headers into the CPU oracle
 * (oracle/_ref).  The hot path never communicates through MPI in the oracle: halo
 * buffers are moved between chunks by memcpy ("loopback"), exactly as the reference's
 * own halo tests do (unittest/test_xtensor_halo3d.cpp:85-93).  Every function is a no-op.
 * TEST INFRASTRUCTURE - not part of the product. */
#ifndef NIXB200_ORACLE_STUB_MPI_H
#define NIXB200_ORACLE_STUB_MPI_H
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_Datatype;
typedef int MPI_Comm;
typedef int MPI_Request;
typedef int MPI_Op;
typedef int MPI_Info;
typedef long MPI_Offset;
typedef long MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count_; } MPI_Status;
#define MPI_SUCCESS 0
#define MPI_BYTE 1
#define MPI_CHAR 2
#define MPI_INT 3
#define MPI_FLOAT 4
#define MPI_DOUBLE 5
#define MPI_INT64_T 6
#define MPI_PROC_NULL (-2)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_REQUEST_NULL (-1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3
static inline int MPI_Isend(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* r) { (void)b;(void)n;(void)t;(void)dst;(void)tag;(void)c; if (r) *r = MPI_REQUEST_NULL; return 0; }
static inline int MPI_Irecv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* r) { (void)b;(void)n;(void)t;(void)src;(void)tag;(void)c; if (r) *r = MPI_REQUEST_NULL; return 0; }
static inline int MPI_Waitall(int n, MPI_Request* r, MPI_Status* s) { (void)n;(void)r;(void)s; return 0; }
static inline int MPI_Testall(int n, MPI_Request* r, int* flag, MPI_Status* s) { (void)n;(void)r;(void)s; if (flag) *flag = 1; return 0; }
static inline int MPI_Iprobe(int src, int tag, MPI_Comm c, int* flag, MPI_Status* s) { (void)src;(void)tag;(void)c;(void)s; if (flag) *flag = 0; return 0; }
static inline int MPI_Get_count(const MPI_Status* s, MPI_Datatype t, int* n) { (void)s;(void)t; if (n) *n = 0; return 0; }
static inline int MPI_Type_size(MPI_Datatype t, int* n) { (void)t; if (n) *n = 1; return 0; }
#ifdef __cplusplus
}
#endif
#endif
