/* Single-rank stand-in for <mpi.h>, used ONLY to compile-check host/ against the reference's
 * headers in a container without MPI (tests/test_host_cpp.py).  A real nix application builds
 * host/ with its own MPI; nothing in libnixb200.so depends on this file. */
#ifndef NIXB200_HOST_STUB_MPI_H
#define NIXB200_HOST_STUB_MPI_H
#include "../../oracle/stub/mpi.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_File;
typedef int MPI_Group;
#define MPI_IN_PLACE ((void*)1)
#define MPI_INFO_NULL 0
#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_SUM 3
#define MPI_LAND 4
#define MPI_ORDER_C 0
#define MPI_MODE_CREATE 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_RDONLY 4
#define MPI_MODE_RDWR 8
#define MPI_MODE_APPEND 16
#define MPI_SEEK_SET 0
#define MPI_COMM_TYPE_SHARED 0
#define MPI_UNDEFINED (-32766)
#define MPI_DATATYPE_NULL 0
#define MPI_FILE_NULL 0
#define MPI_MAX_PROCESSOR_NAME 256
static inline int MPI_Init_thread(int* a, char*** b, int req, int* prov) { (void)a;(void)b; if (prov) *prov = req; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Abort(MPI_Comm c, int e) { (void)c; __builtin_trap(); return e; }
static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int* n) { (void)c; *n = 1; return 0; }
static inline int MPI_Comm_dup(MPI_Comm c, MPI_Comm* d) { *d = c; return 0; }
static inline int MPI_Comm_free(MPI_Comm* c) { (void)c; return 0; }
static inline int MPI_Comm_split(MPI_Comm c, int a, int b, MPI_Comm* d) { (void)a;(void)b; *d = c; return 0; }
static inline int MPI_Comm_split_type(MPI_Comm c, int a, int b, MPI_Info i, MPI_Comm* d) { (void)a;(void)b;(void)i; *d = c; return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline int MPI_Wait(MPI_Request* r, MPI_Status* s) { (void)r;(void)s; return 0; }
static inline int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b;(void)n;(void)t;(void)root;(void)c; return 0; }
static inline int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c) { (void)s;(void)r;(void)n;(void)t;(void)o;(void)c; return 0; }
static inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op o, int root, MPI_Comm c) { (void)s;(void)r;(void)n;(void)t;(void)o;(void)root;(void)c; return 0; }
static inline int MPI_Allgatherv(const void* s, int n, MPI_Datatype t, void* r, const int* rc, const int* d, MPI_Datatype rt, MPI_Comm c) { (void)s;(void)n;(void)t;(void)r;(void)rc;(void)d;(void)rt;(void)c; return 0; }
static inline int MPI_Gather(const void* s, int n, MPI_Datatype t, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) { (void)s;(void)n;(void)t;(void)r;(void)rn;(void)rt;(void)root;(void)c; return 0; }
static inline int MPI_Gatherv(const void* s, int n, MPI_Datatype t, void* r, const int* rc, const int* d, MPI_Datatype rt, int root, MPI_Comm c) { (void)s;(void)n;(void)t;(void)r;(void)rc;(void)d;(void)rt;(void)root;(void)c; return 0; }
static inline int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) { (void)b;(void)n;(void)t;(void)dst;(void)tag;(void)c; return 0; }
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* s) { (void)b;(void)n;(void)t;(void)src;(void)tag;(void)c;(void)s; return 0; }
static inline double MPI_Wtime(void) { return 0.0; }
#ifdef __cplusplus
}
#endif
#endif
