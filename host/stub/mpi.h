/* Single-process stand-in for <mpi.h>: enough of MPI, with the semantics of a ONE-rank world, to compile
 * host/ against the reference's headers AND to run nix::Application::main() here, where no MPI exists
 * (tests/test_host_cpp.py).  Collectives are the identity on one rank, point-to-point goes nowhere
 * (a one-rank balancer never ships a chunk), MPI-IO is POSIX pread/pwrite for the contiguous calls the
 * checkpoint path uses.  A real nix application builds host/ with its own MPI; nothing in libnixb200.so
 * depends on this file. */
#ifndef NIXB200_HOST_STUB_MPI_H
#define NIXB200_HOST_STUB_MPI_H
#include "../../oracle/stub/mpi.h"

#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>

#ifdef __cplusplus
extern "C" {
#endif
typedef struct nixstub_file {
  int  fd;
  long pos;
}* MPI_File;
typedef int MPI_Group;
#define MPI_IN_PLACE ((void*)1)
#define MPI_INFO_NULL 0
#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_SUM 3
#define MPI_LAND 4
#define MPI_ORDER_C 0
#define MPI_ORDER_FORTRAN 1
#define MPI_MODE_CREATE 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_RDONLY 4
#define MPI_MODE_RDWR 8
#define MPI_MODE_APPEND 16
#define MPI_SEEK_SET 0
#define MPI_COMM_TYPE_SHARED 0
#define MPI_UNDEFINED (-32766)
#define MPI_DATATYPE_NULL 0
#define MPI_FILE_NULL ((MPI_File)0)
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_CXX_BOOL 7
#define MPI_ERR_OTHER 15

/* datatypes: the predefined ones by size; MPI_Type_contiguous(n, BYTE) is encoded as 1000 + n */
static inline long nixstub_type_bytes(MPI_Datatype t)
{
  switch (t) {
  case MPI_BYTE: case MPI_CHAR: case MPI_CXX_BOOL: return 1;
  case MPI_INT: case MPI_FLOAT: return 4;
  case MPI_DOUBLE: case MPI_INT64_T: return 8;
  default: return t >= 1000 ? (long)t - 1000 : 1;
  }
}
static inline void nixstub_copy(const void* s, void* r, long bytes) { if (s != MPI_IN_PLACE && s != r && bytes > 0) memcpy(r, s, (size_t)bytes); }

static inline int MPI_Initialized(int* f) { *f = 1; return 0; } /* "already initialised": nobody calls MPI_Init/Finalize */
static inline int MPI_Query_thread(int* p) { *p = MPI_THREAD_MULTIPLE; return 0; }
static inline int MPI_Init_thread(int* a, char*** b, int req, int* prov) { (void)a;(void)b; if (prov) *prov = req; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Abort(MPI_Comm c, int e) { (void)c; fprintf(stderr, "MPI_Abort(%d)\n", e); abort(); return e; }
static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int* n) { (void)c; *n = 1; return 0; }
static inline int MPI_Comm_dup(MPI_Comm c, MPI_Comm* d) { *d = c; return 0; }
static inline int MPI_Comm_free(MPI_Comm* c) { (void)c; return 0; }
static inline int MPI_Comm_split(MPI_Comm c, int a, int b, MPI_Comm* d) { (void)a;(void)b; *d = c; return 0; }
static inline int MPI_Comm_split_type(MPI_Comm c, int a, int b, MPI_Info i, MPI_Comm* d) { (void)a;(void)b;(void)i; *d = c; return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline int MPI_Wait(MPI_Request* r, MPI_Status* s) { (void)r;(void)s; return 0; }
static inline int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b;(void)n;(void)t;(void)root;(void)c; return 0; }
static inline int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c) { (void)o;(void)c; nixstub_copy(s, r, n * nixstub_type_bytes(t)); return 0; }
static inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op o, int root, MPI_Comm c) { (void)o;(void)root;(void)c; nixstub_copy(s, r, n * nixstub_type_bytes(t)); return 0; }
static inline int MPI_Allgather(const void* s, int n, MPI_Datatype t, void* r, int rn, MPI_Datatype rt, MPI_Comm c) { (void)rn;(void)rt;(void)c; nixstub_copy(s, r, n * nixstub_type_bytes(t)); return 0; }
static inline int MPI_Allgatherv(const void* s, int n, MPI_Datatype t, void* r, const int* rc, const int* d, MPI_Datatype rt, MPI_Comm c) { (void)rc;(void)rt;(void)c; nixstub_copy(s, (char*)r + (d ? d[0] : 0) * nixstub_type_bytes(rt), n * nixstub_type_bytes(t)); return 0; }
static inline int MPI_Gather(const void* s, int n, MPI_Datatype t, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) { (void)rn;(void)rt;(void)root;(void)c; nixstub_copy(s, r, n * nixstub_type_bytes(t)); return 0; }
static inline int MPI_Gatherv(const void* s, int n, MPI_Datatype t, void* r, const int* rc, const int* d, MPI_Datatype rt, int root, MPI_Comm c) { (void)rc;(void)root;(void)c; nixstub_copy(s, (char*)r + (d ? d[0] : 0) * nixstub_type_bytes(rt), n * nixstub_type_bytes(t)); return 0; }
static inline int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) { (void)b;(void)n;(void)t;(void)dst;(void)tag;(void)c; return 0; }
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* s) { (void)b;(void)n;(void)t;(void)src;(void)tag;(void)c;(void)s; return 0; }
static inline double MPI_Wtime(void) { struct timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + 1e-6 * tv.tv_usec; }

static inline int MPI_Type_contiguous(int n, MPI_Datatype t, MPI_Datatype* out) { *out = (MPI_Datatype)(1000 + n * nixstub_type_bytes(t)); return 0; }
static inline int MPI_Type_commit(MPI_Datatype* t) { (void)t; return 0; }
static inline int MPI_Type_free(MPI_Datatype* t) { *t = MPI_DATATYPE_NULL; return 0; }
/* file views over derived types are not modelled (the collective subarray writers of nixio.cpp): they fail */
static inline int MPI_Type_create_hindexed(int n, const int* bl, const MPI_Aint* d, MPI_Datatype t, MPI_Datatype* out) { (void)n;(void)bl;(void)d;(void)t; *out = MPI_DATATYPE_NULL; return MPI_ERR_OTHER; }
static inline int MPI_Type_create_subarray(int n, const int* a, const int* b, const int* c2, int o, MPI_Datatype t, MPI_Datatype* out) { (void)n;(void)a;(void)b;(void)c2;(void)o;(void)t; *out = MPI_DATATYPE_NULL; return MPI_ERR_OTHER; }
static inline int MPI_File_set_view(MPI_File f, MPI_Offset d, MPI_Datatype e, MPI_Datatype ft, const char* rep, MPI_Info i) { (void)f;(void)d;(void)e;(void)ft;(void)rep;(void)i; return MPI_ERR_OTHER; }
static inline int MPI_File_iread_all(MPI_File f, void* b, int n, MPI_Datatype t, MPI_Request* r) { (void)f;(void)b;(void)n;(void)t;(void)r; return MPI_ERR_OTHER; }
static inline int MPI_File_iwrite_all(MPI_File f, const void* b, int n, MPI_Datatype t, MPI_Request* r) { (void)f;(void)b;(void)n;(void)t;(void)r; return MPI_ERR_OTHER; }

static inline int MPI_File_open(MPI_Comm c, const char* name, int amode, MPI_Info i, MPI_File* fh)
{
  (void)c;(void)i;
  int flags = 0;
  if (amode & MPI_MODE_RDWR) flags |= O_RDWR;
  else if (amode & MPI_MODE_WRONLY) flags |= O_WRONLY;
  else flags |= O_RDONLY;
  if (amode & MPI_MODE_CREATE) flags |= O_CREAT;
  int fd = open(name, flags, 0644);
  if (fd < 0) { *fh = MPI_FILE_NULL; return MPI_ERR_OTHER; }
  *fh = (MPI_File)malloc(sizeof(struct nixstub_file));
  (*fh)->fd = fd;
  (*fh)->pos = 0;
  return 0;
}
static inline int MPI_File_close(MPI_File* fh) { if (*fh) { close((*fh)->fd); free(*fh); *fh = MPI_FILE_NULL; } return 0; }
static inline int MPI_File_delete(const char* name, MPI_Info i) { (void)i; return unlink(name) == 0 ? 0 : MPI_ERR_OTHER; }
static inline int MPI_File_seek(MPI_File fh, MPI_Offset off, int whence) { (void)whence; fh->pos = off; return 0; }
static inline int MPI_File_get_position(MPI_File fh, MPI_Offset* off) { *off = fh->pos; return 0; }
static inline int MPI_File_get_size(MPI_File fh, MPI_Offset* size) { struct stat st; if (fstat(fh->fd, &st)) return MPI_ERR_OTHER; *size = st.st_size; return 0; }
static inline int MPI_File_iread_at(MPI_File fh, MPI_Offset off, void* b, int n, MPI_Datatype t, MPI_Request* r) { if (r) *r = MPI_REQUEST_NULL; long bytes = n * nixstub_type_bytes(t); return pread(fh->fd, b, (size_t)bytes, off) == bytes ? 0 : MPI_ERR_OTHER; }
static inline int MPI_File_iwrite_at(MPI_File fh, MPI_Offset off, const void* b, int n, MPI_Datatype t, MPI_Request* r) { if (r) *r = MPI_REQUEST_NULL; long bytes = n * nixstub_type_bytes(t); return pwrite(fh->fd, b, (size_t)bytes, off) == bytes ? 0 : MPI_ERR_OTHER; }
#ifdef __cplusplus
}
#endif
#endif
