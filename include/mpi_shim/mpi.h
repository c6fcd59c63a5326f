/*
 * mpi.h -- bootstrap shim for machines without an MPI installation.
 *
 * cudecomp.h takes an MPI_Comm by value (reference include/cudecomp.h:31,249), so
 * a caller cannot even compile without <mpi.h>.  This header plus the functions of
 * the same names exported by libcudecomp.so give callers the MPI subset that the
 * reference's own callers use (tests/cc/transpose_test.cc:569-669,
 * benchmark/benchmark.cu) on top of the library's TCP bootstrap: ranks come from
 * the torchrun-style environment (RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT).
 *
 * Only small host-side control messages travel this way.  Pencil data never does:
 * it moves GPU to GPU over NVLink inside the transpose kernels.
 *
 * Handles are plain ints (MPICH convention), so MPI_Fint conversion is the identity.
 */
#ifndef CUDECOMP_B200_MPI_SHIM_H
#define CUDECOMP_B200_MPI_SHIM_H

#include <stddef.h>
#include <stdint.h>

#define CUDECOMP_B200_MPI_SHIM 1

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Fint;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef struct {
  int MPI_SOURCE, MPI_TAG, MPI_ERROR;
} MPI_Status;

#define MPI_SUCCESS 0
#define MPI_ERR_OTHER 15

#define MPI_COMM_NULL ((MPI_Comm)0)
#define MPI_COMM_WORLD ((MPI_Comm)1)
#define MPI_COMM_SELF ((MPI_Comm)2)
#define MPI_INFO_NULL ((MPI_Info)0)
#define MPI_COMM_TYPE_SHARED 1
#define MPI_UNDEFINED (-32766)
#define MPI_IN_PLACE ((void*)-1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_MAX_ERROR_STRING 256
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3

/* datatype = (kind << 8) | size_in_bytes; kind: 1 signed int, 2 unsigned int, 3 float, 4 complex, 5 raw */
#define CUDECOMP_SHIM_DT(kind, size) (((kind) << 8) | (size))
#define MPI_CHAR CUDECOMP_SHIM_DT(1, 1)
#define MPI_SIGNED_CHAR CUDECOMP_SHIM_DT(1, 1)
#define MPI_SHORT CUDECOMP_SHIM_DT(1, 2)
#define MPI_INT CUDECOMP_SHIM_DT(1, 4)
#define MPI_INT32_T CUDECOMP_SHIM_DT(1, 4)
#define MPI_LONG CUDECOMP_SHIM_DT(1, 8)
#define MPI_LONG_LONG CUDECOMP_SHIM_DT(1, 8)
#define MPI_LONG_LONG_INT CUDECOMP_SHIM_DT(1, 8)
#define MPI_INT64_T CUDECOMP_SHIM_DT(1, 8)
#define MPI_UNSIGNED_CHAR CUDECOMP_SHIM_DT(2, 1)
#define MPI_UINT8_T CUDECOMP_SHIM_DT(2, 1)
#define MPI_C_BOOL CUDECOMP_SHIM_DT(2, 1)
#define MPI_UNSIGNED CUDECOMP_SHIM_DT(2, 4)
#define MPI_UINT32_T CUDECOMP_SHIM_DT(2, 4)
#define MPI_UNSIGNED_LONG CUDECOMP_SHIM_DT(2, 8)
#define MPI_UNSIGNED_LONG_LONG CUDECOMP_SHIM_DT(2, 8)
#define MPI_UINT64_T CUDECOMP_SHIM_DT(2, 8)
#define MPI_FLOAT CUDECOMP_SHIM_DT(3, 4)
#define MPI_DOUBLE CUDECOMP_SHIM_DT(3, 8)
#define MPI_C_FLOAT_COMPLEX CUDECOMP_SHIM_DT(4, 8)
#define MPI_C_DOUBLE_COMPLEX CUDECOMP_SHIM_DT(4, 16)
#define MPI_BYTE CUDECOMP_SHIM_DT(5, 1)

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_LOR 4
#define MPI_LAND 5
#define MPI_BOR 6
#define MPI_PROD 7

int MPI_Init(int* argc, char*** argv);
int MPI_Init_thread(int* argc, char*** argv, int required, int* provided);
int MPI_Initialized(int* flag);
int MPI_Finalize(void);
int MPI_Finalized(int* flag);
int MPI_Abort(MPI_Comm comm, int errorcode);
double MPI_Wtime(void);
int MPI_Get_processor_name(char* name, int* resultlen);
int MPI_Error_string(int errorcode, char* string, int* resultlen);

int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm* newcomm);
int MPI_Comm_split_type(MPI_Comm comm, int split_type, int key, MPI_Info info, MPI_Comm* newcomm);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm* newcomm);
int MPI_Comm_free(MPI_Comm* comm);
MPI_Fint MPI_Comm_c2f(MPI_Comm comm);
MPI_Comm MPI_Comm_f2c(MPI_Fint comm);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void* buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm);
int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
                  MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Gather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
               MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int root,
               MPI_Comm comm);

#ifdef __cplusplus
}
#endif

#endif
