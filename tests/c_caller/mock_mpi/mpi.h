/* TEST DOUBLE of a "real" MPI installation with Open MPI's handle convention: MPI_Comm is a POINTER to an opaque
 * struct, so nothing that assumes integer handles (as this repo's own mpi_shim does) can accidentally work. Implements
 * just what include/cudecomp_b200_mpi.h and the test caller need; ranks come from MOCK_MPI_RANK / MOCK_MPI_SIZE and the
 * broadcast goes through files in MOCK_MPI_DIR (tests/c_caller/mock_mpi/mock_mpi.c). */
#ifndef MOCK_REAL_MPI_H
#define MOCK_REAL_MPI_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mock_mpi_communicator_t* MPI_Comm;
typedef struct mock_mpi_datatype_t* MPI_Datatype;
typedef int MPI_Fint;

extern struct mock_mpi_communicator_t mock_mpi_comm_world;
extern struct mock_mpi_datatype_t mock_mpi_char, mock_mpi_int;

#define MPI_COMM_WORLD (&mock_mpi_comm_world)
#define MPI_CHAR (&mock_mpi_char)
#define MPI_INT (&mock_mpi_int)
#define MPI_SUCCESS 0
#define MPI_MAX_PROCESSOR_NAME 256

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Bcast(void* buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);
int MPI_Get_processor_name(char* name, int* resultlen);

#ifdef __cplusplus
}
#endif

#endif
