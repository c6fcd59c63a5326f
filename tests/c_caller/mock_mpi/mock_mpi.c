/* See mpi.h in this directory: a test double, not an MPI. */
#include "mpi.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

struct mock_mpi_communicator_t {
  int rank, size, seq;
};
struct mock_mpi_datatype_t {
  int bytes;
};
struct mock_mpi_communicator_t mock_mpi_comm_world = {0, 1, 0};
struct mock_mpi_datatype_t mock_mpi_char = {1}, mock_mpi_int = {4};
int mock_mpi_calls = 0; /* how often this library's functions were entered: proves nobody shadowed them */

static const char* dir(void) {
  const char* d = getenv("MOCK_MPI_DIR");
  return d ? d : "/tmp";
}

int MPI_Init(int* argc, char*** argv) {
  (void)argc;
  (void)argv;
  ++mock_mpi_calls;
  mock_mpi_comm_world.rank = getenv("MOCK_MPI_RANK") ? atoi(getenv("MOCK_MPI_RANK")) : 0;
  mock_mpi_comm_world.size = getenv("MOCK_MPI_SIZE") ? atoi(getenv("MOCK_MPI_SIZE")) : 1;
  return MPI_SUCCESS;
}
int MPI_Finalize(void) {
  ++mock_mpi_calls;
  return MPI_SUCCESS;
}
int MPI_Comm_rank(MPI_Comm comm, int* rank) {
  ++mock_mpi_calls;
  *rank = comm->rank;
  return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm comm, int* size) {
  ++mock_mpi_calls;
  *size = comm->size;
  return MPI_SUCCESS;
}
int MPI_Get_processor_name(char* name, int* resultlen) {
  ++mock_mpi_calls;
  strcpy(name, "127.0.0.1");
  *resultlen = (int)strlen(name);
  return MPI_SUCCESS;
}
int MPI_Bcast(void* buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm) {
  char path[512], tmp[600];
  const size_t bytes = (size_t)count * (size_t)datatype->bytes;
  int tries;
  ++mock_mpi_calls;
  snprintf(path, sizeof(path), "%s/bcast_%d", dir(), comm->seq++);
  if (comm->size == 1) return MPI_SUCCESS;
  if (comm->rank == root) {
    FILE* f;
    snprintf(tmp, sizeof(tmp), "%s.tmp", path);
    f = fopen(tmp, "wb");
    if (!f || fwrite(buffer, 1, bytes, f) != bytes) return 1;
    fclose(f);
    return rename(tmp, path) == 0 ? MPI_SUCCESS : 1;
  }
  for (tries = 0; tries < 6000; ++tries) {
    FILE* f = fopen(path, "rb");
    if (f) {
      const size_t got = fread(buffer, 1, bytes, f);
      fclose(f);
      return got == bytes ? MPI_SUCCESS : 1;
    }
    usleep(10000);
  }
  return 1;
}
int MPI_Barrier(MPI_Comm comm) {
  int token = 0;
  ++mock_mpi_calls;
  return MPI_Bcast(&token, 1, MPI_INT, 0, comm); /* good enough for the test: rank 0 arrives first */
}
