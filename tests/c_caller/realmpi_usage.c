/* An application that runs on a "real" MPI (here: the pointer-handle test double in mock_mpi/) and calls cuDecomp
 * exactly as the reference's examples do (reference examples/cc/basic_usage/basic_usage.cu): its source contains no
 * trace of this repo. It is compiled with `-include cudecomp_b200_mpi.h` and linked with libcudecomp_realmpi.so plus
 * the MPI it brought; see tests/test_c_caller.py and INTEGRATION.md. Host-only (geometry queries), 1 or 2 ranks. */
#include <stdio.h>

#include <mpi.h>

#include <cudecomp.h>

extern int mock_mpi_calls;

int main(int argc, char** argv) {
  int rank, size;
  cudecompHandle_t handle;
  cudecompGridDescConfig_t config;
  cudecompGridDesc_t grid_desc;
  cudecompPencilInfo_t px, py;
  cudecompResult_t res;

  MPI_Init(&argc, &argv);
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &size);

  res = cudecompInit(&handle, MPI_COMM_WORLD);
  if (res != CUDECOMP_RESULT_SUCCESS) {
    printf("rank %d: cudecompInit failed with %d\n", rank, (int)res);
    return 1;
  }
  cudecompGridDescConfigSetDefaults(&config);
  config.gdims[0] = 8;
  config.gdims[1] = 6;
  config.gdims[2] = 4;
  config.pdims[0] = size;
  config.pdims[1] = 1;
  res = cudecompGridDescCreate(handle, &grid_desc, &config, NULL);
  if (res != CUDECOMP_RESULT_SUCCESS) {
    printf("rank %d: cudecompGridDescCreate failed with %d\n", rank, (int)res);
    return 1;
  }
  cudecompGetPencilInfo(handle, grid_desc, &px, 0, NULL, NULL);
  cudecompGetPencilInfo(handle, grid_desc, &py, 1, NULL, NULL);
  /* X pencil [8, 6/size, 4], Y pencil [8/size, 6, 4]; this rank's slab starts at rank * extent */
  if (px.shape[0] != 8 || px.shape[1] != 6 / size || px.shape[2] != 4 || px.lo[1] != rank * (6 / size) ||
      py.shape[0] != 8 / size || py.shape[1] != 6 || py.lo[0] != rank * (8 / size)) {
    printf("rank %d: unexpected pencil info\n", rank);
    return 1;
  }
  cudecompGridDescDestroy(handle, grid_desc);
  cudecompFinalize(handle);
  MPI_Barrier(MPI_COMM_WORLD);
  MPI_Finalize();
  /* MPI_Init, rank, size (app) + rank, size, [processor name], bcast (adapter) + barrier's bcast, finalize */
  if (mock_mpi_calls < 8) {
    printf("rank %d: the application's MPI was shadowed (%d calls seen)\n", rank, mock_mpi_calls);
    return 1;
  }
  printf("real-MPI caller OK rank %d of %d\n", rank, size);
  return 0;
}
