/* A plain C (not C++) caller of the library, written against include/cudecomp.h only: proves that the header is a
 * valid C header with the reference's struct sizes and that the exported symbols link and behave from C.
 * Runs on one rank without touching the GPU data path (geometry queries + argument checking).
 *   gcc -std=c11 -Iinclude -Iinclude/mpi_shim -I/usr/local/cuda/include tests/c_caller/basic_usage.c \
 *       -Lcudecomp_b200/lib -lcudecomp -Wl,-rpath,$PWD/cudecomp_b200/lib -o basic_usage_c */
#include <stdio.h>
#include <string.h>

#include <cudecomp.h>

_Static_assert(sizeof(cudecompGridDescConfig_t) == 104, "config ABI size (reference src/cudecomp.cc:216)");
_Static_assert(sizeof(cudecompGridDescAutotuneOptions_t) == 320, "options ABI size (reference src/cudecomp.cc:242)");
_Static_assert(sizeof(cudecompPencilInfo_t) == 96, "pencil info ABI size (reference src/cudecomp.cc:268)");

#define CHECK(call, want)                                                                                              \
  do {                                                                                                                 \
    cudecompResult_t r_ = (call);                                                                                      \
    if (r_ != (want)) {                                                                                                \
      printf("FAILED %s: got %d, expected %d\n", #call, (int)r_, (int)(want));                                          \
      return 1;                                                                                                        \
    }                                                                                                                  \
  } while (0)

int main(void) {
  if (MPI_Init(NULL, NULL) != MPI_SUCCESS) return 1;
  int rank = -1, size = -1;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  if (rank != 0 || size != 1) return 1;

  cudecompHandle_t handle;
  CHECK(cudecompInit(&handle, MPI_COMM_WORLD), CUDECOMP_RESULT_SUCCESS);

  cudecompGridDescConfig_t config;
  memset(&config, 0xff, sizeof(config));
  CHECK(cudecompGridDescConfigSetDefaults(&config), CUDECOMP_RESULT_SUCCESS);
  if (config.transpose_comm_backend != CUDECOMP_TRANSPOSE_COMM_MPI_P2P || config.transpose_mem_order[2][2] != -1) return 1;
  config.gdims[0] = 9;
  config.gdims[1] = 10;
  config.gdims[2] = 11;
  config.pdims[0] = 1;
  config.pdims[1] = 1;
  config.transpose_axis_contiguous[2] = true;

  cudecompGridDesc_t grid_desc;
  CHECK(cudecompGridDescCreate(handle, &grid_desc, &config, NULL), CUDECOMP_RESULT_SUCCESS);

  cudecompPencilInfo_t pinfo;
  const int32_t halo[3] = {1, 2, 1}, pad[3] = {1, 0, 2};
  CHECK(cudecompGetPencilInfo(handle, grid_desc, &pinfo, 2, halo, pad), CUDECOMP_RESULT_SUCCESS);
  /* z pencils, axis-contiguous: memory order (z, x, y) */
  if (pinfo.order[0] != 2 || pinfo.order[1] != 0 || pinfo.order[2] != 1) return 1;
  if (pinfo.shape[0] != 11 + 2 + 2 || pinfo.shape[1] != 9 + 2 + 1 || pinfo.shape[2] != 10 + 4) return 1;
  if (pinfo.size != 15LL * 12 * 14 || pinfo.magic != CUDECOMP_PENCIL_INFO_MAGIC) return 1;

  int64_t work = 0, dsize = 0;
  CHECK(cudecompGetTransposeWorkspaceSize(handle, grid_desc, &work), CUDECOMP_RESULT_SUCCESS);
  if (work != 1024 + 990) return 1; /* roundup64(990) + 990 */
  CHECK(cudecompGetDataTypeSize(CUDECOMP_DOUBLE_COMPLEX, &dsize), CUDECOMP_RESULT_SUCCESS);
  if (dsize != 16) return 1;
  int32_t nb = 7;
  CHECK(cudecompGetShiftedRank(handle, grid_desc, 0, 1, 1, true, &nb), CUDECOMP_RESULT_SUCCESS);
  if (nb != 0) return 1;
  if (strcmp(cudecompTransposeCommBackendToString(CUDECOMP_TRANSPOSE_COMM_NVSHMEM_PL), "NVSHMEM (pipelined)") != 0) return 1;

  /* argument checking happens before anything touches a pointer (reference tests/ctest/api_tests.cc:1468-1505) */
  CHECK(cudecompTransposeXToY(handle, grid_desc, NULL, NULL, NULL, CUDECOMP_FLOAT, NULL, NULL, NULL, NULL, 0),
        CUDECOMP_RESULT_INVALID_USAGE);
  CHECK(cudecompUpdateHalosX(handle, grid_desc, NULL, NULL, CUDECOMP_FLOAT, NULL, NULL, 0, NULL, 0),
        CUDECOMP_RESULT_INVALID_USAGE);

  CHECK(cudecompGridDescDestroy(handle, grid_desc), CUDECOMP_RESULT_SUCCESS);
  CHECK(cudecompFinalize(handle), CUDECOMP_RESULT_SUCCESS);
  MPI_Finalize();
  printf("C caller OK\n");
  return 0;
}
