/*
 * cudecomp.h -- C ABI of the B200-native pencil-transpose engine.
 *
 * This header is the drop-in boundary.  Every enum value, struct layout and
 * function signature below is binary compatible with the reference header
 * (reference include/cudecomp.h, v0.7.0); the line of the reference declaration
 * each item replaces is cited as "ref:<line>".  Callers written against the
 * reference (C, C++, or Fortran through src/cudecomp_m.cuf:206-531) link against
 * this library unchanged.
 *
 * <mpi.h> is whatever MPI the caller builds with.  When no MPI exists on the
 * machine, put include/mpi_shim on the include path: it provides the small MPI
 * subset callers of this API use, implemented by the library's own bootstrap.
 */
#ifndef CUDECOMP_H
#define CUDECOMP_H

#include <stdbool.h>
#include <stdint.h>

#include <cuda_runtime.h>
#include <mpi.h>

#include "cudecomp_version.h"

/* struct tags checked by every *Versioned entry point (ref:36-38) */
#define CUDECOMP_GRID_DESC_CONFIG_MAGIC INT32_C(0x434f4e46)
#define CUDECOMP_GRID_DESC_AUTOTUNE_OPTIONS_MAGIC INT32_C(0x4155544f)
#define CUDECOMP_PENCIL_INFO_MAGIC INT32_C(0x50494e46)

#ifdef __cplusplus
extern "C" {
#endif

/* ref:48-57.  The values are kept for ABI compatibility.  In this engine every
 * value selects the same NVSwitch peer-store transport; the value picks the
 * schedule variant (see DESIGN.md "backend values"). */
typedef enum {
  CUDECOMP_TRANSPOSE_COMM_MPI_P2P = 1,
  CUDECOMP_TRANSPOSE_COMM_MPI_P2P_PL = 2,
  CUDECOMP_TRANSPOSE_COMM_MPI_A2A = 3,
  CUDECOMP_TRANSPOSE_COMM_NCCL = 4,
  CUDECOMP_TRANSPOSE_COMM_NCCL_PL = 5,
  CUDECOMP_TRANSPOSE_COMM_NVSHMEM = 6,
  CUDECOMP_TRANSPOSE_COMM_NVSHMEM_PL = 7,
  CUDECOMP_TRANSPOSE_COMM_NVSHMEM_SM = 8
} cudecompTransposeCommBackend_t;

/* ref:62-68 */
typedef enum {
  CUDECOMP_HALO_COMM_MPI = 1,
  CUDECOMP_HALO_COMM_MPI_BLOCKING = 2,
  CUDECOMP_HALO_COMM_NCCL = 3,
  CUDECOMP_HALO_COMM_NVSHMEM = 4,
  CUDECOMP_HALO_COMM_NVSHMEM_BLOCKING = 5
} cudecompHaloCommBackend_t;

/* ref:73-78.  Element sizes 4 / 8 / 8 / 16 bytes. */
typedef enum {
  CUDECOMP_FLOAT = -1,
  CUDECOMP_DOUBLE = -2,
  CUDECOMP_FLOAT_COMPLEX = -3,
  CUDECOMP_DOUBLE_COMPLEX = -4
} cudecompDataType_t;

/* ref:83-86 */
typedef enum { CUDECOMP_AUTOTUNE_GRID_TRANSPOSE = 0, CUDECOMP_AUTOTUNE_GRID_HALO = 1 } cudecompAutotuneGridMode_t;

/* ref:91-96 */
typedef enum {
  CUDECOMP_RANK_ORDER_DEFAULT = 0,
  CUDECOMP_RANK_ORDER_ROW_MAJOR = 1,
  CUDECOMP_RANK_ORDER_COL_MAJOR = 2
} cudecompRankOrder_t;

/* ref:102-113.  CUTENSOR/NCCL/NVSHMEM/NVML codes are never produced here but keep their values. */
typedef enum {
  CUDECOMP_RESULT_SUCCESS = 0,
  CUDECOMP_RESULT_INVALID_USAGE = 1,
  CUDECOMP_RESULT_NOT_SUPPORTED = 2,
  CUDECOMP_RESULT_INTERNAL_ERROR = 3,
  CUDECOMP_RESULT_CUDA_ERROR = 4,
  CUDECOMP_RESULT_CUTENSOR_ERROR = 5,
  CUDECOMP_RESULT_MPI_ERROR = 6,
  CUDECOMP_RESULT_NCCL_ERROR = 7,
  CUDECOMP_RESULT_NVSHMEM_ERROR = 8,
  CUDECOMP_RESULT_NVML_ERROR = 9
} cudecompResult_t;

/* opaque handles (ref:118,123) */
typedef struct cudecompHandle* cudecompHandle_t;
typedef struct cudecompGridDesc* cudecompGridDesc_t;

/* ref:128-155, 104 bytes */
typedef struct {
  int64_t struct_size;
  int32_t magic;
  int32_t version;

  int32_t gdims[3];      /* global grid */
  int32_t gdims_dist[3]; /* grid used to distribute (0 = gdims); the excess goes to the last populated rank */
  int32_t pdims[2];      /* process grid, {0,0} = autotune */
  cudecompRankOrder_t rank_order;

  cudecompTransposeCommBackend_t transpose_comm_backend;
  bool transpose_axis_contiguous[3];
  int32_t transpose_mem_order[3][3]; /* [axis][memory position] -> global axis; -1 = unset */

  cudecompHaloCommBackend_t halo_comm_backend;
} cudecompGridDescConfig_t;

/* ref:160-219, 320 bytes */
typedef struct {
  int64_t struct_size;
  int32_t magic;
  int32_t version;

  int32_t n_warmup_trials;
  int32_t n_trials;
  cudecompAutotuneGridMode_t grid_mode;
  cudecompDataType_t dtype;
  bool allow_uneven_decompositions;
  bool disable_mpi_backends;
  bool disable_nccl_backends;
  bool disable_nvshmem_backends;
  double skip_threshold;

  bool autotune_transpose_backend;
  bool transpose_use_inplace_buffers[4]; /* order: XY, YZ, ZY, YX */
  double transpose_op_weights[4];
  int32_t transpose_input_halo_extents[4][3];
  int32_t transpose_output_halo_extents[4][3];
  int32_t transpose_input_padding[4][3];
  int32_t transpose_output_padding[4][3];

  bool autotune_halo_backend;
  int32_t halo_extents[3];
  bool halo_periods[3];
  int32_t halo_axis;
  int32_t halo_padding[3];
} cudecompGridDescAutotuneOptions_t;

/* ref:224-238, 96 bytes.  shape/lo/hi/order are indexed by memory position, halo_extents/padding by global axis. */
typedef struct {
  int64_t struct_size;
  int32_t magic;
  int32_t version;

  int32_t shape[3];
  int32_t lo[3];
  int32_t hi[3];
  int32_t order[3];
  int32_t halo_extents[3];
  int32_t padding[3];
  int64_t size;
} cudecompPencilInfo_t;

/* lifecycle (ref:249,259,268) -- collective over the communicator */
cudecompResult_t cudecompInit(cudecompHandle_t* handle, MPI_Comm mpi_comm);
cudecompResult_t cudecompInit_F(cudecompHandle_t* handle, MPI_Fint mpi_comm_f);
cudecompResult_t cudecompFinalize(cudecompHandle_t handle);

/* grid descriptor (ref:272-313); the unversioned names are header-inline exactly as in the reference (ref:296-303) */
cudecompResult_t cudecompGridDescCreateVersioned(cudecompHandle_t handle, cudecompGridDesc_t* grid_desc,
                                                 cudecompGridDescConfig_t* config, int64_t config_struct_size,
                                                 int32_t config_version,
                                                 const cudecompGridDescAutotuneOptions_t* options,
                                                 int64_t options_struct_size, int32_t options_version);

static inline cudecompResult_t cudecompGridDescCreate(cudecompHandle_t handle, cudecompGridDesc_t* grid_desc,
                                                      cudecompGridDescConfig_t* config,
                                                      const cudecompGridDescAutotuneOptions_t* options) {
  return cudecompGridDescCreateVersioned(handle, grid_desc, config, (int64_t)sizeof(cudecompGridDescConfig_t),
                                         CUDECOMP_GRID_DESC_CONFIG_VERSION, options,
                                         options ? (int64_t)sizeof(cudecompGridDescAutotuneOptions_t) : (int64_t)0,
                                         options ? CUDECOMP_GRID_DESC_AUTOTUNE_OPTIONS_VERSION : (int32_t)0);
}

cudecompResult_t cudecompGridDescDestroy(cudecompHandle_t handle, cudecompGridDesc_t grid_desc);

/* defaults (ref:317-354) */
cudecompResult_t cudecompGridDescConfigSetDefaultsVersioned(cudecompGridDescConfig_t* config, int64_t struct_size,
                                                            int32_t version);
static inline cudecompResult_t cudecompGridDescConfigSetDefaults(cudecompGridDescConfig_t* config) {
  return cudecompGridDescConfigSetDefaultsVersioned(config, (int64_t)sizeof(cudecompGridDescConfig_t),
                                                    CUDECOMP_GRID_DESC_CONFIG_VERSION);
}

cudecompResult_t cudecompGridDescAutotuneOptionsSetDefaultsVersioned(cudecompGridDescAutotuneOptions_t* options,
                                                                     int64_t struct_size, int32_t version);
static inline cudecompResult_t cudecompGridDescAutotuneOptionsSetDefaults(cudecompGridDescAutotuneOptions_t* options) {
  return cudecompGridDescAutotuneOptionsSetDefaultsVersioned(options,
                                                             (int64_t)sizeof(cudecompGridDescAutotuneOptions_t),
                                                             CUDECOMP_GRID_DESC_AUTOTUNE_OPTIONS_VERSION);
}

/* geometry queries (ref:358-388) */
cudecompResult_t cudecompGetPencilInfoVersioned(cudecompHandle_t handle, cudecompGridDesc_t grid_desc,
                                                cudecompPencilInfo_t* pencil_info, int64_t pencil_info_struct_size,
                                                int32_t pencil_info_version, int32_t axis, const int32_t halo_extents[],
                                                const int32_t padding[]);
static inline cudecompResult_t cudecompGetPencilInfo(cudecompHandle_t handle, cudecompGridDesc_t grid_desc,
                                                     cudecompPencilInfo_t* pencil_info, int32_t axis,
                                                     const int32_t halo_extents[], const int32_t padding[]) {
  return cudecompGetPencilInfoVersioned(handle, grid_desc, pencil_info, (int64_t)sizeof(cudecompPencilInfo_t),
                                        CUDECOMP_PENCIL_INFO_VERSION, axis, halo_extents, padding);
}

/* workspace sizes in ELEMENTS of the dtype the caller will use (ref:401,420) */
cudecompResult_t cudecompGetTransposeWorkspaceSize(cudecompHandle_t handle, cudecompGridDesc_t grid_desc,
                                                   int64_t* workspace_size);
cudecompResult_t cudecompGetHaloWorkspaceSize(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t axis,
                                              const int32_t halo_extents[], int64_t* workspace_size);

cudecompResult_t cudecompGetDataTypeSize(cudecompDataType_t dtype, int64_t* dtype_size); /* ref:430 */

/* collective workspace allocator (ref:447,462) */
cudecompResult_t cudecompMalloc(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, void** buffer,
                                size_t buffer_size_bytes);
cudecompResult_t cudecompFree(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, void* buffer);

const char* cudecompTransposeCommBackendToString(cudecompTransposeCommBackend_t comm_backend); /* ref:472 */
const char* cudecompHaloCommBackendToString(cudecompHaloCommBackend_t comm_backend);           /* ref:481 */

/* ref:484-501 */
cudecompResult_t cudecompGetGridDescConfigVersioned(cudecompHandle_t handle, cudecompGridDesc_t grid_desc,
                                                    cudecompGridDescConfig_t* config, int64_t struct_size,
                                                    int32_t version);
static inline cudecompResult_t cudecompGetGridDescConfig(cudecompHandle_t handle, cudecompGridDesc_t grid_desc,
                                                         cudecompGridDescConfig_t* config) {
  return cudecompGetGridDescConfigVersioned(handle, grid_desc, config, (int64_t)sizeof(cudecompGridDescConfig_t),
                                            CUDECOMP_GRID_DESC_CONFIG_VERSION);
}

/* ref:517 */
cudecompResult_t cudecompGetShiftedRank(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t axis,
                                        int32_t dim, int32_t displacement, bool periodic, int32_t* shifted_rank);

/* The hot path: pencil-to-pencil global transposes (ref:545,574,603,632).  input == output means in place.
 * Halo extents / padding are in global axis order and may be NULL (= zeros).  All work is enqueued on `stream`. */
cudecompResult_t cudecompTransposeXToY(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, void* input, void* output,
                                       void* work, cudecompDataType_t dtype, const int32_t input_halo_extents[],
                                       const int32_t output_halo_extents[], const int32_t input_padding[],
                                       const int32_t output_padding[], cudaStream_t stream);
cudecompResult_t cudecompTransposeYToZ(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, void* input, void* output,
                                       void* work, cudecompDataType_t dtype, const int32_t input_halo_extents[],
                                       const int32_t output_halo_extents[], const int32_t input_padding[],
                                       const int32_t output_padding[], cudaStream_t stream);
cudecompResult_t cudecompTransposeZToY(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, void* input, void* output,
                                       void* work, cudecompDataType_t dtype, const int32_t input_halo_extents[],
                                       const int32_t output_halo_extents[], const int32_t input_padding[],
                                       const int32_t output_padding[], cudaStream_t stream);
cudecompResult_t cudecompTransposeYToX(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, void* input, void* output,
                                       void* work, cudecompDataType_t dtype, const int32_t input_halo_extents[],
                                       const int32_t output_halo_extents[], const int32_t input_padding[],
                                       const int32_t output_padding[], cudaStream_t stream);

/* Halo exchange of one dimension of an X/Y/Z pencil (ref:661,688,715). */
cudecompResult_t cudecompUpdateHalosX(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, void* input, void* work,
                                      cudecompDataType_t dtype, const int32_t halo_extents[], const bool halo_periods[],
                                      int32_t dim, const int32_t padding[], cudaStream_t stream);
cudecompResult_t cudecompUpdateHalosY(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, void* input, void* work,
                                      cudecompDataType_t dtype, const int32_t halo_extents[], const bool halo_periods[],
                                      int32_t dim, const int32_t padding[], cudaStream_t stream);
cudecompResult_t cudecompUpdateHalosZ(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, void* input, void* work,
                                      cudecompDataType_t dtype, const int32_t halo_extents[], const bool halo_periods[],
                                      int32_t dim, const int32_t padding[], cudaStream_t stream);

#ifdef __cplusplus
}
#endif

#endif /* CUDECOMP_H */
