// The C ABI (include/cudecomp.h): argument checking, struct versioning, lifecycle. Each entry point keeps
// the contract of the reference function it replaces (reference src/cudecomp.cc:903-2045): same checks in
// the same order, same result codes, no exception escapes.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <set>

#include <unistd.h>

#include "cudecomp.h"
#include "cudecomp_b200_ext.h"
#include "engine.h"
#include "launch_params.h"
#include "nvtx_ranges.h"
#include "errors.h"
#include "mpi_shim.h"
#include "vmm.h"

using namespace cdb;

namespace {

// ------------------------------------------------------------------------------------------------
// struct versioning (reference src/cudecomp.cc:209-414). Only layout version 1 exists.

constexpr int64_t kConfigSizeV1 = 104;
constexpr int64_t kOptionsSizeV1 = 320;
constexpr int64_t kPencilInfoSizeV1 = 96;
static_assert(sizeof(cudecompGridDescConfig_t) == kConfigSizeV1, "config ABI size changed");
static_assert(sizeof(cudecompGridDescAutotuneOptions_t) == kOptionsSizeV1, "options ABI size changed");
static_assert(sizeof(cudecompPencilInfo_t) == kPencilInfoSizeV1, "pencil info ABI size changed");

int64_t configSizeForVersion(int32_t version) {
  if (version > CUDECOMP_GRID_DESC_CONFIG_VERSION)
    THROW_INVALID_USAGE("config was initialized with a newer cuDecomp header than this runtime library supports");
  if (version != 1) THROW_INVALID_USAGE("config layout version is unsupported");
  return kConfigSizeV1;
}
int64_t optionsSizeForVersion(int32_t version) {
  if (version > CUDECOMP_GRID_DESC_AUTOTUNE_OPTIONS_VERSION)
    THROW_INVALID_USAGE("options were initialized with a newer cuDecomp header than this runtime library supports");
  if (version != 1) THROW_INVALID_USAGE("options layout version is unsupported");
  return kOptionsSizeV1;
}
int64_t pencilInfoSizeForVersion(int32_t version) {
  if (version > CUDECOMP_PENCIL_INFO_VERSION)
    THROW_INVALID_USAGE("pencil_info was created with a newer cuDecomp header than this runtime library supports");
  if (version != 1) THROW_INVALID_USAGE("pencil_info layout version is unsupported");
  return kPencilInfoSizeV1;
}

void setConfigDefaults(cudecompGridDescConfig_t* c, int64_t struct_size, int32_t version) {
  std::memset(c, 0, sizeof(*c));
  c->transpose_comm_backend = CUDECOMP_TRANSPOSE_COMM_MPI_P2P;
  c->halo_comm_backend = CUDECOMP_HALO_COMM_MPI;
  c->rank_order = CUDECOMP_RANK_ORDER_DEFAULT;
  for (auto& row : c->transpose_mem_order)
    for (auto& v : row) v = -1;
  c->struct_size = struct_size;
  c->magic = CUDECOMP_GRID_DESC_CONFIG_MAGIC;
  c->version = version;
}

void setOptionsDefaults(cudecompGridDescAutotuneOptions_t* o, int64_t struct_size, int32_t version) {
  std::memset(o, 0, sizeof(*o));
  o->n_warmup_trials = 3;
  o->n_trials = 5;
  o->grid_mode = CUDECOMP_AUTOTUNE_GRID_TRANSPOSE;
  o->dtype = CUDECOMP_DOUBLE;
  o->allow_uneven_decompositions = true;
  o->skip_threshold = 0.0;
  for (double& w : o->transpose_op_weights) w = 1.0;
  o->struct_size = struct_size;
  o->magic = CUDECOMP_GRID_DESC_AUTOTUNE_OPTIONS_MAGIC;
  o->version = version;
}

void checkConfigStruct(const cudecompGridDescConfig_t* c) {
  if (c->magic != CUDECOMP_GRID_DESC_CONFIG_MAGIC)
    THROW_INVALID_USAGE(
        "config is not initialized; call cudecompGridDescConfigSetDefaults() before cudecompGridDescCreate()");
  if (c->struct_size != configSizeForVersion(c->version))
    THROW_INVALID_USAGE("config struct_size does not match its cuDecomp layout version");
}

void checkOptionsStruct(const cudecompGridDescAutotuneOptions_t* o) {
  if (o->magic != CUDECOMP_GRID_DESC_AUTOTUNE_OPTIONS_MAGIC)
    THROW_INVALID_USAGE("options are not initialized; call cudecompGridDescAutotuneOptionsSetDefaults() before "
                        "cudecompGridDescCreate()");
  if (o->struct_size != optionsSizeForVersion(o->version))
    THROW_INVALID_USAGE("options struct_size does not match its cuDecomp layout version");
}

void copyConfigToCaller(cudecompGridDescConfig_t* dst, int64_t struct_size, int32_t version,
                        const cudecompGridDesc_t gd) {
  if (struct_size != configSizeForVersion(version))
    THROW_INVALID_USAGE("config struct_size does not match its cuDecomp layout version");
  *dst = gd->config;
  dst->struct_size = struct_size;
  dst->magic = CUDECOMP_GRID_DESC_CONFIG_MAGIC;
  dst->version = version;
  // settings the caller left unset are reported as unset (reference src/cudecomp.cc:1250-1265)
  if (!gd->gdims_dist_set)
    for (auto& v : dst->gdims_dist) v = 0;
  if (!gd->transpose_mem_order_set)
    for (auto& row : dst->transpose_mem_order)
      for (auto& v : row) v = -1;
}

// ------------------------------------------------------------------------------------------------
// argument checks (reference src/cudecomp.cc:136-207,445-481)

// Live objects of this process. Callers hand back raw pointers; a pointer is only dereferenced after it has been found
// here, so a stale handle or a grid descriptor that was already destroyed is an INVALID_USAGE error, not a read of
// freed memory (the reference dereferences them, src/cudecomp.cc:136-207).
std::set<cudecompHandle_t>& liveHandles() {
  static std::set<cudecompHandle_t> s;
  return s;
}
std::set<cudecompGridDesc_t>& liveGridDescs() {
  static std::set<cudecompGridDesc_t> s;
  return s;
}

void checkHandle(cudecompHandle_t h) {
  if (!h || !liveHandles().count(h) || !h->initialized) THROW_INVALID_USAGE("invalid handle");
}
void checkGridDesc(cudecompHandle_t h, cudecompGridDesc_t gd) {
  if (!gd || !liveGridDescs().count(gd) || !gd->initialized) THROW_INVALID_USAGE("invalid grid descriptor");
  if (gd->handle != h) THROW_INVALID_USAGE("grid descriptor belongs to a different handle");
}
// Enumerators arrive from C callers inside structs or as arguments and may hold any bit pattern; they are read through
// their integer representation so that validating them is well defined (an out-of-range value must yield
// INVALID_USAGE, not undefined behaviour on the load).
template <typename E> int enumBits(const E& e) {
  static_assert(sizeof(E) == sizeof(int), "cuDecomp enums are int-sized");
  int v;
  std::memcpy(&v, &e, sizeof(v));
  return v;
}

void checkTransposeBackend(int b) {
  if (b < CUDECOMP_TRANSPOSE_COMM_MPI_P2P || b > CUDECOMP_TRANSPOSE_COMM_NVSHMEM_SM)
    THROW_INVALID_USAGE("unknown transpose communication type");
}
void checkHaloBackend(int b) {
  if (b < CUDECOMP_HALO_COMM_MPI || b > CUDECOMP_HALO_COMM_NVSHMEM_BLOCKING)
    THROW_INVALID_USAGE("unknown halo communication type");
}
void checkDataType(const cudecompDataType_t& dtype) {
  const int d = enumBits(dtype);
  if (d != CUDECOMP_FLOAT && d != CUDECOMP_DOUBLE && d != CUDECOMP_FLOAT_COMPLEX && d != CUDECOMP_DOUBLE_COMPLEX)
    THROW_INVALID_USAGE("unknown data type");
}
void checkRankOrder(int r) {
  if (r < CUDECOMP_RANK_ORDER_DEFAULT || r > CUDECOMP_RANK_ORDER_COL_MAJOR) THROW_INVALID_USAGE("unknown rank order");
}

// transpose_mem_order: all unset (negative) or three permutations of {0, 1, 2} (reference src/cudecomp.cc:460-480)
void checkMemOrder(const cudecompGridDescConfig_t* c) {
  const bool set = c->transpose_mem_order[0][0] >= 0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      if ((c->transpose_mem_order[i][j] >= 0) != set) THROW_INVALID_USAGE("transpose_mem_order only partially set");
  if (set) {
    for (int i = 0; i < 3; ++i) {
      std::set<int32_t> vals(c->transpose_mem_order[i], c->transpose_mem_order[i] + 3);
      if (vals.size() != 3 || *vals.begin() != 0 || *vals.rbegin() != 2)
        THROW_INVALID_USAGE("transpose_mem_order setting is invalid");
    }
  }
}

void checkConfig(cudecompHandle_t h, const cudecompGridDescConfig_t* c, bool autotune_transpose, bool autotune_halo) {
  if (!autotune_transpose) checkTransposeBackend(enumBits(c->transpose_comm_backend));
  if (!autotune_halo) checkHaloBackend(enumBits(c->halo_comm_backend));
  checkRankOrder(enumBits(c->rank_order));
  if (c->pdims[0] < 0 || c->pdims[1] < 0) THROW_INVALID_USAGE("pdims values are invalid");
  const int64_t prod = static_cast<int64_t>(c->pdims[0]) * c->pdims[1];
  if (prod == 0) {
    if (c->pdims[0] != 0 || c->pdims[1] != 0) THROW_INVALID_USAGE("pdims values are invalid");
  } else if (prod != h->nranks) {
    THROW_INVALID_USAGE("product of pdims values must equal number of ranks");
  } else if (h->have_device && (c->pdims[0] > kMaxPeers + 1 || c->pdims[1] > kMaxPeers + 1)) {
    // said here rather than at the first transpose: a launch handshakes with at most kMaxPeers peers
    THROW_NOT_SUPPORTED("row / column communicators with more than 72 ranks are not supported");
  }
  checkMemOrder(c);
}

bool envFlag(const char* name) {
  const char* v = std::getenv(name);
  return v && *v && std::strcmp(v, "0") != 0;
}

cudecompResult_t report(const cdb::Error& e) {
  std::cerr << e.what();
  return e.code();
}
cudecompResult_t report(const std::exception& e, bool bootstrap) {
  if (bootstrap) {
    std::cerr << "CUDECOMP:ERROR: Bootstrap (MPI) error. (" << e.what() << ")\n";
    return CUDECOMP_RESULT_MPI_ERROR;
  }
  std::cerr << "CUDECOMP:ERROR: Internal error. (" << e.what() << ")\n";
  return CUDECOMP_RESULT_INTERNAL_ERROR;
}

void destroyGridDescResources(cudecompGridDesc_t gd, bool collective) {
  if (!gd) return;
  cudecompHandle_t h = gd->handle;
  if (gd->pad_slot >= 0) {
    // all ranks are past their last operation on this descriptor before anybody recycles the slot
    if (collective && h->nranks > 1) barrier(*h->comm);
    h->arena.zeroSlot(gd->pad_slot);
    h->free_slots.push_back(gd->pad_slot);
    gd->pad_slot = -1;
  }
  gd->mbox.destroy();
  releaseFusedCache(gd);
  if (collective) {
    // Drop every mapping of peer memory (imports are re-created on demand by the descriptors that stay alive): buffers
    // the CALLER owns may be freed once the descriptor they were used with is gone, and nothing may still map them then.
    if (h->have_device) h->peers.clear();
    drainReleases(h);
  }
  for (cudaEvent_t e : gd->side_events) cudaEventDestroy(e);
  gd->side_events.clear();
  if (gd->side_stream) cudaStreamDestroy(gd->side_stream);
  gd->side_stream = nullptr;
  (void)cudaGetLastError();
}

} // namespace

#define API_TRY try {
#define API_CATCH(...)                                                                                                 \
  }                                                                                                                    \
  catch (const cdb::Error& e) {                                                                                        \
    __VA_ARGS__;                                                                                                       \
    return report(e);                                                                                                  \
  }                                                                                                                    \
  catch (const cdb::BootstrapError& e) {                                                                               \
    __VA_ARGS__;                                                                                                       \
    return report(e, true);                                                                                            \
  }                                                                                                                    \
  catch (const std::exception& e) {                                                                                    \
    __VA_ARGS__;                                                                                                       \
    return report(e, false);                                                                                           \
  }                                                                                                                    \
  catch (...) {                                                                                                        \
    __VA_ARGS__;                                                                                                       \
    std::cerr << "CUDECOMP:ERROR: Internal error. (unknown exception)\n";                                              \
    return CUDECOMP_RESULT_INTERNAL_ERROR;                                                                             \
  }                                                                                                                    \
  return CUDECOMP_RESULT_SUCCESS;

extern "C" {

// ------------------------------------------------------------------------------------------------ lifecycle

static void initHandle(cudecompHandle_t h, const CommPtr& parent) {
  h->comm = dup(*parent); // private copy: library traffic never interleaves with the caller's collectives
  h->rank = h->comm->rank();
  h->nranks = h->comm->size();

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 && cudaGetDevice(&h->device) == cudaSuccess) {
    h->have_device = true;
    CHECK_CUDA(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device));
  } else {
    (void)cudaGetLastError();
  }
  h->env_col_major = envFlag("CUDECOMP_USE_COL_MAJOR_RANK_ORDER");
  h->perf.readEnvironment();
  if (const char* v = std::getenv("CUDECOMP_B200_PIPELINE_CHUNKS")) h->pipeline_chunks = std::max(0, std::atoi(v));
  if (const char* v = std::getenv("CUDECOMP_B200_KERNEL"))
    h->kernel_variant = (std::strcmp(v, "bulk") == 0) ? 1 : (std::strcmp(v, "wide") == 0 ? 2 : 0);
  if (const char* v = std::getenv("CUDECOMP_B200_TILE_BYTES")) h->tile_bytes = std::max(0, std::atoi(v));
  if (const char* v = std::getenv("CUDECOMP_B200_PEER_ORDER"))
    h->peer_order = (std::strcmp(v, "pairwise") == 0 || std::strcmp(v, "1") == 0) ? 1 : 0;
  if (const char* v = std::getenv("CUDECOMP_B200_BALANCE_GRID")) h->balance_grid = std::atoi(v) != 0 ? 1 : 0;
  if (const char* v = std::getenv("CUDECOMP_B200_TRANSFER")) h->pull_mode = (std::strcmp(v, "pull") == 0) ? 1 : 0;
  if (const char* v = std::getenv("CUDECOMP_B200_STAGED")) h->staged_mode = (std::strcmp(v, "launches") == 0) ? 1 : 0;
  if (const char* v = std::getenv("CUDECOMP_B200_FUSED_LAG")) h->fused_lag = std::min(8, std::max(1, std::atoi(v)));
  if (const char* v = std::getenv("CUDECOMP_B200_PHASE_HEAD")) h->phase_head_percent = std::min(90, std::max(0, std::atoi(v)));
  if (const char* v = std::getenv("CUDECOMP_B200_TRANSPOSE_GEOM")) h->transpose_geometry = std::atoi(v) != 0 ? 1 : 0;
  if (const char* v = std::getenv("CUDECOMP_B200_COLUMN_CHUNKS")) h->column_chunks = std::atoi(v) != 0 ? 1 : 0;
  if (const char* v = std::getenv("CUDECOMP_B200_WIRE_WIDE")) h->wire_wide = std::atoi(v) != 0 ? 1 : 0;
  if (const char* v = std::getenv("CUDECOMP_B200_DIRECT")) h->allow_direct = std::strcmp(v, "0") != 0;
  double spin_s = 60.0;
  if (const char* v = std::getenv("CUDECOMP_B200_DEVICE_TIMEOUT")) spin_s = std::atof(v);
  h->spin_timeout_ns = static_cast<uint64_t>(spin_s * 1e9);
  h->token = sharedToken(*h->comm);

  // CUDECOMP_ENABLE_CUMEM (reference src/cudecomp.cc:596-660, docs/env_vars.rst): cudecompMalloc hands out cuMem (VMM)
  // allocations with shareable handles. The reference leaves the sharing to NCCL / MPI; here the peers map the buffers
  // themselves (vmm.h), so besides the device's VMM support the ranks must be able to pass file descriptors to each
  // other. Both are checked once, collectively; a request that cannot be met is dropped with the reference's warning.
  // (CUDECOMP_ENABLE_NCCL_UBR requests the same allocation path in the reference, src/cudecomp.cc:603; there is no NCCL
  // communicator to register the buffers with here, so that is all it does.)
  int64_t want[2] = {(envFlag("CUDECOMP_ENABLE_CUMEM") || envFlag("CUDECOMP_ENABLE_NCCL_UBR")) ? 1 : 0, 0};
  want[1] = -want[0];
  allreduceI64(*h->comm, want, 2, ReduceOp::MIN); // {min, -max}: every rank must ask for it
  if (want[0] != -want[1] && h->rank == 0)
    std::printf("CUDECOMP:WARN: CUDECOMP_ENABLE_CUMEM is not set on every rank. Disabling this feature.\n");
  if (want[0] == 1) {
    bool fabric = false;
    int64_t ok[2] = {(h->have_device && vmmDeviceSupported(&fabric)) ? 1 : 0, probeFdPassing(*h->comm, h->token) ? 1 : 0};
    h->cumem_fabric = fabric && envFlag("CUDECOMP_B200_CUMEM_FABRIC");
    if (h->cumem_fabric) ok[1] = 1; // fabric handles travel in the descriptor itself: no file descriptor needed
    allreduceI64(*h->comm, ok, 2, ReduceOp::MIN);
    h->cumem_state = !ok[0] ? kCumemNoDevice : (!ok[1] ? kCumemNoFdPassing : kCumemOn);
    if (h->rank == 0 && h->cumem_state == kCumemNoDevice)
      std::printf("CUDECOMP:WARN: CUDECOMP_ENABLE_CUMEM is set but the current device does not support CUDA VMM "
                  "allocations with POSIX file-descriptor handles. Disabling this feature.\n");
    if (h->rank == 0 && h->cumem_state == kCumemNoFdPassing)
      std::printf("CUDECOMP:WARN: CUDECOMP_ENABLE_CUMEM is set but the ranks cannot pass file descriptors to each other "
                  "(pidfd_getfd). Disabling this feature.\n");
  }
  h->initialized = true;
}

cudecompResult_t cudecompInit(cudecompHandle_t* handle_in, MPI_Comm mpi_comm) {
  cudecompHandle_t h = nullptr;
  API_TRY
  if (!handle_in) THROW_INVALID_USAGE("handle argument cannot be null");
  CommPtr parent = commFromHandle(static_cast<int>(mpi_comm));
  if (!parent) THROW_INVALID_USAGE("invalid communicator");
  h = new cudecompHandle;
  initHandle(h, parent);
  liveHandles().insert(h);
  *handle_in = h;
  API_CATCH(delete h)
}

// For applications that run on a real MPI (include/cudecomp_b200_mpi.h): the caller says who it is and where rank 0
// listens; the library builds its own control-plane mesh among exactly these processes. No MPI type crosses the ABI.
cudecompResult_t cudecompB200InitBootstrap(cudecompHandle_t* handle_in, int32_t rank, int32_t nranks, const char* root_addr,
                                           int32_t root_port) {
  cudecompHandle_t h = nullptr;
  API_TRY
  if (!handle_in) THROW_INVALID_USAGE("handle argument cannot be null");
  if (nranks < 1 || rank < 0 || rank >= nranks) THROW_INVALID_USAGE("rank / nranks are invalid");
  if (nranks > 1 && (!root_addr || !*root_addr || root_port <= 0 || root_port > 65535))
    THROW_INVALID_USAGE("root_addr / root_port are invalid");
  worldInitExplicit(rank, nranks, root_addr ? root_addr : "127.0.0.1", root_port);
  h = new cudecompHandle;
  initHandle(h, worldComm());
  liveHandles().insert(h);
  *handle_in = h;
  API_CATCH(delete h)
}

cudecompResult_t cudecompB200PickBootstrapPort(int32_t* port) {
  API_TRY
  if (!port) THROW_INVALID_USAGE("port argument cannot be null");
  *port = pickFreePort();
  API_CATCH()
}

cudecompResult_t cudecompInit_F(cudecompHandle_t* handle_in, MPI_Fint mpi_comm_f) {
  return cudecompInit(handle_in, MPI_Comm_f2c(mpi_comm_f));
}

cudecompResult_t cudecompFinalize(cudecompHandle_t handle) {
  API_TRY
  checkHandle(handle);
  handle->peers.clear();
  if (!handle->released_pending.empty() || handle->acks.valid()) {
    // every rank has closed all its imports above; after the barrier nothing of mine is mapped anywhere
    if (handle->nranks > 1 && handle->acks.valid()) barrier(*handle->comm);
    reapReleased(handle, true);
    handle->acks.destroy();
  }
  if (handle->arena.valid()) handle->arena.destroy(handle->comm.get());
  handle->initialized = false;
  liveHandles().erase(handle);
  delete handle;
  API_CATCH()
}

// --------------------------------------------------------------------------------------- grid descriptor

cudecompResult_t cudecompGridDescCreateVersioned(cudecompHandle_t handle, cudecompGridDesc_t* grid_desc_in,
                                                 cudecompGridDescConfig_t* config, int64_t config_struct_size,
                                                 int32_t config_version,
                                                 const cudecompGridDescAutotuneOptions_t* options,
                                                 int64_t options_struct_size, int32_t options_version) {
  cudecompGridDesc_t gd = nullptr;
  API_TRY
  checkHandle(handle);
  if (!grid_desc_in) THROW_INVALID_USAGE("grid_desc argument cannot be null");
  if (!config) THROW_INVALID_USAGE("config argument cannot be null");
  if (config_struct_size != configSizeForVersion(config_version))
    THROW_INVALID_USAGE("config struct_size does not match its cuDecomp layout version");
  checkConfigStruct(config);
  if (config->struct_size != config_struct_size || config->version != config_version)
    THROW_INVALID_USAGE("config metadata does not match the requested cuDecomp layout version");
  if (options) {
    if (options_struct_size != optionsSizeForVersion(options_version))
      THROW_INVALID_USAGE("options struct_size does not match its cuDecomp layout version");
    checkOptionsStruct(options);
    if (options->struct_size != options_struct_size || options->version != options_version)
      THROW_INVALID_USAGE("options metadata does not match the requested cuDecomp layout version");
    // rejected even when nothing is left to tune (reference src/cudecomp.cc:1201-1210)
    if (enumBits(options->grid_mode) != CUDECOMP_AUTOTUNE_GRID_TRANSPOSE &&
        enumBits(options->grid_mode) != CUDECOMP_AUTOTUNE_GRID_HALO)
      THROW_INVALID_USAGE("unknown value of autotune_grid_mode encountered.");
  }
  const bool autotune_transpose = options && options->autotune_transpose_backend;
  const bool autotune_halo = options && options->autotune_halo_backend;
  checkConfig(handle, config, autotune_transpose, autotune_halo);
  const bool autotune_pdims = (config->pdims[0] == 0 && config->pdims[1] == 0);
  if (autotune_pdims && !options) THROW_INVALID_USAGE("options argument cannot be null if autotuning pdims");

  gd = new cudecompGridDesc;
  gd->initialized = true;
  gd->handle = handle;
  gd->config = *config;
  gd->pipeline_chunks = handle->pipeline_chunks;
  gd->kernel_variant = handle->kernel_variant;
  gd->tile_bytes = handle->tile_bytes;
  gd->peer_order = handle->peer_order;
  gd->balance_grid = handle->balance_grid;
  gd->pull_mode = handle->pull_mode;
  gd->staged_mode = handle->staged_mode;
  gd->fused_lag = handle->fused_lag;
  gd->phase_head_percent = handle->phase_head_percent;
  gd->wire_wide = handle->wire_wide;
  gd->column_chunks = handle->column_chunks;
  if (gd->config.rank_order == CUDECOMP_RANK_ORDER_DEFAULT)
    gd->config.rank_order = handle->env_col_major ? CUDECOMP_RANK_ORDER_COL_MAJOR : CUDECOMP_RANK_ORDER_ROW_MAJOR;

  // memory order per pencil axis (reference src/cudecomp.cc:1120-1133)
  gd->transpose_mem_order_set = (config->transpose_mem_order[0][0] >= 0);
  if (!gd->transpose_mem_order_set)
    for (int axis = 0; axis < 3; ++axis)
      for (int i = 0; i < 3; ++i)
        gd->config.transpose_mem_order[axis][i] = gd->config.transpose_axis_contiguous[axis] ? (axis + i) % 3 : i;

  // distribution grid (reference src/cudecomp.cc:1135-1150)
  for (int i = 0; i < 3; ++i)
    if (gd->config.gdims_dist[i] > gd->config.gdims[i])
      THROW_INVALID_USAGE("gdims_dist entries must be less than or equal to gdims entries");
  gd->gdims_dist_set = gd->config.gdims_dist[0] != 0 && gd->config.gdims_dist[1] != 0 && gd->config.gdims_dist[2] != 0;
  if (!gd->gdims_dist_set)
    for (int i = 0; i < 3; ++i) gd->config.gdims_dist[i] = gd->config.gdims[i];

  // Device-side plumbing shared by every operation on this descriptor (collective).
  if (handle->have_device) {
    if (!handle->arena.valid()) handle->arena.create(*handle->comm);
    if (!handle->acks.valid() && handle->nranks > 1) handle->acks.create(*handle->comm, handle->token);
    if (!handle->free_slots.empty()) {
      gd->pad_slot = handle->free_slots.back();
      handle->free_slots.pop_back();
    } else {
      if (handle->next_slot >= SignalArena::kSlots) THROW_NOT_SUPPORTED("too many live grid descriptors");
      gd->pad_slot = handle->next_slot++;
    }
    if (handle->nranks > 1) gd->mbox.create(*handle->comm, handle->token, handle->next_instance);
  }
  handle->next_instance++;

  if (autotune_pdims || autotune_transpose || autotune_halo) {
    autotune(handle, gd, options);
  }
  setGeometry(gd, {gd->config.pdims[0], gd->config.pdims[1]});
  if (gd->mbox.valid()) gd->mbox.reset(*handle->comm);
  // samples taken from here on (autotuning trials are not part of the report, reference src/autotune.cc "reset")
  if (handle->perf.enabled && handle->have_device) gd->perf.reset(new PerfReport(handle->perf));

  copyConfigToCaller(config, config_struct_size, config_version, gd);
  handle->live_grid_descs++;
  liveGridDescs().insert(gd);
  *grid_desc_in = gd;
  API_CATCH(if (gd) {
    destroyGridDescResources(gd, false);
    delete gd;
  })
}

cudecompResult_t cudecompGridDescDestroy(cudecompHandle_t handle, cudecompGridDesc_t grid_desc) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  if (handle->have_device) cudaDeviceSynchronize();
  if (grid_desc->perf) {
    grid_desc->perf->print(handle, grid_desc); // collective, like the reference (src/cudecomp.cc:1278)
    grid_desc->perf.reset();
  }
  destroyGridDescResources(grid_desc, true);
  grid_desc->initialized = false;
  handle->live_grid_descs--;
  liveGridDescs().erase(grid_desc);
  delete grid_desc;
  API_CATCH()
}

cudecompResult_t cudecompGridDescConfigSetDefaultsVersioned(cudecompGridDescConfig_t* config, int64_t struct_size,
                                                            int32_t version) {
  API_TRY
  if (!config) THROW_INVALID_USAGE("config argument cannot be null");
  if (struct_size != configSizeForVersion(version))
    THROW_INVALID_USAGE("config struct_size does not match its cuDecomp layout version");
  setConfigDefaults(config, struct_size, version);
  API_CATCH()
}

cudecompResult_t cudecompGridDescAutotuneOptionsSetDefaultsVersioned(cudecompGridDescAutotuneOptions_t* options,
                                                                     int64_t struct_size, int32_t version) {
  API_TRY
  if (!options) THROW_INVALID_USAGE("options argument cannot be null");
  if (struct_size != optionsSizeForVersion(version))
    THROW_INVALID_USAGE("options struct_size does not match its cuDecomp layout version");
  setOptionsDefaults(options, struct_size, version);
  API_CATCH()
}

// ------------------------------------------------------------------------------------------------ queries

cudecompResult_t cudecompGetPencilInfoVersioned(cudecompHandle_t handle, cudecompGridDesc_t grid_desc,
                                                cudecompPencilInfo_t* pencil_info, int64_t pencil_info_struct_size,
                                                int32_t pencil_info_version, int32_t axis, const int32_t halo_extents[],
                                                const int32_t padding[]) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  if (!pencil_info) THROW_INVALID_USAGE("pencil_info argument cannot be null.");
  if (pencil_info_struct_size != pencilInfoSizeForVersion(pencil_info_version))
    THROW_INVALID_USAGE("pencil_info struct_size does not match its cuDecomp layout version");
  if (axis < 0 || axis > 2) THROW_INVALID_USAGE("axis argument out of range");
  const Pencil p = pencilInfo(grid_desc->geom, grid_desc->pidx, axis, halo_extents, padding);
  cudecompPencilInfo_t out;
  std::memset(&out, 0, sizeof(out));
  for (int i = 0; i < 3; ++i) {
    out.shape[i] = p.shape[i];
    out.lo[i] = p.lo[i];
    out.hi[i] = p.hi[i];
    out.order[i] = p.order[i];
    out.halo_extents[i] = p.halo[i];
    out.padding[i] = p.pad[i];
  }
  out.size = p.size;
  out.struct_size = pencil_info_struct_size;
  out.magic = CUDECOMP_PENCIL_INFO_MAGIC;
  out.version = pencil_info_version;
  *pencil_info = out;
  API_CATCH()
}

cudecompResult_t cudecompGetGridDescConfigVersioned(cudecompHandle_t handle, cudecompGridDesc_t grid_desc,
                                                    cudecompGridDescConfig_t* config, int64_t struct_size,
                                                    int32_t version) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  if (!config) THROW_INVALID_USAGE("config argument cannot be null.");
  copyConfigToCaller(config, struct_size, version, grid_desc);
  API_CATCH()
}

cudecompResult_t cudecompGetTransposeWorkspaceSize(cudecompHandle_t handle, cudecompGridDesc_t grid_desc,
                                                   int64_t* workspace_size) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  if (!workspace_size) THROW_INVALID_USAGE("workspace_size argument cannot be null.");
  *workspace_size = transposeWorkspaceSize(grid_desc->geom);
  API_CATCH()
}

cudecompResult_t cudecompGetHaloWorkspaceSize(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t axis,
                                              const int32_t halo_extents[], int64_t* workspace_size) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  if (axis < 0 || axis > 2) THROW_INVALID_USAGE("axis argument out of range");
  if (!halo_extents) THROW_INVALID_USAGE("halo_extents argument cannot be null.");
  if (!workspace_size) THROW_INVALID_USAGE("workspace_size argument cannot be null.");
  *workspace_size = haloWorkspaceSize(grid_desc->geom, grid_desc->pidx, axis, halo_extents);
  API_CATCH()
}

cudecompResult_t cudecompGetDataTypeSize(cudecompDataType_t dtype, int64_t* dtype_size) {
  API_TRY
  checkDataType(dtype);
  if (!dtype_size) THROW_INVALID_USAGE("dtype_size cannot be null.");
  *dtype_size = dtypeSize(dtype);
  API_CATCH()
}

cudecompResult_t cudecompGetShiftedRank(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t axis,
                                        int32_t dim, int32_t displacement, bool periodic, int32_t* shifted_rank) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  if (axis < 0 || axis > 2) THROW_INVALID_USAGE("axis argument out of range");
  if (dim < 0 || dim > 2) THROW_INVALID_USAGE("dim argument out of range");
  if (!shifted_rank) THROW_INVALID_USAGE("shifted_rank argument cannot be null.");
  *shifted_rank = shiftedRank(grid_desc->geom, handle->rank, axis, dim, displacement, periodic);
  API_CATCH()
}

const char* cudecompTransposeCommBackendToString(cudecompTransposeCommBackend_t comm_backend) {
  switch (enumBits(comm_backend)) {
  case CUDECOMP_TRANSPOSE_COMM_NCCL: return "NCCL";
  case CUDECOMP_TRANSPOSE_COMM_NCCL_PL: return "NCCL (pipelined)";
  case CUDECOMP_TRANSPOSE_COMM_MPI_P2P: return "MPI_P2P";
  case CUDECOMP_TRANSPOSE_COMM_MPI_P2P_PL: return "MPI_P2P (pipelined)";
  case CUDECOMP_TRANSPOSE_COMM_MPI_A2A: return "MPI_A2A";
  case CUDECOMP_TRANSPOSE_COMM_NVSHMEM: return "NVSHMEM";
  case CUDECOMP_TRANSPOSE_COMM_NVSHMEM_PL: return "NVSHMEM (pipelined)";
  case CUDECOMP_TRANSPOSE_COMM_NVSHMEM_SM: return "NVSHMEM_SM";
  default: return "ERROR";
  }
}

const char* cudecompHaloCommBackendToString(cudecompHaloCommBackend_t comm_backend) {
  switch (enumBits(comm_backend)) {
  case CUDECOMP_HALO_COMM_NCCL: return "NCCL";
  case CUDECOMP_HALO_COMM_MPI: return "MPI";
  case CUDECOMP_HALO_COMM_MPI_BLOCKING: return "MPI (blocking)";
  case CUDECOMP_HALO_COMM_NVSHMEM: return "NVSHMEM";
  case CUDECOMP_HALO_COMM_NVSHMEM_BLOCKING: return "NVSHMEM (blocking)";
  default: return "ERROR";
  }
}

// ---------------------------------------------------------------------------------------------- workspaces

cudecompResult_t cudecompMalloc(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, void** buffer,
                                size_t buffer_size_bytes) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  if (!buffer) THROW_INVALID_USAGE("buffer argument cannot be null");
  if (buffer_size_bytes == 0) THROW_INVALID_USAGE("buffer size cannot be zero");
  if (!handle->have_device)
    CDB_THROW(CUDECOMP_RESULT_CUDA_ERROR, "CUDA error.", "no CUDA device is available to this process");
  // Plain device memory: peers import it with CUDA IPC the first time an operation names it, so no
  // symmetric-size agreement is needed (the reference's NVSHMEM arm MAX-reduces the size, src/cudecomp.cc:1474).
  // Rounded up to whole 2 MiB pages: small cudaMalloc requests share a driver block (and with it one IPC handle),
  // which peers cannot import twice.
  const size_t kPage = size_t(2) << 20;
  const size_t rounded = (buffer_size_bytes + kPage - 1) / kPage * kPage;
  void* p = nullptr;
  if (handle->cumem_state == kCumemOn) {
    // CUDECOMP_ENABLE_CUMEM: cuMemCreate (POSIX fd [+ fabric] handles, RDMA flag) + map, reference src/cudecomp.cc:1500-1570
    p = vmmAlloc(rounded, handle->cumem_fabric);
  } else {
    CHECK_CUDA(cudaMalloc(&p, rounded));
  }
  grid_desc->allocations.insert(p);
  *buffer = p;
  API_CATCH()
}

cudecompResult_t cudecompFree(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, void* buffer) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  if (handle->have_device && buffer) {
    // Not collective: the reference's own test drivers free and re-allocate workspaces on the ranks that need a
    // larger one only (tests/cc/halo_test.cc workspace reuse). The release is announced to the peers with the next
    // operation's descriptor exchange, where they drop their imports of this allocation (peer.h, CallMsg).
    // The memory itself is only returned to the driver once every rank that may have imported it has closed its
    // mapping (freeing an allocation that another process still has open through CUDA IPC is undefined): those ranks
    // acknowledge on the handle's AckBoard, and the pending frees are retried at every later exchange, free and
    // descriptor destruction.
    CHECK_CUDA(cudaDeviceSynchronize());
    BufDesc d;
    describeBuffer(buffer, &d);
    std::vector<int> shown;
    if (d.exportable && d.offset == 0) shown = takeDescribedTo(buffer);
    if (shown.empty()) {
      if (!vmmFree(buffer)) CHECK_CUDA(cudaFree(buffer));
    } else {
      for (int k = kReleaseSlots - 1; k > 0; --k) handle->released[k] = handle->released[k - 1];
      handle->released[0] = d.buffer_id;
      handle->release_count++;
      handle->released_pending.push_back({buffer, handle->release_count, std::move(shown)});
    }
    reapReleased(handle);
  }
  grid_desc->allocations.erase(buffer);
  API_CATCH()
}

// ---------------------------------------------------------------------------------------------- hot path

#define TRANSPOSE_ENTRY(NAME, AX, DIR)                                                                                 \
  cudecompResult_t NAME(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, void* input, void* output, void* work,  \
                        cudecompDataType_t dtype, const int32_t input_halo_extents[],                                  \
                        const int32_t output_halo_extents[], const int32_t input_padding[],                            \
                        const int32_t output_padding[], cudaStream_t stream) {                                         \
    API_TRY                                                                                                            \
    checkHandle(handle);                                                                                               \
    checkGridDesc(handle, grid_desc);                                                                                  \
    checkDataType(dtype);                                                                                              \
    if (!input) THROW_INVALID_USAGE("input argument cannot be null");                                                  \
    if (!output) THROW_INVALID_USAGE("output argument cannot be null");                                                \
    if (!work) THROW_INVALID_USAGE("work argument cannot be null");                                                    \
    NvtxRange nvtx_range(#NAME);                                                                                       \
    runTranspose(handle, grid_desc, AX, DIR, input, output, work, dtype, input_halo_extents, output_halo_extents,      \
                 input_padding, output_padding, stream);                                                               \
    API_CATCH()                                                                                                        \
  }

// (ax, dir): XY = (0,+1), YZ = (1,+1), ZY = (2,-1), YX = (1,-1)  (reference include/internal/transpose.h:907-953)
TRANSPOSE_ENTRY(cudecompTransposeXToY, 0, 1)
TRANSPOSE_ENTRY(cudecompTransposeYToZ, 1, 1)
TRANSPOSE_ENTRY(cudecompTransposeZToY, 2, -1)
TRANSPOSE_ENTRY(cudecompTransposeYToX, 1, -1)

#define HALO_ENTRY(NAME, AX)                                                                                           \
  cudecompResult_t NAME(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, void* input, void* work,                \
                        cudecompDataType_t dtype, const int32_t halo_extents[], const bool halo_periods[],             \
                        int32_t dim, const int32_t padding[], cudaStream_t stream) {                                   \
    API_TRY                                                                                                            \
    checkHandle(handle);                                                                                               \
    checkGridDesc(handle, grid_desc);                                                                                  \
    checkDataType(dtype);                                                                                              \
    if (!halo_extents) THROW_INVALID_USAGE("halo_extents argument cannot be null");                                    \
    if (halo_extents[0] == 0 && halo_extents[1] == 0 && halo_extents[2] == 0) return CUDECOMP_RESULT_SUCCESS;          \
    if (!input) THROW_INVALID_USAGE("input argument cannot be null");                                                  \
    if (!work) THROW_INVALID_USAGE("work argument cannot be null");                                                    \
    if (dim < 0 || dim > 2) THROW_INVALID_USAGE("dim argument out of range");                                          \
    NvtxRange nvtx_range(std::string(#NAME "_") + std::to_string(dim));                                                \
    runHalo(handle, grid_desc, AX, input, work, dtype, halo_extents, halo_periods, dim, padding, stream);              \
    API_CATCH()                                                                                                        \
  }

HALO_ENTRY(cudecompUpdateHalosX, 0)
HALO_ENTRY(cudecompUpdateHalosY, 1)
HALO_ENTRY(cudecompUpdateHalosZ, 2)

// ---------------------------------------------------------------------------------------------- extensions

cudecompResult_t cudecompB200GetLaunchCount(uint64_t* count) {
  if (!count) return CUDECOMP_RESULT_INVALID_USAGE;
  *count = launchCount();
  return CUDECOMP_RESULT_SUCCESS;
}

cudecompResult_t cudecompB200GetCumemState(cudecompHandle_t handle, int32_t* state) {
  API_TRY
  checkHandle(handle);
  if (!state) THROW_INVALID_USAGE("state argument cannot be null");
  *state = handle->cumem_state;
  API_CATCH()
}

cudecompResult_t cudecompB200ProbeFdPassing(cudecompHandle_t handle, int32_t* ok) {
  API_TRY
  checkHandle(handle);
  if (!ok) THROW_INVALID_USAGE("ok argument cannot be null");
  *ok = probeFdPassing(*handle->comm, handle->token ^ 0xfd9a55ull) ? 1 : 0;
  API_CATCH()
}

cudecompResult_t cudecompB200GetLastPath(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t* path) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  if (!path) THROW_INVALID_USAGE("path argument cannot be null");
  *path = grid_desc->last_path;
  API_CATCH()
}

cudecompResult_t cudecompB200SetTuning(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t grid_ctas,
                                       int32_t force_staged) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  if (grid_ctas < 0) THROW_INVALID_USAGE("grid_ctas must be >= 0");
  grid_desc->grid_ctas = grid_ctas;
  grid_desc->force_staged = force_staged != 0;
  API_CATCH()
}

cudecompResult_t cudecompB200SetKernelVariant(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t variant) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  if (variant < 0 || variant > 3)
    THROW_INVALID_USAGE("variant must be 0 (default), 1 (TMA bulk), 2 (LDG/STG.256) or 3 (element-wise transpose)");
  grid_desc->kernel_variant = variant;
  API_CATCH()
}

cudecompResult_t cudecompB200SetSchedule(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t tile_bytes,
                                         int32_t peer_order, int32_t balance_grid) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  if (tile_bytes != 0 && (tile_bytes < kMinTileBytes || tile_bytes > kMaxTileBytes || (tile_bytes & (tile_bytes - 1)) != 0))
    THROW_INVALID_USAGE("tile_bytes must be 0 or a power of two in [4096, 262144]");
  if (peer_order < 0 || peer_order > 1) THROW_INVALID_USAGE("peer_order must be 0 (interleaved) or 1 (pairwise rounds)");
  grid_desc->tile_bytes = tile_bytes;
  grid_desc->peer_order = peer_order;
  grid_desc->balance_grid = balance_grid != 0 ? 1 : 0;
  API_CATCH()
}

cudecompResult_t cudecompB200SetTransferMode(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t mode) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  if (mode < 0 || mode > 1) THROW_INVALID_USAGE("mode must be 0 (sender-driven, push) or 1 (receiver-driven, pull)");
  grid_desc->pull_mode = mode; // must be set to the same value on every rank
  API_CATCH()
}

cudecompResult_t cudecompB200SetPipelineChunks(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t nchunks) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  if (nchunks < 0 || nchunks > 64) THROW_INVALID_USAGE("nchunks must be in [0, 64]");
  grid_desc->pipeline_chunks = nchunks; // must be set to the same value on every rank
  API_CATCH()
}

cudecompResult_t cudecompB200SetStagedMode(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t mode, int32_t lag) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  if (mode < 0 || mode > 1) THROW_INVALID_USAGE("mode must be 0 (one phased launch) or 1 (separate launches)");
  if (lag < 0 || lag > 8) THROW_INVALID_USAGE("lag must be in [0, 8] (0 keeps the current value)");
  grid_desc->staged_mode = mode; // must be set to the same value on every rank
  if (lag > 0) grid_desc->fused_lag = lag;
  API_CATCH()
}

cudecompResult_t cudecompB200CheckErrors(cudecompHandle_t handle, cudecompGridDesc_t grid_desc) {
  API_TRY
  checkHandle(handle);
  checkGridDesc(handle, grid_desc);
  checkDeviceError(grid_desc);
  API_CATCH()
}

cudecompResult_t cudecompB200GetAutotuneCandidates(const cudecompGridDescAutotuneOptions_t* options, int32_t nranks,
                                                   cudecompRankOrder_t rank_order, int32_t transpose_backends[8],
                                                   int32_t* n_transpose, int32_t halo_backends[5], int32_t* n_halo,
                                                   int32_t pdims[][2], int32_t max_pdims, int32_t* n_pdims) {
  API_TRY
  if (!options) THROW_INVALID_USAGE("options argument cannot be null");
  checkOptionsStruct(options);
  if (transpose_backends && n_transpose) {
    auto c = autotuneTransposeBackendCandidates(options);
    *n_transpose = static_cast<int32_t>(c.size());
    for (size_t i = 0; i < c.size(); ++i) transpose_backends[i] = c[i];
  }
  if (halo_backends && n_halo) {
    auto c = autotuneHaloBackendCandidates(options);
    *n_halo = static_cast<int32_t>(c.size());
    for (size_t i = 0; i < c.size(); ++i) halo_backends[i] = c[i];
  }
  if (pdims && n_pdims) {
    if (nranks < 1) THROW_INVALID_USAGE("nranks must be positive");
    auto c = autotunePdimCandidates(nranks, rank_order == CUDECOMP_RANK_ORDER_COL_MAJOR);
    *n_pdims = static_cast<int32_t>(c.size());
    for (size_t i = 0; i < c.size() && static_cast<int32_t>(i) < max_pdims; ++i) {
      pdims[i][0] = c[i][0];
      pdims[i][1] = c[i][1];
    }
  }
  API_CATCH()
}

cudecompResult_t cudecompB200SelfTestMailbox(cudecompHandle_t handle, int32_t iterations, uint32_t seed) {
  API_TRY
  checkHandle(handle);
  const int n = handle->nranks, me = handle->rank;
  if (n == 1) return CUDECOMP_RESULT_SUCCESS;
  Mailbox box;
  box.create(*handle->comm, handle->token ^ 0x5e1f7e57ull, 1000000 + handle->next_instance++);
  // Group shapes as the engine uses them: channel 0 = "column" groups {r : r % cols == me % cols},
  // channel 1 = "row" groups {r : r / cols == me / cols}, for a process grid that changes every few iterations.
  uint32_t lcg = seed * 2654435761u + 12345u;
  auto rnd = [&]() { return lcg = lcg * 1664525u + 1013904223u; };
  uint32_t my_lcg = (seed + 17u * static_cast<uint32_t>(me)) * 2246822519u + 1u;
  std::vector<int> divisors;
  for (int d = 1; d <= n; ++d)
    if (n % d == 0) divisors.push_back(d);
  int cols = 1;
  for (int it = 0; it < iterations; ++it) {
    // The process grid (and with it the membership of the row/column groups) only changes together with a reset,
    // exactly as in autotuning; between resets every member of a group issues the same sequence of exchanges.
    if (it % 7 == 0) {
      box.reset(*handle->comm);
      cols = divisors[rnd() % divisors.size()]; // same on every rank
    }
    const int channel = static_cast<int>(rnd() % 2);
    std::vector<int> members;
    int my_index = -1;
    for (int r = 0; r < n; ++r) {
      const bool in = (channel == 0) ? (r % cols == me % cols) : (r / cols == me / cols);
      if (in) {
        if (r == me) my_index = static_cast<int>(members.size());
        members.push_back(r);
      }
    }
    if (members.size() < 2) continue;
    my_lcg = my_lcg * 1664525u + 1013904223u;
    if ((my_lcg >> 28) == 0) usleep((my_lcg >> 8) % 300); // desynchronise the ranks a little
    CallMsg mine;
    std::memset(&mine, 0, sizeof(mine));
    mine.opcode = 0x700u + static_cast<uint32_t>(channel);
    mine.flags = static_cast<uint32_t>(it);
    mine.data.offset = static_cast<uint64_t>(me) * 1000003ull + static_cast<uint64_t>(it);
    mine.release_count = static_cast<uint64_t>(it) ^ 0xabcdu;
    std::vector<CallMsg> got;
    box.exchange(channel, members, my_index, mine, got);
    for (size_t i = 0; i < members.size(); ++i) {
      const uint64_t want = static_cast<uint64_t>(members[i]) * 1000003ull + static_cast<uint64_t>(it);
      if (got[i].flags != static_cast<uint32_t>(it) || got[i].data.offset != want ||
          got[i].release_count != (static_cast<uint64_t>(it) ^ 0xabcdu))
        THROW_INTERNAL_ERROR("mailbox self test: wrong message from rank " + std::to_string(members[i]) + " at iteration " +
                             std::to_string(it));
    }
  }
  barrier(*handle->comm);
  box.destroy();

  // The acknowledgement board behind deferred frees (peer.h AckBoard), host-only as well: every rank acknowledges a
  // growing number of "releases" of every other rank; an owner must see the counts of all readers reach each stamp, and
  // never a count it did not expect (cells of other (reader, owner) pairs must not be touched).
  AckBoard board;
  board.create(*handle->comm, handle->token ^ (0xacc0ull + static_cast<uint64_t>(handle->next_instance++)));
  const int rounds = std::max(1, iterations / 100);
  for (int it = 1; it <= rounds; ++it) {
    for (int owner = 0; owner < n; ++owner)
      if (owner != me) board.publish(owner, static_cast<uint64_t>(it) * 1000ull + static_cast<uint64_t>(owner));
    barrier(*handle->comm);
    for (int reader = 0; reader < n; ++reader) {
      if (reader == me) {
        if (board.seen(me, me) != 0) THROW_INTERNAL_ERROR("ack board self test: a rank acknowledged itself");
        continue;
      }
      const uint64_t want = static_cast<uint64_t>(it) * 1000ull + static_cast<uint64_t>(me);
      if (board.seen(reader, me) != want)
        THROW_INTERNAL_ERROR("ack board self test: wrong count from rank " + std::to_string(reader) + " in round " +
                             std::to_string(it));
    }
    barrier(*handle->comm);
  }
  board.destroy();
  API_CATCH()
}

static int32_t emitBoxes(const std::vector<BoxDesc>& push, const std::vector<BoxDesc>& unpack, cudecompB200Box_t* boxes,
                         int32_t max_boxes) {
  std::vector<BoxDesc> all = push;
  for (auto& u : unpack) all.push_back(u);
  int32_t n = 0;
  for (size_t i = 0; i < all.size() && n < max_boxes; ++i, ++n) {
    cudecompB200Box_t& o = boxes[n];
    o.peer_rank = all[i].peer_world;
    o.is_unpack = (i >= push.size()) ? 1 : 0;
    o.step = 0;
    o.reserved = 0;
    o.src_offset = all[i].src_off;
    o.dst_offset = all[i].dst_off;
    for (int k = 0; k < 3; ++k) {
      o.extent[k] = all[i].ext[k];
      o.src_stride[k] = all[i].sstr[k];
      o.dst_stride[k] = all[i].dstr[k];
    }
  }
  return static_cast<int32_t>(all.size());
}

// geometry of a config without a handle (same resolution rules as cudecompGridDescCreateVersioned)
static GridGeom geomFromConfig(const cudecompGridDescConfig_t* c) {
  if (!c) THROW_INVALID_USAGE("config argument cannot be null");
  if (c->pdims[0] < 1 || c->pdims[1] < 1) THROW_INVALID_USAGE("pdims values are invalid");
  // what cudecompGridDescCreate checks before a descriptor exists (these entry points take the raw config)
  checkMemOrder(c);
  for (int i = 0; i < 3; ++i) {
    if (c->gdims[i] < 1) THROW_INVALID_USAGE("gdims values must be positive");
    if (c->gdims_dist[i] < 0 || c->gdims_dist[i] > c->gdims[i])
      THROW_INVALID_USAGE("gdims_dist entries must be less than or equal to gdims entries");
  }
  GridGeom g;
  const bool dist_set = c->gdims_dist[0] != 0 && c->gdims_dist[1] != 0 && c->gdims_dist[2] != 0;
  const bool order_set = c->transpose_mem_order[0][0] >= 0;
  for (int i = 0; i < 3; ++i) {
    g.gdims[i] = c->gdims[i];
    g.gdims_dist[i] = dist_set ? c->gdims_dist[i] : c->gdims[i];
    for (int j = 0; j < 3; ++j)
      g.order[i][j] = order_set ? c->transpose_mem_order[i][j] : (c->transpose_axis_contiguous[i] ? (i + j) % 3 : j);
  }
  g.pdims = {c->pdims[0], c->pdims[1]};
  checkRankOrder(enumBits(c->rank_order));
  g.col_major = (enumBits(c->rank_order) == CUDECOMP_RANK_ORDER_COL_MAJOR);
  return g;
}

int32_t cudecompB200PlanTransposeBoxes(const cudecompGridDescConfig_t* config, int32_t rank, int32_t ax, int32_t dir,
                                       const int32_t input_halo_extents[], const int32_t output_halo_extents[],
                                       const int32_t input_padding[], const int32_t output_padding[], int32_t staged,
                                       cudecompB200Box_t* boxes, int32_t max_boxes) {
  try {
    GridGeom g = geomFromConfig(config);
    if (rank < 0 || rank >= g.pdims[0] * g.pdims[1]) THROW_INVALID_USAGE("rank out of range");
    if (ax < 0 || ax > 2 || dir == 0 || !boxes || max_boxes < 0) THROW_INVALID_USAGE("axis / direction / boxes arguments are invalid");
    if (staged == 2 || staged == 3) { // receiver-driven plans (3: through my workspace); peer_rank owns the SOURCE
      TransposePlan pull = buildPullTransposePlan(g, pidxOfRank(g, rank), ax, dir, input_halo_extents, output_halo_extents,
                                                  input_padding, output_padding,
                                                  staged == 3 ? DstKind::STAGE : DstKind::FINAL, false);
      return emitBoxes(pull.push, pull.unpack, boxes, max_boxes);
    }
    TransposePlan plan = buildTransposePlan(g, pidxOfRank(g, rank), ax, dir, input_halo_extents, output_halo_extents,
                                            input_padding, output_padding, staged ? DstKind::STAGE : DstKind::FINAL, false);
    return emitBoxes(plan.push, plan.unpack, boxes, max_boxes);
  } catch (const cdb::Error& e) {
    return -static_cast<int32_t>(e.code());
  } catch (const std::exception&) {
    return -static_cast<int32_t>(CUDECOMP_RESULT_INTERNAL_ERROR);
  }
}

int32_t cudecompB200PlanPipelinedTransposeBoxes(const cudecompGridDescConfig_t* config, int32_t rank, int32_t ax,
                                                int32_t dir, const int32_t input_halo_extents[],
                                                const int32_t output_halo_extents[], const int32_t input_padding[],
                                                const int32_t output_padding[], int32_t inplace, int32_t nchunks,
                                                cudecompB200Box_t* boxes, int32_t max_boxes) {
  try {
    GridGeom g = geomFromConfig(config);
    if (rank < 0 || rank >= g.pdims[0] * g.pdims[1]) THROW_INVALID_USAGE("rank out of range");
    if (ax < 0 || ax > 2 || dir == 0 || !boxes || max_boxes < 0) THROW_INVALID_USAGE("axis / direction / boxes arguments are invalid");
    // inplace: bit 0 = in place, bit 1 = receiver-driven (peer_rank of a push box then owns the SOURCE), bits 8-15 =
    // element size in bytes (0: plane chunks only; else column chunks where they apply, as the engine plans them),
    // bit 2 = column chunks whatever the row length (tests on small grids)
    PipelinedPlan pp = buildPipelinedTransposePlan(g, pidxOfRank(g, rank), ax, dir, input_halo_extents,
                                                   output_halo_extents, input_padding, output_padding, (inplace & 1) != 0,
                                                   nchunks, (inplace & 2) != 0, (inplace >> 8) & 0xff,
                                                   (inplace & 4) ? 1 : kMinChunkRowBytes);
    int32_t n = 0, total = 0;
    for (size_t s = 0; s < pp.steps.size(); ++s) {
      for (int pass = 0; pass < 2; ++pass) {
        const auto& list = pass == 0 ? pp.steps[s].push : pp.steps[s].unpack;
        for (const BoxDesc& bx : list) {
          ++total;
          if (n >= max_boxes) continue;
          cudecompB200Box_t& o = boxes[n++];
          o.peer_rank = bx.peer_world;
          o.is_unpack = pass;
          o.step = static_cast<int32_t>(s);
          o.reserved = 0;
          o.src_offset = bx.src_off;
          o.dst_offset = bx.dst_off;
          for (int k = 0; k < 3; ++k) {
            o.extent[k] = bx.ext[k];
            o.src_stride[k] = bx.sstr[k];
            o.dst_stride[k] = bx.dstr[k];
          }
        }
      }
    }
    return total;
  } catch (const cdb::Error& e) {
    return -static_cast<int32_t>(e.code());
  } catch (const std::exception&) {
    return -static_cast<int32_t>(CUDECOMP_RESULT_INTERNAL_ERROR);
  }
}

int32_t cudecompB200PlanHaloBoxes(const cudecompGridDescConfig_t* config, int32_t rank, int32_t ax, int32_t dim,
                                  const int32_t halo_extents[], const bool halo_periods[], const int32_t padding[],
                                  int32_t staged, cudecompB200Box_t* boxes, int32_t max_boxes) {
  try {
    GridGeom g = geomFromConfig(config);
    if (rank < 0 || rank >= g.pdims[0] * g.pdims[1]) THROW_INVALID_USAGE("rank out of range");
    if (ax < 0 || ax > 2 || dim < 0 || dim > 2 || !halo_extents || !boxes || max_boxes < 0)
      THROW_INVALID_USAGE("axis / dim / halo_extents / boxes arguments are invalid");
    HaloPlan plan = buildHaloPlan(g, pidxOfRank(g, rank), ax, dim, halo_extents, halo_periods, padding,
                                  staged ? DstKind::STAGE : DstKind::FINAL);
    if (plan.nothing) return 0;
    return emitBoxes(plan.push, plan.unpack, boxes, max_boxes);
  } catch (const cdb::Error& e) {
    return -static_cast<int32_t>(e.code());
  } catch (const std::exception&) {
    return -static_cast<int32_t>(CUDECOMP_RESULT_INTERNAL_ERROR);
  }
}

int32_t cudecompB200DescribeTransposeBoxes(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t ax,
                                           int32_t dir, const int32_t input_halo_extents[],
                                           const int32_t output_halo_extents[], const int32_t input_padding[],
                                           const int32_t output_padding[], int32_t staged,
                                           cudecompB200Box_t* boxes, int32_t max_boxes) {
  try {
    checkHandle(handle);
    checkGridDesc(handle, grid_desc);
    TransposePlan plan =
        buildTransposePlan(grid_desc->geom, grid_desc->pidx, ax, dir, input_halo_extents, output_halo_extents,
                           input_padding, output_padding, staged ? DstKind::STAGE : DstKind::FINAL, false);
    return emitBoxes(plan.push, plan.unpack, boxes, max_boxes);
  } catch (const std::exception& e) {
    std::cerr << e.what();
    return -1;
  }
}

int32_t cudecompB200DescribeHaloBoxes(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t ax, int32_t dim,
                                      const int32_t halo_extents[], const bool halo_periods[], const int32_t padding[],
                                      int32_t staged, cudecompB200Box_t* boxes, int32_t max_boxes) {
  try {
    checkHandle(handle);
    checkGridDesc(handle, grid_desc);
    HaloPlan plan = buildHaloPlan(grid_desc->geom, grid_desc->pidx, ax, dim, halo_extents, halo_periods, padding,
                                  staged ? DstKind::STAGE : DstKind::FINAL);
    if (plan.nothing) return 0;
    return emitBoxes(plan.push, plan.unpack, boxes, max_boxes);
  } catch (const std::exception& e) {
    std::cerr << e.what();
    return -1;
  }
}

} // extern "C"
