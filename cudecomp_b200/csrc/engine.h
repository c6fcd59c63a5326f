// Internal state behind the two opaque handles of the C ABI and the execution of one transpose / halo
// call (the orchestration layer the reference keeps in include/internal/transpose.h and halo.h).
#ifndef CUDECOMP_B200_ENGINE_H
#define CUDECOMP_B200_ENGINE_H

#include <array>
#include <cstdint>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "bootstrap.h"
#include "cudecomp.h"
#include "geometry.h"
#include "kernels.h"
#include "peer.h"
#include "perf_report.h"
#include "plan.h"

// Which way the last operation on a grid descriptor moved its data (cudecompB200GetLastPath).
enum {
  CUDECOMP_B200_PATH_NONE = 0,   // nothing to do (no-op transpose, zero halo, no neighbour)
  CUDECOMP_B200_PATH_LOCAL = 1,  // single-rank communicator: one local kernel
  CUDECOMP_B200_PATH_DIRECT = 2, // peer stores straight into the destination buffers, one kernel
  CUDECOMP_B200_PATH_STAGED = 3  // peer stores into the peers' workspace + local unpack kernel
};

namespace cdb {
// One cached phased launch: tables resident in device memory, uploaded once from a pinned host copy that stays alive
// with the entry (so the upload is a plain asynchronous copy and the steady state does no host work but the lookup).
struct FusedPlanEntry {
  std::string key;
  void* dev = nullptr;     // [KBox boxes...][SegDesc segments...][PhaseDesc phases...]
  void* host = nullptr;    // pinned copy
  size_t bytes = 0;
  PhasedParams params{};   // sync block filled per call
  cudaEvent_t uploaded = nullptr;      // recorded behind the upload: a call on ANOTHER stream waits for it
  cudaStream_t upload_stream = nullptr;
  uint64_t total_slots = 0;
  uint64_t last_use = 0;
};
} // namespace cdb

struct cudecompHandle {
  bool initialized = false;
  cdb::CommPtr comm;
  int rank = 0;
  int nranks = 1;
  bool have_device = false;
  int device = -1;
  int sm_count = 0;
  bool env_col_major = false;
  bool allow_direct = true;      // CUDECOMP_B200_DIRECT=0 forces staging through the workspace
  uint64_t spin_timeout_ns = 0;  // device-side wait limit
  uint64_t token = 0;            // names shared-memory segments
  int next_instance = 0;
  int live_grid_descs = 0;
  cdb::PeerCache peers;          // imported peer allocations, shared by all grid descriptors
  cdb::SignalArena arena;        // device flag pages, one slot per grid descriptor (created with the first one)
  std::vector<int> free_slots;   // recycled slots (create/destroy are collective, so all ranks agree)
  int next_slot = 0;
  cdb::PerfSettings perf;        // CUDECOMP_ENABLE_PERFORMANCE_REPORT and friends
  int pipeline_chunks = 0;       // CUDECOMP_B200_PIPELINE_CHUNKS: chunked schedule of staged transposes (0/1 = off)
  int kernel_variant = 0;        // CUDECOMP_B200_KERNEL=bulk -> 1: TMA bulk row copy where it applies (0 = LDG/STG)
  int tile_bytes = 0;            // CUDECOMP_B200_TILE_BYTES
  int peer_order = 0;            // CUDECOMP_B200_PEER_ORDER=pairwise -> 1
  int balance_grid = 0;          // CUDECOMP_B200_BALANCE_GRID=1
  int pull_mode = 0;             // CUDECOMP_B200_TRANSFER=pull
  int staged_mode = 0;           // CUDECOMP_B200_STAGED=launches -> 1 (separate push / unpack launches)
  int fused_lag = 1;             // CUDECOMP_B200_FUSED_LAG
  int phase_head_percent = 25;   // CUDECOMP_B200_PHASE_HEAD
  int wire_wide = 1;             // CUDECOMP_B200_WIRE_WIDE=0: 128-bit accesses on the wire too
  int column_chunks = 1;         // CUDECOMP_B200_COLUMN_CHUNKS=0: plane chunks only
  int transpose_geometry = 0;    // CUDECOMP_B200_TRANSPOSE_GEOM=1: 64 x 32 tiles for 8-byte vectorised transposes
  int cumem_state = 0;           // CUDECOMP_ENABLE_CUMEM outcome (vmm.h kCumem*): 1 = cudecompMalloc uses cuMem allocations
  bool cumem_fabric = false;     // ... created with fabric handles as well where the platform has them
  cdb::AckBoard acks;            // which of my release announcements every rank has processed
  struct Released {              // freed by the caller, still mapped by peers: the real cudaFree waits for their acks
    void* ptr;
    uint64_t stamp;              // my release_count right after announcing it
    std::vector<int> waiting_for;
  };
  std::vector<Released> released_pending;
  uint64_t release_count = 0;    // buffers freed through cudecompFree so far
  uint64_t released[cdb::kReleaseSlots] = {0}; // ids of the most recent ones, newest first
};

struct cudecompGridDesc {
  bool initialized = false;
  cudecompHandle_t handle = nullptr;
  cudecompGridDescConfig_t config{}; // normalised, orders and gdims_dist resolved
  bool gdims_dist_set = false;
  bool transpose_mem_order_set = false;
  cdb::GridGeom geom;
  std::array<int, 2> pidx{0, 0};
  int pad_slot = -1; // slot of handle->arena
  cdb::Mailbox mbox;
  uint64_t epoch = 0;
  std::set<void*> allocations; // from cudecompMalloc
  // tuning knobs (autotuner / cudecompB200SetTuning)
  int grid_ctas = 0;       // 0: all resident CTAs
  bool force_staged = false;
  int last_path = CUDECOMP_B200_PATH_NONE;
  bool warned_unmappable = false; // the "output cannot be mapped by the peers" warning is printed once per descriptor
  std::unique_ptr<cdb::PerfReport> perf; // only when the performance report is enabled
  // chunked (pipelined) staged schedule: unpack kernels run on a side stream beside the next chunk's push
  int pipeline_chunks = 0;
  int kernel_variant = 0; // 0: SIMT row copy, 1: TMA bulk row copy for 16-byte aligned rows of at least 2 KiB
  int tile_bytes = 0;     // row-copy tile size (0: 32 KiB), see launch_params.h
  int peer_order = 0;     // 0: slots interleaved over the peers (one-shot), 1: one peer after the other (pairwise rounds)
  int balance_grid = 0;   // 1: pick the CTA count whose last grid-stride round is fullest (kernels.h chooseGrid)
  int pull_mode = 0;      // 1: direct transposes are receiver-driven (each rank LOADS its blocks from the peers' inputs)
  cudaStream_t side_stream = nullptr;
  std::vector<cudaEvent_t> side_events;
  // Staged transposes (in place, forced staging, unmappable outputs): 0 = ONE phased launch that pushes chunk s while it
  // unpacks chunk s - lag (kernels.h PhasedParams; the default), 1 = separate launches (push, unpack; chunked: K pushes
  // on the caller's stream, unpacks on a side stream).
  int staged_mode = 0;
  int fused_lag = 1;
  int wire_wide = 1;           // 256-bit accesses in launches that store into peers (kernel_variant 0 only)
  int column_chunks = 1;       // fused staged schedule: chunk along the fastest axis when it takes no part (plan.cc)
  int phase_head_percent = 25; // share of a step's pushes that runs before the unpacks of the earlier chunk join in
  std::vector<cdb::FusedPlanEntry> fused_cache; // device tables of phased launches, keyed by everything they depend on
};

namespace cdb {

// chunks of the fused staged schedule when pipeline_chunks == 0 (auto): at least 256 MiB of pencil per chunk, 1..16
int autoFusedChunks(int64_t pencil_bytes);

void releaseFusedCache(cudecompGridDesc_t gd);
// cudaFree every released allocation whose importers have all acknowledged (all of them when `everything`).
void reapReleased(cudecompHandle_t h, bool everything = false);
// Collective over the handle's communicator: every rank processes every rank's release announcements, then the owners
// free what was pending (used where the API is collective anyway: cudecompGridDescDestroy).
void drainReleases(cudecompHandle_t h);
uint64_t epochStride(const cudecompGridDesc_t gd);

void setGeometry(cudecompGridDesc_t gd, const std::array<int32_t, 2>& pdims);

int64_t dtypeSize(cudecompDataType_t dtype);

void runTranspose(cudecompHandle_t h, cudecompGridDesc_t gd, int ax, int dir, void* input, void* output, void* work,
                  cudecompDataType_t dtype, const int32_t in_halo[], const int32_t out_halo[], const int32_t in_pad[],
                  const int32_t out_pad[], cudaStream_t stream);

void runHalo(cudecompHandle_t h, cudecompGridDesc_t gd, int ax, void* input, void* work, cudecompDataType_t dtype,
             const int32_t halo[], const bool periods[], int dim, const int32_t pad[], cudaStream_t stream);

// Raises INTERNAL_ERROR if a device-side wait of an earlier operation timed out.
void checkDeviceError(cudecompGridDesc_t gd);

// grid-descriptor autotuning (autotune.cc)
std::vector<int> autotuneTransposeBackendCandidates(const cudecompGridDescAutotuneOptions_t* options);
std::vector<int> autotuneHaloBackendCandidates(const cudecompGridDescAutotuneOptions_t* options);
std::vector<std::array<int32_t, 2>> autotunePdimCandidates(int nranks, bool col_major);
void autotune(cudecompHandle_t h, cudecompGridDesc_t gd, const cudecompGridDescAutotuneOptions_t* options);

} // namespace cdb

#endif
