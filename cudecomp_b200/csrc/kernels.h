// Device side of the hot path: box-copy kernels that read local HBM and store into (possibly peer)
// memory, with the cross-GPU entry/exit handshake folded into the same launch.
//
// Replaces reference include/internal/cudecomp_kernels.cuh:125-270 (the batched strided copy used for
// pack and unpack), the cuTENSOR permute call in include/internal/transpose.h:80-157, and the exchange
// backends of include/internal/comm_routines.h:260-425 (nearest prior art: the SM-driven NVSHMEM put
// kernel, cudecomp_kernels.cuh:86-122).
#ifndef CUDECOMP_B200_KERNELS_H
#define CUDECOMP_B200_KERNELS_H

#include <cuda_runtime.h>

#include <cstdint>

namespace cdb {

constexpr int kMaxBoxes = 16; // boxes per launch (one per peer of a row/column communicator)
constexpr int kMaxPeers = 16; // peers a launch can handshake with

// Signal pad: one 4 KiB page per grid descriptor on every rank, mapped into every peer.
// Slots are indexed by GLOBAL rank and hold monotonically increasing epochs, so they never need a reset.
constexpr int kPadMaxRanks = 128;
constexpr int kPadEntry = 0;                    // [r]: rank r has reached operation `epoch` in its stream
constexpr int kPadExit = kPadMaxRanks;          // [r]: all of rank r's stores for operation `epoch` have landed here
constexpr int kPadCounter = 2 * kPadMaxRanks;   // local: CTAs finished in the running launch
constexpr int kPadWords = 2 * kPadMaxRanks + 8; // uint64 words

struct KBox {
  const char* src;
  char* dst;
  // ROWCOPY: n[0] = row length, n[1], n[2] = row grid. TRANSPOSE: n[0] = source-contiguous axis,
  // n[1] = destination-contiguous axis, n[2] = the rest. Strides in elements.
  int64_t n[3];
  int64_t ss[3];
  int64_t ds[3];
  // tiling chosen on the host
  uint32_t tiles;         // tiles of this box
  uint32_t row_vecs;      // ROWCOPY: vectors per row
  uint32_t seg_vecs;      // ROWCOPY: vectors per row segment of a tile
  uint32_t segs_per_row;  // ROWCOPY
  uint32_t rows_per_tile; // ROWCOPY
  uint32_t tiles0;        // TRANSPOSE: tiles along n[0]
  uint32_t tiles1;        // TRANSPOSE: tiles along n[1]
  uint32_t pad_;
};

struct SyncParams {
  uint64_t* my_pad;              // local signal pad (nullptr: no handshake in this launch)
  uint64_t* peer_pad[kMaxPeers]; // peers' pads, peer-mapped
  int32_t peer_world[kMaxPeers]; // their global ranks
  int32_t npeers;
  int32_t my_world;
  uint64_t epoch;
  uint32_t do_entry; // signal "arrived" to the peers and wait for theirs before the first remote store
  uint32_t do_exit;  // after the last CTA: signal "done" to the peers and wait for theirs
  uint32_t* error_word;    // host-mapped; set to a nonzero code when a wait times out
  uint64_t timeout_ns;
};

struct CopyParams {
  KBox box[kMaxBoxes];
  SyncParams sync;
  uint32_t nboxes;
  uint32_t max_tiles; // max over boxes; slot t of the launch -> (box, tile) by slotToBoxTile (tiling.h)
  uint32_t elem_size; // 4, 8 or 16
  uint32_t vec_size;  // ROWCOPY: 4, 8, 16 or 32
  uint32_t peer_order; // slot order: 0 interleaved over the boxes (one-shot), 1 rounds, one peer after the other (pairwise)
  uint32_t pad_[3];
};

// ROWCOPY_BULK: the row copy driven by the TMA unit instead of LDG/STG: one elected thread per CTA moves row segments
// global -> shared -> global with cp.async.bulk and an mbarrier ring. Same boxes and tiling fields as ROWCOPY, but a
// tile is ONE row segment of at most kBulkChunkBytes (rows_per_tile = 1) and everything must be 16-byte aligned.
enum class KernelKind { ROWCOPY, TRANSPOSE, ROWCOPY_BULK };

constexpr int kBulkStages = 4;
constexpr uint32_t kBulkChunkBytes = 16384;

struct LaunchConfig {
  int grid = 0;    // CTAs (0: library default, see defaultGrid)
  int threads = 256;
  int balance = 0; // 1: shrink the grid (by at most 20 %) to the CTA count whose last round of slots is fullest
};

// CTA count of a launch: `requested` (0: `dflt`), capped by the resident CTAs and the slot count; with balance != 0
// the count in [0.8 * that, that] that wastes the least of its last grid-stride round (ties: the larger count).
int chooseGrid(int requested, int dflt, int resident, uint64_t total_slots, int balance);

// Enqueues the copy described by `p` on `stream`. Returns the CUDA status of the launch.
cudaError_t launchCopy(KernelKind kind, const CopyParams& p, const LaunchConfig& cfg, cudaStream_t stream);

// Upper bound on co-resident CTAs of the given kernel on the current device.
int maxResidentCtas(KernelKind kind, int vec_or_elem_size, int threads, uint32_t peer_order = 0);

// counts launches issued through launchCopy (bench.py reports it as gpu_launches)
uint64_t launchCount();

} // namespace cdb

#endif
