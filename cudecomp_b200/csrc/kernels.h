// Device side of the hot path: box-copy kernels that read local HBM and store into (possibly peer)
// memory, with the cross-GPU entry/exit handshake folded into the same launch.
//
// Replaces reference include/internal/cudecomp_kernels.cuh:125-270 (the batched strided copy used for
// pack and unpack), the cuTENSOR permute call in include/internal/transpose.h:80-157, and the exchange
// backends of include/internal/comm_routines.h:260-425 (nearest prior art: the SM-driven NVSHMEM put
// kernel, cudecomp_kernels.cuh:86-122), and -- the phased launch -- the pipelined exchange that overlaps per-peer
// pack / send / unpack on auxiliary streams (comm_routines.h:427-631, cudecompAlltoallPipelined).
#ifndef CUDECOMP_B200_KERNELS_H
#define CUDECOMP_B200_KERNELS_H

#include <cuda_runtime.h>

#include <cstdint>

namespace cdb {

constexpr int kMaxBoxes = 16; // boxes per launch (one per peer of a row/column communicator)
constexpr int kMaxPeers = 71; // peers a launch can handshake with: a row / column communicator of up to 72 ranks (one
                              // NVLink domain); the parameter block stays under the 4 KiB every launch path accepts

// Signal pad: one 4 KiB page per grid descriptor on every rank, mapped into every peer.
// Slots are indexed by GLOBAL rank and hold monotonically increasing epochs, so they never need a reset.
constexpr int kPadMaxRanks = 128;
constexpr int kPadEntry = 0;                    // [r]: rank r has reached operation `epoch` in its stream
constexpr int kPadExit = kPadMaxRanks;          // [r]: all of rank r's stores for operation `epoch` have landed here
constexpr int kPadCounter = 2 * kPadMaxRanks;   // local: CTAs finished in the running launch
constexpr int kPadStep = 2 * kPadMaxRanks + 8;  // [r]: phased launches, rank r's pushes of (epoch, step) have landed here
constexpr int kPadPhaseCounter = 3 * kPadMaxRanks + 8; // local: per phase of a phased launch, CTAs that finished it
constexpr int kMaxPhases = 64;                  // phases of one phased launch (chunks + lag)
constexpr int kPadWords = 3 * kPadMaxRanks + 8 + kMaxPhases; // uint64 words (must fit the 4 KiB slot)

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
  uint32_t geometry;   // TRANSPOSE_VEC: tile geometry (tiling.h TransVecGeom kAlt)
  uint32_t pad_[2];
};

// Phased launch (the fused in-place / staged schedule, see engine.cc runFusedStaged): ONE persistent launch walks
// `nphases` phases in order. Phase s holds the push boxes of chunk s (stores into the peers' workspaces) and the unpack
// boxes that may run once chunk `wait_step` has been exchanged by EVERY member (workspace -> my pencil). Slots of a
// phase interleave its boxes, so link-bound pushes and HBM-bound unpacks are in flight together on every SM.
// After its share of a phase a CTA fences and bumps the phase counter; the last one publishes (epoch, s) to the peers'
// kPadStep slot. A CTA that reaches a box with wait_step >= 0 first waits until every peer has published that step
// and every local CTA has finished it (all local reads of that chunk are done, all remote data of it has landed).
// The boxes of a phase differ widely in size (a push box is 1/K of a peer's share, an unpack piece may be 1/16 of
// that), so a phase is cut into SEGMENTS of at most kSegTiles tiles of one box and its slots interleave the segments:
// slot t -> segment t % nsegs, tile t / nsegs of it. Every window of nsegs slots then holds the boxes in proportion to
// their sizes, and at most one segment per box is partly empty.
// The tables live in device memory (too many boxes for the parameter space).
constexpr uint32_t kSegTiles = 64;

struct alignas(16) SegDesc {
  uint32_t box;        // index into PhasedParams::boxes
  uint32_t first_tile; // of that box
  uint32_t count;      // tiles in this segment (<= kSegTiles)
  uint32_t wait;       // wait_step + 1 of the box (0: no dependency)
};

struct alignas(16) PhaseDesc {
  uint32_t first_seg;
  uint32_t nsegs;
  uint32_t seg_tiles; // slots of the phase = nsegs * seg_tiles (seg_tiles = longest segment of the phase)
  uint32_t publish;   // step + 1 whose pushes are complete (on this CTA) at the end of this phase, 0: none
};

struct PhasedParams {
  const KBox* boxes;       // device table; KBox::pad_ holds wait_step + 1 (0: no dependency)
  const SegDesc* segs;     // device table
  const PhaseDesc* phases; // device table
  SyncParams sync;         // do_exit is ignored: the last phase's waits subsume the exit handshake
  uint32_t nphases;
  uint32_t nsteps;         // chunks: steps [0, nsteps) are published to the peers, one counter each
  uint32_t elem_size;
  uint32_t vec_size;
};

// ROWCOPY_BULK: the row copy driven by the TMA unit instead of LDG/STG: one elected thread per CTA moves row segments
// global -> shared -> global with cp.async.bulk and an mbarrier ring. Same boxes and tiling fields as ROWCOPY, but a
// tile is ONE row segment of at most kBulkChunkBytes (rows_per_tile = 1) and everything must be 16-byte aligned.
// TRANSPOSE_VEC: the transpose with 16-byte global accesses on both sides (micro-tiles of VEC x VEC elements, VEC =
// 16 / element size, are transposed in registers and exchanged through a swizzled shared-memory tile of 16-byte vectors;
// see tiling.h TransVecGeom). Needs 16-byte aligned rows on both sides and extents that are multiples of VEC; other
// launches keep the element-wise TRANSPOSE kernel.
enum class KernelKind { ROWCOPY, TRANSPOSE, ROWCOPY_BULK, TRANSPOSE_VEC };

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

// Enqueues a phased row-copy launch (all boxes ROWCOPY-shaped, vector width p.vec_size). `total_slots`: sum over the
// phases of nboxes * max_tiles (sizes the grid).
cudaError_t launchPhased(const PhasedParams& p, uint64_t total_slots, const LaunchConfig& cfg, cudaStream_t stream);

// counts launches issued through launchCopy (bench.py reports it as gpu_launches)
uint64_t launchCount();

} // namespace cdb

#endif
