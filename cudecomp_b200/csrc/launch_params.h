// From transfer boxes to kernel launch parameters: kernel selection, vector width, tiling. Pure host arithmetic on
// whatever base pointers it is given, so the same code prepares real launches (engine.cc) and the launches that the
// host-side emulator under tests/host_emu/ walks on numpy buffers.
#ifndef CUDECOMP_B200_LAUNCH_PARAMS_H
#define CUDECOMP_B200_LAUNCH_PARAMS_H

#include <vector>

#include "kernels.h"
#include "plan.h"

namespace cdb {

struct LaunchBox {
  BoxDesc d;
  const char* src_base;
  char* dst_base;
};

// Per-descriptor schedule knobs (cudecompB200SetSchedule / SetKernelVariant); the autotuner sweeps them.
struct LaunchTuning {
  int tile_bytes = 0;     // bytes of a ROWCOPY tile (0: kDefaultTileBytes); power of two in [4 KiB, 256 KiB]
  int peer_order = 0;     // CopyParams::peer_order
  int transpose_geometry = 0;  // vectorised transpose of 8-byte elements: 0 = 32 x 64 tiles, 1 = 64 x 32 (tiling.h)
  int phase_head_percent = 25; // phased launches: share of a step's pushes that runs before its unpacks join in
  int kernel_variant = 0; // 1: TMA bulk row copy where every row is 16-byte aligned and at least 2 KiB long;
                          // 2: 256-bit LDG/STG where every address and stride is 32-byte aligned (else 128-bit)
};

constexpr int kDefaultTileBytes = 32768;
constexpr int kMinTileBytes = 4096;
constexpr int kMaxTileBytes = 262144;

struct PreparedLaunch {
  KernelKind kind;
  CopyParams params; // sync block zeroed: the caller fills it in
};

// Splits `boxes` (all of element size `es`) into launches of at most kMaxBoxes boxes. Empty boxes are dropped; the
// result always holds at least one launch (possibly with nboxes == 0) because a launch also carries the handshake.
// `me` / `comm_size`: this rank's index in the communicator the boxes' `peer` fields refer to; only used to order the
// boxes for peer_order == 1 (self first, then me+1, me+2, ... cyclically). Pass comm_size <= 1 to keep the given order.
std::vector<PreparedLaunch> prepareLaunches(const std::vector<LaunchBox>& boxes, int es, const LaunchTuning& tuning, int me,
                                            int comm_size);

// Tables of a phased launch (kernels.h PhasedParams), host copies.
struct PhasedLaunch {
  std::vector<KBox> boxes;
  std::vector<SegDesc> segs;
  std::vector<PhaseDesc> phases;
  uint32_t nsteps = 0; // chunks (each published to the peers once)
  int vec_size = 16;
  uint64_t total_slots = 0;
};

// push[s] / unpack[s]: the boxes of step s of a chunked staged schedule (plan.h PipelinedPlan), resolved to base
// pointers. Step s of the launch holds push[s] and unpack[s - lag] (which waits for step s - lag), lag >= 1; a step is
// one phase, or two when a head of pushes runs before the unpacks join (LaunchTuning::phase_head_percent). Returns false when the schedule cannot run as one phased launch (a box that is not a row
// copy, too many phases); the caller then falls back to separate launches.
bool preparePhased(const std::vector<std::vector<LaunchBox>>& push, const std::vector<std::vector<LaunchBox>>& unpack, int es,
                   const LaunchTuning& tuning, int lag, PhasedLaunch* out);

} // namespace cdb

#endif
