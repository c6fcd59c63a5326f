// Execution of one transpose / halo call. See engine.h, plan.h, kernels.h.
#include "engine.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "errors.h"
#include "launch_params.h"
#include "nvtx_ranges.h"
#include "vmm.h"

namespace cdb {

int64_t dtypeSize(cudecompDataType_t dtype) {
  switch (dtype) {
  case CUDECOMP_FLOAT: return 4;
  case CUDECOMP_DOUBLE:
  case CUDECOMP_FLOAT_COMPLEX: return 8;
  case CUDECOMP_DOUBLE_COMPLEX: return 16;
  }
  THROW_INVALID_USAGE("unknown data type");
}

// Epochs one collective call owns: (epoch - stride, epoch]. Single-launch paths use the last one, the chunked
// separate-launch schedule one per step. A function of the descriptor's settings only, never of the path taken.
uint64_t epochStride(const cudecompGridDesc_t gd) { return static_cast<uint64_t>(std::max(1, gd->pipeline_chunks)); }

// Measured on B200 (profiles/r2_n2_schedules.md): a chunk costs about 10 us of pipeline bubbles (CTAs drift apart by a
// tile within a step and wait for the slowest one anywhere before they unpack), so chunks are kept at >= 256 MiB of
// pencil: 8 chunks for the 2 GiB pencils of 1024^3 complex128 on 8 GPUs (9.01 ms per round trip against 9.34 with 16 and
// 10.5 with 32, profiles/r2_n8_*.json), 2 for the 512 MiB ones of 512^3 complex64 on 2 GPUs.
int autoFusedChunks(int64_t pencil_bytes) {
  const int64_t k = pencil_bytes / (int64_t(256) << 20);
  return static_cast<int>(std::min<int64_t>(std::max<int64_t>(k, 1), 16));
}

void releaseFusedCache(cudecompGridDesc_t gd) {
  if (gd->fused_cache.empty()) return;
  cudaDeviceSynchronize(); // a running phased launch may still read its tables
  for (auto& e : gd->fused_cache) {
    if (e.dev) cudaFree(e.dev);
    if (e.host) cudaFreeHost(e.host);
    if (e.uploaded) cudaEventDestroy(e.uploaded);
  }
  gd->fused_cache.clear();
  (void)cudaGetLastError();
}

void reapReleased(cudecompHandle_t h, bool everything) {
  auto& list = h->released_pending;
  for (size_t i = 0; i < list.size();) {
    bool ready = true;
    if (!everything)
      for (int r : list[i].waiting_for)
        if (h->acks.seen(r, h->rank) < list[i].stamp) ready = false;
    if (ready) {
      if (!vmmFree(list[i].ptr)) cudaFree(list[i].ptr);
      list.erase(list.begin() + i);
    } else {
      ++i;
    }
  }
  (void)cudaGetLastError();
}

void drainReleases(cudecompHandle_t h) {
  if (h->nranks == 1 || !h->have_device) {
    reapReleased(h, true);
    return;
  }
  struct Note {
    uint64_t release_count;
    uint64_t released[kReleaseSlots];
  } mine;
  mine.release_count = h->release_count;
  for (int k = 0; k < kReleaseSlots; ++k) mine.released[k] = h->released[k];
  std::vector<Note> all(h->nranks);
  allgather(*h->comm, &mine, sizeof(Note), all.data());
  bool any = false;
  for (int r = 0; r < h->nranks; ++r) {
    if (all[r].release_count) any = true;
    if (r != h->rank && h->peers.noteReleases(r, all[r].release_count, all[r].released)) h->acks.publish(r, all[r].release_count);
  }
  if (!any) return; // nobody has ever released anything: nothing can be pending anywhere
  barrier(*h->comm);
  reapReleased(h);
}

void setGeometry(cudecompGridDesc_t gd, const std::array<int32_t, 2>& pdims) {
  releaseFusedCache(gd); // plans depend on the process grid
  gd->config.pdims[0] = pdims[0];
  gd->config.pdims[1] = pdims[1];
  GridGeom& g = gd->geom;
  for (int i = 0; i < 3; ++i) {
    g.gdims[i] = gd->config.gdims[i];
    g.gdims_dist[i] = gd->config.gdims_dist[i];
    for (int j = 0; j < 3; ++j) g.order[i][j] = gd->config.transpose_mem_order[i][j];
  }
  g.pdims = pdims;
  g.col_major = (gd->config.rank_order == CUDECOMP_RANK_ORDER_COL_MAJOR);
  gd->pidx = pidxOfRank(g, gd->handle->rank);
}

void checkDeviceError(cudecompGridDesc_t gd) {
  SignalArena& arena = gd->handle->arena;
  if (!arena.valid()) return;
  const uint32_t e = arena.errorWordHost();
  if (e == 0) return;
  // launches that gave up left without their bookkeeping: drain the device and zero this rank's counters
  cudaDeviceSynchronize();
  if (gd->pad_slot >= 0) {
    cudaMemset(arena.mine(gd->pad_slot) + kPadCounter, 0, 8 * sizeof(uint64_t));
    cudaMemset(arena.mine(gd->pad_slot) + kPadPhaseCounter, 0, kMaxPhases * sizeof(uint64_t));
  }
  (void)cudaGetLastError();
  arena.clearError();
  if (e == 3) THROW_INTERNAL_ERROR("a TMA bulk copy of an earlier operation did not complete within 20 s");
  THROW_INTERNAL_ERROR(std::string("a device-side wait for a peer rank timed out during an earlier operation (") +
                       (e == 1 ? "entry" : (e == 2 ? "exit" : "chunk")) +
                       " handshake); a rank of the communicator did not enter the same operation");
}

namespace {

using ResolvedBox = LaunchBox;

// Fills the per-launch handshake block. `peers` are global ranks other than mine.
SyncParams makeSync(cudecompGridDesc_t gd, const std::vector<int>& peers) {
  SyncParams s;
  std::memset(&s, 0, sizeof(s));
  if (peers.empty()) return s;
  SignalArena& arena = gd->handle->arena;
  if (!arena.valid() || gd->pad_slot < 0) THROW_INTERNAL_ERROR("signal pads are not initialised");
  if (peers.size() > static_cast<size_t>(kMaxPeers))
    THROW_NOT_SUPPORTED("row / column communicators with more than 72 ranks are not supported");
  s.my_pad = arena.mine(gd->pad_slot);
  s.npeers = static_cast<int32_t>(peers.size());
  for (size_t i = 0; i < peers.size(); ++i) {
    s.peer_pad[i] = arena.of(peers[i], gd->pad_slot);
    s.peer_world[i] = peers[i];
  }
  s.my_world = gd->handle->rank;
  s.epoch = gd->epoch;
  s.do_entry = 1;
  s.do_exit = 1;
  s.error_word = arena.errorWordDevice();
  s.timeout_ns = gd->handle->spin_timeout_ns;
  return s;
}

// Enqueue the copy of `boxes` (all with the same element size). `sync` handshakes with peers (may be empty).
// `me` / `comm_size`: my index in the communicator the boxes' peers belong to (orders the boxes for the pairwise
// slot order; -1 / 0 for purely local launches).
void launchBoxes(cudecompGridDesc_t gd, const std::vector<ResolvedBox>& boxes, int es, const SyncParams& sync,
                 cudaStream_t stream, int me = -1, int comm_size = 0) {
  LaunchTuning tuning;
  tuning.tile_bytes = gd->tile_bytes;
  tuning.peer_order = gd->peer_order;
  tuning.kernel_variant = gd->kernel_variant;
  // Launches that store into peers: 256-bit accesses where the alignment allows (measured on B200, profiles/
  // r2_n2_schedules.md: +3 % on the wire; local HBM-bound copies are 8 % SLOWER with them and keep 128-bit accesses)
  if (tuning.kernel_variant == 0 && sync.npeers > 0 && gd->wire_wide) tuning.kernel_variant = 2;
  tuning.transpose_geometry = gd->handle->transpose_geometry;
  std::vector<PreparedLaunch> launches = prepareLaunches(boxes, es, tuning, me, comm_size);

  LaunchConfig cfg;
  cfg.grid = gd->grid_ctas;
  cfg.balance = gd->balance_grid;

  for (size_t l = 0; l < launches.size(); ++l) {
    CopyParams& p = launches[l].params;
    p.sync = sync;
    p.sync.do_entry = (sync.npeers > 0 && sync.do_entry && l == 0) ? 1 : 0;
    p.sync.do_exit = (sync.npeers > 0 && sync.do_exit && l + 1 == launches.size()) ? 1 : 0;
    if (p.nboxes == 0 && sync.npeers == 0) continue;
    cudaError_t err = launchCopy(launches[l].kind, p, cfg, stream);
    if (err != cudaSuccess) THROW_CUDA_ERROR(std::string("kernel launch failed: ") + cudaGetErrorString(err));
  }
}

void requireDevice(cudecompHandle_t h) {
  if (!h->have_device)
    CDB_THROW(CUDECOMP_RESULT_CUDA_ERROR, "CUDA error.", "no CUDA device is available to this process");
}

void fillReleases(cudecompHandle_t h, CallMsg* m) {
  m->release_count = h->release_count;
  for (int k = 0; k < kReleaseSlots; ++k) m->released[k] = h->released[k];
}

void processReleases(cudecompHandle_t h, const std::vector<int>& group_world, int me, const std::vector<CallMsg>& msgs) {
  for (size_t i = 0; i < msgs.size(); ++i)
    if (static_cast<int>(i) != me &&
        h->peers.noteReleases(group_world[i], msgs[i].release_count, msgs[i].released))
      h->acks.publish(group_world[i], msgs[i].release_count); // my imports of what that rank freed are closed
  if (!h->released_pending.empty()) reapReleased(h);
}

// Ends the current performance sample on every way out of a call (including exceptions).
struct PerfGuard {
  PerfSample* sample;
  cudaStream_t stream;
  bool exchange = false;
  ~PerfGuard() { PerfReport::end(sample, exchange, stream); }
};

bool isManaged(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeManaged;
}

uint32_t transposeOpcode(int ax, int dir) { return 0x100u + static_cast<uint32_t>(ax) * 4u + (dir > 0 ? 1u : 0u); }
uint32_t haloOpcode(int ax, int dim) { return 0x200u + static_cast<uint32_t>(ax) * 4u + static_cast<uint32_t>(dim); }

// Chunked schedule of the staged path (plan.h PipelinedPlan): K push launches on the caller's stream, each with its
// own handshake epoch; the unpack pieces that become writable after push s run on a side stream beside push s+1.
// The caller's stream rejoins the side stream at the end, so stream semantics are unchanged. Returns false when
// chunking does not apply (the caller then runs the unchunked schedule). The steps use the sub-epochs the call owns
// (epochStride): the descriptor's epoch itself advances by the same amount on every rank whichever path a row or column
// group takes (the path depends on that group's buffers, the epoch is shared by both communicators).
bool runPipelinedStaged(cudecompHandle_t h, cudecompGridDesc_t gd, int ax, int dir, void* input, void* output, void* work,
                        int es, const int32_t in_halo[], const int32_t out_halo[], const int32_t in_pad[],
                        const int32_t out_pad[], bool inplace, const std::vector<CallMsg>& msgs,
                        const std::vector<int>& peers, PerfSample* perf, cudaStream_t stream, bool pull) {
  PipelinedPlan pp = buildPipelinedTransposePlan(gd->geom, gd->pidx, ax, dir, in_halo, out_halo, in_pad, out_pad, inplace,
                                                 gd->pipeline_chunks, pull);
  if (pp.steps.empty()) return false;
  const size_t K = pp.steps.size();
  if (!gd->side_stream) {
    int lo = 0, hi = 0;
    CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CHECK_CUDA(cudaStreamCreateWithPriority(&gd->side_stream, cudaStreamNonBlocking, hi));
  }
  while (gd->side_events.size() < K + 1) {
    cudaEvent_t e;
    CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    gd->side_events.push_back(e);
  }
  SyncParams nosync;
  std::memset(&nosync, 0, sizeof(nosync));
  bool used_side = false;
  const uint64_t first_epoch = gd->epoch - epochStride(gd) + 1; // K <= stride
  for (size_t s = 0; s < K; ++s) {
    SyncParams sync = makeSync(gd, peers);
    sync.epoch = first_epoch + s;
    // Peers only have to be met on the way in once per call: after step 0 everybody is inside this operation and the
    // chunks land in disjoint parts of the workspace. The way out is needed per step (unpack(s) reads what peers pushed).
    if (s > 0) sync.do_entry = 0;
    std::vector<ResolvedBox> push, unpack;
    for (auto& b : pp.steps[s].push) {
      if (pull) { // load the slice from its owner's input into my workspace
        const char* src = (b.peer == pp.base.me) ? static_cast<const char*>(input)
                                                 : static_cast<const char*>(h->peers.resolve(b.peer_world, msgs[b.peer].src));
        push.push_back({b, src, static_cast<char*>(work)});
        continue;
      }
      char* dst = (b.peer == pp.base.me) ? static_cast<char*>(work)
                                         : static_cast<char*>(h->peers.resolve(b.peer_world, msgs[b.peer].work));
      push.push_back({b, static_cast<const char*>(input), dst});
    }
    launchBoxes(gd, push, es, sync, stream, pp.base.me, pp.base.comm_size);
    if (s + 1 == K) PerfReport::markExchangeDone(perf, stream);
    if (pp.steps[s].unpack.empty()) continue;
    for (auto& b : pp.steps[s].unpack) unpack.push_back({b, static_cast<const char*>(work), static_cast<char*>(output)});
    CHECK_CUDA(cudaEventRecord(gd->side_events[s], stream));
    CHECK_CUDA(cudaStreamWaitEvent(gd->side_stream, gd->side_events[s], 0));
    launchBoxes(gd, unpack, es, nosync, gd->side_stream);
    used_side = true;
  }
  if (used_side) {
    CHECK_CUDA(cudaEventRecord(gd->side_events[K], gd->side_stream));
    CHECK_CUDA(cudaStreamWaitEvent(stream, gd->side_events[K], 0));
  }
  return true;
}


// Fused staged schedule: ONE phased launch (kernels.h PhasedParams) pushes chunk s into the peers' workspaces while it
// unpacks chunk s - lag from mine, with per-chunk flags instead of per-chunk launches. Replaces the pack / a2a / unpack
// pipeline of reference include/internal/comm_routines.h:427-631 (cudecompAlltoallPipelined) for in-place and other
// staged calls. Returns false when the schedule does not apply (boxes that are not row copies: differing memory
// orders); the caller then runs separate launches. Every member takes the same decision: it depends on the geometry,
// the layouts and K only... and on buffer alignment, which can only lower the vector width, never the shape.
bool runFusedStaged(cudecompHandle_t h, cudecompGridDesc_t gd, int ax, int dir, void* input, void* output, void* work, int es,
                    const int32_t in_halo[], const int32_t out_halo[], const int32_t in_pad[], const int32_t out_pad[],
                    bool inplace, const std::vector<CallMsg>& msgs, const TransposePlan& probe, const SyncParams& sync,
                    PerfSample* perf, cudaStream_t stream) {
  const int P = probe.comm_size;
  // Rank-independent decisions only (every member must run the same number of phases): equal memory orders make every
  // box a row copy whatever its extents; the chunk count comes from the average pencil size.
  for (int k = 0; k < 3; ++k)
    if (gd->geom.order[probe.axes.a][k] != gd->geom.order[probe.axes.b][k]) return false;
  int K = gd->pipeline_chunks;
  if (K <= 0)
    K = autoFusedChunks(static_cast<int64_t>(gd->geom.gdims[0]) * gd->geom.gdims[1] * gd->geom.gdims[2] /
                        std::max(1, gd->handle->nranks) * es);
  const int lag = std::max(1, gd->fused_lag);
  K = std::min(K, kMaxPhases - lag);

  // resolve the peers' workspaces first: they are part of the cache key (a peer may have re-allocated)
  std::vector<char*> peer_work(P, nullptr);
  for (int i = 0; i < P; ++i)
    peer_work[i] = (i == probe.me) ? static_cast<char*>(work)
                                   : static_cast<char*>(h->peers.resolve(probe.group_world[i], msgs[i].work));

  std::string key;
  auto put = [&](const void* p, size_t n) { key.append(static_cast<const char*>(p), n); };
  const int32_t zero3[3] = {0, 0, 0};
  const int32_t head[10] = {ax, dir, es, K, lag, gd->tile_bytes, inplace ? 1 : 0, P, gd->phase_head_percent,
                            gd->kernel_variant * 4 + gd->wire_wide * 2 + gd->column_chunks};
  put(head, sizeof(head));
  put(&input, sizeof(input));
  put(&output, sizeof(output));
  put(in_halo ? in_halo : zero3, 12);
  put(out_halo ? out_halo : zero3, 12);
  put(in_pad ? in_pad : zero3, 12);
  put(out_pad ? out_pad : zero3, 12);
  put(peer_work.data(), peer_work.size() * sizeof(char*));

  static uint64_t tick = 0;
  FusedPlanEntry* entry = nullptr;
  for (auto& e : gd->fused_cache)
    if (e.key == key) entry = &e;
  if (!entry) {
    PipelinedPlan pp;
    if (K > 1)
      pp = buildPipelinedTransposePlan(gd->geom, gd->pidx, ax, dir, in_halo, out_halo, in_pad, out_pad, inplace, K, false,
                                       gd->column_chunks ? es : 0);
    std::vector<std::vector<ResolvedBox>> push, unpack;
    if (K > 1 && !pp.steps.empty()) {
      push.resize(pp.steps.size());
      unpack.resize(pp.steps.size());
      for (size_t s = 0; s < pp.steps.size(); ++s) {
        for (auto& b : pp.steps[s].push) push[s].push_back({b, static_cast<const char*>(input), peer_work[b.peer]});
        for (auto& b : pp.steps[s].unpack) unpack[s].push_back({b, static_cast<const char*>(work), static_cast<char*>(output)});
      }
    } else {
      TransposePlan st = buildTransposePlan(gd->geom, gd->pidx, ax, dir, in_halo, out_halo, in_pad, out_pad, DstKind::STAGE, inplace);
      push.resize(1);
      unpack.resize(1);
      for (auto& b : st.push) push[0].push_back({b, static_cast<const char*>(input), peer_work[b.peer]});
      for (auto& b : st.unpack) unpack[0].push_back({b, static_cast<const char*>(work), static_cast<char*>(output)});
    }
    LaunchTuning tuning;
    tuning.tile_bytes = gd->tile_bytes;
    tuning.phase_head_percent = gd->phase_head_percent;
    tuning.kernel_variant = (gd->kernel_variant == 2 || (gd->kernel_variant == 0 && gd->wire_wide)) ? 2 : 0;
    PhasedLaunch pl;
    if (!preparePhased(push, unpack, es, tuning, lag, &pl))
      THROW_INTERNAL_ERROR("staged schedule with equal memory orders is not a row copy");

    if (gd->fused_cache.size() >= 64) releaseFusedCache(gd);
    FusedPlanEntry e;
    e.key = key;
    const size_t box_bytes = (pl.boxes.size() * sizeof(KBox) + 15) / 16 * 16; // the tables behind it are 16-byte records
    const size_t seg_bytes = pl.segs.size() * sizeof(SegDesc);
    e.bytes = box_bytes + seg_bytes + pl.phases.size() * sizeof(PhaseDesc);
    CHECK_CUDA(cudaMalloc(&e.dev, std::max<size_t>(e.bytes, 256)));
    CHECK_CUDA(cudaHostAlloc(&e.host, std::max<size_t>(e.bytes, 256), cudaHostAllocDefault));
    std::memcpy(e.host, pl.boxes.data(), pl.boxes.size() * sizeof(KBox));
    std::memcpy(static_cast<char*>(e.host) + box_bytes, pl.segs.data(), seg_bytes);
    std::memcpy(static_cast<char*>(e.host) + box_bytes + seg_bytes, pl.phases.data(), pl.phases.size() * sizeof(PhaseDesc));
    CHECK_CUDA(cudaMemcpyAsync(e.dev, e.host, e.bytes, cudaMemcpyHostToDevice, stream));
    CHECK_CUDA(cudaEventCreateWithFlags(&e.uploaded, cudaEventDisableTiming));
    CHECK_CUDA(cudaEventRecord(e.uploaded, stream));
    e.upload_stream = stream;
    std::memset(&e.params, 0, sizeof(e.params));
    e.params.boxes = static_cast<const KBox*>(e.dev);
    e.params.segs = reinterpret_cast<const SegDesc*>(static_cast<char*>(e.dev) + box_bytes);
    e.params.phases = reinterpret_cast<const PhaseDesc*>(static_cast<char*>(e.dev) + box_bytes + seg_bytes);
    e.params.nphases = static_cast<uint32_t>(pl.phases.size());
    e.params.nsteps = pl.nsteps;
    e.params.elem_size = static_cast<uint32_t>(es);
    e.params.vec_size = static_cast<uint32_t>(pl.vec_size);
    e.total_slots = pl.total_slots;
    gd->fused_cache.push_back(e);
    entry = &gd->fused_cache.back();
  }
  entry->last_use = ++tick;
  if (stream != entry->upload_stream) CHECK_CUDA(cudaStreamWaitEvent(stream, entry->uploaded, 0));
  PhasedParams params = entry->params;
  params.sync = sync;
  LaunchConfig cfg;
  cfg.grid = gd->grid_ctas;
  cudaError_t err = launchPhased(params, entry->total_slots, cfg, stream);
  if (err != cudaSuccess) THROW_CUDA_ERROR(std::string("kernel launch failed: ") + cudaGetErrorString(err));
  PerfReport::markExchangeDone(perf, stream);
  return true;
}

} // namespace

void runTranspose(cudecompHandle_t h, cudecompGridDesc_t gd, int ax, int dir, void* input, void* output, void* work,
                  cudecompDataType_t dtype, const int32_t in_halo[], const int32_t out_halo[], const int32_t in_pad[],
                  const int32_t out_pad[], cudaStream_t stream) {
  gd->epoch += epochStride(gd); // advanced identically on every rank, whatever path the call takes
  const int es = static_cast<int>(dtypeSize(dtype));
  const bool inplace = (input == output);

  // Geometry and argument errors surface here, before any communication (every rank fails alike).
  TransposePlan probe = buildTransposePlan(gd->geom, gd->pidx, ax, dir, in_halo, out_halo, in_pad, out_pad,
                                           DstKind::FINAL, inplace);
  PerfGuard perf{nullptr, stream};
  if (gd->perf && h->have_device)
    perf.sample = gd->perf->beginTranspose(ax, dir, dtype, in_halo, out_halo, in_pad, out_pad, inplace,
                                           isManaged(input) || isManaged(output),
                                           pencilInfo(gd->geom, gd->pidx, probe.axes.a, nullptr, nullptr).size * es, stream);
  if (probe.noop) {
    gd->last_path = CUDECOMP_B200_PATH_NONE;
    return;
  }
  requireDevice(h);
  checkDeviceError(gd);
  const int P = probe.comm_size;

  if (P == 1) {
    gd->last_path = CUDECOMP_B200_PATH_LOCAL;
    SyncParams nosync;
    std::memset(&nosync, 0, sizeof(nosync));
    if (!inplace) {
      std::vector<ResolvedBox> boxes;
      for (auto& b : probe.push) boxes.push_back({b, static_cast<const char*>(input), static_cast<char*>(output)});
      launchBoxes(gd, boxes, es, nosync, stream);
    } else {
      // in place with differing layouts: through the workspace (reference transpose.h:323-362)
      TransposePlan st = buildTransposePlan(gd->geom, gd->pidx, ax, dir, in_halo, out_halo, in_pad, out_pad,
                                            DstKind::STAGE, inplace);
      std::vector<ResolvedBox> push, unpack;
      for (auto& b : st.push) push.push_back({b, static_cast<const char*>(input), static_cast<char*>(work)});
      for (auto& b : st.unpack) unpack.push_back({b, static_cast<const char*>(work), static_cast<char*>(output)});
      launchBoxes(gd, push, es, nosync, stream);
      launchBoxes(gd, unpack, es, nosync, stream);
    }
    return;
  }
  perf.exchange = true;
  NvtxRange nvtx_exchange("cudecompAlltoall"); // descriptor exchange + the peer-store launch(es): the reference's a2a phase

  // Tell the other members which buffers this call uses and learn theirs.
  CallMsg mine;
  std::memset(&mine, 0, sizeof(mine));
  mine.opcode = transposeOpcode(ax, dir);
  mine.flags = inplace ? 1u : 0u;
  describeBuffer(output, &mine.data);
  describeBuffer(work, &mine.work);
  if (gd->pull_mode) describeBuffer(input, &mine.src); // else left zeroed: not exportable
  fillReleases(h, &mine);
  std::vector<CallMsg> msgs;
  gd->mbox.exchange(probe.axes.comm == COMM_COL ? 0 : 1, probe.group_world, probe.me, mine, msgs);
  processReleases(h, probe.group_world, probe.me, msgs);
  noteDescribed(mine.data, output, probe.group_world, probe.me);
  noteDescribed(mine.work, work, probe.group_world, probe.me);
  noteDescribed(mine.src, input, probe.group_world, probe.me);

  bool direct = h->allow_direct && !gd->force_staged &&
                gd->config.transpose_comm_backend < CUDECOMP_TRANSPOSE_COMM_NVSHMEM; // NVSHMEM* values = staged schedule
  bool work_ok = true;
  bool src_ok = gd->pull_mode != 0; // receiver-driven: every member's INPUT must be mappable
  bool anybody_inplace = false, unmappable = false;
  for (auto& m : msgs) {
    if (m.flags & 1u) anybody_inplace = true; // peers may not overwrite a buffer that is still being read
    if (!m.data.exportable) unmappable = true;
    if (!m.work.exportable) work_ok = false;
    if (!m.src.exportable) src_ok = false;
  }
  if (direct && !anybody_inplace && unmappable && !gd->warned_unmappable) {
    // The call would have taken the one-kernel path: say once why it does not (the reference warns on rank 0 too).
    gd->warned_unmappable = true;
    if (h->rank == 0)
      std::printf("CUDECOMP:WARN: an output buffer of this transpose cannot be mapped by the peer GPUs (managed or pool memory, "
                  "or an allocation below 2 MiB): the exchange is staged through the workspace. Pencils from cudecompMalloc "
                  "or plain cudaMalloc (>= 2 MiB) take the direct path.\n");
  }
  if (anybody_inplace || unmappable) direct = false;
  // Receiver-driven staged schedule: every rank loads its blocks from the peers' inputs into its OWN workspace and
  // unpacks locally, so nothing is written into a peer's memory (the workspace need not be mappable). In place this is
  // safe for the same reason the sender-driven staged schedule is: a pencil is only overwritten by its owner's unpack,
  // after the exit handshake has told it that every peer has finished reading.
  const bool pull_staged = !direct && src_ok;
  if (!direct && !work_ok && !pull_staged)
    THROW_INVALID_USAGE("the workspace must be device memory that peers can map: allocate it with cudecompMalloc");

  std::vector<int> peers;
  for (int i = 0; i < P; ++i)
    if (i != probe.me) peers.push_back(probe.group_world[i]);
  const SyncParams sync = makeSync(gd, peers);

  // Receiver-driven variant of the direct path: every member's INPUT must be mappable, nobody in place (a peer's input
  // is read while that peer writes its own output). Same kernel, same handshake: the entry flag says "my input is
  // ready", the exit flag "I have read everything I needed from you"; the output is written locally.
  const bool pull = direct && src_ok;

  if (pull) {
    gd->last_path = CUDECOMP_B200_PATH_DIRECT;
    TransposePlan pl = buildPullTransposePlan(gd->geom, gd->pidx, ax, dir, in_halo, out_halo, in_pad, out_pad);
    std::vector<ResolvedBox> boxes;
    for (auto& b : pl.push) {
      const char* src = (b.peer == pl.me) ? static_cast<const char*>(input)
                                          : static_cast<const char*>(h->peers.resolve(b.peer_world, msgs[b.peer].src));
      boxes.push_back({b, src, static_cast<char*>(output)});
    }
    launchBoxes(gd, boxes, es, sync, stream, pl.me, P);
  } else if (direct) {
    gd->last_path = CUDECOMP_B200_PATH_DIRECT;
    std::vector<ResolvedBox> boxes;
    for (auto& b : probe.push) {
      char* dst = (b.peer == probe.me) ? static_cast<char*>(output)
                                       : static_cast<char*>(h->peers.resolve(b.peer_world, msgs[b.peer].data));
      boxes.push_back({b, static_cast<const char*>(input), dst});
    }
    launchBoxes(gd, boxes, es, sync, stream, probe.me, P);
  } else if (pull_staged) {
    gd->last_path = CUDECOMP_B200_PATH_STAGED;
    if (gd->pipeline_chunks > 1 &&
        runPipelinedStaged(h, gd, ax, dir, input, output, work, es, in_halo, out_halo, in_pad, out_pad, inplace, msgs,
                           peers, perf.sample, stream, true))
      return;
    TransposePlan pl = buildPullTransposePlan(gd->geom, gd->pidx, ax, dir, in_halo, out_halo, in_pad, out_pad,
                                              DstKind::STAGE, inplace);
    std::vector<ResolvedBox> pullb, unpack;
    for (auto& b : pl.push) {
      const char* src = (b.peer == pl.me) ? static_cast<const char*>(input)
                                          : static_cast<const char*>(h->peers.resolve(b.peer_world, msgs[b.peer].src));
      pullb.push_back({b, src, static_cast<char*>(work)});
    }
    for (auto& b : pl.unpack) unpack.push_back({b, static_cast<const char*>(work), static_cast<char*>(output)});
    SyncParams nosync;
    std::memset(&nosync, 0, sizeof(nosync));
    launchBoxes(gd, pullb, es, sync, stream, pl.me, P);
    PerfReport::markExchangeDone(perf.sample, stream);
    launchBoxes(gd, unpack, es, nosync, stream);
  } else {
    gd->last_path = CUDECOMP_B200_PATH_STAGED;
    if (gd->staged_mode == 0 && runFusedStaged(h, gd, ax, dir, input, output, work, es, in_halo, out_halo, in_pad, out_pad,
                                               inplace, msgs, probe, sync, perf.sample, stream))
      return;
    if (gd->pipeline_chunks > 1 &&
        runPipelinedStaged(h, gd, ax, dir, input, output, work, es, in_halo, out_halo, in_pad, out_pad, inplace, msgs,
                           peers, perf.sample, stream, false))
      return;
    TransposePlan st = buildTransposePlan(gd->geom, gd->pidx, ax, dir, in_halo, out_halo, in_pad, out_pad,
                                          DstKind::STAGE, inplace);
    std::vector<ResolvedBox> push, unpack;
    for (auto& b : st.push) {
      char* dst = (b.peer == st.me) ? static_cast<char*>(work)
                                    : static_cast<char*>(h->peers.resolve(b.peer_world, msgs[b.peer].work));
      push.push_back({b, static_cast<const char*>(input), dst});
    }
    for (auto& b : st.unpack) unpack.push_back({b, static_cast<const char*>(work), static_cast<char*>(output)});
    SyncParams nosync;
    std::memset(&nosync, 0, sizeof(nosync));
    launchBoxes(gd, push, es, sync, stream, st.me, P);
    PerfReport::markExchangeDone(perf.sample, stream);
    launchBoxes(gd, unpack, es, nosync, stream);
  }
}

void runHalo(cudecompHandle_t h, cudecompGridDesc_t gd, int ax, void* input, void* work, cudecompDataType_t dtype,
             const int32_t halo[], const bool periods[], int dim, const int32_t pad[], cudaStream_t stream) {
  gd->epoch += epochStride(gd);
  const int es = static_cast<int>(dtypeSize(dtype));

  // Make the "halo wider than a neighbour's slab" error collective: the reference raises it only on the
  // ranks that touch the thin slab (halo.h:120-145), which would leave the others waiting.
  if (dim != ax && halo[dim] > 0 && !hasEmptyPencils(gd->geom, ax)) {
    const int P = gd->geom.pdims[haloCommAxis(ax, dim)];
    if (P > 1) {
      const auto splits = getSplits(gd->geom.gdims_dist[dim], P, gd->geom.gdims[dim] - gd->geom.gdims_dist[dim]);
      if (halo[dim] > *std::min_element(splits.begin(), splits.end()))
        THROW_INVALID_USAGE(
            "halo includes ranks other than nearest neighbor processes, this is not currently supported.");
    }
  }

  HaloPlan probe = buildHaloPlan(gd->geom, gd->pidx, ax, dim, halo, periods, pad, DstKind::FINAL);
  if (probe.nothing) {
    gd->last_path = CUDECOMP_B200_PATH_NONE;
    return;
  }
  requireDevice(h);
  checkDeviceError(gd);
  PerfGuard perf{nullptr, stream};
  if (gd->perf)
    perf.sample = gd->perf->beginHalo(ax, dim, dtype, halo, periods, pad, isManaged(input),
                                      probe.face_elems * es * ((probe.neighbor[0] >= 0) + (probe.neighbor[1] >= 0)),
                                      stream);

  SyncParams nosync;
  std::memset(&nosync, 0, sizeof(nosync));

  if (probe.comm_size == 1) {
    // periodic wrap inside one rank: two local face copies (reference halo.h:165-193)
    gd->last_path = CUDECOMP_B200_PATH_LOCAL;
    std::vector<ResolvedBox> boxes;
    for (auto& b : probe.push) boxes.push_back({b, static_cast<const char*>(input), static_cast<char*>(input)});
    launchBoxes(gd, boxes, es, nosync, stream);
    return;
  }

  perf.exchange = true;
  NvtxRange nvtx_exchange("cudecompSendRecvPair");
  CallMsg mine;
  std::memset(&mine, 0, sizeof(mine));
  mine.opcode = haloOpcode(ax, dim);
  describeBuffer(input, &mine.data);
  describeBuffer(work, &mine.work);
  fillReleases(h, &mine);
  std::vector<CallMsg> msgs;
  gd->mbox.exchange(probe.comm == COMM_COL ? 0 : 1, probe.group_world, probe.me, mine, msgs);
  processReleases(h, probe.group_world, probe.me, msgs);
  noteDescribed(mine.data, input, probe.group_world, probe.me);
  noteDescribed(mine.work, work, probe.group_world, probe.me);

  bool direct = h->allow_direct && !gd->force_staged && gd->config.halo_comm_backend < CUDECOMP_HALO_COMM_NVSHMEM;
  bool work_ok = true;
  for (auto& m : msgs) {
    if (!m.data.exportable) direct = false;
    if (!m.work.exportable) work_ok = false;
  }
  if (!direct && !work_ok)
    THROW_INVALID_USAGE("the workspace must be device memory that peers can map: allocate it with cudecompMalloc");

  std::vector<int> peers;
  for (int nb : probe.neighbor) {
    if (nb < 0 || nb == probe.me) continue;
    const int w = probe.group_world[nb];
    if (std::find(peers.begin(), peers.end(), w) == peers.end()) peers.push_back(w);
  }
  const SyncParams sync = makeSync(gd, peers);

  if (direct) {
    gd->last_path = CUDECOMP_B200_PATH_DIRECT;
    std::vector<ResolvedBox> boxes;
    for (auto& b : probe.push) {
      char* dst = (b.peer == probe.me) ? static_cast<char*>(input)
                                       : static_cast<char*>(h->peers.resolve(b.peer_world, msgs[b.peer].data));
      boxes.push_back({b, static_cast<const char*>(input), dst});
    }
    launchBoxes(gd, boxes, es, sync, stream, probe.me, probe.comm_size);
  } else {
    gd->last_path = CUDECOMP_B200_PATH_STAGED;
    HaloPlan st = buildHaloPlan(gd->geom, gd->pidx, ax, dim, halo, periods, pad, DstKind::STAGE);
    std::vector<ResolvedBox> push, unpack;
    for (auto& b : st.push) {
      char* dst = (b.peer == st.me) ? static_cast<char*>(work)
                                    : static_cast<char*>(h->peers.resolve(b.peer_world, msgs[b.peer].work));
      push.push_back({b, static_cast<const char*>(input), dst});
    }
    for (auto& b : st.unpack) unpack.push_back({b, static_cast<const char*>(work), static_cast<char*>(input)});
    launchBoxes(gd, push, es, sync, stream, st.me, st.comm_size);
    PerfReport::markExchangeDone(perf.sample, stream);
    launchBoxes(gd, unpack, es, nosync, stream);
  }
}

} // namespace cdb
