// Execution of one transpose / halo call. See engine.h, plan.h, kernels.h.
#include "engine.h"

#include <algorithm>
#include <cstring>

#include "errors.h"

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

void setGeometry(cudecompGridDesc_t gd, const std::array<int32_t, 2>& pdims) {
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
  arena.clearError();
  THROW_INTERNAL_ERROR(std::string("a device-side wait for a peer rank timed out during an earlier operation (") +
                       (e == 1 ? "entry" : "exit") +
                       " handshake); a rank of the communicator did not enter the same operation");
}

namespace {

struct ResolvedBox {
  BoxDesc d;
  const char* src_base;
  char* dst_base;
};

uint64_t lowBit(uint64_t x) { return x ? (x & (~x + 1)) : (1ull << 62); }

// Fills the per-launch handshake block. `peers` are global ranks other than mine.
SyncParams makeSync(cudecompGridDesc_t gd, const std::vector<int>& peers) {
  SyncParams s;
  std::memset(&s, 0, sizeof(s));
  if (peers.empty()) return s;
  SignalArena& arena = gd->handle->arena;
  if (!arena.valid() || gd->pad_slot < 0) THROW_INTERNAL_ERROR("signal pads are not initialised");
  if (peers.size() > static_cast<size_t>(kMaxPeers))
    THROW_NOT_SUPPORTED("communicators with more than 17 ranks are not supported yet");
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

void fillRowCopy(KBox& kb, const CanonBox& c, int es, int V, bool bulk = false) {
  // SIMT: tiles of ~32 KiB (several short rows or one segment of a long row); bulk: one row segment of <= 16 KiB
  const uint32_t tile_vecs = (bulk ? kBulkChunkBytes : 32768u) / static_cast<uint32_t>(V);
  kb.row_vecs = static_cast<uint32_t>(c.n[0] * es / V);
  kb.seg_vecs = std::min(kb.row_vecs, tile_vecs);
  if (kb.seg_vecs == 0) kb.seg_vecs = 1;
  kb.segs_per_row = (kb.row_vecs + kb.seg_vecs - 1) / kb.seg_vecs;
  kb.rows_per_tile = bulk ? 1u : std::max(1u, tile_vecs / kb.seg_vecs);
  const int64_t rows = c.n[1] * c.n[2];
  const int64_t row_tiles = (rows + kb.rows_per_tile - 1) / kb.rows_per_tile;
  const int64_t tiles = (c.n[0] == 0) ? 0 : row_tiles * kb.segs_per_row;
  if (tiles > 0x7fffffff) THROW_NOT_SUPPORTED("box too large for one launch");
  kb.tiles = static_cast<uint32_t>(tiles);
  kb.tiles0 = kb.tiles1 = 0;
}

// Enqueue the copy of `boxes` (all with the same element size). `sync` handshakes with peers (may be empty).
void launchBoxes(cudecompGridDesc_t gd, const std::vector<ResolvedBox>& boxes, int es, const SyncParams& sync,
                 cudaStream_t stream) {
  std::vector<CanonBox> canon;
  std::vector<const ResolvedBox*> live;
  for (auto& b : boxes) {
    if (b.d.count() == 0) continue;
    canon.push_back(canonicalize(b.d, true));
    live.push_back(&b);
  }
  bool all_rows = true;
  for (auto& c : canon)
    if (!c.rowCopy()) all_rows = false;
  if (all_rows) {
    // keep row lengths addressable with 32-bit vector indices
    for (size_t i = 0; i < canon.size(); ++i)
      if (canon[i].n[0] * es / std::min(es, 16) >= (1ll << 31)) canon[i] = canonicalize(live[i]->d, false);
    for (auto& c : canon)
      if (!c.rowCopy()) all_rows = false;
  }
  KernelKind kind = all_rows ? KernelKind::ROWCOPY : KernelKind::TRANSPOSE;

  int V = 16;
  if (kind == KernelKind::ROWCOPY) {
    uint64_t a = 16;
    for (size_t i = 0; i < canon.size(); ++i) {
      const CanonBox& c = canon[i];
      const uint64_t sa = reinterpret_cast<uint64_t>(live[i]->src_base) + static_cast<uint64_t>(live[i]->d.src_off) * es;
      const uint64_t da = reinterpret_cast<uint64_t>(live[i]->dst_base) + static_cast<uint64_t>(live[i]->d.dst_off) * es;
      a = std::min({a, lowBit(sa), lowBit(da), lowBit(static_cast<uint64_t>(c.n[0]) * es)});
      for (int k = 1; k < 3; ++k) {
        if (c.n[k] > 1) a = std::min({a, lowBit(static_cast<uint64_t>(c.ss[k]) * es), lowBit(static_cast<uint64_t>(c.ds[k]) * es)});
      }
    }
    V = static_cast<int>(std::min<uint64_t>(a, 16));
    if (V < 4) THROW_INVALID_USAGE("buffers must be aligned to the element size");
  }

  // TMA bulk variant: only when every row is 16-byte aligned and long enough for one bulk copy to pay off
  if (kind == KernelKind::ROWCOPY && gd->kernel_variant == 1 && V == 16) {
    bool ok = !canon.empty();
    for (auto& c : canon)
      if (c.n[0] * es < 2048) ok = false;
    if (ok) kind = KernelKind::ROWCOPY_BULK;
  }

  LaunchConfig cfg;
  cfg.grid = gd->grid_ctas;

  const size_t nlaunch = std::max<size_t>(1, (canon.size() + kMaxBoxes - 1) / kMaxBoxes);
  for (size_t l = 0; l < nlaunch; ++l) {
    CopyParams p;
    std::memset(&p, 0, sizeof(p));
    p.elem_size = static_cast<uint32_t>(es);
    p.vec_size = static_cast<uint32_t>(V);
    p.sync = sync;
    p.sync.do_entry = (sync.npeers > 0 && l == 0) ? 1 : 0;
    p.sync.do_exit = (sync.npeers > 0 && l + 1 == nlaunch) ? 1 : 0;
    const size_t lo = l * kMaxBoxes, hi = std::min(canon.size(), lo + kMaxBoxes);
    for (size_t i = lo; i < hi; ++i) {
      const CanonBox& c = canon[i];
      KBox& kb = p.box[p.nboxes++];
      kb.src = live[i]->src_base + live[i]->d.src_off * es;
      kb.dst = live[i]->dst_base + live[i]->d.dst_off * es;
      if (kind == KernelKind::ROWCOPY || kind == KernelKind::ROWCOPY_BULK) {
        for (int k = 0; k < 3; ++k) {
          kb.n[k] = c.n[k];
          kb.ss[k] = c.ss[k];
          kb.ds[k] = c.ds[k];
        }
        fillRowCopy(kb, c, es, V, kind == KernelKind::ROWCOPY_BULK);
      } else {
        // axis 0: contiguous in the source; axis 1: contiguous in the destination when there is one
        int a1 = c.dstUnitAxis();
        if (a1 <= 0) a1 = (c.nd > 1) ? 1 : -1;
        int a2 = -1;
        for (int k = 1; k < c.nd; ++k)
          if (k != a1) a2 = k;
        const int map[3] = {0, a1, a2};
        for (int k = 0; k < 3; ++k) {
          kb.n[k] = (map[k] >= 0) ? c.n[map[k]] : 1;
          kb.ss[k] = (map[k] >= 0) ? c.ss[map[k]] : 0;
          kb.ds[k] = (map[k] >= 0) ? c.ds[map[k]] : 0;
        }
        kb.tiles0 = static_cast<uint32_t>((kb.n[0] + 31) / 32);
        kb.tiles1 = static_cast<uint32_t>((kb.n[1] + 31) / 32);
        const int64_t tiles = static_cast<int64_t>(kb.tiles0) * kb.tiles1 * kb.n[2];
        if (tiles > 0x7fffffff) THROW_NOT_SUPPORTED("box too large for one launch");
        kb.tiles = static_cast<uint32_t>(tiles);
      }
      p.max_tiles = std::max(p.max_tiles, kb.tiles);
    }
    if (static_cast<uint64_t>(p.nboxes) * p.max_tiles > 0xffffffffull) THROW_NOT_SUPPORTED("launch too large");
    if (p.nboxes == 0 && sync.npeers == 0) continue;
    cudaError_t err = launchCopy(kind, p, cfg, stream);
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
    if (static_cast<int>(i) != me) h->peers.noteReleases(group_world[i], msgs[i].release_count, msgs[i].released);
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
// chunking does not apply (the caller then runs the unchunked schedule). Every rank of the job takes the same
// decision and advances the epoch by the same amount: it only depends on the geometry and on K.
bool runPipelinedStaged(cudecompHandle_t h, cudecompGridDesc_t gd, int ax, int dir, void* input, void* output, void* work,
                        int es, const int32_t in_halo[], const int32_t out_halo[], const int32_t in_pad[],
                        const int32_t out_pad[], bool inplace, const std::vector<CallMsg>& msgs,
                        const std::vector<int>& peers, PerfSample* perf, cudaStream_t stream) {
  PipelinedPlan pp = buildPipelinedTransposePlan(gd->geom, gd->pidx, ax, dir, in_halo, out_halo, in_pad, out_pad, inplace,
                                                 gd->pipeline_chunks);
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
  for (size_t s = 0; s < K; ++s) {
    if (s > 0) gd->epoch++; // the first step uses the epoch the call was given
    const SyncParams sync = makeSync(gd, peers);
    std::vector<ResolvedBox> push, unpack;
    for (auto& b : pp.steps[s].push) {
      char* dst = (b.peer == pp.base.me) ? static_cast<char*>(work)
                                         : static_cast<char*>(h->peers.resolve(b.peer_world, msgs[b.peer].work));
      push.push_back({b, static_cast<const char*>(input), dst});
    }
    launchBoxes(gd, push, es, sync, stream);
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

} // namespace

void runTranspose(cudecompHandle_t h, cudecompGridDesc_t gd, int ax, int dir, void* input, void* output, void* work,
                  cudecompDataType_t dtype, const int32_t in_halo[], const int32_t out_halo[], const int32_t in_pad[],
                  const int32_t out_pad[], cudaStream_t stream) {
  gd->epoch++; // one epoch per collective call, advanced identically on every rank
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

  // Tell the other members which buffers this call uses and learn theirs.
  CallMsg mine;
  std::memset(&mine, 0, sizeof(mine));
  mine.opcode = transposeOpcode(ax, dir);
  mine.flags = inplace ? 1u : 0u;
  describeBuffer(output, &mine.data);
  describeBuffer(work, &mine.work);
  fillReleases(h, &mine);
  std::vector<CallMsg> msgs;
  gd->mbox.exchange(probe.axes.comm == COMM_COL ? 0 : 1, probe.group_world, probe.me, mine, msgs);
  processReleases(h, probe.group_world, probe.me, msgs);

  bool direct = h->allow_direct && !gd->force_staged &&
                gd->config.transpose_comm_backend < CUDECOMP_TRANSPOSE_COMM_NVSHMEM; // NVSHMEM* values = staged schedule
  bool work_ok = true;
  for (auto& m : msgs) {
    if (m.flags & 1u) direct = false; // anybody in place: peers may not overwrite a buffer that is still being read
    if (!m.data.exportable) direct = false;
    if (!m.work.exportable) work_ok = false;
  }
  if (!direct && !work_ok)
    THROW_INVALID_USAGE("the workspace must be device memory that peers can map: allocate it with cudecompMalloc");

  std::vector<int> peers;
  for (int i = 0; i < P; ++i)
    if (i != probe.me) peers.push_back(probe.group_world[i]);
  const SyncParams sync = makeSync(gd, peers);

  if (direct) {
    gd->last_path = CUDECOMP_B200_PATH_DIRECT;
    std::vector<ResolvedBox> boxes;
    for (auto& b : probe.push) {
      char* dst = (b.peer == probe.me) ? static_cast<char*>(output)
                                       : static_cast<char*>(h->peers.resolve(b.peer_world, msgs[b.peer].data));
      boxes.push_back({b, static_cast<const char*>(input), dst});
    }
    launchBoxes(gd, boxes, es, sync, stream);
  } else {
    gd->last_path = CUDECOMP_B200_PATH_STAGED;
    if (gd->pipeline_chunks > 1 &&
        runPipelinedStaged(h, gd, ax, dir, input, output, work, es, in_halo, out_halo, in_pad, out_pad, inplace, msgs,
                           peers, perf.sample, stream))
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
    launchBoxes(gd, push, es, sync, stream);
    PerfReport::markExchangeDone(perf.sample, stream);
    launchBoxes(gd, unpack, es, nosync, stream);
  }
}

void runHalo(cudecompHandle_t h, cudecompGridDesc_t gd, int ax, void* input, void* work, cudecompDataType_t dtype,
             const int32_t halo[], const bool periods[], int dim, const int32_t pad[], cudaStream_t stream) {
  gd->epoch++;
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
  CallMsg mine;
  std::memset(&mine, 0, sizeof(mine));
  mine.opcode = haloOpcode(ax, dim);
  describeBuffer(input, &mine.data);
  describeBuffer(work, &mine.work);
  fillReleases(h, &mine);
  std::vector<CallMsg> msgs;
  gd->mbox.exchange(probe.comm == COMM_COL ? 0 : 1, probe.group_world, probe.me, mine, msgs);
  processReleases(h, probe.group_world, probe.me, msgs);

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
    launchBoxes(gd, boxes, es, sync, stream);
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
    launchBoxes(gd, push, es, sync, stream);
    PerfReport::markExchangeDone(perf.sample, stream);
    launchBoxes(gd, unpack, es, nosync, stream);
  }
}

} // namespace cdb
