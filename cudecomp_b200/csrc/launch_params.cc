// See launch_params.h.
#include "launch_params.h"

#include <algorithm>
#include <cstring>

#include "errors.h"
#include "tiling.h"

namespace cdb {

namespace {

uint64_t lowBit(uint64_t x) { return x ? (x & (~x + 1)) : (1ull << 62); }

void fillRowCopy(KBox& kb, const CanonBox& c, int es, int V, uint32_t tile_bytes, bool bulk) {
  // SIMT: tiles of tile_bytes (several short rows or one segment of a long row); bulk: one row segment of <= 16 KiB
  const uint32_t tile_vecs = (bulk ? kBulkChunkBytes : tile_bytes) / static_cast<uint32_t>(V);
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

// TRANSPOSE boxes: axis 0 = contiguous in the source, axis 1 = contiguous in the destination when there is one
struct TransAxes {
  int64_t n[3], ss[3], ds[3];
};
TransAxes transposeAxesOf(const CanonBox& c) {
  int a1 = c.dstUnitAxis();
  if (a1 <= 0) a1 = (c.nd > 1) ? 1 : -1;
  int a2 = -1;
  for (int k = 1; k < c.nd; ++k)
    if (k != a1) a2 = k;
  const int map[3] = {0, a1, a2};
  TransAxes t;
  for (int k = 0; k < 3; ++k) {
    t.n[k] = (map[k] >= 0) ? c.n[map[k]] : 1;
    t.ss[k] = (map[k] >= 0) ? c.ss[map[k]] : 0;
    t.ds[k] = (map[k] >= 0) ? c.ds[map[k]] : 0;
  }
  return t;
}

// 16-byte accesses on both sides: unit strides where the kernel assumes them, every row start 16-byte aligned, whole
// vectors only
bool transposeVecOk(const TransAxes& t, const LaunchBox& b, int es) {
  const int vec = 16 / es;
  if (t.ss[0] != 1 || t.ds[1] != 1 || t.n[1] <= 1) return false;
  if (t.n[0] % vec || t.n[1] % vec) return false;
  const uint64_t sa = reinterpret_cast<uint64_t>(b.src_base) + static_cast<uint64_t>(b.d.src_off) * es;
  const uint64_t da = reinterpret_cast<uint64_t>(b.dst_base) + static_cast<uint64_t>(b.d.dst_off) * es;
  if (sa % 16 || da % 16) return false;
  if ((t.ss[1] * es) % 16 || (t.ds[0] * es) % 16) return false;
  if (t.n[2] > 1 && ((t.ss[2] * es) % 16 || (t.ds[2] * es) % 16)) return false;
  return true;
}

} // namespace

std::vector<PreparedLaunch> prepareLaunches(const std::vector<LaunchBox>& boxes, int es, const LaunchTuning& tuning, int me,
                                            int comm_size) {
  std::vector<CanonBox> canon;
  std::vector<const LaunchBox*> live;
  for (auto& b : boxes) {
    if (b.d.count() == 0) continue;
    live.push_back(&b);
  }
  if (tuning.peer_order == 1 && comm_size > 1 && me >= 0)
    std::stable_sort(live.begin(), live.end(), [&](const LaunchBox* x, const LaunchBox* y) {
      return ((x->d.peer - me) % comm_size + comm_size) % comm_size < ((y->d.peer - me) % comm_size + comm_size) % comm_size;
    });
  for (auto* b : live) canon.push_back(canonicalize(b->d, true));

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
    uint64_t a = (tuning.kernel_variant == 2) ? 32 : 16; // variant 2: 256-bit accesses where everything is 32-byte aligned
    for (size_t i = 0; i < canon.size(); ++i) {
      const CanonBox& c = canon[i];
      const uint64_t sa = reinterpret_cast<uint64_t>(live[i]->src_base) + static_cast<uint64_t>(live[i]->d.src_off) * es;
      const uint64_t da = reinterpret_cast<uint64_t>(live[i]->dst_base) + static_cast<uint64_t>(live[i]->d.dst_off) * es;
      a = std::min({a, lowBit(sa), lowBit(da), lowBit(static_cast<uint64_t>(c.n[0]) * es)});
      for (int k = 1; k < 3; ++k) {
        if (c.n[k] > 1) a = std::min({a, lowBit(static_cast<uint64_t>(c.ss[k]) * es), lowBit(static_cast<uint64_t>(c.ds[k]) * es)});
      }
    }
    V = static_cast<int>(std::min<uint64_t>(a, 32));
    if (V < 4) THROW_INVALID_USAGE("buffers must be aligned to the element size");
  }

  // TMA bulk variant: only when every row is 16-byte aligned and long enough for one bulk copy to pay off
  if (kind == KernelKind::ROWCOPY && tuning.kernel_variant == 1 && V == 16) {
    bool ok = !canon.empty();
    for (auto& c : canon)
      if (c.n[0] * es < 2048) ok = false;
    if (ok) kind = KernelKind::ROWCOPY_BULK;
  }

  // vectorised transpose where every box allows it (kernel_variant 3 keeps the element-wise kernel, for comparison)
  if (kind == KernelKind::TRANSPOSE && tuning.kernel_variant != 3) {
    bool ok = !canon.empty();
    for (size_t i = 0; i < canon.size(); ++i)
      if (!transposeVecOk(transposeAxesOf(canon[i]), *live[i], es)) ok = false;
    if (ok) kind = KernelKind::TRANSPOSE_VEC;
  }

  uint32_t tile_bytes = static_cast<uint32_t>(tuning.tile_bytes > 0 ? tuning.tile_bytes : kDefaultTileBytes);
  tile_bytes = std::min<uint32_t>(std::max<uint32_t>(tile_bytes, kMinTileBytes), kMaxTileBytes);

  std::vector<PreparedLaunch> out;
  const size_t nlaunch = std::max<size_t>(1, (canon.size() + kMaxBoxes - 1) / kMaxBoxes);
  for (size_t l = 0; l < nlaunch; ++l) {
    PreparedLaunch pl;
    pl.kind = kind;
    CopyParams& p = pl.params;
    std::memset(&p, 0, sizeof(p));
    p.elem_size = static_cast<uint32_t>(es);
    p.vec_size = static_cast<uint32_t>(V);
    p.peer_order = (tuning.peer_order == 1) ? 1u : 0u;
    p.geometry = (kind == KernelKind::TRANSPOSE_VEC && es == 8 && tuning.transpose_geometry) ? 1u : 0u;
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
        fillRowCopy(kb, c, es, V, tile_bytes, kind == KernelKind::ROWCOPY_BULK);
      } else {
        const TransAxes t = transposeAxesOf(c);
        for (int k = 0; k < 3; ++k) {
          kb.n[k] = t.n[k];
          kb.ss[k] = t.ss[k];
          kb.ds[k] = t.ds[k];
        }
        int e0 = 32, e1 = 32;
        if (kind == KernelKind::TRANSPOSE_VEC) transVecTileExtents(es, tuning.transpose_geometry, e0, e1);
        kb.tiles0 = static_cast<uint32_t>((kb.n[0] + e0 - 1) / e0);
        kb.tiles1 = static_cast<uint32_t>((kb.n[1] + e1 - 1) / e1);
        const int64_t tiles = static_cast<int64_t>(kb.tiles0) * kb.tiles1 * kb.n[2];
        if (tiles > 0x7fffffff) THROW_NOT_SUPPORTED("box too large for one launch");
        kb.tiles = static_cast<uint32_t>(tiles);
      }
      p.max_tiles = std::max(p.max_tiles, kb.tiles);
    }
    if (static_cast<uint64_t>(p.nboxes) * p.max_tiles > 0xffffffffull) THROW_NOT_SUPPORTED("launch too large");
    out.push_back(pl);
  }
  return out;
}

bool preparePhased(const std::vector<std::vector<LaunchBox>>& push, const std::vector<std::vector<LaunchBox>>& unpack, int es,
                   const LaunchTuning& tuning, int lag, PhasedLaunch* out) {
  const size_t K = push.size();
  if (K == 0 || unpack.size() != K || lag < 1 || K > static_cast<size_t>(kMaxPhases)) return false;
  uint32_t tile_bytes = static_cast<uint32_t>(tuning.tile_bytes > 0 ? tuning.tile_bytes : kDefaultTileBytes);
  tile_bytes = std::min<uint32_t>(std::max<uint32_t>(tile_bytes, kMinTileBytes), kMaxTileBytes);

  struct Item {
    CanonBox c;
    const LaunchBox* b;
    int wait; // step that must be complete, -1: none
  };
  std::vector<std::vector<Item>> pushes(K), unpacks(K);
  uint64_t a = (tuning.kernel_variant == 2) ? 32 : 16;
  auto add = [&](const LaunchBox& b, std::vector<Item>& list, int wait) -> bool {
    if (b.d.count() == 0) return true;
    CanonBox c = canonicalize(b.d, true);
    if (c.rowCopy() && c.n[0] * es / std::min(es, 16) >= (1ll << 31)) c = canonicalize(b.d, false);
    if (!c.rowCopy()) return false;
    const uint64_t sa = reinterpret_cast<uint64_t>(b.src_base) + static_cast<uint64_t>(b.d.src_off) * es;
    const uint64_t da = reinterpret_cast<uint64_t>(b.dst_base) + static_cast<uint64_t>(b.d.dst_off) * es;
    a = std::min({a, lowBit(sa), lowBit(da), lowBit(static_cast<uint64_t>(c.n[0]) * es)});
    for (int k = 1; k < 3; ++k)
      if (c.n[k] > 1) a = std::min({a, lowBit(static_cast<uint64_t>(c.ss[k]) * es), lowBit(static_cast<uint64_t>(c.ds[k]) * es)});
    list.push_back({c, &b, wait});
    return true;
  };
  for (size_t s = 0; s < K; ++s) {
    for (auto& b : push[s])
      if (!add(b, pushes[s], -1)) return false;
    for (auto& b : unpack[s])
      if (!add(b, unpacks[s], static_cast<int>(s))) return false;
  }
  if (a < 4) THROW_INVALID_USAGE("buffers must be aligned to the element size");
  const int V = static_cast<int>(a);

  out->boxes.clear();
  out->segs.clear();
  out->phases.clear();
  out->nsteps = static_cast<uint32_t>(K);
  out->vec_size = V;
  out->total_slots = 0;

  // the segments of one box list, box by box
  auto segmentsOf = [&](const std::vector<Item>& items) {
    std::vector<std::vector<SegDesc>> per_box;
    for (auto& it : items) {
      KBox kb;
      std::memset(&kb, 0, sizeof(kb));
      kb.src = it.b->src_base + it.b->d.src_off * es;
      kb.dst = it.b->dst_base + it.b->d.dst_off * es;
      for (int k = 0; k < 3; ++k) {
        kb.n[k] = it.c.n[k];
        kb.ss[k] = it.c.ss[k];
        kb.ds[k] = it.c.ds[k];
      }
      fillRowCopy(kb, it.c, es, V, tile_bytes, false);
      kb.pad_ = static_cast<uint32_t>(it.wait + 1);
      const uint32_t box_index = static_cast<uint32_t>(out->boxes.size());
      out->boxes.push_back(kb);
      std::vector<SegDesc> segs;
      for (uint32_t j0 = 0; j0 < kb.tiles; j0 += kSegTiles) segs.push_back({box_index, j0, std::min(kSegTiles, kb.tiles - j0), kb.pad_});
      per_box.push_back(std::move(segs));
    }
    return per_box;
  };
  // one phase from segment lists: round robin over the boxes (segment q of every box before segment q + 1 of any)
  auto emitPhase = [&](const std::vector<std::vector<SegDesc>>& a_, size_t a_from, size_t a_to_frac_num, size_t a_to_frac_den,
                       bool take_tail, const std::vector<std::vector<SegDesc>>* b_, uint32_t publish) {
    PhaseDesc pd{};
    pd.first_seg = static_cast<uint32_t>(out->segs.size());
    pd.publish = publish;
    size_t longest = 0;
    for (auto& segs : a_) longest = std::max(longest, segs.size());
    if (b_)
      for (auto& segs : *b_) longest = std::max(longest, segs.size());
    for (size_t q = 0; q < longest; ++q) {
      for (auto& segs : a_) {
        // head: segments [0, n * num / den); tail: the rest
        const size_t cut = segs.size() * a_to_frac_num / a_to_frac_den;
        const size_t lo = take_tail ? cut : a_from, hi = take_tail ? segs.size() : cut;
        if (lo + q < hi) {
          out->segs.push_back(segs[lo + q]);
          pd.seg_tiles = std::max(pd.seg_tiles, segs[lo + q].count);
        }
      }
      if (b_)
        for (auto& segs : *b_)
          if (q < segs.size()) {
            out->segs.push_back(segs[q]);
            pd.seg_tiles = std::max(pd.seg_tiles, segs[q].count);
          }
    }
    pd.nsegs = static_cast<uint32_t>(out->segs.size()) - pd.first_seg;
    if (static_cast<uint64_t>(pd.nsegs) * pd.seg_tiles > 0xffffffffull) THROW_NOT_SUPPORTED("launch too large");
    out->total_slots += static_cast<uint64_t>(pd.nsegs) * pd.seg_tiles;
    out->phases.push_back(pd);
  };

  // Step s = the pushes of chunk s and the unpacks of chunk s - lag. Unpack tiles wait for chunk s - lag to be complete
  // on every member; CTAs drift apart by about a tile within a phase, so with a short lag the step opens with a head of
  // pushes only (head_percent of every push box) and the unpacks join for the rest: by then the slowest CTA anywhere
  // has long finished the earlier chunk and nobody waits.
  const int head = std::min(90, std::max(0, tuning.phase_head_percent));
  std::vector<std::vector<std::vector<SegDesc>>> unpack_segs(K);
  const std::vector<std::vector<SegDesc>> none;
  for (size_t s = 0; s < K + static_cast<size_t>(lag); ++s) {
    const std::vector<std::vector<SegDesc>> ps = (s < K) ? segmentsOf(pushes[s]) : none;
    if (s < K) unpack_segs[s] = segmentsOf(unpacks[s]);
    const std::vector<std::vector<SegDesc>>* us = (s >= static_cast<size_t>(lag)) ? &unpack_segs[s - lag] : nullptr;
    const bool have_unpack = us && !us->empty();
    const uint32_t publish = (s < K) ? static_cast<uint32_t>(s + 1) : 0u;
    if (have_unpack && s < K && head > 0) {
      emitPhase(ps, 0, static_cast<size_t>(head), 100, false, nullptr, 0u);
      emitPhase(ps, 0, static_cast<size_t>(head), 100, true, us, publish);
    } else {
      emitPhase(ps, 0, 1, 1, false, us, publish);
    }
  }
  return true;
}

int chooseGrid(int requested, int dflt, int resident, uint64_t total_slots, int balance) {
  int grid = requested > 0 ? requested : dflt;
  if (grid > resident) grid = resident;
  if (static_cast<uint64_t>(grid) > total_slots) grid = static_cast<int>(total_slots);
  if (grid < 1) return 1; // still runs the handshake when this rank has nothing to move
  if (balance && total_slots > static_cast<uint64_t>(grid)) {
    // rounds(g) = ceil(slots / g); utilisation of a launch = slots / (g * rounds(g)). Among the counts within 20 % of
    // the chosen one take the best utilisation: the slowest CTA sets the launch time, and a last round that occupies a
    // handful of CTAs costs a whole tile time (a 512^3 complex64 pencil is only 11 rounds of 32 KiB tiles).
    int best = grid;
    double best_u = 0.0;
    for (int g = grid; g >= std::max(1, grid - grid / 5); --g) {
      const uint64_t rounds = (total_slots + g - 1) / g;
      const double u = static_cast<double>(total_slots) / (static_cast<double>(g) * static_cast<double>(rounds));
      if (u > best_u + 1e-12) {
        best_u = u;
        best = g;
      }
    }
    grid = best;
  }
  return grid;
}

} // namespace cdb
