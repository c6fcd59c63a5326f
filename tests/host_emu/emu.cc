// Host-side emulator of the copy kernels' index arithmetic -- TEST INFRASTRUCTURE, never part of libcudecomp.so.
//
// It prepares launches with the product's own host code (csrc/launch_params.cc: kernel selection, vector width,
// tiling) and then walks every launch the way the device does: CTA by CTA over the grid-stride slot list, decoding
// slots and tiles with the SAME inline functions the kernels use (csrc/tiling.h), and inside a tile warp by warp and
// lane by lane with loops that restate those of csrc/kernels.cu. Every vector access is checked for alignment and
// against the legal address ranges, and bytes are copied on numpy buffers, so a Python test can compare the outcome
// with the oracle without a GPU. What it cannot show: memory ordering, the cross-GPU handshake, TMA/mbarrier protocol.
#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "cudecomp_b200_ext.h"
#include "launch_params.h"
#include "tiling.h"

using namespace cdb;

namespace {

struct Fail : std::exception {
  std::string msg;
  explicit Fail(std::string m) : msg(std::move(m)) {}
  const char* what() const noexcept override { return msg.c_str(); }
};

std::string fmt(const char* f, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof(buf), f, ap);
  va_end(ap);
  return buf;
}

struct Ranges {
  const char* const* lo;
  const int64_t* len;
  int n;
  void check(const void* p, int64_t bytes, const char* what) const {
    const char* c = static_cast<const char*>(p);
    for (int i = 0; i < n; ++i)
      if (c >= lo[i] && c + bytes <= lo[i] + len[i]) return;
    throw Fail(fmt("%s of %lld bytes outside every buffer", what, static_cast<long long>(bytes)));
  }
};

struct Walk {
  const Ranges& ranges;
  int64_t bytes_written = 0;
  int64_t accesses = 0;
  void vec(char* dst, const char* src, int V) {
    if (reinterpret_cast<uintptr_t>(dst) % V || reinterpret_cast<uintptr_t>(src) % V)
      throw Fail(fmt("misaligned %d-byte access", V));
    ranges.check(src, V, "load");
    ranges.check(dst, V, "store");
    std::memcpy(dst, src, V);
    bytes_written += V;
    ++accesses;
  }
};

// rowCopyKernel<V, order>: see csrc/kernels.cu
void walkRowCopy(const CopyParams& p, int grid, int threads, Walk& w) {
  const int V = static_cast<int>(p.vec_size);
  const uint32_t nwarps = static_cast<uint32_t>(threads) >> 5;
  const uint32_t total = p.nboxes * p.max_tiles;
  constexpr uint32_t kUnroll = 4, kPiece = 32 * kUnroll;
  for (uint32_t cta = 0; cta < static_cast<uint32_t>(grid); ++cta) {
    for (uint32_t t = cta; t < total; t += grid) {
      uint32_t b, j;
      slotToBoxTile(t, p.nboxes, p.max_tiles, p.peer_order, b, j);
      if (b >= p.nboxes) throw Fail("slot decodes to a box that does not exist");
      const KBox& bx = p.box[b];
      if (j >= bx.tiles) continue;
      const RowTile rt = decodeRowTile(bx, j);
      const uint32_t pieces_per_row = (rt.nvec + kPiece - 1) / kPiece;
      const uint32_t npieces = rt.rows_here * pieces_per_row;
      const int64_t esz = p.elem_size;
      if (bx.row_vecs <= 32u) {
        const uint32_t total_vecs = rt.rows_here * rt.nvec;
        for (uint32_t tid = 0; tid < static_cast<uint32_t>(threads); ++tid)
          for (uint32_t e0 = tid; e0 < total_vecs; e0 += threads * kUnroll)
            for (uint32_t k = 0; k < kUnroll; ++k) {
              const uint32_t e = e0 + k * threads;
              if (e >= total_vecs) continue;
              const uint32_t r = e / rt.nvec, c = e - r * rt.nvec;
              int64_t so, dof;
              rowOffsets(bx, rt.row0 + r, esz, so, dof);
              w.vec(bx.dst + dof + static_cast<int64_t>(rt.c0 + c) * V, bx.src + so + static_cast<int64_t>(rt.c0 + c) * V, V);
            }
        continue;
      }
      for (uint32_t warp = 0; warp < nwarps; ++warp)
        for (uint32_t pc = warp; pc < npieces; pc += nwarps) {
          const uint32_t r = pc / pieces_per_row, q = pc - r * pieces_per_row;
          int64_t so, dof;
          rowOffsets(bx, rt.row0 + r, esz, so, dof);
          for (uint32_t lane = 0; lane < 32; ++lane)
            for (uint32_t k = 0; k < kUnroll; ++k) {
              const uint32_t idx = q * kPiece + lane + 32u * k;
              if (idx >= rt.nvec) continue;
              w.vec(bx.dst + dof + static_cast<int64_t>(rt.c0 + idx) * V, bx.src + so + static_cast<int64_t>(rt.c0 + idx) * V, V);
            }
        }
    }
  }
}

// transposeKernel<T, order>: 32 x 32 tiles through a (here: host) staging tile
void walkTranspose(const CopyParams& p, int grid, Walk& w) {
  const int T = static_cast<int>(p.elem_size);
  const uint32_t total = p.nboxes * p.max_tiles;
  std::vector<char> tile(32 * 33 * 16);
  std::vector<char> filled(32 * 33);
  for (uint32_t cta = 0; cta < static_cast<uint32_t>(grid); ++cta) {
    for (uint32_t t = cta; t < total; t += grid) {
      uint32_t b, j;
      slotToBoxTile(t, p.nboxes, p.max_tiles, p.peer_order, b, j);
      if (b >= p.nboxes) throw Fail("slot decodes to a box that does not exist");
      const KBox& bx = p.box[b];
      if (j >= bx.tiles) continue;
      const TransTile tt = decodeTransposeTile(bx, j);
      const char* s = bx.src + tt.i2 * bx.ss[2] * T;
      char* d = bx.dst + tt.i2 * bx.ds[2] * T;
      std::fill(filled.begin(), filled.end(), 0);
      for (uint32_t wrow = 0; wrow < 8; ++wrow)
        for (uint32_t lane = 0; lane < 32; ++lane) {
          const int64_t i0 = static_cast<int64_t>(tt.j0) * 32 + lane;
          for (uint32_t r = 0; r < 32; r += 8) {
            const int64_t i1 = static_cast<int64_t>(tt.j1) * 32 + r + wrow;
            if (i0 < bx.n[0] && i1 < bx.n[1]) {
              const char* src = s + (i0 * bx.ss[0] + i1 * bx.ss[1]) * T;
              if (reinterpret_cast<uintptr_t>(src) % T) throw Fail("misaligned element load");
              w.ranges.check(src, T, "load");
              std::memcpy(&tile[((r + wrow) * 33 + lane) * 16], src, T);
              filled[(r + wrow) * 33 + lane] = 1;
            }
          }
        }
      for (uint32_t wrow = 0; wrow < 8; ++wrow)
        for (uint32_t lane = 0; lane < 32; ++lane) {
          const int64_t i1 = static_cast<int64_t>(tt.j1) * 32 + lane;
          for (uint32_t r = 0; r < 32; r += 8) {
            const int64_t i0 = static_cast<int64_t>(tt.j0) * 32 + r + wrow;
            if (i0 < bx.n[0] && i1 < bx.n[1]) {
              if (!filled[lane * 33 + r + wrow]) throw Fail("transpose tile cell read before it was written");
              char* dst = d + (i0 * bx.ds[0] + i1 * bx.ds[1]) * T;
              if (reinterpret_cast<uintptr_t>(dst) % T) throw Fail("misaligned element store");
              w.ranges.check(dst, T, "store");
              std::memcpy(dst, &tile[(lane * 33 + r + wrow) * 16], T);
              w.bytes_written += T;
              ++w.accesses;
            }
          }
        }
    }
  }
}

// transposeVecKernel<T, order>: micro-tiles transposed in registers, swizzled shared tile of 16-byte vectors
void walkTransposeVec(const CopyParams& p, int grid, Walk& w) {
  const int T = static_cast<int>(p.elem_size);
  const int VEC = 16 / T;
  int E0, E1;
  transVecTileExtents(T, static_cast<int>(p.geometry), E0, E1);
  const uint32_t C0 = E0 / VEC, G1 = E1 / VEC, NM = C0 * G1 / 256;
  const uint32_t total = p.nboxes * p.max_tiles;
  std::vector<char> tile(1024 * 16);
  std::vector<char> filled(1024);
  for (uint32_t cta = 0; cta < static_cast<uint32_t>(grid); ++cta) {
    for (uint32_t t = cta; t < total; t += grid) {
      uint32_t b, j;
      slotToBoxTile(t, p.nboxes, p.max_tiles, p.peer_order, b, j);
      if (b >= p.nboxes) throw Fail("slot decodes to a box that does not exist");
      const KBox& bx = p.box[b];
      if (j >= bx.tiles) continue;
      if (bx.ss[0] != 1 || bx.ds[1] != 1) throw Fail("vectorised transpose needs unit strides on its contiguous axes");
      const TransTile tt = decodeTransposeTile(bx, j);
      const int64_t base0 = static_cast<int64_t>(tt.j0) * E0, base1 = static_cast<int64_t>(tt.j1) * E1;
      const char* s = bx.src + tt.i2 * bx.ss[2] * T;
      char* d = bx.dst + tt.i2 * bx.ds[2] * T;
      std::fill(filled.begin(), filled.end(), 0);
      for (uint32_t tid = 0; tid < 256; ++tid)
        for (uint32_t q = 0; q < NM; ++q) {
          const uint32_t m = tid + 256u * q, c = m % C0, g = m / C0;
          const int64_t i0 = base0 + c * VEC;
          char in[4][16];
          bool have[4] = {false, false, false, false};
          for (int k = 0; k < VEC; ++k) {
            const int64_t i1 = base1 + g * VEC + k;
            std::memset(in[k], 0, 16);
            if (i0 < bx.n[0] && i1 < bx.n[1]) {
              const char* src = s + (i0 + i1 * bx.ss[1]) * T;
              if (reinterpret_cast<uintptr_t>(src) % 16) throw Fail("misaligned 16-byte load");
              w.ranges.check(src, 16, "load");
              std::memcpy(in[k], src, 16);
              have[k] = true;
            }
          }
          for (int e = 0; e < VEC; ++e) {
            char out[16];
            bool all = true;
            for (int k = 0; k < VEC; ++k) {
              std::memcpy(out + k * T, in[k] + e * T, T);
              all = all && have[k];
            }
            const uint32_t slot = transVecSlot(c * VEC + e, g, VEC, G1);
            if (slot >= 1024) throw Fail("shared tile index out of range");
            if (filled[slot] & 2) throw Fail("shared tile vector written twice");
            std::memcpy(&tile[slot * 16], out, 16);
            filled[slot] = static_cast<char>(2 | (all ? 1 : 0));
          }
        }
      for (uint32_t tid = 0; tid < 256; ++tid)
        for (uint32_t pass = 0; pass < static_cast<uint32_t>(VEC) * NM; ++pass) {
          const uint32_t n = tid + 256u * pass, v = n % G1, row = n / G1;
          const int64_t i0 = base0 + row, i1 = base1 + static_cast<int64_t>(v) * VEC;
          if (i0 < bx.n[0] && i1 < bx.n[1]) {
            const uint32_t slot = transVecSlot(row, v, VEC, G1);
            if (!(filled[slot] & 1)) throw Fail("transpose tile vector read before it was completely written");
            char* dst = d + (i0 * bx.ds[0] + i1) * T;
            if (reinterpret_cast<uintptr_t>(dst) % 16) throw Fail("misaligned 16-byte store");
            w.ranges.check(dst, 16, "store");
            std::memcpy(dst, &tile[slot * 16], 16);
            w.bytes_written += 16;
            ++w.accesses;
          }
        }
    }
  }
}

// rowCopyBulkKernel: one thread per CTA, one bulk copy per slot
void walkBulk(const CopyParams& p, int grid, Walk& w) {
  const uint32_t total = p.nboxes * p.max_tiles;
  const int64_t esz = p.elem_size;
  for (uint32_t cta = 0; cta < static_cast<uint32_t>(grid); ++cta) {
    for (uint32_t t = cta; t < total; t += grid) {
      uint32_t b, j;
      slotToBoxTile(t, p.nboxes, p.max_tiles, p.peer_order, b, j);
      if (b >= p.nboxes) throw Fail("slot decodes to a box that does not exist");
      const KBox& bx = p.box[b];
      if (j >= bx.tiles) continue;
      if (bx.rows_per_tile != 1) throw Fail("bulk tiles must be single row segments");
      const RowTile rt = decodeRowTile(bx, j);
      int64_t so, dof;
      rowOffsets(bx, rt.row0, esz, so, dof);
      const char* src = bx.src + so + static_cast<int64_t>(rt.c0) * 16;
      char* dst = bx.dst + dof + static_cast<int64_t>(rt.c0) * 16;
      const uint32_t bytes = rt.nvec * 16u;
      // cp.async.bulk: 16-byte aligned addresses, size a multiple of 16, and the stage buffer must hold it
      if (reinterpret_cast<uintptr_t>(src) % 16 || reinterpret_cast<uintptr_t>(dst) % 16 || bytes % 16 || bytes == 0)
        throw Fail("bulk copy violates the 16-byte rules of cp.async.bulk");
      if (bytes > kBulkChunkBytes) throw Fail("bulk copy larger than a pipeline stage");
      w.ranges.check(src, bytes, "bulk load");
      w.ranges.check(dst, bytes, "bulk store");
      std::memcpy(dst, src, bytes);
      w.bytes_written += bytes;
      ++w.accesses;
    }
  }
}

} // namespace

// boxes[i] moves from src_bases[i] to dst_bases[i] (element offsets inside the box). peer_index[i]: communicator index
// of the destination rank. ranges: every address a launch may touch. grid <= 0: 370 CTAs (2.5 x 148 SMs).
// stats: [0] launches, [1] bytes written, [2] vector / bulk accesses, [3] bitmask of kernel kinds (1 row copy,
// 2 transpose, 4 bulk), [4] vector width of the row copy, [5] slots over all launches, [6] balanced grid for those slots,
// [7] longest row of a row-copy box in bytes (shows whether contiguous axes were merged).
extern "C" int cdb_emu_run_boxes(const cudecompB200Box_t* boxes, const int32_t* peer_index, int nboxes,
                                 const void* const* src_bases, void* const* dst_bases, const char* const* range_lo,
                                 const int64_t* range_len, int nranges, int es, int tile_bytes, int peer_order,
                                 int kernel_variant, int me, int comm_size, int grid, int threads, int64_t* stats,
                                 char* err, int err_len) {
  try {
    std::vector<LaunchBox> lb(nboxes);
    for (int i = 0; i < nboxes; ++i) {
      BoxDesc& d = lb[i].d;
      d.peer = peer_index ? peer_index[i] : 0;
      d.peer_world = boxes[i].peer_rank;
      d.src_off = boxes[i].src_offset;
      d.dst_off = boxes[i].dst_offset;
      for (int k = 0; k < 3; ++k) {
        d.ext[k] = boxes[i].extent[k];
        d.sstr[k] = boxes[i].src_stride[k];
        d.dstr[k] = boxes[i].dst_stride[k];
      }
      lb[i].src_base = static_cast<const char*>(src_bases[i]);
      lb[i].dst_base = static_cast<char*>(dst_bases[i]);
    }
    LaunchTuning tuning;
    tuning.tile_bytes = tile_bytes;
    tuning.peer_order = peer_order;
    tuning.kernel_variant = kernel_variant & 0xff;
    tuning.transpose_geometry = (kernel_variant >> 8) & 1; // bit 8: alternate tile geometry of the vectorised transpose
    std::vector<PreparedLaunch> launches = prepareLaunches(lb, es, tuning, me, comm_size);
    Ranges ranges{range_lo, range_len, nranges};
    Walk w{ranges};
    int64_t kinds = 0, slots = 0, longest_row = 0;
    for (auto& l : launches) {
      const CopyParams& p = l.params;
      if (l.kind != KernelKind::TRANSPOSE && l.kind != KernelKind::TRANSPOSE_VEC)
        for (uint32_t b = 0; b < p.nboxes; ++b)
          longest_row = std::max<int64_t>(longest_row, static_cast<int64_t>(p.box[b].row_vecs) * p.vec_size);
      const uint64_t total = static_cast<uint64_t>(p.nboxes) * p.max_tiles;
      slots += static_cast<int64_t>(total);
      const int dflt = (l.kind == KernelKind::ROWCOPY) ? 370 : (l.kind == KernelKind::ROWCOPY_BULK ? 148 : 592);
      const int g = chooseGrid(grid, dflt, 1 << 20, total, 0);
      if (l.kind == KernelKind::ROWCOPY) {
        kinds |= 1;
        walkRowCopy(p, g, threads > 0 ? threads : 256, w);
      } else if (l.kind == KernelKind::TRANSPOSE) {
        kinds |= 2;
        walkTranspose(p, g, w);
      } else if (l.kind == KernelKind::TRANSPOSE_VEC) {
        kinds |= 8;
        walkTransposeVec(p, g, w);
      } else {
        kinds |= 4;
        walkBulk(p, g, w);
      }
      if (stats) stats[4] = p.vec_size;
    }
    if (stats) {
      stats[0] = static_cast<int64_t>(launches.size());
      stats[1] = w.bytes_written;
      stats[2] = w.accesses;
      stats[3] = kinds;
      stats[5] = slots;
      stats[6] = chooseGrid(grid, 370, 1 << 20, static_cast<uint64_t>(slots), 1);
      stats[7] = longest_row;
    }
    return 0;
  } catch (const std::exception& e) {
    if (err && err_len > 0) std::snprintf(err, static_cast<size_t>(err_len), "%s", e.what());
    return -1;
  }
}

namespace {

// rowCopyPhasedKernel: the tile body is copyRowTile (the same loops as walkRowCopy), slots of a phase interleave its boxes.
// Walks the boxes selected by (want_unpack, step): the push boxes of phase `step`, or the unpack boxes that wait for
// `step`, so that the caller can interleave the ranks in dependency order. Also checks the phase structure itself.
void walkPhasedSelection(const PhasedLaunch& pl, int es, int lag, int want_unpack, int step, int grid, int threads, Walk& w) {
  const uint32_t nwarps = static_cast<uint32_t>(threads) >> 5;
  constexpr uint32_t kUnroll = 4, kPiece = 32 * kUnroll;
  const int V = pl.vec_size;
  for (uint32_t s = 0; s < pl.phases.size(); ++s) {
    const PhaseDesc& ph = pl.phases[s];
    const uint32_t total = ph.nsegs * ph.seg_tiles;
    for (uint32_t cta = 0; cta < static_cast<uint32_t>(grid); ++cta)
      for (uint32_t t = cta; t < total; t += grid) {
        if (ph.first_seg + t % ph.nsegs >= pl.segs.size()) throw Fail("slot decodes to a segment that does not exist");
        const SegDesc& sg = pl.segs[ph.first_seg + t % ph.nsegs];
        const uint32_t jj = t / ph.nsegs;
        if (jj >= sg.count) continue;
        if (sg.box >= pl.boxes.size()) throw Fail("segment of a box that does not exist");
        const KBox& bx = pl.boxes[sg.box];
        const uint32_t j = sg.first_tile + jj;
        if (j >= bx.tiles) throw Fail("segment reaches past the last tile of its box");
        if (sg.wait != bx.pad_) throw Fail("segment and box disagree on the dependency");
        const int need = static_cast<int>(sg.wait) - 1;
        // the step a phase belongs to = the next publication at or after it (phases behind the last one: drain)
        int phase_step = static_cast<int>(pl.nsteps);
        for (size_t q = s; q < pl.phases.size(); ++q)
          if (pl.phases[q].publish) {
            phase_step = static_cast<int>(pl.phases[q].publish) - 1;
            break;
          }
        if (need >= 0) {
          if (need >= static_cast<int>(pl.nsteps)) throw Fail("unpack box waits for a step that is never published");
          // its step must have been published by an EARLIER phase
          bool published = false;
          for (size_t q = 0; q < s; ++q)
            if (static_cast<int>(pl.phases[q].publish) == need + 1) published = true;
          if (!published) throw Fail("unpack box scheduled before the phase that publishes its step");
          if (phase_step < static_cast<int>(pl.nsteps) && phase_step != need + lag) throw Fail("unpack box in the wrong step");
        } else if (phase_step >= static_cast<int>(pl.nsteps)) {
          throw Fail("push box after the last published phase");
        }
        const bool sel = want_unpack ? (need == step) : (need < 0 && phase_step == step);
        if (!sel) continue;
        const RowTile rt = decodeRowTile(bx, j);
        const uint32_t pieces_per_row = (rt.nvec + kPiece - 1) / kPiece;
        const uint32_t npieces = rt.rows_here * pieces_per_row;
        if (bx.row_vecs <= 32u) {
          const uint32_t total_vecs = rt.rows_here * rt.nvec;
          for (uint32_t e = 0; e < total_vecs; ++e) {
            const uint32_t r = e / rt.nvec, c = e - r * rt.nvec;
            int64_t so, dof;
            rowOffsets(bx, rt.row0 + r, es, so, dof);
            w.vec(bx.dst + dof + static_cast<int64_t>(rt.c0 + c) * V, bx.src + so + static_cast<int64_t>(rt.c0 + c) * V, V);
          }
          continue;
        }
        for (uint32_t warp = 0; warp < nwarps; ++warp)
          for (uint32_t pc = warp; pc < npieces; pc += nwarps) {
            const uint32_t r = pc / pieces_per_row, q = pc - r * pieces_per_row;
            int64_t so, dof;
            rowOffsets(bx, rt.row0 + r, es, so, dof);
            for (uint32_t lane = 0; lane < 32; ++lane)
              for (uint32_t k = 0; k < kUnroll; ++k) {
                const uint32_t idx = q * kPiece + lane + 32u * k;
                if (idx >= rt.nvec) continue;
                w.vec(bx.dst + dof + static_cast<int64_t>(rt.c0 + idx) * V, bx.src + so + static_cast<int64_t>(rt.c0 + idx) * V, V);
              }
          }
      }
  }
}

} // namespace

// The fused staged schedule of ONE rank (engine.cc runFusedStaged): boxes carry step and is_unpack (a chunked plan from
// cudecompB200PlanPipelinedTransposeBoxes, or a plain staged plan with every step 0). nsteps = K. Executes only the
// selection (want_unpack, step), see walkPhasedSelection. stats: [0] phases, [1] bytes written, [2] accesses,
// [3] vector width, [4] total slots, [5] boxes. Returns 1 when the schedule cannot run as a phased launch.
extern "C" int cdb_emu_run_phased(const cudecompB200Box_t* boxes, int nboxes, int nsteps, const void* const* src_bases,
                                  void* const* dst_bases, const char* const* range_lo, const int64_t* range_len, int nranges,
                                  int es, int tile_bytes, int lag, int want_unpack, int step, int grid, int threads,
                                  int kernel_variant, int head_percent, int64_t* stats, char* err, int err_len) {
  try {
    std::vector<std::vector<LaunchBox>> push(nsteps), unpack(nsteps);
    for (int i = 0; i < nboxes; ++i) {
      LaunchBox lb;
      BoxDesc& d = lb.d;
      d.peer = 0;
      d.peer_world = boxes[i].peer_rank;
      d.src_off = boxes[i].src_offset;
      d.dst_off = boxes[i].dst_offset;
      for (int k = 0; k < 3; ++k) {
        d.ext[k] = boxes[i].extent[k];
        d.sstr[k] = boxes[i].src_stride[k];
        d.dstr[k] = boxes[i].dst_stride[k];
      }
      lb.src_base = static_cast<const char*>(src_bases[i]);
      lb.dst_base = static_cast<char*>(dst_bases[i]);
      if (boxes[i].step < 0 || boxes[i].step >= nsteps) throw Fail("box with a step outside the schedule");
      (boxes[i].is_unpack ? unpack : push)[boxes[i].step].push_back(lb);
    }
    LaunchTuning tuning;
    tuning.tile_bytes = tile_bytes;
    tuning.kernel_variant = kernel_variant;
    tuning.phase_head_percent = head_percent;
    PhasedLaunch pl;
    if (!preparePhased(push, unpack, es, tuning, lag, &pl)) return 1;
    if (pl.nsteps != static_cast<uint32_t>(nsteps) || pl.phases.size() < static_cast<size_t>(nsteps + lag))
      throw Fail("wrong number of steps / phases");
    Ranges ranges{range_lo, range_len, nranges};
    Walk w{ranges};
    walkPhasedSelection(pl, es, lag, want_unpack, step, grid > 0 ? grid : 370, threads > 0 ? threads : 256, w);
    if (stats) {
      stats[0] = static_cast<int64_t>(pl.phases.size());
      stats[1] = w.bytes_written;
      stats[2] = w.accesses;
      stats[3] = pl.vec_size;
      stats[4] = static_cast<int64_t>(pl.total_slots);
      stats[5] = static_cast<int64_t>(pl.boxes.size());
    }
    return 0;
  } catch (const std::exception& e) {
    if (err && err_len > 0) std::snprintf(err, static_cast<size_t>(err_len), "%s", e.what());
    return -1;
  }
}

// chooseGrid as the library computes it (kernels.h)
extern "C" int cdb_emu_choose_grid(int requested, int dflt, int resident, uint64_t total_slots, int balance) {
  return chooseGrid(requested, dflt, resident, total_slots, balance);
}

// slot order as the kernels decode it (tiling.h)
extern "C" void cdb_emu_slot(uint32_t t, uint32_t nboxes, uint32_t max_tiles, uint32_t peer_order, uint32_t* b, uint32_t* j) {
  slotToBoxTile(t, nboxes, max_tiles, peer_order, *b, *j);
}
