// Transfer plans: which 3-D sub-block ("box") of my pencil goes where.
//
// The reference runs every transpose as pack -> all-to-all -> unpack through two staging regions
// (reference include/internal/transpose.h:196-905). Here a transpose is described once as P boxes,
// one per rank of the row/column communicator: the part of my source pencil whose `a` coordinate
// belongs to peer i, addressed with my strides on the source side and with PEER i's strides on the
// destination side. A single kernel then gathers each box from local HBM and stores it straight into
// the peer's memory in its final layout, so pack, exchange and unpack are one pass.
//
// Halo exchange (reference include/internal/halo.h:40-315) uses the same representation with at most
// two boxes (the faces for the left and right neighbour).
#ifndef CUDECOMP_B200_PLAN_H
#define CUDECOMP_B200_PLAN_H

#include <array>
#include <cstdint>
#include <vector>

#include "geometry.h"

namespace cdb {

// One box in element units; extents and strides are indexed by GLOBAL axis.
struct BoxDesc {
  int peer = 0;        // index of the destination rank inside the communicator
  int peer_world = 0;  // its global rank
  int64_t src_off = 0; // element offset of the first element in the source buffer
  int64_t dst_off = 0; // element offset of the first element in the destination buffer
  std::array<int64_t, 3> ext{};
  std::array<int64_t, 3> sstr{};
  std::array<int64_t, 3> dstr{};
  int64_t count() const { return ext[0] * ext[1] * ext[2]; }
};

enum class DstKind {
  FINAL, // the peer's user buffer (output pencil / halo region), final layout incl. halos and padding
  STAGE  // the peer's workspace: dense destination pencil (transpose) or dense face slots (halo)
};

struct TransposePlan {
  TransposeAxes axes{};
  int comm_size = 1;
  int me = 0;                   // my index in the communicator
  std::vector<int> group_world; // global rank of each communicator member
  bool noop = false;            // single rank, in place, identical layouts: nothing to do
  std::vector<BoxDesc> push;    // one per peer (self included), source = input
  std::vector<BoxDesc> unpack;  // STAGE only: dense pencil in my workspace -> output
  int64_t src_elems = 0;        // elements of my source pencil interior (= what this rank sends, self included)
  int64_t wire_elems = 0;       // elements that leave this GPU
};

// Builds the sender-side plan of transpose (ax, dir) for the rank at `pidx`.
// Throws NOT_SUPPORTED for decompositions with empty pencils (reference transpose.h:257-259).
TransposePlan buildTransposePlan(const GridGeom& g, const std::array<int, 2>& pidx, int ax, int dir,
                                 const int32_t in_halo[3], const int32_t out_halo[3], const int32_t in_pad[3],
                                 const int32_t out_pad[3], DstKind kind, bool inplace);

// Receiver-driven ("pull") variant of the direct plan: one box per SOURCE rank of my communicator, addressed with that
// rank's strides on the source side (its input pencil, mapped into my address space) and with MY strides on the
// destination side (my output pencil). Box j is exactly the box rank j would push to me, so the union over all ranks
// moves the same cells as the sender-driven plan. `peer` / `peer_world` name the rank that owns the SOURCE buffer.
// With kind == STAGE the destination is the dense pencil in MY workspace (then `unpack` holds the local copy into the
// output, as in the sender-driven staged plan): nobody writes into a peer's memory at all.
TransposePlan buildPullTransposePlan(const GridGeom& g, const std::array<int, 2>& pidx, int ax, int dir,
                                     const int32_t in_halo[3], const int32_t out_halo[3], const int32_t in_pad[3],
                                     const int32_t out_pad[3], DstKind kind = DstKind::FINAL, bool inplace = false);

// Chunked schedule of a staged transpose: the pencil is cut into K chunks along the slowest axis of the SOURCE memory
// order; step s pushes chunk s to every peer's workspace and, once that has landed, unpacks every piece whose
// destination memory is free by then. In place, "free" means the push has already consumed the source planes the piece
// overwrites (pieces that land beyond the consumed prefix wait for a later step); out of place a piece is unpacked in the
// step it arrives. Consecutive steps overlap on the device: unpack(s) runs beside push(s+1).
struct PipelineStep {
  std::vector<BoxDesc> push;   // sub-boxes of TransposePlan::push restricted to chunk s
  std::vector<BoxDesc> unpack; // pieces of the dense pencil in my workspace -> output that become writable in step s
};
struct PipelinedPlan {
  TransposePlan base;              // the unchunked staged plan (geometry, group, noop)
  int chunk_axis = -1;             // global axis the chunks run along
  std::vector<PipelineStep> steps; // empty: not applicable, run `base` as is
};
// pull == true: the receiver-driven mirror image. Step s then holds, per source rank j, the slice rank j would have
// pushed to me in its step s (its source strides, `peer` = j); the unpack lists are unchanged, because the same data
// lands in my workspace at the same step and the planes of my pencil that are still unread after step s are the same
// (they are now read by the peers' later steps instead of by mine).
// elem_bytes > 0 allows column chunks (chunks along the fastest axis when it takes no part in the transpose, see
// plan.cc) with the largest chunk count <= nchunks whose rows are whole multiples of min_row_bytes; 0 keeps plane chunks.
constexpr int64_t kMinChunkRowBytes = 2048;
PipelinedPlan buildPipelinedTransposePlan(const GridGeom& g, const std::array<int, 2>& pidx, int ax, int dir,
                                          const int32_t in_halo[3], const int32_t out_halo[3], const int32_t in_pad[3],
                                          const int32_t out_pad[3], bool inplace, int nchunks, bool pull = false,
                                          int elem_bytes = 0, int64_t min_row_bytes = kMinChunkRowBytes);

struct HaloPlan {
  bool nothing = false; // zero halo width, or no neighbour in this dimension
  CommAxis comm = COMM_COL;
  int comm_size = 1;
  int me = 0;
  std::vector<int> group_world;
  std::array<int, 2> neighbor{-1, -1}; // communicator index of the left / right neighbour (-1: none)
  std::vector<BoxDesc> push;           // faces I send; peer may be myself (periodic, single rank)
  std::vector<BoxDesc> unpack;         // STAGE only: face slots in my workspace -> my halo cells
  int64_t face_elems = 0;
};

HaloPlan buildHaloPlan(const GridGeom& g, const std::array<int, 2>& pidx, int ax, int dim, const int32_t halo[3],
                       const bool periods[3], const int32_t pad[3], DstKind kind);

// A box reduced to what a kernel needs: axes sorted by source stride, unit axes dropped, adjacent
// axes merged when both sides allow it.
struct CanonBox {
  int nd = 1;                  // 1..3
  std::array<int64_t, 3> n{1, 1, 1};
  std::array<int64_t, 3> ss{1, 0, 0};
  std::array<int64_t, 3> ds{1, 0, 0};
  bool rowCopy() const { return ss[0] == 1 && ds[0] == 1; } // contiguous runs on both sides
  int dstUnitAxis() const {                                 // axis with unit destination stride, -1 if none
    for (int k = 0; k < nd; ++k)
      if (ds[k] == 1) return k;
    return -1;
  }
};
CanonBox canonicalize(const BoxDesc& b, bool merge);

} // namespace cdb

#endif
