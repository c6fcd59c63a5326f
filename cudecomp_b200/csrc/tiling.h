// Tile decoding shared by the device kernels (kernels.cu) and by host code that needs to know exactly which bytes a
// launch moves in which order (launch_params.cc; the host-side launch emulator under tests/host_emu/ that checks the
// index arithmetic of every kernel against the oracle without a GPU).
//
// A launch is a list of `nboxes * max_tiles` slots walked grid-stride by persistent CTAs. A slot maps to one tile of
// one box; slots past a box's last tile are empty. Two slot orders exist (CopyParams::peer_order):
//   0  interleaved ("one-shot"): consecutive slots belong to different boxes, so all peers are fed all the time;
//   1  rounds ("pairwise"): every nboxes-th slot belongs to box 0 (the host puts the local box there, so local HBM
//      traffic stays spread over the whole launch) and the slots in between walk boxes 1, 2, ... one after the other:
//      at any moment the whole GPU stores into ONE peer, and the host orders the boxes (me+1, me+2, ...) so that no
//      two ranks of a communicator target the same peer in the same round.
#ifndef CUDECOMP_B200_TILING_H
#define CUDECOMP_B200_TILING_H

#include "kernels.h"

#ifdef __CUDACC__
#define CDB_HD __host__ __device__ __forceinline__
#else
#define CDB_HD inline
#endif

namespace cdb {

CDB_HD void slotToBoxTile(uint32_t t, uint32_t nboxes, uint32_t max_tiles, uint32_t peer_order, uint32_t& b, uint32_t& j) {
  if (peer_order == 0u || nboxes == 1u) {
    b = t % nboxes;
    j = t / nboxes;
    return;
  }
  const uint32_t slot = t % nboxes;
  const uint32_t round = t / nboxes;
  if (slot == 0u) {
    b = 0u;
    j = round;
    return;
  }
  const uint64_t q = static_cast<uint64_t>(round) * (nboxes - 1u) + (slot - 1u);
  b = 1u + static_cast<uint32_t>(q / max_tiles);
  j = static_cast<uint32_t>(q % max_tiles);
}

// ROWCOPY / ROWCOPY_BULK: tile j of a box = rows [row0, row0 + rows_here) x vectors [c0, c0 + nvec) of each row
struct RowTile {
  int64_t row0;
  uint32_t rows_here;
  uint32_t c0;
  uint32_t nvec;
};

CDB_HD RowTile decodeRowTile(const KBox& bx, uint32_t j) {
  RowTile rt;
  const uint32_t seg = j % bx.segs_per_row;
  const uint32_t row_tile = j / bx.segs_per_row;
  rt.row0 = static_cast<int64_t>(row_tile) * bx.rows_per_tile;
  const int64_t nrows = bx.n[1] * bx.n[2];
  rt.c0 = seg * bx.seg_vecs;
  const uint32_t rest = bx.row_vecs - rt.c0;
  rt.nvec = bx.seg_vecs < rest ? bx.seg_vecs : rest;
  const int64_t left = nrows - rt.row0;
  rt.rows_here = static_cast<uint32_t>(static_cast<int64_t>(bx.rows_per_tile) < left ? static_cast<int64_t>(bx.rows_per_tile) : left);
  return rt;
}

// byte offsets of row `row` of a ROWCOPY box on the source / destination side
CDB_HD void rowOffsets(const KBox& bx, int64_t row, int64_t esz, int64_t& src_bytes, int64_t& dst_bytes) {
  const int64_t i1 = row % bx.n[1];
  const int64_t i2 = row / bx.n[1];
  src_bytes = (i1 * bx.ss[1] + i2 * bx.ss[2]) * esz;
  dst_bytes = (i1 * bx.ds[1] + i2 * bx.ds[2]) * esz;
}

// TRANSPOSE: tile j = 32 x 32 elements at (j0 * 32, j1 * 32) of plane i2
struct TransTile {
  uint32_t j0;
  uint32_t j1;
  int64_t i2;
};

// TRANSPOSE_VEC: a tile is kE0 elements along the source-contiguous axis x kE1 along the destination-contiguous axis,
// 16 KiB for every element size. It is loaded as kC0 x kG1 micro-tiles of VEC x VEC elements (VEC rows of one 16-byte
// vector each), kNM per thread of a 256-thread CTA, and written out as kE0 rows of kG1 16-byte vectors. The shared tile
// holds those rows; vector v of row i0 sits at v ^ ((i0 / VEC) & 7), which makes both the micro-tile stores (8
// consecutive lanes = 8 consecutive vector columns c = i0 / VEC) and the row reads (8 consecutive v) conflict-free.
// kAlt (8-byte elements only): the tile is 64 x 32 instead of 32 x 64 elements, i.e. 512-byte row segments on the LOAD
// side and 256-byte ones on the store side instead of the other way round.
template <int kElemBytes, int kAlt = 0> struct TransVecGeom {
  static constexpr int kVec = 16 / kElemBytes;
  static constexpr int kE0 = (kElemBytes == 16) ? 32 : ((kElemBytes == 8 && kAlt) ? 64 : 16 * kVec);
  static constexpr int kE1 = (kElemBytes == 16) ? 32 : ((kElemBytes == 8 && kAlt) ? 32 : 64);
  static constexpr int kC0 = kE0 / kVec;
  static constexpr int kG1 = kE1 / kVec;
  static constexpr int kNM = kC0 * kG1 / 256;
  static_assert(kC0 * kG1 % 256 == 0 && kE0 * kG1 == 1024 && kC0 % 8 == 0 && kG1 % 8 == 0, "tile geometry");
};
CDB_HD void transVecTileExtents(int elem_bytes, int alt, int& e0, int& e1) {
  e0 = (elem_bytes == 16) ? 32 : ((elem_bytes == 8 && alt) ? 64 : 16 * (16 / elem_bytes));
  e1 = (elem_bytes == 16) ? 32 : ((elem_bytes == 8 && alt) ? 32 : 64);
}
CDB_HD uint32_t transVecSlot(uint32_t row, uint32_t v, uint32_t vec, uint32_t g1) { return row * g1 + (v ^ ((row / vec) & 7u)); }

CDB_HD TransTile decodeTransposeTile(const KBox& bx, uint32_t j) {
  TransTile tt;
  tt.j0 = j % bx.tiles0;
  tt.j1 = (j / bx.tiles0) % bx.tiles1;
  tt.i2 = j / (bx.tiles0 * bx.tiles1);
  return tt;
}

} // namespace cdb

#endif
