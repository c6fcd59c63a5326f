// See geometry.h for the reference lines each function restates.
#include "geometry.h"

#include <algorithm>
#include <limits>

#include "errors.h"

namespace cdb {

std::array<int, 2> pidxOfRank(const GridGeom& g, int rank) {
  // reference include/internal/common.h:318-331
  if (g.col_major) return {rank % g.pdims[0], rank / g.pdims[0]};
  return {rank / g.pdims[1], rank % g.pdims[1]};
}

int rankOfPidx(const GridGeom& g, const std::array<int, 2>& pidx) {
  // inverse of pidxOfRank (reference include/internal/common.h:334-346)
  if (g.col_major) return pidx[0] + pidx[1] * g.pdims[0];
  return pidx[0] * g.pdims[1] + pidx[1];
}

static int32_t checkedShape(int64_t v) {
  if (v < 0) THROW_INVALID_USAGE("computed pencil shape values must be non-negative");
  if (v > std::numeric_limits<int32_t>::max()) THROW_INVALID_USAGE("computed pencil shape exceeds int32_t limit");
  return static_cast<int32_t>(v);
}

Pencil pencilInfo(const GridGeom& g, const std::array<int, 2>& pidx, int axis, const int32_t halo[3],
                  const int32_t pad[3]) {
  Pencil p;
  std::array<int, 3> inv{};
  for (int i = 0; i < 3; ++i) {
    p.order[i] = g.order[axis][i];
    inv[p.order[i]] = i;
  }
  p.size = 1;
  int j = 0; // counts the non-axis dims: the first is split by pdims[0], the second by pdims[1]
  for (int i = 0; i < 3; ++i) {
    const int pos = inv[i];
    int64_t extent, lo;
    if (i != axis) {
      const int64_t nd = g.gdims_dist[i], np = g.pdims[j], me = pidx[j];
      const int64_t q = nd / np, r = nd % np;
      extent = q + (me < r ? 1 : 0);
      // whatever gdims has beyond gdims_dist goes to the last rank that owns anything
      if (me == std::min<int64_t>(np, nd) - 1) extent += g.gdims[i] - nd;
      lo = me * q + std::min<int64_t>(me, r);
      ++j;
    } else {
      extent = g.gdims[i];
      lo = 0;
    }
    const int32_t e32 = checkedShape(extent);
    p.lo[pos] = static_cast<int32_t>(lo);
    p.hi[pos] = static_cast<int32_t>(lo + e32 - 1);
    p.halo[i] = halo ? halo[i] : 0;
    p.pad[i] = pad ? pad[i] : 0;
    if (p.halo[i] < 0) THROW_INVALID_USAGE("halo_extents values must be non-negative");
    if (p.pad[i] < 0) THROW_INVALID_USAGE("padding values must be non-negative");
    p.shape[pos] = checkedShape(static_cast<int64_t>(e32) + 2 * static_cast<int64_t>(p.halo[i]) + p.pad[i]);
    if (p.size != 0 && p.shape[pos] != 0 && p.shape[pos] > std::numeric_limits<int64_t>::max() / p.size)
      THROW_INVALID_USAGE("computed pencil size exceeds int64_t limit");
    p.size *= p.shape[pos];
  }
  return p;
}

std::vector<int64_t> getSplits(int64_t N, int nchunks, int64_t pad) {
  std::vector<int64_t> s(nchunks, N / nchunks);
  for (int i = 0; i < N % nchunks; ++i) s[i] += 1;
  s[std::min<int64_t>(N, nchunks) - 1] += pad;
  return s;
}

std::vector<int64_t> prefixOffsets(const std::vector<int64_t>& splits) {
  std::vector<int64_t> o(splits.size(), 0);
  for (size_t i = 1; i < splits.size(); ++i) o[i] = o[i - 1] + splits[i - 1];
  return o;
}

bool hasEmptyPencils(const GridGeom& g, int axis) {
  int j = 0;
  for (int i = 0; i < 3; ++i) {
    if (i == axis) continue;
    if (g.gdims_dist[i] / g.pdims[j] == 0) return true;
    ++j;
  }
  return false;
}

int64_t globalMaxPencilSize(const GridGeom& g, int axis) {
  int64_t size = 1;
  int j = 0;
  for (int i = 0; i < 3; ++i) {
    if (i != axis) {
      int64_t d = (g.gdims_dist[i] + g.pdims[j] - 1) / g.pdims[j];
      d += g.gdims[i] - g.gdims_dist[i];
      size *= d;
      ++j;
    } else {
      size *= g.gdims[i];
    }
  }
  return size;
}

int64_t alignCount(int64_t count) { return (count + 63) / 64 * 64; }

int64_t transposeWorkspaceSize(const GridGeom& g) {
  const int64_t x = globalMaxPencilSize(g, 0), y = globalMaxPencilSize(g, 1), z = globalMaxPencilSize(g, 2);
  return std::max({alignCount(x) + y, alignCount(y) + x, alignCount(y) + z, alignCount(z) + y});
}

int64_t haloWorkspaceSize(const GridGeom& g, const std::array<int, 2>& pidx, int axis, const int32_t halo[3]) {
  Pencil p = pencilInfo(g, pidx, axis, halo, nullptr);
  auto s = p.shapeG();
  int64_t best = 0;
  for (int d = 0; d < 3; ++d) {
    int64_t face = static_cast<int64_t>(s[(d + 1) % 3]) * s[(d + 2) % 3] * p.halo[d];
    best = std::max(best, 4 * alignCount(face));
  }
  return best;
}

CommAxis haloCommAxis(int axis, int dim) {
  // dim is the first or the second non-axis dimension
  int count = 0;
  for (int i = 0; i < 3; ++i) {
    if (i == axis) continue;
    if (i == dim) break;
    ++count;
  }
  return count == 0 ? COMM_COL : COMM_ROW;
}

int shiftedRank(const GridGeom& g, int rank, int axis, int dim, int displacement, bool periodic) {
  if (displacement == 0) return rank;
  if (dim == axis) return periodic ? rank : -1;
  const CommAxis ca = haloCommAxis(axis, dim);
  auto pidx = pidxOfRank(g, rank);
  const int n = g.pdims[ca];
  int shifted = pidx[ca] + displacement;
  if (!periodic && (shifted < 0 || shifted >= n)) return -1;
  // the reference adds one period before the modulo, which leaves displacements below -n negative
  // (reference src/cudecomp.cc:1747); do a true modulo instead, identical for |displacement| <= n
  shifted = ((shifted % n) + n) % n;
  pidx[ca] = shifted;
  return rankOfPidx(g, pidx);
}

TransposeAxes transposeAxes(int ax, int dir) {
  TransposeAxes t;
  t.a = ax;
  t.b = (dir > 0 ? ax + 1 : ax + 2) % 3;
  t.c = (dir > 0 ? ax + 2 : ax + 1) % 3;
  t.comm = (t.a == 2 || t.b == 2) ? COMM_ROW : COMM_COL;
  return t;
}

std::vector<std::array<int32_t, 2>> pdimCandidates(int nranks, bool col_major) {
  std::vector<int> factors;
  for (int i = 1; static_cast<int64_t>(i) * i <= nranks; ++i) {
    if (nranks % i == 0) {
      factors.push_back(i);
      if (nranks / i != i) factors.push_back(nranks / i);
    }
  }
  std::sort(factors.begin(), factors.end());
  std::vector<std::array<int32_t, 2>> out;
  for (int f : factors) {
    if (col_major)
      out.push_back({f, nranks / f});
    else
      out.push_back({nranks / f, f});
  }
  return out;
}

} // namespace cdb
