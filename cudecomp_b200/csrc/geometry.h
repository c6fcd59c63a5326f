// Pencil geometry: pure integer host arithmetic, a function of (grid, process-grid index) only, so any
// rank can compute any other rank's pencil without communication -- the push kernels need the
// destination rank's strides.
//
// Restates the arithmetic of reference src/cudecomp.cc:1317-1379 (pencil info), :1120-1150 (default
// orders, gdims_dist), :1411-1459 (workspace sizes), :1710-1755 (shifted rank) and
// include/internal/common.h:318-366,579-589,620-640 (rank <-> grid index, splits, empty pencils).
#ifndef CUDECOMP_B200_GEOMETRY_H
#define CUDECOMP_B200_GEOMETRY_H

#include <array>
#include <cstdint>
#include <vector>

#include "cudecomp.h"

namespace cdb {

enum CommAxis { COMM_COL = 0, COMM_ROW = 1 }; // COL: ranks sharing pidx[1], size pdims[0]; ROW: sharing pidx[0], size pdims[1]

struct GridGeom {
  std::array<int32_t, 3> gdims{};
  std::array<int32_t, 3> gdims_dist{};
  std::array<int32_t, 2> pdims{};
  bool col_major = false;
  int32_t order[3][3]; // [axis][memory position] -> global axis
};

struct Pencil {
  std::array<int32_t, 3> shape{}, lo{}, hi{}, order{}; // by memory position
  std::array<int32_t, 3> halo{}, pad{};                // by global axis
  int64_t size = 0;

  // extent (incl. halo and padding) of global axis g
  std::array<int32_t, 3> shapeG() const {
    std::array<int32_t, 3> s{};
    for (int i = 0; i < 3; ++i) s[order[i]] = shape[i];
    return s;
  }
  // element stride of global axis g
  std::array<int64_t, 3> strideG() const {
    std::array<int64_t, 3> s{};
    int64_t acc = 1;
    for (int i = 0; i < 3; ++i) {
      s[order[i]] = acc;
      acc *= shape[i];
    }
    return s;
  }
};

std::array<int, 2> pidxOfRank(const GridGeom& g, int rank);
int rankOfPidx(const GridGeom& g, const std::array<int, 2>& pidx);

// throws INVALID_USAGE on negative halo/padding or int32/int64 overflow
Pencil pencilInfo(const GridGeom& g, const std::array<int, 2>& pidx, int axis, const int32_t halo[3],
                  const int32_t pad[3]);

// N/n (+1 for the first N%n chunks); `pad` is added to the last populated chunk
std::vector<int64_t> getSplits(int64_t N, int nchunks, int64_t pad);
std::vector<int64_t> prefixOffsets(const std::vector<int64_t>& splits);

bool hasEmptyPencils(const GridGeom& g, int axis);
int64_t globalMaxPencilSize(const GridGeom& g, int axis);
// round an element count up to a multiple of 64 (256 bytes of the smallest dtype)
int64_t alignCount(int64_t count);
int64_t transposeWorkspaceSize(const GridGeom& g);
int64_t haloWorkspaceSize(const GridGeom& g, const std::array<int, 2>& pidx, int axis, const int32_t halo[3]);

// Which communicator a halo exchange of `dim` on `axis`-pencils uses (dim != axis)
CommAxis haloCommAxis(int axis, int dim);
// global rank of the neighbour `displacement` steps along `dim`, -1 outside a non-periodic domain
int shiftedRank(const GridGeom& g, int rank, int axis, int dim, int displacement, bool periodic);

// (ax, dir) -> axes a (source pencil), b (destination pencil), c (the other one) and the communicator
struct TransposeAxes {
  int a, b, c;
  CommAxis comm;
};
TransposeAxes transposeAxes(int ax, int dir);

// candidate process grids for `nranks`, in the autotuner's sweep order (reference src/autotune.cc:94-106)
std::vector<std::array<int32_t, 2>> pdimCandidates(int nranks, bool col_major);

} // namespace cdb

#endif
