// Two live grid descriptors on one handle, one of them destroyed while the other keeps working (C++ host code, no
// kernels of its own): the known-answer round trip of the reference's tests (every element carries its global linear
// index, tests/cc/transpose_test.cc) on both descriptors, out of place (direct peer stores) and in place (fused staged
// schedule); then descriptor A and its buffers go away and descriptor B runs the same round trips again -- its mappings
// of the peers' buffers were dropped with A and must be re-created on demand -- the in-place one on a second stream, so
// that the cached schedule tables are used from a stream other than the one that uploaded them.
// With CUDECOMP_ENABLE_CUMEM=1 in the environment all of this runs on cuMem (VMM) allocations that the peers map through
// POSIX file descriptors instead of CUDA IPC handles (csrc/vmm.h).
// Built and run by tests/test_c_caller.py (g++ against include/ and cudecomp_b200/lib, plus the CUDA runtime).
// Run on 2 (or 4) ranks with RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT set; ranks share GPUs when there are fewer.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>
#include <cudecomp.h>
#include <cudecomp_b200_ext.h>

#define CHECK(call)                                                                                                    \
  do {                                                                                                                 \
    cudecompResult_t r_ = (call);                                                                                      \
    if (r_ != CUDECOMP_RESULT_SUCCESS) {                                                                               \
      std::printf("FAILED %s: %d (line %d)\n", #call, static_cast<int>(r_), __LINE__);                                 \
      std::exit(1);                                                                                                    \
    }                                                                                                                  \
  } while (0)
#define CHECK_CUDA(call)                                                                                               \
  do {                                                                                                                 \
    cudaError_t e_ = (call);                                                                                           \
    if (e_ != cudaSuccess) {                                                                                           \
      std::printf("FAILED %s: %s (line %d)\n", #call, cudaGetErrorString(e_), __LINE__);                               \
      std::exit(1);                                                                                                    \
    }                                                                                                                  \
  } while (0)

namespace {

constexpr int NX = 96, NY = 80, NZ = 72;

struct Desc {
  cudecompGridDesc_t gd = nullptr;
  cudecompPencilInfo_t pinfo[3];
  double *a = nullptr, *b = nullptr, *work = nullptr;
  int64_t max_elems = 0;
};

std::vector<double> expected(const cudecompPencilInfo_t& p) {
  std::vector<double> v(static_cast<size_t>(p.size));
  size_t n = 0;
  for (int i2 = 0; i2 < p.shape[2]; ++i2)
    for (int i1 = 0; i1 < p.shape[1]; ++i1)
      for (int i0 = 0; i0 < p.shape[0]; ++i0) {
        int64_t g[3];
        g[p.order[0]] = p.lo[0] + i0;
        g[p.order[1]] = p.lo[1] + i1;
        g[p.order[2]] = p.lo[2] + i2;
        v[n++] = static_cast<double>(g[0] + NX * (g[1] + static_cast<int64_t>(NY) * g[2]));
      }
  return v;
}

void make(cudecompHandle_t h, int p0, int p1, Desc* d) {
  cudecompGridDescConfig_t cfg;
  CHECK(cudecompGridDescConfigSetDefaults(&cfg));
  cfg.gdims[0] = NX, cfg.gdims[1] = NY, cfg.gdims[2] = NZ;
  cfg.pdims[0] = p0, cfg.pdims[1] = p1;
  cfg.transpose_comm_backend = CUDECOMP_TRANSPOSE_COMM_NCCL;
  CHECK(cudecompGridDescCreate(h, &d->gd, &cfg, nullptr));
  for (int ax = 0; ax < 3; ++ax) {
    CHECK(cudecompGetPencilInfo(h, d->gd, &d->pinfo[ax], ax, nullptr, nullptr));
    if (d->pinfo[ax].size > d->max_elems) d->max_elems = d->pinfo[ax].size;
  }
  int64_t work_elems = 0;
  CHECK(cudecompGetTransposeWorkspaceSize(h, d->gd, &work_elems));
  CHECK(cudecompMalloc(h, d->gd, reinterpret_cast<void**>(&d->a), d->max_elems * sizeof(double)));
  CHECK(cudecompMalloc(h, d->gd, reinterpret_cast<void**>(&d->b), d->max_elems * sizeof(double)));
  CHECK(cudecompMalloc(h, d->gd, reinterpret_cast<void**>(&d->work), work_elems * sizeof(double)));
}

bool same(const double* dev, const cudecompPencilInfo_t& p, cudaStream_t s, const char* what, int rank) {
  std::vector<double> got(static_cast<size_t>(p.size)), want = expected(p);
  CHECK_CUDA(cudaStreamSynchronize(s));
  CHECK_CUDA(cudaMemcpy(got.data(), dev, got.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < got.size(); ++i)
    if (got[i] != want[i]) {
      std::printf("rank %d: %s differs at %zu: got %.1f, expected %.1f\n", rank, what, i, got[i], want[i]);
      return false;
    }
  return true;
}

// X -> Y -> Z -> Y -> X, every pencil checked against the known answer.
bool roundTrip(cudecompHandle_t h, Desc& d, bool inplace, cudaStream_t s, int rank) {
  std::vector<double> x = expected(d.pinfo[0]);
  CHECK_CUDA(cudaMemcpy(d.a, x.data(), x.size() * sizeof(double), cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemset(d.b, 0xff, d.max_elems * sizeof(double)));
  CHECK_CUDA(cudaDeviceSynchronize()); // `s` may be a non-blocking stream: the fills are done before the first launch
  double* cur = d.a;
  double* other = inplace ? d.a : d.b;
  bool ok = true;
  auto step = [&](cudecompResult_t r, int ax, const char* what) {
    CHECK(r);
    ok = same(other, d.pinfo[ax], s, what, rank) && ok;
    if (!inplace) std::swap(cur, other);
  };
  const cudecompDataType_t T = CUDECOMP_DOUBLE;
  step(cudecompTransposeXToY(h, d.gd, cur, other, d.work, T, nullptr, nullptr, nullptr, nullptr, s), 1, "XToY");
  step(cudecompTransposeYToZ(h, d.gd, cur, other, d.work, T, nullptr, nullptr, nullptr, nullptr, s), 2, "YToZ");
  step(cudecompTransposeZToY(h, d.gd, cur, other, d.work, T, nullptr, nullptr, nullptr, nullptr, s), 1, "ZToY");
  step(cudecompTransposeYToX(h, d.gd, cur, other, d.work, T, nullptr, nullptr, nullptr, nullptr, s), 0, "YToX");
  return ok;
}

} // namespace

int main() {
  if (MPI_Init(nullptr, nullptr) != MPI_SUCCESS) return 1;
  int rank = -1, size = -1;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  if (size != 2 && size != 4) {
    std::printf("needs 2 or 4 ranks\n");
    return 1;
  }
  int ndev = 0;
  CHECK_CUDA(cudaGetDeviceCount(&ndev));
  CHECK_CUDA(cudaSetDevice(rank % ndev));

  cudecompHandle_t h;
  CHECK(cudecompInit(&h, MPI_COMM_WORLD));
  int32_t cumem = -1; // CUDECOMP_ENABLE_CUMEM=1: cudecompMalloc hands out cuMem allocations, mapped by the peers through fds
  CHECK(cudecompB200GetCumemState(h, &cumem));
  Desc A, B;
  make(h, size, 1, &A);     // X<->Y on the wire
  make(h, size / 2, 2, &B); // 2 ranks: Y<->Z on the wire; 4 ranks: both
  cudaStream_t side;
  CHECK_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));

  bool ok = true;
  ok = roundTrip(h, A, false, nullptr, rank) && ok;
  ok = roundTrip(h, B, false, nullptr, rank) && ok;
  ok = roundTrip(h, A, true, nullptr, rank) && ok;
  ok = roundTrip(h, B, true, nullptr, rank) && ok;

  // A goes away, buffers first (deferred frees: the peers still map them), then the descriptor (collective).
  CHECK(cudecompFree(h, A.gd, A.a));
  CHECK(cudecompFree(h, A.gd, A.b));
  CHECK(cudecompFree(h, A.gd, A.work));
  CHECK(cudecompGridDescDestroy(h, A.gd));

  ok = roundTrip(h, B, false, nullptr, rank) && ok;
  ok = roundTrip(h, B, true, side, rank) && ok; // cached tables, uploaded on the default stream
  ok = roundTrip(h, B, true, nullptr, rank) && ok;
  ok = roundTrip(h, B, false, side, rank) && ok;

  CHECK(cudecompFree(h, B.gd, B.a));
  CHECK(cudecompFree(h, B.gd, B.b));
  CHECK(cudecompFree(h, B.gd, B.work));
  CHECK(cudecompGridDescDestroy(h, B.gd));
  CHECK_CUDA(cudaStreamDestroy(side));
  CHECK(cudecompFinalize(h));
  MPI_Finalize();
  if (!ok) return 1;
  std::printf("two descriptors OK rank %d of %d, cumem state %d\n", rank, size, static_cast<int>(cumem));
  return 0;
}
