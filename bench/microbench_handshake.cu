// How long does the in-kernel cross-GPU entry/exit handshake of cdb::rowCopyKernel take?
// Single process, 2 GPUs with peer access; both GPUs launch the same (almost empty) kernel with handshake in a loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I include/mpi_shim -I cudecomp_b200/csrc \
//        bench/microbench_handshake.cu cudecomp_b200/csrc/{launch_params,plan,geometry}.cc -o /tmp/hs && /tmp/hs
// --profile-remote: three launches without handshake for ncu (kernel replay is safe): the product's row-copy kernel
// storing 1 GiB into the peer GPU with 128-bit and with 256-bit accesses, then the same box locally.
#include "../cudecomp_b200/csrc/kernels.cu"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CK(x)                                                                                                          \
  do {                                                                                                                 \
    cudaError_t e = (x);                                                                                               \
    if (e != cudaSuccess) {                                                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);                                   \
      exit(1);                                                                                                         \
    }                                                                                                                  \
  } while (0)

using namespace cdb;

int main(int argc, char** argv) {
  const bool profile_remote = argc > 1 && std::string(argv[1]) == "--profile-remote";
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 2) {
    printf("needs 2 GPUs\n");
    return 0;
  }
  uint64_t* pad[2];
  char *src[2], *dst[2];
  cudaStream_t st[2];
  cudaEvent_t e0[2], e1[2];
  const size_t bytes = profile_remote ? (1024ull << 20) : (64ull << 20);
  for (int d = 0; d < 2; ++d) {
    CK(cudaSetDevice(d));
    CK(cudaDeviceEnablePeerAccess(1 - d, 0));
    CK(cudaMalloc(&pad[d], 4096));
    CK(cudaMemset(pad[d], 0, 4096));
    CK(cudaMalloc(&src[d], bytes));
    CK(cudaMalloc(&dst[d], bytes));
    CK(cudaStreamCreate(&st[d]));
    CK(cudaEventCreate(&e0[d]));
    CK(cudaEventCreate(&e1[d]));
  }
  for (int d = 0; d < 2; ++d) {
    CK(cudaSetDevice(d));
    CK(cudaDeviceSynchronize());
  }
  uint64_t epoch = 0;
  auto run = [&](const char* name, int tiles, bool handshake, bool remote, int grid, int iters, int vec = 16) {
    for (int rep = 0; rep < 2; ++rep) {
      for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaEventRecord(e0[d], st[d]));
      }
      for (int it = 0; it < iters; ++it) {
        ++epoch;
        for (int d = 0; d < 2; ++d) {
          CK(cudaSetDevice(d));
          CopyParams p;
          memset(&p, 0, sizeof(p));
          p.elem_size = 16;
          p.vec_size = vec;
          if (tiles > 0) {
            KBox& b = p.box[0];
            b.src = src[d];
            b.dst = remote ? dst[1 - d] : dst[d];
            b.n[0] = 2048ll * tiles;
            b.n[1] = b.n[2] = 1;
            b.row_vecs = 2048u * 16u / vec * tiles;
            b.seg_vecs = 2048u * 16u / vec;
            b.segs_per_row = tiles;
            b.rows_per_tile = 1;
            b.tiles = tiles;
            p.nboxes = 1;
            p.max_tiles = tiles;
          }
          if (handshake) {
            p.sync.my_pad = pad[d];
            p.sync.peer_pad[0] = pad[1 - d];
            p.sync.peer_world[0] = 1 - d;
            p.sync.npeers = 1;
            p.sync.my_world = d;
            p.sync.epoch = epoch;
            p.sync.do_entry = p.sync.do_exit = 1;
            p.sync.timeout_ns = 5000000000ull;
          }
          LaunchConfig cfg;
          cfg.grid = grid;
          CK(launchCopy(KernelKind::ROWCOPY, p, cfg, st[d]));
        }
      }
      float worst = 0;
      for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaEventRecord(e1[d], st[d]));
      }
      for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaStreamSynchronize(st[d]));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0[d], e1[d]));
        if (ms > worst) worst = ms;
      }
      if (rep == 1) printf("%-60s %8.2f us per launch\n", name, worst * 1e3 / iters);
    }
  };
  if (profile_remote) {
    // for ncu (single process, no handshake so that kernel replay is safe): cdb::rowCopyKernel<uint4> storing 1 GiB
    // into the peer GPU, then the same box locally
    // 4 launches each (2 repetitions x 2 GPUs): launches 0-3, 4-7, 8-11 of the process
    run("remote copy 32768 tiles (1 GiB), no handshake, default grid", 32768, false, true, 0, 1);
    run("remote copy 32768 tiles (1 GiB), 256-bit accesses, no handshake, default grid", 32768, false, true, 0, 1, 32);
    run("local copy 32768 tiles (1 GiB), no handshake, default grid", 32768, false, false, 0, 1);
    return 0;
  }
  const int it = 300;
  run("empty kernel, no handshake, 1 CTA", 0, false, false, 1, it);
  run("handshake only, 1 CTA", 0, true, false, 1, it);
  run("local copy 370 tiles (12 MB), no handshake, 370 CTAs", 370, false, false, 370, it);
  run("local copy 370 tiles + handshake, 370 CTAs", 370, true, false, 370, it);
  run("local copy 370 tiles + handshake, 148 CTAs", 370, true, false, 148, it);
  run("local copy 370 tiles + handshake, 37 CTAs", 370, true, false, 37, it);
  run("remote copy 370 tiles, no handshake, 370 CTAs", 370, false, true, 370, it);
  run("remote copy 370 tiles + handshake, 370 CTAs", 370, true, true, 370, it);
  run("remote copy 37 tiles + handshake, 37 CTAs", 37, true, true, 37, it);
  run("remote copy 2048 tiles (67 MB) no handshake, 370 CTAs", 2048, false, true, 370, it);
  run("remote copy 2048 tiles (67 MB) + handshake, 370 CTAs", 2048, true, true, 370, it);
  return 0;
}
