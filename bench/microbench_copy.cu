// Microbenchmark behind the kernel design choices in DESIGN.md: what is the fastest way for SMs to copy a large
// contiguous region (a) inside one GPU's HBM and (b) into a peer GPU over NVLink?
// Single process; with >= 2 GPUs it enables peer access and also measures both directions at once (the transposes
// always send and receive simultaneously).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo bench/microbench_copy.cu -o /tmp/mb && /tmp/mb
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                                          \
  do {                                                                                                                 \
    cudaError_t e = (x);                                                                                               \
    if (e != cudaSuccess) {                                                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);                                   \
      exit(1);                                                                                                         \
    }                                                                                                                  \
  } while (0)

// ------------------------------------------------------------------------------------------------ SIMT variants
enum Hint { CS = 0, PLAIN = 1, NC = 2 };

template <int H> __device__ __forceinline__ uint4 ld(const uint4* p) {
  if (H == CS) return __ldcs(p);
  if (H == NC) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
  }
  return *p;
}
template <int H> __device__ __forceinline__ void st(uint4* p, uint4 v) {
  if (H == CS)
    __stcs(p, v);
  else if (H == NC)
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
  else
    *p = v;
}

// tile = 32 KiB per CTA iteration, warp pieces of 32*U vectors (the shape of cdb::rowCopyKernel)
template <int U, int H> __global__ void __launch_bounds__(512) simtCopy(const uint4* src, uint4* dst, size_t nvec) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t tile = 2048;
  const size_t ntiles = (nvec + tile - 1) / tile;
  for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const size_t base = t * tile;
    const uint32_t n = (uint32_t)((nvec - base < tile) ? nvec - base : tile);
    for (uint32_t p = warp * 32 * U; p < n; p += nw * 32 * U) {
      uint4 v[U];
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (p + lane + 32 * k < n) v[k] = ld<H>(src + base + p + lane + 32 * k);
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (p + lane + 32 * k < n) st<H>(dst + base + p + lane + 32 * k, v[k]);
    }
  }
}

// 256-bit accesses (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256): the same tile shape with 32-byte vectors, i.e. one
// warp instruction covers 1 KiB. Question for the NVLink path: do wider stores raise the SM-store ceiling?
struct alignas(32) V32 {
  uint64_t a, b, c, d;
};
__device__ __forceinline__ V32 ld256(const V32* p) {
  V32 v;
  asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
  return v;
}
__device__ __forceinline__ void st256(V32* p, V32 v) {
  asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v.a), "l"(v.b), "l"(v.c), "l"(v.d) : "memory");
}
template <int U> __global__ void __launch_bounds__(256) simtCopy256(const V32* src, V32* dst, size_t nvec) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t tile = 1024; // 32 KiB
  const size_t ntiles = (nvec + tile - 1) / tile;
  for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const size_t base = t * tile;
    const uint32_t n = (uint32_t)((nvec - base < tile) ? nvec - base : tile);
    for (uint32_t p = warp * 32 * U; p < n; p += nw * 32 * U) {
      V32 v[U];
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (p + lane + 32 * k < n) v[k] = ld256(src + base + p + lane + 32 * k);
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (p + lane + 32 * k < n) st256(dst + base + p + lane + 32 * k, v[k]);
    }
  }
}

// P=2 transpose shape: half of the tiles stay local, half go to the peer.
// mode 0: tiles alternate local/peer inside every CTA; mode 1: the first `peer_ctas` CTAs only push, the rest only copy.
template <int U> __global__ void __launch_bounds__(256) mixedCopy(const uint4* src, uint4* dst_local, uint4* dst_peer,
                                                                  size_t nvec, int mode, int peer_ctas) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t tile = 2048;
  const size_t ntiles = (nvec + tile - 1) / tile; // first half of the tiles -> peer, second half -> local
  const size_t half = ntiles / 2;
  size_t t0, step, tend;
  bool peer_only = false, local_only = false;
  if (mode == 0) {
    t0 = blockIdx.x;
    step = gridDim.x;
    tend = ntiles;
  } else if ((int)blockIdx.x < peer_ctas) {
    t0 = blockIdx.x;
    step = peer_ctas;
    tend = half;
    peer_only = true;
  } else {
    t0 = blockIdx.x - peer_ctas;
    step = gridDim.x - peer_ctas;
    tend = ntiles - half;
    local_only = true;
  }
  for (size_t t = t0; t < tend; t += step) {
    size_t tt;
    bool to_peer;
    if (peer_only) {
      tt = t;
      to_peer = true;
    } else if (local_only) {
      tt = half + t;
      to_peer = false;
    } else {
      to_peer = (t & 1) == 0;
      tt = to_peer ? t / 2 : half + t / 2;
      if (tt >= ntiles) continue;
    }
    uint4* dst = to_peer ? dst_peer : dst_local;
    const size_t base = tt * tile;
    const uint32_t n = (uint32_t)((nvec - base < tile) ? nvec - base : tile);
    for (uint32_t p = warp * 32 * U; p < n; p += nw * 32 * U) {
      uint4 v[U];
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (p + lane + 32 * k < n) v[k] = __ldcs(src + base + p + lane + 32 * k);
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (p + lane + 32 * k < n) __stcs(dst + base + p + lane + 32 * k, v[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------------- TMA variant
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(b)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}"
               ::"r"(smemAddr(b)), "r"(parity)
               : "memory");
}
__device__ __forceinline__ void bulkLoad(void* smem, const void* g, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smemAddr(smem)),
               "l"(g), "r"(bytes), "r"(smemAddr(b))
               : "memory");
}
__device__ __forceinline__ void bulkStore(void* g, const void* smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smemAddr(smem)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void bulkWaitRead() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// One thread per CTA drives a ring of STAGES chunks: bulk load global->shared, bulk store shared->global.
template <int STAGES> __global__ void __launch_bounds__(32) tmaCopy(const char* src, char* dst, size_t bytes, uint32_t chunk) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t full[STAGES];
  if (threadIdx.x != 0) return;
  for (int s = 0; s < STAGES; ++s) mbarInit(&full[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const size_t nchunks = (bytes + chunk - 1) / chunk;
  // chunks of this CTA: blockIdx.x, +gridDim.x, ...
  const size_t mine = (nchunks > blockIdx.x) ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto chunkOff = [&](size_t i) { return (blockIdx.x + i * gridDim.x) * (size_t)chunk; };
  auto chunkLen = [&](size_t i) {
    size_t o = chunkOff(i);
    return (uint32_t)((bytes - o < chunk) ? bytes - o : chunk);
  };
  const size_t pre = mine < (size_t)(STAGES - 1) ? mine : (size_t)(STAGES - 1);
  for (size_t i = 0; i < pre; ++i) {
    mbarExpectTx(&full[i % STAGES], chunkLen(i));
    bulkLoad(smem + (i % STAGES) * (size_t)chunk, src + chunkOff(i), chunkLen(i), &full[i % STAGES]);
  }
  for (size_t i = 0; i < mine; ++i) {
    const int s = (int)(i % STAGES);
    mbarWait(&full[s], (uint32_t)((i / STAGES) & 1));
    bulkStore(dst + chunkOff(i), smem + s * (size_t)chunk, chunkLen(i));
    // refill the slot whose store was issued one iteration ago
    const size_t nxt = i + STAGES - 1;
    if (nxt < mine) {
      bulkWaitRead<1>();
      const int ns = (int)(nxt % STAGES);
      mbarExpectTx(&full[ns], chunkLen(nxt));
      bulkLoad(smem + ns * (size_t)chunk, src + chunkOff(nxt), chunkLen(nxt), &full[ns]);
    }
  }
  bulkWaitRead<0>();
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------------ driver
struct Dev {
  int id;
  char *a, *b;
  cudaStream_t st;
  cudaEvent_t e0, e1;
};

template <typename F> double timeIt(std::vector<Dev*> devs, F launch, int reps = 5) {
  for (auto d : devs) {
    CK(cudaSetDevice(d->id));
    launch(*d);
  }
  for (auto d : devs) {
    CK(cudaSetDevice(d->id));
    CK(cudaStreamSynchronize(d->st));
  }
  for (auto d : devs) {
    CK(cudaSetDevice(d->id));
    CK(cudaEventRecord(d->e0, d->st));
    for (int r = 0; r < reps; ++r) launch(*d);
    CK(cudaEventRecord(d->e1, d->st));
  }
  double worst = 0;
  for (auto d : devs) {
    CK(cudaSetDevice(d->id));
    CK(cudaStreamSynchronize(d->st));
    float ms;
    CK(cudaEventElapsedTime(&ms, d->e0, d->e1));
    if (ms / reps > worst) worst = ms / reps;
  }
  return worst;
}

int main(int argc, char** argv) {
  size_t bytes = (argc > 1 ? atoll(argv[1]) : 4096ll) << 20; // MiB
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev > 2) ndev = 2;
  std::vector<Dev> dv(ndev);
  int sms = 0;
  for (int i = 0; i < ndev; ++i) {
    dv[i].id = i;
    CK(cudaSetDevice(i));
    CK(cudaMalloc(&dv[i].a, bytes));
    CK(cudaMalloc(&dv[i].b, bytes));
    CK(cudaMemset(dv[i].a, 1, bytes));
    CK(cudaMemset(dv[i].b, 0, bytes));
    CK(cudaStreamCreate(&dv[i].st));
    CK(cudaEventCreate(&dv[i].e0));
    CK(cudaEventCreate(&dv[i].e1));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, i));
    if (ndev == 2) CK(cudaDeviceEnablePeerAccess(1 - i, 0));
  }
  const size_t nvec = bytes / 16;
  printf("bytes per copy: %.2f GB, SMs %d, GPUs %d\n", bytes / 1e9, sms, ndev);
  printf("%-44s %10s %10s %10s\n", "variant", "local GB/s", "peer 1dir", "peer 2dir");

  auto report = [&](const char* name, auto launchTo) {
    // local: a -> b on dev 0 (read + write bytes counted, like MEASURED_PEAKS)
    double loc = timeIt({&dv[0]}, [&](Dev& d) { launchTo(d, d.a, d.b); });
    double p1 = 0, p2 = 0;
    if (ndev == 2) {
      p1 = timeIt({&dv[0]}, [&](Dev& d) { launchTo(d, d.a, dv[1 - d.id].b); });
      p2 = timeIt({&dv[0], &dv[1]}, [&](Dev& d) { launchTo(d, d.a, dv[1 - d.id].b); });
    }
    printf("%-44s %10.1f %10.1f %10.1f\n", name, 2.0 * bytes / loc / 1e6, p1 > 0 ? bytes / p1 / 1e6 : 0.0,
           p2 > 0 ? bytes / p2 / 1e6 : 0.0);
    fflush(stdout);
  };

  // PULL instead of push: the same kernels reading the PEER's buffer and writing local memory (only meaningful for the
  // two peer columns; "local" repeats the local copy). If remote loads beat remote stores, a receiver-driven transpose
  // would have a higher ceiling than the sender-driven one the engine uses.
  auto reportPull = [&](const char* name, auto launchTo) {
    if (ndev != 2) return;
    double p1 = timeIt({&dv[0]}, [&](Dev& d) { launchTo(d, dv[1 - d.id].a, d.b); });
    double p2 = timeIt({&dv[0], &dv[1]}, [&](Dev& d) { launchTo(d, dv[1 - d.id].a, d.b); });
    printf("%-44s %10s %10.1f %10.1f\n", name, "(pull)", bytes / p1 / 1e6, bytes / p2 / 1e6);
    fflush(stdout);
  };

  report("cudaMemcpyAsync (copy engine)", [&](Dev& d, char* s, char* t) {
    CK(cudaMemcpyAsync(t, s, bytes, cudaMemcpyDeviceToDevice, d.st));
  });

  for (int grid : {sms * 1, sms * 3 / 2, sms * 2, sms * 5 / 2, sms * 3, sms * 4, sms * 6, sms * 8}) {
    char nm[96];
    snprintf(nm, sizeof nm, "simt U=4 cs grid=%d", grid);
    report(nm, [&](Dev& d, char* s, char* t) { simtCopy<4, CS><<<grid, 256, 0, d.st>>>((const uint4*)s, (uint4*)t, nvec); });
  }
  for (int grid : {sms * 1, sms * 2, sms * 5 / 2, sms * 4}) {
    char nm[96];
    snprintf(nm, sizeof nm, "simt 256-bit U=2 grid=%d", grid);
    report(nm, [&](Dev& d, char* s, char* t) { simtCopy256<2><<<grid, 256, 0, d.st>>>((const V32*)s, (V32*)t, bytes / 32); });
    snprintf(nm, sizeof nm, "simt 256-bit U=4 grid=%d", grid);
    report(nm, [&](Dev& d, char* s, char* t) { simtCopy256<4><<<grid, 256, 0, d.st>>>((const V32*)s, (V32*)t, bytes / 32); });
  }
  for (int grid : {sms * 1, sms * 2, sms * 5 / 2, sms * 4, sms * 8}) {
    char nm[96];
    snprintf(nm, sizeof nm, "PULL simt U=4 cs grid=%d", grid);
    reportPull(nm, [&](Dev& d, char* s, char* t) { simtCopy<4, CS><<<grid, 256, 0, d.st>>>((const uint4*)s, (uint4*)t, nvec); });
    snprintf(nm, sizeof nm, "PULL simt U=8 cs grid=%d", grid);
    reportPull(nm, [&](Dev& d, char* s, char* t) { simtCopy<8, CS><<<grid, 256, 0, d.st>>>((const uint4*)s, (uint4*)t, nvec); });
    snprintf(nm, sizeof nm, "PULL simt 256-bit U=4 grid=%d", grid);
    reportPull(nm, [&](Dev& d, char* s, char* t) { simtCopy256<4><<<grid, 256, 0, d.st>>>((const V32*)s, (V32*)t, bytes / 32); });
  }
  for (int grid : {sms * 1, sms * 3 / 2, sms * 2, sms * 5 / 2, sms * 3}) {
    char nm[96];
    snprintf(nm, sizeof nm, "simt U=8 cs grid=%d", grid);
    report(nm, [&](Dev& d, char* s, char* t) { simtCopy<8, CS><<<grid, 256, 0, d.st>>>((const uint4*)s, (uint4*)t, nvec); });
  }
  for (int grid : {sms * 1, sms * 2, sms * 3}) {
    char nm[96];
    snprintf(nm, sizeof nm, "simt U=6 cs grid=%d", grid);
    report(nm, [&](Dev& d, char* s, char* t) { simtCopy<6, CS><<<grid, 256, 0, d.st>>>((const uint4*)s, (uint4*)t, nvec); });
  }
  for (int grid : {sms * 1, sms * 2}) {
    char nm[96];
    snprintf(nm, sizeof nm, "simt U=4 cs 512thr grid=%d", grid);
    report(nm, [&](Dev& d, char* s, char* t) { simtCopy<4, CS><<<grid, 512, 0, d.st>>>((const uint4*)s, (uint4*)t, nvec); });
  }
  if (ndev == 2) {
    // P=2 transpose shape, both GPUs at once: time for "S bytes read, S/2 local + S/2 peer"
    printf("%-44s %10s %10s %10s\n", "mixed (half local, half peer), 2 GPUs", "ms", "wire GB/s", "");
    auto mixed = [&](const char* name, int U, int grid, int mode, int peer_ctas) {
      double ms = timeIt({&dv[0], &dv[1]}, [&](Dev& d) {
        if (U == 4)
          mixedCopy<4><<<grid, 256, 0, d.st>>>((const uint4*)d.a, (uint4*)d.b, (uint4*)dv[1 - d.id].b, nvec, mode, peer_ctas);
        else
          mixedCopy<8><<<grid, 256, 0, d.st>>>((const uint4*)d.a, (uint4*)d.b, (uint4*)dv[1 - d.id].b, nvec, mode, peer_ctas);
      });
      printf("%-44s %10.3f %10.1f\n", name, ms, bytes / 2.0 / ms / 1e6);
      fflush(stdout);
    };
    for (int g : {2, 3, 4, 6}) {
      char nm[96];
      snprintf(nm, sizeof nm, "interleaved U=4 grid=%d*SM", g);
      mixed(nm, 4, sms * g, 0, 0);
    }
    mixed("interleaved U=8 grid=2*SM", 8, sms * 2, 0, 0);
    for (int g : {2, 3, 4}) {
      for (int frac : {50, 67, 75, 85}) {
        char nm[96];
        snprintf(nm, sizeof nm, "split U=4 grid=%d*SM peer CTAs %d%%", g, frac);
        mixed(nm, 4, sms * g, 1, sms * g * frac / 100);
      }
    }
  }
  report("simt U=4 plain grid=6*SM", [&](Dev& d, char* s, char* t) {
    simtCopy<4, PLAIN><<<sms * 6, 256, 0, d.st>>>((const uint4*)s, (uint4*)t, nvec);
  });
  report("simt U=4 nc/no_allocate grid=6*SM", [&](Dev& d, char* s, char* t) {
    simtCopy<4, NC><<<sms * 6, 256, 0, d.st>>>((const uint4*)s, (uint4*)t, nvec);
  });

  auto tma = [&](auto kern, int stages, uint32_t chunk, int ctas_per_sm) {
    size_t sm = (size_t)stages * chunk;
    CK(cudaSetDevice(0));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    if (ndev == 2) {
      CK(cudaSetDevice(1));
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    }
    char nm[96];
    snprintf(nm, sizeof nm, "tma bulk stages=%d chunk=%uK ctas/SM=%d", stages, chunk >> 10, ctas_per_sm);
    report(nm, [&](Dev& d, char* s, char* t) { kern<<<sms * ctas_per_sm, 32, sm, d.st>>>(s, t, bytes, chunk); });
  };
  tma(tmaCopy<4>, 4, 16384, 1);
  tma(tmaCopy<4>, 4, 16384, 2);
  tma(tmaCopy<4>, 4, 8192, 4);
  tma(tmaCopy<8>, 8, 16384, 1);
  tma(tmaCopy<8>, 8, 8192, 2);
  tma(tmaCopy<8>, 8, 4096, 4);
  tma(tmaCopy<8>, 8, 24576, 1);
  tma(tmaCopy<4>, 4, 32768, 1);
  tma(tmaCopy<4>, 4, 49152, 1);
  tma(tmaCopy<8>, 8, 2048, 8);
  for (int i = 0; i < ndev; ++i) {
    CK(cudaSetDevice(i));
    CK(cudaFuncSetAttribute(tmaCopy<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384));
  }
  for (int cps : {1, 2}) {
    char nm[96];
    snprintf(nm, sizeof nm, "PULL tma bulk stages=8 chunk=16K ctas/SM=%d", cps);
    reportPull(nm, [&](Dev& d, char* s, char* t) { tmaCopy<8><<<sms * cps, 32, 8 * 16384, d.st>>>(s, t, bytes, 16384); });
  }

  // verify the last TMA copy
  CK(cudaSetDevice(0));
  std::vector<char> h(1 << 20);
  CK(cudaMemcpy(h.data(), dv[0].b + bytes - h.size(), h.size(), cudaMemcpyDeviceToHost));
  for (char c : h)
    if (c != 1) {
      printf("VERIFY FAILED\n");
      return 1;
    }
  printf("verify ok\n");
  return 0;
}
