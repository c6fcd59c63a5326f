// Box-copy kernels for sm_100a. See kernels.h.
//
// Data path per box: coalesced 16-byte loads from the local pencil (LDG.128, streaming), 16-byte stores
// to the destination (STG.128) which is either local HBM or a peer GPU's memory mapped over NVLink
// (32-byte LDG/STG.256 in launches that store into peers: +3 % on the wire, measured).
// The kernels are persistent: a fixed grid of CTAs walks the tile list, and consecutive tiles belong
// to different peers so every peer's ingress is fed evenly for the whole duration of the launch.
// Kernels: rowCopyKernel (contiguous rows on both sides), transposeVecKernel / transposeKernel (memory
// orders differ: 16-byte micro-tile transpose through a swizzled shared tile / element-wise fallback),
// rowCopyPhasedKernel (the fused staged schedule of in-place calls: push chunk s, unpack chunk s - lag,
// per-chunk flags), rowCopyBulkKernel (TMA variant, opt-in).
//
// Cross-GPU ordering lives in the same launch (no host synchronisation, no second kernel):
//   entry : CTA 0 stores the operation's epoch into slot [me] of each peer's signal pad; every CTA then
//           waits until the local pad shows that epoch for each peer. A peer's kernel only starts after
//           that peer's earlier stream work, so from here on its buffers may be overwritten.
//   exit  : each CTA fences its stores (one thread, after a CTA barrier: the fence is cumulative) and bumps a
//           local counter; the last CTA publishes "done" to every peer and waits for every peer's "done".
//           Kernel completion therefore means both "my data has landed everywhere" and "everyone's data
//           has landed here".
//   step  : (phased kernel) after its share of chunk s a CTA fences and bumps counter[s]; the last CTA
//           publishes (epoch, s) to every peer with red.max; unpack tiles of chunk s wait for every peer's
//           (epoch, s) and for counter[s] == grid. No exit handshake: the last chunk's waits subsume it.
// A wait that times out records an error for the host and the launch leaves WITHOUT moving (more) data.
#include "kernels.h"

#include <atomic>
#include <cstring>

#include "tiling.h"

namespace cdb {

namespace {

std::atomic<uint64_t> g_launches{0};

__device__ __forceinline__ uint64_t ldAcquireSys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void stReleaseSys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Monotonic publish: a later step's flag may overtake an earlier one on its way to the peer (they are written by
// different CTAs), so the slot takes the maximum instead of the last value.
__device__ __forceinline__ void redMaxReleaseSys(uint64_t* p, uint64_t v) {
  asm volatile("red.release.sys.global.max.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint64_t globalTimerNs() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Spin until *flag >= want. Gives up after timeout_ns and records the failure for the host: a peer that
// never arrives (crashed rank, mismatched call sequence) must not wedge the GPU. Returns false on timeout; callers
// then skip their data movement (the buffers of a peer that never confirmed must not be touched).
__device__ __noinline__ bool waitFlagValue(const uint64_t* flag, uint64_t want, const SyncParams& s, uint32_t code) {
  if (ldAcquireSys(flag) >= want) return true;
  const uint64_t t0 = globalTimerNs();
  uint32_t spins = 0;
  while (ldAcquireSys(flag) < want) {
    ++spins;
    if (spins > 32) __nanosleep(spins > 4096 ? 1000 : 50);
    if ((spins & 255u) == 0 && s.timeout_ns != 0 && globalTimerNs() - t0 > s.timeout_ns) {
      if (s.error_word) {
        *reinterpret_cast<volatile uint32_t*>(s.error_word) = code;
        __threadfence_system();
      }
      return false;
    }
  }
  return true;
}

__device__ __forceinline__ bool waitFlag(const uint64_t* flag, const SyncParams& s, uint32_t code) {
  return waitFlagValue(flag, s.epoch, s, code);
}

// Returns false (for the whole CTA) when a peer did not arrive in time.
__device__ __forceinline__ bool syncEntry(const SyncParams& s) {
  if (s.my_pad == nullptr || s.npeers == 0 || !s.do_entry) return true;
  const int t = threadIdx.x;
  int failed = 0;
  if (t < s.npeers) {
    if (blockIdx.x == 0) stReleaseSys(s.peer_pad[t] + kPadEntry + s.my_world, s.epoch);
    failed = waitFlag(s.my_pad + kPadEntry + s.peer_world[t], s, 1u) ? 0 : 1;
  }
  return __syncthreads_or(failed) == 0;
}

__device__ __forceinline__ void syncExit(const SyncParams& s) {
  if (s.my_pad == nullptr || s.npeers == 0 || !s.do_exit) return;
  __shared__ uint32_t is_last;
  __syncthreads(); // every warp's stores (local and peer) are issued ...
  if (threadIdx.x == 0) {
    __threadfence_system(); // ... and performed before the count below (the barrier makes the fence cover the whole CTA)
    unsigned long long* ctr = reinterpret_cast<unsigned long long*>(s.my_pad + kPadCounter);
    const unsigned long long old = atomicAdd(ctr, 1ull);
    is_last = (old == static_cast<unsigned long long>(gridDim.x) - 1ull) ? 1u : 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence_system(); // order the other CTAs' (already fenced) stores before the flags below
  __syncthreads();        // ... for every flag-writing thread: thread 0 observed the counter, its fence precedes all flags
  const int t = threadIdx.x;
  if (t < s.npeers) {
    stReleaseSys(s.peer_pad[t] + kPadExit + s.my_world, s.epoch);
    waitFlag(s.my_pad + kPadExit + s.peer_world[t], s, 2u);
  }
  if (t == 0) *reinterpret_cast<volatile unsigned long long*>(s.my_pad + kPadCounter) = 0ull;
}

} // namespace

template <typename V> __device__ __forceinline__ V loadStream(const V* p) { return __ldcs(p); }
template <typename V> __device__ __forceinline__ void storeStream(V* p, const V& v) { __stcs(p, v); }

// 32-byte vectors: sm_100 has 256-bit global loads and stores (SASS LDG.E.EF.ENL2.256 / STG.E.EF.ENL2.256), one warp
// instruction then covers 1 KiB. Opt-in (kernel variant 2) until measured against the 128-bit default.
struct alignas(32) Vec32 {
  uint64_t a, b, c, d;
};
template <> __device__ __forceinline__ Vec32 loadStream<Vec32>(const Vec32* p) {
  Vec32 v;
  asm volatile("ld.global.cs.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
  return v;
}
template <> __device__ __forceinline__ void storeStream<Vec32>(Vec32* p, const Vec32& v) {
  asm volatile("st.global.cs.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v.a), "l"(v.b), "l"(v.c), "l"(v.d) : "memory");
}

// ---------------------------------------------------------------------------------------------
// ROWCOPY: every box is a grid of rows that are contiguous on both sides. V is the widest vector all
// addresses and strides of the launch are aligned to (16, 8 or 4 bytes).
// A tile is rows_per_tile rows x seg_vecs vectors (about 32 KiB); inside a tile each warp takes
// 128-vector pieces (4 independent loads per lane in flight, then 4 stores).
// ---------------------------------------------------------------------------------------------
// One tile of a ROWCOPY box, executed by the whole CTA.
template <typename V> __device__ __forceinline__ void copyRowTile(const KBox& bx, uint32_t j, int64_t esz) {
  constexpr uint32_t kUnroll = 4;
  constexpr uint32_t kPiece = 32 * kUnroll;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t nwarps = blockDim.x >> 5;
  const RowTile rt = decodeRowTile(bx, j);
  const int64_t row0 = rt.row0;
  const uint32_t c0 = rt.c0;
  const uint32_t nvec = rt.nvec;
  const uint32_t rows_here = rt.rows_here;
  const uint32_t pieces_per_row = (nvec + kPiece - 1) / kPiece;
  const uint32_t npieces = rows_here * pieces_per_row;

  if (bx.row_vecs <= 32u) {
    // Short rows (e.g. the 2-element faces of a halo along the contiguous axis): a warp per row would leave most
    // lanes idle, so lanes take (row, vector) pairs of the tile instead.
    const uint32_t total_vecs = rows_here * nvec;
    for (uint32_t e0 = threadIdx.x; e0 < total_vecs; e0 += blockDim.x * kUnroll) {
      V v[kUnroll];
      V* dptr[kUnroll];
#pragma unroll
      for (uint32_t k = 0; k < kUnroll; ++k) {
        const uint32_t e = e0 + k * blockDim.x;
        dptr[k] = nullptr;
        if (e < total_vecs) {
          const uint32_t r = e / nvec;
          const uint32_t c = e - r * nvec;
          int64_t so, dof;
          rowOffsets(bx, row0 + r, esz, so, dof);
          v[k] = loadStream(reinterpret_cast<const V*>(bx.src + so) + c0 + c);
          dptr[k] = reinterpret_cast<V*>(bx.dst + dof) + c0 + c;
        }
      }
#pragma unroll
      for (uint32_t k = 0; k < kUnroll; ++k)
        if (dptr[k]) storeStream(dptr[k], v[k]);
    }
    return;
  }

  for (uint32_t pc = warp; pc < npieces; pc += nwarps) {
    const uint32_t r = pc / pieces_per_row;
    const uint32_t q = pc - r * pieces_per_row;
    int64_t so, dof;
    rowOffsets(bx, row0 + r, esz, so, dof);
    const V* s = reinterpret_cast<const V*>(bx.src + so) + c0;
    V* d = reinterpret_cast<V*>(bx.dst + dof) + c0;
    const uint32_t base = q * kPiece + lane;
    V v[kUnroll];
#pragma unroll
    for (uint32_t k = 0; k < kUnroll; ++k) {
      const uint32_t idx = base + 32u * k;
      if (idx < nvec) v[k] = loadStream(s + idx);
    }
#pragma unroll
    for (uint32_t k = 0; k < kUnroll; ++k) {
      const uint32_t idx = base + 32u * k;
      if (idx < nvec) storeStream(d + idx, v[k]);
    }
  }
}

// kOrder: slot order (CopyParams::peer_order), a template parameter so that the default order keeps its register budget.
template <typename V, int kOrder> __global__ void __launch_bounds__(256) rowCopyKernel(const __grid_constant__ CopyParams p) {
  const bool go = syncEntry(p.sync);
  const uint32_t total = go ? p.nboxes * p.max_tiles : 0u; // a peer that never arrived: move nothing, leave through the exit

  for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
    uint32_t b, j;
    slotToBoxTile(t, p.nboxes, p.max_tiles, static_cast<uint32_t>(kOrder), b, j);
    const KBox& bx = p.box[b];
    if (j >= bx.tiles) continue;
    copyRowTile<V>(bx, j, p.elem_size);
  }

  if (go) syncExit(p.sync);
}

// ---------------------------------------------------------------------------------------------
// PHASED ROWCOPY: the fused staged / in-place schedule in one launch (kernels.h PhasedParams).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t stepKey(uint64_t epoch, uint32_t step) { return (epoch << 8) | (step + 1u); }

template <typename V> __global__ void __launch_bounds__(256) rowCopyPhasedKernel(const __grid_constant__ PhasedParams p) {
  const SyncParams& sy = p.sync;
  if (!syncEntry(sy)) return;
  unsigned long long* counters = reinterpret_cast<unsigned long long*>(sy.my_pad + kPadPhaseCounter);
  int32_t known = -1; // highest step known to be exchanged by everybody (uniform over the CTA)

  for (uint32_t s = 0; s < p.nphases; ++s) {
    const PhaseDesc ph = p.phases[s];
    const uint32_t total = ph.nsegs * ph.seg_tiles;
    for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
      const SegDesc sg = p.segs[ph.first_seg + t % ph.nsegs];
      const uint32_t jj = t / ph.nsegs;
      if (jj >= sg.count) continue;
      const KBox& bx = p.boxes[sg.box];
      const uint32_t j = sg.first_tile + jj;
      const int32_t need = static_cast<int32_t>(sg.wait) - 1;
      if (need > known) {
        // every peer's pushes of step `need` have landed here, and every local CTA has finished reading that chunk
        int failed = 0;
        const int tid = threadIdx.x;
        if (tid < sy.npeers) {
          failed = waitFlagValue(sy.my_pad + kPadStep + sy.peer_world[tid], stepKey(sy.epoch, need), sy, 4u) ? 0 : 1;
        } else if (tid == sy.npeers) {
          failed = waitFlagValue(reinterpret_cast<const uint64_t*>(counters + need), gridDim.x, sy, 5u) ? 0 : 1;
        }
        if (__syncthreads_or(failed)) return;
        known = need;
      }
      copyRowTile<V>(bx, j, p.elem_size);
    }
    if (ph.publish) {
      // My share of this step's pushes is issued. One thread fences for the CTA (the barrier orders every warp's stores
      // before it, the fence is cumulative), counts the CTA, and the last CTA of the grid tells the peers that the
      // chunk has landed; the other warps go straight on to the next phase.
      __syncthreads();
      if (threadIdx.x == 0) {
        const uint32_t step = ph.publish - 1u;
        __threadfence_system();
        const unsigned long long old = atomicAdd(counters + step, 1ull);
        if (old == static_cast<unsigned long long>(gridDim.x) - 1ull) {
          __threadfence_system();
          for (int i = 0; i < sy.npeers; ++i) redMaxReleaseSys(sy.peer_pad[i] + kPadStep + sy.my_world, stepKey(sy.epoch, step));
        }
      }
    }
  }

  // leave the phase counters zeroed for the next phased launch: the last CTA out resets them
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long* ctr = reinterpret_cast<unsigned long long*>(sy.my_pad + kPadCounter);
    const unsigned long long old = atomicAdd(ctr, 1ull);
    if (old == static_cast<unsigned long long>(gridDim.x) - 1ull) {
      for (uint32_t s = 0; s < p.nsteps; ++s) counters[s] = 0ull;
      *reinterpret_cast<volatile unsigned long long*>(ctr) = 0ull;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TRANSPOSE: source is contiguous along axis 0, destination along axis 1 (memory orders differ, e.g.
// axis-contiguous layouts). 32x32 element tiles through shared memory: coalesced on both sides.
// T is the element itself (4, 8 or 16 bytes).
// ---------------------------------------------------------------------------------------------
template <typename T, int kOrder> __global__ void __launch_bounds__(256) transposeKernel(const __grid_constant__ CopyParams p) {
  __shared__ T tile[32][33];
  const bool go = syncEntry(p.sync);

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t wrow = threadIdx.x >> 5; // 0..7
  const uint32_t total = go ? p.nboxes * p.max_tiles : 0u;

  for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
    uint32_t b, j;
    slotToBoxTile(t, p.nboxes, p.max_tiles, static_cast<uint32_t>(kOrder), b, j);
    const KBox& bx = p.box[b];
    if (j < bx.tiles) { // uniform per CTA
      const TransTile tt = decodeTransposeTile(bx, j);
      const uint32_t j0 = tt.j0;
      const uint32_t j1 = tt.j1;
      const int64_t i2 = tt.i2;
      const T* s = reinterpret_cast<const T*>(bx.src) + i2 * bx.ss[2];
      T* d = reinterpret_cast<T*>(bx.dst) + i2 * bx.ds[2];
      {
        const int64_t i0 = static_cast<int64_t>(j0) * 32 + lane;
#pragma unroll
        for (uint32_t r = 0; r < 32; r += 8) {
          const int64_t i1 = static_cast<int64_t>(j1) * 32 + r + wrow;
          if (i0 < bx.n[0] && i1 < bx.n[1]) tile[r + wrow][lane] = s[i0 * bx.ss[0] + i1 * bx.ss[1]];
        }
      }
      __syncthreads();
      {
        const int64_t i1 = static_cast<int64_t>(j1) * 32 + lane;
#pragma unroll
        for (uint32_t r = 0; r < 32; r += 8) {
          const int64_t i0 = static_cast<int64_t>(j0) * 32 + r + wrow;
          if (i0 < bx.n[0] && i1 < bx.n[1]) d[i0 * bx.ds[0] + i1 * bx.ds[1]] = tile[lane][r + wrow];
        }
      }
      __syncthreads();
    }
  }

  if (go) syncExit(p.sync);
}

// ---------------------------------------------------------------------------------------------
// TRANSPOSE_VEC: same boxes as TRANSPOSE, 16-byte accesses on both sides (kernels.h, tiling.h TransVecGeom).
// ---------------------------------------------------------------------------------------------
template <typename T, int kOrder, int kAlt = 0>
__global__ void __launch_bounds__(256) transposeVecKernel(const __grid_constant__ CopyParams p) {
  using G = TransVecGeom<sizeof(T), kAlt>;
  constexpr int VEC = G::kVec;
  union Vec {
    uint4 u;
    T t[VEC];
  };
  __shared__ uint4 tile[1024];
  const bool go = syncEntry(p.sync);
  const uint32_t total = go ? p.nboxes * p.max_tiles : 0u;
  const uint32_t tid = threadIdx.x;

  for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
    uint32_t b, j;
    slotToBoxTile(t, p.nboxes, p.max_tiles, static_cast<uint32_t>(kOrder), b, j);
    const KBox& bx = p.box[b];
    if (j < bx.tiles) { // uniform per CTA
      const TransTile tt = decodeTransposeTile(bx, j);
      const int64_t base0 = static_cast<int64_t>(tt.j0) * G::kE0;
      const int64_t base1 = static_cast<int64_t>(tt.j1) * G::kE1;
      const T* s = reinterpret_cast<const T*>(bx.src) + tt.i2 * bx.ss[2];
      T* d = reinterpret_cast<T*>(bx.dst) + tt.i2 * bx.ds[2];
      Vec in[G::kNM][VEC];
#pragma unroll
      for (int q = 0; q < G::kNM; ++q) {
        const uint32_t m = tid + 256u * q;
        const uint32_t c = m % G::kC0, g = m / G::kC0;
        const int64_t i0 = base0 + c * VEC;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const int64_t i1 = base1 + g * VEC + k;
          in[q][k].u = make_uint4(0u, 0u, 0u, 0u);
          if (i0 < bx.n[0] && i1 < bx.n[1]) in[q][k].u = __ldcs(reinterpret_cast<const uint4*>(s + i0 + i1 * bx.ss[1]));
        }
      }
#pragma unroll
      for (int q = 0; q < G::kNM; ++q) {
        const uint32_t m = tid + 256u * q;
        const uint32_t c = m % G::kC0, g = m / G::kC0;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          Vec out;
#pragma unroll
          for (int k = 0; k < VEC; ++k) out.t[k] = in[q][k].t[e];
          tile[transVecSlot(c * VEC + e, g, VEC, G::kG1)] = out.u;
        }
      }
      __syncthreads();
#pragma unroll
      for (int pass = 0; pass < VEC * G::kNM; ++pass) {
        const uint32_t n = tid + 256u * pass;
        const uint32_t v = n % G::kG1, row = n / G::kG1;
        const int64_t i0 = base0 + row;
        const int64_t i1 = base1 + v * VEC;
        if (i0 < bx.n[0] && i1 < bx.n[1])
          __stcs(reinterpret_cast<uint4*>(d + i0 * bx.ds[0] + i1), tile[transVecSlot(row, v, VEC, G::kG1)]);
      }
      __syncthreads();
    }
  }

  if (go) syncExit(p.sync);
}

// ---------------------------------------------------------------------------------------------
// ROWCOPY_BULK: TMA-driven variant of the row copy (cp.async.bulk, SASS UBLKCP). One thread per CTA keeps a ring of
// kBulkStages row segments in flight: bulk load global->shared completes on an mbarrier, bulk store shared->global
// (local HBM or the peer over NVLink) is tracked by bulk groups. Measured on B200 (profiles/r1_microbench_copy_2gpu.txt)
// it moves data exactly as fast as the LDG/STG kernel on both paths while occupying one warp per SM; it needs 16-byte
// alignment everywhere, so the SIMT kernel stays the default and the fallback.
// ---------------------------------------------------------------------------------------------
namespace bulk {
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbarInit(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(b)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbarTryWait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok)
               : "r"(smemAddr(b)), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void load(void* smem, const void* g, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(smem)),
               "l"(g), "r"(bytes), "r"(smemAddr(b))
               : "memory");
}
__device__ __forceinline__ void store(void* g, const void* smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smemAddr(smem)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
} // namespace bulk

__global__ void __launch_bounds__(32) rowCopyBulkKernel(const __grid_constant__ CopyParams p) {
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  __shared__ uint64_t full[kBulkStages];
  // destination and size of what each stage holds (only the driving thread touches them; shared memory instead of a
  // dynamically indexed local array)
  __shared__ char* dst_of[kBulkStages];
  __shared__ uint32_t bytes_of[kBulkStages];
  if (!syncEntry(p.sync)) return;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kBulkStages; ++s) bulk::mbarInit(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t total = p.nboxes * p.max_tiles;
    const int64_t esz = p.elem_size;

    // decode launch tile t -> (source, destination, bytes); false when the slot is empty for that box
    auto decode = [&](uint32_t t, const char*& src, char*& dst, uint32_t& bytes) -> bool {
      uint32_t b, j;
      slotToBoxTile(t, p.nboxes, p.max_tiles, p.peer_order, b, j);
      const KBox& bx = p.box[b];
      if (j >= bx.tiles) return false;
      const RowTile rt = decodeRowTile(bx, j); // rows_per_tile == 1: one row segment
      int64_t so, dof;
      rowOffsets(bx, rt.row0, esz, so, dof);
      src = bx.src + so + static_cast<int64_t>(rt.c0) * 16;
      dst = bx.dst + dof + static_cast<int64_t>(rt.c0) * 16;
      bytes = rt.nvec * 16u;
      return true;
    };

    uint32_t t_load = blockIdx.x;   // next tile to look at for loading
    uint32_t n_loaded = 0, n_stored = 0;
    auto issueLoad = [&]() -> bool {
      while (t_load < total) {
        const char* src;
        char* dst;
        uint32_t bytes;
        const bool ok = decode(t_load, src, dst, bytes);
        t_load += gridDim.x;
        if (!ok) continue;
        const int s = static_cast<int>(n_loaded % kBulkStages);
        dst_of[s] = dst;
        bytes_of[s] = bytes;
        bulk::mbarExpectTx(&full[s], bytes);
        bulk::load(bulk_smem + static_cast<size_t>(s) * kBulkChunkBytes, src, bytes, &full[s]);
        ++n_loaded;
        return true;
      }
      return false;
    };
    for (int s = 0; s < kBulkStages - 1; ++s)
      if (!issueLoad()) break;
    while (n_stored < n_loaded) {
      const int s = static_cast<int>(n_stored % kBulkStages);
      const uint32_t parity = (n_stored / kBulkStages) & 1u;
      // a bulk load that never completes (bad address, driver fault) must not spin forever: give up after 20 s,
      // flag the launch and leave
      if (!bulk::mbarTryWait(&full[s], parity)) {
        const uint64_t t0 = globalTimerNs();
        bool ok = false;
        while (!(ok = bulk::mbarTryWait(&full[s], parity)))
          if (globalTimerNs() - t0 > 20000000000ull) break;
        if (!ok) {
          if (p.sync.error_word) {
            *reinterpret_cast<volatile uint32_t*>(p.sync.error_word) = 3u;
            __threadfence_system();
          }
          break;
        }
      }
      bulk::store(dst_of[s], bulk_smem + static_cast<size_t>(s) * kBulkChunkBytes, bytes_of[s]);
      ++n_stored;
      // the slot stored one iteration ago may be refilled once its store has finished reading shared memory
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      issueLoad();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // all stores performed
    asm volatile("fence.proxy.async;" ::: "memory");            // ... and ordered before the generic-proxy flags below
  }
  __syncwarp();
  syncExit(p.sync);
}

namespace {

using KernelFn = void (*)(const CopyParams);

template <int kOrder> KernelFn pickKernelOrdered(KernelKind kind, int size, uint32_t geometry) {
  if (kind == KernelKind::ROWCOPY) {
    switch (size) {
    case 32: return rowCopyKernel<Vec32, kOrder>;
    case 16: return rowCopyKernel<uint4, kOrder>;
    case 8: return rowCopyKernel<uint2, kOrder>;
    case 4: return rowCopyKernel<uint32_t, kOrder>;
    }
  } else if (kind == KernelKind::TRANSPOSE) {
    switch (size) {
    case 16: return transposeKernel<uint4, kOrder>;
    case 8: return transposeKernel<uint2, kOrder>;
    case 4: return transposeKernel<uint32_t, kOrder>;
    }
  } else if (kind == KernelKind::TRANSPOSE_VEC) {
    switch (size) {
    case 16: return transposeVecKernel<uint4, kOrder>;
    case 8: return geometry ? transposeVecKernel<uint2, kOrder, 1> : transposeVecKernel<uint2, kOrder, 0>;
    case 4: return transposeVecKernel<uint32_t, kOrder>;
    }
  } else if (size == 16) {
    return rowCopyBulkKernel;
  }
  return nullptr;
}

KernelFn pickKernel(KernelKind kind, int size, uint32_t peer_order = 0, uint32_t geometry = 0) {
  return peer_order ? pickKernelOrdered<1>(kind, size, geometry) : pickKernelOrdered<0>(kind, size, geometry);
}

} // namespace

int maxResidentCtas(KernelKind kind, int size, int threads, uint32_t peer_order) {
  static int cache[2][3][4] = {};
  static int cache_dev = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev != cache_dev) {
    std::memset(cache, 0, sizeof(cache));
    cache_dev = dev;
  }
  if (kind == KernelKind::ROWCOPY_BULK) return 0; // not used: launchBulk sizes its own grid
  const int oi = peer_order ? 1 : 0;
  const int ki = (kind == KernelKind::ROWCOPY) ? 0 : (kind == KernelKind::TRANSPOSE ? 1 : 2);
  const int si = (size == 32) ? 3 : (size == 16) ? 2 : (size == 8 ? 1 : 0);
  if (threads == 256 && cache[oi][ki][si] > 0) return cache[oi][ki][si];
  KernelFn fn = pickKernel(kind, size, peer_order);
  int per_sm = 0, sms = 0;
  if (!fn || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, 0) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  const int total = per_sm * sms;
  if (threads == 256) cache[oi][ki][si] = total;
  return total;
}

static cudaError_t launchBulk(const CopyParams& p, const LaunchConfig& cfg, cudaStream_t stream) {
  const int smem = kBulkStages * static_cast<int>(kBulkChunkBytes);
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  // the opt-in for more than 48 KiB of dynamic shared memory is per device (a process may drive several)
  static std::atomic<uint64_t> configured_devices{0};
  if (dev >= 64 || !(configured_devices.load(std::memory_order_relaxed) & (1ull << dev))) {
    cudaError_t e = cudaFuncSetAttribute(rowCopyBulkKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    if (dev < 64) configured_devices.fetch_or(1ull << dev, std::memory_order_relaxed);
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const uint64_t total = static_cast<uint64_t>(p.nboxes) * p.max_tiles;
  // one TMA-driving CTA per SM by default (3 fit by shared memory)
  const int grid = chooseGrid(cfg.grid, sms, 3 * sms, total, cfg.balance);
  rowCopyBulkKernel<<<grid, 32, smem, stream>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

cudaError_t launchCopy(KernelKind kind, const CopyParams& p, const LaunchConfig& cfg, cudaStream_t stream) {
  if (kind == KernelKind::ROWCOPY_BULK) return launchBulk(p, cfg, stream);
  const int size = (kind == KernelKind::ROWCOPY) ? static_cast<int>(p.vec_size) : static_cast<int>(p.elem_size);
  KernelFn fn = pickKernel(kind, size, p.peer_order, p.geometry);
  if (!fn) return cudaErrorInvalidValue;
  const uint64_t total = static_cast<uint64_t>(p.nboxes) * p.max_tiles;
  const int resident = maxResidentCtas(kind, size, cfg.threads, p.peer_order);
  if (resident <= 0) return cudaErrorInvalidDevice;
  // Measured on B200 (profiles/r1_n1_cta_sweep.txt): HBM streams best with a moderate number of CTAs in flight;
  // filling every resident slot costs 6-8 % of copy bandwidth. 2.5 CTAs/SM for the row copy, 4/SM for the
  // shared-memory transpose. NVLink-bound launches are insensitive to the count (profiles/r1_n2_sweep.txt).
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int dflt = (kind == KernelKind::ROWCOPY) ? (5 * sms) / 2 : 4 * sms;
  const int grid = chooseGrid(cfg.grid, dflt, resident, total, cfg.balance);
  fn<<<grid, cfg.threads, 0, stream>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

cudaError_t launchPhased(const PhasedParams& p, uint64_t total_slots, const LaunchConfig& cfg, cudaStream_t stream) {
  using Fn = void (*)(const PhasedParams);
  Fn fn = nullptr;
  switch (p.vec_size) {
  case 32: fn = rowCopyPhasedKernel<Vec32>; break;
  case 16: fn = rowCopyPhasedKernel<uint4>; break;
  case 8: fn = rowCopyPhasedKernel<uint2>; break;
  case 4: fn = rowCopyPhasedKernel<uint32_t>; break;
  }
  if (!fn) return cudaErrorInvalidValue;
  int sms = 0, dev = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, cfg.threads, 0) != cudaSuccess || per_sm <= 0)
    return cudaErrorInvalidDevice;
  // CTAs wait for each other inside the launch: every one of them must be resident
  const int grid = chooseGrid(cfg.grid, (5 * sms) / 2, per_sm * sms, total_slots, 0);
  fn<<<grid, cfg.threads, 0, stream>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

uint64_t launchCount() { return g_launches.load(std::memory_order_relaxed); }

} // namespace cdb
