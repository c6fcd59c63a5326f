// See peer.h.
#include "peer.h"

#include <cuda.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "errors.h"
#include "kernels.h"
#include "vmm.h"

namespace cdb {

namespace {

// Driver entry points are fetched through the runtime so the library has no link-time dependency on
// libcuda.so and still loads on a machine without a GPU.
using PfnGetAddressRange = CUresult (*)(CUdeviceptr*, size_t*, CUdeviceptr);
using PfnPointerGetAttribute = CUresult (*)(void*, CUpointer_attribute, CUdeviceptr);

struct DriverApi {
  PfnGetAddressRange getAddressRange = nullptr;
  PfnPointerGetAttribute pointerGetAttribute = nullptr;
  bool ok = false;
};

const DriverApi& driverApi() {
  static DriverApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      api.getAddressRange = reinterpret_cast<PfnGetAddressRange>(fn);
    fn = nullptr;
    if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      api.pointerGetAttribute = reinterpret_cast<PfnPointerGetAttribute>(fn);
    (void)cudaGetLastError();
    api.ok = api.getAddressRange && api.pointerGetAttribute;
  });
  return api;
}

struct ExportEntry {
  uint64_t buffer_id;
  uint64_t size;
  cudaIpcMemHandle_t handle;
  uint32_t kind; // BufDesc::kind
  bool exportable;
  std::vector<int> described_to; // world ranks that have been shown this allocation in a CallMsg
};
std::map<uint64_t, ExportEntry> g_exports; // by allocation base

double nowSeconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

void describeBuffer(const void* ptr, BufDesc* d) {
  std::memset(d, 0, sizeof(*d));
  if (!ptr) return;
  static const bool disabled = [] {
    const char* v = std::getenv("CUDECOMP_B200_DISABLE_IPC");
    return v && *v && std::strcmp(v, "0") != 0;
  }();
  if (disabled) return;
  static const bool fabric = [] {
    const char* v = std::getenv("CUDECOMP_B200_CUMEM_FABRIC");
    return v && *v && std::strcmp(v, "0") != 0;
  }();
  uint64_t vbase = 0, vsize = 0, vid = 0;
  if (vmmFind(ptr, &vbase, &vsize, &vid)) {
    // cudecompMalloc'ed with cuMem (CUDECOMP_ENABLE_CUMEM): peers map it through its shareable handle
    auto it = g_exports.find(vbase);
    if (it == g_exports.end() || it->second.buffer_id != vid) {
      ExportEntry e{};
      e.buffer_id = vid;
      e.size = vsize;
      static_assert(sizeof(e.handle) == 64, "the descriptor's handle bytes hold a CUDA IPC handle, {pid, fd} or a fabric handle");
      e.kind = vmmExport(vbase, fabric, reinterpret_cast<unsigned char*>(&e.handle));
      e.exportable = e.kind != kShareIpc;
      g_exports[vbase] = e;
      it = g_exports.find(vbase);
    }
    const ExportEntry& e = it->second;
    d->handle = e.handle;
    d->offset = reinterpret_cast<uint64_t>(ptr) - vbase;
    d->alloc_size = e.size;
    d->buffer_id = e.buffer_id;
    d->exportable = e.exportable ? 1u : 0u;
    d->kind = e.kind;
    return;
  }
  const DriverApi& api = driverApi();
  if (!api.ok) return;

  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess || attr.type != cudaMemoryTypeDevice) {
    (void)cudaGetLastError();
    return;
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  if (api.getAddressRange(&base, &size, reinterpret_cast<CUdeviceptr>(ptr)) != CUDA_SUCCESS) return;
  unsigned long long buffer_id = 0;
  if (api.pointerGetAttribute(&buffer_id, CU_POINTER_ATTRIBUTE_BUFFER_ID, reinterpret_cast<CUdeviceptr>(ptr)) !=
      CUDA_SUCCESS)
    return;

  auto it = g_exports.find(static_cast<uint64_t>(base));
  if (it == g_exports.end() || it->second.buffer_id != buffer_id) {
    ExportEntry e{};
    e.buffer_id = buffer_id;
    e.size = size;
    // Allocations below 2 MiB may be carved out of a shared driver block: all of them would present the same IPC
    // handle, which a peer can import only once. They are treated as not exportable (staged path instead).
    e.exportable = false;
    if (size >= (size_t(2) << 20)) {
      e.exportable = (cudaIpcGetMemHandle(&e.handle, reinterpret_cast<void*>(base)) == cudaSuccess);
      if (!e.exportable) (void)cudaGetLastError();
    }
    g_exports[static_cast<uint64_t>(base)] = e;
    it = g_exports.find(static_cast<uint64_t>(base));
  }
  const ExportEntry& e = it->second;
  d->handle = e.handle;
  d->offset = reinterpret_cast<uint64_t>(ptr) - static_cast<uint64_t>(base);
  d->alloc_size = e.size;
  d->buffer_id = e.buffer_id;
  d->exportable = e.exportable ? 1u : 0u;
  d->kind = kShareIpc;
}

void noteDescribed(const BufDesc& d, const void* ptr, const std::vector<int>& group_world, int me) {
  if (!d.exportable || !ptr) return;
  auto it = g_exports.find(reinterpret_cast<uint64_t>(ptr) - d.offset);
  if (it == g_exports.end() || it->second.buffer_id != d.buffer_id) return;
  std::vector<int>& to = it->second.described_to;
  for (size_t i = 0; i < group_world.size(); ++i) {
    if (static_cast<int>(i) == me) continue;
    if (std::find(to.begin(), to.end(), group_world[i]) == to.end()) to.push_back(group_world[i]);
  }
}

std::vector<int> takeDescribedTo(const void* alloc_base) {
  std::vector<int> out;
  auto it = g_exports.find(reinterpret_cast<uint64_t>(alloc_base));
  if (it == g_exports.end()) return out;
  out = std::move(it->second.described_to);
  g_exports.erase(it);
  return out;
}

// ------------------------------------------------------------------------------------------- AckBoard

AckBoard::~AckBoard() {
  if (base_) munmap(base_, bytes_);
}

std::atomic<uint64_t>* AckBoard::cell(int reader, int owner) const {
  return reinterpret_cast<std::atomic<uint64_t>*>(base_) + static_cast<size_t>(reader) * nranks_ + owner;
}

void AckBoard::create(Comm& comm, uint64_t token) {
  nranks_ = comm.size();
  me_ = comm.rank();
  bytes_ = static_cast<size_t>(nranks_) * nranks_ * sizeof(uint64_t);
  char name[96];
  std::snprintf(name, sizeof(name), "/cudecomp_b200_%016llx_acks", static_cast<unsigned long long>(token));
  int32_t ok = 1;
  if (me_ == 0) {
    int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, static_cast<off_t>(bytes_)) != 0) ok = 0;
    if (ok) {
      base_ = mmap(nullptr, bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      if (base_ == MAP_FAILED) {
        base_ = nullptr;
        ok = 0;
      } else {
        std::memset(base_, 0, bytes_);
      }
    }
    if (fd >= 0) close(fd);
  }
  bcast(comm, &ok, sizeof(ok), 0);
  if (ok && me_ != 0) {
    int fd = shm_open(name, O_RDWR, 0600);
    if (fd >= 0) {
      base_ = mmap(nullptr, bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      if (base_ == MAP_FAILED) base_ = nullptr;
      close(fd);
    }
  }
  int64_t all_ok = (ok && base_) ? 1 : 0;
  allreduceI64(comm, &all_ok, 1, ReduceOp::MIN);
  if (me_ == 0) shm_unlink(name);
  if (!all_ok) {
    if (base_) munmap(base_, bytes_);
    base_ = nullptr;
    THROW_INTERNAL_ERROR("cannot create the shared-memory acknowledgement board (is /dev/shm available to all ranks?)");
  }
}

void AckBoard::destroy() {
  if (base_) munmap(base_, bytes_);
  base_ = nullptr;
}

void AckBoard::publish(int owner, uint64_t count) {
  if (base_) cell(me_, owner)->store(count, std::memory_order_release);
}

uint64_t AckBoard::seen(int reader, int owner) const {
  return base_ ? cell(reader, owner)->load(std::memory_order_acquire) : ~0ull;
}

// ------------------------------------------------------------------------------------------ PeerCache

PeerCache::~PeerCache() { clear(); }

void* PeerCache::resolve(int owner, const BufDesc& d) {
  if (!d.exportable) THROW_INTERNAL_ERROR("peer buffer is not exportable");
  Key k{owner, d.buffer_id, std::string(reinterpret_cast<const char*>(&d.handle), sizeof(d.handle))};
  auto it = map_.find(k);
  if (it == map_.end()) {
    evictIfNeeded();
    void* base = nullptr;
    if (d.kind != kShareIpc) {
      base = vmmImport(d.kind, reinterpret_cast<const unsigned char*>(&d.handle), d.alloc_size);
    } else {
      cudaError_t err = cudaIpcOpenMemHandle(&base, d.handle, cudaIpcMemLazyEnablePeerAccess);
      if (err != cudaSuccess) {
        (void)cudaGetLastError();
        THROW_CUDA_ERROR(std::string("cudaIpcOpenMemHandle failed for a buffer of rank ") + std::to_string(owner) +
                         ": " + cudaGetErrorString(err) +
                         " (peer-to-peer access between the ranks' GPUs is required)");
      }
    }
    it = map_.emplace(k, Entry{base, 0, d.kind, d.alloc_size}).first;
  }
  it->second.last_use = ++tick_;
  return static_cast<char*>(it->second.base) + d.offset;
}

void PeerCache::closeImport(const Entry& e) {
  if (e.kind != kShareIpc)
    vmmUnimport(e.base, e.size);
  else
    cudaIpcCloseMemHandle(e.base);
}

void PeerCache::forgetBuffer(int owner, uint64_t buffer_id) {
  for (auto it = map_.begin(); it != map_.end();) {
    if (it->first.owner == owner && it->first.buffer_id == buffer_id) {
      closeImport(it->second);
      it = map_.erase(it);
    } else {
      ++it;
    }
  }
  (void)cudaGetLastError();
}

void PeerCache::forgetOwner(int owner) {
  for (auto it = map_.begin(); it != map_.end();) {
    if (it->first.owner == owner) {
      closeImport(it->second);
      it = map_.erase(it);
    } else {
      ++it;
    }
  }
  (void)cudaGetLastError();
}

// No in-flight kernel of this process can still target a released buffer: its owner only frees it after every
// operation that used it has completed there, and the owner's kernel completes only after it has received this
// rank's "all my stores have landed" flag.
bool PeerCache::noteReleases(int owner, uint64_t release_count, const uint64_t* recent_ids) {
  uint64_t& seen = releases_seen_[owner];
  if (release_count <= seen) return false;
  const uint64_t fresh = release_count - seen;
  if (fresh > static_cast<uint64_t>(kReleaseSlots)) {
    forgetOwner(owner); // missed some announcements: drop everything, imports are re-created on demand
  } else {
    for (uint64_t k = 0; k < fresh; ++k) forgetBuffer(owner, recent_ids[k]);
  }
  seen = release_count;
  return true;
}

void PeerCache::evictIfNeeded() {
  constexpr size_t kMaxEntries = 256;
  if (map_.size() < kMaxEntries) return;
  // an import may still be the target of a running kernel: drain the device before unmapping
  cudaDeviceSynchronize();
  while (map_.size() >= kMaxEntries / 2) {
    auto victim = map_.begin();
    for (auto it = map_.begin(); it != map_.end(); ++it)
      if (it->second.last_use < victim->second.last_use) victim = it;
    closeImport(victim->second);
    map_.erase(victim);
  }
  (void)cudaGetLastError();
}

void PeerCache::clear() {
  if (map_.empty()) return;
  cudaDeviceSynchronize();
  for (auto& kv : map_) closeImport(kv.second);
  (void)cudaGetLastError();
  map_.clear();
}

// -------------------------------------------------------------------------------------------- Mailbox

Mailbox::~Mailbox() {
  if (base_) munmap(base_, bytes_);
}

Mailbox::Slot* Mailbox::slot(int rank, int channel, int parity) {
  constexpr size_t kSlotBytes = 384;
  static_assert(sizeof(Slot) <= kSlotBytes, "mailbox slot too small");
  char* p = static_cast<char*>(base_) + (static_cast<size_t>(rank) * 4 + channel * 2 + parity) * kSlotBytes;
  return reinterpret_cast<Slot*>(p);
}

void Mailbox::create(Comm& comm, uint64_t token, int instance) {
  nranks_ = comm.size();
  me_ = comm.rank();
  bytes_ = static_cast<size_t>(nranks_) * 4 * 384;
  if (const char* v = std::getenv("CUDECOMP_B200_HOST_TIMEOUT")) timeout_s_ = std::atof(v);
  char name[96];
  std::snprintf(name, sizeof(name), "/cudecomp_b200_%016llx_%d", static_cast<unsigned long long>(token), instance);
  name_ = name;

  int32_t ok = 1;
  if (me_ == 0) {
    int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, static_cast<off_t>(bytes_)) != 0) ok = 0;
    if (ok) {
      base_ = mmap(nullptr, bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      if (base_ == MAP_FAILED) {
        base_ = nullptr;
        ok = 0;
      } else {
        std::memset(base_, 0, bytes_);
      }
    }
    if (fd >= 0) close(fd);
  }
  bcast(comm, &ok, sizeof(ok), 0);
  if (ok && me_ != 0) {
    int fd = shm_open(name, O_RDWR, 0600);
    if (fd >= 0) {
      base_ = mmap(nullptr, bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      if (base_ == MAP_FAILED) base_ = nullptr;
      close(fd);
    }
  }
  int64_t all_ok = (ok && base_) ? 1 : 0;
  allreduceI64(comm, &all_ok, 1, ReduceOp::MIN);
  if (me_ == 0) shm_unlink(name); // the mappings keep the segment alive; nothing is left behind on a crash
  if (!all_ok) {
    if (base_) munmap(base_, bytes_);
    base_ = nullptr;
    THROW_INTERNAL_ERROR("cannot create the shared-memory mailbox (is /dev/shm available to all ranks of the node?)");
  }
  seq_[0] = seq_[1] = 0;
}

void Mailbox::destroy() {
  if (base_) munmap(base_, bytes_);
  base_ = nullptr;
}

void Mailbox::reset(Comm& comm) {
  if (!base_) return;
  barrier(comm);
  for (int ch = 0; ch < 2; ++ch) {
    for (int par = 0; par < 2; ++par) slot(me_, ch, par)->seq.store(0, std::memory_order_release);
    seq_[ch] = 0;
  }
  barrier(comm);
}

void Mailbox::exchange(int channel, const std::vector<int>& members, int my_index, const CallMsg& mine,
                       std::vector<CallMsg>& out) {
  if (!base_) THROW_INTERNAL_ERROR("mailbox not initialised");
  const uint64_t n = ++seq_[channel];
  const int par = static_cast<int>(n & 1);
  Slot* s = slot(me_, channel, par);
  s->msg = mine;
  s->seq.store(n, std::memory_order_release);

  out.resize(members.size());
  for (size_t i = 0; i < members.size(); ++i) {
    if (static_cast<int>(i) == my_index) {
      out[i] = mine;
      continue;
    }
    Slot* ps = slot(members[i], channel, par);
    double t0 = 0;
    uint32_t spins = 0;
    for (;;) {
      const uint64_t v = ps->seq.load(std::memory_order_acquire);
      if (v == n) break;
      if (v > n) THROW_INTERNAL_ERROR("mailbox sequence overrun: ranks issued collective calls in different orders");
      if (++spins > 2000) {
        sched_yield();
        if ((spins & 1023) == 0) {
          if (t0 == 0) t0 = nowSeconds();
          if (timeout_s_ > 0 && nowSeconds() - t0 > timeout_s_)
            THROW_INTERNAL_ERROR("timed out waiting for rank " + std::to_string(members[i]) +
                                 " to enter the same collective call");
        }
      }
    }
    out[i] = ps->msg;
    if (out[i].opcode != mine.opcode)
      THROW_INVALID_USAGE("ranks of one communicator entered different cuDecomp operations");
  }
}

// ---------------------------------------------------------------------------------------- SignalArena

SignalArena::~SignalArena() {
  // process teardown without cudecompFinalize: leave the memory to the driver
}

void SignalArena::create(Comm& comm) {
  const int n = comm.size();
  if (n > kPadMaxRanks) THROW_NOT_SUPPORTED("more ranks than the signal pad supports");
  static_assert(kPadWords * sizeof(uint64_t) <= kSlotBytes, "signal pad larger than its slot");
  const size_t bytes = kSlotBytes * kSlots;
  void* p = nullptr;
  CHECK_CUDA(cudaMalloc(&p, bytes));
  mine_ = static_cast<char*>(p);
  CHECK_CUDA(cudaMemset(mine_, 0, bytes));
  CHECK_CUDA(cudaDeviceSynchronize());
  void* eh = nullptr;
  CHECK_CUDA(cudaHostAlloc(&eh, 64, cudaHostAllocMapped));
  err_host_ = static_cast<uint32_t*>(eh);
  std::memset(eh, 0, 64);
  void* ed = nullptr;
  CHECK_CUDA(cudaHostGetDevicePointer(&ed, eh, 0));
  err_dev_ = static_cast<uint32_t*>(ed);

  bases_.assign(n, nullptr);
  imported_.assign(n, false);
  bases_[comm.rank()] = mine_;
  if (n == 1) return;

  struct Msg {
    cudaIpcMemHandle_t h;
    int32_t ok;
  } mine{};
  mine.ok = (cudaIpcGetMemHandle(&mine.h, mine_) == cudaSuccess) ? 1 : 0;
  if (!mine.ok) (void)cudaGetLastError();
  std::vector<Msg> all(n);
  allgather(comm, &mine, sizeof(Msg), all.data());
  std::string failure;
  for (int r = 0; r < n && failure.empty(); ++r) {
    if (r == comm.rank()) continue;
    if (!all[r].ok) {
      failure = "rank " + std::to_string(r) + " cannot export device memory with CUDA IPC";
      break;
    }
    void* q = nullptr;
    cudaError_t err = cudaIpcOpenMemHandle(&q, all[r].h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) {
      (void)cudaGetLastError();
      failure = std::string("cannot map the signal arena of rank ") + std::to_string(r) + ": " + cudaGetErrorString(err);
      break;
    }
    bases_[r] = static_cast<char*>(q);
    imported_[r] = true;
  }
  int64_t bad = failure.empty() ? 0 : 1;
  allreduceI64(comm, &bad, 1, ReduceOp::MAX);
  if (bad) {
    if (failure.empty()) failure = "another rank failed to map the signal arenas";
    THROW_CUDA_ERROR(failure + " (all ranks must be on GPUs of one node with peer-to-peer access)");
  }
}

void SignalArena::zeroSlot(int slot) {
  if (!mine_) return;
  cudaMemset(mine_ + static_cast<size_t>(slot) * kSlotBytes, 0, kSlotBytes);
  cudaDeviceSynchronize();
  (void)cudaGetLastError();
}

void SignalArena::destroy(Comm* comm) {
  if (!mine_) return;
  cudaDeviceSynchronize();
  for (size_t r = 0; r < bases_.size(); ++r)
    if (imported_[r] && bases_[r]) cudaIpcCloseMemHandle(bases_[r]);
  bases_.clear();
  imported_.clear();
  if (comm && comm->size() > 1) barrier(*comm);
  cudaFree(mine_);
  mine_ = nullptr;
  if (err_host_) cudaFreeHost(err_host_);
  err_host_ = nullptr;
  err_dev_ = nullptr;
  (void)cudaGetLastError();
}

} // namespace cdb
