// See vmm.h.
#include "vmm.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "errors.h"

#ifndef SYS_pidfd_open
#define SYS_pidfd_open 434
#endif
#ifndef SYS_pidfd_getfd
#define SYS_pidfd_getfd 438
#endif
#ifndef SYS_memfd_create
#define SYS_memfd_create 319
#endif

namespace cdb {

namespace {

// Driver entry points through the runtime: no link-time dependency on libcuda.so (the library loads without a GPU).
struct VmmApi {
  CUresult (*deviceGet)(CUdevice*, int) = nullptr;
  CUresult (*deviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*getGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*addressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*exportHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*importHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*getErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
};

template <typename F> bool fetch(const char* name, F& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    (void)cudaGetLastError();
    return false;
  }
  fn = reinterpret_cast<F>(p);
  return true;
}

const VmmApi& api() {
  static VmmApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    bool ok = true;
    ok = fetch("cuDeviceGet", a.deviceGet) && ok;
    ok = fetch("cuDeviceGetAttribute", a.deviceGetAttribute) && ok;
    ok = fetch("cuMemGetAllocationGranularity", a.getGranularity) && ok;
    ok = fetch("cuMemCreate", a.create) && ok;
    ok = fetch("cuMemRelease", a.release) && ok;
    ok = fetch("cuMemAddressReserve", a.addressReserve) && ok;
    ok = fetch("cuMemAddressFree", a.addressFree) && ok;
    ok = fetch("cuMemMap", a.map) && ok;
    ok = fetch("cuMemUnmap", a.unmap) && ok;
    ok = fetch("cuMemSetAccess", a.setAccess) && ok;
    ok = fetch("cuMemExportToShareableHandle", a.exportHandle) && ok;
    ok = fetch("cuMemImportFromShareableHandle", a.importHandle) && ok;
    ok = fetch("cuGetErrorString", a.getErrorString) && ok;
    a.ok = ok;
  });
  return a;
}

std::string drvError(CUresult r) {
  const char* s = nullptr;
  if (api().getErrorString && api().getErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  return "CUDA driver error " + std::to_string(static_cast<int>(r));
}

#define CHECK_DRV(call)                                                                                                \
  do {                                                                                                                 \
    CUresult r__ = (call);                                                                                             \
    if (r__ != CUDA_SUCCESS) THROW_CUDA_ERROR(std::string(#call) + ": " + drvError(r__));                              \
  } while (0)

bool currentDevice(CUdevice* cu_dev) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return api().ok && api().deviceGet(cu_dev, dev) == CUDA_SUCCESS;
}

struct Allocation {
  CUmemGenericAllocationHandle handle;
  uint64_t size;
  uint64_t id;
  int exported_fd = -1; // POSIX-fd handle, exported on first use, open until the allocation is freed
  bool fabric = false;  // created with CU_MEM_HANDLE_TYPE_FABRIC as well
};
std::map<uint64_t, Allocation> g_allocations; // by base address
struct Import {
  CUmemGenericAllocationHandle handle;
  uint64_t size;
};
std::map<uint64_t, Import> g_imports;         // by mapped base address
uint64_t g_next_id = 1;

// reserve + map + read/write access for the current device; undone on failure
CUdeviceptr mapHandle(CUmemGenericAllocationHandle h, size_t size, size_t granularity, CUdevice cu_dev) {
  const VmmApi& a = api();
  CUdeviceptr ptr = 0;
  CHECK_DRV(a.addressReserve(&ptr, size, granularity, 0, 0));
  CUresult r = a.map(ptr, size, 0, h, 0);
  if (r != CUDA_SUCCESS) {
    a.addressFree(ptr, size);
    THROW_CUDA_ERROR("cuMemMap: " + drvError(r));
  }
  CUmemAccessDesc access = {};
  access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  access.location.id = cu_dev;
  access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = a.setAccess(ptr, size, &access, 1);
  if (r != CUDA_SUCCESS) {
    a.unmap(ptr, size);
    a.addressFree(ptr, size);
    THROW_CUDA_ERROR("cuMemSetAccess: " + drvError(r) + " (peer-to-peer access between the ranks' GPUs is required)");
  }
  return ptr;
}

} // namespace

int duplicateFdOf(int pid, int fd) {
  const int pidfd = static_cast<int>(syscall(SYS_pidfd_open, pid, 0));
  if (pidfd < 0) return -1;
  const int got = static_cast<int>(syscall(SYS_pidfd_getfd, pidfd, fd, 0));
  const int saved = errno; // what the caller reports; close() below may overwrite it
  close(pidfd);
  errno = saved;
  return got;
}

bool probeFdPassing(Comm& comm, uint64_t token) {
  // Every rank offers a small memory file holding a value only it knows; its left neighbour must be able to read it.
  struct Offer {
    int32_t pid;
    int32_t fd;
  } mine{static_cast<int32_t>(getpid()), -1};
  const uint64_t secret = token ^ (0x9e3779b97f4a7c15ull * static_cast<uint64_t>(comm.rank() + 1));
  mine.fd = static_cast<int32_t>(syscall(SYS_memfd_create, "cudecomp_b200_fdprobe", 0u));
  if (mine.fd >= 0 && write(mine.fd, &secret, sizeof(secret)) != static_cast<ssize_t>(sizeof(secret))) {
    close(mine.fd);
    mine.fd = -1;
  }
  struct CloseOnExit { // the offer is closed on every way out, exceptions of the collectives included
    int fd;
    ~CloseOnExit() {
      if (fd >= 0) close(fd);
    }
  } offer{mine.fd};
  std::vector<Offer> all(comm.size());
  allgather(comm, &mine, sizeof(mine), all.data());
  int64_t ok = mine.fd >= 0 ? 1 : 0;
  if (comm.size() > 1) {
    const int nb = (comm.rank() + 1) % comm.size();
    const uint64_t want = token ^ (0x9e3779b97f4a7c15ull * static_cast<uint64_t>(nb + 1));
    uint64_t got = 0;
    const int fd = all[nb].fd >= 0 ? duplicateFdOf(all[nb].pid, all[nb].fd) : -1;
    if (fd < 0 || pread(fd, &got, sizeof(got), 0) != static_cast<ssize_t>(sizeof(got)) || got != want) ok = 0;
    if (fd >= 0) close(fd);
  }
  allreduceI64(comm, &ok, 1, ReduceOp::MIN); // also: nobody closes its offer before everybody has read
  return ok != 0;
}

bool vmmDeviceSupported(bool* fabric_supported) {
  if (fabric_supported) *fabric_supported = false;
  CUdevice cu_dev;
  if (!currentDevice(&cu_dev)) return false;
  int vmm = 0, posix_fd = 0, fabric = 0;
  if (api().deviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_ADDRESS_MANAGEMENT_SUPPORTED, cu_dev) != CUDA_SUCCESS ||
      api().deviceGetAttribute(&posix_fd, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, cu_dev) !=
          CUDA_SUCCESS)
    return false;
  if (api().deviceGetAttribute(&fabric, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, cu_dev) != CUDA_SUCCESS)
    fabric = 0;
  if (fabric_supported) *fabric_supported = fabric != 0;
  return vmm != 0 && posix_fd != 0;
}

void* vmmAlloc(size_t bytes, bool want_fabric) {
  const VmmApi& a = api();
  CUdevice cu_dev;
  (void)cudaFree(nullptr); // the device's primary context exists from here on
  if (!currentDevice(&cu_dev)) THROW_CUDA_ERROR("the CUDA driver's virtual memory management API is not available");
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = cu_dev;
  int rdma = 0;
  if (a.deviceGetAttribute(&rdma, CU_DEVICE_ATTRIBUTE_GPU_DIRECT_RDMA_WITH_CUDA_VMM_SUPPORTED, cu_dev) != CUDA_SUCCESS) rdma = 0;

  Allocation al{};
  size_t granularity = 0;
  CUresult r = CUDA_ERROR_NOT_SUPPORTED;
  // Most capable request first: fabric handles when asked for (a platform without an IMEX channel refuses them at
  // creation time and the allocation falls back to POSIX-fd export only, as the reference does,
  // src/cudecomp.cc:1546-1558), the GPUDirect-RDMA flag when the device reports it; then without them.
  const bool tries[3][2] = {{true, true}, {false, true}, {false, false}}; // {fabric, rdma}
  for (int attempt = 0; attempt < 3; ++attempt) {
    const bool with_fabric = tries[attempt][0], with_rdma = tries[attempt][1];
    if ((with_fabric && !want_fabric) || (with_rdma && !rdma)) continue;
    int types = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    if (with_fabric) types |= CU_MEM_HANDLE_TYPE_FABRIC;
    prop.requestedHandleTypes = static_cast<CUmemAllocationHandleType>(types);
    prop.allocFlags.gpuDirectRDMACapable = with_rdma ? 1 : 0;
    CHECK_DRV(a.getGranularity(&granularity, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    al.size = (bytes + granularity - 1) / granularity * granularity;
    r = a.create(&al.handle, al.size, &prop, 0);
    al.fabric = with_fabric;
    if (r == CUDA_SUCCESS || r == CUDA_ERROR_OUT_OF_MEMORY) break;
  }
  if (r != CUDA_SUCCESS) THROW_CUDA_ERROR("cuMemCreate: " + drvError(r));
  CUdeviceptr ptr = 0;
  try {
    ptr = mapHandle(al.handle, al.size, granularity, cu_dev);
  } catch (...) {
    a.release(al.handle);
    throw;
  }
  al.id = (1ull << 62) | g_next_id++; // never collides with the driver's buffer ids of cudaMalloc'ed memory in a PeerCache key
  g_allocations[static_cast<uint64_t>(ptr)] = al;
  return reinterpret_cast<void*>(ptr);
}

bool vmmFind(const void* ptr, uint64_t* base, uint64_t* size, uint64_t* id) {
  if (g_allocations.empty()) return false;
  const uint64_t p = reinterpret_cast<uint64_t>(ptr);
  auto it = g_allocations.upper_bound(p);
  if (it == g_allocations.begin()) return false;
  --it;
  if (p >= it->first + it->second.size) return false;
  if (base) *base = it->first;
  if (size) *size = it->second.size;
  if (id) *id = it->second.id;
  return true;
}

uint32_t vmmExport(uint64_t base, bool fabric, unsigned char handle_bytes[64]) {
  std::memset(handle_bytes, 0, 64);
  auto it = g_allocations.find(base);
  if (it == g_allocations.end()) return kShareIpc;
  Allocation& al = it->second;
  if (fabric && al.fabric) {
    CUmemFabricHandle fh;
    static_assert(sizeof(fh) == 64, "fabric handles fill the descriptor's handle bytes exactly");
    if (api().exportHandle(&fh, al.handle, CU_MEM_HANDLE_TYPE_FABRIC, 0) == CUDA_SUCCESS) {
      std::memcpy(handle_bytes, &fh, sizeof(fh));
      return kShareFabric;
    }
  }
  if (al.exported_fd < 0) {
    int fd = -1;
    if (api().exportHandle(&fd, al.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS || fd < 0)
      return kShareIpc;
    al.exported_fd = fd;
  }
  FdShare s{static_cast<int32_t>(getpid()), static_cast<int32_t>(al.exported_fd)};
  std::memcpy(handle_bytes, &s, sizeof(s));
  return kSharePosixFd;
}

bool vmmFree(void* ptr) {
  auto it = g_allocations.find(reinterpret_cast<uint64_t>(ptr));
  if (it == g_allocations.end()) return false;
  const VmmApi& a = api();
  const Allocation al = it->second;
  g_allocations.erase(it);
  if (al.exported_fd >= 0) close(al.exported_fd);
  a.unmap(static_cast<CUdeviceptr>(reinterpret_cast<uint64_t>(ptr)), al.size);
  a.release(al.handle);
  a.addressFree(static_cast<CUdeviceptr>(reinterpret_cast<uint64_t>(ptr)), al.size);
  return true;
}

void* vmmImport(uint32_t kind, const unsigned char handle_bytes[64], uint64_t size) {
  const VmmApi& a = api();
  CUdevice cu_dev;
  if (!currentDevice(&cu_dev)) THROW_CUDA_ERROR("the CUDA driver's virtual memory management API is not available");
  CUmemGenericAllocationHandle h;
  if (kind == kSharePosixFd) {
    FdShare s;
    std::memcpy(&s, handle_bytes, sizeof(s));
    const int fd = duplicateFdOf(s.pid, s.fd);
    if (fd < 0)
      THROW_INTERNAL_ERROR("cannot duplicate the memory handle of process " + std::to_string(s.pid) +
                           " (pidfd_getfd): " + std::strerror(errno));
    const CUresult r = a.importHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                      CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    close(fd);
    if (r != CUDA_SUCCESS) THROW_CUDA_ERROR("cuMemImportFromShareableHandle: " + drvError(r));
  } else if (kind == kShareFabric) {
    CUmemFabricHandle fh;
    std::memcpy(&fh, handle_bytes, sizeof(fh));
    CHECK_DRV(a.importHandle(&h, &fh, CU_MEM_HANDLE_TYPE_FABRIC));
  } else {
    THROW_INTERNAL_ERROR("unknown kind of shared allocation");
  }
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = cu_dev;
  size_t granularity = 0;
  CUdeviceptr ptr = 0;
  try {
    CHECK_DRV(a.getGranularity(&granularity, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    ptr = mapHandle(h, size, granularity, cu_dev);
  } catch (...) {
    a.release(h);
    throw;
  }
  g_imports[static_cast<uint64_t>(ptr)] = Import{h, size};
  return reinterpret_cast<void*>(ptr);
}

void vmmUnimport(void* base, uint64_t size) {
  const VmmApi& a = api();
  auto it = g_imports.find(reinterpret_cast<uint64_t>(base));
  if (it == g_imports.end()) return;
  const Import im = it->second;
  g_imports.erase(it);
  (void)size;
  a.unmap(static_cast<CUdeviceptr>(reinterpret_cast<uint64_t>(base)), im.size);
  a.release(im.handle);
  a.addressFree(static_cast<CUdeviceptr>(reinterpret_cast<uint64_t>(base)), im.size);
}

} // namespace cdb
