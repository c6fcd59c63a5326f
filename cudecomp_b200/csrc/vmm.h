// cuMem (VMM) allocations for cudecompMalloc and their cross-process mapping: the reference's CUDECOMP_ENABLE_CUMEM
// path (src/cudecomp.cc:596-660 availability checks, :1500-1570 allocation, :1640-1655 free; docs/env_vars.rst
// "CUDECOMP_ENABLE_CUMEM").
//
// The reference only ALLOCATES this way (NCCL / MPI do the sharing). Here the library maps peers' buffers itself, so
// it also carries the handle to the peer: a POSIX file descriptor, duplicated into the importing process with
// pidfd_getfd(2) (same host, same user; no socket round trip, no helper thread), or -- on systems with an IMEX
// domain -- the 64-byte fabric handle, which fits the descriptor slot CUDA IPC handles use and needs no file
// descriptor at all (multi-node NVLink groundwork; opt-in, CUDECOMP_B200_CUMEM_FABRIC=1).
// Off by default: plain cudaMalloc + CUDA IPC (peer.cc) stays the default transport.
#ifndef CUDECOMP_B200_VMM_H
#define CUDECOMP_B200_VMM_H

#include <cstddef>
#include <cstdint>

#include "bootstrap.h"

namespace cdb {

// How a described buffer is shared (BufDesc::kind).
enum : uint32_t { kShareIpc = 0, kSharePosixFd = 1, kShareFabric = 2 };

// What travels in the 64 handle bytes of a BufDesc for kSharePosixFd.
struct FdShare {
  int32_t pid;
  int32_t fd;
};

// CUDECOMP_ENABLE_CUMEM outcome of a handle (cudecompB200GetCumemState).
enum : int32_t {
  kCumemOff = 0,        // not requested
  kCumemOn = 1,         // cudecompMalloc uses cuMemCreate / cuMemMap; peers map the buffers through the handles
  kCumemNoFdPassing = 2, // requested, but the ranks cannot pass file descriptors to each other (and no fabric handles)
  kCumemNoDevice = 3,   // requested, but the device / driver has no VMM support with POSIX-fd handles (or there is no device)
};

// Host-only, collective over `comm`: can every rank duplicate a file descriptor of its neighbour with pidfd_getfd?
bool probeFdPassing(Comm& comm, uint64_t token);

// Does the current device support VMM allocations with POSIX-fd handles (and fabric handles)?
bool vmmDeviceSupported(bool* fabric_supported);

// Allocation with cuMemCreate (POSIX fd, + fabric when `want_fabric` and the platform allows it) + reserve + map +
// read/write access for the current device. `bytes` is rounded up to the allocation granularity. Throws on failure.
void* vmmAlloc(size_t bytes, bool want_fabric);

// Looks `ptr` up among this process's VMM allocations. On success: base / mapped size / the process-unique id.
bool vmmFind(const void* ptr, uint64_t* base, uint64_t* size, uint64_t* id);

// Fills the 64 handle bytes a peer needs to map allocation `base`: {pid, fd} (the fd is exported once and stays open
// until the allocation is freed) or the fabric handle. Returns the share kind, or kShareIpc (0) when nothing works.
uint32_t vmmExport(uint64_t base, bool fabric, unsigned char handle_bytes[64]);

// Unmaps, releases and frees a VMM allocation of this process; false when `ptr` is not one (the caller then cudaFree's).
bool vmmFree(void* ptr);

// Importer side: maps a peer's allocation into this process (read/write for the current device). Throws on failure.
void* vmmImport(uint32_t kind, const unsigned char handle_bytes[64], uint64_t size);
void vmmUnimport(void* base, uint64_t size);

// Duplicates file descriptor `fd` of process `pid` into this process (pidfd_open + pidfd_getfd). -1 on failure.
int duplicateFdOf(int pid, int fd);

} // namespace cdb

#endif
