// Cross-process peer memory: how a rank obtains a pointer it can store through into another rank's
// buffer. Ranks are separate processes on one NVSwitch domain.
//
// Replaces the reference's transport setup (NCCL communicators src/cudecomp.cc:59-72,1152-1182; NVSHMEM
// symmetric heap :1470-1496; cuMem fabric handles :1508-1602).  Design:
//   * every buffer a peer must write (workspace, and the user's output / halo buffer when it is plain
//     device memory) is exported lazily with CUDA IPC and imported once per peer (PeerCache); cudecompMalloc'ed
//     buffers of the CUDECOMP_ENABLE_CUMEM path travel as POSIX file descriptors or fabric handles instead (vmm.h);
//   * per collective call the members of the row/column communicator swap one small descriptor through a
//     shared-memory mailbox (Mailbox) -- which buffer + offset each rank is using in THIS call and
//     whether it could be exported -- so all members take the same direct/staged decision without a
//     socket round trip;
//   * a 4 KiB signal pad per grid descriptor carries the device-side entry/exit flags (SignalPads).
#ifndef CUDECOMP_B200_PEER_H
#define CUDECOMP_B200_PEER_H

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "bootstrap.h"

namespace cdb {

// What a rank tells its peers about one buffer argument of the current call.
struct BufDesc {
  cudaIpcMemHandle_t handle; // of the allocation containing the pointer (kind 0); {pid, fd} or the fabric handle (vmm.h)
  uint64_t offset;           // pointer - allocation base
  uint64_t alloc_size;
  uint64_t buffer_id;        // CU_POINTER_ATTRIBUTE_BUFFER_ID: unique per allocation, guards handle reuse
  uint32_t exportable;       // 0: cannot be mapped by peers (managed, host, pool memory, ...)
  uint32_t kind;             // how peers map it: 0 CUDA IPC, 1 cuMem POSIX fd, 2 cuMem fabric handle (vmm.h)
};

constexpr int kReleaseSlots = 6;

struct CallMsg {
  uint32_t opcode; // must match on all members (catches mismatched collective calls)
  uint32_t flags;
  BufDesc data;    // transpose: output buffer; halo: the pencil buffer
  BufDesc work;    // workspace
  BufDesc src;     // transpose: input buffer (only described when the receiver-driven mode is on)
  // cudecompFree is not collective in practice (the reference's own tests free workspaces on a subset of the
  // ranks), so releases are announced here: how many buffers this rank has freed so far and the ids of the most
  // recent ones. Readers drop their imports of those buffers (all imports of the rank if they missed some).
  uint64_t release_count;
  uint64_t released[kReleaseSlots];
};

// Fills `d` for a local device pointer. Never throws: failure means exportable = 0.
void describeBuffer(const void* ptr, BufDesc* d);

// Bookkeeping for cudecompFree: an exported allocation must outlive every peer's import of it (freeing memory that
// another process still has open with cudaIpcOpenMemHandle is undefined). noteDescribed records which ranks were shown a
// buffer (the only ones that can have imported it); takeDescribedTo hands that set over when the buffer is freed and
// forgets the export.
void noteDescribed(const BufDesc& d, const void* ptr, const std::vector<int>& group_world, int me);
std::vector<int> takeDescribedTo(const void* alloc_base);

// Per-handle board in shared memory: cell (reader, owner) = how many of `owner`'s release announcements `reader` has
// processed, i.e. closed its imports for. An owner frees a released allocation once every rank it was described to has
// caught up (engine.cc reapReleased).
class AckBoard {
public:
  ~AckBoard();
  void create(Comm& comm, uint64_t token); // collective
  void destroy();
  bool valid() const { return base_ != nullptr; }
  void publish(int owner, uint64_t count);
  uint64_t seen(int reader, int owner) const;

private:
  std::atomic<uint64_t>* cell(int reader, int owner) const;
  void* base_ = nullptr;
  size_t bytes_ = 0;
  int nranks_ = 0;
  int me_ = 0;
};

// Imported peer allocations, keyed by (rank, buffer id, handle bytes).
class PeerCache {
public:
  ~PeerCache();
  // Pointer in MY address space for `d` owned by world rank `owner`. Throws CUDA_ERROR if the import fails.
  void* resolve(int owner, const BufDesc& d);
  // Drop the imports of buffers the owner has released (see CallMsg::release_count).
  // Returns true when something new was processed (the caller then acknowledges on the AckBoard).
  bool noteReleases(int owner, uint64_t release_count, const uint64_t* recent_ids);
  void forgetBuffer(int owner, uint64_t buffer_id);
  void forgetOwner(int owner);
  void clear();

private:
  struct Key {
    int owner;
    uint64_t buffer_id;
    std::string handle;
    bool operator<(const Key& o) const {
      if (owner != o.owner) return owner < o.owner;
      if (buffer_id != o.buffer_id) return buffer_id < o.buffer_id;
      return handle < o.handle;
    }
  };
  struct Entry {
    void* base;
    uint64_t last_use;
    uint32_t kind; // BufDesc::kind: decides how the mapping is closed
    uint64_t size;
  };
  static void closeImport(const Entry& e);
  std::map<Key, Entry> map_;
  std::map<int, uint64_t> releases_seen_; // per owner
  uint64_t tick_ = 0;
  void evictIfNeeded();
};

// Shared-memory mailbox: a host-side all-gather of CallMsg among the members of one communicator,
// a few hundred nanoseconds when everybody is already there. One segment per grid descriptor.
class Mailbox {
public:
  Mailbox() = default;
  ~Mailbox();
  Mailbox(const Mailbox&) = delete;
  Mailbox& operator=(const Mailbox&) = delete;

  // Collective over `comm` (creates and maps the segment).
  void create(Comm& comm, uint64_t token, int instance);
  void destroy();
  bool valid() const { return base_ != nullptr; }

  // Exchange on channel 0 (column groups) or 1 (row groups). `members` are indices into the creating
  // communicator; out[i] receives member i's message. Every member of the group must call this for every
  // operation on that channel, in the same order.
  void exchange(int channel, const std::vector<int>& members, int my_index, const CallMsg& mine,
                std::vector<CallMsg>& out);
  // forget all sequence numbers (collective; used when the process grid changes during autotuning)
  void reset(Comm& comm);

private:
  struct Slot {
    std::atomic<uint64_t> seq;
    CallMsg msg;
  };
  Slot* slot(int rank, int channel, int parity);
  void* base_ = nullptr;
  size_t bytes_ = 0;
  int nranks_ = 0;
  int me_ = 0;
  uint64_t seq_[2] = {0, 0};
  std::string name_;
  // A rank that arrives late (checkpoint I/O, load imbalance) is waited for, like the reference's MPI / NCCL backends
  // do. CUDECOMP_B200_HOST_TIMEOUT=<seconds> turns a longer wait into INTERNAL_ERROR (a debugging aid for mismatched
  // call sequences, which otherwise hang).
  double timeout_s_ = 0.0;
};

// Device-side flag pages (see kernels.h). One arena per handle, exported/imported ONCE: CUDA IPC hands out one
// handle per underlying driver allocation, and small cudaMalloc'ed buffers share a driver block, so per-descriptor
// 4 KiB allocations cannot be imported twice ("resource already mapped"). Every grid descriptor gets one 4 KiB slot.
class SignalArena {
public:
  static constexpr size_t kSlotBytes = 4096;
  static constexpr int kSlots = 512; // 2 MiB arena: a dedicated driver block
  ~SignalArena();
  void create(Comm& comm); // collective: allocate, zero, export, import everybody's
  void destroy(Comm* comm); // collective when comm != nullptr: nobody frees before everybody unmapped
  bool valid() const { return mine_ != nullptr; }
  uint64_t* mine(int slot) const { return reinterpret_cast<uint64_t*>(mine_ + static_cast<size_t>(slot) * kSlotBytes); }
  uint64_t* of(int rank, int slot) const {
    return reinterpret_cast<uint64_t*>(bases_[rank] + static_cast<size_t>(slot) * kSlotBytes);
  }
  void zeroSlot(int slot);
  uint32_t* errorWordDevice() const { return err_dev_; }
  uint32_t errorWordHost() const { return err_host_ ? *reinterpret_cast<volatile uint32_t*>(err_host_) : 0u; }
  void clearError() {
    if (err_host_) *reinterpret_cast<volatile uint32_t*>(err_host_) = 0u;
  }

private:
  char* mine_ = nullptr;
  std::vector<char*> bases_;
  std::vector<bool> imported_;
  uint32_t* err_host_ = nullptr; // pinned, mapped
  uint32_t* err_dev_ = nullptr;
};

} // namespace cdb

#endif
