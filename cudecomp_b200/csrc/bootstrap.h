// Host-side bootstrap: process rendezvous and small control-plane collectives.
//
// Replaces what the reference obtains from MPI for its control plane
// (reference src/cudecomp.cc:61-67,510-589,910-915; include/internal/common.h:502-508):
// rank/size, allgather, bcast, barrier, allreduce and communicator split.  Ranks are
// separate processes (one per GPU) that find each other through the torchrun-style
// environment (RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT) and keep a full TCP mesh.
// Pencil data never goes through here.
#ifndef CUDECOMP_B200_BOOTSTRAP_H
#define CUDECOMP_B200_BOOTSTRAP_H

#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace cdb {

struct BootstrapError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// A communicator: an ordered list of world ranks. Collectives are blocking and must be
// issued in the same order by every member (the MPI rule).
struct Comm {
  int id = 0;               // identical on all members; used to tag messages
  std::vector<int> members; // world ranks
  int me = 0;               // my index in members
  uint32_t seq = 0;         // collective counter (message tag)
  uint32_t nsplits = 0;     // how many children were derived from this communicator
  int rank() const { return me; }
  int size() const { return static_cast<int>(members.size()); }
};
using CommPtr = std::shared_ptr<Comm>;

void worldInit();     // idempotent; reads the environment and builds the mesh
// Same with the rendezvous given explicitly (an application that has a real MPI: cudecompB200InitBootstrap): rank 0
// listens on `port`, everybody else dials `addr`:`port`.
void worldInitExplicit(int rank, int size, const std::string& addr, int port);
// A TCP port that was free a moment ago (rank 0 of an explicit rendezvous picks one and broadcasts it).
int pickFreePort();
void worldFinalize(); // closes the mesh
bool worldInitialized();
int worldRank();
int worldSize();
CommPtr worldComm();
CommPtr selfComm();

void allgather(Comm& c, const void* in, size_t bytes, void* out); // out holds size()*bytes
void bcast(Comm& c, void* buf, size_t bytes, int root);
void barrier(Comm& c);
CommPtr split(Comm& c, int color, int key); // color < 0 -> nullptr (MPI_UNDEFINED)
CommPtr dup(Comm& c);

enum class ReduceOp { SUM, MAX, MIN, LOR, LAND, BOR, PROD };
void allreduceF64(Comm& c, double* v, int n, ReduceOp op);
void allreduceI64(Comm& c, int64_t* v, int n, ReduceOp op);

// job-unique token shared by all members (used to name shared-memory segments)
uint64_t sharedToken(Comm& c);

} // namespace cdb

#endif
