// MPI subset exported for callers built against include/mpi_shim/mpi.h.
// Function list = what the reference's callers use (tests/cc/transpose_test.cc:569-669,
// tests/cc/halo_test.cc, benchmark/benchmark.cu) plus what this library's own API needs.
#include "mpi_shim.h"

#include <sys/time.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

namespace cdb {

namespace {
std::map<int, CommPtr> g_comms;
int g_next_handle = 3;
bool g_finalized = false;

int dtSize(MPI_Datatype dt) { return dt & 0xff; }
int dtKind(MPI_Datatype dt) { return dt >> 8; }

ReduceOp toOp(MPI_Op op) {
  switch (op) {
  case MPI_SUM: return ReduceOp::SUM;
  case MPI_MAX: return ReduceOp::MAX;
  case MPI_MIN: return ReduceOp::MIN;
  case MPI_LOR: return ReduceOp::LOR;
  case MPI_LAND: return ReduceOp::LAND;
  case MPI_BOR: return ReduceOp::BOR;
  case MPI_PROD: return ReduceOp::PROD;
  default: throw BootstrapError("unsupported MPI_Op");
  }
}

// Reduce via int64/double widening. `buf` holds `count` elements of `dt`.
void allreduceTyped(Comm& c, void* buf, int count, MPI_Datatype dt, MPI_Op op) {
  const int kind = dtKind(dt), size = dtSize(dt);
  if (kind == 3) {
    std::vector<double> v(count);
    for (int i = 0; i < count; ++i)
      v[i] = (size == 4) ? static_cast<double>(static_cast<float*>(buf)[i]) : static_cast<double*>(buf)[i];
    allreduceF64(c, v.data(), count, toOp(op));
    for (int i = 0; i < count; ++i) {
      if (size == 4)
        static_cast<float*>(buf)[i] = static_cast<float>(v[i]);
      else
        static_cast<double*>(buf)[i] = v[i];
    }
  } else if (kind == 1 || kind == 2 || kind == 5) {
    std::vector<int64_t> v(count);
    for (int i = 0; i < count; ++i) {
      switch (size) {
      case 1: v[i] = (kind == 1) ? static_cast<int8_t*>(buf)[i] : static_cast<uint8_t*>(buf)[i]; break;
      case 2: v[i] = static_cast<int16_t*>(buf)[i]; break;
      case 4: v[i] = (kind == 1) ? static_cast<int64_t>(static_cast<int32_t*>(buf)[i]) : static_cast<uint32_t*>(buf)[i]; break;
      default: v[i] = static_cast<int64_t*>(buf)[i]; break;
      }
    }
    allreduceI64(c, v.data(), count, toOp(op));
    for (int i = 0; i < count; ++i) {
      switch (size) {
      case 1: static_cast<int8_t*>(buf)[i] = static_cast<int8_t>(v[i]); break;
      case 2: static_cast<int16_t*>(buf)[i] = static_cast<int16_t>(v[i]); break;
      case 4: static_cast<int32_t*>(buf)[i] = static_cast<int32_t>(v[i]); break;
      default: static_cast<int64_t*>(buf)[i] = v[i]; break;
      }
    }
  } else {
    throw BootstrapError("unsupported datatype for reduction");
  }
}
} // namespace

CommPtr commFromHandle(int handle) {
  if (handle == 1) {
    if (!worldInitialized()) worldInit();
    return worldComm();
  }
  if (handle == 2) {
    if (!worldInitialized()) worldInit();
    return selfComm();
  }
  auto it = g_comms.find(handle);
  if (it == g_comms.end()) return nullptr;
  return it->second;
}

int registerComm(const CommPtr& c) {
  if (!c) return 0;
  int h = g_next_handle++;
  g_comms[h] = c;
  return h;
}

} // namespace cdb

using namespace cdb;

#define SHIM_TRY try {
#define SHIM_CATCH                                                                                                     \
  }                                                                                                                    \
  catch (const std::exception& e) {                                                                                    \
    std::fprintf(stderr, "CUDECOMP:ERROR: mpi shim: %s\n", e.what());                                                  \
    return MPI_ERR_OTHER;                                                                                              \
  }                                                                                                                    \
  return MPI_SUCCESS;

extern "C" {

int MPI_Init(int*, char***) {
  SHIM_TRY
  worldInit();
  g_finalized = false;
  SHIM_CATCH
}

int MPI_Init_thread(int* argc, char*** argv, int required, int* provided) {
  if (provided) *provided = required;
  return MPI_Init(argc, argv);
}

int MPI_Initialized(int* flag) {
  *flag = worldInitialized() ? 1 : 0;
  return MPI_SUCCESS;
}

int MPI_Finalized(int* flag) {
  *flag = g_finalized ? 1 : 0;
  return MPI_SUCCESS;
}

int MPI_Finalize(void) {
  SHIM_TRY
  g_comms.clear();
  worldFinalize();
  g_finalized = true;
  SHIM_CATCH
}

int MPI_Abort(MPI_Comm, int errorcode) {
  std::fflush(nullptr);
  _exit(errorcode ? errorcode : 1);
}

double MPI_Wtime(void) {
  timeval tv;
  gettimeofday(&tv, nullptr);
  return static_cast<double>(tv.tv_sec) + 1e-6 * static_cast<double>(tv.tv_usec);
}

int MPI_Get_processor_name(char* name, int* resultlen) {
  if (gethostname(name, MPI_MAX_PROCESSOR_NAME) != 0) std::strcpy(name, "localhost");
  name[MPI_MAX_PROCESSOR_NAME - 1] = 0;
  if (resultlen) *resultlen = static_cast<int>(std::strlen(name));
  return MPI_SUCCESS;
}

int MPI_Error_string(int errorcode, char* string, int* resultlen) {
  std::snprintf(string, MPI_MAX_ERROR_STRING, "cudecomp-b200 mpi shim error %d", errorcode);
  if (resultlen) *resultlen = static_cast<int>(std::strlen(string));
  return MPI_SUCCESS;
}

int MPI_Comm_rank(MPI_Comm comm, int* rank) {
  SHIM_TRY
  auto c = commFromHandle(comm);
  if (!c) throw BootstrapError("invalid communicator");
  *rank = c->rank();
  SHIM_CATCH
}

int MPI_Comm_size(MPI_Comm comm, int* size) {
  SHIM_TRY
  auto c = commFromHandle(comm);
  if (!c) throw BootstrapError("invalid communicator");
  *size = c->size();
  SHIM_CATCH
}

int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm* newcomm) {
  SHIM_TRY
  auto c = commFromHandle(comm);
  if (!c) throw BootstrapError("invalid communicator");
  auto n = split(*c, color == MPI_UNDEFINED ? -1 : color, key);
  *newcomm = registerComm(n);
  SHIM_CATCH
}

int MPI_Comm_split_type(MPI_Comm comm, int, int key, MPI_Info, MPI_Comm* newcomm) {
  // Single node: every rank shares memory, so SHARED == a copy of comm ordered by key. Applications commonly pick their
  // GPU as cudaSetDevice(rank in the SHARED communicator); CUDECOMP_B200_SHIM_RANKS_PER_NODE=n makes the shim report
  // "nodes" of n consecutive ranks, the way a launcher placing n ranks per node would, so that more ranks than GPUs can
  // share the devices of one box.
  int per_node = 0;
  if (const char* v = std::getenv("CUDECOMP_B200_SHIM_RANKS_PER_NODE")) per_node = std::atoi(v);
  if (per_node <= 0) return MPI_Comm_split(comm, 0, key, newcomm);
  int rank = 0;
  if (MPI_Comm_rank(comm, &rank) != MPI_SUCCESS) return MPI_ERR_OTHER;
  return MPI_Comm_split(comm, rank / per_node, key, newcomm);
}

int MPI_Comm_dup(MPI_Comm comm, MPI_Comm* newcomm) {
  SHIM_TRY
  auto c = commFromHandle(comm);
  if (!c) throw BootstrapError("invalid communicator");
  *newcomm = registerComm(dup(*c));
  SHIM_CATCH
}

int MPI_Comm_free(MPI_Comm* comm) {
  if (comm && *comm >= 3) g_comms.erase(*comm);
  if (comm) *comm = MPI_COMM_NULL;
  return MPI_SUCCESS;
}

MPI_Fint MPI_Comm_c2f(MPI_Comm comm) { return comm; }
MPI_Comm MPI_Comm_f2c(MPI_Fint comm) { return comm; }

int MPI_Barrier(MPI_Comm comm) {
  SHIM_TRY
  auto c = commFromHandle(comm);
  if (!c) throw BootstrapError("invalid communicator");
  barrier(*c);
  SHIM_CATCH
}

int MPI_Bcast(void* buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm) {
  SHIM_TRY
  auto c = commFromHandle(comm);
  if (!c) throw BootstrapError("invalid communicator");
  bcast(*c, buffer, static_cast<size_t>(count) * dtSize(datatype), root);
  SHIM_CATCH
}

int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
                  MPI_Datatype recvtype, MPI_Comm comm) {
  SHIM_TRY
  auto c = commFromHandle(comm);
  if (!c) throw BootstrapError("invalid communicator");
  size_t bytes = static_cast<size_t>(recvcount) * dtSize(recvtype);
  const void* in = sendbuf;
  if (sendbuf == MPI_IN_PLACE) {
    in = static_cast<char*>(recvbuf) + bytes * c->rank();
  } else if (static_cast<size_t>(sendcount) * dtSize(sendtype) != bytes) {
    throw BootstrapError("MPI_Allgather send/recv size mismatch");
  }
  allgather(*c, in, bytes, recvbuf);
  SHIM_CATCH
}

int MPI_Gather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
               MPI_Datatype recvtype, int root, MPI_Comm comm) {
  SHIM_TRY
  auto c = commFromHandle(comm);
  if (!c) throw BootstrapError("invalid communicator");
  size_t bytes = static_cast<size_t>(sendcount) * dtSize(sendtype);
  (void)recvcount;
  (void)recvtype;
  std::vector<char> all(bytes * c->size());
  const void* in = (sendbuf == MPI_IN_PLACE) ? static_cast<char*>(recvbuf) + bytes * c->rank() : sendbuf;
  allgather(*c, in, bytes, all.data());
  if (c->rank() == root) std::memcpy(recvbuf, all.data(), all.size());
  SHIM_CATCH
}

int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm) {
  SHIM_TRY
  auto c = commFromHandle(comm);
  if (!c) throw BootstrapError("invalid communicator");
  if (sendbuf != MPI_IN_PLACE) std::memcpy(recvbuf, sendbuf, static_cast<size_t>(count) * dtSize(datatype));
  allreduceTyped(*c, recvbuf, count, datatype, op);
  SHIM_CATCH
}

int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int root,
               MPI_Comm comm) {
  SHIM_TRY
  auto c = commFromHandle(comm);
  if (!c) throw BootstrapError("invalid communicator");
  size_t bytes = static_cast<size_t>(count) * dtSize(datatype);
  std::vector<char> tmp(bytes);
  // MPI_IN_PLACE is only legal at the root, where the contribution sits in recvbuf
  std::memcpy(tmp.data(), (sendbuf == MPI_IN_PLACE) ? recvbuf : sendbuf, bytes);
  allreduceTyped(*c, tmp.data(), count, datatype, op);
  if (c->rank() == root) std::memcpy(recvbuf, tmp.data(), bytes);
  SHIM_CATCH
}

} // extern "C"
