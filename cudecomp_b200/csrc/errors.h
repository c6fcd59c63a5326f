// Error convention of the C ABI: nothing throws across it. Internals throw cdb::Error carrying the
// cudecompResult_t; every extern "C" entry catches, prints "CUDECOMP:ERROR: file:line kind (detail)"
// to stderr and returns the code (same contract as reference src/cudecomp.cc:416-443,
// include/internal/exceptions.h:63-146).
#ifndef CUDECOMP_B200_ERRORS_H
#define CUDECOMP_B200_ERRORS_H

#include <cuda_runtime.h>

#include <exception>
#include <sstream>
#include <string>

#include "cudecomp.h"

namespace cdb {

class Error : public std::exception {
public:
  Error(cudecompResult_t code, const char* kind, const char* file, int line, const std::string& detail) : code_(code) {
    std::ostringstream os;
    os << "CUDECOMP:ERROR: " << file << ":" << line << " " << kind;
    if (!detail.empty()) os << " (" << detail << ")";
    os << "\n";
    what_ = os.str();
  }
  const char* what() const noexcept override { return what_.c_str(); }
  cudecompResult_t code() const { return code_; }

private:
  cudecompResult_t code_;
  std::string what_;
};

} // namespace cdb

#define CDB_THROW(code, kind, msg) throw cdb::Error(code, kind, __FILE__, __LINE__, msg)
#define THROW_INVALID_USAGE(msg) CDB_THROW(CUDECOMP_RESULT_INVALID_USAGE, "Invalid usage.", msg)
#define THROW_NOT_SUPPORTED(msg) CDB_THROW(CUDECOMP_RESULT_NOT_SUPPORTED, "Not supported.", msg)
#define THROW_INTERNAL_ERROR(msg) CDB_THROW(CUDECOMP_RESULT_INTERNAL_ERROR, "Internal error.", msg)
#define THROW_CUDA_ERROR(msg) CDB_THROW(CUDECOMP_RESULT_CUDA_ERROR, "CUDA error.", msg)
#define THROW_MPI_ERROR(msg) CDB_THROW(CUDECOMP_RESULT_MPI_ERROR, "Bootstrap (MPI) error.", msg)

#define CHECK_CUDA(call)                                                                                               \
  do {                                                                                                                 \
    cudaError_t err__ = (call);                                                                                        \
    if (err__ != cudaSuccess) {                                                                                        \
      (void)cudaGetLastError();                                                                                        \
      THROW_CUDA_ERROR(std::string(#call) + ": " + cudaGetErrorString(err__));                                         \
    }                                                                                                                  \
  } while (0)

#endif
