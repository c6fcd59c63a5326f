// Stand-in for the reference's PRIVATE header include/internal/exceptions.h: the one exception type its API tests
// name (EXPECT_THROW(..., cudecomp::InvalidUsage)). See internal/autotune.h in this directory. TEST INFRASTRUCTURE.
#ifndef CUDECOMP_B200_STUB_INTERNAL_EXCEPTIONS_H
#define CUDECOMP_B200_STUB_INTERNAL_EXCEPTIONS_H

#include <stdexcept>

#include "cudecomp.h"

namespace cudecomp {

class InvalidUsage : public std::runtime_error {
public:
  explicit InvalidUsage(const char* what) : std::runtime_error(what) {}
  cudecompResult_t getResult() const { return CUDECOMP_RESULT_INVALID_USAGE; }
};

} // namespace cudecomp

#endif
