// Stand-in for the reference's PRIVATE header include/internal/autotune.h, so that its API test suite
// (reference tests/ctest/api_tests.cc, which includes "internal/autotune.h" to test the candidate filters white-box)
// compiles unmodified against this library. Only the three candidate queries the tests call are declared; they are
// implemented in oracle/stubs_ctest/internal_adapter.cc on top of the public extension cudecompB200GetAutotuneCandidates.
// TEST INFRASTRUCTURE, not part of the product.
#ifndef CUDECOMP_B200_STUB_INTERNAL_AUTOTUNE_H
#define CUDECOMP_B200_STUB_INTERNAL_AUTOTUNE_H

#include <array>
#include <cstdint>
#include <vector>

#include "cudecomp.h"

namespace cudecomp {

std::vector<cudecompTransposeCommBackend_t> getAutotuneTransposeBackendCandidates(const cudecompGridDescAutotuneOptions_t* options);
std::vector<cudecompHaloCommBackend_t> getAutotuneHaloBackendCandidates(const cudecompGridDescAutotuneOptions_t* options);
std::vector<std::array<int32_t, 2>> getAutotunePdimCandidates(int nranks, cudecompRankOrder_t rank_order);

} // namespace cudecomp

#endif
