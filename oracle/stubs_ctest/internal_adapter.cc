// The reference's white-box candidate queries (declared in its private internal/autotune.h) expressed through this
// library's public extension entry point, for the reference's unmodified tests/ctest/api_tests.cc. The exception type
// the tests expect is the reference's own (its header-only include/internal/exceptions.h, found on the include path at
// build time). TEST INFRASTRUCTURE.
#include "cudecomp_b200_ext.h"
#include "internal/autotune.h"
#include <stdexcept>

#include "internal/exceptions.h"

namespace cudecomp {

namespace {
void check(cudecompResult_t res) {
  if (res == CUDECOMP_RESULT_INVALID_USAGE) throw InvalidUsage(__FILE__, __LINE__, "autotune candidate filters rejected");
  if (res != CUDECOMP_RESULT_SUCCESS) throw std::runtime_error("cudecompB200GetAutotuneCandidates failed");
}
} // namespace

std::vector<cudecompTransposeCommBackend_t> getAutotuneTransposeBackendCandidates(const cudecompGridDescAutotuneOptions_t* options) {
  int32_t values[8], n = 0;
  check(cudecompB200GetAutotuneCandidates(options, 1, CUDECOMP_RANK_ORDER_ROW_MAJOR, values, &n, nullptr, nullptr, nullptr, 0, nullptr));
  std::vector<cudecompTransposeCommBackend_t> out;
  for (int32_t i = 0; i < n; ++i) out.push_back(static_cast<cudecompTransposeCommBackend_t>(values[i]));
  return out;
}

std::vector<cudecompHaloCommBackend_t> getAutotuneHaloBackendCandidates(const cudecompGridDescAutotuneOptions_t* options) {
  int32_t values[5], n = 0;
  check(cudecompB200GetAutotuneCandidates(options, 1, CUDECOMP_RANK_ORDER_ROW_MAJOR, nullptr, nullptr, values, &n, nullptr, 0, nullptr));
  std::vector<cudecompHaloCommBackend_t> out;
  for (int32_t i = 0; i < n; ++i) out.push_back(static_cast<cudecompHaloCommBackend_t>(values[i]));
  return out;
}

std::vector<std::array<int32_t, 2>> getAutotunePdimCandidates(int nranks, cudecompRankOrder_t rank_order) {
  cudecompGridDescAutotuneOptions_t options;
  check(cudecompGridDescAutotuneOptionsSetDefaults(&options));
  int32_t pdims[256][2], n = 0;
  check(cudecompB200GetAutotuneCandidates(&options, nranks, rank_order, nullptr, nullptr, nullptr, nullptr, pdims, 256, &n));
  std::vector<std::array<int32_t, 2>> out;
  for (int32_t i = 0; i < n && i < 256; ++i) out.push_back({pdims[i][0], pdims[i][1]});
  return out;
}

} // namespace cudecomp
