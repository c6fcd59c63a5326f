// Opt-in performance report, compatible with the reference's (reference src/performance.cc, docs/env_vars.rst:39-95):
// same environment variables, same summary table and CSV columns, so tooling that parses cuDecomp reports
// (benchmark/benchmark_runner.py, heatmap scripts) keeps working. Disabled unless CUDECOMP_ENABLE_PERFORMANCE_REPORT=1;
// when disabled the hot path pays one null-pointer test.
//
// Mapping of the reference's columns onto this engine: "A2A" (transposes) / "SR" (halos) is the kernel that carries
// the exchange (the push launch); "local" is whatever else the call launched (the unpack kernel of the staged
// schedule, or the whole call when the communicator has one rank). A2A BW = source pencil bytes / A2A time, the
// reference's own definition (include/internal/transpose.h:316, src/performance.cc:391).
#ifndef CUDECOMP_B200_PERF_REPORT_H
#define CUDECOMP_B200_PERF_REPORT_H

#include <cuda_runtime.h>

#include <array>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "cudecomp.h"

struct cudecompHandle;
struct cudecompGridDesc;

namespace cdb {

struct PerfSettings {
  bool enabled = false;
  int detail = 0;
  int samples = 20;
  int warmup = 3;
  std::string write_dir;
  void readEnvironment();
};

struct PerfSample {
  cudaEvent_t start = nullptr, mid = nullptr, end = nullptr;
  bool has_mid = false;     // an exchange kernel was followed by local work
  bool exchange = false;    // the call crossed ranks at all
  bool valid = false;
};

struct PerfSeries {
  // key fields, already formatted the way the report prints them
  std::string operation, dtype, halos_a, halos_b, pads_a, pads_b, flag_a, flag_b; // flags: inplace/managed or periods
  int dim = -1;
  int64_t bytes = 0; // bytes the A2A/SR bandwidth column is computed from
  std::vector<PerfSample> ring;
  int64_t seen = 0; // calls so far (warm-up ones included)
};

class PerfReport {
public:
  explicit PerfReport(const PerfSettings& s) : s_(s) {}
  ~PerfReport();
  // Returns nullptr while the configuration is still in its warm-up calls.
  PerfSample* beginTranspose(int ax, int dir, cudecompDataType_t dtype, const int32_t* ih, const int32_t* oh,
                             const int32_t* ip, const int32_t* op, bool inplace, bool managed, int64_t bytes,
                             cudaStream_t stream);
  PerfSample* beginHalo(int ax, int dim, cudecompDataType_t dtype, const int32_t* halo, const bool* periods,
                        const int32_t* pad, bool managed, int64_t bytes, cudaStream_t stream);
  static void markExchangeDone(PerfSample* s, cudaStream_t stream) {
    if (s) {
      cudaEventRecord(s->mid, stream);
      s->has_mid = true;
    }
  }
  static void end(PerfSample* s, bool exchange, cudaStream_t stream) {
    if (s) {
      cudaEventRecord(s->end, stream);
      s->exchange = exchange;
      s->valid = true;
    }
  }
  // Collective over the handle's communicator: prints the summary on rank 0 (and writes CSV files when a directory is set).
  void print(cudecompHandle* h, cudecompGridDesc* gd);

private:
  PerfSample* next(PerfSeries& series, cudaStream_t stream);
  PerfSettings s_;
  std::map<std::string, PerfSeries> transposes_, halos_;
};

} // namespace cdb

#endif
