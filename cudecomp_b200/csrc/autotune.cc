// Grid-descriptor autotuning. Keeps the reference protocol (reference src/autotune.cc:275-769 transposes,
// :771-1124 halos): sweep the factorisations of nranks in locality-first order, skip empty / disallowed
// uneven grids, n_warmup_trials untimed + n_trials timed runs of the enabled operations with CUDA events
// on one stream, score = rank-average of the weighted per-trial time, smallest wins, identical on all
// ranks. What is swept per process grid is different: there are no MPI/NCCL/NVSHMEM libraries to choose
// between, so a "backend" is a schedule of the one peer-store engine:
//   backend values 1..5 (MPI*, NCCL*)  -> direct: peers store straight into the destination buffers
//   backend values 6..8 (NVSHMEM*)     -> staged: peers store into the workspace, local unpack
// and for each of them the CTA count of the copy kernels is swept too.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>

#include "engine.h"
#include "errors.h"
#include "launch_params.h"

namespace cdb {

namespace {

bool backendIsStaged(int b) { return b >= CUDECOMP_TRANSPOSE_COMM_NVSHMEM; }

// "min,max" with nonnegative min, positive max, min <= max (reference src/autotune.cc:192-213)
std::pair<int32_t, int32_t> parseRange(const char* name) {
  const char* v = std::getenv(name);
  if (!v) return {1, std::numeric_limits<int32_t>::max()};
  const std::string s(v);
  const size_t comma = s.find(',');
  bool valid = comma != std::string::npos && comma > 0 && comma + 1 < s.size() && s.find(',', comma + 1) == std::string::npos;
  long lo = 0, hi = 0;
  if (valid) {
    char* e1 = nullptr;
    char* e2 = nullptr;
    const std::string a = s.substr(0, comma), b = s.substr(comma + 1);
    lo = std::strtol(a.c_str(), &e1, 10);
    hi = std::strtol(b.c_str(), &e2, 10);
    valid = *e1 == 0 && *e2 == 0 && a.find_first_not_of("0123456789") == std::string::npos &&
            b.find_first_not_of("0123456789") == std::string::npos && lo >= 0 && hi > 0 && lo <= hi;
  }
  if (!valid)
    THROW_INVALID_USAGE(std::string(name) + " must be comma-separated nonnegative min and positive max with min <= max");
  return {static_cast<int32_t>(lo), static_cast<int32_t>(hi)};
}

// CUDECOMP_AUTOTUNE_{TRANSPOSE,HALO}_BACKENDS: comma-separated backend names, a leading '^' turns the list into an
// exclusion list; unknown or empty names are errors (reference src/autotune.cc:108-143)
std::vector<int> filterByEnvironment(const char* env_name, const std::vector<std::pair<std::string, int>>& supported,
                                     std::vector<int> available) {
  const char* env = std::getenv(env_name);
  if (!env) return available;
  std::string value(env);
  const bool exclude = !value.empty() && value[0] == '^';
  if (exclude) value.erase(0, 1);
  std::vector<int> listed;
  size_t start = 0;
  for (;;) {
    size_t end = value.find(',', start);
    if (end == std::string::npos) end = value.size();
    const std::string name = value.substr(start, end - start);
    auto it = std::find_if(supported.begin(), supported.end(), [&](const auto& e) { return e.first == name; });
    if (it == supported.end())
      THROW_INVALID_USAGE(std::string(env_name) + " contains unknown or empty backend name '" + name + "'");
    listed.push_back(it->second);
    if (end == value.size()) break;
    start = end + 1;
  }
  available.erase(std::remove_if(available.begin(), available.end(),
                                 [&](int b) {
                                   const bool is_listed = std::find(listed.begin(), listed.end(), b) != listed.end();
                                   return exclude ? is_listed : !is_listed;
                                 }),
                  available.end());
  return available;
}

} // namespace

std::vector<int> autotuneTransposeBackendCandidates(const cudecompGridDescAutotuneOptions_t* o) {
  static const std::vector<std::pair<std::string, int>> supported = {
      {"MPI_P2P", CUDECOMP_TRANSPOSE_COMM_MPI_P2P},       {"MPI_P2P_PL", CUDECOMP_TRANSPOSE_COMM_MPI_P2P_PL},
      {"MPI_A2A", CUDECOMP_TRANSPOSE_COMM_MPI_A2A},       {"NCCL", CUDECOMP_TRANSPOSE_COMM_NCCL},
      {"NCCL_PL", CUDECOMP_TRANSPOSE_COMM_NCCL_PL},       {"NVSHMEM", CUDECOMP_TRANSPOSE_COMM_NVSHMEM},
      {"NVSHMEM_PL", CUDECOMP_TRANSPOSE_COMM_NVSHMEM_PL}, {"NVSHMEM_SM", CUDECOMP_TRANSPOSE_COMM_NVSHMEM_SM}};
  std::vector<int> c = {1, 2, 3, 4, 5, 6, 7, 8}; // every value is a schedule of the one engine, so all are available
  c = filterByEnvironment("CUDECOMP_AUTOTUNE_TRANSPOSE_BACKENDS", supported, c);
  c.erase(std::remove_if(c.begin(), c.end(),
                         [&](int b) {
                           return (o->disable_mpi_backends && b <= CUDECOMP_TRANSPOSE_COMM_MPI_A2A) ||
                                  (o->disable_nccl_backends &&
                                   (b == CUDECOMP_TRANSPOSE_COMM_NCCL || b == CUDECOMP_TRANSPOSE_COMM_NCCL_PL)) ||
                                  (o->disable_nvshmem_backends && b >= CUDECOMP_TRANSPOSE_COMM_NVSHMEM);
                         }),
          c.end());
  if (c.empty()) THROW_INVALID_USAGE("Transpose backend autotuning has no usable candidates after applying filters");
  return c;
}

std::vector<int> autotuneHaloBackendCandidates(const cudecompGridDescAutotuneOptions_t* o) {
  static const std::vector<std::pair<std::string, int>> supported = {{"MPI", CUDECOMP_HALO_COMM_MPI},
                                                                     {"MPI_BLOCKING", CUDECOMP_HALO_COMM_MPI_BLOCKING},
                                                                     {"NCCL", CUDECOMP_HALO_COMM_NCCL},
                                                                     {"NVSHMEM", CUDECOMP_HALO_COMM_NVSHMEM},
                                                                     {"NVSHMEM_BLOCKING", CUDECOMP_HALO_COMM_NVSHMEM_BLOCKING}};
  std::vector<int> c = {1, 2, 3, 4, 5};
  c = filterByEnvironment("CUDECOMP_AUTOTUNE_HALO_BACKENDS", supported, c);
  c.erase(std::remove_if(c.begin(), c.end(),
                         [&](int b) {
                           return (o->disable_mpi_backends && b <= CUDECOMP_HALO_COMM_MPI_BLOCKING) ||
                                  (o->disable_nccl_backends && b == CUDECOMP_HALO_COMM_NCCL) ||
                                  (o->disable_nvshmem_backends && b >= CUDECOMP_HALO_COMM_NVSHMEM);
                         }),
          c.end());
  if (c.empty()) THROW_INVALID_USAGE("Halo backend autotuning has no usable candidates after applying filters");
  return c;
}

std::vector<std::array<int32_t, 2>> autotunePdimCandidates(int nranks, bool col_major) {
  auto cand = pdimCandidates(nranks, col_major);
  auto rows = parseRange("CUDECOMP_AUTOTUNE_P_ROW_RANGE");
  auto cols = parseRange("CUDECOMP_AUTOTUNE_P_COL_RANGE");
  cand.erase(std::remove_if(cand.begin(), cand.end(),
                            [&](const std::array<int32_t, 2>& p) {
                              return p[0] < rows.first || p[0] > rows.second || p[1] < cols.first || p[1] > cols.second;
                            }),
             cand.end());
  if (cand.empty()) THROW_INVALID_USAGE("Process-grid autotuning has no usable candidates after applying filters");
  return cand;
}

namespace {

std::vector<std::array<int32_t, 2>> autotunePdims(cudecompHandle_t h, cudecompGridDesc_t gd) {
  return autotunePdimCandidates(h->nranks, gd->config.rank_order == CUDECOMP_RANK_ORDER_COL_MAJOR);
}

bool gridUsable(const cudecompGridDesc_t gd, const std::array<int32_t, 2>& p, bool allow_uneven) {
  const auto& d = gd->config.gdims_dist;
  if (p[0] > kMaxPeers + 1 || p[1] > kMaxPeers + 1) return false; // larger than a launch can handshake with
  if (p[0] > std::min(d[0], d[1]) || p[1] > std::min(d[1], d[2])) return false; // empty pencils
  if (!allow_uneven && (d[0] % p[0] != 0 || d[1] % p[0] != 0 || d[1] % p[1] != 0 || d[2] % p[1] != 0)) return false;
  return true;
}

struct Stats {
  double min, max, avg, std;
};

// min/max/avg/std over (ranks x trials) (reference src/autotune.cc:167-188)
Stats reduceTimes(cudecompHandle_t h, const std::vector<double>& t) {
  double mn = std::numeric_limits<double>::max(), mx = 0, sum = 0, sq = 0;
  for (double v : t) {
    mn = std::min(mn, v);
    mx = std::max(mx, v);
    sum += v;
    sq += v * v;
  }
  double n = static_cast<double>(t.size());
  double red[3] = {sum, sq, n};
  allreduceF64(*h->comm, red, 3, ReduceOp::SUM);
  allreduceF64(*h->comm, &mn, 1, ReduceOp::MIN);
  allreduceF64(*h->comm, &mx, 1, ReduceOp::MAX);
  Stats s;
  s.min = mn;
  s.max = mx;
  s.avg = red[0] / red[2];
  s.std = std::sqrt(std::max(0.0, red[1] / red[2] - s.avg * s.avg));
  return s;
}

struct DeviceBuffer {
  void* p = nullptr;
  ~DeviceBuffer() {
    if (p) cudaFree(p);
  }
  // whole 2 MiB pages, like cudecompMalloc: peers must be able to import the buffer (see peer.cc describeBuffer)
  void alloc(size_t bytes) {
    const size_t page = size_t(2) << 20;
    CHECK_CUDA(cudaMalloc(&p, (std::max<size_t>(bytes, 1) + page - 1) / page * page));
  }
};

struct Events {
  std::vector<cudaEvent_t> ev;
  explicit Events(size_t n) : ev(n) {
    for (auto& e : ev) CHECK_CUDA(cudaEventCreate(&e));
  }
  ~Events() {
    for (auto& e : ev) cudaEventDestroy(e);
  }
};

const int kOps[4][2] = {{0, 1}, {1, 1}, {2, -1}, {1, -1}}; // XY, YZ, ZY, YX as (ax, dir)

// What is swept per (process grid, schedule family) besides the reference's dimensions: the launch schedule of the one
// engine (BASELINE north star: "tile shape / CTA count / one-shot-vs-pairwise").
struct Schedule {
  int ctas = 0;       // CTAs per launch (0: library default)
  int tile_bytes = 0; // row-copy tile (0: 32 KiB)
  int peer_order = 0; // 0 one-shot interleaved, 1 pairwise rounds
  int balance = 0;    // balanced grid
  int chunks = 0;     // chunked staged schedule (in-place / staged calls)
  int variant = 0;    // 1: TMA bulk row copy, 2: 256-bit LDG/STG
  int pull = 0;       // 1: receiver-driven direct transposes
};

void applySchedule(cudecompGridDesc_t gd, const Schedule& s) {
  gd->grid_ctas = s.ctas;
  gd->tile_bytes = s.tile_bytes;
  gd->peer_order = s.peer_order;
  gd->balance_grid = s.balance;
  gd->pipeline_chunks = s.chunks;
  gd->kernel_variant = s.variant;
  gd->pull_mode = s.pull;
}

Schedule currentSchedule(const cudecompGridDesc_t gd) {
  Schedule s;
  s.ctas = gd->grid_ctas;
  s.tile_bytes = gd->tile_bytes;
  s.peer_order = gd->peer_order;
  s.balance = gd->balance_grid;
  s.chunks = gd->pipeline_chunks;
  s.variant = gd->kernel_variant;
  s.pull = gd->pull_mode;
  return s;
}

std::string describeSchedule(const Schedule& s) {
  std::string d = "CTAs: " + std::to_string(s.ctas) + ", tile: " + std::to_string(s.tile_bytes ? s.tile_bytes : kDefaultTileBytes) +
                  ", order: " + (s.peer_order ? "pairwise" : "one-shot");
  if (s.balance) d += ", balanced grid";
  if (s.chunks > 1) d += ", chunks: " + std::to_string(s.chunks);
  if (s.variant == 1) d += ", TMA bulk";
  if (s.variant == 2) d += ", 256-bit LDG/STG";
  if (s.pull) d += ", receiver-driven";
  return d;
}

// CUDECOMP_B200_AUTOTUNE_SCHEDULES: which schedule dimensions the second tuning phase explores on the winning
// (grid, family): comma list of tile, order, balance, chunks, bulk, pull, or "all" / "none". The CTA count is always
// swept (phase 1). Default: tile, order, chunks -- the north star's "tile shape / CTA count / one-shot-vs-pairwise", plus
// the chunk count of staged (in-place) calls. The other three are measured (profiles/r2_n2_schedules.md: balanced grid
// within noise, TMA bulk equal on the wire and slower locally, receiver-driven slower in both directions at once) and
// stay opt-in.
struct ScheduleDims {
  bool tile = true, order = true, balance = false, chunks = true, bulk = false, pull = false;
};

ScheduleDims scheduleDimsFromEnvironment() {
  ScheduleDims d;
  const char* v = std::getenv("CUDECOMP_B200_AUTOTUNE_SCHEDULES");
  if (!v) return d;
  d.tile = d.order = d.chunks = false; // an explicit list replaces the default
  const std::string s(v);
  size_t start = 0;
  while (start <= s.size()) {
    size_t end = s.find(',', start);
    if (end == std::string::npos) end = s.size();
    const std::string name = s.substr(start, end - start);
    if (name == "all") d.tile = d.order = d.balance = d.chunks = d.bulk = d.pull = true;
    else if (name == "none") d = ScheduleDims{false, false, false, false, false, false};
    else if (name == "pull") d.pull = true;
    else if (name == "tile") d.tile = true;
    else if (name == "order") d.order = true;
    else if (name == "balance") d.balance = true;
    else if (name == "chunks") d.chunks = true;
    else if (name == "bulk") d.bulk = true;
    else if (!name.empty())
      THROW_INVALID_USAGE("CUDECOMP_B200_AUTOTUNE_SCHEDULES contains unknown schedule dimension '" + name + "'");
    start = end + 1;
  }
  return d;
}

void autotuneTransposes(cudecompHandle_t h, cudecompGridDesc_t gd, const cudecompGridDescAutotuneOptions_t* o,
                        bool tune_pdims, bool tune_backend) {
  const double t_start = MPI_Wtime();
  if (h->rank == 0) std::printf("CUDECOMP: Running transpose autotuning...\n");

  std::vector<std::array<int32_t, 2>> grids;
  if (tune_pdims)
    grids = autotunePdims(h, gd);
  else
    grids.push_back({gd->config.pdims[0], gd->config.pdims[1]});

  std::vector<int> backends;
  if (tune_backend) {
    // Backend values of one family run the same schedule (1-5 direct, 6-8 staged): time one representative of each
    // family that the filters leave enabled, the first in enum order.
    bool have_direct = false, have_staged = false;
    for (int b : autotuneTransposeBackendCandidates(o)) {
      bool& seen = backendIsStaged(b) ? have_staged : have_direct;
      if (!seen) backends.push_back(b);
      seen = true;
    }
  } else {
    backends.push_back(gd->config.transpose_comm_backend);
  }

  const int64_t es = dtypeSize(o->dtype);

  // buffers large enough for every candidate
  int64_t data_elems = 0, work_elems = 0;
  for (auto& p : grids) {
    if (!gridUsable(gd, p, o->allow_uneven_decompositions)) continue;
    setGeometry(gd, p);
    for (int op = 0; op < 4; ++op) {
      TransposeAxes ax = transposeAxes(kOps[op][0], kOps[op][1]);
      data_elems = std::max(data_elems, pencilInfo(gd->geom, gd->pidx, ax.a, o->transpose_input_halo_extents[op],
                                                   o->transpose_input_padding[op]).size);
      data_elems = std::max(data_elems, pencilInfo(gd->geom, gd->pidx, ax.b, o->transpose_output_halo_extents[op],
                                                   o->transpose_output_padding[op]).size);
    }
    work_elems = std::max(work_elems, transposeWorkspaceSize(gd->geom));
  }
  bool need_data2 = false;
  for (int op = 0; op < 4; ++op)
    if (!o->transpose_use_inplace_buffers[op]) need_data2 = true;

  double t_best = std::numeric_limits<double>::max();
  std::array<int32_t, 2> best_grid{0, 0};
  int best_backend = backends[0];
  const Schedule saved_schedule = currentSchedule(gd);
  Schedule best_schedule = saved_schedule;
  bool valid = false;

  if (!h->have_device) {
    // nothing can be timed: take the first usable grid so that geometry queries still work on CPU-only hosts
    for (auto& p : grids) {
      if (!gridUsable(gd, p, o->allow_uneven_decompositions)) continue;
      best_grid = p;
      valid = true;
      break;
    }
    if (h->rank == 0) std::printf("CUDECOMP:WARN: no CUDA device, autotuning skipped (first usable grid selected)\n");
  } else {
    DeviceBuffer data, data2, work;
    if (data_elems > 0) {
      data.alloc(static_cast<size_t>(data_elems * es));
      CHECK_CUDA(cudaMemset(data.p, 0, static_cast<size_t>(data_elems * es)));
      if (need_data2) {
        data2.alloc(static_cast<size_t>(data_elems * es));
        CHECK_CUDA(cudaMemset(data2.p, 0, static_cast<size_t>(data_elems * es)));
      }
      work.alloc(static_cast<size_t>(work_elems * es));
    }
    cudaStream_t stream = 0;
    Events ev(5);
    const bool saved_staged = gd->force_staged;

    // Times the enabled operations on the geometry / backend / schedule currently set on `gd` with the reference's
    // protocol and prints the candidate block. Returns the rank-averaged weighted time, or a negative value when the
    // candidate was abandoned after its first trial (skip_threshold). Collective.
    auto timeCandidate = [&](const std::array<int32_t, 2>& p, const Schedule& sched) -> double {
      applySchedule(gd, sched);
      auto runOp = [&](int op) {
        void* in = data.p;
        void* out = o->transpose_use_inplace_buffers[op] ? data.p : data2.p;
        runTranspose(h, gd, kOps[op][0], kOps[op][1], in, out, work.p, o->dtype, o->transpose_input_halo_extents[op],
                     o->transpose_output_halo_extents[op], o->transpose_input_padding[op],
                     o->transpose_output_padding[op], stream);
      };
      for (int w = 0; w < o->n_warmup_trials; ++w)
        for (int op = 0; op < 4; ++op)
          if (o->transpose_op_weights[op] != 0.0) runOp(op);
      CHECK_CUDA(cudaStreamSynchronize(stream));

      std::vector<double> total, total_w;
      std::vector<double> per_op[4];
      bool skipped = false;
      for (int t = 0; t < o->n_trials; ++t) {
        CHECK_CUDA(cudaEventRecord(ev.ev[0], stream));
        for (int op = 0; op < 4; ++op) {
          if (o->transpose_op_weights[op] != 0.0) runOp(op);
          CHECK_CUDA(cudaEventRecord(ev.ev[op + 1], stream));
        }
        CHECK_CUDA(cudaStreamSynchronize(stream));
        double tt = 0, tw = 0;
        for (int op = 0; op < 4; ++op) {
          float ms = 0;
          CHECK_CUDA(cudaEventElapsedTime(&ms, ev.ev[op], ev.ev[op + 1]));
          per_op[op].push_back(ms);
          tt += ms;
          tw += ms * o->transpose_op_weights[op];
        }
        total.push_back(tt);
        total_w.push_back(tw);
        if (t == 0 && o->skip_threshold > 0.0) {
          // agree across ranks whether this configuration is hopeless (reference src/autotune.cc:578-602)
          double first = tw;
          allreduceF64(*h->comm, &first, 1, ReduceOp::SUM);
          first /= h->nranks;
          if (o->skip_threshold * first > t_best) {
            skipped = true;
            break;
          }
        }
      }
      const std::string desc = describeSchedule(sched);
      if (skipped) {
        if (h->rank == 0)
          std::printf("CUDECOMP:\tgrid: %d x %d, backend: %s, %s \nCUDECOMP:\t(skipped) \n", p[0], p[1],
                      cudecompTransposeCommBackendToString(gd->config.transpose_comm_backend), desc.c_str());
        return -1.0;
      }
      Stats st = reduceTimes(h, total), sw = reduceTimes(h, total_w);
      Stats so[4];
      for (int op = 0; op < 4; ++op) so[op] = reduceTimes(h, per_op[op]);
      if (h->rank == 0) {
        std::printf("CUDECOMP:\tgrid: %d x %d, backend: %s, %s \n"
                    "CUDECOMP:\tTotal time min/max/avg/std [ms]: %f/%f/%f/%f\n"
                    "CUDECOMP:\t           min/max/avg/std [ms]: %f/%f/%f/%f (weighted)\n",
                    p[0], p[1], cudecompTransposeCommBackendToString(gd->config.transpose_comm_backend), desc.c_str(), st.min,
                    st.max, st.avg, st.std, sw.min, sw.max, sw.avg, sw.std);
        const char* names[4] = {"XY", "YZ", "ZY", "YX"};
        for (int op = 0; op < 4; ++op)
          std::printf("CUDECOMP:\tTranspose%s time min/max/avg/std [ms]: %f/%f/%f/%f%s\n", names[op], so[op].min, so[op].max,
                      so[op].avg, so[op].std, o->transpose_op_weights[op] == 0.0 ? " (skipped)" : "");
      }
      return sw.avg;
    };

    // ---- phase 1 (reference protocol): process grid x schedule family x CTA count
    for (auto& p : grids) {
      if (!gridUsable(gd, p, o->allow_uneven_decompositions)) continue;
      valid = true;
      setGeometry(gd, p);
      if (gd->mbox.valid()) gd->mbox.reset(*h->comm);

      for (int backend : backends) {
        gd->config.transpose_comm_backend = static_cast<cudecompTransposeCommBackend_t>(backend);
        gd->force_staged = backendIsStaged(backend);
        // CTA counts to try: the library default, and two / one CTAs per SM
        std::vector<int> cta_list = {saved_schedule.ctas};
        if (tune_backend && h->sm_count > 0) {
          cta_list = {0, 2 * h->sm_count, h->sm_count};
        }
        for (int ctas : cta_list) {
          Schedule sched = saved_schedule;
          sched.ctas = ctas;
          const double t = timeCandidate(p, sched);
          if (t >= 0 && t < t_best) {
            t_best = t;
            best_grid = p;
            best_backend = backend;
            best_schedule = sched;
          }
        }
      }
    }

    // ---- phase 2: the launch schedule of the winner, one dimension after the other (greedy); an alternative is kept
    // when it is faster than everything seen so far
    const ScheduleDims dims = scheduleDimsFromEnvironment();
    if (valid && tune_backend && t_best < std::numeric_limits<double>::max() &&
        (dims.tile || dims.order || dims.balance || dims.chunks || dims.bulk || dims.pull)) {
      setGeometry(gd, best_grid);
      if (gd->mbox.valid()) gd->mbox.reset(*h->comm);
      gd->config.transpose_comm_backend = static_cast<cudecompTransposeCommBackend_t>(best_backend);
      gd->force_staged = backendIsStaged(best_backend);
      bool any_inplace = false;
      for (int op = 0; op < 4; ++op)
        if (o->transpose_op_weights[op] != 0.0 && o->transpose_use_inplace_buffers[op]) any_inplace = true;
      std::vector<Schedule> alternatives;
      auto tryAlternatives = [&]() {
        for (const Schedule& alt : alternatives) {
          const double t = timeCandidate(best_grid, alt);
          if (t >= 0 && t < t_best) {
            t_best = t;
            best_schedule = alt;
          }
        }
        alternatives.clear();
      };
      if (dims.tile) {
        for (int tb : {16384, 65536}) {
          Schedule a = best_schedule;
          a.tile_bytes = tb;
          alternatives.push_back(a);
        }
        tryAlternatives();
      }
      if (dims.order) {
        Schedule a = best_schedule;
        a.peer_order = 1;
        alternatives.push_back(a);
        tryAlternatives();
      }
      if (dims.balance) {
        Schedule a = best_schedule;
        a.balance = 1;
        alternatives.push_back(a);
        tryAlternatives();
      }
      if (dims.pull && !backendIsStaged(best_backend)) {
        Schedule a = best_schedule;
        a.pull = 1;
        alternatives.push_back(a);
        tryAlternatives();
      }
      if (dims.chunks && (any_inplace || backendIsStaged(best_backend))) {
        // around what the library would pick from the pencil size (engine.cc autoFusedChunks): half and double
        const int auto_k = autoFusedChunks(static_cast<int64_t>(gd->geom.gdims[0]) * gd->geom.gdims[1] * gd->geom.gdims[2] /
                                           std::max(1, h->nranks) * es);
        for (int k : {std::max(1, auto_k / 2), std::min(32, auto_k * 2)}) {
          if (k == auto_k) continue;
          Schedule a = best_schedule;
          a.chunks = k;
          alternatives.push_back(a);
        }
        tryAlternatives();
      }
      if (dims.bulk) {
        for (int variant : {1, 2}) { // TMA bulk, 256-bit LDG/STG
          Schedule a = best_schedule;
          a.variant = variant;
          alternatives.push_back(a);
        }
        tryAlternatives();
      }
    }

    applySchedule(gd, saved_schedule);
    gd->force_staged = saved_staged;
    CHECK_CUDA(cudaDeviceSynchronize());
    // our device buffers are about to be freed: every peer drops its imports of them first
    h->peers.clear();
    barrier(*h->comm);
  }

  if (!valid) THROW_NOT_SUPPORTED("No valid decomposition found during autotuning with provided arguments.");

  gd->config.pdims[0] = best_grid[0];
  gd->config.pdims[1] = best_grid[1];
  gd->config.transpose_comm_backend = static_cast<cudecompTransposeCommBackend_t>(best_backend);
  if (tune_backend) applySchedule(gd, best_schedule);
  gd->force_staged = false;
  if (h->rank == 0) {
    std::printf("CUDECOMP: SELECTED: grid: %d x %d, backend: %s, Avg. time (weighted) [ms]: %f\n", best_grid[0],
                best_grid[1], cudecompTransposeCommBackendToString(gd->config.transpose_comm_backend),
                h->have_device ? t_best : 0.0);
    if (tune_backend && h->have_device) std::printf("CUDECOMP: SELECTED schedule: %s\n", describeSchedule(best_schedule).c_str());
  }
  barrier(*h->comm);
  if (h->rank == 0) std::printf("CUDECOMP: transpose autotuning time [s]: %f\n", MPI_Wtime() - t_start);
}

void autotuneHalos(cudecompHandle_t h, cudecompGridDesc_t gd, const cudecompGridDescAutotuneOptions_t* o,
                   bool tune_pdims, bool tune_backend) {
  const double t_start = MPI_Wtime();
  if (h->rank == 0) std::printf("CUDECOMP: Running halo autotuning...\n");
  if (o->halo_axis < 0 || o->halo_axis > 2) THROW_INVALID_USAGE("halo_axis out of range");

  std::vector<std::array<int32_t, 2>> grids;
  if (tune_pdims)
    grids = autotunePdims(h, gd);
  else
    grids.push_back({gd->config.pdims[0], gd->config.pdims[1]});

  // all halo backend values run the same peer-store schedule; keep the caller's value unless it asked to tune
  int backend = gd->config.halo_comm_backend;
  if (tune_backend) backend = autotuneHaloBackendCandidates(o).front();
  const int64_t es = dtypeSize(o->dtype);
  const int ax = o->halo_axis;

  int64_t data_elems = 0, work_elems = 0;
  for (auto& p : grids) {
    if (!gridUsable(gd, p, o->allow_uneven_decompositions)) continue;
    setGeometry(gd, p);
    data_elems = std::max(data_elems, pencilInfo(gd->geom, gd->pidx, ax, o->halo_extents, o->halo_padding).size);
    work_elems = std::max(work_elems, haloWorkspaceSize(gd->geom, gd->pidx, ax, o->halo_extents));
  }
  // the workspace size depends on the rank's slab; peers only need each rank's own size to be sufficient

  double t_best = std::numeric_limits<double>::max();
  std::array<int32_t, 2> best_grid{0, 0};
  bool valid = false;

  if (!h->have_device) {
    for (auto& p : grids) {
      if (!gridUsable(gd, p, o->allow_uneven_decompositions)) continue;
      best_grid = p;
      valid = true;
      break;
    }
    if (h->rank == 0) std::printf("CUDECOMP:WARN: no CUDA device, autotuning skipped (first usable grid selected)\n");
  } else {
    DeviceBuffer data, work;
    if (data_elems > 0) {
      data.alloc(static_cast<size_t>(data_elems * es));
      CHECK_CUDA(cudaMemset(data.p, 0, static_cast<size_t>(data_elems * es)));
      work.alloc(static_cast<size_t>(std::max<int64_t>(work_elems, 64) * es));
    }
    cudaStream_t stream = 0;
    Events ev(4);
    for (auto& p : grids) {
      if (!gridUsable(gd, p, o->allow_uneven_decompositions)) continue;
      setGeometry(gd, p);
      if (gd->mbox.valid()) gd->mbox.reset(*h->comm);
      // a halo wider than some rank's slab cannot be exchanged on this grid
      bool too_wide = false;
      for (int dim = 0; dim < 3 && !too_wide; ++dim) {
        if (dim == ax || o->halo_extents[dim] == 0) continue;
        const int P = gd->geom.pdims[haloCommAxis(ax, dim)];
        if (P == 1) continue;
        auto splits = getSplits(gd->geom.gdims_dist[dim], P, gd->geom.gdims[dim] - gd->geom.gdims_dist[dim]);
        if (o->halo_extents[dim] > *std::min_element(splits.begin(), splits.end())) too_wide = true;
      }
      if (too_wide) continue;
      valid = true;
      auto runAll = [&]() {
        for (int dim = 0; dim < 3; ++dim)
          runHalo(h, gd, ax, data.p, work.p, o->dtype, o->halo_extents, o->halo_periods, dim, o->halo_padding, stream);
      };
      for (int w = 0; w < o->n_warmup_trials; ++w) runAll();
      CHECK_CUDA(cudaStreamSynchronize(stream));
      std::vector<double> total;
      for (int t = 0; t < o->n_trials; ++t) {
        CHECK_CUDA(cudaEventRecord(ev.ev[0], stream));
        runAll();
        CHECK_CUDA(cudaEventRecord(ev.ev[1], stream));
        CHECK_CUDA(cudaStreamSynchronize(stream));
        float ms = 0;
        CHECK_CUDA(cudaEventElapsedTime(&ms, ev.ev[0], ev.ev[1]));
        total.push_back(ms);
      }
      Stats st = reduceTimes(h, total);
      if (h->rank == 0)
        std::printf("CUDECOMP:\tgrid: %d x %d, halo backend: %s \n"
                    "CUDECOMP:\tTotal time min/max/avg/std [ms]: %f/%f/%f/%f\n",
                    p[0], p[1], cudecompHaloCommBackendToString(static_cast<cudecompHaloCommBackend_t>(backend)),
                    st.min, st.max, st.avg, st.std);
      if (st.avg < t_best) {
        t_best = st.avg;
        best_grid = p;
      }
    }
    CHECK_CUDA(cudaDeviceSynchronize());
    h->peers.clear();
    barrier(*h->comm);
  }
  if (!valid) THROW_NOT_SUPPORTED("No valid decomposition found during autotuning with provided arguments.");
  gd->config.pdims[0] = best_grid[0];
  gd->config.pdims[1] = best_grid[1];
  gd->config.halo_comm_backend = static_cast<cudecompHaloCommBackend_t>(backend);
  if (h->rank == 0)
    std::printf("CUDECOMP: SELECTED: grid: %d x %d, halo backend: %s, Avg. time [ms]: %f\n", best_grid[0], best_grid[1],
                cudecompHaloCommBackendToString(gd->config.halo_comm_backend), h->have_device ? t_best : 0.0);
  barrier(*h->comm);
  if (h->rank == 0) std::printf("CUDECOMP: halo autotuning time [s]: %f\n", MPI_Wtime() - t_start);
}

} // namespace

void autotune(cudecompHandle_t h, cudecompGridDesc_t gd, const cudecompGridDescAutotuneOptions_t* o) {
  if (!o) THROW_INVALID_USAGE("options argument cannot be null if autotuning");
  const bool tune_pdims = (gd->config.pdims[0] == 0 && gd->config.pdims[1] == 0);
  const bool tune_t = o->autotune_transpose_backend;
  const bool tune_h = o->autotune_halo_backend;
  if (o->n_trials < 1 || o->n_warmup_trials < 0) THROW_INVALID_USAGE("invalid number of autotuning trials");
  // the grid is chosen by the first pass; the second pass (if any) tunes its backend on that grid
  // (reference src/cudecomp.cc:1201-1211)
  if (o->grid_mode == CUDECOMP_AUTOTUNE_GRID_TRANSPOSE) {
    if (tune_t || tune_pdims) autotuneTransposes(h, gd, o, tune_pdims, tune_t);
    if (tune_h) autotuneHalos(h, gd, o, false, true);
  } else if (o->grid_mode == CUDECOMP_AUTOTUNE_GRID_HALO) {
    if (tune_h || tune_pdims) autotuneHalos(h, gd, o, tune_pdims, tune_h);
    if (tune_t) autotuneTransposes(h, gd, o, false, true);
  } else {
    THROW_INVALID_USAGE("unknown value of autotune_grid_mode encountered.");
  }
  std::fflush(stdout);
}

} // namespace cdb
