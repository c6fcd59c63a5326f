// See perf_report.h.
#include "perf_report.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <sstream>

#include "engine.h"

namespace cdb {

namespace {

int envInt(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

std::string fmt3(const int32_t* a) {
  std::ostringstream os;
  os << "[" << (a ? a[0] : 0) << "," << (a ? a[1] : 0) << "," << (a ? a[2] : 0) << "]";
  return os.str();
}

std::string fmtBool3(const bool* a) {
  std::ostringstream os;
  os << "[" << ((a && a[0]) ? 1 : 0) << "," << ((a && a[1]) ? 1 : 0) << "," << ((a && a[2]) ? 1 : 0) << "]";
  return os.str();
}

const char* dtypeLetter(cudecompDataType_t d) {
  switch (d) {
  case CUDECOMP_FLOAT: return "S";
  case CUDECOMP_DOUBLE: return "D";
  case CUDECOMP_FLOAT_COMPLEX: return "C";
  case CUDECOMP_DOUBLE_COMPLEX: return "Z";
  }
  return "unknown";
}

struct Row {
  const PerfSeries* series;
  double count = 0, total = 0, exch = 0, local = 0; // sums over samples, then over ranks
  // per-sample values of this rank in ring order (CUDECOMP_PERFORMANCE_REPORT_DETAIL > 0), 4 floats per sample:
  // total, exchange, local [ms], exchange bandwidth [GB/s]
  std::vector<float> samples;
  // detail level 2: every rank's samples, `slots` samples per rank, unused slots hold a negative total
  std::vector<float> all_samples;
  int slots = 0;
};

void printSamples(const PerfSettings& s_, cudecompHandle* h, cudecompGridDesc* gd, const std::vector<Row>& rows, bool transposes);

std::string csvName(const cudecompGridDesc* gd, const char* kind) {
  std::ostringstream f;
  f << "cudecomp-perf-report-" << kind << "-aggregated-tcomm_" << gd->config.transpose_comm_backend << "-hcomm_"
    << gd->config.halo_comm_backend << "-pdims_" << gd->config.pdims[0] << "x" << gd->config.pdims[1] << "-gdims_"
    << gd->config.gdims[0] << "x" << gd->config.gdims[1] << "x" << gd->config.gdims[2] << "-memorder_";
  for (int ax = 0; ax < 3; ++ax)
    for (int i = 0; i < 3; ++i) f << gd->config.transpose_mem_order[ax][i];
  f << ".csv";
  return f.str();
}

void csvHeader(std::ofstream& file, const cudecompGridDesc* gd) {
  file << "# Transpose backend: " << cudecompTransposeCommBackendToString(gd->config.transpose_comm_backend) << "\n";
  file << "# Halo backend: " << cudecompHaloCommBackendToString(gd->config.halo_comm_backend) << "\n";
  file << "# Process grid: [" << gd->config.pdims[0] << ", " << gd->config.pdims[1] << "]\n";
  file << "# Global dimensions: [" << gd->config.gdims[0] << ", " << gd->config.gdims[1] << ", " << gd->config.gdims[2]
       << "]\n";
  file << "# Memory order: ";
  for (int ax = 0; ax < 3; ++ax) {
    file << "[" << gd->config.transpose_mem_order[ax][0] << "," << gd->config.transpose_mem_order[ax][1] << ","
         << gd->config.transpose_mem_order[ax][2] << "]";
    if (ax < 2) file << "; ";
  }
  file << "\n#\n";
}

} // namespace

void PerfSettings::readEnvironment() {
  enabled = envInt("CUDECOMP_ENABLE_PERFORMANCE_REPORT", 0) != 0;
  detail = envInt("CUDECOMP_PERFORMANCE_REPORT_DETAIL", 0);
  samples = std::max(1, envInt("CUDECOMP_PERFORMANCE_REPORT_SAMPLES", 20));
  warmup = std::max(0, envInt("CUDECOMP_PERFORMANCE_REPORT_WARMUP_SAMPLES", 3));
  const char* dir = std::getenv("CUDECOMP_PERFORMANCE_REPORT_WRITE_DIR");
  write_dir = dir ? dir : "";
}

PerfReport::~PerfReport() {
  for (auto* m : {&transposes_, &halos_})
    for (auto& kv : *m)
      for (auto& s : kv.second.ring) {
        if (s.start) cudaEventDestroy(s.start);
        if (s.mid) cudaEventDestroy(s.mid);
        if (s.end) cudaEventDestroy(s.end);
      }
  (void)cudaGetLastError();
}

PerfSample* PerfReport::next(PerfSeries& series, cudaStream_t stream) {
  const int64_t call = series.seen++;
  if (call < s_.warmup) return nullptr;
  if (series.ring.empty()) {
    series.ring.resize(static_cast<size_t>(s_.samples));
    for (auto& s : series.ring) {
      cudaEventCreate(&s.start);
      cudaEventCreate(&s.mid);
      cudaEventCreate(&s.end);
    }
  }
  PerfSample& s = series.ring[static_cast<size_t>((call - s_.warmup) % s_.samples)];
  s.has_mid = false;
  s.exchange = false;
  s.valid = false;
  cudaEventRecord(s.start, stream);
  return &s;
}

PerfSample* PerfReport::beginTranspose(int ax, int dir, cudecompDataType_t dtype, const int32_t* ih, const int32_t* oh,
                                       const int32_t* ip, const int32_t* op, bool inplace, bool managed, int64_t bytes,
                                       cudaStream_t stream) {
  static const char* names[4] = {"TransposeXY", "TransposeYZ", "TransposeZY", "TransposeYX"};
  const int opi = (ax == 0) ? 0 : (ax == 2 ? 2 : (dir > 0 ? 1 : 3));
  std::ostringstream key;
  key << opi << "|" << -static_cast<int>(dtype) << "|" << fmt3(ih) << fmt3(oh) << fmt3(ip) << fmt3(op) << inplace << managed;
  PerfSeries& series = transposes_[key.str()];
  if (series.operation.empty()) {
    series.operation = names[opi];
    series.dtype = dtypeLetter(dtype);
    series.halos_a = fmt3(ih);
    series.halos_b = fmt3(oh);
    series.pads_a = fmt3(ip);
    series.pads_b = fmt3(op);
    series.flag_a = inplace ? "Y" : "N";
    series.flag_b = managed ? "Y" : "N";
    series.bytes = bytes;
  }
  return next(series, stream);
}

PerfSample* PerfReport::beginHalo(int ax, int dim, cudecompDataType_t dtype, const int32_t* halo, const bool* periods,
                                  const int32_t* pad, bool managed, int64_t bytes, cudaStream_t stream) {
  static const char* names[3] = {"HaloX", "HaloY", "HaloZ"};
  std::ostringstream key;
  key << ax << "|" << dim << "|" << -static_cast<int>(dtype) << "|" << fmt3(halo) << fmtBool3(periods) << fmt3(pad) << managed;
  PerfSeries& series = halos_[key.str()];
  if (series.operation.empty()) {
    series.operation = names[ax];
    series.dtype = dtypeLetter(dtype);
    series.dim = dim;
    series.halos_a = fmt3(halo);
    series.flag_a = fmtBool3(periods);
    series.pads_a = fmt3(pad);
    series.flag_b = managed ? "Y" : "N";
    series.bytes = bytes;
  }
  return next(series, stream);
}

void PerfReport::print(cudecompHandle* h, cudecompGridDesc* gd) {
  cudaDeviceSynchronize();
  auto collect = [&](std::map<std::string, PerfSeries>& m) {
    std::vector<Row> rows;
    for (auto& kv : m) {
      Row r;
      r.series = &kv.second;
      for (auto& s : kv.second.ring) {
        if (!s.valid) continue;
        float total = 0, first = 0;
        if (cudaEventElapsedTime(&total, s.start, s.end) != cudaSuccess) continue;
        double exch = 0, local = total;
        if (s.exchange) {
          if (s.has_mid && cudaEventElapsedTime(&first, s.start, s.mid) == cudaSuccess) {
            exch = first;
            local = total - first;
          } else {
            exch = total;
            local = 0;
          }
        }
        r.count += 1;
        r.total += total;
        r.exch += exch;
        r.local += local;
        const double bw = exch > 0 ? static_cast<double>(kv.second.bytes) / 1e6 / exch : 0.0;
        r.samples.insert(r.samples.end(), {total, static_cast<float>(exch), static_cast<float>(local), static_cast<float>(bw)});
      }
      rows.push_back(r);
    }
    (void)cudaGetLastError();
    // every rank must hold the same configurations (calls are collective); otherwise report this rank alone
    double n[2] = {static_cast<double>(rows.size()), -static_cast<double>(rows.size())};
    allreduceF64(*h->comm, n, 2, ReduceOp::MAX);
    const bool same = (n[0] == -n[1]);
    if (same && !rows.empty()) {
      std::vector<double> buf;
      for (auto& r : rows) {
        buf.push_back(r.count);
        buf.push_back(r.total);
        buf.push_back(r.exch);
        buf.push_back(r.local);
      }
      allreduceF64(*h->comm, buf.data(), static_cast<int>(buf.size()), ReduceOp::SUM);
      for (size_t i = 0; i < rows.size(); ++i) {
        rows[i].count = buf[4 * i];
        rows[i].total = buf[4 * i + 1];
        rows[i].exch = buf[4 * i + 2];
        rows[i].local = buf[4 * i + 3];
      }
    } else if (!same && h->rank == 0) {
      std::printf("CUDECOMP:WARN: ranks recorded different operation sets; the report shows rank 0 only\n");
    }
    if (same && s_.detail >= 2) {
      // every rank contributes `samples` slots per configuration (reference: gatherSampleData, src/performance.cc)
      for (auto& r : rows) {
        r.slots = s_.samples;
        std::vector<float> mine(static_cast<size_t>(r.slots) * 4, -1.0f);
        std::copy(r.samples.begin(), r.samples.begin() + std::min(r.samples.size(), mine.size()), mine.begin());
        r.all_samples.resize(mine.size() * static_cast<size_t>(h->nranks));
        allgather(*h->comm, mine.data(), mine.size() * sizeof(float), r.all_samples.data());
      }
    }
    return rows;
  };
  std::vector<Row> trows = collect(transposes_);
  std::vector<Row> hrows = collect(halos_);
  if (h->rank != 0) return;

  std::printf("CUDECOMP:\nCUDECOMP: ===== Performance Summary =====\nCUDECOMP: Grid Configuration:\n");
  std::printf("CUDECOMP:\tTranspose backend: %s\n", cudecompTransposeCommBackendToString(gd->config.transpose_comm_backend));
  std::printf("CUDECOMP:\tHalo backend: %s\n", cudecompHaloCommBackendToString(gd->config.halo_comm_backend));
  std::printf("CUDECOMP:\tProcess grid: [%d, %d]\n", gd->config.pdims[0], gd->config.pdims[1]);
  std::printf("CUDECOMP:\tGlobal dimensions: [%d, %d, %d]\n", gd->config.gdims[0], gd->config.gdims[1], gd->config.gdims[2]);
  std::printf("CUDECOMP:\tMemory order: ");
  for (int ax = 0; ax < 3; ++ax) {
    std::printf("[%d,%d,%d]", gd->config.transpose_mem_order[ax][0], gd->config.transpose_mem_order[ax][1],
                gd->config.transpose_mem_order[ax][2]);
    if (ax < 2) std::printf("; ");
  }
  std::printf("\nCUDECOMP:\n");

  auto avg = [](double sum, double n) { return n > 0 ? sum / n : 0.0; };
  auto bw = [&](const Row& r) {
    const double ms = avg(r.exch, r.count);
    return ms > 0 ? static_cast<double>(r.series->bytes) / 1e6 / ms : 0.0; // bytes / ms -> GB/s
  };

  if (!trows.empty()) {
    std::printf("CUDECOMP: Transpose Performance Data:\nCUDECOMP:\n");
    std::printf("CUDECOMP: %-12s %-6s %-15s %-15s %-8s %-8s %-8s %-9s %-9s %-9s %-9s\n", "operation", "dtype",
                "halo extents", "padding", "inplace", "managed", "samples", "total", "A2A", "local", "A2A BW");
    std::printf("CUDECOMP: %-12s %-6s %-15s %-15s %-8s %-8s %-8s %-9s %-9s %-9s %-9s\n", "", "", "", "", "", "", "", "[ms]",
                "[ms]", "[ms]", "[GB/s]");
    std::printf("CUDECOMP: %s\n", std::string(120, '-').c_str());
    for (auto& r : trows) {
      if (r.count <= 0) continue;
      const PerfSeries& s = *r.series;
      std::printf("CUDECOMP: %-12s %-6s %-7s/%-7s %-7s/%-7s %-8s %-8s %-8d %-9.3f %-9.3f %-9.3f %-9.3f\n", s.operation.c_str(),
                  s.dtype.c_str(), s.halos_a.c_str(), s.halos_b.c_str(), s.pads_a.c_str(), s.pads_b.c_str(),
                  s.flag_a.c_str(), s.flag_b.c_str(), static_cast<int>(r.count / h->nranks), avg(r.total, r.count),
                  avg(r.exch, r.count), avg(r.local, r.count), bw(r));
    }
  }
  if (!hrows.empty()) {
    std::printf("CUDECOMP:\nCUDECOMP: Halo Performance Data:\nCUDECOMP:\n");
    std::printf("CUDECOMP: %-12s %-6s %-5s %-12s %-12s %-12s %-8s %-8s %-9s %-9s %-9s %-9s\n", "operation", "dtype", "dim",
                "halo extent", "periods", "padding", "managed", "samples", "total", "SR", "local", "SR BW");
    std::printf("CUDECOMP: %-12s %-6s %-5s %-12s %-12s %-12s %-8s %-8s %-9s %-9s %-9s %-9s\n", "", "", "", "", "", "", "", "",
                "[ms]", "[ms]", "[ms]", "[GB/s]");
    std::printf("CUDECOMP: %s\n", std::string(120, '-').c_str());
    for (auto& r : hrows) {
      if (r.count <= 0) continue;
      const PerfSeries& s = *r.series;
      std::printf("CUDECOMP: %-12s %-6s %-5d %-12s %-12s %-12s %-8s %-8d %-9.3f %-9.3f %-9.3f %-9.3f\n", s.operation.c_str(),
                  s.dtype.c_str(), s.dim, s.halos_a.c_str(), s.flag_a.c_str(), s.pads_a.c_str(), s.flag_b.c_str(),
                  static_cast<int>(r.count / h->nranks), avg(r.total, r.count), avg(r.exch, r.count),
                  avg(r.local, r.count), bw(r));
    }
  }
  std::printf("CUDECOMP:\n");

  if (!s_.write_dir.empty()) {
    if (!trows.empty()) {
      const std::string path = s_.write_dir + "/" + csvName(gd, "transpose");
      std::ofstream file(path);
      if (!file) {
        std::printf("CUDECOMP:WARN: Could not open file %s for writing\n", path.c_str());
      } else {
        csvHeader(file, gd);
        file << "operation,dtype,input_halo_extents,output_halo_extents,input_padding,output_padding,inplace,managed,"
                "samples,total_ms,A2A_ms,local_ms,A2A_BW_GBps\n";
        file << std::fixed << std::setprecision(3);
        for (auto& r : trows) {
          if (r.count <= 0) continue;
          const PerfSeries& s = *r.series;
          file << s.operation << "," << s.dtype << ",\"" << s.halos_a << "\",\"" << s.halos_b << "\",\"" << s.pads_a
               << "\",\"" << s.pads_b << "\"," << s.flag_a << "," << s.flag_b << "," << static_cast<int>(r.count / h->nranks)
               << "," << avg(r.total, r.count) << "," << avg(r.exch, r.count) << "," << avg(r.local, r.count) << ","
               << bw(r) << "\n";
        }
        std::printf("CUDECOMP: Wrote transpose performance data to %s\n", path.c_str());
      }
    }
    if (!hrows.empty()) {
      const std::string path = s_.write_dir + "/" + csvName(gd, "halo");
      std::ofstream file(path);
      if (!file) {
        std::printf("CUDECOMP:WARN: Could not open file %s for writing\n", path.c_str());
      } else {
        csvHeader(file, gd);
        file << "operation,dtype,dim,halo_extent,periods,padding,managed,samples,total_ms,SR_ms,local_ms,SR_BW_GBps\n";
        file << std::fixed << std::setprecision(3);
        for (auto& r : hrows) {
          if (r.count <= 0) continue;
          const PerfSeries& s = *r.series;
          file << s.operation << "," << s.dtype << "," << s.dim << ",\"" << s.halos_a << "\",\"" << s.flag_a << "\",\""
               << s.pads_a << "\"," << s.flag_b << "," << static_cast<int>(r.count / h->nranks) << ","
               << avg(r.total, r.count) << "," << avg(r.exch, r.count) << "," << avg(r.local, r.count) << "," << bw(r)
               << "\n";
        }
        std::printf("CUDECOMP: Wrote halo performance data to %s\n", path.c_str());
      }
    }
  }
  if (s_.detail > 0) {
    std::printf("CUDECOMP:\nCUDECOMP: Per-Sample Details:\nCUDECOMP:\n");
    printSamples(s_, h, gd, trows, true);
    printSamples(s_, h, gd, hrows, false);
  }
  std::printf("CUDECOMP: ================================\nCUDECOMP:\n");
  std::fflush(stdout);
}

// Per-sample tables and CSV files (CUDECOMP_PERFORMANCE_REPORT_DETAIL 1: this rank's samples, 2: every rank's), in the
// reference's format (src/performance.cc:560-770). Rank 0 only.
namespace {
void printSamples(const PerfSettings& s_, cudecompHandle* h, cudecompGridDesc* gd, const std::vector<Row>& rows, bool transposes) {
  bool any = false;
  for (auto& r : rows)
    if (r.count > 0) any = true;
  if (!any) return;
  std::ofstream csv;
  std::string path;
  if (!s_.write_dir.empty()) {
    path = s_.write_dir + "/" + csvName(gd, transposes ? "transpose-samples" : "halo-samples");
    csv.open(path);
    if (!csv) {
      std::printf("CUDECOMP: Warning: Could not open file %s for writing\n", path.c_str());
    } else {
      csvHeader(csv, gd);
      csv << (transposes ? "operation,dtype,input_halo_extents,output_halo_extents,input_padding,output_padding,inplace,"
                           "managed,rank,sample,total_ms,A2A_ms,local_ms,A2A_BW_GBps\n"
                         : "operation,dtype,dim,halo_extent,periods,padding,managed,rank,sample,total_ms,SR_ms,local_ms,"
                           "SR_BW_GBps\n");
      csv << std::fixed << std::setprecision(3);
    }
  }
  for (auto& r : rows) {
    if (r.count <= 0) continue;
    const PerfSeries& s = *r.series;
    if (transposes)
      std::printf("CUDECOMP: %s (dtype=%s, halo extents=%s/%s, padding=%s/%s, inplace=%s, managed=%s) samples:\n",
                  s.operation.c_str(), s.dtype.c_str(), s.halos_a.c_str(), s.halos_b.c_str(), s.pads_a.c_str(),
                  s.pads_b.c_str(), s.flag_a.c_str(), s.flag_b.c_str());
    else
      std::printf("CUDECOMP: %s (dtype=%s, dim=%d, halos=%s, periods=%s, padding=%s, managed=%s) samples:\n",
                  s.operation.c_str(), s.dtype.c_str(), s.dim, s.halos_a.c_str(), s.flag_a.c_str(), s.pads_a.c_str(),
                  s.flag_b.c_str());
    const char* ex = transposes ? "A2A" : "SR";
    std::printf("CUDECOMP: %-6s %-12s %-9s %-9s %-9s %-9s\n", "rank", "sample", "total", ex, "local",
                (std::string(ex) + " BW").c_str());
    std::printf("CUDECOMP: %-6s %-12s %-9s %-9s %-9s %-9s\n", "", "", "[ms]", "[ms]", "[ms]", "[GB/s]");
    const bool all = s_.detail >= 2 && !r.all_samples.empty();
    const int nranks = all ? h->nranks : 1;
    for (int rank = 0; rank < nranks; ++rank) {
      const float* v = all ? r.all_samples.data() + static_cast<size_t>(rank) * r.slots * 4 : r.samples.data();
      const int n = all ? r.slots : static_cast<int>(r.samples.size() / 4);
      for (int k = 0; k < n; ++k) {
        if (v[4 * k] < 0) continue; // unused slot
        std::printf("CUDECOMP: %-6d %-12d %-9.3f %-9.3f %-9.3f %-9.3f\n", rank, k, v[4 * k], v[4 * k + 1], v[4 * k + 2],
                    v[4 * k + 3]);
        if (csv.is_open()) {
          csv << s.operation << "," << s.dtype << ",";
          if (transposes)
            csv << "\"" << s.halos_a << "\",\"" << s.halos_b << "\",\"" << s.pads_a << "\",\"" << s.pads_b << "\"," << s.flag_a
                << "," << s.flag_b;
          else
            csv << s.dim << ",\"" << s.halos_a << "\",\"" << s.flag_a << "\",\"" << s.pads_a << "\"," << s.flag_b;
          csv << "," << rank << "," << k << "," << v[4 * k] << "," << v[4 * k + 1] << "," << v[4 * k + 2] << "," << v[4 * k + 3]
              << "\n";
        }
      }
    }
    std::printf("CUDECOMP:\n");
  }
  if (csv.is_open()) {
    csv.close();
    std::printf("CUDECOMP:\nCUDECOMP: Wrote per-sample %s data to %s\n", transposes ? "transpose" : "halo", path.c_str());
  }
}
} // namespace

} // namespace cdb
