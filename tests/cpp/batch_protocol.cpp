// C++ host-side test and example of the batched entry points (ddlo_batch_*, BASELINE config C5): no Python in the
// data path.  Independent scan-to-scan units (source = scan f+1, target = scan f, cycled) are sharded contiguously
// over the visible devices, one batch and one host thread per device, exactly as one process per GPU would do it;
// every unit's result is then recomputed by a single engine (ddlo_gicp_align with the same align-block limit) and must
// match bit for bit.
//
//   batch_protocol scans.bin units lanes host_threads [max_devices [waves|lanes [wave_units]]]
//   scans.bin: int32 count, then per scan int32 n and n*4 float32 (x y z 1)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/ddlo_gicp.h"

struct Scan {
  std::vector<float> xyzw;
  int n = 0;
};

static std::vector<Scan> load(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) {
    std::perror(path);
    std::exit(2);
  }
  int count = 0;
  if (std::fread(&count, 4, 1, f) != 1) std::exit(2);
  std::vector<Scan> scans(count);
  for (auto& s : scans) {
    if (std::fread(&s.n, 4, 1, f) != 1) std::exit(2);
    s.xyzw.resize(4 * (size_t)s.n);
    if (std::fread(s.xyzw.data(), 4, s.xyzw.size(), f) != s.xyzw.size()) std::exit(2);
  }
  std::fclose(f);
  return scans;
}

#define CHECK(expr)                                                                         \
  do {                                                                                      \
    int rc__ = (expr);                                                                      \
    if (rc__ != DDLO_OK) {                                                                  \
      std::fprintf(stderr, "%s failed (%d): %s\n", #expr, rc__, ddlo_last_error());         \
      std::exit(1);                                                                         \
    }                                                                                       \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: %s scans.bin units lanes host_threads [max_devices [waves|lanes [wave_units]]]\n", argv[0]);
    return 2;
  }
  const std::vector<Scan> scans = load(argv[1]);
  const int units = std::atoi(argv[2]), lanes = std::atoi(argv[3]), host_threads = std::atoi(argv[4]);
  int n_dev = 0;
  CHECK(ddlo_device_count(&n_dev));
  if (argc > 5) n_dev = std::min(n_dev, std::atoi(argv[5]));
  const bool waves = !(argc > 6 && std::strcmp(argv[6], "lanes") == 0);
  const int wave_units = argc > 7 ? std::atoi(argv[7]) : 0;
  if (n_dev < 1 || scans.size() < 2) return 2;
  const int n_pairs = (int)scans.size() - 1;

  std::vector<ddlo_align_result> results(units);
  std::vector<int> unit_dev(units, 0), blocks_of_dev(n_dev, 0);
  std::vector<double> seconds(n_dev, 0.0);
  auto shard = [&](int d, int& begin, int& end) {  // contiguous shards, sizes differ by at most one (sharding.shard_range)
    const int base = units / n_dev, extra = units % n_dev;
    begin = d * base + std::min(d, extra);
    end = begin + base + (d < extra ? 1 : 0);
  };
  auto run_device = [&](int d) {
    int begin, end;
    shard(d, begin, end);
    ddlo_batch* b = nullptr;
    CHECK(ddlo_batch_create(d, lanes, 0, host_threads, &b));
    CHECK(ddlo_batch_set_mode(b, waves ? DDLO_BATCH_WAVES : DDLO_BATCH_LANES, wave_units));
    CHECK(ddlo_batch_info(b, nullptr, &blocks_of_dev[d], nullptr));
    if (waves) blocks_of_dev[d] = 0;  // the batched kernels reproduce an ordinary engine's full-size align
    std::vector<int> ids(scans.size());
    for (size_t s = 0; s < scans.size(); ++s) CHECK(ddlo_batch_stage_cloud(b, scans[s].xyzw.data(), scans[s].n, 16, &ids[s]));
    std::vector<ddlo_batch_job> jobs(end - begin);
    for (int u = begin; u < end; ++u) {
      ddlo_batch_job& j = jobs[u - begin];
      j.source = ids[u % n_pairs + 1];
      j.target = ids[u % n_pairs];
      std::memset(j.guess, 0, sizeof(j.guess));
      j.guess[0] = j.guess[5] = j.guess[10] = j.guess[15] = 1.0f;
      unit_dev[u] = d;
    }
    const int warm = std::min<int>((int)jobs.size(), 2 * lanes);
    std::vector<ddlo_align_result> scratch(warm);
    CHECK(ddlo_batch_run(b, jobs.data(), warm, scratch.data()));
    const auto t0 = std::chrono::steady_clock::now();
    CHECK(ddlo_batch_submit(b, jobs.data(), (int)jobs.size(), results.data() + begin));
    CHECK(ddlo_batch_wait(b));
    seconds[d] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    CHECK(ddlo_batch_destroy(b));
  };
  std::vector<std::thread> th;
  for (int d = 1; d < n_dev; ++d) th.emplace_back(run_device, d);
  run_device(0);
  for (auto& t : th) t.join();

  // the same units, one at a time, on an ordinary engine of the same device with the same block limit
  int mismatches = 0;
  for (int d = 0; d < n_dev; ++d) {
    ddlo_runtime* rt = nullptr;
    ddlo_gicp* g = nullptr;
    CHECK(ddlo_runtime_create(d, &rt));
    CHECK(ddlo_runtime_set_align_blocks(rt, blocks_of_dev[d]));
    CHECK(ddlo_gicp_create(rt, &g));
    std::vector<ddlo_align_result> ref(n_pairs);
    for (int p = 0; p < std::min(n_pairs, units); ++p) {
      ddlo_cloud *s = nullptr, *t = nullptr;
      CHECK(ddlo_cloud_create(rt, scans[p + 1].xyzw.data(), scans[p + 1].n, 16, &s));
      CHECK(ddlo_cloud_create(rt, scans[p].xyzw.data(), scans[p].n, 16, &t));
      CHECK(ddlo_gicp_set_input_source(g, s, 1));
      CHECK(ddlo_gicp_set_input_target(g, t));
      CHECK(ddlo_gicp_align(g, nullptr, &ref[p]));
      CHECK(ddlo_cloud_release(s));
      CHECK(ddlo_cloud_release(t));
    }
    for (int u = 0; u < units; ++u) {
      if (unit_dev[u] != d) continue;
      const ddlo_align_result &a = results[u], &r = ref[u % n_pairs];
      if (std::memcmp(a.final_transformation, r.final_transformation, sizeof(a.final_transformation)) != 0 ||
          std::memcmp(a.final_hessian, r.final_hessian, sizeof(a.final_hessian)) != 0 || a.nr_iterations != r.nr_iterations || a.flags != r.flags)
        ++mismatches;
    }
    CHECK(ddlo_gicp_destroy(g));
    CHECK(ddlo_runtime_destroy(rt));
  }
  double tmax = 0.0;
  for (double s : seconds) tmax = std::max(tmax, s);
  for (int u = 0; u < units; ++u) {
    std::printf("unit %d %d %d %d", u, unit_dev[u], results[u].flags & DDLO_FLAG_CONVERGED ? 1 : 0, results[u].nr_iterations);
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) std::printf(" %.9g", results[u].final_transformation[4 * c + r]);
    std::printf("\n");
  }
  std::printf("summary mode %s devices %d units %d lanes %d host_threads %d seconds %.6f units_per_s %.1f mismatches %d\n", waves ? "waves" : "lanes",
              n_dev, units, lanes, host_threads, tmax, units / tmax, mismatches);
  return mismatches == 0 ? 0 : 1;
}
