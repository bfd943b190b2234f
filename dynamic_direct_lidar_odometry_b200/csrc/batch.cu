// Batched registrations behind the C ABI (BASELINE config C5): many independent
// (source, target, guess) units on ONE device, driven from C++.
//
// The reference has no batched entry point: OdomNode calls one engine from one callback
// (src/odometry/odom.cc:745-793).  A single registration is 65k points of latency-bound work and leaves most
// of a B200 idle, so the batch owns S LANES, each a runtime (CUDA stream) + a NanoGICP engine whose align kernel
// is limited to num_SMs / S blocks: the cooperative launches of all lanes are resident side by side and the
// small index kernels of one unit overlap the searches of another.  Units are dealt to the lanes round-robin
// and every unit runs the whole path on its lane's stream without a host synchronisation:
//     fresh source (and target) handles -> index build(s) -> covariances -> LM align -> result copy
// Two ways to run the align stage (ddlo_batch_set_mode):
//   WAVES (default)  the units are taken W at a time.  The lanes prepare a wave (handles, indexes, covariances)
//                    while the previous wave is being aligned by the batched round kernels of batch_align.cu: three
//                    ordinary launches per LM round over ALL problems of the wave, 256-thread blocks scheduled freely
//                    over the SMs, no cooperative launch, no grid barrier.  A C++ driver thread owned by the batch
//                    feeds the rounds and polls one counter per round once the typical iteration count is through.
//   LANES            every unit's align is the single-registration kernel k_align (one cooperative launch) on its
//                    lane's stream, limited to num_SMs / S blocks so that the lanes' launches are resident side by side.
// In both modes the host only enqueues; `ddlo_batch_wait` is the synchronisation of the batch.  Results are
// bit-identical to a single engine's (WAVES: with the default block count; LANES: with the same align-block limit),
// because chunks, groups and summation orders are shared (gicp_dev.cuh).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "engine.cuh"

using namespace ddlo;

struct ddlo_batch {
  int device = 0;
  int align_blocks = 0;
  int host_threads = 1;
  struct Lane {
    ddlo_runtime* rt = nullptr;
    ddlo_gicp* eng = nullptr;
    bool has_shared_target = false;
  };
  std::vector<Lane> lanes;
  // WAVES mode: two wave buffers of W slots (engine on lane slot % S + its "prepared" event), a stream for the rounds
  int mode = DDLO_BATCH_WAVES;
  int wave_units = 64;
  struct Slot {
    ddlo_gicp* eng = nullptr;
    cudaEvent_t ready = nullptr;
    bool has_shared_target = false;
  };
  struct WaveBuf {
    std::vector<Slot> slots;
    void* d_probs = nullptr;
    void* h_probs = nullptr;  // pinned
    AlignOut* d_outs = nullptr;
    int* d_active = nullptr;
    int* h_active = nullptr;  // pinned
  };
  WaveBuf wb[2];
  ddlo_runtime* align_rt = nullptr;
  std::thread driver;
  int driver_rc = DDLO_OK;
  std::string driver_err;
  long long wave_rounds = 0, wave_polls = 0;
  std::vector<ddlo_cloud*> staged;  // inputs resident in HBM (owned by lane 0's runtime), index-less
  bool staging_dirty = false;
  ddlo_cloud* shared_tgt = nullptr;
  ddlo_covs* shared_cov = nullptr;
  // the submission in flight
  AlignOut* h_out = nullptr;  // pinned, one slot per unit
  size_t h_cap = 0;
  std::vector<int> covs_computed;
  std::vector<int> unit_rc;
  ddlo_align_result* results = nullptr;
  int pending = 0;
  // sources that come from HOST memory with the submission (ddlo_batch_submit_host), else null
  const float* const* host_src = nullptr;
  const int* host_n = nullptr;
  int host_stride = 0;
};

static void waves_free(ddlo_batch* b) {
  for (auto& w : b->wb) {
    for (auto& s : w.slots) {
      if (s.eng) ddlo_gicp_destroy(s.eng);
      if (s.ready) cudaEventDestroy(s.ready);
    }
    w.slots.clear();
    if (w.d_probs) cudaFree(w.d_probs);
    if (w.h_probs) cudaFreeHost(w.h_probs);
    if (w.d_outs) cudaFree(w.d_outs);
    if (w.d_active) cudaFree(w.d_active);
    if (w.h_active) cudaFreeHost(w.h_active);
    w = ddlo_batch::WaveBuf();
  }
}

static void batch_free(ddlo_batch* b) {
  if (!b) return;
  if (b->driver.joinable()) b->driver.join();
  cudaSetDevice(b->device);
  for (auto& l : b->lanes)
    if (l.rt) cudaStreamSynchronize(l.rt->stream);
  if (b->align_rt) cudaStreamSynchronize(b->align_rt->stream);
  waves_free(b);
  if (b->align_rt) ddlo_runtime_destroy(b->align_rt);
  for (auto& l : b->lanes)
    if (l.eng) ddlo_gicp_destroy(l.eng);
  for (ddlo_cloud* c : b->staged) ddlo_cloud_release(c);
  if (b->shared_cov) ddlo_covs_release(b->shared_cov);
  if (b->shared_tgt) ddlo_cloud_release(b->shared_tgt);
  for (auto& l : b->lanes)
    if (l.rt) ddlo_runtime_destroy(l.rt);
  if (b->h_out) cudaFreeHost(b->h_out);
  delete b;
}

// one unit, enqueued on its lane's stream; nothing here waits for the device.  args == nullptr: the unit's align is
// enqueued too (k_align, LANES mode); else only the preparation is, and *args / *nchunks describe the align to be run.
static int enqueue_unit(ddlo_batch* b, ddlo_runtime* rt, ddlo_gicp* g, bool& has_shared_target, const ddlo_batch_job& job, int slot,
                        GicpArgs* args, int* nchunks) {
  const int n_staged = (int)b->staged.size();
  const float* host = b->host_src ? b->host_src[slot] : nullptr;
  if (!host && (job.source < 0 || job.source >= n_staged)) return fail(DDLO_E_INVALID, "batch job: unknown source cloud id");
  if (job.target >= n_staged) return fail(DDLO_E_INVALID, "batch job: unknown target cloud id");
  if (job.target < 0 && !b->shared_tgt) return fail(DDLO_E_NOT_READY, "batch job: target < 0 needs ddlo_batch_set_shared_target");
  ddlo_cloud *src = nullptr, *tgt = nullptr;
  // fresh handles: every unit builds its own indexes and covariances, nothing is cached from an earlier unit
  if (host)  // the scan crosses PCIe inside the unit (asynchronous from page-locked memory)
    DDLO_TRY(ddlo_cloud_create(rt, host, b->host_n[slot], b->host_stride, &src));
  else
    DDLO_TRY(ddlo_cloud_create_from_device(rt, b->staged[job.source]->pts, b->staged[job.source]->n, &src));
  int rc = DDLO_OK;
  if (job.target >= 0) rc = ddlo_cloud_create_from_device(rt, b->staged[job.target]->pts, b->staged[job.target]->n, &tgt);
  if (rc == DDLO_OK) rc = ddlo_gicp_clear_source(g);
  if (rc == DDLO_OK) rc = ddlo_gicp_set_input_source(g, src, 1);
  if (rc == DDLO_OK) {
    if (job.target >= 0) {
      rc = ddlo_gicp_set_input_target(g, tgt);
      has_shared_target = false;
    } else if (!has_shared_target) {
      rc = ddlo_gicp_set_input_target(g, b->shared_tgt);
      if (rc == DDLO_OK) rc = ddlo_gicp_set_target_covariances(g, b->shared_cov);
      has_shared_target = rc == DDLO_OK;
    }
  }
  if (args) {
    if (rc == DDLO_OK) rc = prepare_align(g, job.guess, &b->covs_computed[slot], args, nchunks);
  } else {
    if (rc == DDLO_OK) rc = enqueue_align(g, job.guess, &b->covs_computed[slot]);
    if (rc == DDLO_OK && cudaMemcpyAsync(b->h_out + slot, g->d_out, sizeof(AlignOut), cudaMemcpyDeviceToHost, rt->stream) != cudaSuccess)
      rc = fail(DDLO_E_CUDA, "batch: result copy failed");
  }
  ddlo_cloud_release(src);  // the engine holds them until the lane's next unit
  if (tgt) ddlo_cloud_release(tgt);
  return rc;
}

// ---- WAVES mode -----------------------------------------------------------------------------------------------------
static int waves_setup(ddlo_batch* b) {
  const int W = b->wave_units;
  if (!b->align_rt) DDLO_TRY(ddlo_runtime_create(b->device, &b->align_rt));
  for (auto& w : b->wb) {
    if ((int)w.slots.size() == W) continue;
    if (!w.slots.empty()) return fail(DDLO_E_INVALID, "batch: wave size can not change once used");
    w.slots.resize(W);
    for (int s = 0; s < W; ++s) {
      DDLO_TRY(ddlo_gicp_create(b->lanes[s % b->lanes.size()].rt, &w.slots[s].eng));
      DDLO_TRY(ddlo_gicp_set_params(w.slots[s].eng, &b->lanes[0].eng->p));
      DDLO_CUDA(cudaEventCreateWithFlags(&w.slots[s].ready, cudaEventDisableTiming));
    }
    {
      // Grow the device's stream-ordered pool once, here: two waves of units keep their clouds, indexes, covariances
      // and workspaces alive at the same time (about 48 MB per 64x1024 scan-to-scan unit), and growing the pool in the
      // middle of a run costs tens of milliseconds per wave.  DDLO_BATCH_PREWARM_MB overrides (0 disables).
      size_t mb = (size_t)W * 48;
      if (const char* e = std::getenv("DDLO_BATCH_PREWARM_MB")) mb = (size_t)std::max(0L, std::atol(e));
      void* warm = nullptr;
      if (mb > 0 && cudaMallocAsync(&warm, mb << 20, b->lanes[0].rt->stream) == cudaSuccess)
        cudaFreeAsync(warm, b->lanes[0].rt->stream);
      else
        (void)cudaGetLastError();
    }
    DDLO_CUDA(cudaMalloc(&w.d_probs, batch_prob_bytes() * W));
    DDLO_CUDA(cudaMallocHost(&w.h_probs, batch_prob_bytes() * W));
    DDLO_CUDA(cudaMalloc(reinterpret_cast<void**>(&w.d_outs), sizeof(AlignOut) * W));
    DDLO_CUDA(cudaMalloc(reinterpret_cast<void**>(&w.d_active), sizeof(int)));
    DDLO_CUDA(cudaMallocHost(reinterpret_cast<void**>(&w.h_active), sizeof(int)));
  }
  return DDLO_OK;
}

// prepare the units [u0, u1) of a wave on the lanes' streams (handles, indexes, covariances, argument records)
static void wave_prepare(ddlo_batch* b, ddlo_batch::WaveBuf& w, const ddlo_batch_job* jobs, int u0, int u1, std::vector<std::string>& errs) {
  const int S = (int)b->lanes.size();
  const int T = std::min(b->host_threads, S);
  auto drive = [&](int t) {
    cudaSetDevice(b->device);
    for (int u = u0; u < u1; ++u) {
      const int slot = u - u0, lane = slot % S;
      if (lane % T != t) continue;
      ddlo_batch::Slot& sl = w.slots[slot];
      GicpArgs a;
      int nchunks = 0;
      int rc = enqueue_unit(b, b->lanes[lane].rt, sl.eng, sl.has_shared_target, jobs[u], u, &a, &nchunks);
      if (rc == DDLO_OK) {
        a.out = w.d_outs + slot;
        batch_prob_fill(w.h_probs, slot, &a, nchunks);
        if (cudaEventRecord(sl.ready, b->lanes[lane].rt->stream) != cudaSuccess) rc = fail(DDLO_E_CUDA, "batch: event record failed");
      }
      if (rc != DDLO_OK) {
        batch_prob_fill(w.h_probs, slot, nullptr, 0);
        if (errs[t].empty()) errs[t] = ddlo_last_error();
      }
      b->unit_rc[u] = rc;
    }
  };
  if (T <= 1) {
    drive(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(drive, t);
    drive(0);
    for (auto& x : th) x.join();
  }
}

// align the prepared wave: rounds of the batched kernels until every problem is done; blocks until then
static int wave_align(ddlo_batch* b, ddlo_batch::WaveBuf& w, int u0, int u1) {
  const int n = u1 - u0;
  cudaStream_t st = b->align_rt->stream;
  int max_chunks = 0;
  for (int s = 0; s < n; ++s)
    if (b->unit_rc[u0 + s] == DDLO_OK) DDLO_CUDA(cudaStreamWaitEvent(st, w.slots[s].ready, 0));
  DDLO_CUDA(cudaMemcpyAsync(w.d_probs, w.h_probs, batch_prob_bytes() * n, cudaMemcpyHostToDevice, st));
  max_chunks = b->lanes[0].rt->max_coop_blocks_align;  // no problem has more chunks than a full-size align has blocks
  DDLO_TRY(batch_align_begin(st, w.d_probs, n, w.d_active, &b->align_rt->launches));
  const ddlo_params& p = b->lanes[0].eng->p;
  const long long cap = (long long)std::max(p.max_iterations, 0) * (1 + std::max(p.lm_max_iterations, 0)) + 2;
  long long rounds = 0;
  // a registration of consecutive scans typically takes 3-5 outer iterations of one linearize + one accepted trial
  // each: enqueue that many rounds blind, then poll the number of unfinished problems after every further round
  for (int r = 0; r < 4 && rounds < cap; ++r, ++rounds) DDLO_TRY(batch_align_round(st, w.d_probs, n, max_chunks, w.d_active, &b->align_rt->launches));
  for (;;) {
    DDLO_CUDA(cudaMemcpyAsync(w.h_active, w.d_active, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDLO_CUDA(cudaStreamSynchronize(st));
    b->wave_polls += 1;
    if (*w.h_active <= 0 || rounds >= cap) break;
    DDLO_TRY(batch_align_round(st, w.d_probs, n, max_chunks, w.d_active, &b->align_rt->launches));
    ++rounds;
  }
  b->wave_rounds += rounds;
  DDLO_CUDA(cudaMemcpyAsync(b->h_out + u0, w.d_outs, sizeof(AlignOut) * n, cudaMemcpyDeviceToHost, st));
  DDLO_CUDA(cudaStreamSynchronize(st));
  if (*w.h_active > 0) return fail(DDLO_E_CUDA, "batch: problems still active after the round limit");
  return DDLO_OK;
}

static void waves_run(ddlo_batch* b, const ddlo_batch_job* jobs, int m) {
  cudaSetDevice(b->device);
  const int W = b->wave_units;
  const int nw = (m + W - 1) / W;
  std::vector<std::string> errs(std::max(1, b->host_threads));
  int rc = DDLO_OK;
  // DDLO_BATCH_TRACE=1: one line per wave on stderr (host time spent enqueueing the next wave's preparation, host time
  // until this wave's aligns were done, LM rounds so far); DDLO_BATCH_SKIP_ALIGN=1 (timing experiments only): the
  // aligns are not run at all, results are undefined
  static const bool trace = std::getenv("DDLO_BATCH_TRACE") != nullptr;
  static const bool skip_align = std::getenv("DDLO_BATCH_SKIP_ALIGN") != nullptr;
  using clk = std::chrono::steady_clock;
  auto ms = [](clk::time_point a, clk::time_point c) { return std::chrono::duration<double, std::milli>(c - a).count(); };
  wave_prepare(b, b->wb[0], jobs, 0, std::min(m, W), errs);
  for (int w = 0; w < nw && rc == DDLO_OK; ++w) {
    const auto t0 = clk::now();
    if (w + 1 < nw) wave_prepare(b, b->wb[(w + 1) & 1], jobs, (w + 1) * W, std::min(m, (w + 2) * W), errs);
    const auto t1 = clk::now();
    if (skip_align) {
      for (auto& l : b->lanes) cudaStreamSynchronize(l.rt->stream);
    } else {
      rc = wave_align(b, b->wb[w & 1], w * W, std::min(m, (w + 1) * W));
    }
    if (trace) std::fprintf(stderr, "[ddlo_batch] wave %d: prepare-next %.3f ms host, align %.3f ms, rounds so far %lld, polls %lld\n", w, ms(t0, t1), ms(t1, clk::now()), b->wave_rounds, b->wave_polls);
  }
  if (rc != DDLO_OK) {
    b->driver_rc = rc;
    b->driver_err = ddlo_last_error();
    return;
  }
  for (int i = 0; i < m; ++i)
    if (b->unit_rc[i] != DDLO_OK) {
      b->driver_rc = b->unit_rc[i];
      for (auto& e : errs)
        if (!e.empty()) b->driver_err = e;
      return;
    }
}

extern "C" {

int ddlo_batch_create(int device, int n_lanes, int align_blocks_per_lane, int host_threads, ddlo_batch** out) {
  if (!out) return fail(DDLO_E_INVALID, "out is null");
  *out = nullptr;
  if (n_lanes < 1 || n_lanes > 64 || host_threads < 0 || host_threads > 64) return fail(DDLO_E_INVALID, "batch: 1..64 lanes, 0..64 host threads");
  ddlo_batch* b = new (std::nothrow) ddlo_batch();
  if (!b) return fail(DDLO_E_INVALID, "out of host memory");
  b->device = device;
  b->host_threads = std::max(1, std::min(host_threads, n_lanes));
  b->lanes.resize(n_lanes);
  for (auto& l : b->lanes) {
    int rc = ddlo_runtime_create(device, &l.rt);
    if (rc == DDLO_OK) rc = ddlo_gicp_create(l.rt, &l.eng);
    if (rc != DDLO_OK) {
      const std::string why = ddlo_last_error();
      batch_free(b);
      return fail(rc, why);
    }
  }
  const int sms = b->lanes[0].rt->num_sms;
  b->align_blocks = std::min(align_blocks_per_lane > 0 ? align_blocks_per_lane : std::max(1, sms / n_lanes), b->lanes[0].rt->max_coop_blocks_align);
  // the per-lane block limit belongs to LANES mode; in WAVES mode (the default) a problem is cut into as many chunks
  // as an ordinary engine's align has blocks
  *out = b;
  return DDLO_OK;
}

int ddlo_batch_destroy(ddlo_batch* b) {
  batch_free(b);
  return DDLO_OK;
}

int ddlo_batch_info(const ddlo_batch* b, int* n_lanes, int* align_blocks_per_lane, int* host_threads) {
  if (!b) return fail(DDLO_E_INVALID, "batch is null");
  if (n_lanes) *n_lanes = (int)b->lanes.size();
  if (align_blocks_per_lane) *align_blocks_per_lane = b->align_blocks;
  if (host_threads) *host_threads = b->host_threads;
  return DDLO_OK;
}

int ddlo_batch_set_params(ddlo_batch* b, const ddlo_params* p) {
  if (!b || !p) return fail(DDLO_E_INVALID, "null argument");
  if (b->pending) return fail(DDLO_E_NOT_READY, "batch: a submission is in flight");
  for (auto& l : b->lanes) DDLO_TRY(ddlo_gicp_set_params(l.eng, p));
  for (auto& w : b->wb)
    for (auto& sl : w.slots) DDLO_TRY(ddlo_gicp_set_params(sl.eng, p));
  return DDLO_OK;
}

int ddlo_batch_stage_cloud(ddlo_batch* b, const float* xyz, int n, int stride_bytes, int* id) {
  if (!b || !id) return fail(DDLO_E_INVALID, "null argument");
  if (b->pending) return fail(DDLO_E_NOT_READY, "batch: a submission is in flight");
  ddlo_cloud* c = nullptr;
  DDLO_TRY(ddlo_cloud_create(b->lanes[0].rt, xyz, n, stride_bytes, &c));
  b->staged.push_back(c);
  b->staging_dirty = true;
  *id = (int)b->staged.size() - 1;
  return DDLO_OK;
}

int ddlo_batch_staged_count(const ddlo_batch* b, int* count) {
  if (!b || !count) return fail(DDLO_E_INVALID, "null argument");
  *count = (int)b->staged.size();
  return DDLO_OK;
}

int ddlo_batch_set_shared_target(ddlo_batch* b, int cloud_id, const double* covs_mat4x4) {
  if (!b) return fail(DDLO_E_INVALID, "batch is null");
  if (b->pending) return fail(DDLO_E_NOT_READY, "batch: a submission is in flight");
  if (cloud_id < 0 || cloud_id >= (int)b->staged.size()) return fail(DDLO_E_INVALID, "batch: unknown cloud id");
  ddlo_runtime* rt = b->lanes[0].rt;
  for (auto& l : b->lanes) {  // nobody may still read the previous shared target
    DDLO_TRY(ddlo_runtime_synchronize(l.rt));
    if (l.has_shared_target) DDLO_TRY(ddlo_gicp_clear_target(l.eng));
    l.has_shared_target = false;
  }
  for (auto& w : b->wb)
    for (auto& sl : w.slots) {
      if (sl.has_shared_target) DDLO_TRY(ddlo_gicp_clear_target(sl.eng));
      sl.has_shared_target = false;
    }
  if (b->shared_cov) ddlo_covs_release(b->shared_cov);
  if (b->shared_tgt) ddlo_cloud_release(b->shared_tgt);
  b->shared_cov = nullptr, b->shared_tgt = nullptr;
  ddlo_cloud* t = nullptr;
  ddlo_covs* v = nullptr;
  DDLO_TRY(ddlo_cloud_create_from_device(rt, b->staged[cloud_id]->pts, b->staged[cloud_id]->n, &t));
  int rc = ddlo_cloud_build_index(t);
  if (rc == DDLO_OK) {
    ddlo_params p;
    ddlo_gicp_get_params(b->lanes[0].eng, &p);
    rc = covs_mat4x4 ? ddlo_covs_from_host(rt, covs_mat4x4, t->n, &v) : ddlo_covs_compute(t, p.k_correspondences, p.regularization_method, &v);
  }
  if (rc == DDLO_OK) rc = ddlo_cloud_share(t);
  if (rc == DDLO_OK) rc = ddlo_covs_share(v, t);
  if (rc != DDLO_OK) {
    if (v) ddlo_covs_release(v);
    ddlo_cloud_release(t);
    return rc;
  }
  b->shared_tgt = t;
  b->shared_cov = v;
  return DDLO_OK;
}

static int batch_submit(ddlo_batch* b, const ddlo_batch_job* jobs, int m, ddlo_align_result* results);

int ddlo_batch_submit(ddlo_batch* b, const ddlo_batch_job* jobs, int m, ddlo_align_result* results) {
  if (b) b->host_src = nullptr, b->host_n = nullptr, b->host_stride = 0;
  return batch_submit(b, jobs, m, results);
}

int ddlo_batch_submit_host(ddlo_batch* b, const ddlo_batch_job* jobs, int m, const float* const* source_xyz, const int* source_n, int stride_bytes,
                           ddlo_align_result* results) {
  if (!b || (m > 0 && (!source_xyz || !source_n))) return fail(DDLO_E_INVALID, "null argument");
  if (stride_bytes < 12 || (stride_bytes & 3)) return fail(DDLO_E_INVALID, "bad stride");
  if (b->pending) return fail(DDLO_E_NOT_READY, "batch: the previous submission has not been waited for");
  b->host_src = source_xyz;
  b->host_n = source_n;
  b->host_stride = stride_bytes;
  return batch_submit(b, jobs, m, results);
}

static int batch_submit(ddlo_batch* b, const ddlo_batch_job* jobs, int m, ddlo_align_result* results) {
  if (!b || m < 0 || (m > 0 && (!jobs || !results))) return fail(DDLO_E_INVALID, "null argument");
  if (b->pending) return fail(DDLO_E_NOT_READY, "batch: the previous submission has not been waited for");
  if (m == 0) return DDLO_OK;
  DDLO_CUDA(cudaSetDevice(b->device));
  if (b->staging_dirty) {  // uploads run on lane 0's stream; every lane reads them
    DDLO_CUDA(cudaStreamSynchronize(b->lanes[0].rt->stream));
    b->staging_dirty = false;
  }
  if (b->h_cap < (size_t)m) {
    if (b->h_out) DDLO_CUDA(cudaFreeHost(b->h_out));
    b->h_out = nullptr, b->h_cap = 0;
    const size_t cap = std::max<size_t>(m, 64);
    DDLO_CUDA(cudaMallocHost(reinterpret_cast<void**>(&b->h_out), cap * sizeof(AlignOut)));
    b->h_cap = cap;
  }
  b->covs_computed.assign(m, 0);
  b->unit_rc.assign(m, DDLO_OK);
  b->results = results;
  b->pending = m;
  b->driver_rc = DDLO_OK;
  b->driver_err.clear();
  if (b->mode == DDLO_BATCH_WAVES) {
    const int rc = waves_setup(b);
    if (rc != DDLO_OK) {
      b->pending = 0;
      return rc;
    }
    b->driver = std::thread(waves_run, b, jobs, m);  // `jobs` must stay valid until ddlo_batch_wait
    return DDLO_OK;
  }
  const int S = (int)b->lanes.size();
  const int T = std::min(b->host_threads, S);
  std::vector<std::string> errs(T);
  auto drive = [&](int t) {
    cudaSetDevice(b->device);
    // thread t owns the lanes t, t + T, ...; unit i runs on lane i % S, units of one lane stay in submission order
    for (int i = 0; i < m; ++i) {
      const int lane = i % S;
      if (lane % T != t) continue;
      const int rc = enqueue_unit(b, b->lanes[lane].rt, b->lanes[lane].eng, b->lanes[lane].has_shared_target, jobs[i], i, nullptr, nullptr);
      b->unit_rc[i] = rc;
      if (rc != DDLO_OK && errs[t].empty()) errs[t] = ddlo_last_error();
    }
  };
  if (T <= 1) {
    drive(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(drive, t);
    drive(0);
    for (auto& x : th) x.join();
  }
  for (int i = 0; i < m; ++i)
    if (b->unit_rc[i] != DDLO_OK) {
      for (int t = 0; t < T; ++t)
        if (!errs[t].empty()) return fail(b->unit_rc[i], errs[t]);  // ddlo_batch_wait still has to be called
      return fail(b->unit_rc[i], "batch: a unit could not be enqueued");
    }
  return DDLO_OK;
}

int ddlo_batch_wait(ddlo_batch* b) {
  if (!b) return fail(DDLO_E_INVALID, "batch is null");
  if (!b->pending) return DDLO_OK;
  DDLO_CUDA(cudaSetDevice(b->device));
  int rc = DDLO_OK;
  if (b->driver.joinable()) {
    b->driver.join();
    if (b->driver_rc != DDLO_OK) rc = fail(b->driver_rc, b->driver_err.empty() ? "batch: a unit failed" : b->driver_err);
  }
  for (auto& l : b->lanes)
    if (cudaStreamSynchronize(l.rt->stream) != cudaSuccess && rc == DDLO_OK) rc = fail(DDLO_E_CUDA, "batch: synchronisation failed");
  const int m = b->pending;
  b->pending = 0;
  for (int i = 0; i < m; ++i) {
    if (b->unit_rc[i] != DDLO_OK) {
      std::memset(b->results + i, 0, sizeof(ddlo_align_result));
      if (rc == DDLO_OK) rc = fail(b->unit_rc[i], "batch: a unit could not be enqueued (see the status of ddlo_batch_submit)");
      continue;
    }
    fill_result(b->h_out + i, b->covs_computed[i], b->results + i);
    if (rc == DDLO_OK && (b->results[i].flags & DDLO_FLAG_NONFINITE)) rc = fail(DDLO_E_NONFINITE, "batch: an input cloud holds NaN or Inf coordinates");
  }
  b->results = nullptr;
  return rc;
}

int ddlo_batch_run(ddlo_batch* b, const ddlo_batch_job* jobs, int m, ddlo_align_result* results) {
  const int rc = ddlo_batch_submit(b, jobs, m, results);
  const int rw = ddlo_batch_wait(b);
  return rc != DDLO_OK ? rc : rw;
}

int ddlo_batch_set_mode(ddlo_batch* b, int mode, int wave_units) {
  if (!b) return fail(DDLO_E_INVALID, "batch is null");
  if (b->pending) return fail(DDLO_E_NOT_READY, "batch: a submission is in flight");
  if (mode != DDLO_BATCH_WAVES && mode != DDLO_BATCH_LANES) return fail(DDLO_E_INVALID, "batch: unknown mode");
  if (mode == DDLO_BATCH_WAVES) {
    if (wave_units <= 0) wave_units = b->wave_units;
    if (wave_units > 4096) return fail(DDLO_E_INVALID, "batch: at most 4096 units per wave");
    if (!b->wb[0].slots.empty() && (int)b->wb[0].slots.size() != wave_units) {
      cudaSetDevice(b->device);
      for (auto& l : b->lanes) cudaStreamSynchronize(l.rt->stream);
      waves_free(b);
    }
    b->wave_units = wave_units;
  }
  b->mode = mode;
  for (auto& l : b->lanes) ddlo_runtime_set_align_blocks(l.rt, mode == DDLO_BATCH_LANES ? b->align_blocks : 0);
  return DDLO_OK;
}

int ddlo_batch_stats(ddlo_batch* b, long long* wave_rounds, long long* wave_polls) {
  if (!b) return fail(DDLO_E_INVALID, "batch is null");
  if (wave_rounds) *wave_rounds = b->wave_rounds;
  if (wave_polls) *wave_polls = b->wave_polls;
  return DDLO_OK;
}

int ddlo_batch_launch_count(ddlo_batch* b, long long* count) {
  if (!b || !count) return fail(DDLO_E_INVALID, "null argument");
  long long n = b->align_rt ? b->align_rt->launches : 0;
  for (auto& l : b->lanes) n += l.rt->launches;
  *count = n;
  return DDLO_OK;
}

}  // extern "C"
