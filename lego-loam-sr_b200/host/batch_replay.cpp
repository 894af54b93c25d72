// Batched replay of independent sequences (BASELINE.json configs[4]): one MapOptimization mirror — one ll_handle, one
// CUDA stream — per sequence, `n_threads` host threads, each replaying its share of the sequences frame by frame.
// There is no data-path exchange between sequences or GPUs (replicas only); the caller gathers the
// result records. Everything numerical still happens behind the C ABI; this file is the host-side driver the
// reference would run as N nodes in N processes.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <algorithm>
#include <thread>
#include <vector>

#if defined(__linux__)
#include <pthread.h>
#include <sched.h>
#endif

#include "map_optimization.hpp"

using lego_loam_b200::MapOptimization;

extern "C" {

// One frame of one sequence: transformSum and the five clouds of MapOptimization::FrameView (host memory,
// `stride` bytes per point; see map_optimization.hpp).
struct llh_frame {
  const float* odom6;
  const void* cloud[5];  // corner_scan, corner_scan_ds, surf_last_ds, outlier_last_ds, surf_total_ds
  size_t n[5];
};

// Replays frames [first, first + count) of every sequence through its node. records: [n_seq][count][16] floats =
// {pose[6], status, iter_num, sel_num, corner_map_ds, surf_map_ds, min_lambda, cf_mean, is_degenerate, frame ms, 0}.
// cpu_ids (optional): thread t is pinned to cpu_ids[t % n_cpu_ids]. *seconds = wall time from the common start of the
// threads to the last one finishing (every result is on the host by then). Returns 0, or -1 with err filled.
// group: sequences one thread keeps IN FLIGHT together (>= 1). A thread takes its sequences `group` at a time and, frame by
// frame, begins the frame of each (scan kernel and keyframe upload enqueued) before it ends any of them (wait, keyframe
// store): with fewer host cores than scans the GPU can usefully hold (an 8-GPU box with 4 cores per GPU), two or more scans
// per thread keep the SMs busy. The results do not depend on it (every sequence is its own node, handle and stream).
int llh_batch_replay2(void* const* nodes, int n_seq, const llh_frame* const* frames, int first, int count, size_t stride,
                      int n_threads, int group, const int* cpu_ids, int n_cpu_ids, float* records, double* seconds, char* err,
                      size_t err_cap) {
  if (!nodes || !frames || n_seq < 0 || count < 0 || n_threads < 1 || group < 1) return -1;
  n_threads = n_threads > n_seq ? (n_seq > 0 ? n_seq : 1) : n_threads;
  std::atomic<int> ready{0};
  std::atomic<bool> go{false}, failed{false};
  std::string first_error;
  std::vector<std::chrono::steady_clock::time_point> t_end((size_t)n_threads);
  std::chrono::steady_clock::time_point t_start;
  auto worker = [&](int t) {
#if defined(__linux__)
    if (cpu_ids && n_cpu_ids > 0) {
      cpu_set_t set;
      CPU_ZERO(&set);
      CPU_SET(cpu_ids[t % n_cpu_ids], &set);
      pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
#endif
    ready.fetch_add(1);
    while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
    try {
      // this thread's sequences t, t + n_threads, ... in groups of `group`; within a group frame-major, begin-all then end-all
      std::vector<int> mine;
      for (int q = t; q < n_seq; q += n_threads) mine.push_back(q);
      std::vector<MapOptimization::FrameView> views((size_t)group);
      for (size_t g0 = 0; g0 < mine.size() && !failed.load(); g0 += (size_t)group) {
        const size_t ng = std::min((size_t)group, mine.size() - g0);
        for (int k = 0; k < count && !failed.load(); ++k) {
          for (size_t u = 0; u < ng; ++u) {
            const int q = mine[g0 + u];
            const llh_frame& f = frames[q][first + k];
            MapOptimization::FrameView& v = views[u];
            v.corner_scan = f.cloud[0]; v.n_corner_scan = f.n[0];
            v.corner_scan_ds = f.cloud[1]; v.n_corner_scan_ds = f.n[1];
            v.surf_last_ds = f.cloud[2]; v.n_surf_last_ds = f.n[2];
            v.outlier_last_ds = f.cloud[3]; v.n_outlier_last_ds = f.n[3];
            v.surf_total_ds = f.cloud[4]; v.n_surf_total_ds = f.n[4];
            v.device = false;
            v.stride_bytes = stride;
            static_cast<MapOptimization*>(nodes[q])->processFrameViewBegin(f.odom6, v);
          }
          for (size_t u = 0; u < ng; ++u) {
            const int q = mine[g0 + u];
            MapOptimization* M = static_cast<MapOptimization*>(nodes[q]);
            M->processFrameViewEnd();
            if (records) {
              float* r = records + ((size_t)q * (size_t)count + (size_t)k) * 16;
              const ll_result& R = M->lastResult;
              std::memcpy(r, R.pose, 6 * sizeof(float));
              r[6] = (float)R.status; r[7] = (float)R.iter_num; r[8] = (float)R.sel_num;
              r[9] = (float)R.corner_map_ds; r[10] = (float)R.surf_map_ds; r[11] = R.min_lambda; r[12] = R.cf_mean;
              r[13] = (float)R.is_degenerate; r[14] = (float)M->lastFrameMs; r[15] = 0.f;
            }
          }
        }
      }
    } catch (const std::exception& e) {
      if (!failed.exchange(true)) first_error = e.what();
    }
    t_end[(size_t)t] = std::chrono::steady_clock::now();
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; ++t) pool.emplace_back(worker, t);
  while (ready.load() < n_threads) std::this_thread::yield();
  t_start = std::chrono::steady_clock::now();
  go.store(true, std::memory_order_release);
  for (auto& th : pool) th.join();
  double secs = 0.0;
  for (const auto& te : t_end) secs = std::max(secs, std::chrono::duration<double>(te - t_start).count());
  if (seconds) *seconds = secs;
  if (failed.load()) {
    if (err && err_cap) std::snprintf(err, err_cap, "%s", first_error.c_str());
    return -1;
  }
  return 0;
}

int llh_batch_replay(void* const* nodes, int n_seq, const llh_frame* const* frames, int first, int count, size_t stride,
                     int n_threads, const int* cpu_ids, int n_cpu_ids, float* records, double* seconds, char* err,
                     size_t err_cap) {
  return llh_batch_replay2(nodes, n_seq, frames, first, count, stride, n_threads, 1, cpu_ids, n_cpu_ids, records, seconds, err,
                           err_cap);
}

}  // extern "C"
