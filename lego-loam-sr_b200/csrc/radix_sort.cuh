// Stable LSD radix sort of (uint32 key, uint32 value) pairs, one kernel per 8-bit digit
// ("onesweep": per-tile digit counts are chained through a decoupled look-back, so each pass reads
// and writes every pair exactly once). Replaces the std::sort inside pcl::VoxelGrid (the dominant
// CPU cost of a3) and builds the cell order of the search grid that replaces the
// KD-tree (a4). Stability is what makes the within-voxel accumulation order = ascending original
// index, i.e. the canonical order the oracle uses.
//
// Everything data-dependent stays on the device: the element count and the number of significant
// key bits are read from device memory, so a whole map build is enqueued without a host round trip.
// Passes above the significant bits return immediately; because executed passes are always
// 0..P-1, pass p reads buffer (p & 1) and writes buffer ((p + 1) & 1), and the result is in
// buffer (P & 1) (written to *final_buf by the key-generating kernel).
//
// Up to two independent problems (corner map, surf map) are sorted by the same launches
// (blockIdx.y selects the problem).
#pragma once
#include "common.cuh"

namespace ll {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 2048 pairs per tile
constexpr int RS_RADIX = 256;
constexpr int RS_MAX_PASSES = 4;
constexpr int RS_LOOKBACK = 8;  // predecessor tiles inspected per look-back round trip
constexpr uint32_t RS_FLAG_AGG = 1u << 30;     // tile aggregate published
constexpr uint32_t RS_FLAG_PREFIX = 2u << 30;  // inclusive prefix published
constexpr uint32_t RS_VALUE_MASK = (1u << 30) - 1;

struct SortProblem {
  uint32_t* keys[2];
  uint32_t* vals[2];
  const int* n;            // element count (device)
  const int* bits;         // significant key bits (device), 0..32
  uint32_t* ghist;         // [RS_MAX_PASSES][256] digit histograms of the whole input (pre-zeroed, filled by keygen)
  uint32_t* status;        // [max_tiles][RS_MAX_PASSES][256] look-back words (pre-zeroed)
  uint32_t* tile_counter;  // [RS_MAX_PASSES] dynamic tile tickets (pre-zeroed)
  int max_tiles;
};

// Block-level digit histogram helper for key-generating kernels: accumulate in shared memory,
// flush once per block. smem must hold RS_MAX_PASSES * 256 uint32 and be zeroed by the caller.
// Only the passes that will run (8 * p < bits) are counted: the upper digits are zero for every key and
// would serialise all 32 lanes on one shared-memory counter.
__device__ __forceinline__ void rs_hist_add(uint32_t* smem_hist, uint32_t key, int bits) {
#pragma unroll
  for (int p = 0; p < RS_MAX_PASSES; ++p)
    if (8 * p < bits) atomicAdd(&smem_hist[p * RS_RADIX + ((key >> (8 * p)) & 255u)], 1u);
}
__device__ __forceinline__ void rs_hist_flush(const uint32_t* smem_hist, uint32_t* ghist) {
  for (int i = threadIdx.x; i < RS_MAX_PASSES * RS_RADIX; i += blockDim.x) {
    uint32_t c = smem_hist[i];
    if (c) atomicAdd(&ghist[i], c);
  }
}

struct SortBatch {
  SortProblem p[2];
};

__global__ void __launch_bounds__(RS_THREADS) rs_onesweep_kernel(const __grid_constant__ SortBatch batch, int pass) {
  pdl_entry();
  const SortProblem& P = batch.p[blockIdx.y];
  const int bits = *P.bits;
  if (pass * 8 >= bits) return;  // digit is zero for every key: the pass is the identity
  const int n = *P.n;

  __shared__ uint32_t warp_hist[RS_WARPS][RS_RADIX];
  __shared__ uint32_t digit_base[RS_RADIX];
  __shared__ uint32_t scan_buf[2 * RS_WARPS];
  __shared__ uint32_t tile_first[RS_RADIX];
  __shared__ uint32_t stage_keys[RS_TILE], stage_vals[RS_TILE];
  __shared__ int s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = (int)atomicAdd(&P.tile_counter[pass], 1u);
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) warp_hist[w][tid] = 0;
  __syncthreads();
  const int tile = s_tile;
  const long long tile_base = (long long)tile * RS_TILE;
  if (tile_base >= n) return;  // tiles are handed out in order, so nobody waits on an empty tile

  const uint32_t* __restrict__ keys_in = P.keys[pass & 1];
  const uint32_t* __restrict__ vals_in = P.vals[pass & 1];
  uint32_t* __restrict__ keys_out = P.keys[(pass + 1) & 1];
  uint32_t* __restrict__ vals_out = P.vals[(pass + 1) & 1];
  const int shift = pass * 8;

  // Warp-striped load: warp w owns pairs [w*256, (w+1)*256) of the tile; item k of lane l is pair
  // k*32 + l of that range, so (k, lane) order == memory order and the loads coalesce.
  uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
  const long long warp_base = tile_base + (long long)warp * (32 * RS_ITEMS);
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    long long i = warp_base + k * 32 + lane;
    bool valid = i < n;
    key[k] = valid ? keys_in[i] : 0xffffffffu;
    val[k] = valid ? vals_in[i] : 0u;
  }
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const bool valid = (warp_base + k * 32 + lane) < n;
    const uint32_t d = (key[k] >> shift) & 255u;
    const uint32_t mkey = valid ? d : 256u;  // out-of-range items form their own group and are not counted
    const unsigned peers = __match_any_sync(0xffffffffu, mkey);
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader && valid) {
      old = warp_hist[warp][d];
      warp_hist[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[k] = old + __popc(peers & lt_mask);
    __syncwarp();
  }
  __syncthreads();

  // Thread t owns digit t: exclusive scan over the warps, then the look-back over earlier tiles.
  uint32_t tile_count = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) {
    uint32_t c = warp_hist[w][tid];
    warp_hist[w][tid] = tile_count;
    tile_count += c;
  }
  // tile-major layout [tile][pass][digit]: the words a build can touch are a prefix of the array, so the host
  // only has to zero ceil(n / RS_TILE) tiles' worth, not the whole capacity
  constexpr size_t kTileStride = (size_t)RS_MAX_PASSES * RS_RADIX;
  volatile uint32_t* status = P.status + (size_t)pass * RS_RADIX;
  uint32_t excl = 0;
  if (tile == 0) {
    status[(size_t)tile * kTileStride + tid] = tile_count | RS_FLAG_PREFIX;
  } else {
    status[(size_t)tile * kTileStride + tid] = tile_count | RS_FLAG_AGG;
    // Windowed look-back: RS_LOOKBACK predecessor words are fetched together (independent loads, one round
    // trip) and consumed in order. When every tile of a pass is resident at once they all finish ranking at
    // the same moment, and a one-word-at-a-time walk would serialise ~tiles/2 L2 round trips.
    int j = tile - 1;
    bool done = false;
    while (!done) {
      uint32_t w[RS_LOOKBACK];
#pragma unroll
      for (int u = 0; u < RS_LOOKBACK; ++u) {
        w[u] = 2u << 30;  // RS_FLAG_PREFIX: virtual empty prefix before tile 0
        if (j - u >= 0) w[u] = status[(size_t)(j - u) * kTileStride + tid];
      }
      int consumed = 0;
#pragma unroll
      for (int u = 0; u < RS_LOOKBACK; ++u) {
        if (done || consumed != u) continue;  // stopped at an unpublished tile: re-read from there
        if (w[u] & RS_FLAG_PREFIX) {
          excl += w[u] & RS_VALUE_MASK;
          done = true;
        } else if (w[u] & RS_FLAG_AGG) {
          excl += w[u] & RS_VALUE_MASK;
          ++consumed;
        }
        // else: that tile has its ticket but has not published yet; it is running, so retry
      }
      j -= consumed;
    }
    status[(size_t)tile * kTileStride + tid] = (excl + tile_count) | RS_FLAG_PREFIX;
  }
  // Exclusive prefixes over the 256 digits of (a) the global digit histogram and (b) this tile's digit counts:
  // warp-shuffle scans of the eight 32-digit groups + one pass over the eight group totals.
  const uint32_t g = P.ghist[pass * RS_RADIX + tid];
  uint32_t gi = g, ti = tile_count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t ug = __shfl_up_sync(0xffffffffu, gi, o), ut = __shfl_up_sync(0xffffffffu, ti, o);
    if (lane >= o) { gi += ug; ti += ut; }
  }
  if (lane == 31) { scan_buf[warp] = gi; scan_buf[RS_WARPS + warp] = ti; }
  __syncthreads();
  uint32_t goff = 0, toff = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) {
    goff += w < warp ? scan_buf[w] : 0u;
    toff += w < warp ? scan_buf[RS_WARPS + w] : 0u;
  }
  const uint32_t tile_excl = toff + ti - tile_count;   // items of this tile with a smaller digit
  digit_base[tid] = goff + gi - g + excl - tile_excl;   // global position of tile-local position 0 ... shifted per digit:
  tile_first[tid] = tile_excl;                          // item at tile-local position j with digit d goes to digit_base[d] + j
  __syncthreads();

  // Stage the tile in digit order in shared memory, then write it out with consecutive threads on consecutive tile-local
  // positions: runs of equal digits become contiguous global stores (full sectors instead of 4-byte scatters).
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const bool valid = (warp_base + k * 32 + lane) < n;
    if (valid) {
      const uint32_t d = (key[k] >> shift) & 255u;
      const uint32_t lpos = tile_first[d] + warp_hist[warp][d] + rank[k];
      stage_keys[lpos] = key[k];
      stage_vals[lpos] = val[k];
    }
  }
  __syncthreads();
  const int tile_n = (int)min((long long)RS_TILE, (long long)n - tile_base);
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const int j = k * RS_THREADS + tid;
    if (j < tile_n) {
      const uint32_t kk = stage_keys[j];
      const uint32_t pos = digit_base[(kk >> shift) & 255u] + (uint32_t)j;
      keys_out[pos] = kk;
      vals_out[pos] = stage_vals[j];
    }
  }
}

}  // namespace ll
