// Local-map assembly on the device:
//   transform_minmax_kernel  a2  transformPointCloud of every selected keyframe cloud, concatenated
//                                in cache order, fused with the bounding-box reduction VoxelGrid needs
//   voxel_keys_kernel        a3  voxel index per point (PCL arithmetic, bit-exact) + digit histograms
//   rs_onesweep_kernel       a3  stable radix sort by voxel index (radix_sort.cuh)
//   voxel_reduce_kernel      a3  one output point per occupied voxel, in ascending voxel index,
//                                = mean of x,y,z,intensity accumulated in ascending original index
//   grid_keys_kernel         a4  1 m cell key per downsampled point + table clear + histograms
//   rs_onesweep_kernel       a4  sort by cell key
//   grid_finalize_kernel     a4  cell-ordered point array (w = map index) + per-cell [start,end)
// The corner map and the surf map go through the same launches (blockIdx.y = map).
// These stand-alone kernels are the general path (any size, stage-level ABI, persistent_lm = 0); a VLP-16-sized local
// map is normally assembled inside the scan kernel instead (map_inline.cuh), with bit-identical results.
//
// Why a 1 m uniform grid is EXACT for this path: the reference only uses a neighbour set when
// pointSearchSqDis[4] < 1.0 (mapOptmization.cpp:1279, :1389). cell(p) = floorf(p) is exact in float,
// and a float-evaluated squared distance below 1 implies |dx|,|dy|,|dz| < 1 in exact arithmetic, so
// every such neighbour lies in the 3x3x3 cells around the query's cell.
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"

namespace ll {

constexpr int MB_THREADS = 256;
constexpr int MB_ITEMS = 4;
constexpr int MB_TILE = MB_THREADS * MB_ITEMS;  // 1024 points per block

// Device-resident, data-dependent state of one map (zeroed at the start of every build).
struct MapState {
  uint32_t mm[6];        // bounding box of the raw cloud, ordered-uint encoded:
                         // mm[0..2] = ~ordered(min xyz) (so atomicMax tracks the minimum), mm[3..5] = ordered(max xyz)
  int n_raw;             // points in the concatenated raw cloud
  int n_ds;              // points after VoxelGrid
  int voxel_bits;        // significant bits of the voxel keys
  int voxel_overflow;    // PCL's "leaf size too small" pass-through was taken
  int grid_bits;
  int grid_too_large;    // search grid exceeds the cell budget
  int gox, goy, goz;     // grid origin (integer metres) and dimensions in cells
  int gdx, gdy, gdz;
  int ncells;
  uint32_t head_ticket;  // dynamic tile ticket of voxel_reduce_kernel
  int inc_fail;          // the incremental assembly (map_incremental.cuh) met a case it cannot represent: rebuild in full
  int cell_tag;          // stamp of the cell-table entries this build wrote (cell_entry_* below)
};

// ---- the cell table is never cleared between builds --------------------------------------------------------------------
// A dense table over the bounding box of a 200 x 300 m map has millions of cells of which a few ten thousand are occupied;
// clearing it cost more than everything else in the map assembly (89 us of 240 per scan at the benchmark's sizes,
// profiles/r02_*). Every entry therefore carries the stamp of the build that wrote it in its top byte:
//   x = ((255 - tag) << 24) | start      y = (tag << 24) | end          (positions < 2^24)
// An entry is valid when y's stamp is the current build's. Stamps grow from build to build, so with UNSIGNED atomics a
// stale y is always below and a stale x always above anything the current build writes: atomicMax / atomicMin (inline
// build) and plain stores (stand-alone build) both work on uncleared memory. The host resets the table to (~0, 0) when it
// is allocated and every 255 builds, when the stamp wraps (abi.cu: next_cell_tag).
constexpr uint32_t CELL_POS_MASK = 0x00ffffffu;
constexpr int CELL_MAX_POINTS = 1 << 24;
__device__ __forceinline__ uint32_t cell_entry_x(int tag, uint32_t start) { return ((uint32_t)(255 - tag) << 24) | start; }
__device__ __forceinline__ uint32_t cell_entry_y(int tag, uint32_t end) { return ((uint32_t)tag << 24) | end; }
// [start, end) of a loaded entry; (0, 0) when the entry is not of this build
__device__ __forceinline__ int2 cell_decode(int2 e, int tag) {
  const bool ok = ((uint32_t)e.y >> 24) == (uint32_t)tag;
  return ok ? make_int2((int)((uint32_t)e.x & CELL_POS_MASK), (int)((uint32_t)e.y & CELL_POS_MASK)) : make_int2(0, 0);
}

struct MapBuffers {
  float4* raw;            // concatenated surrounding cloud in the map frame (laserCloud*FromMap)
  float4* ds;             // laserCloud*FromMapDS (== raw when map_downsample = 0)
  float4* sorted;         // ds reordered by cell key; w = index into ds (bit pattern)
  int2* cells;            // per cell [start, end) into sorted, stamped (cell_entry_x / _y)
  MapState* st;
  uint32_t* head_status;  // look-back words of voxel_reduce_kernel
  int* member;            // optional: output slot per raw point (stage-level parity tests)
  int n_raw;              // host-known size of raw
  int cell_cap;           // capacity of cells
  int cell_tag;           // stamp of the next grid build over this map
  float inv_leaf;         // 1 / leaf (computed on the host in float, like PCL)
  int downsample;
};

struct MapBatch {
  MapBuffers m[2];
  SortProblem sort[2];
  int count;  // 1 or 2 maps in this batch (= gridDim.y of the batched kernels)
};

// One contiguous piece of one keyframe cloud.
struct Segment {
  const float4* src;
  int count;
  int dst_off;  // offset in the destination raw cloud
  int map;      // 0 corner, 1 surf
  int identity; // 1 = already in the map frame
  Trig t;
  int pad[3];
};
struct BlockJob {
  int seg;
  int start;
};

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(MB_THREADS) transform_minmax_kernel(const __grid_constant__ MapBatch B,
                                                                      const Segment* __restrict__ segs,
                                                                      const BlockJob* __restrict__ jobs) {
  pdl_entry();
  const BlockJob job = jobs[blockIdx.x];
  const Segment s = segs[job.seg];
  const MapBuffers& M = B.m[s.map];
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  bool any = false;
#pragma unroll
  for (int k = 0; k < MB_ITEMS; ++k) {
    const int i = job.start + k * MB_THREADS + threadIdx.x;
    if (i < s.count) {
      const float4 p = ld_stream(s.src + i);
      const float4 q = s.identity ? p : rigid_transform(s.t, p);
      M.raw[s.dst_off + i] = q;
      if (finite3(q)) {
        any = true;
        mn[0] = fminf(mn[0], q.x); mx[0] = fmaxf(mx[0], q.x);
        mn[1] = fminf(mn[1], q.y); mx[1] = fmaxf(mx[1], q.y);
        mn[2] = fminf(mn[2], q.z); mx[2] = fmaxf(mx[2], q.z);
      }
    }
  }
  __shared__ float s_mn[3][MB_THREADS / 32], s_mx[3][MB_THREADS / 32];
  __shared__ int s_any;
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  if (any) s_any = 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float v0 = warp_min(mn[a]), v1 = warp_max(mx[a]);
    if (lane == 0) { s_mn[a][warp] = v0; s_mx[a][warp] = v1; }
  }
  __syncthreads();
  if (threadIdx.x < 3 && s_any) {
    const int a = threadIdx.x;
    float v0 = s_mn[a][0], v1 = s_mx[a][0];
#pragma unroll
    for (int w = 1; w < MB_THREADS / 32; ++w) { v0 = fminf(v0, s_mn[a][w]); v1 = fmaxf(v1, s_mx[a][w]); }
    atomicMax(&M.st->mm[a], ~float_to_ordered(v0));
    atomicMax(&M.st->mm[3 + a], float_to_ordered(v1));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int m = 0; m < B.count; ++m) {
      B.m[m].st->n_raw = B.m[m].n_raw;
      if (!B.m[m].downsample) B.m[m].st->n_ds = B.m[m].n_raw;
    }
  }
}

// pcl::VoxelGrid<PointXYZI>::applyFilter index arithmetic (PCL 1.8-1.12 voxel_grid.hpp; not vendored in
// the reference, restated from the published algorithm — SURVEY.md §8c). Everything a thread needs is
// recomputed from the bounding box; no setup kernel.
struct VoxelMeta {
  int min_b[3];
  int mul[3];
  long long total;  // number of voxels in the box; the key of non-finite points
  bool overflow;
  bool empty;
};

// kCoherent: read the box through L2 (__ldcg) — needed in the single-kernel scan (map_inline.cuh), where it is produced
// by atomics of the same launch; the stand-alone kernels read it through L1 like any other input.
template <bool kCoherent = false>
__device__ __forceinline__ VoxelMeta voxel_meta(const MapState* st, float inv) {
  VoxelMeta v;
  auto ld = [](const uint32_t* p) { return kCoherent ? __ldcg(p) : *p; };
  const uint32_t e0 = ld(&st->mm[0]);
  v.empty = (e0 == 0u);  // no finite point seen
  float mn[3], mx[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mn[a] = ordered_to_float(~ld(&st->mm[a]));
    mx[a] = ordered_to_float(ld(&st->mm[3 + a]));
  }
  if (v.empty) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { mn[a] = 0.f; mx[a] = 0.f; }
  }
  const long long dx = (long long)((mx[0] - mn[0]) * inv) + 1;
  const long long dy = (long long)((mx[1] - mn[1]) * inv) + 1;
  const long long dz = (long long)((mx[2] - mn[2]) * inv) + 1;
  v.overflow = (dx * dy * dz) > 2147483647LL;
  int div_b[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    v.min_b[a] = (int)floorf(mn[a] * inv);
    const int max_b = (int)floorf(mx[a] * inv);
    div_b[a] = max_b - v.min_b[a] + 1;
  }
  v.mul[0] = 1;
  v.mul[1] = div_b[0];
  v.mul[2] = div_b[0] * div_b[1];
  v.total = (long long)div_b[0] * div_b[1] * div_b[2];
  return v;
}

__global__ void __launch_bounds__(MB_THREADS) voxel_keys_kernel(const __grid_constant__ MapBatch B) {
  pdl_entry();
  const MapBuffers& M = B.m[blockIdx.y];
  if (!M.downsample) return;
  const int n = M.n_raw;
  const int base = blockIdx.x * MB_TILE;
  if (base >= n && !(blockIdx.x == 0)) return;
  const SortProblem& S = B.sort[blockIdx.y];
  __shared__ uint32_t s_hist[RS_MAX_PASSES * RS_RADIX];
  for (int i = threadIdx.x; i < RS_MAX_PASSES * RS_RADIX; i += MB_THREADS) s_hist[i] = 0;
  __syncthreads();
  const VoxelMeta v = voxel_meta(M.st, M.inv_leaf);
  const uint32_t sentinel = v.overflow ? (uint32_t)n : (uint32_t)v.total;
  const int bits = sentinel == 0 ? 0 : 32 - __clz(sentinel);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    M.st->voxel_overflow = v.overflow ? 1 : 0;
    M.st->voxel_bits = bits;
  }
  const float inv = M.inv_leaf;
#pragma unroll
  for (int k = 0; k < MB_ITEMS; ++k) {
    const int i = base + k * MB_THREADS + threadIdx.x;
    if (i < n) {
      const float4 p = M.raw[i];
      uint32_t key;
      if (v.overflow) {
        key = (uint32_t)i;  // PCL returns the input unchanged (non-finite points included): every point its own voxel
      } else if (!finite3(p)) {
        key = sentinel;  // sorted behind every voxel and dropped, like PCL skips non-finite points
      } else {
        const int i0 = (int)(floorf(p.x * inv) - (float)v.min_b[0]);
        const int i1 = (int)(floorf(p.y * inv) - (float)v.min_b[1]);
        const int i2 = (int)(floorf(p.z * inv) - (float)v.min_b[2]);
        key = (uint32_t)(i0 * v.mul[0] + i1 * v.mul[1] + i2 * v.mul[2]);
      }
      S.keys[0][i] = key;
      S.vals[0][i] = (uint32_t)i;
      rs_hist_add(s_hist, key, bits);
    }
  }
  __syncthreads();
  rs_hist_flush(s_hist, S.ghist);
}

// Sorted (voxel, original index) pairs -> one centroid per voxel, written in ascending voxel index.
// Per tile of 1024 sorted pairs:
//   1. every thread gathers the points of its four pairs into shared memory (independent loads: one
//      round trip) and flags segment heads (first pair of a voxel);
//   2. block scan + single-word decoupled look-back give every head its output slot;
//   3. the head's thread adds its segment up IN SORTED ORDER (= ascending original index, the canonical
//      order: the float sum is the same sequence of additions the CPU performs) from shared memory —
//      the segment length is known from the next head, so the loop is unrolled with the loads ahead of
//      the adds; a segment that runs past the tile continues from global memory.
__global__ void __launch_bounds__(MB_THREADS) voxel_reduce_kernel(const __grid_constant__ MapBatch B) {
  pdl_entry();
  const MapBuffers& M = B.m[blockIdx.y];
  if (!M.downsample) return;
  const SortProblem& S = B.sort[blockIdx.y];
  MapState* st = M.st;
  const int n = M.n_raw;
  __shared__ float4 s_pts[MB_TILE];
  __shared__ uint32_t s_val[MB_TILE];
  __shared__ int s_head_pos[MB_TILE + 1];
  __shared__ int s_tile;
  __shared__ uint32_t s_warp_sum[MB_THREADS / 32];
  __shared__ uint32_t s_tile_excl;
  __shared__ int s_sent_pos;  // local position of the first sentinel-keyed (non-finite) pair, if any
  __shared__ float4 s_cross_sum;
  __shared__ int s_cross, s_cross_cnt, s_first_miss;
  __shared__ uint32_t s_cross_slot, s_cross_key;
  if (threadIdx.x == 0) {
    s_tile = (int)atomicAdd(&st->head_ticket, 1u);
    s_sent_pos = MB_TILE;
    s_cross = 0;
  }
  __syncthreads();
  const int tile = s_tile;
  const int base = tile * MB_TILE;
  if (base >= n) return;
  const int fin = ((st->voxel_bits + 7) >> 3) & 1;
  const uint32_t* __restrict__ keys = S.keys[fin];
  const uint32_t* __restrict__ vals = S.vals[fin];
  const VoxelMeta v = voxel_meta(st, M.inv_leaf);
  const uint32_t sentinel = v.overflow ? (uint32_t)n : (uint32_t)v.total;

  const int l0 = threadIdx.x * MB_ITEMS;  // local index of this thread's first pair
  const int i0 = base + l0;
  uint32_t key[MB_ITEMS], val[MB_ITEMS];
  bool head[MB_ITEMS];
  uint32_t prev = (i0 > 0 && i0 < n) ? keys[i0 - 1] : 0xffffffffu;
#pragma unroll
  for (int k = 0; k < MB_ITEMS; ++k) {
    const int i = i0 + k;
    key[k] = (i < n) ? keys[i] : 0xffffffffu;
    val[k] = (i < n) ? vals[i] : 0u;
  }
#pragma unroll
  for (int k = 0; k < MB_ITEMS; ++k) {  // gather: four independent loads in flight per thread
    const int i = i0 + k;
    s_pts[l0 + k] = (i < n) ? M.raw[val[k]] : make_float4(0.f, 0.f, 0.f, 0.f);
    s_val[l0 + k] = val[k];
  }
  uint32_t cnt = 0;
#pragma unroll
  for (int k = 0; k < MB_ITEMS; ++k) {
    const int i = i0 + k;
    const bool boundary = (i < n) && (i == 0 || key[k] != prev);
    head[k] = boundary && key[k] != sentinel;
    if (boundary && key[k] == sentinel) s_sent_pos = l0 + k;  // at most one such pair per tile (sorted)
    prev = key[k];
    cnt += head[k] ? 1u : 0u;
  }
  // block exclusive scan of cnt
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp_sum[warp] = incl;
  __syncthreads();
  uint32_t warp_off = 0, tile_total = 0;
#pragma unroll
  for (int w = 0; w < MB_THREADS / 32; ++w) {
    const uint32_t c = s_warp_sum[w];
    if (w < warp) warp_off += c;
    tile_total += c;
  }
  if (warp == 0) {  // warp-wide decoupled look-back: 32 predecessor tiles per round trip
    volatile uint32_t* status = M.head_status;
    uint32_t excl = 0;
    if (tile == 0) {
      if (lane == 0) status[0] = tile_total | RS_FLAG_PREFIX;
    } else {
      if (lane == 0) status[tile] = tile_total | RS_FLAG_AGG;
      int j = tile - 1;
      bool done = false;
      while (!done) {
        const int idx = j - lane;
        uint32_t w = 2u << 30;  // RS_FLAG_PREFIX: virtual empty prefix before tile 0
        if (idx >= 0) w = status[idx];
        const unsigned pre = __ballot_sync(0xffffffffu, (w & RS_FLAG_PREFIX) != 0);
        const unsigned nrd = __ballot_sync(0xffffffffu, (w & (RS_FLAG_PREFIX | RS_FLAG_AGG)) == 0);
        const int first_pre = pre ? __ffs(pre) - 1 : 32;
        const int first_nrd = nrd ? __ffs(nrd) - 1 : 32;
        done = first_pre < first_nrd;
        const int stop = done ? first_pre + 1 : first_nrd;  // lanes [0, stop) contribute
        uint32_t v = lane < stop ? (w & RS_VALUE_MASK) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        j -= stop;
      }
      if (lane == 0) status[tile] = (excl + tile_total) | RS_FLAG_PREFIX;
    }
    if (lane == 0) {
      s_tile_excl = excl;
      if (base + MB_TILE >= n) st->n_ds = (int)(excl + tile_total);  // last tile: laserCloud*FromMapDSNum
    }
  }
  // local head positions, so that a head knows where its segment ends (the next head, or the tile end)
  const uint32_t local0 = warp_off + (incl - cnt);
  {
    uint32_t ls = local0;
#pragma unroll
    for (int k = 0; k < MB_ITEMS; ++k)
      if (head[k]) s_head_pos[ls++] = l0 + k;
  }
  const int tile_n = min(MB_TILE, n - base);
  if (threadIdx.x == 0) s_head_pos[tile_total] = tile_n;
  __syncthreads();
  uint32_t ls = local0;
  const uint32_t slot0 = s_tile_excl;
#pragma unroll
  for (int k = 0; k < MB_ITEMS; ++k) {
    if (!head[k]) continue;
    const int beg = l0 + k;
    const int end = min(s_head_pos[ls + 1], s_sent_pos);  // sentinel-keyed pairs only ever trail the last head
    const uint32_t slot = slot0 + ls;
    ++ls;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    int c = 0;
    int j = beg;
    const uint32_t kk = key[k];
    for (; j + 4 <= end; j += 4) {  // loads first, then the four dependent adds per component
      const float4 p0 = s_pts[j], p1 = s_pts[j + 1], p2 = s_pts[j + 2], p3 = s_pts[j + 3];
      sx += p0.x; sy += p0.y; sz += p0.z; si += p0.w;
      sx += p1.x; sy += p1.y; sz += p1.z; si += p1.w;
      sx += p2.x; sy += p2.y; sz += p2.z; si += p2.w;
      sx += p3.x; sy += p3.y; sz += p3.z; si += p3.w;
    }
    for (; j < end; ++j) {
      const float4 p = s_pts[j];
      sx += p.x; sy += p.y; sz += p.z; si += p.w;
    }
    c = end - beg;
    if (M.member)
      for (int t = beg; t < end; ++t) M.member[s_val[t]] = (int)slot;
    if (end == tile_n && s_sent_pos >= tile_n && base + tile_n < n) {
      // The segment may continue in the following tiles: park the running sums; the whole block finishes it below.
      s_cross_sum = make_float4(sx, sy, sz, si);
      s_cross_cnt = c;
      s_cross_slot = slot;
      s_cross_key = kk;
      s_cross = 1;
      continue;
    }
    const float fc = (float)c;
    M.ds[slot] = make_float4(sx / fc, sy / fc, sz / fc, si / fc);
  }
  __syncthreads();
  if (!s_cross) return;
  // Continuation of the one segment that crosses the tile end. The block finds its end and gathers its points
  // MB_THREADS pairs at a time (coalesced key reads, independent gathers); thread 0 keeps adding in sorted order.
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  int c = 0;
  if (threadIdx.x == 0) {
    sx = s_cross_sum.x; sy = s_cross_sum.y; sz = s_cross_sum.z; si = s_cross_sum.w;
    c = s_cross_cnt;
  }
  const uint32_t ck = s_cross_key;
  const uint32_t cslot = s_cross_slot;
  for (int g0 = base + tile_n; g0 < n; g0 += MB_THREADS) {
    const int g = g0 + threadIdx.x;
    const bool match = g < n && keys[g] == ck;
    if (threadIdx.x == 0) s_first_miss = MB_THREADS;
    __syncthreads();
    if (!match) atomicMin(&s_first_miss, (int)threadIdx.x);
    __syncthreads();
    const int m = s_first_miss;  // pairs [g0, g0 + m) still belong to the segment
    if ((int)threadIdx.x < m) {
      const uint32_t src = vals[g];
      s_pts[threadIdx.x] = M.raw[src];
      if (M.member) M.member[src] = (int)cslot;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int t = 0; t < m; ++t) {
        const float4 p = s_pts[t];
        sx += p.x; sy += p.y; sz += p.z; si += p.w;
      }
      c += m;
    }
    __syncthreads();
    if (m < MB_THREADS) break;
  }
  if (threadIdx.x == 0) {
    const float fc = (float)c;
    M.ds[cslot] = make_float4(sx / fc, sy / fc, sz / fc, si / fc);
  }
}

__global__ void __launch_bounds__(MB_THREADS) grid_keys_kernel(const __grid_constant__ MapBatch B) {
  pdl_entry();
  const MapBuffers& M = B.m[blockIdx.y];
  const SortProblem& S = B.sort[blockIdx.y];
  MapState* st = M.st;
  const int n = st->n_ds;
  __shared__ uint32_t s_hist[RS_MAX_PASSES * RS_RADIX];
  for (int i = threadIdx.x; i < RS_MAX_PASSES * RS_RADIX; i += MB_THREADS) s_hist[i] = 0;
  __syncthreads();
  // grid geometry from the raw bounding box (a superset of the downsampled one)
  int o[3], d[3];
  const bool empty = (st->mm[0] == 0u);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float mn = empty ? 0.f : ordered_to_float(~st->mm[a]);
    const float mx = empty ? 0.f : ordered_to_float(st->mm[3 + a]);
    o[a] = (int)floorf(mn);
    d[a] = (int)floorf(mx) - o[a] + 1;
  }
  const long long ncells64 = (long long)d[0] * d[1] * d[2];
  const bool too_large = ncells64 > (long long)M.cell_cap || n >= CELL_MAX_POINTS;
  const int ncells = too_large ? 0 : (int)ncells64;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->cell_tag = M.cell_tag;
    st->gox = o[0]; st->goy = o[1]; st->goz = o[2];
    st->gdx = d[0]; st->gdy = d[1]; st->gdz = d[2];
    st->ncells = ncells;
    st->grid_too_large = too_large ? 1 : 0;
    st->grid_bits = ncells <= 1 ? 0 : 32 - __clz((uint32_t)(ncells - 1));
  }
  const int bits = ncells <= 1 ? 0 : 32 - __clz((uint32_t)(ncells - 1));
  // (no cell-table clear: entries are stamped, see cell_entry_x)
  const int base = blockIdx.x * MB_TILE;
#pragma unroll
  for (int k = 0; k < MB_ITEMS; ++k) {
    const int i = base + k * MB_THREADS + threadIdx.x;
    if (i < n) {
      const float4 p = M.ds[i];
      uint32_t key = 0;
      if (!too_large) {
        const int cx = (int)floorf(p.x) - o[0], cy = (int)floorf(p.y) - o[1], cz = (int)floorf(p.z) - o[2];
        key = (uint32_t)((cz * d[1] + cy) * d[0] + cx);
      }
      S.keys[0][i] = key;
      S.vals[0][i] = (uint32_t)i;
      rs_hist_add(s_hist, key, bits);
    }
  }
  __syncthreads();
  rs_hist_flush(s_hist, S.ghist);
}

__global__ void __launch_bounds__(MB_THREADS) grid_finalize_kernel(const __grid_constant__ MapBatch B) {
  pdl_entry();
  const MapBuffers& M = B.m[blockIdx.y];
  const SortProblem& S = B.sort[blockIdx.y];
  const MapState* st = M.st;
  const int n = st->n_ds;
  if (st->grid_too_large) return;
  const int fin = ((st->grid_bits + 7) >> 3) & 1;
  const uint32_t* __restrict__ keys = S.keys[fin];
  const uint32_t* __restrict__ vals = S.vals[fin];
  const int base = blockIdx.x * MB_TILE;
#pragma unroll
  for (int k = 0; k < MB_ITEMS; ++k) {
    const int i = base + k * MB_THREADS + threadIdx.x;
    if (i < n) {
      const uint32_t key = keys[i];
      const uint32_t v = vals[i];
      const float4 p = M.ds[v];
      M.sorted[i] = make_float4(p.x, p.y, p.z, __uint_as_float(v));
      if (i == 0 || keys[i - 1] != key) M.cells[key].x = (int)cell_entry_x(M.cell_tag, (uint32_t)i);
      if (i == n - 1 || keys[i + 1] != key) M.cells[key].y = (int)cell_entry_y(M.cell_tag, (uint32_t)(i + 1));
    }
  }
}

}  // namespace ll
