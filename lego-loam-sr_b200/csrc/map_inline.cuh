// Local-map assembly INSIDE the persistent scan kernel, by its co-resident grid of
// 512-thread blocks and grid barriers instead of twelve separate launches. The stand-alone kernels of map_build.cuh /
// radix_sort.cuh remain the general path (maps above the capacity below, stage-level entry points) and produce
// bit-identical arrays.
//
//   a2  transform + bounding box        blocks stride over the 1024-point jobs of the segment table   | barrier
//   a3  VoxelGrid                       every block owns ONE tile of 4/8/16 x 512 raw points (both maps share the
//                                       grid) and keeps its (voxel key, index) pairs in registers:
//         per 8-bit pass                rank the digits (warp match_any in memory order => stable), publish the
//                                       tile's 512-bin histogram                                      | barrier
//                                       sum the histograms of the tiles before this one / of all tiles, exclusive
//                                       digit prefix, scatter                                          | barrier
//         reduce                        segment heads per tile                                         | barrier
//                                       output slot = heads in earlier tiles + rank; the head's thread adds its
//                                       segment in sorted order (= ascending original index)           | barrier
//   a4  search grid                     the same sort on the 1 m cell key of the downsampled points; the LAST pass
//                                       writes the cell-ordered point itself and extends its cell's [start, end)
//                                       with atomicMin / atomicMax, so there is no finalize stage      | barrier
// A grid barrier costs ~1.3 us and a pass ~1.5 us of work, against a kernel boundary plus a decoupled look-back
// chain (~9-15 us) per pass of the stand-alone sort. `cells` marks an empty cell as (INT_MAX, 0) instead of (0, 0) —
// both read as "end <= start" by the search.
// Memory-ordering notes for the in-kernel path. (1) Everything one stage writes and a later stage of ANOTHER block reads is
// separated by a coop_barrier (release-reduction arrive, acquire poll) and read with __ldcg where the reading block may
// hold an older copy of the line in its L1 (MapState fields, sort buffers, raw / ds). (2) The LM loop afterwards reads
// `sorted` and `cells` through the read-only path (__ldg) although this launch wrote them: no SM touches those lines
// with a load before the barrier that follows their last write, and L1 is invalidated between launches, so no stale
// copy can exist — the single-kernel path is compared bit for bit with the stand-alone kernels in the tests.
#pragma once
#include "common.cuh"
#include "map_build.cuh"

namespace ll {

constexpr int GI_THREADS = 512;
constexpr int GI_WARPS = GI_THREADS / 32;
constexpr int GI_MAX_ITEMS = 16;                 // points per thread: 4, 8 or 16 (one tile per block)
constexpr int GI_TILE = GI_THREADS * 4;          // host-side unit for the capacity bound
constexpr int GI_DIGIT_BITS = 9;                 // 512 bins = one per thread: a 26-bit voxel key takes 3 passes, not 4
constexpr int GI_RADIX = 1 << GI_DIGIT_BITS;
constexpr int GI_SMEM_WORDS = GI_WARPS * GI_RADIX + GI_RADIX + 16;
constexpr int GI_REDUCE_ITEMS = 4;               // pairs per thread and sub-tile of the inline voxel reduce (2048-pair sub-tiles)
constexpr int GI_SMEM_WORDS_REDUCE = 5 * GI_REDUCE_ITEMS * GI_THREADS + 8 + 64 + 8;

struct GridDims {
  int o[3], d[3];
  int ncells, bits;
  bool too_large;
};

// Grid geometry from the raw bounding box (a superset of the downsampled one), as grid_keys_kernel.
__device__ __forceinline__ GridDims grid_dims(const MapState* st, int cell_cap) {
  GridDims g;
  const bool empty = (__ldcg(&st->mm[0]) == 0u);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float mn = empty ? 0.f : ordered_to_float(~__ldcg(&st->mm[a]));
    const float mx = empty ? 0.f : ordered_to_float(__ldcg(&st->mm[3 + a]));
    g.o[a] = (int)floorf(mn);
    g.d[a] = (int)floorf(mx) - g.o[a] + 1;
  }
  const long long ncells64 = (long long)g.d[0] * g.d[1] * g.d[2];
  g.too_large = ncells64 > (long long)cell_cap;
  g.ncells = g.too_large ? 0 : (int)ncells64;
  g.bits = g.ncells <= 1 ? 0 : 32 - __clz((uint32_t)(g.ncells - 1));
  return g;
}

// Points per thread such that both maps fit one tile per block; 0 = they do not (the caller must not have deferred
// the build: the host bounds the sizes, kInlineGridMaxPoints).
__device__ __forceinline__ int gi_items(int n0, int n1, int nblocks) {
#pragma unroll
  for (int items = 4; items <= GI_MAX_ITEMS; items *= 2) {
    const int t = GI_THREADS * items;
    if ((n0 + t - 1) / t + (n1 + t - 1) / t <= nblocks) return items;
  }
  return 0;
}

// This block's tile: tiles of map 0 come first, then the tiles of map 1.
struct TileJob {
  bool has;
  int m, tile, ntiles, n, warp_base;
};
template <int ITEMS>
__device__ __forceinline__ TileJob gi_tile(int n0, int n1) {
  constexpr int T = GI_THREADS * ITEMS;
  const int nt0 = (n0 + T - 1) / T, nt1 = (n1 + T - 1) / T;
  TileJob J;
  const int t = blockIdx.x;
  J.has = t < nt0 + nt1;
  J.m = (J.has && t >= nt0) ? 1 : 0;
  J.tile = J.m ? t - nt0 : t;
  J.ntiles = J.m ? nt1 : nt0;
  J.n = J.m ? n1 : n0;
  J.warp_base = J.tile * T + (threadIdx.x >> 5) * (32 * ITEMS);
  return J;
}

// Stable LSD radix sort of the tiles' (key, value) pairs, GI_DIGIT_BITS bits per pass; key / val hold this block's tile (item k of
// lane l of warp w = element warp_base + k*32 + l). `active_map` = this block has a tile and its map takes part;
// bits = significant key bits of this block's map, max_bits = the larger of the two maps (grid-uniform).
// kCells = false: the result is in S.keys / S.vals[(passes & 1)] (passes = ceil(bits / GI_DIGIT_BITS)).
// kCells = true : the last pass writes M.sorted and M.cells instead (the cell sort of the search grid).
template <int ITEMS, bool kCells>
__device__ __forceinline__ void gi_sort_core(const MapBatch& B, const TileJob& J, bool active_map, int bits, int max_bits,
                                             uint32_t (&key)[ITEMS], uint32_t (&val)[ITEMS], unsigned* barrier,
                                             unsigned& epoch, uint32_t* smem) {
  uint32_t (*warp_hist)[GI_RADIX] = reinterpret_cast<uint32_t (*)[GI_RADIX]>(smem);
  uint32_t* gbase = smem + GI_WARPS * GI_RADIX;
  uint32_t* wsum = gbase + GI_RADIX;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const MapBuffers& M = B.m[J.m];
  const SortProblem& S = B.sort[J.m];
  const int nm = J.n, warp_base = J.warp_base, tile = J.tile, ntiles = J.ntiles;
  uint32_t rank[ITEMS];
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int pass = 0; pass * GI_DIGIT_BITS < max_bits; ++pass) {
    const int shift = pass * GI_DIGIT_BITS;
    const bool active = active_map && shift < bits;  // block-uniform
    const bool last = shift + GI_DIGIT_BITS >= bits;
    if (active) {
#pragma unroll
      for (int j = 0; j < GI_RADIX / 32; ++j) warp_hist[warp][j * 32 + lane] = 0;
      __syncwarp();
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const bool valid = (warp_base + k * 32 + lane) < nm;
        const uint32_t d = (key[k] >> shift) & (uint32_t)(GI_RADIX - 1);
        // (eight ballots instead of match.any were measured: no faster here, slower in rs_onesweep_kernel)
        const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : (uint32_t)GI_RADIX);
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
      if (tid < GI_RADIX) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < GI_WARPS; ++w) {
          const uint32_t c = warp_hist[w][tid];
          warp_hist[w][tid] = run;
          run += c;
        }
        S.status[(size_t)tile * GI_RADIX + tid] = run;
      }
    }
    coop_barrier(barrier, epoch);
    if (active) {
      constexpr bool kPrefetch = kCells && ITEMS <= 4;  // (more would not fit the register budget of the host kernel)
      float4 pt[kPrefetch ? ITEMS : 1];
      if (kPrefetch && last) {  // the gathers of the final pass go out before the histogram sums are waited for
#pragma unroll
        for (int k = 0; k < (kPrefetch ? ITEMS : 1); ++k)
          pt[k] = (warp_base + k * 32 + lane) < nm ? __ldcg(&M.ds[val[k]]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      uint32_t before = 0, total = 0;
      if (tid < GI_RADIX) {
        const uint32_t* hist = S.status + tid;
        for (int tt = 0; tt < ntiles; tt += 16) {  // sixteen loads in flight
          uint32_t hv[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) hv[u] = tt + u < ntiles ? __ldcg(hist + (size_t)(tt + u) * GI_RADIX) : 0u;
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            total += hv[u];
            before += (tt + u < tile) ? hv[u] : 0u;
          }
        }
      }
      uint32_t incl = total;  // exclusive scan of `total` over the digits (thread d = digit d)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      if (lane == 31 && warp < GI_RADIX / 32) wsum[warp] = incl;
      __syncthreads();
      if (tid < GI_RADIX) {
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < GI_RADIX / 32; ++w) woff += (w < warp) ? wsum[w] : 0u;
        gbase[tid] = woff + incl - total + before;
      }
      __syncthreads();
      uint32_t* __restrict__ keys_out = S.keys[(pass + 1) & 1];
      uint32_t* __restrict__ vals_out = S.vals[(pass + 1) & 1];
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        if (warp_base + k * 32 + lane < nm) {
          const uint32_t d = (key[k] >> shift) & (uint32_t)(GI_RADIX - 1);
          const uint32_t pos = gbase[d] + warp_hist[warp][d] + rank[k];
          if (kCells && last) {
            const float4 pk = kPrefetch ? pt[kPrefetch ? k : 0] : __ldcg(&M.ds[val[k]]);
            M.sorted[pos] = make_float4(pk.x, pk.y, pk.z, __uint_as_float(val[k]));
            atomicMin(reinterpret_cast<unsigned*>(&M.cells[key[k]].x), cell_entry_x(M.cell_tag, pos));
            atomicMax(reinterpret_cast<unsigned*>(&M.cells[key[k]].y), cell_entry_y(M.cell_tag, pos + 1u));
          } else {
            keys_out[pos] = key[k];
            vals_out[pos] = val[k];
          }
        }
      }
    }
    coop_barrier(barrier, epoch);
    if (active && !last) {
      const uint32_t* __restrict__ keys_in = S.keys[(pass + 1) & 1];
      const uint32_t* __restrict__ vals_in = S.vals[(pass + 1) & 1];
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const int i = warp_base + k * 32 + lane;
        key[k] = i < nm ? __ldcg(keys_in + i) : 0xffffffffu;
        val[k] = i < nm ? __ldcg(vals_in + i) : 0u;
      }
    }
  }
}

// ---- a4: the cell sort of the downsampled points -----------------------------------------------------------------
template <int ITEMS>
__device__ __forceinline__ void gi_sort_tiles(const MapBatch& B, const GridDims (&gd)[2], const int (&n)[2], unsigned* barrier,
                                              unsigned& epoch, uint32_t* smem) {
  const TileJob J = gi_tile<ITEMS>(n[0], n[1]);
  const GridDims& g = gd[J.m];
  const MapBuffers& M = B.m[J.m];
  const bool mine = J.has && !g.too_large;
  const int lane = threadIdx.x & 31;
  uint32_t key[ITEMS], val[ITEMS];
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {  // pass 0 reads the points themselves: no key array round trip
    const int i = J.warp_base + k * 32 + lane;
    key[k] = 0xffffffffu;
    val[k] = (uint32_t)i;
    if (mine && i < J.n) {
      const float4 p = __ldcg(&M.ds[i]);
      const int cx = (int)floorf(p.x) - g.o[0], cy = (int)floorf(p.y) - g.o[1], cz = (int)floorf(p.z) - g.o[2];
      key[k] = (uint32_t)((cz * g.d[1] + cy) * g.d[0] + cx);
    }
  }
  const int max_bits = max(gd[0].too_large ? 0 : gd[0].bits, gd[1].too_large ? 0 : gd[1].bits);
  gi_sort_core<ITEMS, true>(B, J, mine, g.bits, max_bits, key, val, barrier, epoch, smem);
  if (mine && g.bits == 0) {  // a single cell: the order is the identity
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int i = J.warp_base + k * 32 + lane;
      if (i < J.n) {
        const float4 p = __ldcg(&M.ds[i]);
        M.sorted[i] = make_float4(p.x, p.y, p.z, __uint_as_float((uint32_t)i));
      }
    }
    if (J.tile == 0 && threadIdx.x == 0)
      M.cells[0] = make_int2((int)cell_entry_x(M.cell_tag, 0u), (int)cell_entry_y(M.cell_tag, (uint32_t)J.n));
  }
}

// ---- a2: transform + bounding box ---------------------------------------------------------------------------------
// Body of transform_minmax_kernel for a co-resident grid: blocks stride over the 1024-point jobs.
__device__ __forceinline__ void gi_transform(const MapBatch& B, const Segment* __restrict__ segs,
                                             const BlockJob* __restrict__ jobs, int njobs, uint32_t* smem) {
  float (*s_mn)[GI_WARPS] = reinterpret_cast<float (*)[GI_WARPS]>(smem);
  float (*s_mx)[GI_WARPS] = reinterpret_cast<float (*)[GI_WARPS]>(smem + 3 * GI_WARPS);
  int* s_any = reinterpret_cast<int*>(smem + 6 * GI_WARPS);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int j = blockIdx.x; j < njobs; j += gridDim.x) {
    const BlockJob job = jobs[j];
    const Segment s = segs[job.seg];
    const MapBuffers& M = B.m[s.map];
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool any = false;
#pragma unroll
    for (int k = 0; k < MB_TILE / GI_THREADS; ++k) {
      const int i = job.start + k * GI_THREADS + tid;
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
    if (tid == 0) *s_any = 0;
    __syncthreads();
    if (any) *s_any = 1;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v0 = warp_min(mn[a]), v1 = warp_max(mx[a]);
      if (lane == 0) { s_mn[a][warp] = v0; s_mx[a][warp] = v1; }
    }
    __syncthreads();
    if (tid < 3 && *s_any) {
      float v0 = s_mn[tid][0], v1 = s_mx[tid][0];
#pragma unroll
      for (int w = 1; w < GI_WARPS; ++w) { v0 = fminf(v0, s_mn[tid][w]); v1 = fmaxf(v1, s_mx[tid][w]); }
      atomicMax(&M.st->mm[tid], ~float_to_ordered(v0));
      atomicMax(&M.st->mm[3 + tid], float_to_ordered(v1));
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && tid == 0) {
    for (int m = 0; m < B.count; ++m) {
      B.m[m].st->n_raw = B.m[m].n_raw;
      if (!B.m[m].downsample) B.m[m].st->n_ds = B.m[m].n_raw;
    }
  }
}

// ---- a3: VoxelGrid ---------------------------------------------------------------------------------------------------
template <int ITEMS>
__device__ __forceinline__ void gi_voxel(const MapBatch& B, unsigned* barrier, unsigned& epoch, uint32_t* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int n[2], bits[2];
  uint32_t sentinel[2];
  VoxelMeta vm[2];
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const MapBuffers& M = B.m[m];
    n[m] = M.downsample ? M.n_raw : 0;  // a map that is not downsampled does not take part (ds aliases raw)
    vm[m] = voxel_meta<true>(M.st, M.inv_leaf);
    sentinel[m] = vm[m].overflow ? (uint32_t)M.n_raw : (uint32_t)vm[m].total;
    bits[m] = sentinel[m] == 0 ? 0 : 32 - __clz(sentinel[m]);
    if (blockIdx.x == 0 && tid == 0 && M.downsample) {
      M.st->voxel_overflow = vm[m].overflow ? 1 : 0;
      M.st->voxel_bits = bits[m];
    }
  }
  const TileJob J = gi_tile<ITEMS>(n[0], n[1]);
  const MapBuffers& M = B.m[J.m];
  const SortProblem& S = B.sort[J.m];
  const VoxelMeta& v = vm[J.m];
  const uint32_t sent = sentinel[J.m];
  const bool mine = J.has;
  uint32_t key[ITEMS], val[ITEMS];
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {  // voxel_keys_kernel's arithmetic (PCL's), straight from the transformed point
    const int i = J.warp_base + k * 32 + lane;
    key[k] = 0xffffffffu;
    val[k] = (uint32_t)i;
    if (mine && i < J.n) {
      const float4 p = __ldcg(&M.raw[i]);
      if (v.overflow) {
        key[k] = (uint32_t)i;  // (PCL's pass-through keeps the non-finite points too)
      } else if (!finite3(p)) {
        key[k] = sent;
      } else {
        const float inv = M.inv_leaf;
        const int i0 = (int)(floorf(p.x * inv) - (float)v.min_b[0]);
        const int i1 = (int)(floorf(p.y * inv) - (float)v.min_b[1]);
        const int i2 = (int)(floorf(p.z * inv) - (float)v.min_b[2]);
        key[k] = (uint32_t)(i0 * v.mul[0] + i1 * v.mul[1] + i2 * v.mul[2]);
      }
    }
  }
  const int max_bits = max(n[0] ? bits[0] : 0, n[1] ? bits[1] : 0);
  gi_sort_core<ITEMS, false>(B, J, mine, bits[J.m], max_bits, key, val, barrier, epoch, smem);
  // ---- one centroid per voxel, in ascending voxel index. The sorted pairs are in buffer (passes & 1).
  const int fin = ((bits[J.m] + GI_DIGIT_BITS - 1) / GI_DIGIT_BITS) & 1;
  const uint32_t* __restrict__ keys = S.keys[fin];
  const uint32_t* __restrict__ vals = S.vals[fin];
  uint32_t* s_w = smem;  // [GI_WARPS] heads per warp, then scratch for the block sums
  uint32_t hk[ITEMS];
  unsigned bal[ITEMS];
  uint32_t warp_heads = 0;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int i = J.warp_base + k * 32 + lane;
    bool head = false;
    hk[k] = 0;
    if (mine && i < J.n) {
      hk[k] = __ldcg(keys + i);
      const bool boundary = i == 0 || __ldcg(keys + i - 1) != hk[k];
      head = boundary && hk[k] != sent;  // sentinel-keyed (non-finite) pairs trail the sorted array and are dropped
    }
    bal[k] = __ballot_sync(0xffffffffu, head);
    warp_heads += __popc(bal[k]);
  }
  if (lane == 0) s_w[warp] = warp_heads;
  __syncthreads();
  uint32_t warp_off = 0, tile_heads = 0;
#pragma unroll
  for (int w = 0; w < GI_WARPS; ++w) {
    const uint32_t c = s_w[w];
    warp_off += w < warp ? c : 0u;
    tile_heads += c;
  }
  if (mine && tid == 0) S.status[J.tile] = tile_heads;
  coop_barrier(barrier, epoch);
  if (!mine) return;
  // heads in the tiles before this one, and in all tiles of this map (= laserCloud*FromMapDSNum)
  uint32_t before = 0, total = 0;
  for (int tt = tid; tt < J.ntiles; tt += GI_THREADS) {
    const uint32_t c = __ldcg(S.status + tt);
    total += c;
    before += tt < J.tile ? c : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    before += __shfl_xor_sync(0xffffffffu, before, o);
    total += __shfl_xor_sync(0xffffffffu, total, o);
  }
  __syncthreads();  // (s_w is reused)
  if (lane == 0) { s_w[warp] = before; s_w[GI_WARPS + warp] = total; }
  __syncthreads();
  before = 0; total = 0;
#pragma unroll
  for (int w = 0; w < GI_WARPS; ++w) { before += s_w[w]; total += s_w[GI_WARPS + w]; }
  if (J.tile == 0 && tid == 0) M.st->n_ds = (int)total;
  // Sub-tiles of 2048 sorted pairs, organised as voxel_reduce_kernel: the points are gathered into shared memory by all
  // threads (independent loads), then every head's thread adds its own segment up IN SORTED ORDER (= ascending
  // original index, the canonical order) from shared memory; a segment that runs past the sub-tile is finished by
  // the whole block, 512 pairs at a time.
  constexpr int R = GI_REDUCE_ITEMS;   // thread t owns the consecutive pairs R t .. R t + R - 1 of the sub-tile
  constexpr int SUB = R * GI_THREADS;
  float4* s_pts = reinterpret_cast<float4*>(smem);                       // [SUB]
  int* s_head_pos = reinterpret_cast<int*>(smem + 4 * SUB);              // [SUB + 1]
  uint32_t* s_wsum = smem + 4 * SUB + SUB + 8;                           // [GI_WARPS]
  int* s_misc = reinterpret_cast<int*>(smem + 4 * SUB + SUB + 8 + 32);   // sentinel position, cross-segment state
  float4* s_cross_sum = reinterpret_cast<float4*>(smem + 4 * SUB + SUB + 8 + 64);
  static_assert(4 * SUB + SUB + 8 + 64 + 4 <= GI_SMEM_WORDS_REDUCE, "shared-memory budget of the inline voxel reduce");
  uint32_t slot_base = before;
  constexpr int T = GI_THREADS * ITEMS;
  const int tile_begin = J.tile * T, tile_end = min(J.n, tile_begin + T);
  for (int base = tile_begin; base < tile_end; base += SUB) {
    const int sub_n = min(SUB, tile_end - base);
    __syncthreads();  // (shared memory of the previous sub-tile / of the head count is reused)
    if (tid == 0) { s_misc[0] = SUB; s_misc[1] = 0; }  // [0] first sentinel-keyed pair, [1] a segment crosses the end
    __syncthreads();
    const int l0 = R * tid, i0 = base + l0;
    uint32_t k2[R], v2[R];
    bool head[R];
    uint32_t prev = (i0 > 0 && l0 < sub_n) ? __ldcg(keys + i0 - 1) : 0xffffffffu;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const bool in = l0 + k < sub_n;
      k2[k] = in ? __ldcg(keys + i0 + k) : 0xffffffffu;
      v2[k] = in ? __ldcg(vals + i0 + k) : 0u;
    }
#pragma unroll
    for (int k = 0; k < R; ++k)
      s_pts[l0 + k] = l0 + k < sub_n ? __ldcg(&M.raw[v2[k]]) : make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const bool in = l0 + k < sub_n;
      const bool boundary = in && (i0 + k == 0 || k2[k] != prev);
      head[k] = boundary && k2[k] != sent;
      if (boundary && k2[k] == sent) s_misc[0] = l0 + k;  // at most one such pair (the array is sorted)
      prev = k2[k];
      cnt += head[k] ? 1u : 0u;
    }
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, sub_total = 0;
#pragma unroll
    for (int w = 0; w < GI_WARPS; ++w) {
      const uint32_t c = s_wsum[w];
      woff += w < warp ? c : 0u;
      sub_total += c;
    }
    const uint32_t local0 = woff + incl - cnt;
    {
      uint32_t ls = local0;
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (head[k]) s_head_pos[ls++] = l0 + k;
    }
    if (tid == 0) s_head_pos[sub_total] = sub_n;
    __syncthreads();
    const int sent_pos = s_misc[0];
    uint32_t ls = local0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (!head[k]) continue;
      const int beg = l0 + k;
      const int end = min(s_head_pos[ls + 1], sent_pos);  // sentinel-keyed pairs only ever trail the last head
      const uint32_t slot = slot_base + ls;
      ++ls;
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      int jj = beg;
      for (; jj + 4 <= end; jj += 4) {  // loads first, then the four dependent adds per component
        const float4 p0 = s_pts[jj], p1 = s_pts[jj + 1], p2 = s_pts[jj + 2], p3 = s_pts[jj + 3];
        sx += p0.x; sy += p0.y; sz += p0.z; si += p0.w;
        sx += p1.x; sy += p1.y; sz += p1.z; si += p1.w;
        sx += p2.x; sy += p2.y; sz += p2.z; si += p2.w;
        sx += p3.x; sy += p3.y; sz += p3.z; si += p3.w;
      }
      for (; jj < end; ++jj) {
        const float4 p = s_pts[jj];
        sx += p.x; sy += p.y; sz += p.z; si += p.w;
      }
      const int c = end - beg;
      if (end == sub_n && sent_pos >= sub_n && base + sub_n < J.n) {
        // the segment may continue behind this sub-tile: park the running sums, the whole block finishes it below
        *s_cross_sum = make_float4(sx, sy, sz, si);
        s_misc[2] = c;
        s_misc[3] = (int)slot;
        s_misc[4] = (int)k2[k];
        s_misc[1] = 1;
        continue;
      }
      const float fc = (float)c;
      M.ds[slot] = make_float4(sx / fc, sy / fc, sz / fc, si / fc);
    }
    slot_base += sub_total;
    __syncthreads();
    if (!s_misc[1]) continue;  // block-uniform
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    int c = 0;
    if (tid == 0) {
      sx = s_cross_sum->x; sy = s_cross_sum->y; sz = s_cross_sum->z; si = s_cross_sum->w;
      c = s_misc[2];
    }
    const uint32_t ck = (uint32_t)s_misc[4];
    const uint32_t cslot = (uint32_t)s_misc[3];
    for (int g0 = base + sub_n; g0 < J.n; g0 += GI_THREADS) {
      const int g = g0 + tid;
      const bool match = g < J.n && __ldcg(keys + g) == ck;
      if (tid == 0) s_misc[5] = GI_THREADS;
      __syncthreads();
      if (!match) atomicMin(&s_misc[5], tid);
      __syncthreads();
      const int mcount = s_misc[5];  // pairs [g0, g0 + mcount) still belong to the segment
      if (tid < mcount) s_pts[tid] = __ldcg(&M.raw[__ldcg(vals + g)]);
      __syncthreads();
      if (tid == 0) {
        for (int u = 0; u < mcount; ++u) {
          const float4 p = s_pts[u];
          sx += p.x; sy += p.y; sz += p.z; si += p.w;
        }
        c += mcount;
      }
      __syncthreads();
      if (mcount < GI_THREADS) break;
    }
    if (tid == 0) {
      const float fc = (float)c;
      M.ds[cslot] = make_float4(sx / fc, sy / fc, sz / fc, si / fc);
    }
  }
}

// Rows a1-a3 for the two maps of B: raw clouds (transformed, concatenated), VoxelGrid. Every thread of every block of
// the co-resident grid must call; returns with a grid barrier passed.
// Returns true when the cell tables of the search grids were cleared on the side (see below).
__device__ __forceinline__ bool map_build_inline(const MapBatch& B, const Segment* segs, const BlockJob* jobs, int njobs,
                                                 unsigned* barrier, unsigned& epoch, uint32_t* smem,
                                                 bool may_clear_cells = true) {
  gi_transform(B, segs, jobs, njobs, smem);
  coop_barrier(barrier, epoch);
  const int n0 = B.m[0].downsample ? B.m[0].n_raw : 0, n1 = B.m[1].downsample ? B.m[1].n_raw : 0;
  const int items = gi_items(n0, n1, (int)gridDim.x);  // > 0: the host bounds the raw size
  (void)may_clear_cells;  // (the cell tables are stamped, not cleared: map_build.cuh)
  const bool side_clear = false;
  if (items == 4) gi_voxel<4>(B, barrier, epoch, smem);
  else if (items == 8) gi_voxel<8>(B, barrier, epoch, smem);
  else if (items == 16) gi_voxel<16>(B, barrier, epoch, smem);
  coop_barrier(barrier, epoch);
  return side_clear;
}

// Builds `sorted` and `cells` of the two maps of B. Every thread of every block of the co-resident grid must call.
// smem: GI_SMEM_WORDS uint32. Returns with a grid barrier passed (results visible to every block).
__device__ __forceinline__ void grid_build_inline(const MapBatch& B, unsigned* barrier, unsigned& epoch, uint32_t* smem,
                                                  bool cells_cleared = false) {
  const int gtid = blockIdx.x * GI_THREADS + threadIdx.x;
  GridDims gd[2];
  int n[2];
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    gd[m] = grid_dims(B.m[m].st, B.m[m].cell_cap);
    n[m] = __ldcg(&B.m[m].st->n_ds);  // (possibly written by another block of this launch)
  }
  const int items = gi_items(gd[0].too_large ? 0 : n[0], gd[1].too_large ? 0 : n[1], (int)gridDim.x);
  if (items == 0) {  // cannot happen within the host's bound; refuse rather than build a wrong grid
    gd[0].too_large = true;
    gd[1].too_large = true;
  }
  // geometry + cell-table clear ("empty" = start INT_MAX, end 0, so that the final pass can use atomicMin / atomicMax)
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const MapBuffers& M = B.m[m];
    const GridDims& g = gd[m];
    if (gtid == 0) {
      MapState* st = M.st;
      st->gox = g.o[0]; st->goy = g.o[1]; st->goz = g.o[2];
      st->gdx = g.d[0]; st->gdy = g.d[1]; st->gdz = g.d[2];
      st->ncells = g.ncells;
      st->grid_too_large = g.too_large ? 1 : 0;
      st->grid_bits = g.bits;
      st->cell_tag = M.cell_tag;
    }
    (void)cells_cleared;  // (no clear: stamped entries)
  }
  if (gd[0].bits == 0 || gd[1].bits == 0)  // (warp-, block- and grid-uniform)
    coop_barrier(barrier, epoch);  // a single-cell map writes its cell entry without an intervening pass barrier
  if (items == 4) gi_sort_tiles<4>(B, gd, n, barrier, epoch, smem);
  else if (items == 8) gi_sort_tiles<8>(B, gd, n, barrier, epoch, smem);
  else if (items == 16) gi_sort_tiles<16>(B, gd, n, barrier, epoch, smem);
  coop_barrier(barrier, epoch);
}

// Rows a2-a3 alone (ll_voxel_downsample on a cloud of up to 2^20 points: downsampleCurrentScan's six filters): the same
// in-kernel pipeline as the head of the scan kernel, as ONE cooperative launch instead of seven launches.
// VoxelGrid of small clouds (<= VS_MAX points: the corner clouds and outliers of downsampleCurrentScan, the key poses
// of extractSurroundingKeyFrames) entirely in one block's shared memory, one block per cloud: the multi-block
// pipeline above is a chain of ~30 grid-wide round trips, which for a few hundred points is all it costs (27 us).
// Same arithmetic, same order: bounding box of the finite points, PCL's voxel index, (voxel, original index) pairs
// sorted — a bitonic network on the 64-bit composite, whose keys are unique, so the order is the stable one — and
// one thread per voxel adding its points up in that order.
constexpr int VS_THREADS = 512;
constexpr int VS_MAX = 1024;

__global__ void __launch_bounds__(VS_THREADS, 1) voxel_small_kernel(const __grid_constant__ MapBatch B,
                                                                   const Segment* __restrict__ segs, int n_segs) {
  const MapBuffers& M = B.m[blockIdx.x];
  __shared__ float4 s_pts[VS_MAX];
  __shared__ unsigned long long s_key[VS_MAX];
  __shared__ int s_head_pos[VS_MAX + 1];
  __shared__ MapState s_st;
  __shared__ uint32_t s_warp_sum[VS_THREADS / 32];
  __shared__ int s_valid;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = M.n_raw;
  if (tid < (int)(sizeof(MapState) / 4)) reinterpret_cast<uint32_t*>(&s_st)[tid] = 0u;
  if (tid == 0) s_valid = 0;
  __syncthreads();
  for (int sg = 0; sg < n_segs; ++sg) {  // a2 for this cloud's pieces (identity for clouds already in their frame)
    const Segment& g = segs[sg];
    if (g.map != (int)blockIdx.x) continue;
    for (int i = tid; i < g.count; i += VS_THREADS) {
      const float4 p = __ldg(g.src + i);
      s_pts[g.dst_off + i] = g.identity ? p : rigid_transform(g.t, p);
    }
  }
  __syncthreads();
  {  // bounding box of the finite points (MapState encoding, see transform_minmax_kernel)
    uint32_t lo[3] = {0u, 0u, 0u}, hi[3] = {0u, 0u, 0u};
    int valid = 0;
    for (int i = tid; i < n; i += VS_THREADS) {
      const float4 q = s_pts[i];
      if (finite3(q)) {
        ++valid;
        lo[0] = max(lo[0], ~float_to_ordered(q.x)); hi[0] = max(hi[0], float_to_ordered(q.x));
        lo[1] = max(lo[1], ~float_to_ordered(q.y)); hi[1] = max(hi[1], float_to_ordered(q.y));
        lo[2] = max(lo[2], ~float_to_ordered(q.z)); hi[2] = max(hi[2], float_to_ordered(q.z));
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lo[a] = __reduce_max_sync(0xffffffffu, lo[a]);
      hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
    }
    valid = __reduce_add_sync(0xffffffffu, valid);
    if (lane == 0 && valid) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        atomicMax(&s_st.mm[a], lo[a]);
        atomicMax(&s_st.mm[3 + a], hi[a]);
      }
      atomicAdd(&s_valid, valid);
    }
  }
  __syncthreads();
  const VoxelMeta v = voxel_meta(&s_st, M.inv_leaf);
  const uint32_t sentinel = v.overflow ? (uint32_t)n : (uint32_t)v.total;
  int N = 64;
  while (N < n) N <<= 1;
  {
    const float inv = M.inv_leaf;
    for (int i = tid; i < N; i += VS_THREADS) {
      unsigned long long ck = 0xffffffffffffffffull;  // padding sorts behind everything
      if (i < n) {
        const float4 q = s_pts[i];
        uint32_t key;
        if (v.overflow) {
          key = (uint32_t)i;  // PCL's "leaf size too small": the input passes through (non-finite points too)
        } else if (!finite3(q)) {
          key = sentinel;  // behind every voxel, dropped (PCL skips non-finite points)
        } else {
          const int i0 = (int)(floorf(q.x * inv) - (float)v.min_b[0]);
          const int i1 = (int)(floorf(q.y * inv) - (float)v.min_b[1]);
          const int i2 = (int)(floorf(q.z * inv) - (float)v.min_b[2]);
          key = (uint32_t)(i0 * v.mul[0] + i1 * v.mul[1] + i2 * v.mul[2]);
        }
        ck = ((unsigned long long)key << 32) | (uint32_t)i;
      }
      s_key[i] = ck;
    }
  }
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int t = tid; t < (N >> 1); t += VS_THREADS) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const unsigned long long a = s_key[i], b = s_key[l];
        const bool up = (i & k) == 0;
        if ((a > b) == up) {
          s_key[i] = b;
          s_key[l] = a;
        }
      }
    }
  }
  __syncthreads();
  // segment heads -> output slots: every thread owns two consecutive sorted positions
  const int n_valid = v.overflow ? n : s_valid;  // sorted positions [0, n_valid) hold the finite points (pass-through: all)
  bool head[2];
  uint32_t cnt = 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int pos = 2 * tid + k;
    head[k] = pos < n_valid && (pos == 0 || (uint32_t)(s_key[pos] >> 32) != (uint32_t)(s_key[pos - 1] >> 32));
    cnt += head[k] ? 1u : 0u;
  }
  uint32_t incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp_sum[warp] = incl;
  __syncthreads();
  uint32_t off = incl - cnt, total = 0;
#pragma unroll
  for (int w = 0; w < VS_THREADS / 32; ++w) {
    const uint32_t c = s_warp_sum[w];
    if (w < warp) off += c;
    total += c;
  }
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (head[k]) s_head_pos[off++] = 2 * tid + k;
  if (tid == 0) s_head_pos[total] = n_valid;
  __syncthreads();
  for (int slot = tid; slot < (int)total; slot += VS_THREADS) {  // the centroid: additions in sorted order
    const int beg = s_head_pos[slot], end = s_head_pos[slot + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (int jj = beg; jj < end; ++jj) {
      const float4 q = s_pts[(uint32_t)s_key[jj]];
      sx += q.x; sy += q.y; sz += q.z; si += q.w;
    }
    const float fc = (float)(end - beg);
    M.ds[slot] = make_float4(sx / fc, sy / fc, sz / fc, si / fc);
  }
  if (tid == 0) {
    s_st.n_raw = n;
    s_st.n_ds = (int)total;
    s_st.voxel_bits = sentinel == 0 ? 0 : 32 - __clz(sentinel);
    s_st.voxel_overflow = v.overflow ? 1 : 0;
    *M.st = s_st;
  }
}

__global__ void __launch_bounds__(GI_THREADS, 1) voxel_inline_kernel(const __grid_constant__ MapBatch B,
                                                                     const Segment* segs, const BlockJob* jobs, int n_jobs,
                                                                     unsigned* barrier, unsigned* barrier_clear) {
  __shared__ __align__(16) uint32_t smem[(GI_SMEM_WORDS > GI_SMEM_WORDS_REDUCE ? GI_SMEM_WORDS : GI_SMEM_WORDS_REDUCE)];
  unsigned epoch = 0;
  if (blockIdx.x == 0 && threadIdx.x < 128) reinterpret_cast<uint32_t*>(B.m[threadIdx.x >> 6].st)[threadIdx.x & 63] = 0u;
  coop_barrier(barrier, epoch);
  map_build_inline(B, segs, jobs, n_jobs, barrier, epoch, smem, /*may_clear_cells=*/false);
  if (blockIdx.x == 0 && threadIdx.x == 0) *barrier_clear = 0u;
}

}  // namespace ll
