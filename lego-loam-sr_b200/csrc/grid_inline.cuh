// The 1 m search grids built INSIDE the persistent scan kernel, by its co-resident grid of
// 512-thread blocks, instead of five separate launches (grid_keys, three sort passes, grid_finalize). Every block
// owns ONE tile of 4/8/16 x 512 downsampled points (both maps share the grid) and keeps it in registers:
//   prologue  geometry, cell-table clear, cell key of every point straight from the point (no key array)
//   pass p    rank the digits, publish the tile's 256-bin histogram                    | grid barrier
//             sum the histograms (tiles before this one / all tiles), digit prefix, scatter (key, value);
//             the LAST pass writes the cell-ordered point itself and extends its cell's [start, end) with
//             atomicMin / atomicMax, so there is no finalize stage                      | grid barrier
// With ~60 k points: 6 barriers (~1.3 us each) plus ~1.5 us of work per pass, against five kernel boundaries and
// three decoupled look-back chains. The sort is the same stable LSD radix sort as rs_onesweep_kernel (ranks by warp
// match_any in memory order), so `sorted` is bit-identical to what the separate kernels produce; `cells` marks an
// empty cell as (INT_MAX, 0) instead of (0, 0) — both read as "end <= start" by the search.
#pragma once
#include "common.cuh"
#include "map_build.cuh"

namespace ll {

constexpr int GI_THREADS = 512;
constexpr int GI_WARPS = GI_THREADS / 32;
constexpr int GI_MAX_ITEMS = 16;                 // points per thread: 4, 8 or 16 (one tile per block)
constexpr int GI_TILE = GI_THREADS * 4;          // host-side unit for the capacity bound
constexpr int GI_SMEM_WORDS = GI_WARPS * RS_RADIX + RS_RADIX + 16;

struct GridDims {
  int o[3], d[3];
  int ncells, bits;
  bool too_large;
};

// Grid geometry from the raw bounding box (a superset of the downsampled one), as grid_keys_kernel.
__device__ __forceinline__ GridDims grid_dims(const MapState* st, int cell_cap) {
  GridDims g;
  const bool empty = (st->mm[0] == 0u);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float mn = empty ? 0.f : ordered_to_float(~st->mm[a]);
    const float mx = empty ? 0.f : ordered_to_float(st->mm[3 + a]);
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

// One map's worth of work for this block: its tile of ITEMS * 512 points stays in registers across the passes.
//   per pass   rank the digits (warp match_any, memory order => stable), publish the tile's 256-bin histogram
//              | grid barrier |
//              add up the histograms of the tiles before this one and of all tiles, exclusive digit prefix,
//              scatter (key, value) — or, in the last pass, write the cell-ordered point itself and extend its
//              cell's [start, end) with atomicMin / atomicMax
//              | grid barrier | reload the tile
template <int ITEMS>
__device__ __forceinline__ void gi_sort_tiles(const MapBatch& B, const GridDims (&gd)[2], const int (&n)[2], unsigned* barrier,
                                              unsigned& epoch, uint32_t* smem) {
  uint32_t (*warp_hist)[RS_RADIX] = reinterpret_cast<uint32_t (*)[RS_RADIX]>(smem);
  uint32_t* gbase = smem + GI_WARPS * RS_RADIX;
  uint32_t* wsum = gbase + RS_RADIX;
  constexpr int T = GI_THREADS * ITEMS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntiles0 = (n[0] + T - 1) / T, ntiles1 = (n[1] + T - 1) / T;
  const int t = blockIdx.x;
  const bool has = t < ntiles0 + ntiles1;
  const int m = (has && t >= ntiles0) ? 1 : 0;
  const int tile = m ? t - ntiles0 : t;
  const int ntiles = m ? ntiles1 : ntiles0;
  const int nm = n[m];
  const GridDims& g = gd[m];
  const MapBuffers& M = B.m[m];
  const SortProblem& S = B.sort[m];
  const bool mine = has && !g.too_large;
  const int warp_base = tile * T + warp * (32 * ITEMS);
  uint32_t key[ITEMS], val[ITEMS], rank[ITEMS];
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {  // pass 0 reads the points themselves: no key array round trip
    const int i = warp_base + k * 32 + lane;
    key[k] = 0xffffffffu;
    val[k] = (uint32_t)i;
    if (mine && i < nm) {
      const float4 p = M.ds[i];
      const int cx = (int)floorf(p.x) - g.o[0], cy = (int)floorf(p.y) - g.o[1], cz = (int)floorf(p.z) - g.o[2];
      key[k] = (uint32_t)((cz * g.d[1] + cy) * g.d[0] + cx);
    }
  }
  const int max_bits = max(gd[0].too_large ? 0 : gd[0].bits, gd[1].too_large ? 0 : gd[1].bits);
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int pass = 0; pass * 8 < max_bits; ++pass) {
    const int shift = pass * 8;
    const bool active = mine && shift < g.bits;  // block-uniform
    const bool last = shift + 8 >= g.bits;
    if (active) {
#pragma unroll
      for (int j = 0; j < RS_RADIX / 32; ++j) warp_hist[warp][j * 32 + lane] = 0;
      __syncwarp();
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const bool valid = (warp_base + k * 32 + lane) < nm;
        const uint32_t d = (key[k] >> shift) & 255u;
        const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : 256u);
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
      if (tid < RS_RADIX) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < GI_WARPS; ++w) {
          const uint32_t c = warp_hist[w][tid];
          warp_hist[w][tid] = run;
          run += c;
        }
        S.status[(size_t)tile * RS_RADIX + tid] = run;
      }
    }
    coop_barrier(barrier, epoch);
    if (active) {
      constexpr bool kPrefetch = ITEMS <= 4;  // (more would not fit the register budget of the host kernel)
      float4 pt[kPrefetch ? ITEMS : 1];
      if (kPrefetch && last) {  // the gathers of the final pass go out before the histogram sums are waited for
#pragma unroll
        for (int k = 0; k < (kPrefetch ? ITEMS : 1); ++k)
          pt[k] = (warp_base + k * 32 + lane) < nm ? M.ds[val[k]] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      uint32_t before = 0, total = 0;
      if (tid < RS_RADIX) {
        const uint32_t* hist = S.status + tid;
        for (int tt = 0; tt < ntiles; tt += 16) {  // sixteen loads in flight
          uint32_t hv[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) hv[u] = tt + u < ntiles ? __ldcg(hist + (size_t)(tt + u) * RS_RADIX) : 0u;
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            total += hv[u];
            before += (tt + u < tile) ? hv[u] : 0u;
          }
        }
      }
      uint32_t incl = total;  // exclusive scan of `total` over the 256 digits (threads 0..255 = warps 0..7)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      if (lane == 31 && warp < RS_RADIX / 32) wsum[warp] = incl;
      __syncthreads();
      if (tid < RS_RADIX) {
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < RS_RADIX / 32; ++w) woff += (w < warp) ? wsum[w] : 0u;
        gbase[tid] = woff + incl - total + before;
      }
      __syncthreads();
      uint32_t* __restrict__ keys_out = S.keys[(pass + 1) & 1];
      uint32_t* __restrict__ vals_out = S.vals[(pass + 1) & 1];
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        if (warp_base + k * 32 + lane < nm) {
          const uint32_t d = (key[k] >> shift) & 255u;
          const uint32_t pos = gbase[d] + warp_hist[warp][d] + rank[k];
          if (last) {
            const float4 pk = kPrefetch ? pt[kPrefetch ? k : 0] : M.ds[val[k]];
            M.sorted[pos] = make_float4(pk.x, pk.y, pk.z, __uint_as_float(val[k]));
            atomicMin(&M.cells[key[k]].x, (int)pos);
            atomicMax(&M.cells[key[k]].y, (int)pos + 1);
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
  if (mine && g.bits == 0) {  // a single cell: the order is the identity
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int i = warp_base + k * 32 + lane;
      if (i < nm) {
        const float4 p = M.ds[i];
        M.sorted[i] = make_float4(p.x, p.y, p.z, __uint_as_float((uint32_t)i));
      }
    }
    if (tile == 0 && tid == 0) M.cells[0] = make_int2(0, nm);
  }
}

// Builds `sorted` and `cells` of the two maps of B. Every thread of every block of the co-resident grid must call.
// smem: GI_SMEM_WORDS uint32. Returns with a grid barrier passed (results visible to every block).
__device__ __forceinline__ void grid_build_inline(const MapBatch& B, unsigned* barrier, unsigned& epoch, uint32_t* smem) {
  const int gtid = blockIdx.x * GI_THREADS + threadIdx.x, gthreads = gridDim.x * GI_THREADS;
  GridDims gd[2];
  int n[2];
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    gd[m] = grid_dims(B.m[m].st, B.m[m].cell_cap);
    n[m] = B.m[m].st->n_ds;
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
    }
    for (int c = gtid; c < g.ncells; c += gthreads) M.cells[c] = make_int2(0x7fffffff, 0);
  }
  if (gd[0].bits == 0 || gd[1].bits == 0)  // (warp-, block- and grid-uniform)
    coop_barrier(barrier, epoch);  // a single-cell map writes its cell entry without an intervening pass barrier
  if (items == 4) gi_sort_tiles<4>(B, gd, n, barrier, epoch, smem);
  else if (items == 8) gi_sort_tiles<8>(B, gd, n, barrier, epoch, smem);
  else if (items == 16) gi_sort_tiles<16>(B, gd, n, barrier, epoch, smem);
  coop_barrier(barrier, epoch);
}

}  // namespace ll
