// Incremental assembly of the local map for a SLIDING set of keyframes.
//
// The reference re-filters the whole surrounding cloud every frame (mapOptmization.cpp:1218-1231) although, frame to
// frame, the set of keyframes only loses a few members and gains the newest one (deque pop / push :1127-1144, erase /
// push_back :1171-1216) and their transformed clouds are cached (:1124-1144, :1193-1215). What makes an incremental
// VoxelGrid possible with BIT-IDENTICAL output:
//   * pcl::VoxelGrid emits one centroid per occupied voxel in ascending idx = i + j*dx + k*dx*dy, i.e. in lexicographic
//     (k, j, i) order of the voxel coordinates relative to the bounding box — which is the lexicographic order of the
//     ABSOLUTE coordinates (floor(z/leaf), floor(y/leaf), floor(x/leaf)), whatever the box is;
//   * the canonical accumulation order inside a voxel is the order of the points in the concatenated cloud (oracle and
//     full path: stable sort), i.e. keyframes in list order, per keyframe corner | surf then outlier, then point index.
// So the window W of a map is kept on the device as ONE array of (absolute voxel key, keyframe tag, point) sorted by
// (key, list order). A frame then needs:
//   R  the added keyframes' clouds: transform, key, RANK (every block holds all composites in shared memory and ranks  ┐ one
//      its own slice of the points: inc_build_run)                                                                    │
//   C  the old window: count the entries whose keyframe stays, per tile (its loads are issued before R)               ┘ barrier
//   M  stable merge of the surviving entries with the new run into the other buffer (old before new on equal keys,
//      exactly the order a rebuild would produce), bounding box of what is written                                    barrier
//   V  segment heads per tile | barrier | ordered centroid sums -> laserCloud*FromMapDS, exactly as the full path     barrier
// instead of transform (all 50 keyframes) + 3 radix passes over ~0.6 M pairs + reduce. Anything the scheme cannot
// represent (PCL's int32-overflow pass-through, voxel coordinates beyond +-2^20, an added point more than 2048 voxels
// from the added keyframe's position) raises MapState::inc_fail; the host then rebuilds through the full path (abi.cu).
// A longer list of added keyframes (the default branch at start-up, a rebuild after a pose correction) is fed in groups
// of <= INC_RUN_MAX points per map, one merge each (map_inc_kernel); only the last group computes the centroids.
#pragma once
#include "common.cuh"
#include "map_build.cuh"
#include "map_inline.cuh"

namespace ll {

constexpr int INC_RUN_MAX = 16384;  // points per map the keyframes added in one merge may hold (one bitonic network)
constexpr int INC_RUN_BITS = 14;
constexpr int INC_MAX_DEAD = 8;     // keyframes dropped in one merge
constexpr int INC_MAX_SEGS = 12;    // clouds of the added keyframes (three per keyframe)
constexpr int INC_SUB = 2560;       // window entries per sub-tile of the merge / centroid pass (five per thread); a block cuts its tile into EQUAL sub-tiles
constexpr int INC_MAX_BLOCKS = 160;
constexpr unsigned long long INC_KEY_NONFINITE = ~0ull;
constexpr int INC_SMEM_BYTES = INC_RUN_MAX * 8;  // the run's sort network; the other stages need less

struct WindowBuffers {
  float4* pts[2];
  unsigned long long* key[2];
  uint32_t* tag[2];
  float4* run_pts;              // the added keyframes' points, sorted by (key, order) — [INC_RUN_MAX]
  unsigned long long* run_key;
  uint32_t* run_tag;
  float4* stage_pts;            // their transformed points in arrival order — [INC_RUN_MAX]
  uint32_t* stage_tag;
  uint32_t* counts;             // [2 * INC_MAX_BLOCKS]: surviving entries per merge tile | heads per reduce tile
};

struct IncJob {
  WindowBuffers w[2];
  int cur;                       // buffer that holds the window now; the merged window goes to cur ^ 1
  int n_old[2];                  // entries in the window now
  int n_new[2];                  // points of the added keyframes
  int n_out[2];                  // entries after the merge (host-known: every raw point stays an entry)
  int n_dead;
  uint32_t dead[INC_MAX_DEAD];   // tags (keyframe ids) to drop
  int n_segs;
  Segment segs[INC_MAX_SEGS];    // Segment::pad[0] = tag
  int run_org[2][3];             // per map: voxel coordinates the added points' composites are relative to (inc_build_run)
  int final_step;                // 1 = compute laserCloud*FromMapDS after the merge
  int keep_fail;                 // 1 = the window this merge starts from comes from a merge whose verdict the host has not
                                 //     read yet: MapState::inc_fail survives the zeroing at the head of the launch
};

// Zero both MapStates at the head of a launch that assembles the map (threads 0..127 of one block), keeping inc_fail when asked.
__device__ __forceinline__ void inc_zero_map_states(const MapBatch& B, bool keep_fail) {
  const int t = threadIdx.x;
  if (t < 128) {
    constexpr int kFailWord = (int)(offsetof(MapState, inc_fail) / 4);
    static_assert(sizeof(MapState) <= 256 && offsetof(MapState, inc_fail) % 4 == 0, "MapState is zeroed as 64 words");
    if (!(keep_fail && (t & 63) == kFailWord)) reinterpret_cast<uint32_t*>(B.m[t >> 6].st)[t & 63] = 0u;
  }
}

// Absolute voxel coordinates of a point, packed (z | y | x, 21 bits each, offset 2^20): ascending key = VoxelGrid order.
__device__ __forceinline__ unsigned long long inc_abs_key(const float4 p, float inv, bool& ok) {
  if (!finite3(p)) return INC_KEY_NONFINITE;
  const float fx = floorf(p.x * inv), fy = floorf(p.y * inv), fz = floorf(p.z * inv);
  if (!(fabsf(fx) < 1048576.f && fabsf(fy) < 1048576.f && fabsf(fz) < 1048576.f)) {
    ok = false;
    return INC_KEY_NONFINITE - 1ull;
  }
  const unsigned long long ix = (unsigned long long)((int)fx + 1048576), iy = (unsigned long long)((int)fy + 1048576),
                           iz = (unsigned long long)((int)fz + 1048576);
  return (iz << 42) | (iy << 21) | ix;
}

// First index in [lo, hi) of the ascending array a whose element is >= v (kUpper: > v). kGlobal: the array is in
// global memory and may have been written earlier in this launch by another block (read through L2).
template <bool kUpper, bool kGlobal>
__device__ __forceinline__ int inc_bound(const unsigned long long* a, int lo, int hi, unsigned long long v) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const unsigned long long x = kGlobal ? __ldcg(a + mid) : a[mid];
    if (kUpper ? (x <= v) : (x < v)) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// lower_bound over an ascending array in global memory by the WHOLE block: every thread probes one position per round
// (two rounds cover 2^18 entries with 512 threads) and __syncthreads_count tells how many probes lie below v — two L2
// round trips instead of the fourteen dependent ones of a one-thread binary search (three such searches sat at the head
// of every merge tile). All threads of the block must call; the result is block-uniform.
__device__ __forceinline__ int inc_bound_block(const unsigned long long* a, int lo, int hi, unsigned long long v) {
  const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
#pragma unroll 1
  while (hi - lo > 0) {
    const int len = hi - lo;
    // probe positions split [lo, hi) into nt + 1 nearly equal pieces; with len <= nt every element is probed
    const int p = len <= nt ? lo + tid : lo + (int)(((long long)len * (tid + 1)) / (nt + 1));
    const bool below = p < hi && __ldcg(a + p) < v;
    const int c = __syncthreads_count(below);  // probes are ascending, so the first c of them are the ones below v
    if (len <= nt) return lo + c;
    const int new_lo = c == 0 ? lo : lo + (int)(((long long)len * c) / (nt + 1)) + 1;
    const int new_hi = c == nt ? hi : lo + (int)(((long long)len * (c + 1)) / (nt + 1));
    lo = new_lo;
    hi = new_hi;
  }
  return lo;
}

// Which blocks merge which map: workers [first, nb) are split between the two maps in proportion to their windows.
struct IncTiling {
  int first, nw[2];
  __device__ __forceinline__ IncTiling(int nb, const int (&n)[2], bool reserve_two) {
    first = (reserve_two && nb >= 8) ? 2 : 0;
    const int w = nb - first;
    const long long tot = (long long)n[0] + n[1];
    if (w < 2) {
      nw[0] = nw[1] = w;  // a single block takes both maps in turn (see inc_worker)
    } else {
      int a = tot > 0 ? (int)(((long long)w * n[0] + tot / 2) / tot) : w / 2;
      a = a < 1 ? 1 : (a > w - 1 ? w - 1 : a);
      nw[0] = a;
      nw[1] = w - a;
    }
  }
};

__device__ __forceinline__ bool inc_alive(uint32_t tag, const IncJob& J) {
  bool alive = true;
#pragma unroll
  for (int k = 0; k < INC_MAX_DEAD; ++k) alive = alive && !(k < J.n_dead && tag == J.dead[k]);
  return alive;
}

// Ranks of the added points [lo, hi) of map m among all n composites in s_key (shared memory), and the points to their
// places in the run. One warp per E points: the lanes stride over the keys (four loads in flight), every key is compared
// with the E composites, the counts are reduced over the warp.
template <int E>
__device__ __forceinline__ void inc_rank_slice(const IncJob& J, const MapBatch& B, int m, const unsigned long long* s_key, int n,
                                               int lo, int hi) {
  const WindowBuffers& W = J.w[m];
  const float inv = B.m[m].inv_leaf;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int e0 = lo + E * warp; e0 < hi; e0 += E * nwarps) {
    unsigned long long ke[E];
    int cnt[E];
#pragma unroll
    for (int u = 0; u < E; ++u) { ke[u] = e0 + u < hi ? s_key[e0 + u] : 0ull; cnt[u] = 0; }
    int j = lane;
    for (; j + 96 < n; j += 128) {  // four key loads in flight
      const unsigned long long k0 = s_key[j], k1 = s_key[j + 32], k2 = s_key[j + 64], k3 = s_key[j + 96];
#pragma unroll
      for (int u = 0; u < E; ++u)
        cnt[u] += (k0 < ke[u] ? 1 : 0) + (k1 < ke[u] ? 1 : 0) + (k2 < ke[u] ? 1 : 0) + (k3 < ke[u] ? 1 : 0);
    }
    for (; j < n; j += 32) {
      const unsigned long long k = s_key[j];
#pragma unroll
      for (int u = 0; u < E; ++u) cnt[u] += k < ke[u] ? 1 : 0;
    }
#pragma unroll
    for (int u = 0; u < E; ++u) cnt[u] = __reduce_add_sync(0xffffffffu, cnt[u]);
    if (lane < E && e0 + lane < hi) {
      const int e = e0 + lane;
      int rank = cnt[0];
#pragma unroll
      for (int u = 1; u < E; ++u) rank = lane == u ? cnt[u] : rank;
      // the point itself again (its cloud and offset from the segment table), to its place in the run
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t tag = 0u;
      for (int sg = 0; sg < J.n_segs; ++sg) {
        const Segment& g = J.segs[sg];
        if (g.map == m && e >= g.dst_off && e < g.dst_off + g.count) {
          const float4 p = __ldg(g.src + (e - g.dst_off));
          q = g.identity ? p : rigid_transform(g.t, p);
          tag = (uint32_t)g.pad[0];
        }
      }
      bool ok = true;
      W.run_pts[rank] = q;
      W.run_key[rank] = inc_abs_key(q, inv, ok);
      W.run_tag[rank] = tag;
    }
  }
}

// ---- R: the run of the added keyframes of map m = their points sorted by (voxel key, arrival order), by ALL blocks.
// A sorting network in one block's shared memory took 92 us for a 10 k-point keyframe (profiles/r02_*): the run is instead
// RANKED. Every block transforms all the added points itself (a few thousand: cheaper than a grid barrier between a
// transform stage and a sort stage) and keeps their 64-bit composites (voxel coordinates relative to a host-given origin,
// arrival index: unique, so the order is the stable one) in shared memory; it then owns a contiguous slice of the points and
// counts, for each of them, the composites below its own — E points per pass over the keys, one warp per E points —
// which is its position in the run. n^2 / blocks comparisons: 13-16 us for 10 k points on 148 blocks (64-bit compares, three
// instructions each; a sample-splitter bucket ranking was built and measured: not faster, profiles/r02_ab_records.txt); the
// host takes the full path when that budget is exceeded (few blocks, abi.cu). smem: INC_RUN_MAX u64.
__device__ __forceinline__ void inc_build_run(const IncJob& J, const MapBatch& B, int m, unsigned long long* s_key) {
  const MapBuffers& M = B.m[m];
  const int tid = threadIdx.x, nwarps = blockDim.x >> 5, n = J.n_new[m];
  __shared__ int s_bad;
  __syncthreads();  // (the previous map's pass is over: its keys may be overwritten)
  if (tid == 0) s_bad = 0;
  __syncthreads();
  const float inv = M.inv_leaf;
  // The composites of ALL added points: voxel coordinates relative to an origin the HOST derives from the first added
  // keyframe's position (IncJob::run_org: its voxel minus 2048 per axis), twelve bits per axis — a cloud reaches at most
  // ~130 m from its sensor, 650 voxels of 0.2 m — so that no pass over the points is needed to find their box first. A
  // point outside that cube (or beyond +-2^20 voxels) gives the merge up: the host rebuilds in full.
  const int ox = J.run_org[m][0], oy = J.run_org[m][1], oz = J.run_org[m][2];
  for (int sg = 0; sg < J.n_segs; ++sg) {
    const Segment& g = J.segs[sg];
    if (g.map != m) continue;
    for (int i0 = tid; i0 < g.count; i0 += 4 * blockDim.x) {  // four loads in flight (the clouds are cold: fresh from the host)
      float4 pv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * (int)blockDim.x;
        pv[u] = i < g.count ? __ldg(g.src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * (int)blockDim.x;
        if (i >= g.count) continue;
        const float4 q = g.identity ? pv[u] : rigid_transform(g.t, pv[u]);
        unsigned long long rel = 0xfffffffffull;  // non-finite points: behind every voxel (36 bits), in arrival order
        if (finite3(q)) {
          const float f[3] = {floorf(q.x * inv), floorf(q.y * inv), floorf(q.z * inv)};
          bool in = fabsf(f[0]) < 1048576.f && fabsf(f[1]) < 1048576.f && fabsf(f[2]) < 1048576.f;
          const int ix = (int)f[0] - ox, iy = (int)f[1] - oy, iz = (int)f[2] - oz;
          in = in && ix >= 0 && ix < 4096 && iy >= 0 && iy < 4096 && iz >= 0 && iz < 4095;  // (z = 4095 is kept for the sentinel)
          if (!in) s_bad = 1;
          rel = ((unsigned long long)(unsigned)iz << 24) | ((unsigned long long)(unsigned)iy << 12) | (unsigned long long)(unsigned)ix;
        }
        const int e = g.dst_off + i;
        s_key[e] = (rel << INC_RUN_BITS) | (unsigned long long)e;  // unique => the rank order is the stable order
      }
    }
  }
  __syncthreads();
  LL_TL_SUB(2 + 2 * m);  // composites of the added points in shared memory
  if (s_bad) {  // (the same verdict in every block: they all looked at the same points)
    if (tid == 0 && blockIdx.x == 0) M.st->inc_fail = 1;
    return;
  }
  static_assert(INC_RUN_MAX == (1 << INC_RUN_BITS), "the composite key keeps the arrival index in INC_RUN_BITS bits");
  // ranks of this block's slice [lo, hi) of the added points: E points per pass over the keys, one warp per E points, E
  // chosen so that the slice is one round of the block's warps (69 points per block at 10 k: E = 5, not two rounds of 4)
  const int per = (n + (int)gridDim.x - 1) / (int)gridDim.x;
  const int lo = min(n, (int)blockIdx.x * per), hi = min(n, lo + per);
  // (E from the LARGER of the two maps' runs: both maps then go through the same instance of the loop, whose code is
  //  fetched once — these stages run once per scan, straight after an L2 flush in the benchmark, and every instruction is a miss)
  const int per_max = (max(J.n_new[0], J.n_new[1]) + (int)gridDim.x - 1) / (int)gridDim.x;
  const int want = (per_max + nwarps - 1) / nwarps;
  if (want <= 4) inc_rank_slice<4>(J, B, m, s_key, n, lo, hi);
  else if (want <= 6) inc_rank_slice<6>(J, B, m, s_key, n, lo, hi);
  else inc_rank_slice<8>(J, B, m, s_key, n, lo, hi);
}

// ---- C: surviving entries of this worker's tile of the old window.
// (block sum of the threads' counts -> *out)
__device__ __forceinline__ void inc_count_finish(uint32_t c, uint32_t* out, uint32_t* s_w) {
  c = __reduce_add_sync(0xffffffffu, c);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_w[warp] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_w[w];
    *out = tot;
  }
}
__device__ __forceinline__ void inc_count(const IncJob& J, int m, int t, int ntiles, uint32_t* out, uint32_t* s_w) {
  const WindowBuffers& W = J.w[m];
  const int n = J.n_old[m];
  const int len = (n + ntiles - 1) / ntiles;
  const int a = min(n, t * len), b = min(n, a + len);
  uint32_t c = 0;
  if (J.n_dead == 0) {
    c = threadIdx.x == 0 ? (uint32_t)(b - a) : 0u;
  } else {
    const uint32_t* __restrict__ tag = W.tag[J.cur];
    for (int i0 = a + (int)threadIdx.x; i0 < b; i0 += 10 * (int)blockDim.x) {  // ten loads in flight: one round for a 0.62 M-entry window on 148 blocks
      uint32_t tv[10];
#pragma unroll
      for (int u = 0; u < 10; ++u) tv[u] = i0 + u * (int)blockDim.x < b ? __ldcg(tag + i0 + u * (int)blockDim.x) : 0u;
#pragma unroll
      for (int u = 0; u < 10; ++u) c += (i0 + u * (int)blockDim.x < b && inc_alive(tv[u], J)) ? 1u : 0u;
    }
  }
  inc_count_finish(c, out, s_w);
}

// ---- M: this worker's tile of the old window merged with its slice of the run, into buffer cur ^ 1.
// smem (uint32 words): [keys 2 * INC_SUB | prefix INC_SUB + 1 (+ pad) | slice keys 2 * INC_SUB | 64]
constexpr int INC_MERGE_SMEM_WORDS = 2 * INC_SUB + (INC_SUB + 8) + 2 * INC_SUB + 64;
__device__ __forceinline__ void inc_merge(const IncJob& J, const MapBatch& B, int m, int t, int ntiles, const uint32_t* counts,
                                          uint32_t* smem) {
  const WindowBuffers& W = J.w[m];
  const MapBuffers& M = B.m[m];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem);
  uint32_t* s_pref = smem + 2 * INC_SUB;
  unsigned long long* s_slice = reinterpret_cast<unsigned long long*>(smem + 2 * INC_SUB + INC_SUB + 8);
  uint32_t* s_w = smem + 2 * INC_SUB + INC_SUB + 8 + 2 * INC_SUB;
  const int n = J.n_old[m], n_new = J.n_new[m];
  const int len = n > 0 ? (n + ntiles - 1) / ntiles : 0;
  const int a0 = min(n, t * len), b0 = min(n, a0 + len);
  const bool first_tile = t == 0;
  const bool last_tile = b0 >= n;            // (also true for the tiles behind the end of a short window)
  if (a0 >= n && !(first_tile)) return;      // nothing here; tile 0 of an EMPTY window still places the run
  const float4* __restrict__ o_pts = W.pts[J.cur];
  const unsigned long long* __restrict__ o_key = W.key[J.cur];
  const uint32_t* __restrict__ o_tag = W.tag[J.cur];
  float4* __restrict__ d_pts = W.pts[J.cur ^ 1];
  unsigned long long* __restrict__ d_key = W.key[J.cur ^ 1];
  uint32_t* __restrict__ d_tag = W.tag[J.cur ^ 1];
  const unsigned long long* r_key = W.run_key;  // (written earlier in this launch: no read-only path)
  // surviving entries in the tiles before this one
  uint32_t before = 0;
  for (int w = tid; w < t; w += blockDim.x) before += __ldcg(counts + w);
  before = __reduce_add_sync(0xffffffffu, before);
  __syncthreads();
  if (lane == 0) s_w[warp] = before;
  __syncthreads();
  before = 0;
  for (int w = 0; w < nwarps; ++w) before += s_w[w];
  __syncthreads();
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  bool any = false;
  auto track = [&](const float4 q) {
    if (finite3(q)) {
      any = true;
      mn[0] = fminf(mn[0], q.x); mx[0] = fmaxf(mx[0], q.x);
      mn[1] = fminf(mn[1], q.y); mx[1] = fmaxf(mx[1], q.y);
      mn[2] = fminf(mn[2], q.z); mx[2] = fmaxf(mx[2], q.z);
    }
  };
  uint32_t surv_before = before;  // surviving entries before the current sub-tile
  // the run entries of this tile: [r_lo, r_hi) = those whose key is >= this tile's first key and < the next tile's
  const int r_lo = first_tile ? 0 : inc_bound_block(r_key, 0, n_new, __ldcg(o_key + a0));
  const int r_hi = last_tile ? n_new : inc_bound_block(r_key, 0, n_new, __ldcg(o_key + b0));
  int sub_lo = r_lo;
  LL_TL_SUB(10);  // merge: tile offsets and run bounds known
  const int m_nsub = max(1, (b0 - a0 + INC_SUB - 1) / INC_SUB), m_step = max(1, (b0 - a0 + m_nsub - 1) / m_nsub);
  for (int a = a0; a < b0 || (a == a0 && a0 >= b0); a += m_step) {  // (one pass with an empty range when the window is empty)
    const int b = min(b0, a + m_step), cnt = max(0, b - a);
    // old entries of the sub-tile: R consecutive ones per thread; their keys, tags and points all go out before the search
    // for the run's share of the sub-tile (two more round trips), so the three latencies overlap
    constexpr int R = INC_SUB / 512;
    static_assert(R * 512 == INC_SUB && GI_THREADS == 512, "sub-tile = R entries per thread of a 512-thread block");
    const int l0 = R * tid;
    unsigned long long k4[R];
    uint32_t tg[R];
    float4 q4[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const bool in = l0 + k < cnt;
      k4[k] = in ? __ldcg(o_key + a + l0 + k) : INC_KEY_NONFINITE;
      tg[k] = in ? __ldcg(o_tag + a + l0 + k) : 0u;
      q4[k] = in ? __ldcg(o_pts + a + l0 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int sub_hi = b >= b0 ? r_hi : inc_bound_block(r_key, sub_lo, r_hi, __ldcg(o_key + b));
    const int ns = sub_hi - sub_lo;
    const bool slice_in_smem = ns <= INC_SUB;
    bool al[R];
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const bool in = l0 + k < cnt;
      al[k] = in && inc_alive(tg[k], J);
      s_keys[l0 + k] = k4[k];
      c += al[k] ? 1u : 0u;
    }
    if (slice_in_smem)
      for (int j = tid; j < ns; j += blockDim.x) s_slice[j] = __ldcg(r_key + sub_lo + j);
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, sub_total = 0;
    for (int w = 0; w < nwarps; ++w) {
      const uint32_t cw = s_w[w];
      woff += w < warp ? cw : 0u;
      sub_total += cw;
    }
    {
      uint32_t run = woff + incl - c;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        s_pref[l0 + k] = run;  // surviving entries of the sub-tile before entry l0 + k
        run += al[k] ? 1u : 0u;
      }
      if (tid == blockDim.x - 1) s_pref[INC_SUB] = run;
    }
    __syncthreads();
    // surviving old entries -> (surviving entries before) + (run entries with a smaller key)
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (!al[k]) continue;
      const int nb = slice_in_smem ? inc_bound<false, false>(s_slice, 0, ns, k4[k]) : inc_bound<false, true>(r_key + sub_lo, 0, ns, k4[k]);
      const size_t pos = (size_t)surv_before + s_pref[l0 + k] + (size_t)(sub_lo + nb);
      const float4 q = q4[k];
      d_pts[pos] = q;
      d_key[pos] = k4[k];
      d_tag[pos] = tg[k];
      track(q);
    }
    // run entries of the slice -> (their index in the run) + (surviving old entries with a key <= theirs)
    for (int j = tid; j < ns; j += blockDim.x) {
      const unsigned long long kx = slice_in_smem ? s_slice[j] : __ldcg(r_key + sub_lo + j);
      const int ub = inc_bound<true, false>(s_keys, 0, cnt, kx);
      const uint32_t surv = ub >= INC_SUB ? s_pref[INC_SUB] : (ub < cnt ? s_pref[ub] : s_pref[INC_SUB]);
      const size_t pos = (size_t)(sub_lo + j) + surv_before + surv;
      const float4 q = __ldcg(W.run_pts + sub_lo + j);
      d_pts[pos] = q;
      d_key[pos] = kx;
      d_tag[pos] = __ldcg(W.run_tag + sub_lo + j);
      track(q);
    }
    surv_before += sub_total;
    sub_lo = sub_hi;
    __syncthreads();
    LL_TL_SUB(11 + min(4, (a - a0) / m_step));  // merge: sub-tile written
    if (a0 >= b0) break;
  }
  // bounding box of everything written (finite points), as transform_minmax_kernel
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    const float v0 = warp_min(mn[ax]), v1 = warp_max(mx[ax]);
    if (lane == 0 && __any_sync(0x1u, true)) {
      if (v0 <= v1) {
        atomicMax(&M.st->mm[ax], ~float_to_ordered(v0));
        atomicMax(&M.st->mm[3 + ax], float_to_ordered(v1));
      }
    }
  }
  (void)any;
}

// ---- V: heads of the merged window per tile, then the centroids (the reduce of gi_voxel on keyed, already ordered points).
constexpr int INC_REDUCE_SMEM_WORDS = 4 * INC_SUB + INC_SUB + 8 + 64 + 16;
__device__ __forceinline__ uint32_t inc_count_heads(const IncJob& J, int m, int t, int ntiles, uint32_t* s_w) {
  const unsigned long long* key = J.w[m].key[J.cur ^ 1];
  const int n = J.n_out[m];
  const int len = (n + ntiles - 1) / ntiles;
  const int a = min(n, t * len), b = min(n, a + len);
  uint32_t c = 0;
  for (int i0 = a + (int)threadIdx.x; i0 < b; i0 += 10 * (int)blockDim.x) {  // twenty loads in flight: one round per tile
    unsigned long long kv[10], pv[10];
#pragma unroll
    for (int u = 0; u < 10; ++u) {
      const int i = i0 + u * (int)blockDim.x;
      kv[u] = i < b ? __ldcg(key + i) : INC_KEY_NONFINITE;
      pv[u] = (i < b && i > 0) ? __ldcg(key + i - 1) : INC_KEY_NONFINITE;
    }
#pragma unroll
    for (int u = 0; u < 10; ++u) {
      const int i = i0 + u * (int)blockDim.x;
      const bool head = i < b && kv[u] < INC_KEY_NONFINITE - 1ull && (i == 0 || pv[u] != kv[u]);
      c += head ? 1u : 0u;
    }
  }
  c = __reduce_add_sync(0xffffffffu, c);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_w[warp] = c;
  __syncthreads();
  uint32_t tot = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_w[w];
  __syncthreads();
  return tot;
}

__device__ __forceinline__ void inc_reduce(const IncJob& J, const MapBatch& B, int m, int t, int ntiles, const uint32_t* heads,
                                           uint32_t* smem) {
  const MapBuffers& M = B.m[m];
  const unsigned long long* keys = J.w[m].key[J.cur ^ 1];  // (written earlier in this launch: read with __ldcg)
  const float4* pts = J.w[m].pts[J.cur ^ 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int n = J.n_out[m];
  const int len = (n + ntiles - 1) / ntiles;
  const int tile_begin = min(n, t * len), tile_end = min(n, tile_begin + len);
  float4* s_pts = reinterpret_cast<float4*>(smem);                           // [INC_SUB]
  int* s_head_pos = reinterpret_cast<int*>(smem + 4 * INC_SUB);              // [INC_SUB + 1]
  uint32_t* s_wsum = smem + 4 * INC_SUB + INC_SUB + 8;                       // [nwarps]
  int* s_misc = reinterpret_cast<int*>(smem + 4 * INC_SUB + INC_SUB + 8 + 32);
  float4* s_cross_sum = reinterpret_cast<float4*>(smem + 4 * INC_SUB + INC_SUB + 8 + 64);
  unsigned long long* s_cross_key = reinterpret_cast<unsigned long long*>(smem + 4 * INC_SUB + INC_SUB + 8 + 64 + 4);
  // output slots before this tile, and the size of laserCloud*FromMapDS
  uint32_t before = 0, total = 0;
  for (int w = tid; w < ntiles; w += blockDim.x) {
    const uint32_t c = __ldcg(heads + w);
    total += c;
    before += w < t ? c : 0u;
  }
  before = __reduce_add_sync(0xffffffffu, before);
  total = __reduce_add_sync(0xffffffffu, total);
  __syncthreads();
  if (lane == 0) { s_wsum[warp] = before; s_wsum[16 + warp] = total; }
  __syncthreads();
  before = 0; total = 0;
  for (int w = 0; w < nwarps; ++w) { before += s_wsum[w]; total += s_wsum[16 + w]; }
  if (t == 0 && tid == 0) {
    M.st->n_ds = (int)total;
    M.st->n_raw = n;
  }
  constexpr int R = INC_SUB / 512;
  uint32_t slot_base = before;
  LL_TL_SUB(20);  // centroids: output offsets known
  const int r_nsub = max(1, (tile_end - tile_begin + INC_SUB - 1) / INC_SUB), r_step = max(1, (tile_end - tile_begin + r_nsub - 1) / r_nsub);
  for (int base = tile_begin; base < tile_end; base += r_step) {
    const int sub_n = min(r_step, tile_end - base);
    __syncthreads();
    if (tid == 0) { s_misc[0] = INC_SUB; s_misc[1] = 0; }  // [0] first non-finite entry, [1] a segment crosses the end
    __syncthreads();
    const int l0 = R * tid, i0 = base + l0;
    unsigned long long k2[R];
    bool head[R];
    unsigned long long prev = (i0 > 0 && l0 < sub_n) ? __ldcg(keys + i0 - 1) : INC_KEY_NONFINITE;
#pragma unroll
    for (int k = 0; k < R; ++k) k2[k] = l0 + k < sub_n ? __ldcg(keys + i0 + k) : INC_KEY_NONFINITE;
#pragma unroll
    for (int k = 0; k < R; ++k) s_pts[l0 + k] = l0 + k < sub_n ? __ldcg(pts + i0 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const bool in = l0 + k < sub_n;
      const bool boundary = in && (i0 + k == 0 || k2[k] != prev);
      const bool dropped = k2[k] >= INC_KEY_NONFINITE - 1ull;
      head[k] = boundary && !dropped;
      if (in && dropped && (i0 + k == 0 || prev < INC_KEY_NONFINITE - 1ull)) s_misc[0] = l0 + k;  // first trailing entry
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
    for (int w = 0; w < nwarps; ++w) {
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
      const int end = min(s_head_pos[ls + 1], sent_pos);
      const uint32_t slot = slot_base + ls;
      ++ls;
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      int jj = beg;
      for (; jj + 4 <= end; jj += 4) {
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
      if (end == sub_n && sent_pos >= sub_n && base + sub_n < n) {
        // the segment may continue behind this sub-tile (or tile): park the running sums, the block finishes it below
        *s_cross_sum = make_float4(sx, sy, sz, si);
        s_misc[2] = c;
        s_misc[3] = (int)slot;
        *s_cross_key = k2[k];
        s_misc[1] = 1;
        continue;
      }
      const float fc = (float)c;
      M.ds[slot] = make_float4(sx / fc, sy / fc, sz / fc, si / fc);
    }
    slot_base += sub_total;
    __syncthreads();
    LL_TL_SUB(21 + min(4, (base - tile_begin) / r_step));  // centroids: sub-tile done
    if (!s_misc[1]) continue;  // block-uniform
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    int c = 0;
    if (tid == 0) {
      sx = s_cross_sum->x; sy = s_cross_sum->y; sz = s_cross_sum->z; si = s_cross_sum->w;
      c = s_misc[2];
    }
    const unsigned long long ck = *s_cross_key;
    const uint32_t cslot = (uint32_t)s_misc[3];
    for (int g0 = base + sub_n; g0 < n; g0 += (int)blockDim.x) {
      const int g = g0 + tid;
      const bool match = g < n && __ldcg(keys + g) == ck;
      if (tid == 0) s_misc[5] = (int)blockDim.x;
      __syncthreads();
      if (!match) atomicMin(&s_misc[5], tid);
      __syncthreads();
      const int mcount = s_misc[5];
      if (tid < mcount) s_pts[tid] = __ldcg(pts + g);
      __syncthreads();
      if (tid == 0) {
        for (int u = 0; u < mcount; ++u) {
          const float4 p = s_pts[u];
          sx += p.x; sy += p.y; sz += p.z; si += p.w;
        }
        c += mcount;
      }
      __syncthreads();
      if (mcount < (int)blockDim.x) break;
    }
    if (tid == 0) {
      const float fc = (float)c;
      M.ds[cslot] = make_float4(sx / fc, sy / fc, sz / fc, si / fc);
    }
  }
}

// The whole incremental step for the two maps of B. Every thread of every block of the co-resident grid calls; the caller
// has zeroed both MapStates and passed a barrier. Returns with a grid barrier passed. smem: INC_SMEM_BYTES.
__device__ __forceinline__ void map_incremental_inline(const IncJob& J, const MapBatch& B, unsigned* barrier, unsigned& epoch,
                                                       uint32_t* smem) {
  const int nb = (int)gridDim.x, blk = (int)blockIdx.x;
  const IncTiling T(nb, J.n_old, false);
  uint32_t* s_w = smem + INC_MERGE_SMEM_WORDS - 64;
  // ---- R | C
  LL_TL_SUB(1);
  // C's loads go out first: the tags of this block's tile of the old window (cold: DRAM) are in registers by the time the
  // run is built (one round of ten loads per thread covers a 0.62 M-entry window on 148 blocks; a longer tile counts below)
  constexpr int kPre = 10;
  uint32_t tv[kPre];
  bool pre = false;
  int c_a = 0, c_b = 0;
  if (blk >= T.first && nb - T.first >= 2 && J.n_dead > 0) {
    const int w = blk - T.first, m = w < T.nw[0] ? 0 : 1, t = m ? w - T.nw[0] : w;
    const int n = J.n_old[m], len = (n + T.nw[m] - 1) / T.nw[m];
    c_a = min(n, t * len);
    c_b = min(n, c_a + len);
    pre = c_b - c_a <= kPre * (int)blockDim.x;  // block-uniform
    if (pre) {
      const uint32_t* __restrict__ tag = J.w[m].tag[J.cur];
#pragma unroll
      for (int u = 0; u < kPre; ++u) {
        const int i = c_a + (int)threadIdx.x + u * (int)blockDim.x;
        tv[u] = i < c_b ? __ldcg(tag + i) : 0u;
      }
    }
  }
  for (int m = 0; m < 2; ++m) {
    if (J.n_new[m] > 0) inc_build_run(J, B, m, reinterpret_cast<unsigned long long*>(smem));
    LL_TL_SUB(3 + 2 * m);  // ... ranked and placed
  }
  __syncthreads();
  LL_TL_SUB(6);
  if (blk >= T.first) {
    const int w = blk - T.first;
    if (nb - T.first < 2) {
      for (int m = 0; m < 2; ++m) inc_count(J, m, 0, 1, J.w[m].counts, s_w);
    } else if (pre) {
      const int m = w < T.nw[0] ? 0 : 1, t = m ? w - T.nw[0] : w;
      uint32_t c = 0;
#pragma unroll
      for (int u = 0; u < kPre; ++u) c += (c_a + (int)threadIdx.x + u * (int)blockDim.x < c_b && inc_alive(tv[u], J)) ? 1u : 0u;
      inc_count_finish(c, J.w[m].counts + t, s_w);
    } else {
      const int m = w < T.nw[0] ? 0 : 1, t = m ? w - T.nw[0] : w;
      inc_count(J, m, t, T.nw[m], J.w[m].counts + t, s_w);
    }
  }
  LL_TL_SUB(7);  // surviving entries counted
  coop_barrier(barrier, epoch);
  // ---- M
  if (blk >= T.first) {
    const int w = blk - T.first;
    if (nb - T.first < 2) {
      for (int m = 0; m < 2; ++m) {
        inc_merge(J, B, m, 0, 1, J.w[m].counts, smem);
        __syncthreads();
      }
    } else {
      const int m = w < T.nw[0] ? 0 : 1, t = m ? w - T.nw[0] : w;
      inc_merge(J, B, m, t, T.nw[m], J.w[m].counts, smem);
    }
  }
  coop_barrier(barrier, epoch);
  if (!J.final_step) return;
  // ---- V: every block takes part; the tiling is over the merged windows
  const IncTiling TV(nb, J.n_out, false);
  bool fail = false;
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const VoxelMeta v = voxel_meta<true>(B.m[m].st, B.m[m].inv_leaf);
    fail = fail || v.overflow || __ldcg(&B.m[m].st->inc_fail) != 0;
  }
  if (fail) {  // grid-uniform: the centroids are not computed; lm_begin_state reports it and the host rebuilds
    if (blk == 0 && threadIdx.x == 0) { B.m[0].st->inc_fail = 1; B.m[1].st->inc_fail = 1; }
    coop_barrier(barrier, epoch);
    return;
  }
  uint32_t* s_wr = smem + INC_REDUCE_SMEM_WORDS - 80;
  if (nb < 2) {
    for (int m = 0; m < 2; ++m) {
      const uint32_t c = inc_count_heads(J, m, 0, 1, s_wr);
      if (threadIdx.x == 0) J.w[m].counts[INC_MAX_BLOCKS] = c;
    }
  } else {
    const int m = blk < TV.nw[0] ? 0 : 1, t = m ? blk - TV.nw[0] : blk;
    const uint32_t c = inc_count_heads(J, m, t, TV.nw[m], s_wr);
    if (threadIdx.x == 0) J.w[m].counts[INC_MAX_BLOCKS + t] = c;
  }
  coop_barrier(barrier, epoch);
  if (nb < 2) {
    for (int m = 0; m < 2; ++m) {
      inc_reduce(J, B, m, 0, 1, J.w[m].counts + INC_MAX_BLOCKS, smem);
      __syncthreads();
    }
  } else {
    const int m = blk < TV.nw[0] ? 0 : 1, t = m ? blk - TV.nw[0] : blk;
    inc_reduce(J, B, m, t, TV.nw[m], J.w[m].counts + INC_MAX_BLOCKS, smem);
  }
  if (blk == 0 && threadIdx.x == 0) {
    for (int m = 0; m < 2; ++m) {
      B.m[m].st->voxel_overflow = 0;
      B.m[m].st->voxel_bits = 0;
    }
  }
  coop_barrier(barrier, epoch);
}

// One merge as its own cooperative launch: the groups of a long list of added keyframes before the last, and a pending
// step that another consumer than ll_scan2map needs finished (ensure_map in abi.cu).
__global__ void __launch_bounds__(GI_THREADS, 1) map_inc_kernel(const __grid_constant__ IncJob J, const __grid_constant__ MapBatch B,
                                                                unsigned* barrier, unsigned* barrier_clear) {
  extern __shared__ __align__(16) unsigned char s_inc[];
  unsigned epoch = 0;
  if (blockIdx.x == 0) inc_zero_map_states(B, J.keep_fail != 0);
  coop_barrier(barrier, epoch);
  map_incremental_inline(J, B, barrier, epoch, reinterpret_cast<uint32_t*>(s_inc));
  if (blockIdx.x == 0 && threadIdx.x == 0) *barrier_clear = 0u;
}

}  // namespace ll
