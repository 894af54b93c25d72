# E5: the added keyframes' run is ranked through 256 value-range buckets (sample splitters) instead of n^2 / blocks comparisons.
p='csrc/map_incremental.cuh'
s=open(p).read()
a="""// ---- R: the run of the added keyframes of map m = their points sorted by (voxel key, arrival order), by ALL blocks."""
assert s.count(a)==1
new = r'''// Ranks through value-range buckets (inc_build_run, n >= INC_BKT_MIN points that fit the layout below): 256 sample keys are
// rank-sorted into splitters, every key finds its bucket by binary search (bucket = number of splitters <= key: monotone in
// the key), a counting pass lists each bucket's members, and the rank of a point is (keys in lower buckets) + (members of
// its own bucket below it) — ~40 comparisons per point instead of n. Every block does all of this for all keys (they are
// in its shared memory anyway) and then ranks and places only its own slice [lo, hi).
// smem behind the n composites: splitters u64[256] | hist int[264] | cursor int[256] | members u16[n] | bucket u8[n]
constexpr int INC_BKT = 256;
constexpr int INC_BKT_MIN = 2048;
__host__ __device__ constexpr int inc_bucket_smem_bytes(int n) { return n * 8 + INC_BKT * 8 + 264 * 4 + INC_BKT * 4 + ((n + 7) & ~7) * 2 + ((n + 15) & ~15); }

__device__ __forceinline__ void inc_rank_buckets(const IncJob& J, const MapBatch& B, int m, unsigned long long* s_key, int n, int lo, int hi) {
  const WindowBuffers& W = J.w[m];
  const float inv = B.m[m].inv_leaf;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5, nt = blockDim.x;
  unsigned long long* s_split = s_key + n;                       // [INC_BKT], ascending; [0] is not used as a splitter
  int* s_hist = reinterpret_cast<int*>(s_split + INC_BKT);       // [INC_BKT + 1] (+ pad): members per bucket, then exclusive prefix
  int* s_cur = s_hist + 264;                                     // [INC_BKT]
  unsigned short* s_mem = reinterpret_cast<unsigned short*>(s_cur + INC_BKT);  // [n] arrival indices, bucket by bucket
  unsigned char* s_bkt = reinterpret_cast<unsigned char*>(s_mem + ((n + 7) & ~7));  // [n]
  // ---- splitters: sample t = key at arrival index t * n / 256, rank-sorted (the composites are unique); the 512 threads
  //      split each sample's 256 comparisons in two halves
  if (tid < 264) s_hist[tid] = 0;
  if (tid < INC_BKT) s_cur[tid] = 0;
  __syncthreads();
  {
    const int t = tid & (INC_BKT - 1), half = tid >> 8;  // (blockDim.x == 512)
    const unsigned long long mine = s_key[(int)(((long long)t * n) >> 8)];
    int c = 0;
    for (int u = half * (INC_BKT / 2); u < (half + 1) * (INC_BKT / 2); ++u) c += s_key[(int)(((long long)u * n) >> 8)] < mine ? 1 : 0;
    atomicAdd(&s_cur[t], c);
  }
  __syncthreads();
  if (tid < INC_BKT) s_split[s_cur[tid]] = s_key[(int)(((long long)tid * n) >> 8)];
  __syncthreads();
  if (tid < INC_BKT) s_cur[tid] = 0;
  // ---- bucket of every key (number of splitters [1..255] that are <= key), members per bucket
  for (int j = tid; j < n; j += nt) {
    const unsigned long long k = s_key[j];
    int b = 0;
#pragma unroll
    for (int step = INC_BKT / 2; step > 0; step >>= 1) b += s_split[b + step] <= k ? step : 0;
    s_bkt[j] = (unsigned char)b;
    atomicAdd(&s_hist[b + 1], 1);
  }
  __syncthreads();
  // ---- exclusive prefix over the buckets: s_hist[b] = members of the buckets below b (warp 0: eight bins per lane)
  if (warp == 0) {
    int v[8], sum = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) { v[u] = s_hist[1 + lane * 8 + u]; sum += v[u]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    int run = incl - sum;
#pragma unroll
    for (int u = 0; u < 8; ++u) { run += v[u]; s_hist[1 + lane * 8 + u] = run; }   // s_hist[b + 1] = members of buckets <= b
    if (lane == 0) s_hist[0] = 0;
  }
  __syncthreads();
  // ---- the members of every bucket (any order within a bucket)
  for (int j = tid; j < n; j += nt) {
    const int b = s_bkt[j];
    s_mem[s_hist[b] + atomicAdd(&s_cur[b], 1)] = (unsigned short)j;
  }
  __syncthreads();
  // ---- this block's slice, dealt to the warps: the whole warp counts for one point at a time, then its lanes place the points
  const int per_w = (hi - lo + nwarps - 1) / nwarps;  // points per warp (five at 10 k points on 148 blocks)
  const int w_lo = min(hi, lo + warp * per_w), w_hi = min(hi, w_lo + per_w);
  for (int e0 = w_lo; e0 < w_hi; e0 += 32) {
    const int ne = min(32, w_hi - e0);
    int my_rank = 0;
    for (int u = 0; u < ne; ++u) {
      const int e = e0 + u;
      const unsigned long long ke = s_key[e];
      const int b = s_bkt[e], beg = s_hist[b], end = s_hist[b + 1];
      int c = 0;
      for (int t = beg + lane; t < end; t += 32) c += s_key[s_mem[t]] < ke ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (lane == u) my_rank = beg + c;
    }
    if (lane < ne) {
      const int e = e0 + lane;
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
      W.run_pts[my_rank] = q;
      W.run_key[my_rank] = inc_abs_key(q, inv, ok);
      W.run_tag[my_rank] = tag;
    }
  }
}

'''
s=s.replace(a, new+a)
a="""  const int want = (hi - lo + nwarps - 1) / nwarps;
  if (want <= 4) inc_rank_slice<4>(J, B, m, s_key, n, lo, hi);"""
assert s.count(a)==1
s=s.replace(a,"""  if (n >= INC_BKT_MIN && n <= 65535 && inc_bucket_smem_bytes(n) <= INC_SMEM_BYTES) {  // (grid-uniform)
    inc_rank_buckets(J, B, m, s_key, n, lo, hi);
    return;
  }
  const int want = (hi - lo + nwarps - 1) / nwarps;
  if (want <= 4) inc_rank_slice<4>(J, B, m, s_key, n, lo, hi);""")
open(p,'w').write(s)
