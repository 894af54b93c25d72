# E8: the runs of the two maps are built TOGETHER when their points fit the shared-memory array at once: one pass over
# all added clouds for the composites, then the corner slice is ranked by the warps the surf slice leaves idle.
p='csrc/map_incremental.cuh'
s=open(p).read()
a="""template <int E>
__device__ __forceinline__ void inc_rank_slice(const IncJob& J, const MapBatch& B, int m, const unsigned long long* s_key, int n,
                                               int lo, int hi) {
  const WindowBuffers& W = J.w[m];
  const float inv = B.m[m].inv_leaf;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int e0 = lo + E * warp; e0 < hi; e0 += E * nwarps) {"""
assert s.count(a)==1
s=s.replace(a,"""// Warps [wf, wf + wc) of the block take part (the others return at once).
template <int E>
__device__ __forceinline__ void inc_rank_slice(const IncJob& J, const MapBatch& B, int m, const unsigned long long* s_key, int n,
                                               int lo, int hi, int wf = 0, int wc = GI_THREADS / 32) {
  const WindowBuffers& W = J.w[m];
  const float inv = B.m[m].inv_leaf;
  const int lane = threadIdx.x & 31, warp = (int)(threadIdx.x >> 5) - wf;
  if (warp < 0 || warp >= wc) return;
  for (int e0 = lo + E * warp; e0 < hi; e0 += E * wc) {""")
a="""// ---- C: surviving entries of this worker's tile of the old window."""
assert s.count(a)==1
s=s.replace(a,"""// ---- R for both maps at once (their added points together fit the composite array): one pass over all added clouds
// (all their loads in flight together, one copy of the code), then the surf slice is ranked by the first warps of the block
// and the corner slice — a fifth of the points, a twentieth of the comparisons — by the warps that leaves idle.
__device__ __forceinline__ void inc_build_runs_together(const IncJob& J, const MapBatch& B, unsigned long long* s_key) {
  const int tid = threadIdx.x, nwarps = blockDim.x >> 5;
  const int nn[2] = {J.n_new[0], J.n_new[1]};
  unsigned long long* s_k[2] = {s_key, s_key + nn[0]};
  __shared__ int s_bad2[2];
  __syncthreads();
  if (tid < 2) s_bad2[tid] = 0;
  __syncthreads();
  for (int sg = 0; sg < J.n_segs; ++sg) {
    const Segment& g = J.segs[sg];
    const int m = g.map;
    const float inv = B.m[m].inv_leaf;
    const int ox = J.run_org[m][0], oy = J.run_org[m][1], oz = J.run_org[m][2];
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
          if (!in) s_bad2[m] = 1;
          rel = ((unsigned long long)(unsigned)iz << 24) | ((unsigned long long)(unsigned)iy << 12) | (unsigned long long)(unsigned)ix;
        }
        const int e = g.dst_off + i;
        s_k[m][e] = (rel << INC_RUN_BITS) | (unsigned long long)e;  // unique within its map => the rank order is the stable order
      }
    }
  }
  __syncthreads();
  LL_TL_SUB(4);  // composites of all added points in shared memory
  int lo[2], hi[2], per[2];
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    if (s_bad2[m] && tid == 0 && blockIdx.x == 0) B.m[m].st->inc_fail = 1;  // (the same verdict in every block)
    per[m] = (nn[m] + (int)gridDim.x - 1) / (int)gridDim.x;
    lo[m] = min(nn[m], (int)blockIdx.x * per[m]);
    hi[m] = s_bad2[m] ? lo[m] : min(nn[m], lo[m] + per[m]);
  }
  // E points per warp and pass over the keys: the smallest of 4 / 6 / 8 with which both slices fit the block's warps side
  // by side (surf first); otherwise one map after the other on all warps
  auto warps_for = [&](int E, int m) { return (per[m] + E - 1) / E; };
  if (warps_for(4, 0) + warps_for(4, 1) <= nwarps) {
    inc_rank_slice<4>(J, B, 1, s_k[1], nn[1], lo[1], hi[1], 0, warps_for(4, 1));
    inc_rank_slice<4>(J, B, 0, s_k[0], nn[0], lo[0], hi[0], warps_for(4, 1), nwarps - warps_for(4, 1));
  } else if (warps_for(6, 0) + warps_for(6, 1) <= nwarps) {
    inc_rank_slice<6>(J, B, 1, s_k[1], nn[1], lo[1], hi[1], 0, warps_for(6, 1));
    inc_rank_slice<6>(J, B, 0, s_k[0], nn[0], lo[0], hi[0], warps_for(6, 1), nwarps - warps_for(6, 1));
  } else if (warps_for(8, 0) + warps_for(8, 1) <= nwarps) {
    inc_rank_slice<8>(J, B, 1, s_k[1], nn[1], lo[1], hi[1], 0, warps_for(8, 1));
    inc_rank_slice<8>(J, B, 0, s_k[0], nn[0], lo[0], hi[0], warps_for(8, 1), nwarps - warps_for(8, 1));
  } else {
    inc_rank_slice<8>(J, B, 1, s_k[1], nn[1], lo[1], hi[1]);
    inc_rank_slice<8>(J, B, 0, s_k[0], nn[0], lo[0], hi[0]);
  }
}

"""+a)
a="""  for (int m = 0; m < 2; ++m) {
    if (J.n_new[m] > 0) inc_build_run(J, B, m, reinterpret_cast<unsigned long long*>(smem));
    LL_TL_SUB(3 + 2 * m);  // ... ranked and placed
  }
"""
assert s.count(a)==1
s=s.replace(a,"""  if (J.n_new[0] > 0 && J.n_new[1] > 0 && J.n_new[0] + J.n_new[1] <= INC_RUN_MAX) {
    inc_build_runs_together(J, B, reinterpret_cast<unsigned long long*>(smem));
    LL_TL_SUB(5);  // ... ranked and placed
  } else {
    for (int m = 0; m < 2; ++m) {
      if (J.n_new[m] > 0) inc_build_run(J, B, m, reinterpret_cast<unsigned long long*>(smem));
      LL_TL_SUB(3 + 2 * m);  // ... ranked and placed
    }
  }
""")
open(p,'w').write(s)
