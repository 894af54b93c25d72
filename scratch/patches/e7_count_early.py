# E7: the survivor count's tag loads are issued BEFORE the run is built and consumed after it (their DRAM latency hides
# behind ~25 us of ranking).
p='csrc/map_incremental.cuh'
s=open(p).read()
a="""  c = __reduce_add_sync(0xffffffffu, c);
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
"""
assert s.count(a)==1
s=s.replace(a,"""  inc_count_finish(c, out, s_w);
}
""")
a="// ---- C: surviving entries of this worker's tile of the old window.\n"
assert s.count(a)==1
s=s.replace(a, a+"""// (block sum of the threads' counts -> *out)
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
""")
a="""  LL_TL_SUB(1);
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
    } else {
      const int m = w < T.nw[0] ? 0 : 1, t = m ? w - T.nw[0] : w;
      inc_count(J, m, t, T.nw[m], J.w[m].counts + t, s_w);
    }
  }
"""
assert s.count(a)==1
s=s.replace(a,"""  LL_TL_SUB(1);
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
""")
open(p,'w').write(s)
