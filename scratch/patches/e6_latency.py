# E6: latency chains of the incremental assembly: ten loads in flight in the survivor / head counts (one round per tile
# instead of three), the merge's points and tags fetched together with its keys (before the run-bound search).
p='csrc/map_incremental.cuh'
s=open(p).read()
a="""    for (int i0 = a + (int)threadIdx.x; i0 < b; i0 += 4 * (int)blockDim.x) {  // four loads in flight
      uint32_t tv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) tv[u] = i0 + u * (int)blockDim.x < b ? __ldcg(tag + i0 + u * (int)blockDim.x) : 0u;
#pragma unroll
      for (int u = 0; u < 4; ++u) c += (i0 + u * (int)blockDim.x < b && inc_alive(tv[u], J)) ? 1u : 0u;
    }"""
assert s.count(a)==1
s=s.replace(a,"""    for (int i0 = a + (int)threadIdx.x; i0 < b; i0 += 10 * (int)blockDim.x) {  // ten loads in flight: one round for a 0.62 M-entry window on 148 blocks
      uint32_t tv[10];
#pragma unroll
      for (int u = 0; u < 10; ++u) tv[u] = i0 + u * (int)blockDim.x < b ? __ldcg(tag + i0 + u * (int)blockDim.x) : 0u;
#pragma unroll
      for (int u = 0; u < 10; ++u) c += (i0 + u * (int)blockDim.x < b && inc_alive(tv[u], J)) ? 1u : 0u;
    }""")
a="""  for (int i0 = a + (int)threadIdx.x; i0 < b; i0 += 4 * (int)blockDim.x) {  // eight loads in flight
    unsigned long long kv[4], pv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * (int)blockDim.x;
      kv[u] = i < b ? __ldcg(key + i) : INC_KEY_NONFINITE;
      pv[u] = (i < b && i > 0) ? __ldcg(key + i - 1) : INC_KEY_NONFINITE;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {"""
assert s.count(a)==1
s=s.replace(a,"""  for (int i0 = a + (int)threadIdx.x; i0 < b; i0 += 10 * (int)blockDim.x) {  // twenty loads in flight: one round per tile
    unsigned long long kv[10], pv[10];
#pragma unroll
    for (int u = 0; u < 10; ++u) {
      const int i = i0 + u * (int)blockDim.x;
      kv[u] = i < b ? __ldcg(key + i) : INC_KEY_NONFINITE;
      pv[u] = (i < b && i > 0) ? __ldcg(key + i - 1) : INC_KEY_NONFINITE;
    }
#pragma unroll
    for (int u = 0; u < 10; ++u) {""")
# merge: loads first
a="""    const int sub_hi = b >= b0 ? r_hi : inc_bound_block(r_key, sub_lo, r_hi, __ldcg(o_key + b));
    const int ns = sub_hi - sub_lo;
    const bool slice_in_smem = ns <= INC_SUB;
    // old entries of the sub-tile: 4 consecutive ones per thread
    constexpr int R = INC_SUB / 512;
    static_assert(R * 512 == INC_SUB && GI_THREADS == 512, "sub-tile = four entries per thread of a 512-thread block");
    const int l0 = R * tid;
    unsigned long long k4[R];
    bool al[R];
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const bool in = l0 + k < cnt;
      k4[k] = in ? __ldcg(o_key + a + l0 + k) : INC_KEY_NONFINITE;
      al[k] = in && inc_alive(__ldcg(o_tag + a + l0 + k), J);
      s_keys[l0 + k] = k4[k];
      c += al[k] ? 1u : 0u;
    }
"""
assert s.count(a)==1
s=s.replace(a,"""    // old entries of the sub-tile: R consecutive ones per thread; their keys, tags and points all go out before the search
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
""")
a="""      const float4 q = __ldcg(o_pts + a + l0 + k);
      d_pts[pos] = q;
      d_key[pos] = k4[k];
      d_tag[pos] = __ldcg(o_tag + a + l0 + k);
      track(q);"""
assert s.count(a)==1
s=s.replace(a,"""      const float4 q = q4[k];
      d_pts[pos] = q;
      d_key[pos] = k4[k];
      d_tag[pos] = tg[k];
      track(q);""")
open(p,'w').write(s)
