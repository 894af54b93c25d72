# E4: the merge's bounding box is reduced over the BLOCK before the atomics (six per block instead of six per warp:
# 14 k atomics on twelve words of one sector queued up in front of every block's barrier arrival).
p='csrc/map_incremental.cuh'
s=open(p).read()
a="""#pragma unroll
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
"""
assert s.count(a)==1
s=s.replace(a,"""  float* s_box = reinterpret_cast<float*>(s_pref);  // [6][nwarps] (the prefix array is free: the loop above ended with a barrier)
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    const float v0 = warp_min(mn[ax]), v1 = warp_max(mx[ax]);
    if (lane == 0) {
      s_box[ax * 32 + warp] = v0;
      s_box[(3 + ax) * 32 + warp] = v1;
    }
  }
  __syncthreads();
  if (tid < 3) {
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (int w = 0; w < nwarps; ++w) {
      lo = fminf(lo, s_box[tid * 32 + w]);
      hi = fmaxf(hi, s_box[(3 + tid) * 32 + w]);
    }
    if (lo <= hi) {
      atomicMax(&M.st->mm[tid], ~float_to_ordered(lo));
      atomicMax(&M.st->mm[3 + tid], float_to_ordered(hi));
    }
  }
  (void)any;
""")
open(p,'w').write(s)
