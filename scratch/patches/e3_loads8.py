# E3: eight loads in flight in the pass that transforms the added keyframes' points (every block reads all of them).
p='csrc/map_incremental.cuh'
s=open(p).read()
a="""    for (int i0 = tid; i0 < g.count; i0 += 4 * blockDim.x) {  // four loads in flight (the clouds are cold: fresh from the host)
      float4 pv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * (int)blockDim.x;
        pv[u] = i < g.count ? __ldg(g.src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {"""
assert s.count(a)==1
s=s.replace(a,"""    for (int i0 = tid; i0 < g.count; i0 += 8 * blockDim.x) {  // eight loads in flight (the clouds are cold: fresh from the host)
      float4 pv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * (int)blockDim.x;
        pv[u] = i < g.count ? __ldg(g.src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {""")
open(p,'w').write(s)
