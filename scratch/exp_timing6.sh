#!/bin/bash
# Timing-instrumented copy: per-block arrive times at every grid barrier (distribution: who are the stragglers?).
set -e
cd /root/repo
rm -rf scratch/var/timing && mkdir -p scratch/var/timing/csrc scratch/var/include
cp lego-loam-sr_b200/csrc/* scratch/var/timing/csrc/
cp include/legoloam_b200.h scratch/var/include/
python - <<'PY'
p='scratch/var/timing/csrc/common.cuh'
s=open(p).read()
a="__device__ __forceinline__ void coop_barrier(unsigned* counter, unsigned& epoch) {\n  ++epoch;\n"
assert s.count(a)==1
s=s.replace(a, "__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile(\"mov.u64 %0, %globaltimer;\" : \"=l\"(t)); return t; }\n__device__ unsigned long long g_arr[48][160];\n__device__ unsigned long long g_rel[48];\n"+a+"  if (threadIdx.x == 0 && epoch < 48) g_arr[epoch][blockIdx.x] = gtime();\n")
a="  }\n  __syncthreads();\n}\n\n// Streaming"
assert s.count(a)==1
s=s.replace(a, "  }\n  __syncthreads();\n  if (threadIdx.x == 0 && blockIdx.x == 0 && epoch < 48) g_rel[epoch] = gtime();\n}\n\n// Streaming")
open(p,'w').write(s)
p='scratch/var/timing/csrc/lm.cuh'
s=open(p).read()
a="  unsigned epoch = 0;  // grid barriers passed so far\n"
assert s.count(a)==1
s=s.replace(a, "  unsigned long long t_entry = gtime();\n"+a)
a="    if (tid == 0) *barrier_clear = 0u;\n"
assert s.count(a)==1
s=s.replace(a, a+"""    if (tid == 0) {
      printf("barrier arrivals relative to kernel entry of block 0 (us): epoch: first / median-ish / last (last block id) -> release\\n");
      for (unsigned e = 1; e <= epoch && e < 48; ++e) {
        unsigned long long mn = ~0ull, mx = 0, sum = 0; int who = -1;
        for (int b = 0; b < (int)gridDim.x; ++b) { const unsigned long long t = g_arr[e][b]; if (t < mn) mn = t; if (t > mx) { mx = t; who = b; } sum += t - t_entry; }
        printf("  %2u: %7.2f / %7.2f / %7.2f (blk %3d) -> %7.2f\\n", e, (mn - t_entry) * 1e-3, sum * 1e-3 / gridDim.x, (mx - t_entry) * 1e-3, who, (g_rel[e] - t_entry) * 1e-3);
      }
    }
""")
open(p,'w').write(s)
PY
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -cudart static -o scratch/var/lib_timing.so scratch/var/timing/csrc/abi.cu
ls -la scratch/var/lib_timing.so
