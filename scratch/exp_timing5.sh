#!/bin/bash
# Timing-instrumented copy: %globaltimer stamps (block 0) at every grid barrier of the single-kernel scan.
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
s=s.replace(a, "__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile(\"mov.u64 %0, %globaltimer;\" : \"=l\"(t)); return t; }\n__device__ unsigned long long g_bar[2][128];\n"+a+"  if (threadIdx.x == 0 && blockIdx.x == 0 && epoch < 128) g_bar[0][epoch] = gtime();\n")
a="    while (ld_acquire_u32(counter) < target) {}\n  }\n  __syncthreads();\n"
assert s.count(a)==1
s=s.replace(a, a+"  if (threadIdx.x == 0 && blockIdx.x == 0 && epoch < 128) g_bar[1][epoch] = gtime();\n")
open(p,'w').write(s)
p='scratch/var/timing/csrc/lm.cuh'
s=open(p).read()
a="  unsigned epoch = 0;  // grid barriers passed so far\n"
assert s.count(a)==1
s=s.replace(a, "  unsigned long long t_entry = gtime();\n"+a)
a="    if (tid == 0) *barrier_clear = 0u;\n"
assert s.count(a)==1
s=s.replace(a, a+"    if (tid == 0) { unsigned long long te = gtime(); printf(\"single-kernel scan blk0: entry->end %llu ns; barriers (arrive/leave):\", te - t_entry); for (unsigned i = 1; i <= epoch && i < 128; ++i) printf(\" %llu/%llu\", g_bar[0][i] - t_entry, g_bar[1][i] - t_entry); printf(\"\\n\"); }\n")
open(p,'w').write(s)
PY
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -cudart static -o scratch/var/lib_timing.so scratch/var/timing/csrc/abi.cu
ls -la scratch/var/lib_timing.so
