#!/bin/bash
# Timing-instrumented copy: %globaltimer stamps inside grid_build_inline (block 0 and block 147).
set -e
cd /root/repo
rm -rf scratch/var/timing && mkdir -p scratch/var/timing/csrc scratch/var/include
cp lego-loam-sr_b200/csrc/* scratch/var/timing/csrc/
cp include/legoloam_b200.h scratch/var/include/
python - <<'PY'
p='scratch/var/timing/csrc/grid_inline.cuh'
s=open(p).read()
def rep(a,b,cnt=1):
    global s
    assert s.count(a)==cnt, (s.count(a), a)
    s=s.replace(a,b)
rep("namespace ll {\n\nconstexpr int GI_THREADS = 512;", "namespace ll {\n__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile(\"mov.u64 %0, %globaltimer;\" : \"=l\"(t)); return t; }\n__device__ unsigned long long g_gi[2][64];\n__device__ int g_gin[2];\n#define GI_STAMP() do { if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) { int b_ = blockIdx.x == 0 ? 0 : 1; g_gi[b_][g_gin[b_]++ & 63] = gtime(); } } while (0)\n\nconstexpr int GI_THREADS = 512;")
rep("    coop_barrier(barrier, epoch);\n    if (active) {\n      constexpr bool kPrefetch", "    GI_STAMP();\n    coop_barrier(barrier, epoch);\n    GI_STAMP();\n    if (active) {\n      constexpr bool kPrefetch")
rep("    coop_barrier(barrier, epoch);\n    if (active && !last) {", "    GI_STAMP();\n    coop_barrier(barrier, epoch);\n    GI_STAMP();\n    if (active && !last) {")
rep("  const int gtid = blockIdx.x * GI_THREADS + threadIdx.x, gthreads = gridDim.x * GI_THREADS;\n  GridDims gd[2];", "  if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) g_gin[blockIdx.x == 0 ? 0 : 1] = 0;\n  GI_STAMP();\n  const int gtid = blockIdx.x * GI_THREADS + threadIdx.x, gthreads = gridDim.x * GI_THREADS;\n  GridDims gd[2];")
rep("  if (items == 4) gi_sort_tiles<4>", "  GI_STAMP();\n  if (items == 4) gi_sort_tiles<4>")
rep("  else if (items == 16) gi_sort_tiles<16>(B, gd, n, barrier, epoch, smem);\n  coop_barrier(barrier, epoch);\n", "  else if (items == 16) gi_sort_tiles<16>(B, gd, n, barrier, epoch, smem);\n  GI_STAMP();\n  coop_barrier(barrier, epoch);\n  GI_STAMP();\n")
open(p,'w').write(s)

p2='scratch/var/timing/csrc/lm.cuh'
t=open(p2).read()
a="  unsigned epoch = 0;  // grid barriers passed so far\n"
assert t.count(a)==1
t=t.replace(a, "  unsigned long long t_entry = gtime();\n"+a)
a="    coop_barrier(barrier, epoch);\n    reduce_partials_and_solve"
assert t.count(a)==1
t=t.replace(a, "    GI_STAMP();\n    coop_barrier(barrier, epoch);\n    GI_STAMP();\n    reduce_partials_and_solve")
a="    if (tid == 0) *barrier_clear = 0u;\n"
assert t.count(a)==1
t=t.replace(a, a+"    if (tid == 0) { unsigned long long te = gtime(); printf(\"persistent kernel blk0: entry->end %llu ns; stamps:\", te - t_entry); for (int i = 0; i < g_gin[0] && i < 64; ++i) printf(\" %llu\", g_gi[0][i] - t_entry); printf(\"\\n\"); }\n")
open(p2,'w').write(t)
PY
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -cudart static -o scratch/var/lib_timing.so scratch/var/timing/csrc/abi.cu
ls -la scratch/var/lib_timing.so
