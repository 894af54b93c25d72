#!/bin/bash
# Timing-instrumented copy of the library: %globaltimer stamps inside lm_persistent_kernel (block 100 and max over blocks).
set -e
cd /root/repo
rm -rf scratch/var/timing && mkdir -p scratch/var/timing/csrc scratch/var/include
cp lego-loam-sr_b200/csrc/* scratch/var/timing/csrc/
cp include/legoloam_b200.h scratch/var/include/
python - <<'PY'
p='scratch/var/timing/csrc/lm.cuh'
s=open(p).read()
def rep(a,b):
    global s
    assert s.count(a)==1, (s.count(a), a)
    s=s.replace(a,b)
rep("namespace ll {\n\nconstexpr int LM_THREADS = 512;", "namespace ll {\n__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile(\"mov.u64 %0, %globaltimer;\" : \"=l\"(t)); return t; }\n__device__ unsigned long long g_ts[160][16][8];\n\nconstexpr int LM_THREADS = 512;")
# inside lm_block_pass: stamp after phase 1 and after fit (needs iteration index: use a __shared__-free trick: global g_it per block)
rep("    __syncthreads();\n    // ---- phase 2: fit + Jacobian row, one thread per query", "    __syncthreads();\n    if (threadIdx.x == 0 && dbg_it >= 0 && dbg_it < 16) g_ts[blockIdx.x][dbg_it][1] = gtime();\n    // ---- phase 2: fit + Jacobian row, one thread per query")
rep("      if (A.mode == 0) {\n        if (warp * 32 < nslots) {  // warp-uniform", "      if (threadIdx.x == 0 && dbg_it >= 0 && dbg_it < 16) g_ts[blockIdx.x][dbg_it][2] = gtime();\n      if (A.mode == 0) {\n        if (warp * 32 < nslots) {  // warp-uniform")
rep("                                               int q_begin, int q_end, const float4* preq = nullptr) {\n  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;\n  float run_sum = 0.f;", "                                               int q_begin, int q_end, const float4* preq = nullptr, int dbg_it = -1) {\n  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;\n  float run_sum = 0.f;")
rep("    const float run_sum = lm_block_pass(A, trig, rt, geom, s_slot, s_wacc, q_begin, q_end, one_batch ? preq : nullptr);\n    float* part =", "    if (tid == 0 && it < 16) g_ts[blockIdx.x][it][0] = gtime();\n    const float run_sum = lm_block_pass(A, trig, rt, geom, s_slot, s_wacc, q_begin, q_end, one_batch ? preq : nullptr, (int)it);\n    if (tid == 0 && it < 16) g_ts[blockIdx.x][it][3] = gtime();\n    float* part =")
rep("      while (ld_acquire_u32(barrier) < target) {}\n    }\n    __syncthreads();\n    reduce_partials_and_solve(part, gridDim.x, A.prm, &s_st, false, reinterpret_cast<double*>(s_slot));",
    "      if (it < 16) g_ts[blockIdx.x][it][4] = gtime();\n      while (ld_acquire_u32(barrier) < target) {}\n      if (it < 16) g_ts[blockIdx.x][it][5] = gtime();\n    }\n    __syncthreads();\n    reduce_partials_and_solve(part, gridDim.x, A.prm, &s_st, false, reinterpret_cast<double*>(s_slot));\n    if (tid == 0 && it < 16) g_ts[blockIdx.x][it][6] = gtime();")
rep("    if (tid == 0) *barrier_clear = 0u;\n", "    if (tid == 0) *barrier_clear = 0u;\n    if (tid == 0) {\n      printf(\"per-block (p1,fit,rest) it5:\"); for (int b = 0; b < (int)gridDim.x; ++b) { volatile unsigned long long* g = g_ts[b][5]; printf(\" %d:%llu/%llu/%llu\", b, (g[1]-g[0])/100, (g[2]-g[1])/100, (g[3]-g[2])/100); } printf(\"\\n\");\n      for (int it = 0; it < 10; ++it) {\n        unsigned long long t0 = ~0ull, p1 = 0, p2 = 0, rd = 0, pass = 0, arrive_last = 0, rel_first = ~0ull, rel_last = 0, endmax = 0; int b1 = 0, b2 = 0;\n        for (int b = 0; b < (int)gridDim.x; ++b) { volatile unsigned long long* g = g_ts[b][it]; if (g[0] < t0) t0 = g[0]; if (g[1] - g[0] > p1) { p1 = g[1] - g[0]; b1 = b; } if (g[2] - g[1] > p2) { p2 = g[2] - g[1]; b2 = b; } if (g[3] - g[2] > rd) rd = g[3] - g[2]; if (g[3] - g[0] > pass) pass = g[3] - g[0]; if (g[4] > arrive_last) arrive_last = g[4]; if (g[5] < rel_first) rel_first = g[5]; if (g[5] > rel_last) rel_last = g[5]; if (g[6] > endmax) endmax = g[6]; }\n        volatile unsigned long long* h = g_ts[100][it];\n        printf(\"it %d: p1 max %llu (blk %d) b100 %llu | fit max %llu (blk %d) b100 %llu | blkreduce max %llu | pass max %llu | last arrive at %llu, release first %llu last %llu | b100 solve %llu | iter span %llu\\n\", it, p1, b1, h[1]-h[0], p2, b2, h[2]-h[1], rd, pass, arrive_last - t0, rel_first - t0, rel_last - t0, h[6]-h[5], endmax - t0);\n      }\n    }\n")

rep("__device__ unsigned long long g_ts[160][16][8];", "__device__ unsigned long long g_ts[160][16][8];\n__device__ unsigned long long g_kb[16];")
rep("  constexpr int NB = LM_KNN_BATCH;\n", "  constexpr int NB = LM_KNN_BATCH;\n  const bool dbg = blockIdx.x == 100 && threadIdx.x == 0;\n  if (dbg) g_kb[0] = gtime();\n")
rep("  // ---- stage 2: ranges -> candidate addresses", "  if (dbg) g_kb[1] = gtime() + (ce[0].x & 0) + (ce[1].x & 0) + (ce[2].x & 0);\n  // ---- stage 2: ranges -> candidate addresses")
rep("  // ---- stage 3: distances and the per-lane sorting network", "  if (dbg) g_kb[2] = gtime();\n  // ---- stage 3: distances and the per-lane sorting network")
rep("  // ---- stage 4: five selection rounds", "  if (dbg) g_kb[3] = gtime() + (cdq[0][0] & 0) + (cdq[1][0] & 0) + (cdq[2][0] & 0);\n  // ---- stage 4: five selection rounds")
rep("  // ---- queries with more than one chunk of candidates: the remaining chunks, merged into the five found so far", "  if (dbg) g_kb[4] = gtime() + (od[0] & 0) + (od[1] & 0) + (od[2] & 0);\n  // ---- queries with more than one chunk of candidates: the remaining chunks, merged into the five found so far")
rep("      if (it < 16) g_ts[blockIdx.x][it][4] = gtime();\n", "      if (it < 16) g_ts[blockIdx.x][it][4] = gtime();\n      if (blockIdx.x == 100) printf(\"it %u b100 w0 knn batch: entry +%llu | cells %llu | issue %llu | dist+sort %llu | select %llu | to p1 end %llu\\n\", it, g_kb[0] - g_ts[100][it][0], g_kb[1] - g_kb[0], g_kb[2] - g_kb[1], g_kb[3] - g_kb[2], g_kb[4] - g_kb[3], g_ts[100][it][1] - g_kb[4]);\n")
open(p,'w').write(s)
PY
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -cudart static -o scratch/var/lib_timing.so scratch/var/timing/csrc/abi.cu
ls -la scratch/var/lib_timing.so
