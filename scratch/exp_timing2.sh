#!/bin/bash
set -e
cd /root/repo
rm -rf scratch/var/timing && mkdir -p scratch/var/timing/csrc scratch/var/include
cp lego-loam-sr_b200/csrc/* scratch/var/timing/csrc/
cp include/legoloam_b200.h scratch/var/include/
python - <<'PY'
p='scratch/var/timing/csrc/lm.cuh'
s=open(p).read()
def rep(a,b,cnt=1):
    global s
    assert s.count(a)>=1, a
    s=s.replace(a,b,cnt)
rep("namespace ll {\n\nconstexpr int LM_THREADS = 512;", "namespace ll {\n__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile(\"mov.u64 %0, %globaltimer;\" : \"=l\"(t)); return t; }\n__device__ unsigned long long g_ts[1024][8];\n__device__ unsigned long long g_tail[8];\n\nconstexpr int LM_THREADS = 512;")
rep("  pdl_entry();\n  if (A.mode == 0 && S->done) return;\n  __shared__ __align__(16) float s_slot", "  pdl_entry();\n  if (threadIdx.x == 0) g_ts[blockIdx.x][0] = gtime();\n  if (A.mode == 0 && S->done) return;\n  __shared__ __align__(16) float s_slot")
rep("    __syncthreads();\n    // ---- phase 2: fit + Jacobian row, one thread per query", "    __syncthreads();\n    if (threadIdx.x == 0) g_ts[blockIdx.x][1] = gtime();\n    // ---- phase 2: fit + Jacobian row, one thread per query")
rep("      if (A.mode == 0) {\n        if (warp * 32 < nslots) {  // warp-uniform", "      if (threadIdx.x == 0) g_ts[blockIdx.x][4] = gtime();\n      if (A.mode == 0) {\n        if (warp * 32 < nslots) {  // warp-uniform")
rep("  if (A.mode != 0) return;\n  if (tid < 32) {\n    A.partials[blockIdx.x * 32 + tid]", "  if (A.mode != 0) return;\n  if (threadIdx.x == 0) g_ts[blockIdx.x][2] = gtime();\n  if (tid < 32) {\n    A.partials[blockIdx.x * 32 + tid]")
rep("  if (tid == 0) S->ticket = 0;\n  reduce_partials_and_solve(A.partials, gridDim.x, A.prm, S, false, reinterpret_cast<double*>(s_slot));",
    "  if (tid == 0) S->ticket = 0;\n  if (threadIdx.x == 0) g_ts[blockIdx.x][3] = gtime();\n  const int it_dbg = S->iter;\n  reduce_partials_and_solve(A.partials, gridDim.x, A.prm, S, false, reinterpret_cast<double*>(s_slot));\n  if (threadIdx.x == 0) {\n    unsigned long long t5 = gtime(), t0min = ~0ull, p1 = 0, p2 = 0, pf = 0; int b1 = 0, b2 = 0;\n    for (int b = 0; b < (int)gridDim.x; ++b) { unsigned long long a0 = *(volatile unsigned long long*)&g_ts[b][0], a1 = *(volatile unsigned long long*)&g_ts[b][1], a2 = *(volatile unsigned long long*)&g_ts[b][2], a4 = *(volatile unsigned long long*)&g_ts[b][4]; if (a0 < t0min) t0min = a0; if (a1 - a0 > p1) { p1 = a1 - a0; b1 = b; } if (a2 - a1 > p2) { p2 = a2 - a1; b2 = b; } if (a4 - a1 > pf) pf = a4 - a1; }\n    printf(\"iter %d: phase1 max %llu (blk %d) [blk0 %llu blk100 %llu], phase2 max %llu (blk %d) [blk0 %llu blk100 %llu] fit-only max %llu, tail: reduce %llu, AtA+qr/jacobi %llu, finish %llu, total %llu\\n\", it_dbg, p1, b1, g_ts[0][1]-g_ts[0][0], g_ts[100][1]-g_ts[100][0], p2, b2, g_ts[0][2]-g_ts[0][1], g_ts[100][2]-g_ts[100][1], pf, g_tail[0] - g_ts[blockIdx.x][3], g_tail[1] - g_tail[0], t5 - g_tail[1], t5 - t0min);\n  }")
# stamps inside solve
rep("  __syncthreads();\n  lm_solve_block(s_red, prm, S, stage_mode);", "  __syncthreads();\n  if (threadIdx.x == 0) g_tail[0] = gtime();\n  lm_solve_block(s_red, prm, S, stage_mode);")
rep("  __syncthreads();\n  if (warp != 0) return;\n  // ---- the rest of LMOptimization", "  __syncthreads();\n  if (threadIdx.x == 0) g_tail[1] = gtime();\n  if (warp != 0) return;\n  // ---- the rest of LMOptimization")
rep("  float x = lane < 6 ? xs[lane] : 0.f;\n  if (degenerate && lane < 6) {  // matX = matP * matX2", "  if (threadIdx.x == 0) g_tail[2] = gtime();\n  float x = lane < 6 ? xs[lane] : 0.f;\n  if (degenerate && lane < 6) {  // matX = matP * matX2")
rep("  const bool pose_changed = prm.apply_update != 0;\n  if (pose_changed) {  // updatePointAssociateToMapSinCos", "  const bool pose_changed = prm.apply_update != 0;\n  if (threadIdx.x == 0) g_tail[3] = gtime();\n  if (pose_changed) {  // updatePointAssociateToMapSinCos")
rep("  if (lane == 0) {\n    S->sel_num = n;", "  if (threadIdx.x == 0) g_tail[4] = gtime();\n  if (lane == 0) {\n    S->sel_num = n;")
rep("finish %llu, total %llu\\n\", it_dbg,", "finish %llu (eigstore %llu, update %llu, trig %llu, ctl %llu), total %llu\\n\", it_dbg,")
rep("t5 - g_tail[1], t5 - t0min);", "t5 - g_tail[1], g_tail[2]-g_tail[1], g_tail[3]-g_tail[2], g_tail[4]-g_tail[3], t5-g_tail[4], t5 - t0min);")

rep("  const int cx = (int)floorf(sx) - g.ox, cy = (int)floorf(sy) - g.oy, cz = (int)floorf(sz) - g.oz;\n  int2 ce = make_int2(0, 0);",
    "  const bool dbg = blockIdx.x == 100 && threadIdx.x == 0;\n  if (dbg) g_knn[g_knn_n++ & 63] = gtime();\n  const int cx = (int)floorf(sx) - g.ox, cy = (int)floorf(sy) - g.oy, cz = (int)floorf(sz) - g.oz;\n  int2 ce = make_int2(0, 0);")
rep("  const uint32_t max_bits = __float_as_uint(max_sq_dist);\n  Knn5 K;", "  if (dbg) g_knn[g_knn_n++ & 63] = gtime() + (total & 0);\n  const uint32_t max_bits = __float_as_uint(max_sq_dist);\n  Knn5 K;")
rep("    static_assert(LM_CAND_UNROLL == 4,", "    if (dbg) g_knn[g_knn_n++ & 63] = gtime() + (cd[0] & 0);\n    static_assert(LM_CAND_UNROLL == 4,")
rep("  K.accept = K.d2[4] != LM_EMPTY_D2;", "  if (dbg) g_knn[g_knn_n++ & 63] = gtime() + (K.d2[4] & 0);\n  K.accept = K.d2[4] != LM_EMPTY_D2;")
rep("__device__ unsigned long long g_tail[8];", "__device__ unsigned long long g_tail[8];\n__device__ unsigned long long g_knn[64];\n__device__ int g_knn_n;")
rep("    __syncthreads();\n    if (threadIdx.x == 0) g_ts[blockIdx.x][1] = gtime();", "    if (blockIdx.x == 100 && threadIdx.x == 0) g_knn[g_knn_n++ & 63] = gtime();\n    __syncthreads();\n    if (threadIdx.x == 0) g_ts[blockIdx.x][1] = gtime();")
rep("  const int it_dbg = S->iter;", "  const int it_dbg = S->iter;\n  if (threadIdx.x == 0) { printf(\"knn stamps blk100 warp0 (n=%d):\", g_knn_n); for (int i = 0; i < g_knn_n && i < 20; ++i) printf(\" %llu\", g_knn[i] - g_ts[100][0]); printf(\"\\n\"); g_knn_n = 0; }")
open(p,'w').write(s)
PY
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -cudart static -o scratch/var/lib_timing.so scratch/var/timing/csrc/abi.cu
