#!/bin/bash
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
    assert a in s, a
    s=s.replace(a,b,1)
rep("namespace ll {\n\nconstexpr int LM_THREADS = 512;", "namespace ll {\n__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile(\"mov.u64 %0, %globaltimer;\" : \"=l\"(t)); return t; }\n__device__ unsigned long long g_ts[1024][8];\n\nconstexpr int LM_THREADS = 512;")
# kernel start
rep("  LMState* S = A.st;\n  if (A.mode == 0 && S->done) return;\n  __shared__ __align__(16) float s_slot", "  LMState* S = A.st;\n  if (threadIdx.x == 0) g_ts[blockIdx.x][0] = gtime();\n  if (A.mode == 0 && S->done) return;\n  __shared__ __align__(16) float s_slot")
rep("    __syncthreads();\n    // ---- phase 2: fit + Jacobian row, one thread per query", "    __syncthreads();\n    if (threadIdx.x == 0) g_ts[blockIdx.x][1] = gtime();\n    // ---- phase 2: fit + Jacobian row, one thread per query")
rep("  if (A.mode != 0) return;\n  if (tid < 32) {\n    A.partials[blockIdx.x * 32 + tid]", "  if (A.mode != 0) return;\n  if (threadIdx.x == 0) g_ts[blockIdx.x][2] = gtime();\n  if (tid < 32) {\n    A.partials[blockIdx.x * 32 + tid]")
rep("  if (!s_last) return;\n  __threadfence();\n  if (tid == 0) S->ticket = 0;\n  reduce_partials_and_solve(A.partials, gridDim.x, A.prm, S, false, reinterpret_cast<double*>(s_slot));",
    "  if (!s_last) return;\n  __threadfence();\n  if (tid == 0) S->ticket = 0;\n  if (threadIdx.x == 0) g_ts[blockIdx.x][3] = gtime();\n  const int it_dbg = S->iter;\n  reduce_partials_and_solve(A.partials, gridDim.x, A.prm, S, false, reinterpret_cast<double*>(s_slot));\n  if (threadIdx.x == 0) {\n    unsigned long long t5 = gtime(), t0min = ~0ull, t0max = 0, p1 = 0, p2 = 0, t2max = 0;\n    for (int b = 0; b < (int)gridDim.x; ++b) { unsigned long long a0 = *(volatile unsigned long long*)&g_ts[b][0], a1 = *(volatile unsigned long long*)&g_ts[b][1], a2 = *(volatile unsigned long long*)&g_ts[b][2]; if (a0 < t0min) t0min = a0; if (a0 > t0max) t0max = a0; if (a1 - a0 > p1) p1 = a1 - a0; if (a2 - a1 > p2) p2 = a2 - a1; if (a2 > t2max) t2max = a2; }\n    printf(\"iter %d: start spread %llu ns, phase1 max %llu, phase2 max %llu, all-blocks-done %llu, last-block-ticket %llu, reduce+solve %llu, total %llu\\n\", it_dbg, t0max - t0min, p1, p2, t2max - t0min, g_ts[blockIdx.x][3] - t0min, t5 - g_ts[blockIdx.x][3], t5 - t0min);\n  }")
open(p,'w').write(s)
PY
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -cudart static -o scratch/var/lib_timing.so scratch/var/timing/csrc/abi.cu
ls -la scratch/var/lib_timing.so
