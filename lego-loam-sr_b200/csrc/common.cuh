// Shared declarations for the sm_100a scan-to-map kernels.
//
// Numerics contract: this library is compiled with -fmad=false, so every float
// expression below is evaluated as written, IEEE round-to-nearest, with correctly rounded division
// and square root (nvcc defaults -prec-div=true -prec-sqrt=true). That is what makes voxel
// membership, grid cells and neighbour distances bit-identical to the reference's x86-64 build,
// which has no FMA contraction either (LeGO-LOAM/CMakeLists.txt:9-11: -O3, no -march).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ll {

constexpr int kNumSMsHint = 148;  // B200: 2 dies x 74 SMs; actual count read at ll_create

// One rigid transform, the reference's cached sin/cos members
// (updatePointAssociateToMapSinCos mapOptmization.cpp:591-604, updateTransformPointCloudSinCos :622-635).
struct Trig {
  float cr, sr, cp, sp, cy, sy, tx, ty, tz;
};

// pointAssociateToMap (mapOptmization.cpp:606-620) and the body of transformPointCloud (:649-664):
// rotate about z (rz), then x (rx), then y (ry), then translate. Operation order kept.
__device__ __forceinline__ float4 rigid_transform(const Trig& t, const float4 p) {
  float x1 = t.cy * p.x - t.sy * p.y;
  float y1 = t.sy * p.x + t.cy * p.y;
  float z1 = p.z;
  float x2 = x1;
  float y2 = t.cr * y1 - t.sr * z1;
  float z2 = t.sr * y1 + t.cr * z1;
  float4 o;
  o.x = t.cp * x2 + t.sp * z2 + t.tx;
  o.y = y2 + t.ty;
  o.z = -t.sp * x2 + t.cp * z2 + t.tz;
  o.w = p.w;
  return o;
}

// Monotone float <-> uint mapping for atomicMin / atomicMax on floats.
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ bool finite3(const float4 p) {
  return isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
}

// Programmatic dependent launch (PDL). Every kernel of the per-scan chain is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and starts with pdl_entry(): it waits until the previous
// kernel's memory is visible, then lets the next kernel begin its own launch (block scheduling, parameter
// fetch) while this one is still running. The chain is ~25 short dependent kernels per scan, so launch-to-launch
// latency is a first-order cost. Both instructions are no-ops under a plain launch.
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Barrier across a co-resident grid (cooperative launch): `counter` starts at zero and only grows during the
// launch; `epoch` counts the barriers this thread has passed (the same in every thread of the grid).
#ifdef LL_TIMELINE
// scratch/build_timeline.sh: block 0 stamps %globaltimer when it arrives at and when it leaves every grid barrier of a launch
__device__ unsigned long long g_tl_bar[512][2];
__device__ __forceinline__ unsigned long long tl_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define LL_TL_BAR(e, k) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (e) < 512) g_tl_bar[(e)][(k)] = tl_globaltimer(); } while (0)
// ... and thread 0 of block 0 at named points between the barriers of the map assembly (printed with the barrier list)
__device__ unsigned long long g_tl_sub[64];
#define LL_TL_SUB(id) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_tl_sub[(id)] = tl_globaltimer(); } while (0)
#else
#define LL_TL_BAR(e, k) do {} while (0)
#define LL_TL_SUB(id) do {} while (0)
#endif

__device__ __forceinline__ void coop_barrier(unsigned* counter, unsigned& epoch) {
  ++epoch;
  __syncthreads();
  LL_TL_BAR(epoch, 0);
  if (threadIdx.x == 0) {
    // arrive: one release-reduction (cumulative: it orders the block's earlier writes, observed through bar.sync,
    // before the increment — what a __threadfence() + atomicAdd pair does in two round trips)
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
    const unsigned target = epoch * gridDim.x;
    while (ld_acquire_u32(counter) < target) {}  // (relaxed polls + one fence at the end were measured: +6 us per scan)
  }
  LL_TL_BAR(epoch, 1);
  __syncthreads();
}

// Streaming 128-bit accesses: raw clouds are read once / written once per stage.
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }

}  // namespace ll
