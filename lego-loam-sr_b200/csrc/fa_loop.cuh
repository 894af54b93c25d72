// SURVEY.md §8f ranks 1 + 3 together: FeatureAssociation::updateTransformation (featureAssociation.cpp:2505-2535) as ONE
// cooperative launch — up to 100 iterations of findCorrespondingSurfFeatures + calculateTransformationSurf (:1846-2010;
// unknowns rx, rz, ty), then up to 100 of findCorrespondingCornerFeatures + calculateTransformationCorner (:2013-2143;
// unknowns ry, tx, tz). The correspondence records of fa_corr.cuh feed the Jacobian rows, the rows feed the 3x3 normal
// equations, the solve updates transformCur, and nothing returns to the host until both loops have ended.
//
// Per iteration, per block (every block owns the feature points b, b + gridDim, b + 2 gridDim, ...):
//   search iteration (iterCount % 5 == 0)  one warp per feature point: TransformToStart, exact 1-NN + ring-window walks
//                                          (fa_search_warp), indices kept in global memory for the next four iterations;
//   reuse iteration                        one thread per feature point: TransformToStart + the stored indices;
//   both                                   the (laserCloudOri, coeffSel) record (fa_record), the row of matA and the entry
//                                          of matB (:1926-1948 / :2050-2066; float arithmetic as written, -fmad=false),
//                                          10 running sums in double (6 AtA + 3 AtB + count) -> block partial -> grid
//                                          barrier -> EVERY block adds all partials in the same order and runs the same
//                                          3x3 solve (neq_solve_small: Eigen's column-pivoted QR sequence, projector at
//                                          iterCount == 0), so every block holds the same transformCur and loop decision.
// State that outlives a call in the reference: isDegenerate (class member shared by both loops, :223) and matP (a
// function-local that is assigned only at iterCount == 0 but read whenever isDegenerate is set — kept per function here,
// see oracle.h). The weights switch on at iterCount >= 5 (:1671, :1817), the loop skips the solve below 10 pairs (:2514).
#pragma once
#include "fa_corr.cuh"
#include "normal_eq.cuh"

namespace ll {

constexpr int FAL_THREADS = 256;
constexpr int FAL_WARPS = FAL_THREADS / 32;
constexpr int FAL_NSUM = 10;       // AtA 00 01 02 11 12 22 | AtB 0 1 2 | count
constexpr int FAL_PART = 16;       // doubles per block partial (padded)
constexpr int FAL_MAX_ITERS = 100;

struct FaLoopState {
  float tc[6];        // transformCur
  int is_degenerate;  // isDegenerate
  float matP[2][9];   // [0] calculateTransformationCorner's, [1] calculateTransformationSurf's; row-major
  int iters[2];       // [0] iterCount1 (surf loop), [1] iterCount2 (corner loop) at loop exit
  int skipped;        // the :2506 guard returned
  int n_sel[2];       // laserCloudOri size of the last executed iteration: [0] surf loop, [1] corner loop
  int pad;
};

struct FaLoopArgs {
  const float4* q[2];     // [0] cornerPointsSharp, [1] surfPointsFlat
  const float4* last[2];  // [0] laserCloudCornerLast, [1] laserCloudSurfLast
  int nq[2], n_last[2];
  float dist_sqr;
  int eigen_mode;
  int* ind;               // [3][ind_cap]: pointSearch*Ind1/2/3 of the loop in progress
  int ind_cap;
  double* partials;       // [2][gridDim.x][FAL_PART], double-buffered by iteration parity
  FaLoopState* st;        // in / out
  FaLoopState* st_host;   // pinned mirror (nullable)
  float* trace;           // nullable: [2][FAL_MAX_ITERS][4] = matX, row count (surf loop first)
};

// sin / cos of transformCur and the products calculateTransformationSurf / Corner derive from them (:1884-1923,
// :2025-2044). Float arithmetic as written; sin / cos are the double functions rounded to float.
struct FaRowConst {
  // surf
  float a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, b1, b2, b5, b6, c1, c2, c3, c4, c5, c6, c7, c8, c9, crx;
  // corner
  float kb1, kb2, kb3, kb4, kb5, kb6, kb7, kb8, kc5, srx;
};

// sc = {sin rx, cos rx, sin ry, cos ry, sin rz, cos rz} of transformCur
__device__ __forceinline__ void fa_row_const(const float* tc, const float* sc, FaRowConst& k) {
  const float srx = sc[0], crx = sc[1], sry = sc[2], cry = sc[3], srz = sc[4], crz = sc[5];
  const float tx = tc[3], ty = tc[4], tz = tc[5];
  k.a1 = crx * sry * srz;
  k.a2 = crx * crz * sry;
  k.a3 = srx * sry;
  k.a4 = tx * k.a1 - ty * k.a2 - tz * k.a3;
  k.a5 = srx * srz;
  k.a6 = crz * srx;
  k.a7 = ty * k.a6 - tz * crx - tx * k.a5;
  k.a8 = crx * cry * srz;
  k.a9 = crx * cry * crz;
  k.a10 = cry * srx;
  k.a11 = tz * k.a10 + ty * k.a9 - tx * k.a8;
  k.b1 = -crz * sry - cry * srx * srz;
  k.b2 = cry * crz * srx - sry * srz;
  k.b5 = cry * crz - srx * sry * srz;
  k.b6 = cry * srz + crz * srx * sry;
  k.c1 = -k.b6;
  k.c2 = k.b5;
  k.c3 = tx * k.b6 - ty * k.b5;
  k.c4 = -crx * crz;
  k.c5 = crx * srz;
  k.c6 = ty * k.c5 + tx * -k.c4;
  k.c7 = k.b2;
  k.c8 = -k.b1;
  k.c9 = tx * -k.b2 - ty * -k.b1;
  k.crx = crx;
  k.kb1 = k.b1;
  k.kb2 = k.b2;
  k.kb3 = crx * cry;
  k.kb4 = tx * -k.kb1 + ty * -k.kb2 + tz * k.kb3;
  k.kb5 = k.b5;
  k.kb6 = k.b6;
  k.kb7 = crx * sry;
  k.kb8 = tz * k.kb7 - ty * k.kb6 - tx * k.kb5;
  k.kc5 = crx * srz;
  k.srx = srx;
}

// One row of matA and entry of matB. o = the feature point as pushed to laserCloudOri (not de-skewed), c = coeffSel.
__device__ __forceinline__ void fa_row(const FaRowConst& k, int which, const float4 o, const float (&c)[4], float (&r)[3],
                                       float& b) {
  if (which) {  // calculateTransformationSurf :1926-1948: arx, arz, aty
    r[0] = (-k.a1 * o.x + k.a2 * o.y + k.a3 * o.z + k.a4) * c[0] + (k.a5 * o.x - k.a6 * o.y + k.crx * o.z + k.a7) * c[1] +
           (k.a8 * o.x - k.a9 * o.y - k.a10 * o.z + k.a11) * c[2];
    r[1] = (k.c1 * o.x + k.c2 * o.y + k.c3) * c[0] + (k.c4 * o.x - k.c5 * o.y + k.c6) * c[1] +
           (k.c7 * o.x + k.c8 * o.y + k.c9) * c[2];
    r[2] = -k.b6 * c[0] + k.c4 * c[1] + k.b2 * c[2];
  } else {      // calculateTransformationCorner :2050-2066: ary, atx, atz
    r[0] = (k.kb1 * o.x + k.kb2 * o.y - k.kb3 * o.z + k.kb4) * c[0] + (k.kb5 * o.x + k.kb6 * o.y - k.kb7 * o.z + k.kb8) * c[2];
    r[1] = -k.kb5 * c[0] + k.kc5 * c[1] + k.kb1 * c[2];
    r[2] = k.kb7 * c[0] - k.srx * c[1] - k.kb3 * c[2];
  }
  b = (float)(-0.05 * (double)c[3]);
}

__device__ __forceinline__ void fa_accumulate(const float (&r)[3], float b, double (&acc)[FAL_NSUM]) {
  acc[0] += (double)(r[0] * r[0]);
  acc[1] += (double)(r[0] * r[1]);
  acc[2] += (double)(r[0] * r[2]);
  acc[3] += (double)(r[1] * r[1]);
  acc[4] += (double)(r[1] * r[2]);
  acc[5] += (double)(r[2] * r[2]);
  acc[6] += (double)(r[0] * b);
  acc[7] += (double)(r[1] * b);
  acc[8] += (double)(r[2] * b);
  acc[9] += 1.0;
}

// The end of calculateTransformationSurf / Corner (:1992-2009, :2127-2142) by one thread: transformCur update, NaN reset,
// deltaR / deltaT. Returns true when the function returns false (converged: the caller breaks).
__device__ __forceinline__ bool fa_apply(int which, const float* x, float* tc) {
  float deltaR, deltaT;
  if (which) {
    tc[0] += x[0];
    tc[2] += x[1];
    tc[4] += x[2];
    const double r0 = (double)(57.29578f * x[0]), r1 = (double)(57.29578f * x[1]), t0 = (double)(x[2] * 100.f);
    deltaR = (float)sqrt(r0 * r0 + r1 * r1);
    deltaT = (float)sqrt(t0 * t0);
  } else {
    tc[1] += x[0];
    tc[3] += x[1];
    tc[5] += x[2];
    const double r0 = (double)(57.29578f * x[0]), t0 = (double)(x[1] * 100.f), t1 = (double)(x[2] * 100.f);
    deltaR = (float)sqrt(r0 * r0);
    deltaT = (float)sqrt(t0 * t0 + t1 * t1);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if (isnan(tc[k])) tc[k] = 0.f;
  return (double)deltaR < 0.1 && (double)deltaT < 0.1;
}

// ---- staging of the previous sweep's cloud: a regular, contiguous tile that every block reads in full at every search
// iteration — the case BASELINE.json's north_star names for TMA bulk copies. One thread arms an mbarrier with the byte count
// and issues cp.async.bulk (global -> shared, completion counted on the mbarrier); everybody waits on the phase bit.
__device__ __forceinline__ uint32_t fal_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fal_mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fal_smem_addr(mbar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fal_stage_issue(float4* dst, const float4* src, uint32_t bytes, uint64_t* mbar) {
  // generic-proxy accesses of the destination (the other cloud was read from it) are ordered before the async-proxy writes
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fal_smem_addr(mbar)), "r"(bytes) : "memory");
  constexpr uint32_t kChunk = 65536;
  for (uint32_t off = 0; off < bytes; off += kChunk) {
    const uint32_t n = min(kChunk, bytes - off);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                     "r"(fal_smem_addr(reinterpret_cast<const char*>(dst) + off)), "l"(reinterpret_cast<const char*>(src) + off), "r"(n),
                 "r"(fal_smem_addr(mbar))
                 : "memory");
  }
}
__device__ __forceinline__ void fal_mbar_wait(uint64_t* mbar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(fal_smem_addr(mbar)),
      "r"(parity)
      : "memory");
}

template <bool kStage>
__global__ void __launch_bounds__(FAL_THREADS) fa_loop_kernel(const __grid_constant__ FaLoopArgs A, unsigned* barrier,
                                                              unsigned* barrier_clear) {
  extern __shared__ __align__(128) unsigned char s_stage[];  // kStage: the cloud of the loop in progress
  __shared__ __align__(8) uint64_t s_mbar;
  __shared__ FaLoopState s_st;
  __shared__ FaRowConst s_k;
  __shared__ float s_sc[6];
  __shared__ double s_w[FAL_WARPS][FAL_NSUM];
  __shared__ double s_sum[FAL_NSUM];
  __shared__ int s_break;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = (int)gridDim.x, bid = (int)blockIdx.x;
  {
    static_assert(sizeof(FaLoopState) % 4 == 0, "FaLoopState is copied word by word");
    const uint32_t* src = reinterpret_cast<const uint32_t*>(A.st);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&s_st);
    if (tid < (int)(sizeof(FaLoopState) / 4)) dst[tid] = __ldcg(src + tid);
  }
  __syncthreads();
  unsigned epoch = 0;
  const bool guard = A.n_last[0] < 10 || A.n_last[1] < 100;  // :2506 (the same decision in every block)
  if (tid == 0) {
    s_st.skipped = guard ? 1 : 0;
    s_st.iters[0] = s_st.iters[1] = 0;
    s_st.n_sel[0] = s_st.n_sel[1] = 0;
  }
  if (kStage && tid == 0) fal_mbar_init(&s_mbar, 1);
  for (int loop = 0; loop < 2 && !guard; ++loop) {
    const int which = loop == 0 ? 1 : 0;
    FaArgs F;
    F.q = A.q[which];
    F.last = A.last[which];
    if constexpr (kStage) {
      __syncthreads();  // (the mbarrier is initialised; nobody reads the previous loop's cloud any more)
      if (tid == 0) fal_stage_issue(reinterpret_cast<float4*>(s_stage), A.last[which], (uint32_t)A.n_last[which] * 16u, &s_mbar);
      fal_mbar_wait(&s_mbar, (uint32_t)(loop & 1));
      F.last = reinterpret_cast<const float4*>(s_stage);
    }
    F.nq = A.nq[which];
    F.n_last = A.n_last[which];
    F.which = which;
    F.dist_sqr = A.dist_sqr;
    F.ind1 = A.ind;
    F.ind2 = A.ind + A.ind_cap;
    F.ind3 = A.ind + 2 * A.ind_cap;
    F.coeff = nullptr;
    F.keep = nullptr;
    const int n_slots = F.nq > bid ? (F.nq - bid + nb - 1) / nb : 0;  // this block's feature points: bid + nb * s
    int it = 0;
    for (; it < FAL_MAX_ITERS; ++it) {
      __syncthreads();
      if (tid < 3) {  // sin / cos of the three angles side by side (double precision, rounded to float: DESIGN.md §3)
        double sn, cs;
        sincos((double)s_st.tc[tid], &sn, &cs);
        s_sc[2 * tid] = (float)sn;
        s_sc[2 * tid + 1] = (float)cs;
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) F.tc[k] = s_st.tc[k];
      F.iter = it;
      __syncthreads();
      if (tid == 0) fa_row_const(s_st.tc, s_sc, s_k);
      __syncthreads();
      double acc[FAL_NSUM];
#pragma unroll
      for (int k = 0; k < FAL_NSUM; ++k) acc[k] = 0.0;
      if (it % 5 == 0) {  // search: one warp per feature point
        for (int s = warp; s < n_slots; s += FAL_WARPS) {
          const int i = bid + nb * s;
          const float4 p = __ldg(F.q + i);
          const float4 sel = fa_transform_to_start(p, F.tc);
          int closest, min2, min3;
          fa_search_warp<kStage>(F, sel, lane, closest, min2, min3);
          if (lane == 0) {
            F.ind1[i] = closest;
            F.ind2[i] = min2;
            F.ind3[i] = min3;
            float c[4], r[3], b;
            if (fa_record<kStage>(F, sel, closest, min2, min3, c)) {
              fa_row(s_k, which, p, c, r, b);
              fa_accumulate(r, b, acc);
            }
          }
        }
      } else {            // reuse: one thread per feature point, consecutive slots on different warps
        for (int s = lane * FAL_WARPS + warp; s < n_slots; s += FAL_THREADS) {
          const int i = bid + nb * s;
          const float4 p = __ldg(F.q + i);
          const float4 sel = fa_transform_to_start(p, F.tc);
          const int closest = __ldcg(F.ind1 + i), min2 = __ldcg(F.ind2 + i), min3 = __ldcg(F.ind3 + i);
          float c[4], r[3], b;
          if (fa_record<kStage>(F, sel, closest, min2, min3, c)) {
            fa_row(s_k, which, p, c, r, b);
            fa_accumulate(r, b, acc);
          }
        }
      }
      // ---- block partial (fixed order: shuffle tree, then the warps in turn)
#pragma unroll
      for (int k = 0; k < FAL_NSUM; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_w[warp][k] = v;
      }
      __syncthreads();
      double* part = A.partials + (size_t)(it & 1) * nb * FAL_PART;
      if (tid < FAL_NSUM) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < FAL_WARPS; ++w) t += s_w[w][tid];
        part[bid * FAL_PART + tid] = t;
      }
      coop_barrier(barrier, epoch);
      // ---- every block: all partials in the same order (16 partial sums per quantity, then a fixed tree)
      if (tid < FAL_NSUM * 16) {
        const int k = tid >> 4, j = tid & 15;
        double t = 0.0;
        for (int b0 = j; b0 < nb; b0 += 16 * 8) {  // eight partials in flight per thread, added in block order
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = b0 + 16 * u < nb ? __ldcg(part + (b0 + 16 * u) * FAL_PART + k) : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) t += v[u];
        }
        t += __shfl_xor_sync(0xffffffffu, t, 8, 16);
        t += __shfl_xor_sync(0xffffffffu, t, 4, 16);
        t += __shfl_xor_sync(0xffffffffu, t, 2, 16);
        t += __shfl_xor_sync(0xffffffffu, t, 1, 16);
        if (j == 0) s_sum[k] = t;
      }
      __syncthreads();
      if (tid == 0) {
        const int n = (int)(s_sum[9] + 0.5);
        bool brk = false;
        float x[3] = {0.f, 0.f, 0.f};
        if (n >= 10) {  // :2514 / :2527
          double sums[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) sums[k] = s_sum[k];
          neq_solve_small(sums, 3, it, 10.f, A.eigen_mode, x, s_st.matP[which], &s_st.is_degenerate);
          brk = fa_apply(which, x, s_st.tc);
        }
        s_st.n_sel[loop] = n;
        s_break = brk ? 1 : 0;
        if (A.trace && bid == 0) {
          float* t = A.trace + ((size_t)loop * FAL_MAX_ITERS + it) * 4;
          t[0] = x[0]; t[1] = x[1]; t[2] = x[2]; t[3] = (float)n;
        }
      }
      __syncthreads();
      if (s_break) break;
    }
    if (tid == 0) s_st.iters[loop] = it;
  }
  __syncthreads();
  if (bid == 0) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&s_st);
    if (tid < (int)(sizeof(FaLoopState) / 4)) {
      reinterpret_cast<uint32_t*>(A.st)[tid] = src[tid];
      if (A.st_host) reinterpret_cast<uint32_t*>(A.st_host)[tid] = src[tid];
    }
    if (tid == 0) *barrier_clear = 0u;
  }
}

}  // namespace ll
