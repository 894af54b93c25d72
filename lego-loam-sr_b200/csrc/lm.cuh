// The Levenberg-Marquardt loop of scan2MapOptimization (mapOptmization.cpp:1572-1610) on the device.
//
// One iteration of the reference's loop body, per block of 512 threads (lm_block_pass):
//   cornerOptimization (a7) + surfOptimization (a8), in two phases —
//     phase 1, one warp per query, three queries of a warp searched together (warp_knn5_batch):
//       pointAssociateToMap (a6) -> 5-NN in the 3x3x3 cells of the 1 m grid (a5; lanes fetch the 27 cell
//       descriptors in one round trip; the candidates of a row of three x-adjacent cells are contiguous in the
//       cell-sorted array, so lanes read consecutive float4s) -> per-lane sorting network + warp-level top-5 by
//       (d2, map index) with redux.sync -> neighbours staged in shared memory;
//     phase 2, one thread per query: line / plane fit in registers (fit.cuh) -> Jacobian row ->
//       shuffle-tree reduction of the 29 sums (21 AtA + 6 Atb + sum|d| + count) in a fixed order;
//   LMOptimization (a9): block partials -> global memory -> added in a fixed order -> 6x6 solve, degeneracy
//     projector, pose update and its sin/cos, stop test (lm_solve_block, solve_warp.cuh).
// Two drivers:
//   lm_persistent_kernel  the WHOLE scan in one cooperative launch (default): optional in-kernel map assembly
//                         (map_inline.cuh), then all iterations, separated by grid barriers; every block adds
//                         all partials and solves redundantly, so no pose is ever published or waited for;
//   lm_iter_kernel        one launch per iteration (stage-level entry points, persistent_lm = 0): the last block
//                         to finish (atomic ticket) reduces and solves; the kernel boundary is the barrier.
// The host never reads anything back between iterations in either.
#pragma once
#include "common.cuh"
#include "fit.cuh"
#include "map_build.cuh"
#include "map_inline.cuh"
#include "map_incremental.cuh"
#include "solve_warp.cuh"

namespace ll {

// ---- optional in-kernel timeline (scratch/build_timeline.sh compiles a copy of the library with -DLL_TIMELINE):
// thread 0 of every block stamps clock64() at the sub-phases of each LM iteration and %globaltimer at the grid
// barrier; block 0 prints the per-iteration summary at the end of the launch (profiles/r02_lm_iteration_timeline.txt).
#ifdef LL_TIMELINE
constexpr int TL_ITERS = 12, TL_EVENTS = 12, TL_BLOCKS = 160;
__device__ long long g_tl[TL_ITERS][TL_EVENTS][TL_BLOCKS];
__device__ unsigned long long g_tl_gt[TL_ITERS][2][TL_BLOCKS];  // %globaltimer: [0] barrier arrive, [1] barrier release
__device__ long long g_tl_warp[TL_ITERS][8][16];                // block TL_WATCH's warps: sub-stamps of phase 1
__device__ int g_tl_it;
__device__ unsigned g_tl_cnt[TL_ITERS][8];  // [0] full searches, [1] heavy (> 128 candidates), [2] set > 32 points, [3] rejected & dense, [4] cached ok, [5] cached attempted
#define LL_TL_COUNT(k) do { if ((threadIdx.x & 31) == 0 && g_tl_it < TL_ITERS) atomicAdd(&g_tl_cnt[g_tl_it][(k)], 1u); } while (0)
#define LL_TL_COUNT_THREAD(k) do { if (g_tl_it < TL_ITERS) atomicAdd(&g_tl_cnt[g_tl_it][(k)], 1u); } while (0)
constexpr int TL_WATCH = 100;
__device__ __forceinline__ unsigned long long tl_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define LL_TL(it, ev)                                                                                        \
  do {                                                                                                       \
    if (threadIdx.x == 0 && (it) < TL_ITERS && blockIdx.x < TL_BLOCKS) g_tl[(it)][(ev)][blockIdx.x] = clock64(); \
  } while (0)
#define LL_TL_GT(it, k)                                                                                          \
  do {                                                                                                           \
    if (threadIdx.x == 0 && (it) < TL_ITERS && blockIdx.x < TL_BLOCKS) g_tl_gt[(it)][(k)][blockIdx.x] = tl_gtime(); \
  } while (0)
#define LL_TL_WARP(ev)                                                                                                         \
  do {                                                                                                                         \
    if ((threadIdx.x & 31) == 0 && blockIdx.x == TL_WATCH && g_tl_it < TL_ITERS) g_tl_warp[g_tl_it][(ev)][threadIdx.x >> 5] = clock64(); \
  } while (0)
#else
#define LL_TL(it, ev) do {} while (0)
#define LL_TL_GT(it, k) do {} while (0)
#define LL_TL_WARP(ev) do {} while (0)
#define LL_TL_COUNT(k) do {} while (0)
#define LL_TL_COUNT_THREAD(k) do {} while (0)
#endif

constexpr int LM_THREADS = 512;
constexpr int LM_WARPS = LM_THREADS / 32;
constexpr int LM_MIN_BLOCKS = 1;         // resident blocks per SM the register budget is sized for
constexpr int LM_NSUM = 29;              // 21 AtA + 6 Atb + sum|d| + count
constexpr int LM_SLOT = 21;              // floats staged per query (odd => bank-conflict-free per-thread reads)
constexpr int LM_CHUNK = LM_THREADS - 32;  // queries a block stages per pass (one warp of slack: corner and surf
                                           // queries of a pass are placed in different warps, see lm_iter_kernel)
constexpr int LM_CAND_UNROLL = 4;        // candidate loads in flight per lane during the 5-NN scan
constexpr uint32_t LM_EMPTY_D2 = 0x7f800000u;  // +inf: "no candidate"

struct LMParams {
  FitParams fit;
  int min_residuals;
  float eig_thresh;
  int iter_count_thres;
  float step_size;
  float stop_thres;
  int apply_update;
  int min_corner_map;
  int min_surf_map;
  int lm_solver;  // 0 = Eigen's ColPivHouseholderQR step sequence (qr66_warp), 1 = LDL^T in double (ldlt6_solve)
};

// Persistent per-handle state; the members of MapOptimization the hot path reads and writes
// (mapOptimization.h:245-253, :279-281) plus loop bookkeeping.
struct LMState {
  float pose[6];  // transformTobeMapped
  Trig trig;      // its cached sin / cos (updatePointAssociateToMapSinCos)
  RotTrig rtrig;  // the six values LMOptimization recomputes (:1445-1450)
  int iter;       // iterCount of the next launch
  int done;       // loop has ended (break, threshold, or nothing can change any more)
  int status;     // 0 optimised, 1 skipped by the map-size guard, 2 search grid too large (scan not registered)
  int iter_num;
  int converged;
  int sel_num;
  int flags;
  int lm_ret;     // return value of the last LMOptimization: 1 true, 0 false, -1 "< 50 residuals"
  int is_degenerate;
  float matP[36];  // row-major
  float min_lambda;
  float cf_mean;
  float last_x[6];
  float AtA[36];   // row-major, of the last executed iteration
  float Atb[6];
  float deltaR, deltaT;
  unsigned int ticket;
  int corner_map_ds, surf_map_ds;
};

struct LMArgs {
  const float4* sorted[2];
  const int2* cells[2];
  const MapState* mst[2];
  const float4* q[2];  // corner queries, surf queries (sensor frame)
  int nq[2];
  int q_step;  // distance between two queries in float4 units: 1 = packed {x, y, z, intensity}; 2 = pcl::PointXYZI records
               // uploaded as they are (32 bytes, {x, y, z, pad} first: the path only reads x, y, z of a query)
  LMParams prm;
  LMState* st;
  LMState* st_host;  // optional pinned mirror of *st, written by the persistent kernel (saves the copy behind it)
  float* partials;  // [gridDim.x][32]
  int mode;         // 0 = LM iteration; 1 = residuals only (emit); 2 = k-NN only (emit, queries in map frame)
  int w_corner, w_surf;  // relative cost of a corner / surf query (lm_partition)
  // emit targets (mode != 0), indexed by query (corner queries first)
  float* e_coeff;
  int* e_mask;
  int* e_accept;
  int* e_idx;
  float* e_d2;
  // cache records of the queries beyond a block's LM_QCAP shared-memory slots (dense scans: hundreds of queries per
  // block), indexed by query (corner queries first), LM_GUARD_STRIDE float4 each: the same fields as a KnnCache slot —
  // [0..31] cand, [32] center, [33] qpt, [34] t1, [35] = {cnt, prev, -, -}, [36..37] geo. nullptr: not used.
  float4* g_guard;
  // The query clouds may still be on their way when the launch starts (ll_scan2map_async with host clouds: the upload runs
  // on the copy stream while this kernel assembles the map, which does not need them). q_sync[0] is written by a 4-byte
  // copy queued BEHIND the clouds on that stream (value q_seq); block 0 waits for it before the LM loop and tells the other
  // blocks through q_sync[1] and a grid barrier whether it arrived (a copy that never comes must not hang the GPU).
  // nullptr: the queries were resident before the launch.
  unsigned* q_sync;
  unsigned q_seq;
};
constexpr int LM_GUARD_STRIDE = 38;

// Address of query qi of the concatenated list (corner queries first).
__device__ __forceinline__ const float4* lm_query(const LMArgs& A, int qi) {
  const bool surf = qi >= A.nq[0];
  return (surf ? A.q[1] : A.q[0]) + (size_t)(surf ? qi - A.nq[0] : qi) * (size_t)A.q_step;
}

// ---- 6x6 dense helpers (one thread; column-major like Eigen) -------------------------------------------
struct Mat6 {
  float d[36];
  __device__ float& operator()(int r, int c) { return d[c * 6 + r]; }
  __device__ float operator()(int r, int c) const { return d[c * 6 + r]; }
};

// Eigen 3.3.7 Householder.h makeHouseholderInPlace on v[0..n)
__device__ inline void make_householder(float* v, int n, float& tau, float& beta) {
  float tailSq = 0.f;
  for (int i = 1; i < n; ++i) tailSq += v[i] * v[i];
  const float c0 = v[0];
  if (n == 1 || tailSq <= FLT_MIN) {
    tau = 0.f;
    beta = c0;
    for (int i = 1; i < n; ++i) v[i] = 0.f;
  } else {
    beta = sqrtf(c0 * c0 + tailSq);
    if (c0 >= 0.f) beta = -beta;
    const float denom = c0 - beta;
    for (int i = 1; i < n; ++i) v[i] = v[i] / denom;
    tau = (beta - c0) / beta;
  }
}

// matAtA.colPivHouseholderQr().solve(matAtB) (mapOptmization.cpp:1505) for an n x n system, n <= 6.
// Same step sequence as Eigen 3.3.7 ColPivHouseholderQR::computeInPlace + _solve_impl.
// kN > 0: the size is a compile-time constant (everything unrolls into registers: the 3 x 3 solves of fa_loop.cuh);
// kN = 0: the size is the run-time argument n.
template <int kN>
__device__ __forceinline__ void colpiv_qr_solve_core(const float* Ain /* col-major n x n */, const float* b, float* x, int n_rt) {
  const int n = kN > 0 ? kN : n_rt;
  float qr[36], hC[6], nU[6], nD[6], c[6], col[6];
  int transp[6], perm[6];
  for (int i = 0; i < n * n; ++i) qr[i] = Ain[i];
#define QR(r, cc) qr[(cc) * n + (r)]
  for (int k = 0; k < n; ++k) {
    float s = 0.f;
    for (int r = 0; r < n; ++r) s += QR(r, k) * QR(r, k);
    nD[k] = sqrtf(s);
    nU[k] = nD[k];
  }
  float maxNorm = nU[0];
  for (int k = 1; k < n; ++k) maxNorm = fmaxf(maxNorm, nU[k]);
  const float thr = (maxNorm * FLT_EPSILON) * (maxNorm * FLT_EPSILON) / (float)n;
  const float ndt = sqrtf(FLT_EPSILON);
  int nonzero = n;
  for (int k = 0; k < n; ++k) {
    int big = k;
    float bigv = nU[k];
    for (int j = k + 1; j < n; ++j)
      if (nU[j] > bigv) { bigv = nU[j]; big = j; }
    if (nonzero == n && bigv * bigv < thr * (float)(n - k)) nonzero = k;
    transp[k] = big;
    if (k != big) {
      for (int r = 0; r < n; ++r) { float t = QR(r, k); QR(r, k) = QR(r, big); QR(r, big) = t; }
      float t = nU[k]; nU[k] = nU[big]; nU[big] = t;
      t = nD[k]; nD[k] = nD[big]; nD[big] = t;
    }
    const int m = n - k;
    for (int r = 0; r < m; ++r) col[r] = QR(k + r, k);
    float tau, beta;
    make_householder(col, m, tau, beta);
    for (int r = 1; r < m; ++r) QR(k + r, k) = col[r];
    QR(k, k) = beta;
    hC[k] = tau;
    if (m == 1) {
      for (int cc = k + 1; cc < n; ++cc) QR(k, cc) *= (1.f - tau);
    } else if (tau != 0.f) {
      for (int cc = k + 1; cc < n; ++cc) {
        float tmp = 0.f;
        for (int r = 1; r < m; ++r) tmp += col[r] * QR(k + r, cc);
        tmp += QR(k, cc);
        QR(k, cc) -= tau * tmp;
        for (int r = 1; r < m; ++r) QR(k + r, cc) -= tau * col[r] * tmp;
      }
    }
    for (int j = k + 1; j < n; ++j) {
      if (nU[j] != 0.f) {
        float temp = fabsf(QR(k, j)) / nU[j];
        temp = (1.f + temp) * (1.f - temp);
        temp = temp < 0.f ? 0.f : temp;
        const float ratio = nU[j] / nD[j];
        const float temp2 = temp * (ratio * ratio);
        if (temp2 <= ndt) {
          float s = 0.f;
          for (int r = k + 1; r < n; ++r) s += QR(r, j) * QR(r, j);
          nD[j] = sqrtf(s);
          nU[j] = nD[j];
        } else {
          nU[j] *= sqrtf(temp);
        }
      }
    }
  }
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int k = 0; k < n; ++k) { int t = perm[k]; perm[k] = perm[transp[k]]; perm[transp[k]] = t; }
  for (int i = 0; i < n; ++i) x[i] = 0.f;
  if (nonzero == 0) return;
  for (int r = 0; r < n; ++r) c[r] = b[r];
  for (int k = 0; k < nonzero; ++k) {
    const int m = n - k;
    const float tau = hC[k];
    if (m == 1) {
      c[k] *= (1.f - tau);
    } else if (tau != 0.f) {
      float tmp = 0.f;
      for (int r = 1; r < m; ++r) tmp += QR(k + r, k) * c[k + r];
      tmp += c[k];
      c[k] -= tau * tmp;
      for (int r = 1; r < m; ++r) c[k + r] -= tau * QR(k + r, k) * tmp;
    }
  }
  for (int i = nonzero - 1; i >= 0; --i) {
    float s = c[i];
    for (int j = i + 1; j < nonzero; ++j) s -= QR(i, j) * c[j];
    c[i] = s / QR(i, i);
  }
  for (int i = 0; i < nonzero; ++i) x[perm[i]] = c[i];
#undef QR
}
__device__ __noinline__ void colpiv_qr_solve_n(const float* Ain /* col-major n x n */, const float* b, float* x, int n) {
  colpiv_qr_solve_core<0>(Ain, b, x, n);
}
__device__ __noinline__ void colpiv_qr_solve_3(const float* Ain /* col-major 3 x 3 */, const float* b, float* x) {
  colpiv_qr_solve_core<3>(Ain, b, x, 3);
}

// Cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (n <= 6) in double: eigenvalues
// ascending in w, eigenvectors in the columns of V (col-major n x n). Used for eigen_mode = upstream,
// where only eigenvalues and the span of the kept eigenvectors matter (sign-free projector).
__device__ __noinline__ void jacobi_eig_n(const float* Ain /* col-major, symmetric */, int n, double* w, double* V) {
  double A[36];
  for (int i = 0; i < n * n; ++i) A[i] = (double)Ain[i];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[j * n + i] = (i == j) ? 1.0 : 0.0;
#define AA(r, c) A[(c) * n + (r)]
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) {
      diag += AA(i, i) * AA(i, i);
      for (int j = i + 1; j < n; ++j) off += AA(i, j) * AA(i, j);
    }
    if (off <= 1e-30 * diag || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = AA(p, q);
        if (apq == 0.0) continue;
        const double theta = (AA(q, q) - AA(p, p)) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double cth = 1.0 / sqrt(t * t + 1.0), sth = t * cth;
        for (int k = 0; k < n; ++k) {  // A <- A J
          const double akp = AA(k, p), akq = AA(k, q);
          AA(k, p) = cth * akp - sth * akq;
          AA(k, q) = sth * akp + cth * akq;
        }
        for (int k = 0; k < n; ++k) {  // A <- J^T A
          const double apk = AA(p, k), aqk = AA(q, k);
          AA(p, k) = cth * apk - sth * aqk;
          AA(q, k) = sth * apk + cth * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[p * n + k], vkq = V[q * n + k];
          V[p * n + k] = cth * vkp - sth * vkq;
          V[q * n + k] = sth * vkp + cth * vkq;
        }
      }
  }
  for (int i = 0; i < n; ++i) w[i] = AA(i, i);
#undef AA
  for (int i = 0; i < n - 1; ++i) {  // ascending selection sort, swapping columns
    int k = i;
    for (int j = i + 1; j < n; ++j)
      if (w[j] < w[k]) k = j;
    if (k != i) {
      double t = w[i]; w[i] = w[k]; w[k] = t;
      for (int r = 0; r < n; ++r) { double u = V[i * n + r]; V[i * n + r] = V[k * n + r]; V[k * n + r] = u; }
    }
  }
}

// Eigen-style SelfAdjointEigenSolver<Matrix<float,6,6>> (general Householder tridiagonalisation +
// the shared QR iteration of fit.cuh) and PartialPivLU inverse, for eigen_mode = as shipped
// (mapOptmization.cpp:1512-1530 depends on Eigen's eigenvector signs). Q col-major.
__device__ __noinline__ void eig6_eigen_style(const float* Ain /* col-major symmetric */, float ev[6], float Q[36]) {
  const int N = 6;
  float A[36];
#define AA(r, c) A[(c) * N + (r)]
  float scale = 0.f;
  for (int i = 0; i < 36; ++i) A[i] = 0.f;
  for (int c = 0; c < N; ++c)
    for (int r = c; r < N; ++r) {
      AA(r, c) = Ain[c * N + r];
      scale = fmaxf(scale, fabsf(Ain[c * N + r]));
    }
  if (scale == 0.f) scale = 1.f;
  for (int c = 0; c < N; ++c)
    for (int r = c; r < N; ++r) AA(r, c) /= scale;
  float hC[5], p[6], col[6], diag[6], sub[5];
  for (int i = 0; i < N - 1; ++i) {
    const int rs = N - i - 1;
    for (int r = 0; r < rs; ++r) col[r] = AA(i + 1 + r, i);
    float h, beta;
    make_householder(col, rs, h, beta);
    col[0] = 1.f;
    for (int r = 0; r < rs; ++r) {
      float acc = 0.f;
      for (int c = 0; c < rs; ++c) {
        const float a = (r >= c) ? AA(i + 1 + r, i + 1 + c) : AA(i + 1 + c, i + 1 + r);
        acc += a * (h * col[c]);
      }
      p[r] = acc;
    }
    float dot = 0.f;
    for (int r = 0; r < rs; ++r) dot += p[r] * col[r];
    const float alpha = h * -0.5f * dot;
    for (int r = 0; r < rs; ++r) p[r] += alpha * col[r];
    for (int c = 0; c < rs; ++c)
      for (int r = c; r < rs; ++r) AA(i + 1 + r, i + 1 + c) -= (col[r] * p[c] + p[r] * col[c]);
    AA(i + 1, i) = beta;
    for (int r = 1; r < rs; ++r) AA(i + 1 + r, i) = col[r];
    hC[i] = h;
  }
  for (int i = 0; i < N; ++i) diag[i] = AA(i, i);
  for (int i = 0; i < N - 1; ++i) sub[i] = AA(i + 1, i);
  for (int i = 0; i < 36; ++i) Q[i] = 0.f;
  for (int i = 0; i < N; ++i) Q[i * N + i] = 1.f;
#define QQ(r, c) Q[(c) * N + (r)]
  for (int k = N - 2; k >= 0; --k) {
    const int cs = N - k - 1;
    const int off = N - cs;
    const float tau = hC[k];
    if (cs == 1) {
      QQ(off, off) *= (1.f - tau);
    } else if (tau != 0.f) {
      for (int c = 0; c < cs; ++c) {
        float tmp = 0.f;
        for (int r = 1; r < cs; ++r) tmp += AA(k + 1 + r, k) * QQ(off + r, off + c);
        tmp += QQ(off, off + c);
        QQ(off, off + c) -= tau * tmp;
        for (int r = 1; r < cs; ++r) QQ(off + r, off + c) -= tau * AA(k + 1 + r, k) * tmp;
      }
    }
  }
#undef QQ
#undef AA
  tridiagonal_eig<6>(diag, sub, Q);
  for (int i = 0; i < N; ++i) ev[i] = diag[i] * scale;
}

__device__ __noinline__ void inverse_lu6(const float* Ain /* col-major */, float* inv /* col-major */) {
  const int N = 6;
  float lu[36];
  int piv[6];
  for (int i = 0; i < 36; ++i) lu[i] = Ain[i];
#define LU(r, c) lu[(c) * N + (r)]
  for (int k = 0; k < N; ++k) {
    int pp = k;
    float best = fabsf(LU(k, k));
    for (int r = k + 1; r < N; ++r)
      if (fabsf(LU(r, k)) > best) { best = fabsf(LU(r, k)); pp = r; }
    piv[k] = pp;
    if (pp != k)
      for (int c = 0; c < N; ++c) { float t = LU(k, c); LU(k, c) = LU(pp, c); LU(pp, c) = t; }
    if (LU(k, k) != 0.f) {
      for (int r = k + 1; r < N; ++r) LU(r, k) /= LU(k, k);
      for (int c = k + 1; c < N; ++c)
        for (int r = k + 1; r < N; ++r) LU(r, c) -= LU(r, k) * LU(k, c);
    }
  }
  for (int cc = 0; cc < N; ++cc) {
    float y[6];
    for (int r = 0; r < N; ++r) y[r] = (r == cc) ? 1.f : 0.f;
    for (int k = 0; k < N; ++k) { float t = y[k]; y[k] = y[piv[k]]; y[piv[k]] = t; }
    for (int r = 0; r < N; ++r)
      for (int c = 0; c < r; ++c) y[r] -= LU(r, c) * y[c];
    for (int r = N - 1; r >= 0; --r) {
      for (int c = r + 1; c < N; ++c) y[r] -= LU(r, c) * y[c];
      y[r] /= LU(r, r);
    }
    for (int r = 0; r < N; ++r) inv[cc * N + r] = y[r];
  }
#undef LU
}

__device__ inline void set_trig(LMState* S) {
  // cos / sin of a float evaluated in double and rounded (the repo-wide convention)
  S->trig.cr = (float)cos((double)S->pose[0]);
  S->trig.sr = (float)sin((double)S->pose[0]);
  S->trig.cp = (float)cos((double)S->pose[1]);
  S->trig.sp = (float)sin((double)S->pose[1]);
  S->trig.cy = (float)cos((double)S->pose[2]);
  S->trig.sy = (float)sin((double)S->pose[2]);
  S->trig.tx = S->pose[3];
  S->trig.ty = S->pose[4];
  S->trig.tz = S->pose[5];
  S->rtrig.srx = S->trig.sr; S->rtrig.crx = S->trig.cr;
  S->rtrig.sry = S->trig.sp; S->rtrig.cry = S->trig.cp;
  S->rtrig.srz = S->trig.sy; S->rtrig.crz = S->trig.cy;
}

// The part of LMOptimization after matX is known (mapOptmization.cpp:1533-1569) plus the loop control of
// scan2MapOptimization (:1578-1586). One thread. x == nullptr means "fewer than min_residuals rows": the
// reference returns false before computing anything (:1453).
__device__ __noinline__ void lm_apply_and_control(float* x, int n, double sum_abs_d, const float* AtA_rm,
                                                  const float* Atb, const LMParams& prm, LMState* S, bool stage_mode) {
  const int iter = S->iter;
  S->sel_num = n;
  bool pose_changed = false;
  int ret = -1;
  if (x) {
    if (S->is_degenerate) {  // :1533-1537
      float x2[6];
      for (int r = 0; r < 6; ++r) {
        float acc = 0.f;
        for (int c = 0; c < 6; ++c) acc += S->matP[r * 6 + c] * x[c];
        x2[r] = acc;
      }
      for (int r = 0; r < 6; ++r) x[r] = x2[r];
    }
    if (prm.apply_update) {  // :1540-1545
      for (int r = 0; r < 6; ++r) S->pose[r] += x[r];
      pose_changed = true;
    }
    const double r0 = (double)(x[0] * 57.29578f), r1 = (double)(x[1] * 57.29578f), r2 = (double)(x[2] * 57.29578f);
    const float deltaR = (float)sqrt(r0 * r0 + r1 * r1 + r2 * r2);
    const double t0 = (double)(x[3] * 100.f), t1 = (double)(x[4] * 100.f), t2 = (double)(x[5] * 100.f);
    const float deltaT = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2);
    S->cf_mean = (float)sum_abs_d / (float)n;  // :1555-1568
    for (int r = 0; r < 6; ++r) S->last_x[r] = x[r];
    for (int i = 0; i < 36; ++i) S->AtA[i] = AtA_rm[i];
    for (int r = 0; r < 6; ++r) S->Atb[r] = Atb[r];
    S->deltaR = deltaR;
    S->deltaT = deltaT;
    ret = (deltaR < prm.stop_thres && deltaT < prm.stop_thres) ? 1 : 0;
  }
  S->lm_ret = ret;
  if (stage_mode) {
    if (pose_changed) set_trig(S);
    return;
  }
  const int executed = iter + 1;
  if (ret == 1) {
    S->converged = 1;
    S->iter_num = executed;
    S->done = 1;
  } else if (executed >= prm.iter_count_thres) {
    S->iter_num = executed;
    S->done = 1;
  } else if (!pose_changed) {
    // Nothing can change any more: the pose is not updated (as shipped, or fewer than 50 residuals), so
    // every remaining iteration would recompute exactly the same numbers and fall through to the
    // threshold. Report what the reference reports without replaying them.
    S->iter_num = prm.iter_count_thres;
    S->done = 1;
  }
  S->iter = executed;
  if (pose_changed) set_trig(S);
}

// LMOptimization after the rows are reduced (mapOptmization.cpp:1452-1455, :1502-1569), one thread.
// sums: 21 upper-triangular AtA entries (row-major order), 6 Atb, sum|d|, count.
__device__ __noinline__ void lm_solve(const double* sums, const LMParams& prm, LMState* S, bool stage_mode) {
  const int n = (int)(sums[28] + 0.5);
  const int iter = S->iter;
  if (n >= prm.min_residuals) {
    Mat6 AtA;
    float Atb[6], x[6];
    int k = 0;
    for (int r = 0; r < 6; ++r)
      for (int c = r; c < 6; ++c) {
        const float v = (float)sums[k++];
        AtA(r, c) = v;
        AtA(c, r) = v;
      }
    for (int r = 0; r < 6; ++r) Atb[r] = (float)sums[21 + r];
    colpiv_qr_solve_n(AtA.d, Atb, x, 6);
    if (iter == 0) {  // :1507-1531
      float P[36];    // row-major
      int degenerate = 0;
      if (prm.fit.eigen_mode == 0) {
        float E[6], V[36], V2[36], Vinv[36];
        eig6_eigen_style(AtA.d, E, V);
        S->min_lambda = E[0];
        for (int i = 0; i < 36; ++i) V2[i] = V[i];
        for (int i = 5; i >= 0; --i) {  // as shipped: ascending eigenvalues walked from the largest
          if (E[i] < prm.eig_thresh) {
            for (int j = 0; j < 6; ++j) V2[j * 6 + i] = 0.f;  // zero ROW i of the column matrix
            degenerate = 1;
          } else {
            break;
          }
        }
        inverse_lu6(V, Vinv);
        for (int r = 0; r < 6; ++r)
          for (int c = 0; c < 6; ++c) {
            float acc = 0.f;
            for (int kk = 0; kk < 6; ++kk) acc += Vinv[kk * 6 + r] * V2[c * 6 + kk];
            P[r * 6 + c] = acc;
          }
      } else {
        double w[6], V[36];
        jacobi_eig_n(AtA.d, 6, w, V);
        S->min_lambda = (float)w[0];
        int first_kept = 0;
        while (first_kept < 6 && (float)w[first_kept] < prm.eig_thresh) ++first_kept;
        degenerate = first_kept > 0;
        for (int r = 0; r < 6; ++r)
          for (int c = 0; c < 6; ++c) {
            double acc = 0.0;
            for (int kk = first_kept; kk < 6; ++kk) acc += V[kk * 6 + r] * V[kk * 6 + c];
            P[r * 6 + c] = (float)acc;
          }
      }
      S->is_degenerate = degenerate;
      for (int i = 0; i < 36; ++i) S->matP[i] = P[i];
    }
    float AtA_rm[36];
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) AtA_rm[r * 6 + c] = AtA(r, c);
    lm_apply_and_control(x, n, sums[27], AtA_rm, Atb, prm, S, stage_mode);
    return;
  }
  lm_apply_and_control(nullptr, n, sums[27], nullptr, nullptr, prm, S, stage_mode);
}

// ---- warp-level 5-NN --------------------------------------------------------------------------------------
struct GridGeom {
  int ox, oy, oz, dx, dy, dz;
  int tag;  // stamp of the valid cell-table entries (map_build.cuh: cell_entry_x)
};

struct Knn5 {
  uint32_t d2[5];   // float bit patterns, ascending; LM_EMPTY_D2 = missing
  uint32_t idx[5];  // map (ds) index
  uint32_t pos[5];  // position in the cell-sorted array
  bool accept;
};

// Loads of the search structures (`sorted`, `cells`). kFresh = the arrays were written EARLIER IN THIS LAUNCH (the
// single-kernel scan builds the map and its grids before the LM loop): the read-only path (ld.global.nc) is defined
// only for data that is constant for the whole kernel, so those launches use plain weak loads, which the grid
// barrier's release / acquire pair orders after the writes (PTX memory model) and which still allocate in L1.
template <bool kFresh, typename T>
__device__ __forceinline__ T map_ld(const T* p) {
  if constexpr (kFresh) return *p;
  else return __ldg(p);
}

__device__ __forceinline__ bool pair_less(uint32_t da, uint32_t ia, uint32_t db, uint32_t ib) {
  return da < db || (da == db && ia < ib);
}

// All 32 lanes of a warp cooperate on ONE query (sx, sy, sz) already in the map frame.
//  1. lanes 0..26 each fetch one of the 27 cell descriptors (a single load round trip);
//  2. the three x-adjacent cells of a row are contiguous in the cell-sorted array, so two shuffles merge
//     them into 9 [start, count) ranges; they are broadcast centre row first (the row of the query's own
//     cell holds the nearest points, so the running 5th-best distance is tight from the first chunk on);
//  3. candidates are taken 128 at a time: lane l owns candidates l, l+32, l+64, l+96 of the concatenated
//     ranges (consecutive float4s; all four loads in flight before any is consumed), sorts its four
//     (d2, index) pairs with a 5-comparator network, and five rounds of redux.sync pop the warp-wide
//     minimum by (d2, map index). A typical VLP-16 query has 50-110 candidates: one chunk.
//  4. with more than 128 candidates the best five so far are carried in uniform registers, later
//     candidates are pre-filtered against the carried 5th best, and chunks without survivors are skipped.
// The result is the exact five smallest (d2, index) pairs with d2 < max_sq_dist, independent of the
// order in which candidates are visited.
#define LL_KNN_CE(a, b)                                                  \
  {                                                                      \
    const bool sw_ = pair_less(cd[b], ci[b], cd[a], ci[a]);              \
    const uint32_t d_ = sw_ ? cd[b] : cd[a], i_ = sw_ ? ci[b] : ci[a], p_ = sw_ ? cp[b] : cp[a]; \
    cd[b] = sw_ ? cd[a] : cd[b]; ci[b] = sw_ ? ci[a] : ci[b]; cp[b] = sw_ ? cp[a] : cp[b];        \
    cd[a] = d_; ci[a] = i_; cp[a] = p_;                                  \
  }

// Chunks j0_begin, j0_begin + 128, ... of the candidate list described by (my_start, my_count) on lanes 0, 3, ..., 24,
// merged into K (the best five of the chunks before j0_begin; all empty when j0_begin == 0).
template <bool kFresh>
__device__ __forceinline__ void knn5_chunks(const float4* sorted, int my_start, int my_count, float sx,
                                            float sy, float sz, float max_sq_dist, int lane, int j0_begin, Knn5& K) {
  // pr[k] = candidates before range k, bs[k] = (start of range k) - pr[k]: candidate j of the concatenation
  // lives at sorted[j + bs[k]] for the last k with pr[k] <= j. Range order: centre row first.
  constexpr int kRowLane[9] = {12, 3, 9, 15, 21, 0, 6, 18, 24};
  int pr[9], bs[9];
  int total = 0;
  // Row pruning (dense maps: hundreds of candidates in the 27 cells, the five nearest within centimetres): once the chunks
  // before j0_begin have left a fifth-best distance, a row of cells whose nearest possible point — by its (y, z) cell
  // offsets and the query's position inside its own cell — is farther than that cannot contribute, and its candidates
  // are never loaded. The margin keeps rounding (3 ulp on a float-evaluated d2) and exact ties out of the argument.
  const float fy = sy - floorf(sy), fz = sz - floorf(sz);
  const bool prune = j0_begin > 0 && K.d2[4] != LM_EMPTY_D2;
  const float bound = __uint_as_float(K.d2[4]) * 1.00001f + 1e-12f;
  int consumed = j0_begin;  // candidates of the un-pruned concatenation the earlier chunks have taken
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    int st = __shfl_sync(0xffffffffu, my_start, kRowLane[k]);
    int cn = __shfl_sync(0xffffffffu, my_count, kRowLane[k]);
    const int used = min(consumed, cn);  // this row's share of what is already done (rows are taken in this order)
    consumed -= used;
    st += used;
    cn -= used;
    if (prune) {
      const int oy = (kRowLane[k] / 3) % 3 - 1, oz = kRowLane[k] / 9 - 1;
      const float dy = oy == 0 ? 0.f : (oy < 0 ? fy : 1.f - fy), dz = oz == 0 ? 0.f : (oz < 0 ? fz : 1.f - fz);
      if (dy * dy + dz * dz > bound) cn = 0;
    }
    pr[k] = total;
    bs[k] = st - total;
    total += cn;
  }
  const uint32_t max_bits = __float_as_uint(max_sq_dist);
  for (int j0 = 0; j0 < total; j0 += 32 * LM_CAND_UNROLL) {
    const bool first = j0 == 0 && j0_begin == 0;  // nothing carried yet
    int srcs[LM_CAND_UNROLL];
    float4 ms[LM_CAND_UNROLL];
#pragma unroll
    for (int u = 0; u < LM_CAND_UNROLL; ++u) {
      srcs[u] = -1;
      if (j0 + 32 * u < total) {  // warp-uniform
        const int j = j0 + 32 * u + lane;
        int b = bs[0];
#pragma unroll
        for (int k = 1; k < 9; ++k)
          if (j >= pr[k]) b = bs[k];
        if (j < total) srcs[u] = j + b;
      }
    }
#pragma unroll
    for (int u = 0; u < LM_CAND_UNROLL; ++u)
      ms[u] = srcs[u] >= 0 ? map_ld<kFresh>(&sorted[srcs[u]]) : make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t cd[LM_CAND_UNROLL], ci[LM_CAND_UNROLL], cp[LM_CAND_UNROLL];
    bool any = false;
#pragma unroll
    for (int u = 0; u < LM_CAND_UNROLL; ++u) {
      const float4 m = ms[u];
      // L2_Simple_Adaptor::evalMetric (nanoflann.hpp:431-439): ((dx^2) + dy^2) + dz^2, no contraction
      float df = sx - m.x;
      float d = df * df;
      df = sy - m.y;
      d += df * df;
      df = sz - m.z;
      d += df * df;
      const uint32_t db = __float_as_uint(d);
      const uint32_t ib = __float_as_uint(m.w);
      const bool ok = srcs[u] >= 0 && db < max_bits && (first || pair_less(db, ib, K.d2[4], K.idx[4]));
      cd[u] = ok ? db : LM_EMPTY_D2;
      ci[u] = ok ? ib : 0xffffffffu;
      cp[u] = ok ? (uint32_t)srcs[u] : 0xffffffffu;
      any = any || ok;
    }
    if (!first && !__any_sync(0xffffffffu, any)) continue;  // nothing in this chunk beats the carried five
    static_assert(LM_CAND_UNROLL == 4, "the sorting network below is written for four candidates per lane");
    LL_KNN_CE(0, 1) LL_KNN_CE(2, 3) LL_KNN_CE(0, 2) LL_KNN_CE(1, 3) LL_KNN_CE(1, 2)
    // five rounds: warp-wide arg-min over the list heads (ties broken by the lower map index), merged with the
    // carried list
    Knn5 out;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      const uint32_t m = __reduce_min_sync(0xffffffffu, cd[0]);
      const uint32_t om = __reduce_min_sync(0xffffffffu, cd[0] == m ? ci[0] : 0xffffffffu);
      const bool carried = !first && pair_less(K.d2[0], K.idx[0], m, om);  // warp-uniform
      if (carried) {
        out.d2[r] = K.d2[0]; out.idx[r] = K.idx[0]; out.pos[r] = K.pos[0];
#pragma unroll
        for (int k = 0; k < 4; ++k) { K.d2[k] = K.d2[k + 1]; K.idx[k] = K.idx[k + 1]; K.pos[k] = K.pos[k + 1]; }
        K.d2[4] = LM_EMPTY_D2; K.idx[4] = 0xffffffffu; K.pos[4] = 0xffffffffu;
      } else {
        const bool win = (cd[0] == m) && (ci[0] == om) && (m != LM_EMPTY_D2);
        const uint32_t pw = __reduce_min_sync(0xffffffffu, win ? cp[0] : 0xffffffffu);
        out.d2[r] = m; out.idx[r] = om; out.pos[r] = pw;
        if (win) {
#pragma unroll
          for (int k = 0; k < LM_CAND_UNROLL - 1; ++k) { cd[k] = cd[k + 1]; ci[k] = ci[k + 1]; cp[k] = cp[k + 1]; }
          cd[LM_CAND_UNROLL - 1] = LM_EMPTY_D2; ci[LM_CAND_UNROLL - 1] = 0xffffffffu; cp[LM_CAND_UNROLL - 1] = 0xffffffffu;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) { K.d2[k] = out.d2[k]; K.idx[k] = out.idx[k]; K.pos[k] = out.pos[k]; }
  }
}


// LM_KNN_BATCH queries of one warp searched TOGETHER: the cell descriptors of all of them are fetched in one
// round trip, then all their candidates, and the five selection rounds of the different queries are independent
// dependency chains that the scheduler interleaves — a block's ~47 queries over 16 warps would otherwise be three
// fully serialised search latencies per warp. Handles the first chunk (<= 128 candidates) of every query; a
// query with more candidates is searched by warp_knn5 (same result: the exact five smallest (d2, index) pairs).
// On return lane r < 5 holds neighbour r of query q in (od, oi, op)[q] (bit patterns as in Knn5; empty slots
// LM_EMPTY_D2 / 0xffffffff), and acc[q] (warp-uniform) says whether five neighbours within range exist.
constexpr int LM_KNN_BATCH = 3;

template <bool kFresh, int NB>
__device__ __forceinline__ void warp_knn5_batch(const float4* const (&sorted)[2], const int2* const (&cells)[2],
                                                const GridGeom (&geom)[2], const int (&mq)[NB],
                                                const float4 (&sel)[NB], float max_sq_dist, int lane,
                                                uint32_t (&od)[NB], uint32_t (&oi)[NB],
                                                uint32_t (&op)[NB], bool (&acc)[NB],
                                                int (&q_start)[NB], int (&q_count)[NB],
                                                int (&total)[NB]) {
  // q_start / q_count (per lane: the merged cell row this lane describes, valid on lanes 0, 3, ..., 24) and total (the
  // number of candidates in the 27 cells, warp-uniform) are returned for knn_cache_fill.
  // NB = LM_KNN_BATCH in the first rounds of a warp; a single remaining query (the fourth of a warp's share in iterations
  // 0 and 1, the odd query whose cached set failed later on) takes the NB = 1 instance: the search is instruction-fetch
  // bound, and a third of the unrolled code is a third of the time.
  constexpr int kRowLane[9] = {12, 3, 9, 15, 21, 0, 6, 18, 24};
  const uint32_t max_bits = __float_as_uint(max_sq_dist);
  // ---- stage 1: the 27 cell descriptors of every query (one load per lane per query, all in flight together)
  int2 ce[NB];
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    ce[q] = make_int2(0, 0);
    if (mq[q] >= 0 && lane < 27) {
      const GridGeom& g = mq[q] ? geom[1] : geom[0];
      const int x = (int)floorf(sel[q].x) - g.ox + (lane % 3) - 1, y = (int)floorf(sel[q].y) - g.oy + ((lane / 3) % 3) - 1,
                z = (int)floorf(sel[q].z) - g.oz + (lane / 9) - 1;
      if (x >= 0 && x < g.dx && y >= 0 && y < g.dy && z >= 0 && z < g.dz)
        ce[q] = cell_decode(map_ld<kFresh>(&cells[mq[q] ? 1 : 0][(z * g.dy + y) * g.dx + x]), g.tag);
    }
  }
  // ---- stage 2: ranges -> candidate addresses -> candidate loads (all queries' loads in flight together)
  int srcs[NB][LM_CAND_UNROLL];
  float4 ms[NB][LM_CAND_UNROLL];
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    int s = (ce[q].y > ce[q].x) ? ce[q].x : 0x7fffffff;
    int e = (ce[q].y > ce[q].x) ? ce[q].y : 0;
    const int s1 = __shfl_down_sync(0xffffffffu, s, 1), e1 = __shfl_down_sync(0xffffffffu, e, 1);
    const int s2 = __shfl_down_sync(0xffffffffu, s, 2), e2 = __shfl_down_sync(0xffffffffu, e, 2);
    s = min(s, min(s1, s2));
    e = max(e, max(e1, e2));
    const int my_start = e > 0 ? s : 0, my_count = e > 0 ? e - s : 0;
    q_start[q] = my_start;
    q_count[q] = my_count;
    int pr[9], bs[9];
    int tot = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int st = __shfl_sync(0xffffffffu, my_start, kRowLane[k]);
      const int cn = __shfl_sync(0xffffffffu, my_count, kRowLane[k]);
      pr[k] = tot;
      bs[k] = st - tot;
      tot += cn;
    }
    total[q] = tot;
#pragma unroll
    for (int u = 0; u < LM_CAND_UNROLL; ++u) {
      srcs[q][u] = -1;
      if (32 * u < tot) {  // warp-uniform
        const int j = 32 * u + lane;
        int b = bs[0];
#pragma unroll
        for (int k = 1; k < 9; ++k)
          if (j >= pr[k]) b = bs[k];
        if (j < tot) srcs[q][u] = j + b;
      }
      ms[q][u] = srcs[q][u] >= 0 ? map_ld<kFresh>(&sorted[mq[q] > 0 ? 1 : 0][srcs[q][u]]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // ---- stage 3: distances and the per-lane sorting network
  uint32_t cdq[NB][LM_CAND_UNROLL], ciq[NB][LM_CAND_UNROLL], cpq[NB][LM_CAND_UNROLL];
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    uint32_t (&cd)[LM_CAND_UNROLL] = cdq[q];
    uint32_t (&ci)[LM_CAND_UNROLL] = ciq[q];
    uint32_t (&cp)[LM_CAND_UNROLL] = cpq[q];
#pragma unroll
    for (int u = 0; u < LM_CAND_UNROLL; ++u) {
      const float4 m = ms[q][u];
      float df = sel[q].x - m.x;  // evalMetric order (nanoflann.hpp:431-439), no contraction
      float d = df * df;
      df = sel[q].y - m.y;
      d += df * df;
      df = sel[q].z - m.z;
      d += df * df;
      const uint32_t db = __float_as_uint(d);
      const bool ok = srcs[q][u] >= 0 && db < max_bits;
      cd[u] = ok ? db : LM_EMPTY_D2;
      ci[u] = ok ? __float_as_uint(m.w) : 0xffffffffu;
      cp[u] = ok ? (uint32_t)srcs[q][u] : 0xffffffffu;
    }
    LL_KNN_CE(0, 1) LL_KNN_CE(2, 3) LL_KNN_CE(0, 2) LL_KNN_CE(1, 3) LL_KNN_CE(1, 2)
  }
  // ---- stage 4: five selection rounds, the queries' chains interleaved. One redux.sync finds the smallest d2;
  //      a ballot names the lane(s) holding it; only an exact d2 tie between lanes needs the index redux.
#pragma unroll
  for (int q = 0; q < NB; ++q) { od[q] = LM_EMPTY_D2; oi[q] = 0xffffffffu; op[q] = 0xffffffffu; acc[q] = false; }
#pragma unroll
  for (int r = 0; r < 5; ++r) {
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      uint32_t (&cd)[LM_CAND_UNROLL] = cdq[q];
      uint32_t (&ci)[LM_CAND_UNROLL] = ciq[q];
      uint32_t (&cp)[LM_CAND_UNROLL] = cpq[q];
      const uint32_t m = __reduce_min_sync(0xffffffffu, cd[0]);
      unsigned bal = __ballot_sync(0xffffffffu, cd[0] == m);
      if (m != LM_EMPTY_D2 && (bal & (bal - 1u)) != 0u) {  // warp-uniform; rare
        const uint32_t om = __reduce_min_sync(0xffffffffu, cd[0] == m ? ci[0] : 0xffffffffu);
        bal = __ballot_sync(0xffffffffu, cd[0] == m && ci[0] == om);
      }
      const int wl = __ffs(bal) - 1;
      const uint32_t wi = __shfl_sync(0xffffffffu, ci[0], wl);
      const uint32_t wp = __shfl_sync(0xffffffffu, cp[0], wl);
      if (lane == r) { od[q] = m; oi[q] = wi; op[q] = wp; }
      if (r == 4) acc[q] = m != LM_EMPTY_D2;
      if (lane == wl && m != LM_EMPTY_D2) {
#pragma unroll
        for (int k = 0; k < LM_CAND_UNROLL - 1; ++k) { cd[k] = cd[k + 1]; ci[k] = ci[k + 1]; cp[k] = cp[k + 1]; }
        cd[LM_CAND_UNROLL - 1] = LM_EMPTY_D2; ci[LM_CAND_UNROLL - 1] = 0xffffffffu; cp[LM_CAND_UNROLL - 1] = 0xffffffffu;
      }
    }
  }
  // ---- queries with more than one chunk of candidates: the remaining chunks, merged into the five found so far.
  //      One copy of the chunk loop (not unrolled over q: the kernel is instruction-cache bound enough as it is).
#pragma unroll 1
  for (int q = 0; q < NB; ++q) {
    bool heavy = false;
    int m = 0, qs = 0, qc = 0;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    uint32_t d = 0, i = 0, p = 0;
#pragma unroll
    for (int t = 0; t < NB; ++t)
      if (t == q) {
        heavy = mq[t] >= 0 && total[t] > 32 * LM_CAND_UNROLL;
        m = mq[t] > 0 ? 1 : 0; qs = q_start[t]; qc = q_count[t];
        sx = sel[t].x; sy = sel[t].y; sz = sel[t].z;
        d = od[t]; i = oi[t]; p = op[t];
      }
    if (!heavy) continue;  // warp-uniform
    Knn5 K;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      K.d2[k] = __shfl_sync(0xffffffffu, d, k);
      K.idx[k] = __shfl_sync(0xffffffffu, i, k);
      K.pos[k] = __shfl_sync(0xffffffffu, p, k);
    }
    knn5_chunks<kFresh>(sorted[m], qs, qc, sx, sy, sz, max_sq_dist, lane, 32 * LM_CAND_UNROLL, K);
    d = LM_EMPTY_D2; i = 0xffffffffu; p = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if (lane == k) { d = K.d2[k]; i = K.idx[k]; p = K.pos[k]; }
    const bool a5 = K.d2[4] != LM_EMPTY_D2;
#pragma unroll
    for (int t = 0; t < NB; ++t)
      if (t == q) { od[t] = d; oi[t] = i; op[t] = p; acc[t] = a5; }
  }
}
#undef LL_KNN_CE

// ---- candidate cache of the single-kernel scan -------------------------------------------------------------------
// Between two LM iterations the pose moves by millimetres, so a query's five nearest map points almost never change —
// but "almost" is not parity. What CAN be proven cheaply: let c be the map-frame position of a query when its
// neighbourhood was last searched in full, d5 its 5th-nearest distance then, and S the set of all map points within
// R = min(1, d5 + margin) of c (complete, because the 27 one-metre cells around c hold every point within 1 m).
// Later the query sits at q, delta = |q - c|. Take the five nearest points of q within S, 5th distance D. Every point
// outside S is farther than R - delta from q, so if  D + delta + slack <= R  no outside point can enter the five
// (not even as a tie), they are the exact answer, found among <= 32 points in shared memory instead of ~100 in L2:
// one candidate per lane, no cell table, no address arithmetic, no sorting network. If the test fails (or S has more
// than 32 points, or fewer than five within range) the query takes the full search again and S is rebuilt around q.
// The arithmetic on every candidate (distance, (d2, index) order) is the same as in the full search, so the result
// is bit-identical to it whenever the test passes; slack = 1 mm covers the rounding of the float quantities in the
// test itself (coordinates below 1e4 m: ulp <= 1e-3 / 1000).
constexpr int LM_QCAP = 144;               // queries per block with a cache slot (the rest always search in full)
constexpr int LM_CACHE_CAND = 32;
constexpr float LM_CACHE_MARGIN = 0.25f;   // metres
constexpr float LM_CACHE_SLACK = 1e-3f;    // metres
// The per-thread test below compares quantities that are all computed from exact float coordinates with a relative error of
// a few ulp (distances below 1.5 m: < 1e-6 m absolute, whatever the magnitude of the coordinates), so 2e-5 m is a wide margin.
constexpr float LM_T1_SLACK = 2e-5f;

struct KnnCache {
  float4* cand;    // [LM_QCAP][32]  x, y, z, map-index bits
  float4* center;  // [LM_QCAP]      c and R (R < 0: cell-complete set; 0 < R <= 1: radius set; R > 1: radius set usable only inside the cell of c)
  float4* qpt;     // [LM_QCAP]      the query itself (sensor frame), so that the steady state never leaves the SM
  float4* t1;      // [LM_QCAP]      c' and G: where the five members were last confirmed, and a lower bound of the distance
                   //                from c' to every map point that is NOT one of them (G <= 0: unknown)
  int* cnt;        // [LM_QCAP]      0 = no set yet, -1 = not cacheable, > 0 = points in the set
  uint32_t* prev;  // [LM_QCAP]      positions in cand of the five neighbours found last time (bit mask), 0 = unknown
  float4* geo;     // [LM_QCAP][2]   the fit's geometry (fit.cuh: FitGeom: g[0..5], ok) and the signature of the neighbour
                   //                order it was computed for (0xffffffff: none) — see lm_block_pass_cached, phase 2
};
constexpr uint32_t LM_GEO_NONE = 0xffffffffu;
#ifndef LM_FILL_FROM
#define LM_FILL_FROM 1  // first LM iteration whose full searches also collect the cached sets (0 and 2 measured: profiles/r02_ab_records.txt)
#endif
// A query's record in global memory (LMArgs::g_guard) seen as slot 0 of a KnnCache: the searches below work on either.
__device__ __forceinline__ KnnCache knn_cache_of_record(float4* rec) {
  KnnCache G;
  G.cand = rec;
  G.center = rec + LM_CACHE_CAND;
  G.qpt = rec + LM_CACHE_CAND + 1;
  G.t1 = rec + LM_CACHE_CAND + 2;
  G.cnt = reinterpret_cast<int*>(rec + LM_CACHE_CAND + 3);
  G.prev = reinterpret_cast<uint32_t*>(rec + LM_CACHE_CAND + 3) + 1;
  G.geo = rec + LM_CACHE_CAND + 4;
  return G;
}
constexpr int LM_CACHE_BYTES = LM_QCAP * (LM_CACHE_CAND * 16 + 16 + 16 + 16 + 4 + 4 + 32);
constexpr int LM_LOOP_SMEM = LM_THREADS * LM_SLOT * (int)sizeof(float) + LM_CACHE_BYTES;
// dynamic shared memory of lm_persistent_kernel (the incremental map assembly borrows it for the added keyframes' sort)
constexpr int LM_PERSISTENT_SMEM = LM_LOOP_SMEM > INC_SMEM_BYTES ? LM_LOOP_SMEM : INC_SMEM_BYTES;

// After a full search of a query around `sel` (5th distance d5bits, LM_EMPTY_D2 when it was not accepted; candidates
// described by my_start / my_count as in knn5_chunks, at most 128 of them): collect the query's cached set. Two kinds:
//   cell-complete (R = -1): ALL points of the 27 cells, when there are at most 32 — valid for as long as the query stays
//     in the same one-metre cell, accepted or not (sparse parts of the map, where most rejected queries live);
//   radius (R > 0): the points within R of the query, as described above.
// All 32 lanes.
// Returns (warp-uniform) the signature of the five neighbours' order as knn5_guard_check computes it — the name under which
// phase 2 stores the fit's geometry of THIS iteration, so that the next iteration's guard test can already reuse it — or
// LM_GEO_NONE when no guard was set up.
template <bool kFresh>
__device__ __noinline__ uint32_t knn_cache_fill(const float4* sorted, int my_start, int my_count, float sx, float sy, float sz,
                                            uint32_t d5bits, uint32_t i5bits, float4 qsensor, int lane, float4* cand,
                                            float4* center, float4* qpt, float4* t1, int* cnt_out, uint32_t* prev_out) {
  constexpr int kRowLane[9] = {12, 3, 9, 15, 21, 0, 6, 18, 24};
  int pr[9], bs[9];
  int total = 0;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int st = __shfl_sync(0xffffffffu, my_start, kRowLane[k]);
    const int cn = __shfl_sync(0xffffffffu, my_count, kRowLane[k]);
    pr[k] = total;
    bs[k] = st - total;
    total += cn;
  }
  if (lane == 0) {
    *qpt = qsensor;
    *prev_out = 0u;
    *t1 = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const bool whole = total <= LM_CACHE_CAND;
  const bool heavy = total > 32 * LM_CAND_UNROLL;
  // A query REJECTED in a dense neighbourhood (fewer than five points within 1 m, many beyond) keeps the points within
  // 1 m + margin of it: while it stays in its one-metre cell (so that the 27 cells still hold every point within 1 m of it)
  // and moves less than the margin, every point within 1 m of it is in that set, and the acceptance test and — should it
  // pass — the five neighbours are decided exactly on the set.
  const bool rejected = !whole && d5bits == LM_EMPTY_D2;
  const uint32_t lt_mask = (1u << lane) - 1u;
  // the radius set must fit 32 points: a dense neighbourhood gets a second try with a thinner shell around d5
  const int first_attempt = (total > 256 && !rejected) ? 1 : 0;
#pragma unroll 1
  for (int attempt = first_attempt; attempt < 3; ++attempt) {
    const float margin = attempt == 0 ? LM_CACHE_MARGIN : 0.3f * LM_CACHE_MARGIN;
    // (R may exceed 1 m — a fifth neighbour near the acceptance limit, or a rejected query: the set is then only usable
    //  while the query stays in its cell, knn5_warp_step2.) In a dense map — un-downsampled: the fifth neighbour centimetres
    //  away — a quarter of a metre around it holds hundreds of points, so the radius is also bounded relative to d5: with
    //  N(r) ~ r^2..r^3 a factor 1.7 (then 1.3, then 1.1) keeps the set near 20 points, still far more margin than a query moves.
    const float s5 = rejected ? 0.f : sqrtf(__uint_as_float(d5bits));
    const float R = whole ? -1.f : (rejected ? 1.0f + margin : fminf(s5 + margin, (attempt == 0 ? 1.7f : (attempt == 1 ? 1.3f : 1.1f)) * s5 + 0.004f));
    const float R2 = R * R;
    int base = 0;
    uint32_t members = 0u, d6 = LM_EMPTY_D2;  // positions of the five neighbours in the set; nearest point that is not one of them
    // rows of cells that lie entirely beyond R hold no point of the set: they are not read, and the nearest point one of
    // them could hold bounds d6 from below (conservative: the guard only gets tighter)
    if (heavy) {  // (only where it pays: a sparse neighbourhood's few dozen candidates are read faster than tested)
      const float fy = sy - floorf(sy), fz = sz - floorf(sz);
      total = 0;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const int st = __shfl_sync(0xffffffffu, my_start, kRowLane[k]);
        int cn = __shfl_sync(0xffffffffu, my_count, kRowLane[k]);
        const int oy = (kRowLane[k] / 3) % 3 - 1, oz = kRowLane[k] / 9 - 1;
        const float dy = oy == 0 ? 0.f : (oy < 0 ? fy : 1.f - fy), dz = oz == 0 ? 0.f : (oz < 0 ? fz : 1.f - fz);
        const float rowmin2 = dy * dy + dz * dz;
        if (cn > 0 && rowmin2 * 0.99999f > R2 + 1e-12f) {
          d6 = min(d6, __float_as_uint(rowmin2 * 0.99999f));
          cn = 0;
        }
        pr[k] = total;
        bs[k] = st - total;
        total += cn;
      }
    }
    // four candidate loads in flight per lane (the loop is otherwise one exposed L1 / L2 latency per 32 candidates: a dense
    // map's 600-candidate query spent more time here than in its search)
#pragma unroll 1
    for (int j0 = 0; j0 < total && base <= LM_CACHE_CAND; j0 += 128) {
      float4 mv[4];
      bool hv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 32 * u + lane;
        int b = bs[0];
#pragma unroll
        for (int k = 1; k < 9; ++k)
          if (j >= pr[k]) b = bs[k];
        hv[u] = j < total;
        mv[u] = hv[u] ? map_ld<kFresh>(&sorted[j + b]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j0 + 32 * u >= total) break;  // warp-uniform
        const bool have = hv[u];
        const float4 m = mv[u];
        float df = sx - m.x;
        float d = df * df;
        df = sy - m.y;
        d += df * df;
        df = sz - m.z;
        d += df * df;
        const bool in = have && (whole || d <= R2);
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        const int pos = base + __popc(bal & lt_mask);
        if (in && pos < LM_CACHE_CAND) cand[pos] = m;
        base += __popc(bal);
        // one of the five found by the search <=> (d2, index) <= the fifth's
        const bool member = have && d5bits != LM_EMPTY_D2 && !pair_less(d5bits, i5bits, __float_as_uint(d), __float_as_uint(m.w));
        members |= (member && in && pos < LM_CACHE_CAND) ? (1u << pos) : 0u;   // (per lane: reduced after the loop)
        d6 = min(d6, (have && !member) ? __float_as_uint(d) : LM_EMPTY_D2);
      }
    }
    members = __reduce_or_sync(0xffffffffu, members);
    d6 = __reduce_min_sync(0xffffffffu, d6);
    const bool fits = whole || (base <= LM_CACHE_CAND && (rejected || base >= 5));
    if (fits || attempt == 2) {
      uint32_t sig = LM_GEO_NONE;
      if (fits && __popc(members) == 5) {  // (warp-uniform) rank of each member, in ascending position, by (d2, index)
        __syncwarp();  // the set was written by all lanes
        const int pos = (int)__fns(members, 0u, (lane < 5 ? lane : 0) + 1);
        const float4 m = cand[pos];
        float df = sx - m.x;  // evalMetric order, as everywhere
        float d = df * df;
        df = sy - m.y;
        d += df * df;
        df = sz - m.z;
        d += df * df;
        const uint32_t db = __float_as_uint(d), ib = __float_as_uint(m.w);
        int rank = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const uint32_t dk = __shfl_sync(0xffffffffu, db, k), ik = __shfl_sync(0xffffffffu, ib, k);
          rank += (k != lane && pair_less(dk, ik, db, ib)) ? 1 : 0;
        }
        sig = __reduce_or_sync(0xffffffffu, lane < 5 ? (uint32_t)rank << (3 * lane) : 0u) | (5u << 15);  // (no sixth member: rank 5)
      }
      if (lane == 0) {
        *center = make_float4(sx, sy, sz, R);
        // (an EMPTY cell-complete set is stored as count 33: cnt must be > 0 to be used, and no lane is below 33 <= 32)
        *cnt_out = fits ? (base > 0 ? base : LM_CACHE_CAND + 1) : -1;
        if (fits && __popc(members) == 5) {
          // the guard of knn5_thread_step1, straight from this scan: it covered every point of the 27 cells, so the
          // nearest point that is not one of the five is at distance sqrt(d6) or beyond the 27 cells (>= 1 m + the distance
          // to the nearest face of the query's cell)
          const float fx = sx - floorf(sx), fy = sy - floorf(sy), fz = sz - floorf(sz);
          const float face = fminf(fminf(fminf(fx, 1.f - fx), fminf(fy, 1.f - fy)), fminf(fz, 1.f - fz));
          const float v6 = __uint_as_float(d6);
          const float s6 = d6 != LM_EMPTY_D2 ? (v6 > 0.f ? v6 * rsqrtf(v6) : 0.f) : 1.f;
          *prev_out = members;
          *t1 = make_float4(sx, sy, sz, fminf(fminf(s6, 1.f + face), 1.f));
        }
      }
      return sig;
    }
  }
  return LM_GEO_NONE;
}

// Step 1 of the per-thread cached search (see knn5_thread_step1): up to six member points (absent: d = LM_EMPTY_D2), the guard
// t1 = (c', G). Returns true and fills the slot when the members still hold the five nearest points of sel.
__device__ __forceinline__ bool knn5_guard_check(const float4 sel, const float4 (&nb)[6], const bool (&have)[6], const float4 t1,
                                                 uint32_t max_bits, uint32_t geo_sig, float* slot) {
  uint32_t d[6], ix[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    float df = sel.x - nb[k].x;  // evalMetric order (nanoflann.hpp:431-439), as in the full search
    float dd = df * df;
    df = sel.y - nb[k].y;
    dd += df * df;
    df = sel.z - nb[k].z;
    dd += df * df;
    d[k] = have[k] ? __float_as_uint(dd) : LM_EMPTY_D2;
    ix[k] = have[k] ? __float_as_uint(nb[k].w) : 0xffffffffu;
  }
  // each member's rank among them by (d2, index), counted directly: fifteen independent comparisons instead of a sorting
  // network's dependent layers
  int rank[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) rank[k] = 0;
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = a + 1; b < 6; ++b) {
      const bool lt = pair_less(d[b], ix[b], d[a], ix[a]);  // b before a (keys are distinct: map indices differ)
      rank[a] += lt ? 1 : 0;
      rank[b] += lt ? 0 : 1;
    }
  uint32_t d5 = 0u;
#pragma unroll
  for (int k = 0; k < 6; ++k) d5 = rank[k] == 4 ? d[k] : d5;
  float df = sel.x - t1.x;
  float dl = df * df;
  df = sel.y - t1.y;
  dl += df * df;
  df = sel.z - t1.z;
  dl += df * df;
  const float D = __uint_as_float(d5);
  const float sD = D > 0.f ? D * rsqrtf(D) : 0.f, sl = dl > 0.f ? dl * rsqrtf(dl) : 0.f;  // 2 ulp: the slack dwarfs it
  if (!(d5 < max_bits && sD + sl + LM_T1_SLACK <= t1.w)) return false;
  slot[0] = sel.x; slot[1] = sel.y; slot[2] = sel.z; slot[3] = 1.f;
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if (rank[k] < 5) { slot[4 + 3 * rank[k]] = nb[k].x; slot[5 + 3 * rank[k]] = nb[k].y; slot[6 + 3 * rank[k]] = nb[k].z; }
  // the order of the members (their ranks, three bits each) names the neighbour list: the same signature as when the
  // fit's geometry was stored = the same five points in the same order = the same geometry, bit for bit
  uint32_t sig = 0u;
#pragma unroll
  for (int k = 0; k < 6; ++k) sig |= (uint32_t)rank[k] << (3 * k);
  slot[19] = sig == geo_sig ? 1.f : 0.f;
  slot[20] = __uint_as_float(sig);
  return true;
}

// The cached search of ONE query by ONE thread — the thread that fits the query in phase 2, so a proven answer goes
// straight into that thread's own staging slot and the steady state of the LM loop has no warp-level search at all.
//   step 1  the five neighbours of last time + a guard distance: t1 = (c', G) says that at c' every map point other than
//           those five was at least G away. At q (delta = |q - c'|) they are at least G - delta away, so if the largest of
//           the five distances from q, D5, satisfies  sqrt(D5) + delta + slack <= G  the five are still the answer (their
//           ORDER may have changed: they are re-sorted by (d2, index)). Five loads, a sorting network, one test.
//   step 2  otherwise the whole cached set (<= 32 points): its seven nearest by (d2, index), found by the WARP — one
//           candidate per lane, seven arg-min rounds (knn5_warp_step2; one thread inserting 32 candidates into a sorted
//           list took 4 us, and the grid barrier made every block wait for the one thread of the scan that had to) —
//           then, by the thread again: the set-level proof of knn_cache_fill's comment (radius sets: sqrt(D5) + |q - c| +
//           slack <= R; cell-complete sets: same cell; rejected-query sets: same cell and |q - c| + slack <= R - 1), and
//           a fresh guard for step 1: G = min(sixth distance within the set, what the set's kind guarantees about the
//           points outside it, 1 m).
// Proven exact: slot filled (slot[3] = 1 with the neighbours in ascending (d2, index) order, or slot[3] = 0 for a query
// proven to have fewer than five points within range); otherwise the query is searched in full.
// Distances are evaluated exactly as in the full search, so a proven answer is bit-identical to it.
// Step 1: returns 1 = proven, 2 = scan the set (step 2); sel_out = the query in the map frame.
__device__ __noinline__ int knn5_thread_step1(const KnnCache C, int i, const Trig* trig_p, float max_sq_dist, float* slot, float4* sel_out) {
  const Trig trig = *trig_p;
  const uint32_t max_bits = __float_as_uint(max_sq_dist);
  const float4 sel = rigid_transform(trig, C.qpt[i]);
  *sel_out = sel;
  const float4* cand = C.cand + i * LM_CACHE_CAND;
  const float4 t1 = C.t1[i];
  if (t1.w < 0.f) {
    // a REJECTED query's guard: at c' the fifth-nearest map point was H = -t1.w > 1 m away, so it is still farther than
    // 1 m from q when |q - c'| + slack <= H - 1: fewer than five points within range, as before.
    float df = sel.x - t1.x;
    float d = df * df;
    df = sel.y - t1.y;
    d += df * df;
    df = sel.z - t1.z;
    d += df * df;
    const float dl = d > 0.f ? d * rsqrtf(d) : 0.f;  // 2 ulp: the slack dwarfs it
    if (dl + LM_T1_SLACK <= -t1.w - 1.f) {
      slot[0] = sel.x; slot[1] = sel.y; slot[2] = sel.z; slot[3] = 0.f;
      return 1;
    }
  } else if (t1.w > 0.f) {
    // the members of last time: five, or six when the fifth and sixth were too close to call (step 2 below)
    uint32_t mask = C.prev[i];
    float4 nb[6];
    bool have[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      have[k] = mask != 0u;
      const int j = have[k] ? __ffs(mask) - 1 : 0;
      mask &= mask - 1u;
      nb[k] = cand[j];
    }
    if (knn5_guard_check(sel, nb, have, t1, max_bits, __float_as_uint(C.geo[2 * i + 1].w), slot)) return 1;
  }
  LL_TL_COUNT_THREAD(6);
  if (t1.w == 0.f) LL_TL_COUNT_THREAD(7);
  return 2;
}

// Step 2 for the lanes in `todo` (all 32 lanes call, converged): C / i / sel / slot are each lane's own query. Returns the
// ballot of the lanes whose outcome is proven.
__device__ __noinline__ unsigned knn5_warp_step2(unsigned todo, const KnnCache C, int i, float4 sel, float max_sq_dist, float* slot) {
  const int lane = threadIdx.x & 31;
  const uint32_t max_bits = __float_as_uint(max_sq_dist);
  const bool mine = (todo >> lane) & 1u;
  constexpr int NT = 7;  // five neighbours, the sixth for the guard, the seventh for the guard of a six-member set
  uint32_t td[NT];
  int tc[NT];
#pragma unroll
  for (int k = 0; k < NT; ++k) { td[k] = LM_EMPTY_D2; tc[k] = 0; }
  const float4* my_cand = C.cand + (mine ? i : 0) * LM_CACHE_CAND;
  const int my_cnt_raw = mine ? C.cnt[i] : 0;
  const int my_cnt = my_cnt_raw <= LM_CACHE_CAND ? my_cnt_raw : 0;  // (an EMPTY set is stored as count 33)
  // ---- the seven nearest of each set by (d2, index): lane j takes candidate j, seven arg-min rounds
#pragma unroll 1
  for (unsigned rest = todo; rest; rest &= rest - 1u) {
    const int src = __ffs(rest) - 1;
    const float4* cand = reinterpret_cast<const float4*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(my_cand), src));
    const int cnt = __shfl_sync(0xffffffffu, my_cnt, src);
    const float qx = __shfl_sync(0xffffffffu, sel.x, src), qy = __shfl_sync(0xffffffffu, sel.y, src), qz = __shfl_sync(0xffffffffu, sel.z, src);
    uint32_t db = LM_EMPTY_D2, ib = 0xffffffffu;
    if (lane < cnt) {
      const float4 m = cand[lane];
      float df = qx - m.x;  // evalMetric order (nanoflann.hpp:431-439), as in the full search
      float d = df * df;
      df = qy - m.y;
      d += df * df;
      df = qz - m.z;
      d += df * df;
      db = __float_as_uint(d);  // (all candidates: a rejected query's guard needs the distances beyond 1 m too)
      ib = __float_as_uint(m.w);
    }
#pragma unroll
    for (int r = 0; r < NT; ++r) {
      const uint32_t m = __reduce_min_sync(0xffffffffu, db);
      unsigned bal = __ballot_sync(0xffffffffu, db == m);
      if (m != LM_EMPTY_D2 && (bal & (bal - 1u)) != 0u) {  // warp-uniform; an exact d2 tie: the lower map index first
        const uint32_t om = __reduce_min_sync(0xffffffffu, db == m ? ib : 0xffffffffu);
        bal = __ballot_sync(0xffffffffu, db == m && ib == om);
      }
      const int wl = __ffs(bal) - 1;
      if (lane == src && m != LM_EMPTY_D2) { td[r] = m; tc[r] = wl; }
      if (lane == wl && m != LM_EMPTY_D2) { db = LM_EMPTY_D2; ib = 0xffffffffu; }
    }
  }
  // ---- each lane of `todo`: the proofs on its own seven
  bool proven = false;
  if (mine) {
    auto approx_sqrt = [](float v) { return v > 0.f ? v * rsqrtf(v) : 0.f; };  // 2 ulp: the slack dwarfs it
    const float4* cand = my_cand;
    const float4 t1_old = C.t1[i];
    const uint32_t old_mask = t1_old.w > 0.f ? C.prev[i] : 0u;
    const bool acc = td[4] < max_bits;  // five points within range
    const float4 ctr = C.center[i];
    const bool same_cell = floorf(sel.x) == floorf(ctr.x) && floorf(sel.y) == floorf(ctr.y) && floorf(sel.z) == floorf(ctr.z);
    float df = sel.x - ctr.x;
    float dc = df * df;
    df = sel.y - ctr.y;
    dc += df * df;
    df = sel.z - ctr.z;
    dc += df * df;
    const float sl = approx_sqrt(dc);
    const float sD = td[4] != LM_EMPTY_D2 ? approx_sqrt(__uint_as_float(td[4])) : 3.f;  // (no fifth point in the set at all: any bound above 2 m serves)
    bool ok;
    float outside;  // lower bound of the distance from sel to every map point that is not in the set
    // distance from sel to the nearest face of its one-metre cell: a point outside the 27 cells is at least 1 m + that away
    const float fx = sel.x - floorf(sel.x), fy = sel.y - floorf(sel.y), fz = sel.z - floorf(sel.z);
    const float face = fminf(fminf(fminf(fx, 1.f - fx), fminf(fy, 1.f - fy)), fminf(fz, 1.f - fz));
    if (ctr.w < 0.f) {          // cell-complete
      ok = same_cell;
      outside = 1.f + face;
    } else if (ctr.w <= 1.f) {  // radius set within 1 m: the ball is inside the 27 cells wherever the query goes
      ok = acc && sD + sl + LM_CACHE_SLACK <= ctr.w;
      outside = ctr.w - sl;
    } else {                    // radius set beyond 1 m (fifth neighbour near the limit, or a rejected query): complete only
                                // for points within 1 m of a query that is still in the cell the set was collected for
      ok = same_cell && (acc ? sD + sl + LM_CACHE_SLACK <= ctr.w : sl + LM_CACHE_SLACK <= ctr.w - 1.f);
      outside = fminf(ctr.w - sl, 1.f + face);
    }
    slot[19] = 0.f;  // (unless the list below turns out to be the one the stored geometry was computed for)
    slot[20] = __uint_as_float(LM_GEO_NONE);
    if (ok && !acc) {
      slot[0] = sel.x; slot[1] = sel.y; slot[2] = sel.z; slot[3] = 0.f;
      // guard for step 1: the fifth-nearest point overall is at least H away (within the set: sD; outside it: `outside`)
      const float H = fminf(sD, outside);
      C.t1[i] = H > 1.f ? make_float4(sel.x, sel.y, sel.z, -H) : make_float4(0.f, 0.f, 0.f, 0.f);
      proven = true;
    } else if (ok) {
      uint32_t mask = 0u;
      slot[0] = sel.x; slot[1] = sel.y; slot[2] = sel.z; slot[3] = 1.f;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float4 nb = cand[tc[k]];
        slot[4 + 3 * k] = nb.x; slot[5 + 3 * k] = nb.y; slot[6 + 3 * k] = nb.z;
        mask |= 1u << tc[k];
      }
      // the guard of step 1: the nearest point that is not a member. When the sixth is within a millimetre of the fifth (a
      // few queries per scan: the guard would never hold) the sixth joins the members and the seventh bounds the rest.
      int nmem = 5;
      float d6 = td[5] != LM_EMPTY_D2 ? approx_sqrt(__uint_as_float(td[5])) : 1.f;  // (capped at 1 m below)
      if (d6 - sD < 1e-3f && td[5] != LM_EMPTY_D2) {
        mask |= 1u << tc[5];
        nmem = 6;
        d6 = td[6] != LM_EMPTY_D2 ? approx_sqrt(__uint_as_float(td[6])) : 1.f;
      }
      // The name of this neighbour list as knn5_guard_check spells it: the members in ascending position, three bits of
      // rank by (d2, index) each (an absent sixth: rank 5). The same members in the same order as when the fit's geometry
      // was stored = the same geometry; and phase 2 stores what it computes now under the new name.
      uint32_t sig = 0u, mm = mask;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        int rank = 5;
        if (mm) {
          const int col = __ffs(mm) - 1;
          mm &= mm - 1u;
#pragma unroll
          for (int r = 0; r < 6; ++r) rank = (r < nmem && tc[r] == col) ? r : rank;
        }
        sig |= (uint32_t)rank << (3 * k);
      }
      slot[19] = (mask == old_mask && sig == __float_as_uint(C.geo[2 * i + 1].w)) ? 1.f : 0.f;
      slot[20] = __uint_as_float(sig);
      C.prev[i] = mask;
      C.t1[i] = make_float4(sel.x, sel.y, sel.z, fminf(fminf(d6, outside), 1.f));
      proven = true;
    }
  }
  return __ballot_sync(0xffffffffu, proven);
}

// The 29 running sums of one LM iteration in a fixed order: 21 upper-triangular AtA entries (row-major),
// 6 Atb entries, sum |d|, count. sv = [row0..row5, b, |d|].
__device__ __forceinline__ float lm_term(const float sv[8], float one, int k) {
  // k is a compile-time constant after unrolling
  if (k < 21) {
    int r = 0, rem = k;
    while (rem >= 6 - r) { rem -= 6 - r; ++r; }
    return sv[r] * sv[r + rem];
  }
  if (k < 27) return sv[k - 21] * sv[6];
  if (k == 27) return sv[7];
  return one;
}

// min_lambda is an OUTPUT of the scan (localInfo, :1602-1607), not an input of the loop: when iteration 0's Sturm count at
// the threshold says that nothing is degenerate, the six multisection rounds that pin the smallest eigenvalue down to float
// resolution leave the critical path of iteration 0 (18 us with cold code) and run one per iteration on warp 1, beside
// warp 0's 6x6 solve of iterations 1..6 (single-kernel scan only: the job lives in shared memory). The same rounds on the
// same matrix in the same order: the same number, bit for bit. A loop that ends early finishes the rounds at its exit.
struct MinLambdaJob {
  float A[36];
  double lo, hi;
  int rounds_left;
};
__device__ __forceinline__ void lm_min_lambda_step(MinLambdaJob* J, LMState* S, int lane, int max_rounds) {
  double lo = J->lo, hi = J->hi;
  int left = J->rounds_left;
  for (int r = 0; r < max_rounds && left > 0; ++r, --left) min_eig6_round(J->A, lane, lo, hi);
  __syncwarp();
  if (lane == 0) {
    J->lo = lo;
    J->hi = hi;
    J->rounds_left = left;
    if (left == 0) S->min_lambda = (float)(0.5 * (lo + hi));
  }
  __syncwarp();
}

// Iteration 0 of LMOptimization, upstream semantics (:1507-1531): min_lambda, the degeneracy flag and the projector,
// by one warp. Out of line: it runs once per scan, and the LM loop around it is instruction-fetch bound.
__device__ __noinline__ void lm_iter0_degeneracy(const float* AtA, float eig_thresh, int lane, float* Ps, float* misc,
                                                 MinLambdaJob* defer) {
  struct { float eig_thresh; } prm = {eig_thresh};

  // min_lambda and "is any eigenvalue below the threshold" from Sturm counts (solve_warp.cuh); the full
  // eigen-decomposition only when the projector needs eigenvectors
  double w = 0.0, v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  double wmin;
  int n_below;
  // (float)lambda < eig_thresh  <=>  lambda < the midpoint between eig_thresh and the float below it (ties round to
  // eig_thresh's even mantissa for the default 100): the double shift that reproduces the float comparison
  const double thr_mid = 0.5 * ((double)prm.eig_thresh + (double)nextafterf(prm.eig_thresh, -FLT_MAX));
  if (defer && sturm_count6(AtA, thr_mid) == 0) {  // (warp-uniform) nothing below the threshold: identity projector now,
    for (int i = lane; i < 36; i += 32) {          // min_lambda later (lm_min_lambda_step)
      defer->A[i] = AtA[i];
      Ps[i] = (i / 6 == i % 6) ? 1.f : 0.f;
    }
    if (lane == 0) {
      double lo, hi;
      min_eig6_bracket(AtA, lo, hi);
      defer->lo = lo;
      defer->hi = hi;
      defer->rounds_left = MIN_EIG6_ROUNDS;
      misc[0] = 0.f;
      misc[1] = 0.f;  // (placeholder until the rounds are done)
    }
    __syncwarp();
    return;
  }
  min_eig6_multisection(AtA, thr_mid, lane, wmin, n_below);
  unsigned mb = n_below > 0 ? 1u : 0u;
  if (mb) {  // degenerate
    jacobi6_warp<true>(AtA, lane, w, v);
    mb = __ballot_sync(FULL, lane < 6 && (float)w < prm.eig_thresh);
    double jm = lane < 6 ? w : 1.7976931348623157e308;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) jm = fmin(jm, __shfl_xor_sync(FULL, jm, o));
    wmin = jm;
  }
  const bool below = lane < 6 && (float)w < prm.eig_thresh;
  if (lane == 0) {
    misc[0] = mb ? 1.f : 0.f;
    misc[1] = (float)wmin;
  }
  // matP = sum over kept eigenvectors v v^T (identity when nothing is dropped)
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      float pv = r == c ? 1.f : 0.f;
      if (mb) {
        double t = (lane < 6 && !below) ? v[r] * v[c] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
        pv = (float)t;
      }
      if (lane == 0) Ps[r * 6 + c] = pv;
    }
}

__device__ __noinline__ float qr66_warp_outofline(const float* AtA, const float* Atb, int lane) { return qr66_warp(AtA, Atb, lane); }

// LMOptimization on the reduced sums, executed by the whole (last) block. sums = s_red[0..28] in shared
// memory; s_red must provide at least 64 + 64 doubles. For eigen_mode = upstream the 6x6 solve runs on warp 0
// and, at iteration 0, the eigen-decomposition runs concurrently on warp 1 (solve_warp.cuh); everything
// else (as shipped mode, fewer than min_residuals rows) takes the single-thread path above.
template <int kEigenMode>
__device__ void lm_solve_block(double* s_red, const LMParams& prm, LMState* S, bool stage_mode, MinLambdaJob* defer = nullptr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = (int)(s_red[28] + 0.5);
  if constexpr (kEigenMode == 0) {
    if (tid == 0) lm_solve(s_red, prm, S, stage_mode);
    return;
  }
  if (n < prm.min_residuals) {  // LMOptimization returns false before computing anything (:1453)
    if (tid == 0) lm_apply_and_control(nullptr, n, s_red[27], nullptr, nullptr, prm, S, stage_mode);
    return;
  }
  float* sh = reinterpret_cast<float*>(s_red + 64);
  float* AtA = sh;        // [36] row-major
  float* Atb = sh + 36;   // [6]
  float* xs = sh + 42;    // [6]
  float* Ps = sh + 48;    // [36] row-major
  float* misc = sh + 84;  // [0] = is_degenerate (as float), [1] = min_lambda
  const int iter = S->iter;
  if (tid < 36) {
    const int r = tid / 6, c = tid % 6, i = r < c ? r : c, j = r < c ? c : r;
    AtA[tid] = (float)s_red[i * 6 - (i * (i - 1)) / 2 + (j - i)];
  } else if (tid < 42) {
    Atb[tid - 36] = (float)s_red[21 + tid - 36];
  }
  __syncthreads();
  if (warp == 0) {  // :1505
    float x = 0.f;
    float xv[6];
    if (prm.lm_solver == 1 && ldlt6_solve(AtA, Atb, xv)) {  // every lane computes the same six numbers
#pragma unroll
      for (int r = 0; r < 6; ++r)
        if (lane == r) x = xv[r];
    } else {
      x = qr66_warp_outofline(AtA, Atb, lane);
    }
    if (lane < 6) xs[lane] = x;
  } else if (warp == 1 && iter == 0) {  // :1507-1531, upstream semantics
    lm_iter0_degeneracy(AtA, prm.eig_thresh, lane, Ps, misc, defer);
  } else if (warp == 1 && defer && defer->rounds_left > 0) {
    lm_min_lambda_step(defer, S, lane, 1);  // one deferred multisection round beside warp 0's solve
  }
  __syncthreads();
  if (warp != 0) return;
  // ---- the rest of LMOptimization (:1533-1569) and the loop control (:1578-1586) on warp 0: lane r owns
  //      component r of matX / the pose; lanes 0..2 also own one angle's sin / cos each
  int degenerate;
  if (iter == 0) {
    degenerate = misc[0] != 0.f ? 1 : 0;
    if (lane == 0) {
      S->is_degenerate = degenerate;
      S->min_lambda = misc[1];
    }
    if (lane < 6)
      for (int c = 0; c < 6; ++c) S->matP[lane * 6 + c] = Ps[lane * 6 + c];
  } else {
    degenerate = S->is_degenerate;
    if (degenerate && lane < 6)
      for (int c = 0; c < 6; ++c) Ps[lane * 6 + c] = S->matP[lane * 6 + c];
  }
  float x = lane < 6 ? xs[lane] : 0.f;
  if (degenerate && lane < 6) {  // matX = matP * matX2
    float acc = 0.f;
    for (int c = 0; c < 6; ++c) acc += Ps[lane * 6 + c] * xs[c];
    x = acc;
  }
  float pose = lane < 6 ? S->pose[lane] : 0.f;
  if (prm.apply_update && lane < 6) {  // :1540-1545
    pose += x;
    S->pose[lane] = pose;
  }
  if (lane < 6) {
    S->last_x[lane] = x;
    S->Atb[lane] = Atb[lane];
    for (int c = 0; c < 6; ++c) S->AtA[lane * 6 + c] = AtA[lane * 6 + c];
  }
  // deltaR / deltaT (:1547-1552): pcl::rad2deg(float) then pow(., 2) and sqrt in double
  const float scaled = lane < 3 ? x * 57.29578f : x * 100.f;
  const double sq = (double)scaled * (double)scaled;
  const double q0 = __shfl_sync(FULL, sq, 0), q1 = __shfl_sync(FULL, sq, 1), q2 = __shfl_sync(FULL, sq, 2);
  const double q3 = __shfl_sync(FULL, sq, 3), q4 = __shfl_sync(FULL, sq, 4), q5 = __shfl_sync(FULL, sq, 5);
  const float deltaR = (float)sqrt(q0 + q1 + q2);
  const float deltaT = (float)sqrt(q3 + q4 + q5);
  const int ret = (deltaR < prm.stop_thres && deltaT < prm.stop_thres) ? 1 : 0;
  const bool pose_changed = prm.apply_update != 0;
  if (pose_changed) {  // updatePointAssociateToMapSinCos (:591-604) for the next iteration
    if (lane < 3) {
      double sn, cs;
      sincos((double)pose, &sn, &cs);
      const float fs = (float)sn, fc = (float)cs;
      if (lane == 0) { S->trig.cr = fc; S->trig.sr = fs; S->rtrig.crx = fc; S->rtrig.srx = fs; }
      if (lane == 1) { S->trig.cp = fc; S->trig.sp = fs; S->rtrig.cry = fc; S->rtrig.sry = fs; }
      if (lane == 2) { S->trig.cy = fc; S->trig.sy = fs; S->rtrig.crz = fc; S->rtrig.srz = fs; }
    } else if (lane == 3) {
      S->trig.tx = pose;
    } else if (lane == 4) {
      S->trig.ty = pose;
    } else if (lane == 5) {
      S->trig.tz = pose;
    }
  }
  if (lane == 0) {
    S->sel_num = n;
    S->cf_mean = (float)s_red[27] / (float)n;  // :1555-1568
    S->deltaR = deltaR;
    S->deltaT = deltaT;
    S->lm_ret = ret;
    if (!stage_mode) {
      const int executed = iter + 1;
      if (ret == 1) {
        S->converged = 1;
        S->iter_num = executed;
        S->done = 1;
      } else if (executed >= prm.iter_count_thres) {
        S->iter_num = executed;
        S->done = 1;
      } else if (!pose_changed) {  // nothing can change any more (see lm_apply_and_control)
        S->iter_num = prm.iter_count_thres;
        S->done = 1;
      }
      S->iter = executed;
    }
  }
}

// The 29 sums of a block's rows, from the per-query vectors sv = [row0..row5, b, |d|] and `one` (1 for a kept row, else
// every entry is 0). Every thread < nslots_pad parks its nine numbers in ITS staging slot (stride 21 floats: conflict
// free), then sum k is accumulated by eight threads (8 k + j, j = 0..7), thread j taking slots j, j + 8, ... in order,
// and the eight parts are added pairwise (xor 1, 2, 4). Fixed order => bitwise repeatable, and the same order in the
// single-kernel scan and in lm_iter_kernel. Replaces 29 five-step shuffle trees per warp (1.65 us + 0.3 us per
// iteration, profiles/r02_lm_iteration_timeline_before.txt). Returns sum k on thread 8 k (0 elsewhere / for k >= 29).
// Must be called by all threads of the block; contains one __syncthreads (the caller's trailing one protects reuse).
__device__ __forceinline__ float lm_block_sums(float* s_slot, const float (&sv)[8], float one, int nslots) {
  const int tid = threadIdx.x;
  if (tid < nslots) {
    float* v = &s_slot[tid * LM_SLOT];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = sv[r];
    v[8] = one;
  }
  __syncthreads();
  float acc = 0.f;
  const int k = tid >> 3, j = tid & 7;
  if (k < LM_NSUM) {
    int ia, ib;  // term k = v[ia] * v[ib]  (|d| = v[7] * 1, count = 1 * 1: exact)
    if (k < 21) {
      int r = 0, rem = k;
      while (rem >= 6 - r) { rem -= 6 - r; ++r; }
      ia = r; ib = r + rem;
    } else if (k < 27) {
      ia = k - 21; ib = 6;
    } else if (k == 27) {
      ia = 7; ib = 8;
    } else {
      ia = 8; ib = 8;
    }
    for (int i = j; i < nslots; i += 8) {
      const float* v = &s_slot[i * LM_SLOT];
      acc += v[ia] * v[ib];
    }
  }
  if (tid < 256) {  // (whole warps: 29 x 8 = 232 threads, rounded up)
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  }
  return (j == 0 && k < LM_NSUM) ? acc : 0.f;
}

// Sum the block partials in a fixed order (deterministic regardless of which block arrives last).
template <int kEigenMode>
__device__ void reduce_partials_and_solve(const float* partials, int nblocks, const LMParams& prm, LMState* S,
                                          bool stage_mode, double* s_red /* [nwarps][32] shared */, MinLambdaJob* defer = nullptr) {
  const int lane = threadIdx.x & 31, sub = threadIdx.x >> 5, nsub = blockDim.x >> 5;
  double acc = 0.0;
  constexpr int kInFlight = 10;  // 148 blocks / 16 warps: every partial of a warp's share is fetched in one round trip
  for (int b = sub; b < nblocks; b += kInFlight * nsub) {  // added in the same order as a plain loop
    float v[kInFlight];
#pragma unroll
    for (int u = 0; u < kInFlight; ++u) {
      const int bb = b + u * nsub;
      v[u] = bb < nblocks ? __ldcg(&partials[bb * 32 + lane]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kInFlight; ++u)
      if (b + u * nsub < nblocks) acc += (double)v[u];
  }
  s_red[sub * 32 + lane] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = 0.0;
    for (int w = 0; w < nsub; ++w) t += s_red[w * 32 + lane];
    s_red[lane] = t;
  }
  __syncthreads();
  LL_TL(S->iter, 8);
  lm_solve_block<kEigenMode>(s_red, prm, S, stage_mode, defer);
}

// Block b's share of the queries: groups of LM_QGROUP consecutive corner queries dealt round-robin to the blocks, and the
// same for the surf queries (local order: the corner queries, then the surf queries). Neighbouring queries of a scan fall
// into map regions of similar density: contiguous per-block slices made a few blocks the stragglers of every iteration's
// grid barrier on a dense map (configs[3], hundreds of candidates per query: first arrival -> last arrival 500 us of a
// 780 us iteration), while dealing single queries gives up the L1 / L2 locality of neighbouring queries' candidate rows
// (+1.6 % on the VLP-16 replay). Every block also runs one warp's worth of line fits (closed-form eigen solve in double,
// the slowest piece of phase 2) beside its plane fits.
constexpr int LM_QGROUP = 16;
struct QShare {
  int b, nb, n_c, n_s;
  __device__ __forceinline__ int ncb() const { return n_c; }
  __device__ __forceinline__ int size() const { return n_c + n_s; }
  __device__ __forceinline__ int deal(int L) const { return ((L / LM_QGROUP) * nb + b) * LM_QGROUP + (L % LM_QGROUP); }
  // local index L -> index in the concatenated query list (corner queries first)
  __device__ __forceinline__ int qi(int L, int nq0) const { return L < n_c ? deal(L) : nq0 + deal(L - n_c); }
};
// queries of a list of n that block b of nb owns
__device__ __forceinline__ int lm_share_count(int n, int b, int nb) {
  const int groups = (n + LM_QGROUP - 1) / LM_QGROUP;              // the last one may be partial
  if (groups <= b) return 0;
  const int mine = (groups - b + nb - 1) / nb;                     // groups b, b + nb, ...
  const bool owns_last = (groups - 1 - b) % nb == 0;
  return mine * LM_QGROUP - (owns_last ? groups * LM_QGROUP - n : 0);
}
__device__ __forceinline__ QShare lm_partition(const LMArgs& A, int b, int nb) {
  QShare sh;
  sh.b = b;
  sh.nb = nb;
  sh.n_c = lm_share_count(A.nq[0], b, nb);
  sh.n_s = lm_share_count(A.nq[1], b, nb);
  return sh;
}

// Phases 1 and 2 of one LM iteration over queries [q_begin, q_end) of the concatenated query list (corner
// queries first, then surf):
//   phase 1  one warp per query: pointAssociateToMap + 5-NN; the five neighbours and the transformed query
//            are staged in shared memory (21 floats per query, odd stride => conflict-free);
//   phase 2  one thread per query: line / plane fit, Jacobian row, then a shuffle-tree reduction of the 29
//            sums over the block (fixed order => bitwise repeatable).
// Returns, on thread 8 k (k < 29), this block's total of sum k.
template <int kEigenMode, bool kFresh>
__device__ __forceinline__ float lm_block_pass(const LMArgs& A, const Trig& trig, const RotTrig& rt,
                                               const GridGeom (&geom)[2], float* s_slot, const QShare& sh) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float run_sum = 0.f;  // thread 8 k, k < 29: this block's running total of sum k (see lm_block_sums)
  const int n_local = sh.size();
  for (int c0 = 0; c0 < n_local; c0 += LM_CHUNK) {
    const int cn = min(LM_CHUNK, n_local - c0);
    // Staging slot of query i of this pass = the thread that fits it in phase 2. Corner queries come first; the
    // surf queries start at the next warp boundary so that no warp runs corner_fit and surf_fit back to back.
    const int nc = max(0, min(cn, sh.ncb() - c0));
    const int nc_pad = (nc + 31) & ~31;
    const int nslots = nc_pad + (cn - nc);
    // ---- phase 1: k-NN, one warp per query, LM_KNN_BATCH queries of a warp at a time
    for (int i0 = warp; i0 < cn; i0 += LM_KNN_BATCH * LM_WARPS) {
      int mq[LM_KNN_BATCH];
      float4 sel[LM_KNN_BATCH];
#pragma unroll
      for (int b = 0; b < LM_KNN_BATCH; ++b) {
        const int i = i0 + b * LM_WARPS;
        mq[b] = -1;
        sel[b] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < cn) {
          const int qi = sh.qi(c0 + i, A.nq[0]);
          mq[b] = qi < A.nq[0] ? 0 : 1;
          sel[b] = __ldg(lm_query(A, qi));
        }
      }
#pragma unroll
      for (int b = 0; b < LM_KNN_BATCH; ++b)
        if (A.mode != 2) sel[b] = rigid_transform(trig, sel[b]);
      uint32_t od[LM_KNN_BATCH], oi[LM_KNN_BATCH], op[LM_KNN_BATCH];
      bool acc[LM_KNN_BATCH];
      int q_start[LM_KNN_BATCH], q_count[LM_KNN_BATCH], q_total[LM_KNN_BATCH];
      warp_knn5_batch<kFresh, LM_KNN_BATCH>(A.sorted, A.cells, geom, mq, sel, A.prm.fit.max_sq_dist, lane, od, oi, op, acc, q_start, q_count, q_total);
#pragma unroll
      for (int b = 0; b < LM_KNN_BATCH; ++b) {
        const int i = i0 + b * LM_WARPS;
        if (i >= cn) continue;  // warp-uniform
        const int qi = sh.qi(c0 + i, A.nq[0]);
        float* slot = &s_slot[(i < nc ? i : nc_pad + (i - nc)) * LM_SLOT];
        if (lane < 5) {
          if (acc[b]) {
            const float4 nb = map_ld<kFresh>(&A.sorted[mq[b]][op[b]]);
            slot[4 + 3 * lane] = nb.x;
            slot[5 + 3 * lane] = nb.y;
            slot[6 + 3 * lane] = nb.z;
          }
          if (A.mode != 0) {
            const bool have = od[b] != LM_EMPTY_D2;
            if (A.e_idx) A.e_idx[5 * qi + lane] = have ? (int)oi[b] : -1;
            if (A.e_d2) A.e_d2[5 * qi + lane] = have ? __uint_as_float(od[b]) : FLT_MAX;
          }
        } else if (lane == 5) {
          slot[0] = sel[b].x;
          slot[1] = sel[b].y;
          slot[2] = sel[b].z;
          slot[3] = acc[b] ? 1.f : 0.f;
          if (A.mode != 0 && A.e_accept) A.e_accept[qi] = acc[b] ? 1 : 0;
        }
      }
    }
    __syncthreads();
    // ---- phase 2: fit + Jacobian row, one thread per query
    if (A.mode != 2) {
      float sv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      float one = 0.f;
      const int my_i = tid < nc ? tid : (tid >= nc_pad && tid < nslots ? nc + (tid - nc_pad) : -1);
      if (my_i >= 0) {
        const int qi = sh.qi(c0 + my_i, A.nq[0]);
        const int m = qi < A.nq[0] ? 0 : 1;
        const float* slot = &s_slot[tid * LM_SLOT];
        float coeff[4] = {0.f, 0.f, 0.f, 0.f};
        bool keep = false;
        if (slot[3] != 0.f) {
          float nbx[5], nby[5], nbz[5];
#pragma unroll
          for (int k = 0; k < 5; ++k) { nbx[k] = slot[4 + 3 * k]; nby[k] = slot[5 + 3 * k]; nbz[k] = slot[6 + 3 * k]; }
          keep = m == 0 ? corner_fit<kEigenMode>(nbx, nby, nbz, slot[0], slot[1], slot[2], A.prm.fit, coeff)
                        : surf_fit(nbx, nby, nbz, slot[0], slot[1], slot[2], A.prm.fit, coeff);
        }
        if (A.mode != 0) {
          if (A.e_mask) A.e_mask[qi] = keep ? 1 : 0;
          if (A.e_coeff) {
#pragma unroll
            for (int k = 0; k < 4; ++k) A.e_coeff[4 * qi + k] = keep ? coeff[k] : 0.f;
          }
        } else if (keep) {
          const float4 p = __ldg(lm_query(A, qi));
          jacobian_row(rt, p.x, p.y, p.z, coeff[0], coeff[1], coeff[2], sv);
          sv[6] = -A.prm.step_size * coeff[3];  // matB (:1500)
          sv[7] = fabsf(coeff[3]);
          one = 1.f;
        }
      }
      if (A.mode == 0) run_sum += lm_block_sums(s_slot, sv, one, nslots);
    }
    __syncthreads();
  }
  return run_sum;
}

// One round of step (B): the full search of the next NB queries in `need`, their staging slots, and their cached sets.
template <bool kFresh, int NB>
__device__ __forceinline__ void lm_full_search_round(const LMArgs& A, const Trig& trig, const GridGeom (&geom)[2], float* s_slot,
                                                     const KnnCache& C, const QShare& sh, int c0, int nc, int nc_pad, unsigned& need,
                                                     bool first_chunk, bool do_fill) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int jb[NB], mq[NB];
  float4 sel[NB], qs[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    jb[b] = -1;
    mq[b] = -1;
    sel[b] = make_float4(0.f, 0.f, 0.f, 0.f);
    qs[b] = sel[b];
    if (need) {
      jb[b] = __ffs(need) - 1;
      need &= need - 1u;
      const int qi = sh.qi(c0 + warp + LM_WARPS * jb[b], A.nq[0]);
      mq[b] = qi < A.nq[0] ? 0 : 1;
      qs[b] = __ldcg(lm_query(A, qi));  // (L2: the clouds may arrive during the launch)
      sel[b] = rigid_transform(trig, qs[b]);
    }
  }
  uint32_t od[NB], oi[NB], op[NB];
  bool acc[NB];
  int q_start[NB], q_count[NB], q_total[NB];
  warp_knn5_batch<kFresh, NB>(A.sorted, A.cells, geom, mq, sel, A.prm.fit.max_sq_dist, lane, od, oi, op, acc, q_start, q_count, q_total);
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    if (jb[b] < 0) continue;  // warp-uniform
    LL_TL_COUNT(0);
    if (q_total[b] > 32 * LM_CAND_UNROLL) LL_TL_COUNT(1);
    else if (!acc[b] && q_total[b] > LM_CACHE_CAND) LL_TL_COUNT(3);
    const int i = warp + LM_WARPS * jb[b];
    float* slot = &s_slot[(i < nc ? i : nc_pad + (i - nc)) * LM_SLOT];
    if (lane < 5) {
      if (acc[b]) {
        const float4 nb = map_ld<kFresh>(&A.sorted[mq[b]][op[b]]);
        slot[4 + 3 * lane] = nb.x;
        slot[5 + 3 * lane] = nb.y;
        slot[6 + 3 * lane] = nb.z;
      }
    } else if (lane == 5) {
      slot[0] = sel[b].x;
      slot[1] = sel[b].y;
      slot[2] = sel[b].z;
      slot[3] = acc[b] ? 1.f : 0.f;
      slot[19] = 0.f;  // a new neighbour list: no stored geometry applies
      slot[20] = __uint_as_float(LM_GEO_NONE);
    }
    const bool has_slot = first_chunk && i < LM_QCAP;
    if (do_fill && (has_slot || A.g_guard)) {  // (re)build this query's cached set around where it is now
      const KnnCache Cq = has_slot ? C : knn_cache_of_record(A.g_guard + (size_t)sh.qi(c0 + i, A.nq[0]) * LM_GUARD_STRIDE);
      const int ci = has_slot ? i : 0;
      const uint32_t d5 = __shfl_sync(0xffffffffu, acc[b] ? od[b] : LM_EMPTY_D2, 4);
      const uint32_t i5 = __shfl_sync(0xffffffffu, oi[b], 4);
      const uint32_t sig = knn_cache_fill<kFresh>(A.sorted[mq[b]], q_start[b], q_count[b], sel[b].x, sel[b].y, sel[b].z, d5, i5, qs[b],
                                                  lane, &Cq.cand[ci * LM_CACHE_CAND], &Cq.center[ci], &Cq.qpt[ci], &Cq.t1[ci],
                                                  &Cq.cnt[ci], &Cq.prev[ci]);
      if (lane == 5) slot[20] = __uint_as_float(sig);  // phase 2 stores this iteration's geometry under that name
    }
  }
}

// Step (B) of lm_block_pass_cached for one warp: the full search of the queries in `need` (bit j = query warp + 16 j of
// the chunk starting at c0), three at a time (a single remaining one by itself), their staging slots, and their cached
// sets. Out of line on purpose.
template <bool kFresh>
__device__ __noinline__ void lm_full_search_warp(const LMArgs* Ap, const Trig* trig_p, const GridGeom* geom_p, float* s_slot,
                                                 KnnCache C, QShare sh, int c0, int nc, int nc_pad, unsigned need, bool first_chunk,
                                                 bool do_fill) {
  const LMArgs& A = *Ap;
  const Trig trig = *trig_p;
  const GridGeom geom[2] = {geom_p[0], geom_p[1]};
  while (need) {
    if ((need & (need - 1u)) == 0u) lm_full_search_round<kFresh, 1>(A, trig, geom, s_slot, C, sh, c0, nc, nc_pad, need, first_chunk, do_fill);
    else lm_full_search_round<kFresh, LM_KNN_BATCH>(A, trig, geom, s_slot, C, sh, c0, nc, nc_pad, need, first_chunk, do_fill);
  }
}

// The same two phases for the single-kernel scan, with the candidate cache (KnnCache above) in front of the full search.
// Warp w owns queries w, w + 16, ... of the block's share. Per iteration: (A) the queries with a cached set try the
// cached search, four at a time; (B) the others — all of them in iteration 0 — take the full search three at a time
// and (re)build their sets. Phase 2 and the sums are those of lm_block_pass, so both kernels add in the same order.
template <int kEigenMode, bool kFresh>
__device__ __forceinline__ float lm_block_pass_cached(const LMArgs& A, const Trig& trig_s, const RotTrig& rt,
                                                      const GridGeom* geom_s, float* s_slot, const KnnCache& C,
                                                      const QShare& sh, int tl_it) {
  // trig_s / geom_s live in shared memory (the out-of-line searches take them by address)
  __shared__ unsigned s_need[LM_WARPS];
  const int tid = threadIdx.x, warp = tid >> 5;
  float run_sum = 0.f;  // thread 8 k, k < 29: this block's running total of sum k
  const int n_local = sh.size();
  for (int c0 = 0; c0 < n_local; c0 += LM_CHUNK) {
    const int cn = min(LM_CHUNK, n_local - c0);
    const int nc = max(0, min(cn, sh.ncb() - c0));
    const int nc_pad = (nc + 31) & ~31;
    const int nslots = nc_pad + (cn - nc);
    const bool first_chunk = c0 == 0;  // the cache slots belong to the first LM_QCAP queries of the block
    // ---- (A) cached search, one THREAD per query: the thread that will fit it (knn5_thread_step1; a set that has to be
    //      re-scanned is re-scanned by the warp, knn5_warp_step2). What it cannot prove is
    //      flagged for its warp: bit j of s_need[w] = query w + 16 j of this chunk.
    const int my_i = tid < nc ? tid : (tid >= nc_pad && tid < nslots ? nc + (tid - nc_pad) : -1);
    if (tid < LM_WARPS) s_need[tid] = 0u;
    __syncthreads();
    LL_TL_WARP(0);
    {
      int state = 0;  // 0 = no cached set: full search; 1 = proven; 2 = the set must be scanned (by the warp, below)
      KnnCache Cq = C;
      int ci = 0;
      float4 sel = make_float4(0.f, 0.f, 0.f, 0.f);
      if (my_i >= 0) {
        bool cached = false;
        if (first_chunk && my_i < LM_QCAP) {
          ci = my_i;
          cached = C.cnt[my_i] > 0;
        } else if (A.g_guard) {
          Cq = knn_cache_of_record(A.g_guard + (size_t)sh.qi(c0 + my_i, A.nq[0]) * LM_GUARD_STRIDE);
          cached = Cq.cnt[0] > 0;
        }
        if (cached) {
          state = knn5_thread_step1(Cq, ci, &trig_s, A.prm.fit.max_sq_dist, &s_slot[tid * LM_SLOT], &sel);
          LL_TL_COUNT_THREAD(5);
        }
      }
      const unsigned todo = __ballot_sync(0xffffffffu, state == 2);
      if (todo) {  // warp-uniform
        const unsigned done = knn5_warp_step2(todo, Cq, ci, sel, A.prm.fit.max_sq_dist, &s_slot[tid * LM_SLOT]);
        if (state == 2) state = (done >> (tid & 31)) & 1u ? 1 : 0;
      }
#ifdef LL_TIMELINE
      if (state == 1) LL_TL_COUNT_THREAD(4);
#endif
      if (my_i >= 0 && state != 1) atomicOr(&s_need[my_i % LM_WARPS], 1u << (my_i / LM_WARPS));
    }
    __syncthreads();
    const unsigned need = s_need[warp];
    LL_TL_WARP(1);
    // ---- (B) full search of what is left (everything in iteration 0), out of line: the steady-state loop is
    //      instruction-fetch bound, and this is by far its largest piece of code
    // The cached sets are first collected in iteration 1: iteration 0 starts from the prior, its pose step is by far the
    // largest, and sets collected around the prior's positions mostly failed their proofs one iteration later (55 % of the
    // queries re-scanned their sets and 1-2 % searched again and refilled; profiles/r02_scan_timeline_*.txt) — so
    // iteration 0 only searches, iteration 1 searches again from nearly converged positions and fills.
    if (need) lm_full_search_warp<kFresh>(&A, &trig_s, geom_s, s_slot, C, sh, c0, nc, nc_pad, need, first_chunk, tl_it >= LM_FILL_FROM);
    LL_TL_WARP(5);
    LL_TL(tl_it, 1);
    __syncthreads();
    LL_TL(tl_it, 2);
    // ---- phase 2: fit + Jacobian row, one thread per query
    float sv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float one = 0.f;
    if (my_i >= 0) {
      const int qi = sh.qi(c0 + my_i, A.nq[0]);
      const int m = qi < A.nq[0] ? 0 : 1;
      const float* slot = &s_slot[tid * LM_SLOT];
      float coeff[4] = {0.f, 0.f, 0.f, 0.f};
      bool keep = false;
      if (slot[3] != 0.f) {
        // The geometry of the fit (line points / plane) depends on the neighbour list only: when the guard test of phase A
        // confirmed the SAME five points in the SAME order as when it was last computed (signature in slot[20]), it is
        // taken from the query's record instead of running the eigen solve / QR again — the same numbers, bit for bit.
        const bool has_slot = first_chunk && my_i < LM_QCAP;
        const bool has_rec = has_slot || A.g_guard != nullptr;
        float4* geo = nullptr;
        if (has_rec) geo = has_slot ? C.geo + 2 * my_i : A.g_guard + (size_t)qi * LM_GUARD_STRIDE + LM_CACHE_CAND + 4;
        FitGeom G;
        if (has_rec && slot[19] != 0.f) {
          const float4 g0 = geo[0], g1 = geo[1];
          G.g[0] = g0.x; G.g[1] = g0.y; G.g[2] = g0.z; G.g[3] = g0.w; G.g[4] = g1.x; G.g[5] = g1.y;
          G.ok = g1.z != 0.f;
        } else {
          float nbx[5], nby[5], nbz[5];
#pragma unroll
          for (int k = 0; k < 5; ++k) { nbx[k] = slot[4 + 3 * k]; nby[k] = slot[5 + 3 * k]; nbz[k] = slot[6 + 3 * k]; }
          if (m == 0) corner_geometry<kEigenMode>(nbx, nby, nbz, A.prm.fit, G);
          else surf_geometry(nbx, nby, nbz, A.prm.fit, G);
          if (has_rec) {
            geo[0] = make_float4(G.g[0], G.g[1], G.g[2], G.g[3]);
            geo[1] = make_float4(G.g[4], G.g[5], G.ok ? 1.f : 0.f, slot[20]);
          }
        }
        keep = m == 0 ? corner_residual(G, slot[0], slot[1], slot[2], A.prm.fit, coeff)
                      : surf_residual(G, slot[0], slot[1], slot[2], A.prm.fit, coeff);
      }
      if (keep) {
        const float4 p = __ldcg(lm_query(A, qi));
        jacobian_row(rt, p.x, p.y, p.z, coeff[0], coeff[1], coeff[2], sv);
        sv[6] = -A.prm.step_size * coeff[3];  // matB (:1500)
        sv[7] = fabsf(coeff[3]);
        one = 1.f;
      }
    }
    LL_TL(tl_it, 3);
    run_sum += lm_block_sums(s_slot, sv, one, nslots);
    LL_TL(tl_it, 4);
    __syncthreads();
  }
  return run_sum;
}

__device__ __forceinline__ void lm_load_geom(const LMArgs& A, GridGeom (&geom)[2]) {
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    // __ldcg: in the persistent kernel these fields may have been written by another block moments ago, and this
    // block's L1 can still hold the MapState line from reading the bounding box
    const MapState* ms = A.mst[m];
    geom[m].ox = __ldcg(&ms->gox); geom[m].oy = __ldcg(&ms->goy); geom[m].oz = __ldcg(&ms->goz);
    geom[m].dx = __ldcg(&ms->gdx); geom[m].dy = __ldcg(&ms->gdy); geom[m].dz = __ldcg(&ms->gdz);
    geom[m].tag = __ldcg(&ms->cell_tag);
  }
}

// One LM iteration per launch (stage-level entry points, and scan2map when persistent_lm = 0). Each block owns
// a contiguous chunk of the query list; block partial -> global; the last block to arrive adds all partials in
// a fixed order and solves.
template <int kEigenMode>
__global__ void __launch_bounds__(LM_THREADS, LM_MIN_BLOCKS) lm_iter_kernel(const __grid_constant__ LMArgs A) {
  LMState* S = A.st;
  // Programmatic dependent launch (see enqueue_iterations): wait for the previous grid's results, then let the
  // next grid start its own launch. Both are no-ops for a plain launch.
  pdl_entry();
  if (A.mode == 0 && S->done) return;
  __shared__ __align__(16) float s_slot[LM_THREADS * LM_SLOT];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const Trig trig = S->trig;
  const RotTrig rt = S->rtrig;
  GridGeom geom[2];
  lm_load_geom(A, geom);
  const QShare sh = lm_partition(A, (int)blockIdx.x, (int)gridDim.x);
  const float run_sum = lm_block_pass<kEigenMode, false>(A, trig, rt, geom, s_slot, sh);
  if (A.mode != 0) return;
  if (tid < 256 && (tid & 7) == 0) {  // thread 8 k holds sum k (k >= 29: zero)
    A.partials[blockIdx.x * 32 + (tid >> 3)] = run_sum;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned int t = atomicAdd(&S->ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid == 0) S->ticket = 0;
  reduce_partials_and_solve<kEigenMode>(A.partials, gridDim.x, A.prm, S, false, reinterpret_cast<double*>(s_slot));
}

// The map-size guard and loop state of scan2MapOptimization (:1573), applied to *S by one thread.
__device__ __forceinline__ void lm_begin_state(LMState* S, const MapState* corner, const MapState* surf, const LMParams& prm,
                                               const float pose6[6]) {
#pragma unroll
  for (int i = 0; i < 6; ++i) S->pose[i] = pose6[i];
  set_trig(S);
  S->iter = 0;
  S->iter_num = 0;
  S->converged = 0;
  S->sel_num = 0;
  S->lm_ret = 0;
  S->ticket = 0;
  const int n_corner = __ldcg(&corner->n_ds), n_surf = __ldcg(&surf->n_ds);
  S->corner_map_ds = n_corner;
  S->surf_map_ds = n_surf;
  int flags = 0;
  if (__ldcg(&corner->voxel_overflow)) flags |= 1;
  if (__ldcg(&surf->voxel_overflow)) flags |= 2;
  if (__ldcg(&corner->grid_too_large) || __ldcg(&surf->grid_too_large)) flags |= 4;
  if (__ldcg(&corner->inc_fail) || __ldcg(&surf->inc_fail)) flags |= 8;  // incremental map assembly gave up: the host rebuilds
  S->flags = flags;
  const bool ok = n_corner > prm.min_corner_map && n_surf > prm.min_surf_map && !(flags & 12);
  // 1 = the reference's own skip (:1573); 2 = NOT a reference behaviour: the 1 m search grid over this map's bounding
  // box exceeds the cell budget, so the scan could not be registered here although the reference's KD-tree would
  S->status = ok ? 0 : ((flags & 8) ? 3 : ((flags & 4) ? 2 : 1));  // (3 never leaves the library: abi.cu rebuilds and reruns)
  S->done = ok ? 0 : 1;
  if (prm.iter_count_thres <= 0) S->done = 1;
}

// The whole loop of scan2MapOptimization (:1572-1610) in ONE cooperative launch (persistent_lm = 1): one block
// per SM, all co-resident. Every block keeps its own copy of the loop state (pose, sin/cos, matP, counters) in
// shared memory. Per iteration: phases 1 and 2 over the block's share of the queries -> block partial to global
// (double-buffered by iteration parity) -> grid barrier (one arrive counter, monotonically increasing within the
// scan) -> EVERY block adds all partials in the same fixed order and runs the same 6x6 solve, so all blocks hold
// bit-identical new poses without a publish / observe round trip, and nothing waits for a kernel boundary.
// Block 0 writes the final state back.  barrier: this scan's arrive counter (zero on entry); barrier_clear: the
// counter a later scan will use (ring of counters in the handle, re-armed 32 scans ahead).
template <int kEigenMode, bool kFresh>
__global__ void __launch_bounds__(LM_THREADS, 1) lm_persistent_kernel(const __grid_constant__ LMArgs A, float p0, float p1,
                                                                      float p2, float p3, float p4, float p5,
                                                                      unsigned* barrier, unsigned* barrier_clear,
                                                                      const __grid_constant__ MapBatch GB, int build_map,
                                                                      int build_grid, const Segment* segs,
                                                                      const BlockJob* jobs, int n_jobs,
                                                                      const __grid_constant__ IncJob IJ) {
  // dynamic shared memory (LM_PERSISTENT_SMEM bytes): [staging slots / map-build scratch | candidate cache]
  extern __shared__ __align__(16) unsigned char s_dyn[];
  float* const s_slot = reinterpret_cast<float*>(s_dyn);
  KnnCache C;
  C.cand = reinterpret_cast<float4*>(s_dyn + LM_THREADS * LM_SLOT * sizeof(float));
  C.center = C.cand + LM_QCAP * LM_CACHE_CAND;
  C.qpt = C.center + LM_QCAP;
  C.t1 = C.qpt + LM_QCAP;
  C.geo = C.t1 + LM_QCAP;
  C.cnt = reinterpret_cast<int*>(C.geo + 2 * LM_QCAP);
  C.prev = reinterpret_cast<uint32_t*>(C.cnt + LM_QCAP);
  __shared__ __align__(16) LMState s_st;
  __shared__ __align__(8) MinLambdaJob s_ml;
  const int tid = threadIdx.x;
  if (tid == 0) s_ml.rounds_left = 0;
  {  // prologue: the persistent members (matP, isDegenerate, ...) come from the handle's state
    static_assert(sizeof(LMState) % 4 == 0, "LMState is copied word by word");
    const uint32_t* src = reinterpret_cast<const uint32_t*>(A.st);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&s_st);
    if (tid < (int)(sizeof(LMState) / 4)) dst[tid] = __ldcg(src + tid);
  }
  unsigned epoch = 0;  // grid barriers passed so far
  LL_TL_BAR(0, 1);     // (timeline builds: kernel entry)
  static_assert(GI_SMEM_WORDS <= LM_THREADS * LM_SLOT && GI_SMEM_WORDS_REDUCE <= LM_THREADS * LM_SLOT && GI_THREADS == LM_THREADS,
                "the inline map build borrows the staging slots");
  static_assert((LM_THREADS * LM_SLOT * sizeof(float)) % 16 == 0, "the cache follows the staging slots, 16-byte aligned");
  // a4: the search grids of the local map, when ll_build_local_map left them to this launch (grid_inline.cuh)
  // a1-a3: the local map itself, when ll_build_local_map only prepared it (map_inline.cuh)
  bool cells_cleared = false;
  if (build_map) {
    if (blockIdx.x == 0)  // MapState of both maps (bounding box, counts): zero before the build
      inc_zero_map_states(GB, build_map == 2 && IJ.keep_fail != 0);
    coop_barrier(barrier, epoch);
    if (build_map == 2)  // a1-a3 incrementally: merge the added keyframes into the device-resident window (map_incremental.cuh)
      map_incremental_inline(IJ, GB, barrier, epoch, reinterpret_cast<uint32_t*>(s_dyn));
    else
      cells_cleared = map_build_inline(GB, segs, jobs, n_jobs, barrier, epoch, reinterpret_cast<uint32_t*>(s_slot));
  }
  if (build_grid) grid_build_inline(GB, barrier, epoch, reinterpret_cast<uint32_t*>(s_slot), cells_cleared);
  __syncthreads();
  static_assert(INC_SMEM_BYTES <= LM_PERSISTENT_SMEM && INC_MERGE_SMEM_WORDS * 4 <= LM_PERSISTENT_SMEM &&
                INC_REDUCE_SMEM_WORDS * 4 <= LM_PERSISTENT_SMEM, "the incremental map build borrows the dynamic shared memory");
  bool queries_missing = false;
  if (A.q_sync) {  // the query clouds were uploaded beside the map assembly: make sure they have landed
    if (blockIdx.x == 0 && tid == 0) {
      const long long t0 = clock64();
      unsigned arrived = 1u;
      while (ld_acquire_sys_u32(A.q_sync) != A.q_seq) {
        if (clock64() - t0 > 4000000000LL) { arrived = 0u; break; }  // ~2 s: the copy is never coming
      }
      A.q_sync[1] = arrived;
    }
    coop_barrier(barrier, epoch);
    queries_missing = __ldcg(A.q_sync + 1) == 0u;
  }
  if (tid < LM_QCAP) C.cnt[tid] = 0;  // no cached sets yet (after the map build, which may use this memory as scratch)
  if (A.g_guard) {  // no guards yet for this block's queries beyond its cache slots (only this block ever touches them)
    const QShare sh0 = lm_partition(A, (int)blockIdx.x, (int)gridDim.x);
    for (int L = LM_QCAP + tid; L < sh0.size(); L += LM_THREADS)
      A.g_guard[(size_t)sh0.qi(L, A.nq[0]) * LM_GUARD_STRIDE + LM_CACHE_CAND + 3] = make_float4(0.f, 0.f, 0.f, 0.f);  // cnt = 0
  }
  if (tid == 0) {
    const float pose6[6] = {p0, p1, p2, p3, p4, p5};
    lm_begin_state(&s_st, A.mst[0], A.mst[1], A.prm, pose6);
    if (queries_missing) {  // (grid-uniform) not a reference behaviour: reported as status 4, the scan is not registered
      s_st.status = 4;
      s_st.done = 1;
    }
  }
  __shared__ GridGeom s_geom[2];
  if (tid == 0) {
    GridGeom geom[2];
    lm_load_geom(A, geom);
    s_geom[0] = geom[0];
    s_geom[1] = geom[1];
  }
  const QShare sh = lm_partition(A, (int)blockIdx.x, (int)gridDim.x);
  for (unsigned it = 0;; ++it) {
    __syncthreads();
    if (s_st.done) break;  // block-uniform, and the same decision in every block
    const RotTrig rt = s_st.rtrig;
#ifdef LL_TIMELINE
    if (tid == 0 && blockIdx.x == TL_WATCH) g_tl_it = (int)it;
#endif
    LL_TL(it, 0);
    const float run_sum = lm_block_pass_cached<kEigenMode, kFresh>(A, s_st.trig, rt, s_geom, s_slot, C, sh, (int)it);
    LL_TL(it, 5);
    float* part = A.partials + (size_t)(it & 1u) * gridDim.x * 32;
    // (no __threadfence: the grid barrier's release-reduction is cumulative — it orders these writes, observed by thread 0
    //  through the barrier's bar.sync, before the arrival that the other blocks acquire; the map stages rely on the same)
    if (tid < 256 && (tid & 7) == 0) part[blockIdx.x * 32 + (tid >> 3)] = run_sum;  // thread 8 k holds sum k (k >= 29: zero)
    LL_TL(it, 6);
    LL_TL_GT(it, 0);
    coop_barrier(barrier, epoch);
    LL_TL_GT(it, 1);
    LL_TL(it, 7);
    reduce_partials_and_solve<kEigenMode>(part, gridDim.x, A.prm, &s_st, false, reinterpret_cast<double*>(s_slot), &s_ml);
    LL_TL(it, 9);
  }
  if (blockIdx.x == 0 && (tid >> 5) == 1 && s_ml.rounds_left > 0)  // the loop ended before the deferred rounds did
    lm_min_lambda_step(&s_ml, &s_st, tid & 31, MIN_EIG6_ROUNDS);
  __syncthreads();
#ifdef LL_TIMELINE
  __syncthreads();
  if (blockIdx.x == 0 && tid == 0) {
    printf("grid barriers of this launch (build_map %d, build_grid %d): microseconds from kernel entry to block 0's arrival / release\n", build_map, build_grid);
    for (unsigned e = 1; e <= epoch && e < 512; ++e)
      printf("   barrier %3u  arrive %8.2f  release %8.2f  (stage %6.2f)\n", e, (g_tl_bar[e][0] - g_tl_bar[0][1]) * 1e-3,
             (g_tl_bar[e][1] - g_tl_bar[0][1]) * 1e-3, (g_tl_bar[e][1] - g_tl_bar[e - 1][1]) * 1e-3);
    printf("sub-stage stamps of block 0 (microseconds from kernel entry):");
    for (int k = 0; k < 64; ++k)
      if (g_tl_sub[k] > g_tl_bar[0][1]) printf("  [%d] %.2f", k, (g_tl_sub[k] - g_tl_bar[0][1]) * 1e-3);
    printf("\n");
    for (int k = 0; k < 64; ++k) g_tl_sub[k] = 0ull;
    const int nb = min((int)gridDim.x, TL_BLOCKS);
    const double us = 1e-3 / 1.965;  // cycles -> us at the 1965 MHz the box runs at under load
    const char* names[9] = {"phase1 5-NN (to block barrier)", "block barrier after phase 1", "phase2 fit + Jacobian row", "shuffle reduce 29 sums",
                            "sum warp partials + barrier", "partial -> global + fence", "grid barrier (arrive..release)",
                            "sum all blocks' partials", "6x6 solve + pose update"};
    const int ea[9] = {0, 1, 2, 3, 4, 5, 6, 7, 8}, eb[9] = {1, 2, 3, 4, 5, 6, 7, 8, 9};
    printf("LM iteration timeline, %d blocks x %d threads; per phase: mean / max over blocks, microseconds at 1965 MHz\n", nb, LM_THREADS);
    for (int itx = 0; itx < TL_ITERS && itx < s_st.iter_num; ++itx) {
      double tot_mean = 0.0;
      printf(" iter %d\n", itx);
      for (int p = 0; p < 9; ++p) {
        double sum = 0.0, mx = 0.0;
        for (int b = 0; b < nb; ++b) {
          const double d = (double)(g_tl[itx][eb[p]][b] - g_tl[itx][ea[p]][b]) * us;
          sum += d;
          if (d > mx) mx = d;
        }
        tot_mean += sum / nb;
        printf("   %-34s %7.2f / %7.2f\n", names[p], sum / nb, mx);
      }
      unsigned long long first = ~0ull, last = 0, rel = 0;
      for (int b = 0; b < nb; ++b) {
        const unsigned long long t = g_tl_gt[itx][0][b];
        if (t < first) first = t;
        if (t > last) last = t;
        if (g_tl_gt[itx][1][b] > rel) rel = g_tl_gt[itx][1][b];
      }
      printf("   %-34s %7.2f   (globaltimer: first arrival -> last arrival %.2f us, last arrival -> last release %.2f us)\n",
             "iteration total (mean of blocks)", tot_mean, (last - first) * 1e-3, (rel - last) * 1e-3);
      printf("   queries: %u full searches (%u with > 128 candidates, %u rejected in a dense cell), cached: %u attempted, %u proven; %u went to step 2 (%u of them without a guard)\n",
             g_tl_cnt[itx][0], g_tl_cnt[itx][1], g_tl_cnt[itx][3], g_tl_cnt[itx][5], g_tl_cnt[itx][4], g_tl_cnt[itx][6], g_tl_cnt[itx][7]);
      for (int k = 0; k < 8; ++k) g_tl_cnt[itx][k] = 0;
      if (itx == 0 || itx == 1 || itx == 5) {
        printf("   block %d phase 1 per warp (us from its start): - / - / - / - / thread-level cached search done / full search done\n", TL_WATCH);
        for (int w = 0; w < 16; ++w) {
          const long long t0 = g_tl_warp[itx][0][w];
          printf("     warp %2d: %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f\n", w, (g_tl_warp[itx][6][w] - t0) * us, (g_tl_warp[itx][2][w] - t0) * us,
                 (g_tl_warp[itx][3][w] - t0) * us, (g_tl_warp[itx][4][w] - t0) * us, (g_tl_warp[itx][1][w] - t0) * us, (g_tl_warp[itx][5][w] - t0) * us);
        }
      }
    }
  }
#endif
  if (blockIdx.x == 0) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&s_st);
    uint32_t* dst = reinterpret_cast<uint32_t*>(A.st);
    if (tid < (int)(sizeof(LMState) / 4)) {
      dst[tid] = src[tid];
      if (A.st_host) reinterpret_cast<uint32_t*>(A.st_host)[tid] = src[tid];
    }
    if (tid == 0) *barrier_clear = 0u;
  }
}

// scan2MapOptimization prologue (:1573): the map-size guard and the loop state, one thread.
__global__ void lm_begin_kernel(LMState* S, const MapState* corner, const MapState* surf, LMParams prm,
                                float p0, float p1, float p2, float p3, float p4, float p5) {
  pdl_entry();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float pose6[6] = {p0, p1, p2, p3, p4, p5};
  lm_begin_state(S, corner, surf, prm, pose6);
}

// Stage-level a9 (ll_lm_step): rows from explicit (laserCloudOri, coeffSel) records, one thread per record,
// reduced exactly like phase 2 of lm_iter_kernel.
__global__ void __launch_bounds__(LM_THREADS) lm_rows_kernel(const float4* __restrict__ ori,
                                                             const float4* __restrict__ coeff, int n, LMParams prm,
                                                             LMState* S, float* partials) {
  __shared__ float s_wacc[LM_WARPS][32];
  __shared__ double s_red[LM_WARPS * 32];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const RotTrig rt = S->rtrig;
  float run_sum = 0.f;
  for (int base = blockIdx.x * LM_THREADS; base < n; base += gridDim.x * LM_THREADS) {
    const int i = base + tid;
    float sv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float one = 0.f;
    if (i < n) {
      const float4 p = ori[i];
      const float4 c = coeff[i];
      jacobian_row(rt, p.x, p.y, p.z, c.x, c.y, c.z, sv);
      sv[6] = -prm.step_size * c.w;
      sv[7] = fabsf(c.w);
      one = 1.f;
    }
#pragma unroll
    for (int k = 0; k < LM_NSUM; ++k) {
      float v = lm_term(sv, one, k);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_wacc[warp][k] = v;
    }
    __syncthreads();
    if (tid < LM_NSUM)
      for (int w = 0; w < LM_WARPS; ++w) run_sum += s_wacc[w][tid];
    __syncthreads();
  }
  if (tid < 32) {
    partials[blockIdx.x * 32 + tid] = tid < LM_NSUM ? run_sum : 0.f;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned int t = atomicAdd(&S->ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid == 0) S->ticket = 0;
  if (prm.fit.eigen_mode == 0) reduce_partials_and_solve<0>(partials, gridDim.x, prm, S, true, s_red);
  else reduce_partials_and_solve<1>(partials, gridDim.x, prm, S, true, s_red);
}

}  // namespace ll
