// Warp-parallel versions of the two small dense solves at the end of every LM iteration
// (LMOptimization, mapOptmization.cpp:1502-1531). They run in the LAST block of lm_iter_kernel, i.e. on
// the critical path between two iterations, so they are laid out for latency: lane j owns COLUMN j of
// the 6x6 matrix in registers, every row index is a compile-time constant after unrolling, and columns
// meet only through shuffles.
//
//   qr66_warp      matAtA.colPivHouseholderQr().solve(matAtB): the same step sequence as Eigen 3.3.7
//                  ColPivHouseholderQR (pivot search, Householder reflector, trailing update, norm
//                  down-dating, Q^T b, back substitution) — each column's arithmetic is exactly what the
//                  sequential algorithm does to that column, so the result is bit-identical to it.
//   jacobi6_warp   symmetric eigen-decomposition in double with the round-robin ("tournament") parallel
//                  ordering: 5 rounds of 3 disjoint rotations per sweep. Used for eigen_mode = upstream,
//                  where only the eigenvalues and the span of the kept eigenvectors matter.
#pragma once
#include <float.h>
#include "common.cuh"

namespace ll {

constexpr unsigned FULL = 0xffffffffu;

// All 32 lanes must call. AtA: row-major 6x6 in shared memory (symmetric), Atb: 6 floats.
// Returns this lane's solution component x[lane] on lanes 0..5 (0 elsewhere).
__device__ __forceinline__ float qr66_warp(const float* AtA, const float* Atb, int lane) {
  float a[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) a[r] = lane < 6 ? AtA[r * 6 + lane] : 0.f;
  float s0 = 0.f;
#pragma unroll
  for (int r = 0; r < 6; ++r) s0 += a[r] * a[r];
  float nD = sqrtf(s0), nU = nD;
  float maxNorm = nU;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) maxNorm = fmaxf(maxNorm, __shfl_xor_sync(FULL, maxNorm, o));
  maxNorm = __shfl_sync(FULL, maxNorm, 0);
  const float thr = (maxNorm * FLT_EPSILON) * (maxNorm * FLT_EPSILON) / 6.f;
  const float ndt = sqrtf(FLT_EPSILON);
  int nonzero = 6;
  int cur = lane < 6 ? lane : 100;  // current position of this lane's column in the pivoted order
  float hv[6][6];                   // hv[k][r], r > k: essential part of reflector k (same on every lane)
  float tau[6];
  int owner[6];                     // lane (= original column) sitting at pivot position k
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    // pivot: largest updated norm among positions k..5, lowest position on ties
    // (norms are >= +0, so their bit patterns order like the values: two redux.sync instead of a shuffle tree)
    const bool act = cur >= k && cur < 6;
    const uint32_t nbits = act ? __float_as_uint(nU) : 0u;
    const uint32_t mbits = __reduce_max_sync(FULL, nbits);
    const int bestp = (int)__reduce_min_sync(FULL, (act && nbits == mbits) ? (unsigned)cur : 99u);
    const float bestv = __uint_as_float(mbits);
    if (nonzero == 6 && bestv * bestv < thr * (float)(6 - k)) nonzero = k;
    if (cur == bestp) cur = k;
    else if (cur == k) cur = bestp;
    const int src = __ffs(__ballot_sync(FULL, cur == k)) - 1;
    owner[k] = src;
    // Householder reflector of the pivot column, rows k..5 (makeHouseholderInPlace)
    float tailSq = 0.f;
#pragma unroll
    for (int r = k + 1; r < 6; ++r) tailSq += a[r] * a[r];
    const float c0 = a[k];
    float beta, tk, ess[6];
    if (k == 5 || tailSq <= FLT_MIN) {
      tk = 0.f;
      beta = c0;
#pragma unroll
      for (int r = k + 1; r < 6; ++r) ess[r] = 0.f;
    } else {
      beta = sqrtf(c0 * c0 + tailSq);
      if (c0 >= 0.f) beta = -beta;
      const float denom = c0 - beta;
#pragma unroll
      for (int r = k + 1; r < 6; ++r) ess[r] = a[r] / denom;
      tk = (beta - c0) / beta;
    }
    tk = __shfl_sync(FULL, tk, src);
    beta = __shfl_sync(FULL, beta, src);
#pragma unroll
    for (int r = k + 1; r < 6; ++r) ess[r] = __shfl_sync(FULL, ess[r], src);
    tau[k] = tk;
#pragma unroll
    for (int r = k + 1; r < 6; ++r) hv[k][r] = ess[r];
    if (cur == k) {
      a[k] = beta;
#pragma unroll
      for (int r = k + 1; r < 6; ++r) a[r] = ess[r];
    } else if (cur > k && cur < 6) {
      if (k < 5 && tk != 0.f) {  // applyHouseholderOnTheLeft to a trailing column
        float tmp = 0.f;
#pragma unroll
        for (int r = k + 1; r < 6; ++r) tmp += ess[r] * a[r];
        tmp += a[k];
        a[k] -= tk * tmp;
#pragma unroll
        for (int r = k + 1; r < 6; ++r) a[r] -= tk * ess[r] * tmp;
      }
      if (nU != 0.f) {  // column-norm down-dating
        float temp = fabsf(a[k]) / nU;
        temp = (1.f + temp) * (1.f - temp);
        temp = temp < 0.f ? 0.f : temp;
        const float ratio = nU / nD;
        const float temp2 = temp * (ratio * ratio);
        if (temp2 <= ndt) {
          float s = 0.f;
#pragma unroll
          for (int r = k + 1; r < 6; ++r) s += a[r] * a[r];
          nD = sqrtf(s);
          nU = nD;
        } else {
          nU *= sqrtf(temp);
        }
      }
    }
  }
  // c = Q^T b (every lane computes the same six numbers)
  float c[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) c[r] = Atb[r];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (k < nonzero && tau[k] != 0.f) {
      float tmp = 0.f;
#pragma unroll
      for (int r = k + 1; r < 6; ++r) tmp += hv[k][r] * c[r];
      tmp += c[k];
      c[k] -= tau[k] * tmp;
#pragma unroll
      for (int r = k + 1; r < 6; ++r) c[r] -= tau[k] * hv[k][r] * tmp;
    }
  }
  // back substitution on R: R(i, j) = element i of the column at position j
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    float s = c[i];
#pragma unroll
    for (int j = i + 1; j < 6; ++j) {
      const float rij = __shfl_sync(FULL, a[i], owner[j]);
      if (j < nonzero) s -= rij * c[j];
    }
    const float rii = __shfl_sync(FULL, a[i], owner[i]);
    if (i < nonzero) c[i] = s / rii;
  }
  float x = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i)
    if (cur == i && i < nonzero) x = c[i];
  return lane < 6 ? x : 0.f;
}

// lm_solver = 1: the same 6x6 normal equations solved by an LDL^T factorisation in double (no pivoting: AtA is
// symmetric positive definite whenever the scan constrains all six degrees of freedom). One thread's worth of
// work, ~60 fused multiply-adds and six divisions, so every lane simply computes it redundantly and no
// broadcast is needed. The result is the solution of the float system to double accuracy; Eigen's float
// ColPivHouseholderQR (qr66_warp) differs from it by its own rounding error (<= cond(AtA) * 6e-8 relative).
// Returns false when a pivot is not safely positive (rank-deficient AtA): the caller then takes the
// rank-revealing QR, whose behaviour in that case (zeroed components) is part of the reference's semantics.
__device__ __forceinline__ bool ldlt6_solve(const float* AtA, const float* Atb, float x[6]) {
  double L[6][6], W[6][6], rd[6], y[6];
  double maxdiag = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) maxdiag = fmax(maxdiag, (double)AtA[i * 6 + i]);
  bool ok = maxdiag > 0.0 && maxdiag < 1.7976931348623157e308;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double sj = (double)AtA[j * 6 + j];
#pragma unroll
    for (int k = 0; k < j; ++k) sj -= L[j][k] * W[j][k];
    ok = ok && (sj > 1e-12 * maxdiag) && (sj < 1e37);
    // 1 / sj by a float reciprocal and two Newton steps in double (relative error ~1e-16, i.e. that of the division
    // itself): the six pivots are a serial chain, and a double division is ~10x the latency of this
    double r = (double)__frcp_rn((float)sj);
    r = r * (2.0 - sj * r);
    rd[j] = r * (2.0 - sj * r);
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double t = (double)AtA[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= L[i][k] * W[j][k];
      W[i][j] = t;           // L(i,j) * d(j)
      L[i][j] = t * rd[j];
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {  // L y = b
    double t = (double)Atb[i];
#pragma unroll
    for (int k = 0; k < i; ++k) t -= L[i][k] * y[k];
    y[i] = t;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) y[i] *= rd[i];  // D z = y
#pragma unroll
  for (int i = 5; i >= 0; --i) {  // L^T x = z
    double t = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) t -= L[k][i] * y[k];
    y[i] = t;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    x[i] = (float)y[i];
    ok = ok && isfinite(x[i]);
  }
  return ok;
}

// Number of eigenvalues of the symmetric 6x6 matrix A (row-major float) below sigma: the count of negative
// pivots of the LDL^T factorisation of A - sigma I (Sylvester's law of inertia), in double. One thread.
__device__ __forceinline__ int sturm_count6(const float* A, double sigma) {
  double L[6][6], W[6][6];
  int neg = 0;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double dj = (double)A[j * 6 + j] - sigma;
#pragma unroll
    for (int k = 0; k < j; ++k) dj -= L[j][k] * W[j][k];
    if (fabs(dj) < 1e-30) dj = -1e-30;  // sigma sits on an eigenvalue of a leading block: count it as below
    neg += dj < 0.0 ? 1 : 0;
    // 1 / dj to ~1e-14 (float reciprocal + one Newton step in double): only the SIGNS of the pivots are used, and a
    // full double division per pivot is most of the latency of a count
    double rd = (double)__frcp_rn((float)dj);
    rd = rd * (2.0 - dj * rd);
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double t = (double)A[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= L[i][k] * W[j][k];
      W[i][j] = t;
      L[i][j] = t * rd;
    }
  }
  return neg;
}

// Smallest eigenvalue of A and the number of eigenvalues below `thresh`, by 32-way multisection on Sturm counts:
// every lane tests its own shift, a ballot finds the sub-interval that contains the smallest eigenvalue, and
// six rounds narrow [-(eps) trace, min diagonal] by 33^6 ~ 1.3e9 — at float resolution of the result.
// This replaces the full Jacobi eigen-decomposition on the critical path of iteration 0 whenever no eigenvalue
// is below the threshold (the usual case); the eigenvectors are only needed for the degenerate projector.
// All 32 lanes must call; results are warp-uniform.
// The bracket the multisection starts from, and one of its rounds (32 shifts, one per lane; results warp-uniform).
__device__ __forceinline__ void min_eig6_bracket(const float* A, double& lo, double& hi) {
  double tr = 0.0;
  hi = 1.7976931348623157e308;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    tr += fabs((double)A[i * 6 + i]);
    hi = fmin(hi, (double)A[i * 6 + i]);  // lambda_min <= every diagonal entry
  }
  lo = -1e-6 * tr - 1e-30;                // AtA is positive semi-definite up to rounding
  hi += 1e-12 * tr + 1e-30;
}
__device__ __forceinline__ void min_eig6_round(const float* A, int lane, double& lo, double& hi) {
  const double step = (hi - lo) / 33.0;
  const double sigma = lo + step * (double)(lane + 1);
  const unsigned has = __ballot_sync(FULL, sturm_count6(A, sigma) >= 1);
  // counts are monotone in sigma: the first lane whose shift is above lambda_min
  const int first = has ? __ffs(has) - 1 : 32;
  const double new_lo = lo + step * (double)first;
  hi = first < 32 ? lo + step * (double)(first + 1) : hi;
  lo = new_lo;
}
constexpr int MIN_EIG6_ROUNDS = 6;

__device__ __forceinline__ void min_eig6_multisection(const float* A, double thresh, int lane, double& min_lambda,
                                                      int& n_below) {
  double lo, hi;
  min_eig6_bracket(A, lo, hi);
  n_below = sturm_count6(A, thresh);
#pragma unroll 1
  for (int round = 0; round < MIN_EIG6_ROUNDS; ++round) min_eig6_round(A, lane, lo, hi);
  min_lambda = 0.5 * (lo + hi);
}

// All 32 lanes must call. A: row-major symmetric 6x6 (float, shared memory).
// On return, lane l < 6 holds eigenvalue w and (kVectors) its eigenvector v[0..5] (unsorted; one eigenpair per
// lane). kVectors = false skips the accumulation of V (40 % of the shuffles): the common case only needs the
// eigenvalues (min_lambda, and "is anything below the threshold?"). Sweeps stop when the off-diagonal mass is
// below 1e-11 of the diagonal's, far beyond what the float outputs can see.
template <bool kVectors>
__device__ __forceinline__ void jacobi6_warp(const float* A, int lane, double& w, double v[6]) {
  double a[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    a[r] = lane < 6 ? (double)A[r * 6 + lane] : 0.0;
    v[r] = (r == lane) ? 1.0 : 0.0;
  }
  // round-robin schedule: 5 rounds x 3 disjoint pairs (p < q)
  constexpr int P[5][3] = {{0, 1, 2}, {0, 3, 1}, {0, 2, 1}, {0, 1, 4}, {0, 2, 3}};
  constexpr int Q[5][3] = {{5, 4, 3}, {4, 5, 2}, {3, 4, 5}, {2, 3, 5}, {1, 5, 4}};
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, diag = 0.0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      if (r < lane && lane < 6) off += a[r] * a[r];
      if (r == lane) diag += a[r] * a[r];
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      off += __shfl_xor_sync(FULL, off, o);
      diag += __shfl_xor_sync(FULL, diag, o);
    }
    off = __shfl_sync(FULL, off, 0);
    diag = __shfl_sync(FULL, diag, 0);
    if (off <= 1e-22 * diag || off == 0.0) break;
#pragma unroll
    for (int rd = 0; rd < 5; ++rd) {
      double cs[3], sn[3];
      int partner = lane;
      bool is_p = false;
      // The three rotations of a round are independent: every lane gathers the 2x2 block of ITS pair and the
      // (c, s) of all three pairs are computed at once (one divide + sqrt + rsqrt chain per round, not three).
      double app = 0.0, apq = 0.0, aqq = 0.0;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int p = P[rd][t], q = Q[rd][t];
        // lane q owns a_pq = a[p] and a_qq = a[q]; a_pp comes from lane p
        const double t_pp = __shfl_sync(FULL, a[p], p);
        const double t_pq = __shfl_sync(FULL, a[p], q);
        const double t_qq = __shfl_sync(FULL, a[q], q);
        if (lane == p || lane == q) { app = t_pp; apq = t_pq; aqq = t_qq; }
        if (lane == p) { partner = q; is_p = true; }
        if (lane == q) { partner = p; is_p = false; }
      }
      double myc = 1.0, mys = 0.0;
      if (apq != 0.0) {
        const double theta = (aqq - app) / (2.0 * apq);
        const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        myc = 1.0 / sqrt(tt * tt + 1.0);
        mys = tt * myc;
      }
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        cs[t] = __shfl_sync(FULL, myc, P[rd][t]);
        sn[t] = __shfl_sync(FULL, mys, P[rd][t]);
      }
      // A <- A J and V <- V J: column p' = c p - s q, column q' = s p + c q
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        const double oa = __shfl_sync(FULL, a[r], partner);
        if (lane < 6) a[r] = is_p ? (myc * a[r] - mys * oa) : (mys * oa + myc * a[r]);
        if (kVectors) {
          const double ov = __shfl_sync(FULL, v[r], partner);
          if (lane < 6) v[r] = is_p ? (myc * v[r] - mys * ov) : (mys * ov + myc * v[r]);
        }
      }
      // A <- J^T A: rows p, q of every column
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int p = P[rd][t], q = Q[rd][t];
        const double ap = a[p], aq = a[q];
        a[p] = cs[t] * ap - sn[t] * aq;
        a[q] = sn[t] * ap + cs[t] * aq;
      }
    }
  }
  w = 0.0;
#pragma unroll
  for (int r = 0; r < 6; ++r)
    if (r == lane) w = a[r];
}

}  // namespace ll
