// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into, imported by or executed from the
// product path (lego-loam-sr_b200/, include/). Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may use it.
//
// Restatement of the three Eigen 3.3.x algorithms the reference's scan-to-map path calls.
// Eigen is NOT vendored in /root/reference and not installed here, so these follow Eigen's
// published algorithms (Eigen 3.3.7: SelfAdjointEigenSolver.h, Tridiagonalization.h,
// ColPivHouseholderQR.h, Householder.h, Jacobi.h, PartialPivLU.h) in scalar float arithmetic
// with sequential (non-vectorised) sums.  PARITY UNPINNED: the reference ships no test or
// golden vector for these call sites, and Eigen's SIMD reduction order is not reproducible
// without Eigen, so low-order bits (and eigenvector signs) are this file's convention.
//
// Reference call sites (relative to /root/reference/LeGO-LOAM/src/):
//   SelfAdjointEigenSolver<Matrix3f>          mapOptmization.cpp:1320-1323
//   SelfAdjointEigenSolver<Matrix<float,6,6>> mapOptmization.cpp:1512-1514
//   Matrix<float,5,3>::colPivHouseholderQr    mapOptmization.cpp:1398
//   Matrix<float,6,6>::colPivHouseholderQr    mapOptmization.cpp:1505
//   Matrix<float,6,6>::inverse                mapOptmization.cpp:1530
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace orc_eig {

// Column-major fixed-size matrix, like Eigen's default storage order.
template <int R, int C>
struct Mat {
  float d[R * C];
  float& operator()(int r, int c) { return d[c * R + r]; }
  float operator()(int r, int c) const { return d[c * R + r]; }
  void zero() { std::memset(d, 0, sizeof(d)); }
};

// --- Householder.h: makeHouseholder on v[0..n) -------------------------------------------
// On return v[1..n) holds the essential part, tau and beta as in Eigen.
inline void make_householder_inplace(float* v, int n, float& tau, float& beta) {
  float tailSqNorm = 0.f;
  for (int i = 1; i < n; ++i) tailSqNorm += v[i] * v[i];
  const float c0 = v[0];
  const float tol = FLT_MIN;
  if (n == 1 || tailSqNorm <= tol) {
    tau = 0.f;
    beta = c0;
    for (int i = 1; i < n; ++i) v[i] = 0.f;
  } else {
    beta = std::sqrt(c0 * c0 + tailSqNorm);
    if (c0 >= 0.f) beta = -beta;
    const float denom = c0 - beta;
    for (int i = 1; i < n; ++i) v[i] = v[i] / denom;
    tau = (beta - c0) / beta;
  }
}

// --- Jacobi.h: JacobiRotation::makeGivens (real case) ------------------------------------
inline void make_givens(float p, float q, float& c, float& s) {
  if (q == 0.f) {
    c = p < 0.f ? -1.f : 1.f;
    s = 0.f;
  } else if (p == 0.f) {
    c = 0.f;
    s = q < 0.f ? 1.f : -1.f;
  } else if (std::fabs(p) > std::fabs(q)) {
    float t = q / p;
    float u = std::sqrt(1.f + t * t);
    if (p < 0.f) u = -u;
    c = 1.f / u;
    s = -t * c;
  } else {
    float t = p / q;
    float u = std::sqrt(1.f + t * t);
    if (q < 0.f) u = -u;
    s = -1.f / u;
    c = -t * s;
  }
}

// numext::hypot of Eigen 3.3.x (MathFunctionsImpl.h)
inline float eig_hypot(float x, float y) {
  float ax = std::fabs(x), ay = std::fabs(y), p, qp;
  if (ax > ay) { p = ax; qp = ay / p; } else { p = ay; qp = ax / p; }
  if (p == 0.f) return 0.f;
  return p * std::sqrt(1.f + qp * qp);
}

// --- SelfAdjointEigenSolver.h: tridiagonal_qr_step ---------------------------------------
template <int N>
inline void tridiagonal_qr_step(float* diag, float* subdiag, int start, int end, Mat<N, N>* Q) {
  float td = (diag[end - 1] - diag[end]) * 0.5f;
  float e = subdiag[end - 1];
  float mu = diag[end];
  if (td == 0.f) {
    mu -= std::fabs(e);
  } else if (e != 0.f) {
    const float e2 = e * e;
    const float h = eig_hypot(td, e);
    if (e2 == 0.f) mu -= e / ((td + (td > 0.f ? h : -h)) / e);
    else           mu -= e2 / (td + (td > 0.f ? h : -h));
  }
  float x = diag[start] - mu;
  float z = subdiag[start];
  for (int k = start; k < end && z != 0.f; ++k) {
    float c, s;
    make_givens(x, z, c, s);
    float sdk = s * diag[k] + c * subdiag[k];
    float dkp1 = s * subdiag[k] + c * diag[k + 1];
    diag[k] = c * (c * diag[k] - s * subdiag[k]) - s * (c * subdiag[k] - s * diag[k + 1]);
    diag[k + 1] = s * sdk + c * dkp1;
    subdiag[k] = c * sdk - s * dkp1;
    if (k > start) subdiag[k - 1] = c * subdiag[k - 1] - s * z;
    x = subdiag[k];
    if (k < end - 1) {
      z = -s * subdiag[k + 1];
      subdiag[k + 1] = c * subdiag[k + 1];
    }
    if (Q) {  // q.applyOnTheRight(k, k+1, rot)
      for (int i = 0; i < N; ++i) {
        float xi = (*Q)(i, k), yi = (*Q)(i, k + 1);
        (*Q)(i, k) = c * xi - s * yi;
        (*Q)(i, k + 1) = s * xi + c * yi;
      }
    }
  }
}

// --- Tridiagonalization.h ------------------------------------------------------------------
// 3x3 real special case (tridiagonalization_inplace_selector<MatrixType,3,false>).
inline void tridiagonalize3(Mat<3, 3>& m, float diag[3], float sub[2]) {
  const float tol = FLT_MIN;
  diag[0] = m(0, 0);
  float v1norm2 = m(2, 0) * m(2, 0);
  if (v1norm2 <= tol) {
    diag[1] = m(1, 1);
    diag[2] = m(2, 2);
    sub[0] = m(1, 0);
    sub[1] = m(2, 1);
    m.zero();
    m(0, 0) = m(1, 1) = m(2, 2) = 1.f;
  } else {
    float beta = std::sqrt(m(1, 0) * m(1, 0) + v1norm2);
    float invBeta = 1.f / beta;
    float m01 = m(1, 0) * invBeta;
    float m02 = m(2, 0) * invBeta;
    float q = 2.f * m01 * m(2, 1) + m02 * (m(2, 2) - m(1, 1));
    diag[1] = m(1, 1) + m02 * q;
    diag[2] = m(2, 2) - m02 * q;
    sub[0] = beta;
    sub[1] = m(2, 1) - m01 * q;
    m.zero();
    m(0, 0) = 1.f;
    m(1, 1) = m01;  m(1, 2) = m02;
    m(2, 1) = m02;  m(2, 2) = -m01;
  }
}

// General real case: Householder reduction on the lower triangle, then Q = H_0 ... H_{n-2}.
template <int N>
inline void tridiagonalize_general(Mat<N, N>& A, float diag[N], float sub[N - 1]) {
  float hCoeffs[N - 1];
  float p[N];
  for (int i = 0; i < N - 1; ++i) {
    const int rs = N - i - 1;
    float col[N];
    for (int r = 0; r < rs; ++r) col[r] = A(i + 1 + r, i);
    float h, beta;
    make_householder_inplace(col, rs, h, beta);
    col[0] = 1.f;
    // p = (A_sub.selfadjointView<Lower>() * (h * v))
    for (int r = 0; r < rs; ++r) {
      float acc = 0.f;
      for (int c = 0; c < rs; ++c) {
        float a = (r >= c) ? A(i + 1 + r, i + 1 + c) : A(i + 1 + c, i + 1 + r);
        acc += a * (h * col[c]);
      }
      p[r] = acc;
    }
    // p += (h * -0.5 * (p . v)) * v
    float dot = 0.f;
    for (int r = 0; r < rs; ++r) dot += p[r] * col[r];
    float alpha = h * -0.5f * dot;
    for (int r = 0; r < rs; ++r) p[r] += alpha * col[r];
    // A_sub.selfadjointView<Lower>().rankUpdate(v, p, -1)
    for (int c = 0; c < rs; ++c)
      for (int r = c; r < rs; ++r)
        A(i + 1 + r, i + 1 + c) -= (col[r] * p[c] + p[r] * col[c]);
    A(i + 1, i) = beta;
    for (int r = 1; r < rs; ++r) A(i + 1 + r, i) = col[r];
    hCoeffs[i] = h;
  }
  for (int i = 0; i < N; ++i) diag[i] = A(i, i);
  for (int i = 0; i < N - 1; ++i) sub[i] = A(i + 1, i);
  // HouseholderSequence(A, hCoeffs).setLength(N-1).setShift(1) evaluated to a dense Q
  Mat<N, N> Q;
  Q.zero();
  for (int i = 0; i < N; ++i) Q(i, i) = 1.f;
  for (int k = N - 2; k >= 0; --k) {
    const int cs = N - k - 1;  // corner size
    const int off = N - cs;    // = k + 1
    const float tau = hCoeffs[k];
    if (cs == 1) {
      Q(off, off) *= (1.f - tau);
    } else if (tau != 0.f) {
      // essential = A.col(k).tail(cs - 1) starting at row k + 2
      for (int c = 0; c < cs; ++c) {
        float tmp = 0.f;
        for (int r = 1; r < cs; ++r) tmp += A(k + 1 + r, k) * Q(off + r, off + c);
        tmp += Q(off, off + c);
        Q(off, off + c) -= tau * tmp;
        for (int r = 1; r < cs; ++r) Q(off + r, off + c) -= tau * A(k + 1 + r, k) * tmp;
      }
    }
  }
  A = Q;
}

// --- SelfAdjointEigenSolver<Matrix<float,N,N>>::compute(m, ComputeEigenvectors) -----------
// Input: symmetric matrix (only the lower triangle is read). Output: eigenvalues ascending,
// eigenvectors as COLUMNS of V (V(r,c)). No sign normalisation, as in Eigen.
template <int N>
inline void self_adjoint_eig(const Mat<N, N>& Min, float evals[N], Mat<N, N>& V) {
  Mat<N, N> mat;
  mat.zero();
  float scale = 0.f;
  for (int c = 0; c < N; ++c)
    for (int r = c; r < N; ++r) {
      mat(r, c) = Min(r, c);
      scale = std::max(scale, std::fabs(Min(r, c)));
    }
  if (scale == 0.f) scale = 1.f;
  for (int c = 0; c < N; ++c)
    for (int r = c; r < N; ++r) mat(r, c) /= scale;
  float diag[N], sub[N > 1 ? N - 1 : 1];
  if constexpr (N == 3) {
    tridiagonalize3(reinterpret_cast<Mat<3, 3>&>(mat), diag, sub);
  } else {
    tridiagonalize_general<N>(mat, diag, sub);
  }
  // computeFromTridiagonal_impl
  int end = N - 1, start = 0, iter = 0;
  const int maxIterations = 30;
  const float considerAsZero = FLT_MIN;
  const float precision = 2.f * FLT_EPSILON;
  while (end > 0) {
    for (int i = start; i < end; ++i)
      if (std::fabs(sub[i]) <= (std::fabs(diag[i]) + std::fabs(diag[i + 1])) * precision ||
          std::fabs(sub[i]) <= considerAsZero)
        sub[i] = 0.f;
    while (end > 0 && sub[end - 1] == 0.f) end--;
    if (end <= 0) break;
    iter++;
    if (iter > maxIterations * N) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.f) start--;
    tridiagonal_qr_step<N>(diag, sub, start, end, &mat);
  }
  // selection sort ascending, swapping columns
  for (int i = 0; i < N - 1; ++i) {
    int k = 0;
    float mn = diag[i];
    for (int j = 1; j < N - i; ++j)
      if (diag[i + j] < mn) { mn = diag[i + j]; k = j; }
    if (k > 0) {
      std::swap(diag[i], diag[k + i]);
      for (int r = 0; r < N; ++r) std::swap(mat(r, i), mat(r, k + i));
    }
  }
  for (int i = 0; i < N; ++i) evals[i] = diag[i] * scale;
  V = mat;
}

// --- ColPivHouseholderQR<Matrix<float,R,C>>::compute + solve(b) ----------------------------
template <int R, int C>
inline void colpiv_qr_solve(const Mat<R, C>& Ain, const float b[R], float x[C]) {
  Mat<R, C> qr = Ain;
  constexpr int size = R < C ? R : C;
  float hCoeffs[size];
  int transp[C];
  float normsUpdated[C], normsDirect[C];
  for (int k = 0; k < C; ++k) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += qr(r, k) * qr(r, k);
    normsDirect[k] = std::sqrt(s);
    normsUpdated[k] = normsDirect[k];
  }
  float maxNorm = normsUpdated[0];
  for (int k = 1; k < C; ++k) maxNorm = std::max(maxNorm, normsUpdated[k]);
  const float thr = (maxNorm * FLT_EPSILON) * (maxNorm * FLT_EPSILON) / float(R);
  const float norm_downdate_threshold = std::sqrt(FLT_EPSILON);
  int nonzero_pivots = size;
  for (int k = 0; k < size; ++k) {
    int big = k;
    float bigv = normsUpdated[k];
    for (int j = k + 1; j < C; ++j)
      if (normsUpdated[j] > bigv) { bigv = normsUpdated[j]; big = j; }
    float big_sq = bigv * bigv;
    if (nonzero_pivots == size && big_sq < thr * float(R - k)) nonzero_pivots = k;
    transp[k] = big;
    if (k != big) {
      for (int r = 0; r < R; ++r) std::swap(qr(r, k), qr(r, big));
      std::swap(normsUpdated[k], normsUpdated[big]);
      std::swap(normsDirect[k], normsDirect[big]);
    }
    float col[R];
    const int n = R - k;
    for (int r = 0; r < n; ++r) col[r] = qr(k + r, k);
    float tau, beta;
    make_householder_inplace(col, n, tau, beta);
    for (int r = 1; r < n; ++r) qr(k + r, k) = col[r];
    qr(k, k) = beta;
    hCoeffs[k] = tau;
    // bottomRightCorner(R-k, C-k-1).applyHouseholderOnTheLeft(essential, tau)
    if (n == 1) {
      for (int c = k + 1; c < C; ++c) qr(k, c) *= (1.f - tau);
    } else if (tau != 0.f) {
      for (int c = k + 1; c < C; ++c) {
        float tmp = 0.f;
        for (int r = 1; r < n; ++r) tmp += col[r] * qr(k + r, c);
        tmp += qr(k, c);
        qr(k, c) -= tau * tmp;
        for (int r = 1; r < n; ++r) qr(k + r, c) -= tau * col[r] * tmp;
      }
    }
    for (int j = k + 1; j < C; ++j) {
      if (normsUpdated[j] != 0.f) {
        float temp = std::fabs(qr(k, j)) / normsUpdated[j];
        temp = (1.f + temp) * (1.f - temp);
        temp = temp < 0.f ? 0.f : temp;
        float ratio = normsUpdated[j] / normsDirect[j];
        float temp2 = temp * (ratio * ratio);
        if (temp2 <= norm_downdate_threshold) {
          float s = 0.f;
          for (int r = k + 1; r < R; ++r) s += qr(r, j) * qr(r, j);
          normsDirect[j] = std::sqrt(s);
          normsUpdated[j] = normsDirect[j];
        } else {
          normsUpdated[j] *= std::sqrt(temp);
        }
      }
    }
  }
  // permutation indices from the transpositions
  int perm[C];
  for (int i = 0; i < C; ++i) perm[i] = i;
  for (int k = 0; k < size; ++k) std::swap(perm[k], perm[transp[k]]);
  // solve
  for (int i = 0; i < C; ++i) x[i] = 0.f;
  if (nonzero_pivots == 0) return;
  float c[R];
  for (int r = 0; r < R; ++r) c[r] = b[r];
  for (int k = 0; k < nonzero_pivots; ++k) {  // c = H_k c, k ascending (Q^T b)
    const int n = R - k;
    const float tau = hCoeffs[k];
    if (n == 1) {
      c[k] *= (1.f - tau);
    } else if (tau != 0.f) {
      float tmp = 0.f;
      for (int r = 1; r < n; ++r) tmp += qr(k + r, k) * c[k + r];
      tmp += c[k];
      c[k] -= tau * tmp;
      for (int r = 1; r < n; ++r) c[k + r] -= tau * qr(k + r, k) * tmp;
    }
  }
  for (int i = nonzero_pivots - 1; i >= 0; --i) {  // upper-triangular back substitution
    float s = c[i];
    for (int j = i + 1; j < nonzero_pivots; ++j) s -= qr(i, j) * c[j];
    c[i] = s / qr(i, i);
  }
  for (int i = 0; i < nonzero_pivots; ++i) x[perm[i]] = c[i];
}

// --- PartialPivLU based inverse (Matrix<float,6,6>::inverse()) ----------------------------
template <int N>
inline void inverse_lu(const Mat<N, N>& Ain, Mat<N, N>& inv) {
  Mat<N, N> lu = Ain;
  int piv[N];
  for (int k = 0; k < N; ++k) {
    int p = k;
    float best = std::fabs(lu(k, k));
    for (int r = k + 1; r < N; ++r)
      if (std::fabs(lu(r, k)) > best) { best = std::fabs(lu(r, k)); p = r; }
    piv[k] = p;
    if (p != k)
      for (int c = 0; c < N; ++c) std::swap(lu(k, c), lu(p, c));
    if (lu(k, k) != 0.f) {
      for (int r = k + 1; r < N; ++r) lu(r, k) /= lu(k, k);
      for (int c = k + 1; c < N; ++c)
        for (int r = k + 1; r < N; ++r) lu(r, c) -= lu(r, k) * lu(k, c);
    }
  }
  for (int col = 0; col < N; ++col) {
    float y[N];
    for (int r = 0; r < N; ++r) y[r] = (r == col) ? 1.f : 0.f;
    for (int k = 0; k < N; ++k) std::swap(y[k], y[piv[k]]);
    for (int r = 0; r < N; ++r)
      for (int c = 0; c < r; ++c) y[r] -= lu(r, c) * y[c];
    for (int r = N - 1; r >= 0; --r) {
      for (int c = r + 1; c < N; ++c) y[r] -= lu(r, c) * y[c];
      y[r] /= lu(r, r);
    }
    for (int r = 0; r < N; ++r) inv(r, col) = y[r];
  }
}

// --- Matrix<float,3,3>::inverse(): Eigen's fixed-size path (compute_inverse_size3_helper) = cofactors / det ---
inline void inverse3_cofactor(const Mat<3, 3>& M, Mat<3, 3>& inv) {
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return M(i1, j1) * M(i2, j2) - M(i1, j2) * M(i2, j1);
  };
  const float c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
  const float det = c00 * M(0, 0) + c10 * M(1, 0) + c20 * M(2, 0);
  const float invdet = 1.f / det;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) inv(r, c) = cof(c, r) * invdet;
}

}  // namespace orc_eig
