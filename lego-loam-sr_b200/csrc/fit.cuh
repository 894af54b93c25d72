// Per-point geometry of the scan-to-map path, one thread per point, everything in registers:
//   corner_fit  — body of cornerOptimization   (mapOptmization.cpp:1280-1373): centroid, 3x3
//                 covariance, eigen-decomposition, line test, point-to-line residual + gradient
//   surf_fit    — body of surfOptimization     (mapOptmization.cpp:1390-1438): 5x3 least squares
//                 plane (column-pivoted Householder QR, as Eigen's colPivHouseholderQr), validity,
//                 point-to-plane residual
//   jacobian_row — the row of matA built in LMOptimization (mapOptmization.cpp:1468-1500)
// Compiled with -fmad=false: float expressions round exactly as written (see common.cuh).
// Double is used only where the reference's expression promotes to double (the `0.1 *`, `0.9 *`
// literals and the threshold comparisons) and inside the closed-form eigen solve.
#pragma once
#include <float.h>
#include "common.cuh"

namespace ll {

struct FitParams {
  float max_sq_dist;
  float eig_ratio;
  double plane_tol;
  double weight_k;
  double min_weight;
  int eigen_mode;  // 0 as shipped, 1 upstream
};

// ---- symmetric 3x3 eigen solve -------------------------------------------------------------------
// Closed form (trigonometric solution of the characteristic cubic) evaluated in double, then the
// eigenvector of the largest eigenvalue from the best-conditioned cross product of two rows of
// (A - lambda I). Used for eigen_mode = upstream, where only the eigenvalues (ascending) and the
// principal direction matter and both are sign-free (SURVEY.md §7 "eigenvector-sign dependence").
__device__ __forceinline__ void eig3_closed_form(float a11, float a12, float a13, float a22, float a23, float a33,
                                                 float ev[3], float vmax[3]) {
  const double A11 = a11, A12 = a12, A13 = a13, A22 = a22, A23 = a23, A33 = a33;
  const double p1 = A12 * A12 + A13 * A13 + A23 * A23;
  const double q = (A11 + A22 + A33) / 3.0;
  const double b11 = A11 - q, b22 = A22 - q, b33 = A33 - q;
  const double p2 = b11 * b11 + b22 * b22 + b33 * b33 + 2.0 * p1;
  if (!(p2 > 0.0)) {  // multiple of the identity
    ev[0] = ev[1] = ev[2] = (float)q;
    vmax[0] = 1.f; vmax[1] = 0.f; vmax[2] = 0.f;
    return;
  }
  const double p = sqrt(p2 / 6.0);
  const double ip = 1.0 / p;
  const double c11 = b11 * ip, c22 = b22 * ip, c33 = b33 * ip, c12 = A12 * ip, c13 = A13 * ip, c23 = A23 * ip;
  double r = 0.5 * (c11 * (c22 * c33 - c23 * c23) - c12 * (c12 * c33 - c23 * c13) + c13 * (c12 * c23 - c22 * c13));
  r = fmin(1.0, fmax(-1.0, r));
  const double phi = acos(r) / 3.0;
  const double l2 = q + 2.0 * p * cos(phi);                                  // largest
  const double l0 = q + 2.0 * p * cos(phi + 2.0943951023931954923084289221863);  // smallest
  const double l1 = 3.0 * q - l2 - l0;
  ev[0] = (float)l0; ev[1] = (float)l1; ev[2] = (float)l2;
  // eigenvector of l2
  const double r0x = A11 - l2, r0y = A12, r0z = A13;
  const double r1x = A12, r1y = A22 - l2, r1z = A23;
  const double r2x = A13, r2y = A23, r2z = A33 - l2;
  double ax = r0y * r1z - r0z * r1y, ay = r0z * r1x - r0x * r1z, az = r0x * r1y - r0y * r1x;
  double bx = r0y * r2z - r0z * r2y, by = r0z * r2x - r0x * r2z, bz = r0x * r2y - r0y * r2x;
  double cx = r1y * r2z - r1z * r2y, cy = r1z * r2x - r1x * r2z, cz = r1x * r2y - r1y * r2x;
  double na = ax * ax + ay * ay + az * az, nb = bx * bx + by * by + bz * bz, nc = cx * cx + cy * cy + cz * cz;
  if (nb > na) { ax = bx; ay = by; az = bz; na = nb; }
  if (nc > na) { ax = cx; ay = cy; az = cz; na = nc; }
  if (na > 0.0) {
    const double inv = 1.0 / sqrt(na);
    vmax[0] = (float)(ax * inv); vmax[1] = (float)(ay * inv); vmax[2] = (float)(az * inv);
  } else {
    vmax[0] = 1.f; vmax[1] = 0.f; vmax[2] = 0.f;
  }
}

// ---- Eigen-style SelfAdjointEigenSolver<Matrix3f> (iterative path) ----------------------------------
// Needed only for eigen_mode = as shipped, where the reference reads ROW 0 of the eigenvector matrix
// (mapOptmization.cpp:1329-1334) and the result therefore depends on Eigen's eigenvector signs and
// ordering. Follows Eigen 3.3.7 SelfAdjointEigenSolver::compute: scale by max |coeff|, 3x3 closed
// tridiagonalisation, implicit symmetric QR with Wilkinson shift and Givens rotations accumulated
// on the right, ascending selection sort. V is row-major here: V[r][c] = component r of eigenvector c.
__device__ __forceinline__ void givens(float p, float q, float& c, float& s) {
  if (q == 0.f) {
    c = p < 0.f ? -1.f : 1.f;
    s = 0.f;
  } else if (p == 0.f) {
    c = 0.f;
    s = q < 0.f ? 1.f : -1.f;
  } else if (fabsf(p) > fabsf(q)) {
    float t = q / p;
    float u = sqrtf(1.f + t * t);
    if (p < 0.f) u = -u;
    c = 1.f / u;
    s = -t * c;
  } else {
    float t = p / q;
    float u = sqrtf(1.f + t * t);
    if (q < 0.f) u = -u;
    s = -1.f / u;
    c = -t * s;
  }
}

__device__ __forceinline__ float eig_hypot(float x, float y) {
  float ax = fabsf(x), ay = fabsf(y), p, qp;
  if (ax > ay) { p = ax; qp = ay / p; } else { p = ay; qp = ax / p; }
  if (p == 0.f) return 0.f;
  return p * sqrtf(1.f + qp * qp);
}

template <int N>
__device__ void tridiagonal_qr_step(float* diag, float* sub, int start, int end, float* Q /* N x N col-major */) {
  float td = (diag[end - 1] - diag[end]) * 0.5f;
  float e = sub[end - 1];
  float mu = diag[end];
  if (td == 0.f) {
    mu -= fabsf(e);
  } else if (e != 0.f) {
    const float e2 = e * e;
    const float h = eig_hypot(td, e);
    if (e2 == 0.f) mu -= e / ((td + (td > 0.f ? h : -h)) / e);
    else           mu -= e2 / (td + (td > 0.f ? h : -h));
  }
  float x = diag[start] - mu;
  float z = sub[start];
  for (int k = start; k < end && z != 0.f; ++k) {
    float c, s;
    givens(x, z, c, s);
    float sdk = s * diag[k] + c * sub[k];
    float dkp1 = s * sub[k] + c * diag[k + 1];
    diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
    diag[k + 1] = s * sdk + c * dkp1;
    sub[k] = c * sdk - s * dkp1;
    if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
    x = sub[k];
    if (k < end - 1) {
      z = -s * sub[k + 1];
      sub[k + 1] = c * sub[k + 1];
    }
    for (int i = 0; i < N; ++i) {
      float xi = Q[k * N + i], yi = Q[(k + 1) * N + i];
      Q[k * N + i] = c * xi - s * yi;
      Q[(k + 1) * N + i] = s * xi + c * yi;
    }
  }
}

// diag/sub in, eigenvalues (ascending, still scaled) and Q (col-major) out.
template <int N>
__device__ void tridiagonal_eig(float* diag, float* sub, float* Q) {
  int end = N - 1, start = 0, iter = 0;
  const float precision = 2.f * FLT_EPSILON;
  while (end > 0) {
    for (int i = start; i < end; ++i)
      if (fabsf(sub[i]) <= (fabsf(diag[i]) + fabsf(diag[i + 1])) * precision || fabsf(sub[i]) <= FLT_MIN) sub[i] = 0.f;
    while (end > 0 && sub[end - 1] == 0.f) end--;
    if (end <= 0) break;
    iter++;
    if (iter > 30 * N) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.f) start--;
    tridiagonal_qr_step<N>(diag, sub, start, end, Q);
  }
  for (int i = 0; i < N - 1; ++i) {
    int k = 0;
    float mn = diag[i];
    for (int j = 1; j < N - i; ++j)
      if (diag[i + j] < mn) { mn = diag[i + j]; k = j; }
    if (k > 0) {
      float t = diag[i]; diag[i] = diag[k + i]; diag[k + i] = t;
      for (int r = 0; r < N; ++r) { float u = Q[i * N + r]; Q[i * N + r] = Q[(k + i) * N + r]; Q[(k + i) * N + r] = u; }
    }
  }
}

// Full Eigen-style 3x3 solve. ev ascending; Q col-major (Q[c*3 + r]).
__device__ void eig3_eigen_style(float a11, float a12, float a13, float a22, float a23, float a33, float ev[3],
                                 float Q[9]) {
  // lower triangle, scaled
  float m00 = a11, m10 = a12, m20 = a13, m11 = a22, m21 = a23, m22 = a33;
  float scale = fmaxf(fmaxf(fmaxf(fabsf(m00), fabsf(m10)), fmaxf(fabsf(m20), fabsf(m11))), fmaxf(fabsf(m21), fabsf(m22)));
  if (scale == 0.f) scale = 1.f;
  m00 /= scale; m10 /= scale; m20 /= scale; m11 /= scale; m21 /= scale; m22 /= scale;
  float diag[3], sub[2];
  diag[0] = m00;
  float v1norm2 = m20 * m20;
  for (int i = 0; i < 9; ++i) Q[i] = 0.f;
  if (v1norm2 <= FLT_MIN) {
    diag[1] = m11; diag[2] = m22; sub[0] = m10; sub[1] = m21;
    Q[0] = Q[4] = Q[8] = 1.f;
  } else {
    float beta = sqrtf(m10 * m10 + v1norm2);
    float invBeta = 1.f / beta;
    float m01 = m10 * invBeta;
    float m02 = m20 * invBeta;
    float q = 2.f * m01 * m21 + m02 * (m22 - m11);
    diag[1] = m11 + m02 * q;
    diag[2] = m22 - m02 * q;
    sub[0] = beta;
    sub[1] = m21 - m01 * q;
    Q[0] = 1.f;
    Q[1 * 3 + 1] = m01; Q[2 * 3 + 1] = m02;   // (1,1), (1,2)
    Q[1 * 3 + 2] = m02; Q[2 * 3 + 2] = -m01;  // (2,1), (2,2)
  }
  tridiagonal_eig<3>(diag, sub, Q);
  ev[0] = diag[0] * scale; ev[1] = diag[1] * scale; ev[2] = diag[2] * scale;
}

// The fit is split where the query enters: the GEOMETRY (the two points of the line / the plane) depends on the five
// neighbours and their order only, so the LM loop keeps it from one iteration to the next while the neighbours stay the
// same (lm.cuh); the RESIDUAL is what the query adds. corner_fit / surf_fit are the two halves back to back: the same
// operations in the same order as before the split.
struct FitGeom {
  float g[6];  // corner: x1, y1, z1, x2, y2, z2 (:1329-1334)   surf: pa, pb, pc, pd (normalised, :1400-1409), -, -
  bool ok;     // corner: the line test (:1325) passed          surf: the plane is valid (:1411-1420)
};

template <int kEigenMode>
__device__ __forceinline__ void corner_geometry(const float nbx[5], const float nby[5], const float nbz[5], const FitParams& prm,
                                                FitGeom& G) {
  float cx = 0.f, cy = 0.f, cz = 0.f;
#pragma unroll
  for (int j = 0; j < 5; ++j) { cx += nbx[j]; cy += nby[j]; cz += nbz[j]; }
  cx /= 5.f; cy /= 5.f; cz /= 5.f;
  float a11 = 0.f, a12 = 0.f, a13 = 0.f, a22 = 0.f, a23 = 0.f, a33 = 0.f;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    float ax = nbx[j] - cx, ay = nby[j] - cy, az = nbz[j] - cz;
    a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
    a22 += ay * ay; a23 += ay * az; a33 += az * az;
  }
  a11 /= 5.f; a12 /= 5.f; a13 /= 5.f; a22 /= 5.f; a23 /= 5.f; a33 /= 5.f;
  float ev[3], vx, vy, vz;
  if constexpr (kEigenMode == 0) {
    float Q[9];
    eig3_eigen_style(a11, a12, a13, a22, a23, a33, ev, Q);
    vx = Q[0 * 3 + 0]; vy = Q[1 * 3 + 0]; vz = Q[2 * 3 + 0];  // ROW 0: (V(0,0), V(0,1), V(0,2)) as shipped
  } else {
    float v[3];
    eig3_closed_form(a11, a12, a13, a22, a23, a33, ev, v);
    vx = v[0]; vy = v[1]; vz = v[2];
  }
  G.ok = ev[2] > prm.eig_ratio * ev[1];
  G.g[0] = (float)((double)cx + 0.1 * (double)vx); G.g[1] = (float)((double)cy + 0.1 * (double)vy);
  G.g[2] = (float)((double)cz + 0.1 * (double)vz);
  G.g[3] = (float)((double)cx - 0.1 * (double)vx); G.g[4] = (float)((double)cy - 0.1 * (double)vy);
  G.g[5] = (float)((double)cz - 0.1 * (double)vz);
}

__device__ __forceinline__ bool corner_residual(const FitGeom& G, float x0, float y0, float z0, const FitParams& prm,
                                                float coeff[4]) {
  if (!G.ok) return false;
  const float x1 = G.g[0], y1 = G.g[1], z1 = G.g[2], x2 = G.g[3], y2 = G.g[4], z2 = G.g[5];
  float cxy = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
  float cxz = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
  float cyz = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
  float a012 = sqrtf(cxy * cxy + cxz * cxz + cyz * cyz);
  float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
  float la = ((y1 - y2) * cxy + (z1 - z2) * cxz) / a012 / l12;
  float lb = -((x1 - x2) * cxy - (z1 - z2) * cyz) / a012 / l12;
  float lc = -((x1 - x2) * cxz + (y1 - y2) * cyz) / a012 / l12;
  float ld2 = a012 / l12;
  float s = (float)(1.0 - prm.weight_k * fabs((double)ld2));
  coeff[0] = s * la;
  coeff[1] = s * lb;
  coeff[2] = s * lc;
  coeff[3] = s * ld2;
  return (double)s > prm.min_weight;
}

// ---- a7: corner residual ------------------------------------------------------------------------------
// nb = the five neighbours in ascending (d2, index) order; sel = the query in the map frame.
// Returns true when the residual is kept (s > min_weight); coeff = (s la, s lb, s lc, s ld2).
// kEigenMode: compile-time copy of FitParams::eigen_mode (the hot kernels are instantiated per mode so that the
// Eigen-style iterative solvers of the as-shipped mode stay out of the upstream-mode code).
template <int kEigenMode>
__device__ __forceinline__ bool corner_fit(const float nbx[5], const float nby[5], const float nbz[5], float x0,
                                           float y0, float z0, const FitParams& prm, float coeff[4]) {
  FitGeom G;
  corner_geometry<kEigenMode>(nbx, nby, nbz, prm, G);
  return corner_residual(G, x0, y0, z0, prm, coeff);
}

// ---- a8: surf residual ----------------------------------------------------------------------------------
// 5x3 column-pivoted Householder QR solve of A0 x = -1 (mapOptmization.cpp:1390-1398; matB0 = -1 at :278),
// in registers: a[c][r] = column c, row r. Mirrors Eigen 3.3.7 ColPivHouseholderQR::computeInPlace +
// _solve_impl step by step (norm down-dating included) so that rounding follows the same sequence.
__device__ __forceinline__ void qr53_solve_minus_one(float a[3][5], float x[3]) {
  float nu[3], nd[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 5; ++r) s += a[c][r] * a[c][r];
    nd[c] = sqrtf(s);
    nu[c] = nd[c];
  }
  const float maxNorm = fmaxf(fmaxf(nu[0], nu[1]), nu[2]);
  const float thr = (maxNorm * FLT_EPSILON) * (maxNorm * FLT_EPSILON) / 5.f;
  const float ndt = sqrtf(FLT_EPSILON);
  int nonzero = 3;
  int perm[3] = {0, 1, 2};
  float tau[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int big = k;
    float bigv = nu[k];
#pragma unroll
    for (int j = k + 1; j < 3; ++j)
      if (nu[j] > bigv) { bigv = nu[j]; big = j; }
    if (nonzero == 3 && bigv * bigv < thr * (float)(5 - k)) nonzero = k;
#pragma unroll
    for (int j = k + 1; j < 3; ++j) {
      if (big == j) {
#pragma unroll
        for (int r = 0; r < 5; ++r) { float t = a[k][r]; a[k][r] = a[j][r]; a[j][r] = t; }
        float t = nu[k]; nu[k] = nu[j]; nu[j] = t;
        t = nd[k]; nd[k] = nd[j]; nd[j] = t;
        int ti = perm[k]; perm[k] = perm[j]; perm[j] = ti;
      }
    }
    float tailSq = 0.f;
#pragma unroll
    for (int r = k + 1; r < 5; ++r) tailSq += a[k][r] * a[k][r];
    const float c0 = a[k][k];
    float beta;
    if (tailSq <= FLT_MIN) {
      tau[k] = 0.f;
      beta = c0;
#pragma unroll
      for (int r = k + 1; r < 5; ++r) a[k][r] = 0.f;
    } else {
      beta = sqrtf(c0 * c0 + tailSq);
      if (c0 >= 0.f) beta = -beta;
      const float denom = c0 - beta;
#pragma unroll
      for (int r = k + 1; r < 5; ++r) a[k][r] = a[k][r] / denom;
      tau[k] = (beta - c0) / beta;
    }
    a[k][k] = beta;
    if (tau[k] != 0.f) {
#pragma unroll
      for (int j = k + 1; j < 3; ++j) {
        float tmp = 0.f;
#pragma unroll
        for (int r = k + 1; r < 5; ++r) tmp += a[k][r] * a[j][r];
        tmp += a[j][k];
        a[j][k] -= tau[k] * tmp;
#pragma unroll
        for (int r = k + 1; r < 5; ++r) a[j][r] -= tau[k] * a[k][r] * tmp;
      }
    }
#pragma unroll
    for (int j = k + 1; j < 3; ++j) {
      if (nu[j] != 0.f) {
        float temp = fabsf(a[j][k]) / nu[j];
        temp = (1.f + temp) * (1.f - temp);
        temp = temp < 0.f ? 0.f : temp;
        float ratio = nu[j] / nd[j];
        float temp2 = temp * (ratio * ratio);
        if (temp2 <= ndt) {
          float s = 0.f;
#pragma unroll
          for (int r = k + 1; r < 5; ++r) s += a[j][r] * a[j][r];
          nd[j] = sqrtf(s);
          nu[j] = nd[j];
        } else {
          nu[j] *= sqrtf(temp);
        }
      }
    }
  }
  float c[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k < nonzero && tau[k] != 0.f) {
      float tmp = 0.f;
#pragma unroll
      for (int r = k + 1; r < 5; ++r) tmp += a[k][r] * c[r];
      tmp += c[k];
      c[k] -= tau[k] * tmp;
#pragma unroll
      for (int r = k + 1; r < 5; ++r) c[r] -= tau[k] * a[k][r] * tmp;
    }
  }
#pragma unroll
  for (int i = 2; i >= 0; --i) {
    if (i < nonzero) {
      float s = c[i];
#pragma unroll
      for (int j = i + 1; j < 3; ++j)
        if (j < nonzero) s -= a[j][i] * c[j];
      c[i] = s / a[i][i];
    }
  }
  x[0] = x[1] = x[2] = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (i < nonzero) {
#pragma unroll
      for (int t = 0; t < 3; ++t)
        if (perm[i] == t) x[t] = c[i];
    }
  }
}

__device__ __forceinline__ void surf_geometry(const float nbx[5], const float nby[5], const float nbz[5], const FitParams& prm,
                                              FitGeom& G) {
  float a[3][5];
#pragma unroll
  for (int j = 0; j < 5; ++j) { a[0][j] = nbx[j]; a[1][j] = nby[j]; a[2][j] = nbz[j]; }
  float X[3];
  qr53_solve_minus_one(a, X);
  float pa = X[0], pb = X[1], pc = X[2], pd = 1.f;
  float ps = sqrtf(pa * pa + pb * pb + pc * pc);
  pa /= ps; pb /= ps; pc /= ps; pd /= ps;
  bool valid = true;
#pragma unroll
  for (int j = 0; j < 5; ++j)
    if ((double)fabsf(pa * nbx[j] + pb * nby[j] + pc * nbz[j] + pd) > prm.plane_tol) valid = false;
  G.g[0] = pa; G.g[1] = pb; G.g[2] = pc; G.g[3] = pd; G.g[4] = 0.f; G.g[5] = 0.f;
  G.ok = valid;
}

__device__ __forceinline__ bool surf_residual(const FitGeom& G, float x0, float y0, float z0, const FitParams& prm, float coeff[4]) {
  if (!G.ok) return false;
  const float pa = G.g[0], pb = G.g[1], pc = G.g[2], pd = G.g[3];
  float pd2 = pa * x0 + pb * y0 + pc * z0 + pd;
  float nrm2 = x0 * x0 + y0 * y0 + z0 * z0;
  float s = (float)(1.0 - prm.weight_k * fabs((double)pd2) / sqrt(sqrt((double)nrm2)));
  coeff[0] = s * pa;
  coeff[1] = s * pb;
  coeff[2] = s * pc;
  coeff[3] = s * pd2;
  return (double)s > prm.min_weight;
}

__device__ __forceinline__ bool surf_fit(const float nbx[5], const float nby[5], const float nbz[5], float x0, float y0,
                                         float z0, const FitParams& prm, float coeff[4]) {
  FitGeom G;
  surf_geometry(nbx, nby, nbz, prm, G);
  return surf_residual(G, x0, y0, z0, prm, coeff);
}

// ---- a9: one row of matA and matB (mapOptmization.cpp:1464-1501) ---------------------------------------
struct RotTrig {
  float srx, crx, sry, cry, srz, crz;
};

__device__ __forceinline__ void jacobian_row(const RotTrig& t, float px, float py, float pz, float nx, float ny,
                                             float nz, float row[6]) {
  const float srx = t.srx, crx = t.crx, sry = t.sry, cry = t.cry, srz = t.srz, crz = t.crz;
  row[0] = (crx * sry * srz * px + crx * crz * sry * py - srx * sry * pz) * nx +
           (-srx * srz * px - crz * srx * py - crx * pz) * ny +
           (crx * cry * srz * px + crx * cry * crz * py - cry * srx * pz) * nz;
  row[1] = ((cry * srx * srz - crz * sry) * px + (sry * srz + cry * crz * srx) * py + crx * cry * pz) * nx +
           ((-cry * crz - srx * sry * srz) * px + (cry * srz - crz * srx * sry) * py - crx * sry * pz) * nz;
  row[2] = ((crz * srx * sry - cry * srz) * px + (-cry * crz - srx * sry * srz) * py) * nx +
           (crx * crz * px - crx * srz * py) * ny +
           ((sry * srz + cry * crz * srx) * px + (crz * sry - cry * srx * srz) * py) * nz;
  row[3] = nx;
  row[4] = ny;
  row[5] = nz;
}

}  // namespace ll
