// The second consumer of the normal-equation kernel: FeatureAssociation::calculateTransformationSurf /
// Corner / calculateTransformation (featureAssociation.cpp:1846-2010, :2013-2143, :2145-2289). They have
// the shape of LMOptimization with a subset of the six unknowns: rows of matA restricted to the masked
// columns, matAtA / matAtB, colPivHouseholderQr solve (:1953-1956), and on iterCount == 0 the eigen /
// degeneracy projector with threshold 10 (:1961-1984). The caller owns matP / isDegenerate (in the
// reference matP is a function-local there, so a stale or uninitialised projector is the caller's to
// reproduce or avoid).
#pragma once
#include "lm.cuh"

namespace ll {

// Eigen's fixed-size 3x3 inverse (compute_inverse_size3_helper: cofactors / determinant). col-major.
__device__ inline void inverse3_cofactor(const float* M, float* inv) {
#define MM(r, c) M[(c) * 3 + (r)]
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return MM(i1, j1) * MM(i2, j2) - MM(i1, j2) * MM(i2, j1);
  };
  const float c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
  const float det = c00 * MM(0, 0) + c10 * MM(1, 0) + c20 * MM(2, 0);
  const float invdet = 1.f / det;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) inv[c * 3 + r] = cof(c, r) * invdet;
#undef MM
}

// The solve of calculateTransformation* once matAtA / matAtB are known (featureAssociation.cpp:1953-1990): column-pivoted
// QR, on iterCount == 0 the eigen / degeneracy projector, then matX = matP * matX2 while isDegenerate. One thread.
// sums: the m (m + 1) / 2 upper-triangular entries of AtA (row-major order), then the m entries of AtB. x: [m].
__device__ __noinline__ void neq_solve_small(const double* sums, int m, int iter, float eig_thresh, int eigen_mode, float* x,
                                             float* matP, int* is_degenerate) {
  float AtA[36], Atb[6];
  int k = 0;
  for (int a = 0; a < m; ++a)
    for (int c = a; c < m; ++c) {
      const float v = (float)sums[k++];
      AtA[c * m + a] = v;
      AtA[a * m + c] = v;
    }
  for (int a = 0; a < m; ++a) Atb[a] = (float)sums[k++];
  if (m == 3) colpiv_qr_solve_3(AtA, Atb, x);  // (same step sequence, unrolled: every iteration of fa_loop_kernel pays for it)
  else colpiv_qr_solve_n(AtA, Atb, x, m);
  if (iter == 0) {
    int deg = 0;
    float P[36];
    if (eigen_mode == 0 && (m == 3 || m == 6)) {
      float E[6], V[36], V2[36], Vinv[36];
      if (m == 3) {
        eig3_eigen_style(AtA[0], AtA[1], AtA[2], AtA[4], AtA[5], AtA[8], E, V);
      } else {
        eig6_eigen_style(AtA, E, V);
      }
      for (int i = 0; i < m * m; ++i) V2[i] = V[i];
      for (int i = m - 1; i >= 0; --i) {
        if (E[i] < eig_thresh) {
          for (int j = 0; j < m; ++j) V2[j * m + i] = 0.f;
          deg = 1;
        } else {
          break;
        }
      }
      if (m == 3) inverse3_cofactor(V, Vinv);
      else inverse_lu6(V, Vinv);
      for (int r = 0; r < m; ++r)
        for (int c = 0; c < m; ++c) {
          float a2 = 0.f;
          for (int kk = 0; kk < m; ++kk) a2 += Vinv[kk * m + r] * V2[c * m + kk];
          P[r * m + c] = a2;
        }
    } else {
      double w[6], V[36];
      jacobi_eig_n(AtA, m, w, V);
      int first_kept = 0;
      while (first_kept < m && (float)w[first_kept] < eig_thresh) ++first_kept;
      deg = first_kept > 0;
      for (int r = 0; r < m; ++r)
        for (int c = 0; c < m; ++c) {
          double a2 = 0.0;
          for (int kk = first_kept; kk < m; ++kk) a2 += V[kk * m + r] * V[kk * m + c];
          P[r * m + c] = (float)a2;
        }
    }
    *is_degenerate = deg;
    for (int i = 0; i < m * m; ++i) matP[i] = P[i];
  }
  if (*is_degenerate) {
    float x2[6];
    for (int r = 0; r < m; ++r) {
      float a2 = 0.f;
      for (int c = 0; c < m; ++c) a2 += matP[r * m + c] * x[c];
      x2[r] = a2;
    }
    for (int r = 0; r < m; ++r) x[r] = x2[r];
  }
}

__global__ void __launch_bounds__(LM_THREADS) normal_eq_kernel(const float* __restrict__ rows6,
                                                               const float* __restrict__ b, int n, int dof_mask,
                                                               int iter, float eig_thresh, int eigen_mode,
                                                               float* x_out, float* matP, int* is_degenerate) {
  __shared__ double s_part[LM_WARPS][28];
  __shared__ int s_cols[6];
  __shared__ int s_m;
  if (threadIdx.x == 0) {
    int m = 0;
    for (int k = 0; k < 6; ++k)
      if (dof_mask & (1 << k)) s_cols[m++] = k;
    s_m = m;
  }
  __syncthreads();
  const int m = s_m;
  const int nsum = m * (m + 1) / 2 + m;
  double acc[27];
  for (int k = 0; k < 27; ++k) acc[k] = 0.0;
  for (int i = threadIdx.x; i < n; i += LM_THREADS) {
    float r[6];
    for (int k = 0; k < m; ++k) r[k] = rows6[6 * i + s_cols[k]];
    const float bi = b[i];
    int k = 0;
    for (int a = 0; a < m; ++a)
      for (int c = a; c < m; ++c) acc[k++] += (double)(r[a] * r[c]);
    for (int a = 0; a < m; ++a) acc[k++] += (double)(r[a] * bi);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < nsum; ++k) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_part[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  double sums[27];
  for (int k = 0; k < nsum; ++k) {
    double t = 0.0;
    for (int w = 0; w < LM_WARPS; ++w) t += s_part[w][k];
    sums[k] = t;
  }
  float x[6];
  neq_solve_small(sums, m, iter, eig_thresh, eigen_mode, x, matP, is_degenerate);
  for (int r = 0; r < m; ++r) x_out[r] = x[r];
}

}  // namespace ll
