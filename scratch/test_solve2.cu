#include <cstdio>
#include "../lego-loam-sr_b200/csrc/common.cuh"
#include "../lego-loam-sr_b200/csrc/fit.cuh"
#include "../lego-loam-sr_b200/csrc/map_build.cuh"
#include "lm_dbg.cuh"
using namespace ll;
__global__ void k(const double* sums, LMState* S, int mode) {
  LMParams prm; memset(&prm, 0, sizeof(prm));
  prm.fit.eigen_mode = mode; prm.min_residuals = 50; prm.eig_thresh = 100.f; prm.iter_count_thres = 10; prm.step_size = 1; prm.stop_thres = 0; prm.apply_update = 1;
  S->iter = 0;
  // direct colpiv test
  float AtA[36], Atb[6], x[6];
  int kk = 0;
  for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) { float v = (float)sums[kk++]; AtA[c*6+r] = v; AtA[r*6+c] = v; }
  for (int r = 0; r < 6; ++r) Atb[r] = (float)sums[21+r];
  colpiv_qr_solve_n(AtA, Atb, x, 6);
  printf("mode %d direct x: %g %g %g %g %g %g\n", mode, x[0], x[1], x[2], x[3], x[4], x[5]);
  lm_solve(sums, prm, S, true);
  printf("mode %d lm_solve x: %g %g %g %g %g %g deg %d minlam %g\n", mode, S->last_x[0], S->last_x[1], S->last_x[2], S->last_x[3], S->last_x[4], S->last_x[5], S->is_degenerate, S->min_lambda);
}
int main() {
  // SPD matrix with eigenvalues spread 1e3..1e6
  double M[6][6]; double B[6][6];
  unsigned s = 12345; auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0 - 0.5; };
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) B[i][j] = rnd();
  double d[6] = {1e3, 2e3, 3e3, 1e5, 4e5, 1.4e6};
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += B[k][i] * d[k] * B[k][j]; M[i][j] = a + (i == j ? 500.0 : 0); }
  double sums[29]; int kk = 0;
  for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) sums[kk++] = M[r][c];
  for (int r = 0; r < 6; ++r) sums[21 + r] = 100 * rnd();
  sums[27] = 10; sums[28] = 1000;
  double* ds; LMState* S;
  cudaMalloc(&ds, sizeof(sums)); cudaMalloc(&S, sizeof(LMState)); cudaMemset(S, 0, sizeof(LMState));
  cudaMemcpy(ds, sums, sizeof(sums), cudaMemcpyHostToDevice);
  for (int mode = 1; mode >= 0; --mode) { k<<<1, 1>>>(ds, S, mode); cudaDeviceSynchronize(); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
