#include <cstdio>
#include "../lego-loam-sr_b200/csrc/lm.cuh"
using namespace ll;
__global__ void k(const double* sums, LMState* S, int mode, int iter) {
  LMParams prm; memset(&prm, 0, sizeof(prm));
  prm.fit.eigen_mode = mode; prm.min_residuals = 50; prm.eig_thresh = 100.f; prm.iter_count_thres = 1000000; prm.step_size = 1; prm.stop_thres = 0; prm.apply_update = 1;
  S->iter = iter;
  lm_solve(sums, prm, S, false);
}
__global__ void k_trig(LMState* S) { set_trig(S); }
__global__ void k_empty(LMState* S) { if (S->done == 12345) S->iter = 1; }
int main() {
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
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto timeit = [&](const char* name, auto fn) {
    for (int i = 0; i < 5; ++i) fn();
    cudaEventRecord(a);
    for (int i = 0; i < 200; ++i) fn();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-28s %.2f us/launch\n", name, ms * 1000 / 200);
  };
  timeit("empty kernel", [&]() { k_empty<<<1, 1>>>(S); });
  timeit("set_trig", [&]() { k_trig<<<1, 1>>>(S); });
  timeit("lm_solve mode1 iter>0", [&]() { k<<<1, 1>>>(ds, S, 1, 1); });
  timeit("lm_solve mode1 iter0 (jacobi)", [&]() { k<<<1, 1>>>(ds, S, 1, 0); });
  timeit("lm_solve mode0 iter>0", [&]() { k<<<1, 1>>>(ds, S, 0, 1); });
  timeit("lm_solve mode0 iter0 (eigenQR)", [&]() { k<<<1, 1>>>(ds, S, 0, 0); });
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
