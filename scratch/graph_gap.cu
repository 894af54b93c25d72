#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_small(int* p) { if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1; }
__global__ void k_pdl(int* p) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;"); if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1; }
int main() {
  int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int N = 24, REP = 200;
  float ms;
  for (int grid : {1, 148, 592}) {
    // stream launches
    for (int r = 0; r < 3; ++r) for (int i = 0; i < N; ++i) k_small<<<grid, 256, 0, s>>>(d);
    cudaStreamSynchronize(s);
    cudaEventRecord(a, s);
    for (int r = 0; r < REP; ++r) for (int i = 0; i < N; ++i) k_small<<<grid, 256, 0, s>>>(d);
    cudaEventRecord(b, s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("grid %4d stream : %.2f us/kernel\n", grid, ms * 1000 / (REP * N));
    // PDL launches
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = s;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEventRecord(a, s);
    for (int r = 0; r < REP; ++r) for (int i = 0; i < N; ++i) cudaLaunchKernelEx(&cfg, k_pdl, d);
    cudaEventRecord(b, s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("grid %4d PDL    : %.2f us/kernel\n", grid, ms * 1000 / (REP * N));
    // graph of N dependent kernels
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < N; ++i) k_small<<<grid, 256, 0, s>>>(d);
    cudaStreamEndCapture(s, &g);
    cudaGraphInstantiate(&ge, g, 0);
    for (int r = 0; r < 3; ++r) cudaGraphLaunch(ge, s);
    cudaStreamSynchronize(s);
    cudaEventRecord(a, s);
    for (int r = 0; r < REP; ++r) cudaGraphLaunch(ge, s);
    cudaEventRecord(b, s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("grid %4d graph  : %.2f us/kernel (%d-node graph, %.1f us per graph)\n", grid, ms * 1000 / (REP * N), N, ms * 1000 / REP);
    // single graph launch + sync latency (what a per-scan graph would pay)
    cudaEventRecord(a, s);
    for (int r = 0; r < 50; ++r) { cudaGraphLaunch(ge, s); cudaStreamSynchronize(s); }
    cudaEventRecord(b, s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("grid %4d graph+sync each: %.1f us per graph\n", grid, ms * 1000 / 50);
    cudaEventRecord(a, s);
    for (int r = 0; r < 50; ++r) { for (int i = 0; i < N; ++i) k_small<<<grid, 256, 0, s>>>(d); cudaStreamSynchronize(s); }
    cudaEventRecord(b, s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("grid %4d stream+sync each: %.1f us per %d kernels\n", grid, ms * 1000 / 50, N);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
