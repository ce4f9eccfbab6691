// Feasibility probe: programmatic dependent launch inside a stream-captured CUDA graph.
// Chain of N small dependent kernels (each: prologue smem init, then y[i] = x[i] + 1 over 64K floats).
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__global__ void step_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int pdl) {
  __shared__ float s[256];
  s[threadIdx.x] = threadIdx.x;          // "prologue"
  __syncthreads();
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = x[i] + 1.0f + 0.f * s[(threadIdx.x + 1) & 255];
}

static void launch(const float* x, float* y, int n, int pdl, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((n + 255) / 256); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, step_kernel, x, y, n, pdl);
  if (e != cudaSuccess) printf("launch error %s\n", cudaGetErrorString(e));
}

int main() {
  const int n = 1 << 16, N = 200;
  float *a, *b, *c;
  cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&c, n * 4);
  cudaStream_t st, st2; cudaStreamCreate(&st); cudaStreamCreate(&st2);
  for (int pdl = 0; pdl < 2; ++pdl) {
    for (int fork = 0; fork < 2; ++fork) {
      cudaMemset(a, 0, n * 4);
      cudaGraph_t g; cudaGraphExec_t ge;
      cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
      for (int i = 0; i < N; ++i) {
        launch(i % 2 ? b : a, i % 2 ? a : b, n, pdl, st);
        if (fork && i % 10 == 5) {      // fork/join through a second stream, like the aux streams
          cudaEvent_t e1, e2; cudaEventCreate(&e1); cudaEventCreate(&e2);
          cudaEventRecord(e1, st); cudaStreamWaitEvent(st2, e1);
          launch(i % 2 ? a : b, c, n, pdl, st2);
          cudaEventRecord(e2, st2); cudaStreamWaitEvent(st, e2);
        }
      }
      cudaError_t e = cudaStreamEndCapture(st, &g);
      if (e != cudaSuccess) { printf("capture failed: %s\n", cudaGetErrorString(e)); return 1; }
      e = cudaGraphInstantiate(&ge, g, 0);
      if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); return 1; }
      cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
      for (int w = 0; w < 3; ++w) cudaGraphLaunch(ge, st);
      cudaStreamSynchronize(st);
      cudaMemset(a, 0, n * 4);
      cudaEventRecord(t0, st);
      const int reps = 20;
      for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, st);
      cudaEventRecord(t1, st); cudaStreamSynchronize(st);
      float ms; cudaEventElapsedTime(&ms, t0, t1);
      std::vector<float> h(n);
      cudaMemcpy(h.data(), a, n * 4, cudaMemcpyDeviceToHost);
      bool ok = true;
      for (int i = 0; i < n; ++i) if (h[i] != (float)(N * reps)) { ok = false; printf("mismatch %d: %f\n", i, h[i]); break; }
      printf("pdl=%d fork=%d: %.2f us per kernel, result %s (%s)\n", pdl, fork, 1e3 * ms / reps / N, ok ? "OK" : "WRONG", cudaGetErrorString(cudaGetLastError()));
    }
    // plain stream (no graph)
    cudaMemset(a, 0, n * 4);
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
    for (int i = 0; i < N; ++i) launch(i % 2 ? b : a, i % 2 ? a : b, n, pdl, st);
    cudaStreamSynchronize(st);
    cudaEventRecord(t0, st);
    for (int i = 0; i < N; ++i) launch(i % 2 ? b : a, i % 2 ? a : b, n, pdl, st);
    cudaEventRecord(t1, st); cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, t0, t1);
    printf("pdl=%d stream: %.2f us per kernel\n", pdl, 1e3 * ms / N);
  }
  return 0;
}
