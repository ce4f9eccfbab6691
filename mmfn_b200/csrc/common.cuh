// Shared device/host helpers for libmmfn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define MMFN_API extern "C" __attribute__((visibility("default")))

// Every C-ABI entry returns a cudaError_t-compatible int; 0 == success.
#define MMFN_BAD_ARG ((int)cudaErrorInvalidValue)

void mmfn_set_error(const char* fmt, ...);

#define MMFN_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      mmfn_set_error(__VA_ARGS__);                \
      return MMFN_BAD_ARG;                        \
    }                                             \
  } while (0)

static inline int mmfn_launch_status(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    mmfn_set_error("%s: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
  }
  return (int)e;
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
// Grid-stride launch size: enough CTAs for n items, capped at 8 waves of 148 SMs.
static inline int grid_1d(int64_t n, int threads) {
  int64_t b = ceil_div64(n, threads);
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}

// Counter-based RNG used by every dropout site: the same (seed, index) pair is
// re-evaluated in backward, so no mask tensor is ever stored.
__host__ __device__ __forceinline__ uint32_t mmfn_hash32(uint64_t seed, uint64_t idx) {
  uint64_t z = idx + seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}
// Device-resident RNG offset (mmfn_rng_bind): added to every dropout seed on the device so that a
// captured CUDA graph draws fresh masks on every replay.  One __constant__ pointer per translation unit.
static __constant__ const unsigned long long* c_mmfn_rng = nullptr;
#define MMFN_DEFINE_RNG_BINDER(tu)                                                     \
  int mmfn_bind_rng_##tu(const unsigned long long* p) {                                \
    return (int)cudaMemcpyToSymbol(c_mmfn_rng, &p, sizeof(p));                         \
  }

// keep-scale: 0 if dropped, 1/(1-p) if kept.  p == 0 -> always 1.
__host__ __device__ __forceinline__ float mmfn_dropout_scale(float p, uint64_t seed, uint64_t idx) {
  if (p <= 0.f) return 1.f;
#ifdef __CUDA_ARCH__
  const unsigned long long* rng = c_mmfn_rng;
  if (rng) seed += *rng;
#endif
  float u = (float)(mmfn_hash32(seed, idx) >> 8) * (1.0f / 16777216.0f);
  return (u >= p) ? 1.0f / (1.0f - p) : 0.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum for blockDim.x <= 1024; `sh` must hold 32 floats.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) v = warp_sum(v);
  if (threadIdx.x == 0) sh[0] = v;
  __syncthreads();
  v = sh[0];
  return v;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : -INFINITY;
  if (w == 0) v = warp_max(v);
  if (threadIdx.x == 0) sh[0] = v;
  __syncthreads();
  v = sh[0];
  return v;
}
