// Shared device/host helpers for libmmfn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
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
// re-evaluated in backward, so no mask tensor is ever stored.  ONE 64-bit hash serves FOUR consecutive
// elements (16 random bits each): vector kernels pay a quarter of the hash arithmetic per element.
__host__ __device__ __forceinline__ uint64_t mmfn_hash64(uint64_t seed, uint64_t idx) {
  // Two murmur3-style 32-bit finalisers over (idx, seed)-mixed words: ~20 32-bit integer instructions.  (The first
  // version ran three 64-bit multiplies = ~35 instructions and was 40 % of the fused attention's instruction stream.)
  const uint32_t a = (uint32_t)idx, b = (uint32_t)(idx >> 32), s0 = (uint32_t)seed, s1 = (uint32_t)(seed >> 32);
  uint32_t x = (a * 0x9E3779B1u + s0) ^ (b * 0x85EBCA77u + 0x7F4A7C15u);
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  uint32_t y = (a ^ 0x6C8E9CF5u) * 0xC2B2AE3Du + (s1 ^ (b * 0x27D4EB2Fu)) + x;
  y ^= y >> 15; y *= 0x2C1B3C6Du; y ^= y >> 12; y *= 0x297A2D39u; y ^= y >> 15;
  return ((uint64_t)y << 32) | (uint64_t)x;
}
// Device-resident RNG offset (mmfn_rng_bind): added to every dropout seed on the device so that a
// captured CUDA graph draws fresh masks on every replay.  One __constant__ pointer per translation unit.
static __constant__ const unsigned long long* c_mmfn_rng = nullptr;
#define MMFN_DEFINE_RNG_BINDER(tu)                                                     \
  int mmfn_bind_rng_##tu(const unsigned long long* p) {                                \
    return (int)cudaMemcpyToSymbol(c_mmfn_rng, &p, sizeof(p));                         \
  }

__host__ __device__ __forceinline__ uint64_t mmfn_drop_seed(uint64_t seed) {
#ifdef __CUDA_ARCH__
  const unsigned long long* rng = c_mmfn_rng;
  if (rng) seed += *rng;
#endif
  return seed;
}
// drop iff the element's 16 random bits fall below p * 65536
__host__ __device__ __forceinline__ uint32_t mmfn_drop_threshold(float p) { return (uint32_t)(p * 65536.0f); }

// keep-scale: 0 if dropped, 1/(1-p) if kept.  p == 0 -> always 1.
__host__ __device__ __forceinline__ float mmfn_dropout_scale(float p, uint64_t seed, uint64_t idx) {
  if (p <= 0.f) return 1.f;
  const uint64_t h = mmfn_hash64(mmfn_drop_seed(seed), idx >> 2);
  const uint32_t u = (uint32_t)(h >> (16 * (idx & 3))) & 0xFFFFu;
  return (u >= mmfn_drop_threshold(p)) ? 1.0f / (1.0f - p) : 0.f;
}
// the four keep-scales of elements idx .. idx + 3 (idx % 4 == 0): one hash
__host__ __device__ __forceinline__ void mmfn_dropout_scale4(float p, uint64_t seed, uint64_t idx, float* s) {
  if (p <= 0.f) { s[0] = s[1] = s[2] = s[3] = 1.f; return; }
  const uint64_t h = mmfn_hash64(mmfn_drop_seed(seed), idx >> 2);
  const uint32_t thr = mmfn_drop_threshold(p);
  const float keep = 1.0f / (1.0f - p);
  const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
  s[0] = ((lo & 0xFFFFu) >= thr) ? keep : 0.f;
  s[1] = ((lo >> 16) >= thr) ? keep : 0.f;
  s[2] = ((hi & 0xFFFFu) >= thr) ? keep : 0.f;
  s[3] = ((hi >> 16) >= thr) ? keep : 0.f;
}

// four fp32 values -> four bf16 (round to nearest even), packed for one 8-byte store; and back
__device__ __forceinline__ uint2 mmfn_pack_bf16x4(float a, float b, float c, float d) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 r;
  r.x = *reinterpret_cast<const uint32_t*>(&lo);
  r.y = *reinterpret_cast<const uint32_t*>(&hi);
  return r;
}
__device__ __forceinline__ float4 mmfn_unpack_bf16x4(uint2 r) {
  const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
  const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
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
