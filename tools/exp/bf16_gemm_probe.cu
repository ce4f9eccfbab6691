// Round-2 groundwork (NOT yet run on a GPU: written after the round-1 GPU budget was spent; compile-checked only).
// Stand-alone bf16 GEMM on the same warp-specialised tcgen05 pipeline as mmfn_b200/csrc/tc_kernel.cuh, to answer the
// question DESIGN.md section 9 starts from: what does C[M,N] = A[M,K] B[N,K]^T cost with 2-byte operands?
//   warp 0: TMA producer (128-byte rows = 64 bf16 per k-block), warp 1: MMA issuer (kind::f16, UMMA_K = 16 -> 4 MMAs per
//   k-block), warps 2-5: epilogue (tcgen05.ld, fp32 row stores -- deliberately simple, the probe times the main loop).
// Prints max |error| against a CPU reference on a sampled set of outputs and the achieved TFLOP/s for the transformer4
// MLP shape (4096 x 2048 x 512: the TF32 kernel reaches 357 TFLOP/s) and a larger one (8192 x 4096 x 1024).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o bf16_gemm_probe bf16_gemm_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "../../mmfn_b200/csrc/tc_common.cuh"

void mmfn_set_error(const char*, ...) {}
PFN_encodeTiled mmfn_get_encode_tiled() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return reinterpret_cast<PFN_encodeTiled>(p);
}
int mmfn_make_tmap_f32(CUtensorMap*, const float*, int, const uint64_t*, const uint64_t*, const uint32_t*, const uint32_t*, bool, bool) { return 1; }

static int make_tmap_bf16(CUtensorMap* out, const __nv_bfloat16* base, uint64_t inner, uint64_t outer, uint32_t box_outer) {
  PFN_encodeTiled enc = mmfn_get_encode_tiled();
  cuuint64_t gd[2] = {inner, outer}, gs[1] = {inner * 2};
  cuuint32_t bx[2] = {64, box_outer}, es[2] = {1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 1;
}

constexpr int TBM = 128, TBN = 128, STAGES = 4, THREADS = 192;
constexpr int A_BYTES = TBM * 128, B_BYTES = TBN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM = STAGES * STAGE_BYTES + 256 + 1024;

__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(THREADS, 2)
gemm_bf16(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C, int M, int N, int K) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN, nkb = K / 64;
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(slot, TBN);
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0) {
    if (tc::elect_one()) {
      int st = 0; uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        tc::mbar_wait(&empty[st], ph ^ 1);
        uint8_t* sa = smem + st * STAGE_BYTES;
        tc::mbar_expect_tx(&full[st], STAGE_BYTES);
        tc::tma_load_2d(sa, &tmA, &full[st], kb * 64, m0);
        tc::tma_load_2d(sa + A_BYTES, &tmB, &full[st], kb * 64, n0);
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const uint32_t idesc = idesc_bf16(TBM, TBN);
      int st = 0; uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        tc::mbar_wait(&full[st], ph);
        tc::tc_fence_after();
        const uint32_t sa = tc::smem_u32(smem + st * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)      // 16 bf16 = 32 bytes along the swizzled 128-byte row
          mma_bf16(tmem, tc::smem_desc_kmajor(sa + k * 32), tc::smem_desc_kmajor(sb + k * 32), idesc, (kb | k) ? 1u : 0u);
        tc::mma_commit(&empty[st]);
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
      tc::mma_commit(tmem_full);
    }
  } else {
    const int q = warp & 3;                                  // TMEM lane quarter of this warp (warps 2..5 -> 2,3,0,1)
    tc::mbar_wait(tmem_full, 0);
    tc::tc_fence_after();
    const int row = m0 + q * 32 + lane;
    for (int c = 0; c < TBN / 32; ++c) {
      float v[32];
      tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c * 32, v);
      if (row < M) {
        float4* dst = reinterpret_cast<float4*>(C + (size_t)row * N + n0 + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
  }
  tc::tc_fence_before(); __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem, TBN);
}

static int run(int M, int N, int K) {
  std::vector<__nv_bfloat16> A((size_t)M * K), B((size_t)N * K);
  std::vector<float> Af(A.size()), Bf(B.size());
  for (size_t i = 0; i < A.size(); ++i) { Af[i] = (float)((rand() % 33) - 16) / 16.f; A[i] = __float2bfloat16(Af[i]); }
  for (size_t i = 0; i < B.size(); ++i) { Bf[i] = (float)((rand() % 33) - 16) / 32.f; B[i] = __float2bfloat16(Bf[i]); }
  __nv_bfloat16 *dA, *dB; float* dC;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dC, (size_t)M * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ta, tb;
  if (make_tmap_bf16(&ta, dA, K, M, TBM) || make_tmap_bf16(&tb, dB, K, N, TBN)) { printf("tensor map encode failed\n"); return 1; }
  cudaFuncSetAttribute(gemm_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  dim3 grid(N / TBN, M / TBM);
  gemm_bf16<<<grid, THREADS, SMEM>>>(ta, tb, dC, M, N, K);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%dx%dx%d: ERROR %s\n", M, N, K, cudaGetErrorString(e)); return 2; }
  std::vector<float> C((size_t)M * N);
  cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int t = 0; t < 4096; ++t) {
    int m = rand() % M, n = rand() % N;
    double ref = 0;
    for (int k = 0; k < K; ++k) ref += (double)Af[(size_t)m * K + k] * Bf[(size_t)n * K + k];
    worst = fmax(worst, fabs(ref - C[(size_t)m * N + n]));
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) gemm_bf16<<<grid, THREADS, SMEM>>>(ta, tb, dC, M, N, K);
  cudaEventRecord(e0);
  const int reps = 50;
  for (int i = 0; i < reps; ++i) gemm_bf16<<<grid, THREADS, SMEM>>>(ta, tb, dC, M, N, K);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double us = 1e3 * ms / reps;
  printf("bf16 GEMM %5d x %5d x %5d : max abs err %.3g (%s), %.1f us, %.1f TFLOP/s\n", M, N, K, worst, worst < 1e-2 ? "OK" : "BAD",
         us, 2.0 * M * N * K / us / 1e6);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return 0;
}

int main() {
  if (run(4096, 2048, 512)) return 1;
  if (run(8192, 4096, 1024)) return 1;
  return 0;
}
