// Round-2 groundwork (NOT yet run on a GPU: written after the round-1 GPU budget was spent; compile-checked only).
// Probe for the bf16 tensor-core path (BASELINE configs[2]): tcgen05.mma kind::f16 with bf16 operands fed by TMA.
//   case K : A [128 x 64] and B [N x 64] K-major (rows of 64 bf16 = 128 bytes, SWIZZLE_128B), UMMA_K = 16 -> 4 MMAs
//   case MN: A stored transposed, [64 k-rows][128 m] bf16 (two TMA boxes of [64 k][64 m], SWIZZLE_128B), read as an
//            MN-major operand; the descriptor's LBO / SBO convention for 16-bit MN-major tiles is what the probe is
//            for: it tries (LBO = box stride, SBO = 1024) and (LBO = 1024, SBO = box stride) and prints which matches.
// Expected use next round:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o bf16_desc_probe
//                           bf16_desc_probe.cu -lcuda  &&  gpurun -- ./tools/exp/bf16_desc_probe
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

static int make_tmap_bf16(CUtensorMap* out, const __nv_bfloat16* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                          uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = mmfn_get_encode_tiled();
  cuuint64_t gd[2] = {inner, outer}, gs[1] = {pitch_elems * 2};
  cuuint32_t bx[2] = {box_inner, box_outer}, es[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

constexpr int M = 128, N = 64, K = 64;

// kind::f16 instruction descriptor: c_format F32 (1 << 4), a_format / b_format BF16 = 1 at bits 7 / 10
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: A K-major.  mode 1 / 2: A MN-major with the two LBO / SBO conventions.
__global__ void __launch_bounds__(128)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                       // 16 KB either way
  uint8_t* sb = smem + 16384;               // N * 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + N * 128);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::mbar_init(done, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(slot, 64);
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    tc::mbar_expect_tx(bar, 16384 + N * 128);
    if (mode == 0) tc::tma_load_2d(sa, &tmA, bar, 0, 0);                       // box {64 k, 128 m}
    else { tc::tma_load_2d(sa, &tmA, bar, 0, 0); tc::tma_load_2d(sa + 8192, &tmA, bar, 64, 0); }   // two boxes {64 m, 64 k}
    tc::tma_load_2d(sb, &tmB, bar, 0, 0);
    tc::mbar_wait(bar, 0);
    tc::tc_fence_after();
    const uint32_t a0 = tc::smem_u32(sa), b0 = tc::smem_u32(sb);
    const uint32_t idesc = idesc_bf16(M, N, mode != 0, false);
    for (int k = 0; k < K / 16; ++k) {
      uint64_t ad;
      if (mode == 0) ad = tc::smem_desc(a0 + k * 32, 16, 1024, 2);            // +16 bf16 = 32 bytes inside the swizzled row
      else if (mode == 1) ad = tc::smem_desc(a0 + k * 2048, 8192, 1024, 2);   // 16 k-rows = 2 atoms of 8 rows x 128 B
      else ad = tc::smem_desc(a0 + k * 2048, 1024, 8192, 2);
      uint64_t bd = tc::smem_desc(b0 + k * 32, 16, 1024, 2);
      mma_bf16(tmem, ad, bd, idesc, k ? 1u : 0u);
    }
    tc::mma_commit(done);
  }
  tc::mbar_wait(done, 0);
  tc::tc_fence_after();
  for (int c = 0; c < N / 32; ++c) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    for (int j = 0; j < 32; ++j) out[(size_t)threadIdx.x * N + c * 32 + j] = v[j];
  }
  tc::tc_fence_before(); __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<float> A(M * K), Bm(N * K);
  for (auto& x : A) x = (float)((rand() % 17) - 8) * 0.125f;        // exactly representable in bf16
  for (auto& x : Bm) x = (float)((rand() % 13) - 6) * 0.25f;
  std::vector<__nv_bfloat16> Ak(M * K), At(K * M), Bk(N * K);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { Ak[m * K + k] = __float2bfloat16(A[m * K + k]); At[k * M + m] = Ak[m * K + k]; }
  for (int i = 0; i < N * K; ++i) Bk[i] = __float2bfloat16(Bm[i]);
  __nv_bfloat16 *dAk, *dAt, *dB; float* dO;
  cudaMalloc(&dAk, Ak.size() * 2); cudaMalloc(&dAt, At.size() * 2); cudaMalloc(&dB, Bk.size() * 2); cudaMalloc(&dO, M * N * 4);
  cudaMemcpy(dAk, Ak.data(), Ak.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dAt, At.data(), At.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bk.data(), Bk.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tak, tat, tb;
  if (make_tmap_bf16(&tak, dAk, K, M, K, 64, 128) || make_tmap_bf16(&tat, dAt, M, K, M, 64, 64) || make_tmap_bf16(&tb, dB, K, N, K, 64, N)) {
    printf("tensor map encode failed\n"); return 1;
  }
  const int smem = 16384 + N * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> O(M * N);
  const char* names[3] = {"A K-major", "A MN-major (LBO = box stride 8192, SBO = 1024)", "A MN-major (LBO = 1024, SBO = box stride 8192)"};
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(dO, 0, O.size() * 4);
    probe<<<1, 128, smem>>>(mode == 0 ? tak : tat, tb, dO, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: ERROR %s\n", names[mode], cudaGetErrorString(e)); return 2; }
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * Bm[n * K + k];
        worst = fmax(worst, fabs(ref - O[m * N + n]));
      }
    printf("%-52s : %s (max abs err %.3g)\n", names[mode], worst < 1e-3 ? "OK" : "BAD", worst);
  }
  return 0;
}
