// Feasibility probe for halo-reuse convolutions: can a K-major SWIZZLE_128B UMMA A-operand descriptor start at an
// arbitrary ROW of a larger TMA-written tile (start address = tile + r*128 B, not 1024-B aligned), with the rows of
// the 128-row operand taken in groups of 8 at a stride SBO != 1024?  Tries descriptor base_offset = 0 and (r & 7).
//   D[m][n] = sum_k A[row(m)][k] * B[n][k],  row(m) = r + (m / 8) * G + (m % 8),  G = SBO / 128 rows
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../mmfn_b200/csrc/tc_common.cuh"

void mmfn_set_error(const char*, ...) {}
PFN_encodeTiled mmfn_get_encode_tiled() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return reinterpret_cast<PFN_encodeTiled>(p);
}
int mmfn_make_tmap_f32(CUtensorMap* out, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                       const uint32_t* box, const uint32_t* elem_strides, bool swizzle32, bool as_tf32) {
  PFN_encodeTiled enc = mmfn_get_encode_tiled();
  cuuint64_t gd[5], gs[5]; cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 1; i < rank; ++i) gs[i - 1] = strides_elems[i] * sizeof(float);
  CUresult r = enc(out, as_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<float*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

constexpr int AROWS = 256, N = 64;

__device__ __forceinline__ uint64_t desc_kmajor_ex(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((16 >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int r, int sbo, int use_bo) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                       // AROWS * 128 B
  uint8_t* sb = smem + AROWS * 128;         // N * 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + N * 128);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::mbar_init(done, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(slot, 64);
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    tc::mbar_expect_tx(bar, AROWS * 128 + N * 128);
    tc::tma_load_2d(sa, &tmA, bar, 0, 0);
    tc::tma_load_2d(sb, &tmB, bar, 0, 0);
    tc::mbar_wait(bar, 0);
    tc::tc_fence_after();
    const uint32_t a0 = tc::smem_u32(sa) + r * 128, b0 = tc::smem_u32(sb);
    const uint32_t idesc = tc::idesc_tf32(128, N, false, false);
    for (int k = 0; k < 4; ++k) {
      uint32_t aaddr = a0 + k * 32;
      uint64_t ad = desc_kmajor_ex(aaddr, sbo, use_bo ? ((aaddr >> 7) & 7) : 0);
      uint64_t bd = tc::smem_desc_kmajor(b0 + k * 32);
      tc::mma_tf32(tmem, ad, bd, idesc, k ? 1u : 0u);
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
  std::vector<float> A(AROWS * 32), Bm(N * 32);
  for (auto& x : A) x = (float)((rand() % 17) - 8) * 0.125f;       // exactly representable in TF32
  for (auto& x : Bm) x = (float)((rand() % 13) - 6) * 0.25f;
  float *dA, *dB, *dO;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, Bm.size() * 4); cudaMalloc(&dO, 128 * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bm.data(), Bm.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap ta, tb;
  { uint64_t d[2] = {32, AROWS}, s[2] = {1, 32}; uint32_t b[2] = {32, AROWS}; if (mmfn_make_tmap_f32(&ta, dA, 2, d, s, b, nullptr, false, true)) { printf("tmapA failed\n"); return 1; } }
  { uint64_t d[2] = {32, N}, s[2] = {1, 32}; uint32_t b[2] = {32, N}; if (mmfn_make_tmap_f32(&tb, dB, 2, d, s, b, nullptr, false, true)) { printf("tmapB failed\n"); return 1; } }
  const int smem = AROWS * 128 + N * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> O(128 * N);
  for (int sbo : {1280, 1152, 2304}) {
    const int G = sbo / 128;
    for (int use_bo = 0; use_bo < 1; ++use_bo) {
      printf("SBO=%d base_offset=%s : ", sbo, use_bo ? "(addr>>7)&7" : "0");
      for (int r : {0, 1, 2, 3, 7, 8, 9, 16, 17, 18}) {
        if (r + (15) * G + 8 > AROWS) { printf("r=%d:skip ", r); continue; }
        cudaMemset(dO, 0, O.size() * 4);
        probe<<<1, 128, smem>>>(ta, tb, dO, r, sbo, use_bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("r=%d:ERR(%s) ", r, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            int row = r + (m / 8) * G + (m % 8);
            double ref = 0;
            for (int k = 0; k < 32; ++k) ref += (double)A[row * 32 + k] * Bm[n * 32 + k];
            worst = fmax(worst, fabs(ref - O[m * N + n]));
          }
        printf("r=%d:%s ", r, worst < 1e-3 ? "OK" : "BAD");
      }
      printf("\n");
    }
  }
  return 0;
}
