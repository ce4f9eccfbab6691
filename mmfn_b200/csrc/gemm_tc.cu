// TF32 tensor-core GEMM for the dense Linear layers of the fusion transformers and heads
// (model_rad.py:82-89, :120-125 and their backward passes): tcgen05.mma with the accumulator in
// TMEM, operands streamed by TMA into 128B-swizzled shared memory through an mbarrier ring.
//
//   C[M,N] (+)= alpha * A * B^T, fp32 in HBM, read as TF32 (TMA rounds), fp32 accumulate.
//
// Either operand may be K-major (row = M/N index, reduction contiguous: activations, weights in
// forward) or MN-major (reduction is the slow dimension: W in dgrad, dY/X in wgrad), so forward,
// data-gradient and weight-gradient GEMMs all read the same row-major tensors with no transposes.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane each),
// warps 2-5 = epilogue (TMEM -> registers -> fused bias/ReLU/mask/dropout/residual -> HBM).
#include <stdio.h>
#include "tc_common.cuh"

namespace {

constexpr int TBM = 128;          // tile rows (UMMA M)
constexpr int TBK = 32;           // fp32 elements per k-block = one 128-byte swizzle span
constexpr int UMMA_K = 8;         // tf32
constexpr int TC_THREADS = 192;

struct TcEpilogue {
  float* C;
  int64_t ldc;
  const float* bias;
  const float* res;
  const float* mask;
  float alpha;
  int act;
  int accum;          // 0 store, 1 +=, 2 atomicAdd
  float drop_p;
  uint64_t drop_seed;
};

template <int TBN, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = TBM * 128;
  static constexpr int B_BYTES = TBN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;   // + alignment slack
};

template <int TBN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               int M, int N, int K, int a_mn, int b_mn, int kb_per_split, TcEpilogue e) {
  using L = TcSmem<TBN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;
  const int nkb = (K + TBK - 1) / TBK;
  const int kb0 = blockIdx.z * kb_per_split, kb1 = min(nkb, kb0 + kb_per_split);

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, TBN);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (tc::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        tc::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        tc::mbar_expect_tx(&full[stage], L::STAGE_BYTES);
        if (!a_mn) tc::tma_load_2d(sa, &tmA, &full[stage], kb * TBK, m0);
        else
          for (int i = 0; i < TBM / 32; ++i) tc::tma_load_2d(sa + i * 4096, &tmA, &full[stage], m0 + 32 * i, kb * TBK);
        if (!b_mn) tc::tma_load_2d(sb, &tmB, &full[stage], kb * TBK, n0);
        else
          for (int i = 0; i < TBN / 32; ++i) tc::tma_load_2d(sb + i * 4096, &tmB, &full[stage], n0 + 32 * i, kb * TBK);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (tc::elect_one()) {
      const uint32_t idesc = tc::idesc_tf32(TBM, TBN, a_mn != 0, b_mn != 0);
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        tc::mbar_wait(&full[stage], phase);
        tc::tc_fence_after();
        const uint32_t sa = tc::smem_u32(smem + stage * L::STAGE_BYTES);
        const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < TBK / UMMA_K; ++k) {
          // K-major: 8 tf32 = 32 bytes further along the swizzled row; MN-major: next 8 k-rows = 1024 bytes
          uint64_t ad = a_mn ? tc::smem_desc_mnmajor(sa + k * 1024, 4096) : tc::smem_desc_kmajor(sa + k * 32);
          uint64_t bd = b_mn ? tc::smem_desc_mnmajor(sb + k * 1024, 4096) : tc::smem_desc_kmajor(sb + k * 32);
          tc::mma_tf32(tmem_base, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        tc::mma_commit(&empty[stage]);        // frees the smem slot once these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      tc::mma_commit(tmem_full);              // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5 own TMEM lane quarters (warp % 4) =====
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    tc::mbar_wait(tmem_full, 0);
    tc::tc_fence_after();
    const bool have_k = kb1 > kb0;
#pragma unroll 1
    for (int c = 0; c < TBN / 32; ++c) {
      float v[32];
      __syncwarp();                                   // tcgen05.ld is warp-collective (.sync.aligned)
      tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      const int col0 = n0 + c * 32;
      if (row >= M || col0 >= N) continue;            // re-converges at the __syncwarp above
      const int64_t base = (int64_t)row * e.ldc + col0;
      const bool first = blockIdx.z == 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = have_k ? e.alpha * v[j] : 0.f;
        const int col = col0 + j;
        if (col < N) {
          if (e.bias && first) x += __ldg(e.bias + col);
          if (e.act == 1) x = fmaxf(x, 0.f);
          if (e.mask) x = (__ldg(e.mask + base + j) > 0.f) ? x : 0.f;
          if (e.drop_p > 0.f) x *= mmfn_dropout_scale(e.drop_p, e.drop_seed, (uint64_t)(base + j));
          if (e.res && first) x += __ldg(e.res + base + j);
        }
        v[j] = x;
      }
      float* dst = e.C + base;
      if (e.accum == 0 && col0 + 32 <= N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
        for (int j = 0; j < 32 && col0 + j < N; ++j) {
          if (e.accum == 0) dst[j] = v[j];
          else if (e.accum == 1) dst[j] += v[j];
          else atomicAdd(dst + j, v[j]);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, TBN);
}

}  // namespace

// ---------------------------------------------------------------- host side
PFN_encodeTiled mmfn_get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int mmfn_make_tmap_f32(CUtensorMap* out, const float* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_elems, const uint32_t* box, const uint32_t* elem_strides,
                       bool swizzle32) {
  PFN_encodeTiled enc = mmfn_get_encode_tiled();
  if (!enc) { mmfn_set_error("cuTensorMapEncodeTiled is unavailable"); return (int)cudaErrorNotSupported; }
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 1; i < rank; ++i) gs[i - 1] = strides_elems[i] * sizeof(float);
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, (cuuint32_t)rank, const_cast<float*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mmfn_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu box %u,%u", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
    return (int)cudaErrorInvalidValue;
  }
  return 0;
}

template <int TBN, int STAGES>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, int a_mn, int b_mn,
                     int splitk, const TcEpilogue& e, cudaStream_t stream) {
  using L = TcSmem<TBN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(gemm_tc_kernel<TBN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (ce != cudaSuccess) { mmfn_set_error("gemm_tf32: smem attribute: %s", cudaGetErrorString(ce)); return (int)ce; }
    attr_set = true;
  }
  int nkb = (K + TBK - 1) / TBK;
  int kb_per = (nkb + splitk - 1) / splitk;
  splitk = (nkb + kb_per - 1) / kb_per;
  dim3 grid((N + TBN - 1) / TBN, (M + TBM - 1) / TBM, splitk);
  gemm_tc_kernel<TBN, STAGES><<<grid, TC_THREADS, L::TOTAL, stream>>>(ta, tb, M, N, K, a_mn, b_mn, kb_per, e);
  return mmfn_launch_status("gemm_tf32");
}

// C(M,N) (+)= alpha * op(A) * op(B)^T on the tensor cores (TF32 multiply, FP32 accumulate).
//   a_mn == 0: A is a row-major (M, K) matrix with row pitch lda;  a_mn == 1: A is stored (K, M), pitch lda.
//   b_mn == 0: B is a row-major (N, K) matrix with row pitch ldb;  b_mn == 1: B is stored (K, N), pitch ldb.
// Pitches must be multiples of 4 floats and bases 16-byte aligned (TMA).  Epilogue as mmfn_gemm_f32.
// splitk <= 0 picks a split that fills the 148 SMs (needs accum == 2 and a linear epilogue).
MMFN_API int mmfn_gemm_tf32(const float* A, int64_t lda, int a_mn, const float* B, int64_t ldb, int b_mn,
                            float* C, int64_t ldc, int M, int N, int K,
                            const float* bias, const float* res, const float* mask,
                            float alpha, int act, int accum, float drop_p, uint64_t drop_seed,
                            int splitk, cudaStream_t stream) {
  MMFN_CHECK_ARG(A && B && C, "gemm_tf32: null operand");
  MMFN_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_tf32: bad sizes");
  MMFN_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0, "gemm_tf32: operand pitches must be multiples of 4 floats");
  MMFN_CHECK_ARG((((uintptr_t)A | (uintptr_t)B) & 15) == 0, "gemm_tf32: operands must be 16-byte aligned");
  const bool linear = (act == 0 && !mask && drop_p == 0.f);
  if (splitk <= 0) {
    splitk = 1;
    if (accum == 2 && linear) {
      int tiles = ((M + TBM - 1) / TBM) * ((N + 127) / 128);
      int nkb = (K + TBK - 1) / TBK;
      splitk = max(1, min(nkb / 4, (2 * 148 + tiles - 1) / tiles));
    }
  }
  MMFN_CHECK_ARG(splitk == 1 || (accum == 2 && linear), "gemm_tf32: split-K needs a linear atomic epilogue");
  const int tbn = (N <= 64) ? 64 : 128;
  CUtensorMap ta, tb;
  {
    uint64_t dims[2], strides[2] = {1, (uint64_t)lda};
    uint32_t box[2];
    if (!a_mn) { dims[0] = K; dims[1] = M; box[0] = TBK; box[1] = TBM; }
    else       { dims[0] = M; dims[1] = K; box[0] = 32;  box[1] = TBK; }
    if (int rc = mmfn_make_tmap_f32(&ta, A, 2, dims, strides, box, nullptr, a_mn != 0)) return rc;
  }
  {
    uint64_t dims[2], strides[2] = {1, (uint64_t)ldb};
    uint32_t box[2];
    if (!b_mn) { dims[0] = K; dims[1] = N; box[0] = TBK; box[1] = (uint32_t)tbn; }
    else       { dims[0] = N; dims[1] = K; box[0] = 32;  box[1] = TBK; }
    if (int rc = mmfn_make_tmap_f32(&tb, B, 2, dims, strides, box, nullptr, b_mn != 0)) return rc;
  }
  TcEpilogue e{C, ldc, bias, res, mask, alpha, act, accum, drop_p, drop_seed};
  if (tbn == 64) return launch_tc<64, 4>(ta, tb, M, N, K, a_mn, b_mn, splitk, e, stream);
  return launch_tc<128, 3>(ta, tb, M, N, K, a_mn, b_mn, splitk, e, stream);
}
