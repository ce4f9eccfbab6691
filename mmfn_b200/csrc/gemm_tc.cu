// TF32 tensor-core GEMM for the dense Linear layers of the fusion transformers and heads
// (model_rad.py:82-89, :120-125 and their backward passes), on the generic tcgen05 pipeline of
// tc_kernel.cuh.
//
//   C[M,N] (+)= alpha * A * B^T, fp32 in HBM, read as TF32 (TMA rounds), fp32 accumulate in TMEM.
//
// Either operand may be K-major (row = M/N index, reduction contiguous: activations, weights in
// forward) or MN-major (reduction is the slow dimension: W in dgrad, dY/X in wgrad), so forward,
// data-gradient and weight-gradient GEMMs all read the same row-major tensors with no transposes.
#include "tc_kernel.cuh"

namespace {

template <bool AMN, bool BMN, int TBN>
struct GemmOp {
  static constexpr bool A_MN = AMN, B_MN = BMN;
  int M, N, K, kb_per_split;
  int64_t ldc;
  int m0, n0, kb0, kb1;
  __device__ void setup() {
    m0 = blockIdx.y * tc::TBM;
    n0 = blockIdx.x * TBN;
    int nkb = (K + tc::TBK - 1) / tc::TBK;
    kb0 = blockIdx.z * kb_per_split;
    kb1 = min(nkb, kb0 + kb_per_split);
  }
  __device__ int kb_begin() const { return kb0; }
  __device__ int kb_end() const { return kb1; }
  __device__ void load(int kb, uint8_t* sa, uint8_t* sb, uint64_t* bar, const CUtensorMap* ta, const CUtensorMap* tb) const {
    if constexpr (!AMN) tc::tma_load_2d(sa, ta, bar, kb * tc::TBK, m0);
    else
      for (int i = 0; i < tc::TBM / 32; ++i) tc::tma_load_2d(sa + i * tc::BOX_BYTES, ta, bar, m0 + 32 * i, kb * tc::TBK);
    if constexpr (!BMN) tc::tma_load_2d(sb, tb, bar, kb * tc::TBK, n0);
    else
      for (int i = 0; i < TBN / 32; ++i) tc::tma_load_2d(sb + i * tc::BOX_BYTES, tb, bar, n0 + 32 * i, kb * tc::TBK);
  }
  __device__ bool out_row(int r, int64_t& off) const {
    off = (int64_t)(m0 + r) * ldc;
    return m0 + r < M;
  }
  __device__ int n_cols() const { return N; }
  __device__ int col0() const { return n0; }
  __device__ bool first_split() const { return blockIdx.z == 0; }
};

template <bool AMN, bool BMN>
int run_gemm(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, int64_t ldc, int tbn, int splitk,
             const tc::Epilogue& e, cudaStream_t stream) {
  int nkb = (K + tc::TBK - 1) / tc::TBK;
  int kb_per = (nkb + splitk - 1) / splitk;
  splitk = (nkb + kb_per - 1) / kb_per;
  if (tbn == 64) {
    GemmOp<AMN, BMN, 64> op{M, N, K, kb_per, ldc};
    return tc::launch<GemmOp<AMN, BMN, 64>, 64, 4>(ta, tb, op, e, dim3((N + 63) / 64, (M + tc::TBM - 1) / tc::TBM, splitk), stream, "gemm_tf32");
  }
  GemmOp<AMN, BMN, 128> op{M, N, K, kb_per, ldc};
  return tc::launch<GemmOp<AMN, BMN, 128>, 128, 3>(ta, tb, op, e, dim3((N + 127) / 128, (M + tc::TBM - 1) / tc::TBM, splitk), stream, "gemm_tf32");
}

}  // namespace

// ---------------------------------------------------------------- host side
static unsigned long long* g_trace = nullptr;
unsigned long long* mmfn_tc_trace_ptr() { return g_trace; }

// Developer aid: when buf (8 x uint64, device) is non-null every tensor-core kernel's CTA 0 writes
// %globaltimer stamps of its pipeline phases there.  Pass null to switch it off (default).
MMFN_API int mmfn_tc_set_trace(unsigned long long* buf) {
  g_trace = buf;
  return 0;
}

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
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mmfn_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                   (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                   box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return (int)cudaErrorInvalidValue;
  }
  return 0;
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
      int tiles = ((M + tc::TBM - 1) / tc::TBM) * ((N + 127) / 128);
      int nkb = (K + tc::TBK - 1) / tc::TBK;
      splitk = max(1, min(nkb / 4, (2 * 148 + tiles - 1) / tiles));
    }
  }
  MMFN_CHECK_ARG(splitk == 1 || (accum == 2 && linear), "gemm_tf32: split-K needs a linear atomic epilogue");
  const int tbn = (N <= 64) ? 64 : 128;
  CUtensorMap ta, tb;
  {
    uint64_t dims[2], strides[2] = {1, (uint64_t)lda};
    uint32_t box[2];
    if (!a_mn) { dims[0] = K; dims[1] = M; box[0] = tc::TBK; box[1] = tc::TBM; }
    else       { dims[0] = M; dims[1] = K; box[0] = 32;      box[1] = tc::TBK; }
    if (int rc = mmfn_make_tmap_f32(&ta, A, 2, dims, strides, box, nullptr, a_mn != 0)) return rc;
  }
  {
    uint64_t dims[2], strides[2] = {1, (uint64_t)ldb};
    uint32_t box[2];
    if (!b_mn) { dims[0] = K; dims[1] = N; box[0] = tc::TBK; box[1] = (uint32_t)tbn; }
    else       { dims[0] = N; dims[1] = K; box[0] = 32;      box[1] = tc::TBK; }
    if (int rc = mmfn_make_tmap_f32(&tb, B, 2, dims, strides, box, nullptr, b_mn != 0)) return rc;
  }
  tc::Epilogue e{C, bias, res, mask, alpha, act, accum, drop_p, drop_seed, mmfn_tc_trace_ptr()};
  if (!a_mn && !b_mn) return run_gemm<false, false>(ta, tb, M, N, K, ldc, tbn, splitk, e, stream);
  if (!a_mn && b_mn) return run_gemm<false, true>(ta, tb, M, N, K, ldc, tbn, splitk, e, stream);
  if (a_mn && !b_mn) return run_gemm<true, false>(ta, tb, M, N, K, ldc, tbn, splitk, e, stream);
  return run_gemm<true, true>(ta, tb, M, N, K, ldc, tbn, splitk, e, stream);
}

MMFN_DEFINE_RNG_BINDER(gemm_tc)
