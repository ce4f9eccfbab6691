// TF32 tensor-core GEMM for the dense Linear layers of the fusion transformers and heads
// (model_rad.py:82-89, :120-125 and their backward passes), on the generic tcgen05 pipeline of
// tc_kernel.cuh.
//
//   C[M,N] (+)= alpha * A * B^T, fp32 in HBM, read as TF32 (TMA rounds), fp32 accumulate in TMEM.
//
// Either operand may be K-major (row = M/N index, reduction contiguous: activations, weights in
// forward) or MN-major (reduction is the slow dimension: W in dgrad, dY/X in wgrad), so forward,
// data-gradient and weight-gradient GEMMs all read the same row-major tensors with no transposes.
#include <cstdlib>
#include "tc_kernel.cuh"

namespace {

// Operand addressing: every operand is a rank-4 tensor map (inner, and three outer dimensions sorted
// by stride on the host).  pos[0..2] = position (1..3) of the strided matrix dimension, batch-1 and
// batch-0 coordinates inside that map.
struct OperandPos { int row, b1, b0; };

template <bool AMN, bool BMN, int TBN, int EB_ = 32>
struct GemmOp {
  static constexpr bool A_MN = AMN, B_MN = BMN;
  static constexpr int EB = EB_;                       // elements per 128-byte k-block: 32 (tf32) / 64 (bf16)
  using ET = ElemTraits<EB_>;
  int M, N, K, kb_per_split, splitk, nb1;
  int64_t ldc, c_sb0, c_sb1;
  OperandPos pa, pb;
  int m0, n0, kb0, kb1, b0, b1;
  __device__ void setup() {
    m0 = blockIdx.y * tc::TBM;
    n0 = blockIdx.x * TBN;
    int bz = blockIdx.z / splitk, split = blockIdx.z - bz * splitk;
    b0 = bz / nb1; b1 = bz - b0 * nb1;
    int nkb = (K + EB - 1) / EB;
    kb0 = split * kb_per_split;
    kb1 = min(nkb, kb0 + kb_per_split);
  }
  // persistent kernel: linear tile index (n fastest), whole K, single batch
  __device__ void set_tile(int t) {
    const int tn = (N + TBN - 1) / TBN;
    const int tm = t / tn;
    m0 = tm * tc::TBM;
    n0 = (t - tm * tn) * TBN;
    b0 = 0; b1 = 0; kb0 = 0; kb1 = (K + EB - 1) / EB;
  }
  __device__ int kb_total() const { return (K + EB - 1) / EB; }
  __device__ int kb_begin() const { return kb0; }
  __device__ int kb_end() const { return kb1; }
  __device__ static void issue(uint8_t* dst, const CUtensorMap* t, uint64_t* bar, const OperandPos& p, int inner, int row,
                               int b1_, int b0_) {
    int c[4];
    c[0] = inner; c[p.row] = row; c[p.b1] = b1_; c[p.b0] = b0_;
    tc::tma_load_4d(dst, t, bar, c[0], c[1], c[2], c[3]);
  }
  __device__ void load(int kb, uint8_t* sa, uint8_t* sb, uint64_t* bar, const CUtensorMap* ta, const CUtensorMap* tb) const {
    if constexpr (!AMN) issue(sa, ta, bar, pa, kb * EB, m0, b1, b0);
    else
      for (int i = 0; i < tc::TBM / EB; ++i) issue(sa + i * ET::BOX_BYTES, ta, bar, pa, m0 + EB * i, kb * EB, b1, b0);
    if constexpr (!BMN) issue(sb, tb, bar, pb, kb * EB, n0, b1, b0);
    else
      for (int i = 0; i < TBN / EB; ++i) issue(sb + i * ET::BOX_BYTES, tb, bar, pb, n0 + EB * i, kb * EB, b1, b0);
  }
  __device__ bool out_row(int r, int64_t& off) const {
    off = (int64_t)b0 * c_sb0 + (int64_t)b1 * c_sb1 + (int64_t)(m0 + r) * ldc;
    return m0 + r < M;
  }
  __device__ int n_cols() const { return N; }
  __device__ int col0() const { return n0; }
  __device__ bool first_split() const { return splitk == 1 || (blockIdx.z % splitk) == 0; }
};

// Large plain GEMMs go to the persistent kernel (tc_kernel.cuh): >= 2 tiles per SM, no split-K, no batch.  MMFN_GEMM_PERSIST=0
// switches it off (A/B runs).
static int g_persist_mode = -1;
static int persist_mode() {
  if (g_persist_mode < 0) { const char* v = getenv("MMFN_GEMM_PERSIST"); g_persist_mode = (v && v[0] == '0') ? 0 : 1; }
  return g_persist_mode;
}
static int persist_wide_tiles() {                        // MMFN_GEMM_PERSIST_256=1: 128 x 256 tiles where N % 256 == 0 (measured: no better than 128 x 128 on average)
  static int mode = -1;
  if (mode < 0) { const char* v = getenv("MMFN_GEMM_PERSIST_256"); mode = (v && v[0] == '1') ? 1 : 0; }
  return mode;
}
static int sm_count();
// smallest tile count the persistent kernel takes (MMFN_GEMM_PERSIST_MIN; default two tiles per SM)
static int persist_min_tiles() {
  static int n = 0;
  if (!n) { const char* v = getenv("MMFN_GEMM_PERSIST_MIN"); n = v ? atoi(v) : 0; if (n <= 0) n = 2 * sm_count(); }
  return n;
}
static int sm_count() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
  return n;
}

struct GemmArgs {
  int M, N, K, nb0, nb1, splitk, tbn;
  int64_t ldc, c_sb0, c_sb1;
  OperandPos pa, pb;
};

template <bool AMN, bool BMN, int EB>
int run_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& g, const tc::Epilogue& e, cudaStream_t stream) {
  const char* what = EB == 64 ? "gemm_bf16" : "gemm_tf32";
  int nkb = (g.K + EB - 1) / EB;
  int kb_per = (nkb + g.splitk - 1) / g.splitk;
  int splitk = (nkb + kb_per - 1) / kb_per;
  unsigned gz = (unsigned)(g.nb0 * g.nb1 * splitk);
  if (g.tbn == 64) {
    GemmOp<AMN, BMN, 64, EB> op{g.M, g.N, g.K, kb_per, splitk, g.nb1, g.ldc, g.c_sb0, g.c_sb1, g.pa, g.pb};
    return tc::launch<GemmOp<AMN, BMN, 64, EB>, 64, 4, true>(ta, tb, op, e, dim3((g.N + 63) / 64, (g.M + tc::TBM - 1) / tc::TBM, gz), stream, what);
  }
  if (g.tbn == 256) {                                    // (chosen by gemm_tc_impl only for the persistent kernel)
    GemmOp<AMN, BMN, 256, EB> op{g.M, g.N, g.K, kb_per, splitk, g.nb1, g.ldc, g.c_sb0, g.c_sb1, g.pa, g.pb};
    const int ntiles = (g.N / 256) * ((g.M + tc::TBM - 1) / tc::TBM);
    return tc::launch_persist<GemmOp<AMN, BMN, 256, EB>, 256>(ta, tb, op, e, ntiles, sm_count(), stream, what);
  }
  GemmOp<AMN, BMN, 128, EB> op{g.M, g.N, g.K, kb_per, splitk, g.nb1, g.ldc, g.c_sb0, g.c_sb1, g.pa, g.pb};
  {
    const int ntiles = ((g.N + 127) / 128) * ((g.M + tc::TBM - 1) / tc::TBM);
    if (splitk == 1 && gz == 1 && ntiles >= persist_min_tiles() && e.bn_ws == nullptr && e.trace == nullptr && persist_mode())
      return tc::launch_persist<GemmOp<AMN, BMN, 128, EB>, 128>(ta, tb, op, e, ntiles, sm_count(), stream, what);
  }
  return tc::launch<GemmOp<AMN, BMN, 128, EB>, 128, 3, true>(ta, tb, op, e, dim3((g.N + 127) / 128, (g.M + tc::TBM - 1) / tc::TBM, gz), stream, what);
}

// rank-4 map of one operand.  inner_len x strided_len matrix per batch; outer dims sorted by stride.
int make_operand_tmap(CUtensorMap* out, OperandPos* pos, const void* base, bool mn_major, int rows, int K, int64_t ld,
                      int nb0, int nb1, int64_t sb0, int64_t sb1, int tile_rows, int EB = 32) {
  // logical outer dims: 0 = strided matrix dim, 1 = batch-1, 2 = batch-0
  uint64_t len[3] = {(uint64_t)(mn_major ? K : rows), (uint64_t)nb1, (uint64_t)nb0};
  int64_t str[3] = {ld, sb1, sb0};
  uint32_t bx[3] = {(uint32_t)(mn_major ? EB : tile_rows), 1, 1};
  uint64_t inner_len = (uint64_t)(mn_major ? rows : K);
  // size-1 batch dims get a harmless stride that keeps the map monotonic
  int64_t span = ld * (int64_t)len[0];
  if (nb1 == 1) str[1] = span;
  if (nb0 == 1) str[2] = (nb1 == 1 ? span : (str[1] * nb1 > span ? str[1] * nb1 : span));
  int order[3] = {0, 1, 2};
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (str[order[j]] < str[order[i]]) { int t = order[i]; order[i] = order[j]; order[j] = t; }
  uint64_t dims[4] = {inner_len, 0, 0, 0}, strides[4] = {1, 0, 0, 0};
  uint32_t box[4] = {(uint32_t)EB, 0, 0, 0};
  int where[3];
  for (int i = 0; i < 3; ++i) {
    dims[i + 1] = len[order[i]];
    strides[i + 1] = (uint64_t)str[order[i]];
    box[i + 1] = bx[order[i]];
    where[order[i]] = i + 1;
  }
  pos->row = where[0]; pos->b1 = where[1]; pos->b0 = where[2];
  const int align = EB == 64 ? 8 : 4;                   // TMA global strides are multiples of 16 bytes
  for (int i = 1; i < 4; ++i)
    if (strides[i] % align != 0 || strides[i] == 0) {
      mmfn_set_error("gemm_tc: operand strides must be non-zero multiples of 16 bytes (got %lld elements)", (long long)strides[i]);
      return MMFN_BAD_ARG;
    }
  if (EB == 64) return mmfn_make_tmap_bf16(out, base, 4, dims, strides, box, nullptr);
  return mmfn_make_tmap_f32(out, static_cast<const float*>(base), 4, dims, strides, box, nullptr, mn_major);
}

}  // namespace

// ---------------------------------------------------------------- host side
static unsigned long long* g_trace = nullptr;
unsigned long long* mmfn_tc_trace_ptr() { return g_trace; }

// Developer aid: when buf (8 x uint64, device) is non-null every tensor-core kernel's CTA 0 writes
// %globaltimer stamps of its pipeline phases there.  Pass null to switch it off (default).
// Developer aid: device buffer of 16 x 8 uint64 receiving the per-tile %globaltimer stamps of CTA 0 of the persistent GEMM
// kernel (null = off); see tc_kernel.cuh.
MMFN_API int mmfn_tc_set_persist_trace(unsigned long long* buf) {
  return (int)cudaMemcpyToSymbol(tc::g_tcp_trace, &buf, sizeof(buf));
}

// Route large plain GEMMs (>= 2 output tiles per SM, no split-K, no batch) to the persistent kernel (1, default) or keep
// every GEMM on the one-tile-per-CTA kernel (0).  Both produce bit-identical results; tests and A/B timing use this.
MMFN_API int mmfn_set_gemm_persist(int on) { g_persist_mode = on ? 1 : 0; return 0; }

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
                       bool swizzle32, bool as_tf32) {
  PFN_encodeTiled enc = mmfn_get_encode_tiled();
  if (!enc) { mmfn_set_error("cuTensorMapEncodeTiled is unavailable"); return (int)cudaErrorNotSupported; }
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 1; i < rank; ++i) gs[i - 1] = strides_elems[i] * sizeof(float);
  CUresult r = enc(out, as_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), gd, gs, bx, es,
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

int mmfn_make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_elems, const uint32_t* box, const uint32_t* elem_strides) {
  PFN_encodeTiled enc = mmfn_get_encode_tiled();
  if (!enc) { mmfn_set_error("cuTensorMapEncodeTiled is unavailable"); return (int)cudaErrorNotSupported; }
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 1; i < rank; ++i) gs[i - 1] = strides_elems[i] * 2;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mmfn_set_error("cuTensorMapEncodeTiled(bf16) failed (%d): rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                   (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                   box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return (int)cudaErrorInvalidValue;
  }
  return 0;
}

template <int EB>
static int gemm_tc_impl(const void* A, int64_t lda, int a_mn, int64_t a_sb0, int64_t a_sb1,
                        const void* B, int64_t ldb, int b_mn, int64_t b_sb0, int64_t b_sb1,
                        void* C, int c_bf16, int64_t ldc, int64_t c_sb0, int64_t c_sb1,
                        int M, int N, int K, int nb0, int nb1,
                        const float* bias, const float* res, const float* mask, const void* mask16,
                        float alpha, int act, int accum, float drop_p, uint64_t drop_seed,
                        int splitk, cudaStream_t stream) {
  const char* what = EB == 64 ? "gemm_bf16" : "gemm_tf32";
  const int align = EB == 64 ? 8 : 4;
  MMFN_CHECK_ARG(A && B && C, "%s: null operand", what);
  MMFN_CHECK_ARG(M > 0 && N > 0 && K > 0 && nb0 >= 1 && nb1 >= 1, "%s: bad sizes", what);
  MMFN_CHECK_ARG(lda % align == 0 && ldb % align == 0 && lda > 0 && ldb > 0, "%s: operand pitches must be multiples of 16 bytes", what);
  MMFN_CHECK_ARG((((uintptr_t)A | (uintptr_t)B) & 15) == 0, "%s: operands must be 16-byte aligned", what);
  MMFN_CHECK_ARG(!c_bf16 || accum == 0, "%s: a bf16 result cannot be accumulated", what);
  const bool linear = (act == 0 && !mask && !mask16 && drop_p == 0.f);
  const int nb = nb0 * nb1;
  // Tile width: 128 x 128 unless that leaves most SMs idle -- the epilogue is bound by the per-SM store path
  // (~1 us per 128 x 128 fp32 tile), so small problems finish sooner as twice as many 128 x 64 tiles.
  const int mt = (M + tc::TBM - 1) / tc::TBM;
  const int tbn = (N <= 64 || (int64_t)mt * ((N + 127) / 128) * nb * (splitk > 1 ? splitk : 1) < 148) ? 64 : 128;
  if (splitk <= 0) {
    splitk = 1;
    if (accum == 2 && linear) {
      int tiles = mt * ((N + tbn - 1) / tbn) * nb;
      int nkb = (K + EB - 1) / EB;
      splitk = max(1, min(nkb / 4, (2 * 148) / tiles));
    }
  }
  // 128 x 256 tiles on the persistent kernel when the problem still gives every SM two of them
  int tbn_final = tbn;
  if (tbn == 128 && splitk == 1 && nb == 1 && N % 256 == 0 && (int64_t)mt * (N / 256) >= 2 * sm_count() && mmfn_tc_trace_ptr() == nullptr &&
      persist_mode() && persist_wide_tiles())
    tbn_final = 256;
  MMFN_CHECK_ARG(splitk == 1 || (accum == 2 && linear), "%s: split-K needs a linear atomic epilogue", what);
  MMFN_CHECK_ARG((int64_t)nb * splitk <= 65535, "%s: too many batches x splits", what);
  GemmArgs g{M, N, K, nb0, nb1, splitk, tbn_final, ldc, c_sb0, c_sb1};
  CUtensorMap ta, tb;
  if (int rc = make_operand_tmap(&ta, &g.pa, A, a_mn != 0, M, K, lda, nb0, nb1, a_sb0, a_sb1, tc::TBM, EB)) return rc;
  if (int rc = make_operand_tmap(&tb, &g.pb, B, b_mn != 0, N, K, ldb, nb0, nb1, b_sb0, b_sb1, g.tbn, EB)) return rc;
  tc::Epilogue e{c_bf16 ? nullptr : static_cast<float*>(C), bias, res, mask, alpha, act, accum, drop_p, drop_seed, mmfn_tc_trace_ptr(),
                 c_bf16 ? static_cast<__nv_bfloat16*>(C) : nullptr, static_cast<const __nv_bfloat16*>(mask16)};
  if (!a_mn && !b_mn) return run_gemm<false, false, EB>(ta, tb, g, e, stream);
  if (!a_mn && b_mn) return run_gemm<false, true, EB>(ta, tb, g, e, stream);
  if (a_mn && !b_mn) return run_gemm<true, false, EB>(ta, tb, g, e, stream);
  return run_gemm<true, true, EB>(ta, tb, g, e, stream);
}

// C[b0,b1](M,N) (+)= alpha * op(A) * op(B)^T on the tensor cores (TF32 multiply, FP32 accumulate), batched
// over nb0 x nb1 problems with per-operand batch strides (elements).
//   a_mn == 0: A[b] is a row-major (M, K) matrix with row pitch lda;  a_mn == 1: A[b] is stored (K, M), pitch lda.
//   b_mn == 0: B[b] is a row-major (N, K) matrix with row pitch ldb;  b_mn == 1: B[b] is stored (K, N), pitch ldb.
// Pitches and batch strides must be non-zero multiples of 4 floats and bases 16-byte aligned (TMA).
// K and the MN extents need no alignment: the TMA unit zero-fills past the logical matrix edge (e.g. the
// 16-wide heads of transformer1).  Epilogue as mmfn_gemm_f32.  splitk <= 0 picks a split that fills the SMs
// (needs accum == 2 and a linear epilogue).
MMFN_API int mmfn_gemm_tf32(const float* A, int64_t lda, int a_mn, int64_t a_sb0, int64_t a_sb1,
                            const float* B, int64_t ldb, int b_mn, int64_t b_sb0, int64_t b_sb1,
                            float* C, int64_t ldc, int64_t c_sb0, int64_t c_sb1,
                            int M, int N, int K, int nb0, int nb1,
                            const float* bias, const float* res, const float* mask,
                            float alpha, int act, int accum, float drop_p, uint64_t drop_seed,
                            int splitk, cudaStream_t stream) {
  return gemm_tc_impl<32>(A, lda, a_mn, a_sb0, a_sb1, B, ldb, b_mn, b_sb0, b_sb1, C, 0, ldc, c_sb0, c_sb1, M, N, K, nb0, nb1,
                          bias, res, mask, nullptr, alpha, act, accum, drop_p, drop_seed, splitk, stream);
}

// mmfn_gemm_tf32 with the result written as bf16 (c_bf16 != 0; C is then a bf16 tensor, ldc / batch strides in
// elements, no accumulation): TF32 attention-gradient products whose output feeds a bf16 GEMM (BASELINE configs[2]).
MMFN_API int mmfn_gemm_tf32_out(const float* A, int64_t lda, int a_mn, int64_t a_sb0, int64_t a_sb1,
                                const float* B, int64_t ldb, int b_mn, int64_t b_sb0, int64_t b_sb1,
                                void* C, int c_bf16, int64_t ldc, int64_t c_sb0, int64_t c_sb1,
                                int M, int N, int K, int nb0, int nb1, float alpha, cudaStream_t stream) {
  return gemm_tc_impl<32>(A, lda, a_mn, a_sb0, a_sb1, B, ldb, b_mn, b_sb0, b_sb1, C, c_bf16, ldc, c_sb0, c_sb1, M, N, K, nb0, nb1,
                          nullptr, nullptr, nullptr, nullptr, alpha, 0, 0, 0.f, 0, 1, stream);
}

// The bf16 tensor-core GEMM of BASELINE configs[2] (torch.autocast(bfloat16) over the reference's nn.Linear layers,
// model_rad.py:82-89, :120-125): A and B are bf16 tensors (activations written in bf16 by their producers, the bf16
// shadow of the fp32 master weights), multiplied with tcgen05 kind::f16, accumulated in fp32 in TMEM.  Same operand
// conventions as mmfn_gemm_tf32 (either operand K-major or MN-major, two batch dimensions; pitches / batch strides in
// ELEMENTS, multiples of 8; bases 16-byte aligned).  C: fp32 (c_bf16 == 0; supports accum / split-K) or bf16
// (c_bf16 != 0: tensors that only feed further MMAs).  bias / res are fp32; mask is a BF16 tensor indexed like C.
MMFN_API int mmfn_gemm_bf16(const void* A, int64_t lda, int a_mn, int64_t a_sb0, int64_t a_sb1,
                            const void* B, int64_t ldb, int b_mn, int64_t b_sb0, int64_t b_sb1,
                            void* C, int c_bf16, int64_t ldc, int64_t c_sb0, int64_t c_sb1,
                            int M, int N, int K, int nb0, int nb1,
                            const float* bias, const float* res, const void* mask_bf16,
                            float alpha, int act, int accum, float drop_p, uint64_t drop_seed,
                            int splitk, cudaStream_t stream) {
  return gemm_tc_impl<64>(A, lda, a_mn, a_sb0, a_sb1, B, ldb, b_mn, b_sb0, b_sb1, C, c_bf16, ldc, c_sb0, c_sb1, M, N, K, nb0, nb1,
                          bias, res, nullptr, mask_bf16, alpha, act, accum, drop_p, drop_seed, splitk, stream);
}

MMFN_DEFINE_RNG_BINDER(gemm_tc)
