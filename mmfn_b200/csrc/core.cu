// Library bookkeeping + the GEMM / convolution C-ABI entries of libmmfn_b200.so.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <type_traits>
#include "gemm_simt.cuh"

static thread_local char g_err[512] = "";

void mmfn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

MMFN_API int mmfn_version(void) { return 100; }

// Copies the calling thread's last error string (empty if none) into buf.
MMFN_API int mmfn_last_error(char* buf, int64_t len) {
  if (!buf || len <= 0) return MMFN_BAD_ARG;
  strncpy(buf, g_err, (size_t)len - 1);
  buf[len - 1] = 0;
  return 0;
}

using namespace mmfn;

// Scratch / saved-tensor sizes a caller must allocate (the library owns no device memory).  `op` is one of the
// MMFN_WS_* codes of the header, `dtype` one of MMFN_F32 / MMFN_TF32 / MMFN_BF16 (element size of the tensors the op
// stores: 4 / 4 / 2 bytes); a, b, c, d are the op's sizes:
//   MMFN_WS_BN            a = channels C                     fp64 partial-sum slots + ticket + published coefficients of
//                                                            mmfn_bn_train_fwd / _bwd (zero once; every call leaves it zero)
//   MMFN_WS_STEM_IM2COL   a = N, b = Ho*Wo, c = R*S*C        column matrix (N*Ho*Wo, Kp) of the 7x7/2 stems, Kp = c rounded
//                                                            up to the k-block (32 elements fp32/tf32, 64 bf16)
//   MMFN_WS_STEM_FILTER   a = Co, c = R*S*C                  zero-padded (Co, Kp) filter matrix of the same GEMM
//   MMFN_WS_ATTN_PROB     a = B, b = heads, c = T, d = 1|2   saved attention probabilities (B, heads, T, T); d = 2 with
//                                                            dropout (P and P after dropout)
//   MMFN_WS_GRU_SAVED     a = B, b = steps                   gate activations kept by mmfn_gru_head_fwd for BPTT
//   MMFN_WS_BEV           a = frames                         packed-u16 pillar counters of mmfn_bev_scatter_ws
//                                                            (zero once; every call leaves it zero)
// Replaces nothing in the reference (PyTorch's caching allocator does this implicitly); SURVEY.md section 8(b).
MMFN_API int mmfn_workspace_bytes(int op, int dtype, int64_t a, int64_t b, int64_t c, int64_t d, int64_t* bytes) {
  MMFN_CHECK_ARG(bytes, "workspace_bytes: null output");
  MMFN_CHECK_ARG(dtype >= 0 && dtype <= 2, "workspace_bytes: unknown dtype %d", dtype);
  const int64_t es = dtype == 2 ? 2 : 4, kblock = dtype == 2 ? 64 : 32;
  switch (op) {
    case 0: MMFN_CHECK_ARG(a > 0, "workspace_bytes: C"); *bytes = (34 * a + 8) * 8; return 0;
    case 1: MMFN_CHECK_ARG(a > 0 && b > 0 && c > 0, "workspace_bytes: sizes"); *bytes = a * b * ((c + kblock - 1) / kblock * kblock) * es; return 0;
    case 2: MMFN_CHECK_ARG(a > 0 && c > 0, "workspace_bytes: sizes"); *bytes = a * ((c + kblock - 1) / kblock * kblock) * es; return 0;
    case 3: MMFN_CHECK_ARG(a > 0 && b > 0 && c > 0 && (d == 1 || d == 2), "workspace_bytes: sizes"); *bytes = a * b * c * c * es * d; return 0;
    case 4: MMFN_CHECK_ARG(a > 0 && b > 0, "workspace_bytes: sizes"); *bytes = a * b * 5 * 64 * 4; return 0;
    case 5: MMFN_CHECK_ARG(a > 0, "workspace_bytes: frames"); *bytes = a * 65536 * 4; return 0;
    default: mmfn_set_error("workspace_bytes: unknown op %d", op); return MMFN_BAD_ARG;
  }
}

// C[b0,b1](M,N) (+)= alpha * A(M,K) * B(N,K)^T with fused epilogue; all operands strided.
MMFN_API int mmfn_gemm_f32(const float* A, int64_t a_sr, int64_t a_sk, int64_t a_sb0, int64_t a_sb1,
                           const float* B, int64_t b_sr, int64_t b_sk, int64_t b_sb0, int64_t b_sb1,
                           float* C, int64_t ldc, int64_t c_sb0, int64_t c_sb1,
                           int M, int N, int K, int nb0, int nb1,
                           const float* bias, const float* res, const float* mask,
                           float alpha, int act, int accum, float drop_p, uint64_t drop_seed,
                           int splitk, cudaStream_t stream) {
  MMFN_CHECK_ARG(A && B && C, "gemm: null operand");
  MMFN_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && nb0 >= 1 && nb1 >= 1, "gemm: bad sizes");
  MMFN_CHECK_ARG(splitk >= 1, "gemm: splitk must be >= 1");
  MMFN_CHECK_ARG(splitk == 1 || (accum == 2 && act == 0 && !mask && drop_p == 0.f),
                 "gemm: split-K needs a linear atomic epilogue");
  MMFN_CHECK_ARG((int64_t)nb0 * nb1 * splitk <= 65535, "gemm: too many batches");
  DenseLoader a{A, a_sr, a_sk, a_sb0, a_sb1, M, K};
  DenseLoader b{B, b_sr, b_sk, b_sb0, b_sb1, N, K};
  Epilogue e{C, ldc, c_sb0, c_sb1, bias, res, mask, alpha, act, accum, drop_p, drop_seed};
  return launch_gemm_simt(a, b, M, N, K, e, nb0, nb1, splitk, stream);
}

static int check_geom(const ConvGeom& g) {
  MMFN_CHECK_ARG(g.N > 0 && g.H > 0 && g.W > 0 && g.C > 0 && g.Co > 0, "conv: bad tensor sizes");
  MMFN_CHECK_ARG(g.R > 0 && g.S > 0 && g.stride > 0 && g.pad >= 0, "conv: bad filter");
  MMFN_CHECK_ARG(g.Ho == (g.H + 2 * g.pad - g.R) / g.stride + 1 &&
                 g.Wo == (g.W + 2 * g.pad - g.S) / g.stride + 1, "conv: inconsistent output size");
  MMFN_CHECK_ARG((int64_t)g.N * g.H * g.W < (1ll << 31) && (int64_t)g.N * g.Ho * g.Wo < (1ll << 31),
                 "conv: pixel count overflows int32");
  return 0;
}

// y(N,Ho,Wo,Co) = conv(x(N,H,W,C), w(Co,R,S,C)); NHWC activations, KRSC filters.
MMFN_API int mmfn_conv2d_fwd_f32(const float* x, const float* w, float* y,
                                 int N, int H, int W, int C, int Co, int R, int S, int stride, int pad,
                                 int Ho, int Wo, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && w && y, "conv_fwd: null pointer");
  ConvGeom g{N, H, W, C, R, S, stride, pad, Ho, Wo, Co};
  if (int rc = check_geom(g)) return rc;
  int K = R * S * C;
  Im2colLoader<true> a{x, g};
  DenseLoader b{w, K, 1, 0, 0, Co, K};
  Epilogue e{y, Co, 0, 0, nullptr, nullptr, nullptr, 1.f, 0, 0, 0.f, 0};
  return launch_gemm_simt(a, b, N * Ho * Wo, Co, K, e, 1, 1, 1, stream);
}

// dx(N,H,W,C) = conv_transpose(dy(N,Ho,Wo,Co), wt(C,R,S,Co)) [+ res]; wt is the CRSK
// permutation of the filters (mmfn_filter_krsc_to_crsk).
MMFN_API int mmfn_conv2d_dgrad_f32(const float* dy, const float* wt, float* dx, const float* res,
                                   int N, int H, int W, int C, int Co, int R, int S, int stride, int pad,
                                   int Ho, int Wo, cudaStream_t stream) {
  MMFN_CHECK_ARG(dy && wt && dx, "conv_dgrad: null pointer");
  ConvGeom g{N, H, W, C, R, S, stride, pad, Ho, Wo, Co};
  if (int rc = check_geom(g)) return rc;
  int K = R * S * Co;
  DgradLoader a{dy, g};
  DenseLoader b{wt, K, 1, 0, 0, C, K};
  Epilogue e{dx, C, 0, 0, nullptr, res, nullptr, 1.f, 0, 0, 0.f, 0};
  return launch_gemm_simt(a, b, N * H * W, C, K, e, 1, 1, 1, stream);
}

// dw(Co,R,S,C) += dy^T * im2col(x); split-K over the pixel dimension, atomic accumulate.
MMFN_API int mmfn_conv2d_wgrad_f32(const float* dy, const float* x, float* dw,
                                   int N, int H, int W, int C, int Co, int R, int S, int stride, int pad,
                                   int Ho, int Wo, int splitk, cudaStream_t stream) {
  MMFN_CHECK_ARG(dy && x && dw, "conv_wgrad: null pointer");
  ConvGeom g{N, H, W, C, R, S, stride, pad, Ho, Wo, Co};
  if (int rc = check_geom(g)) return rc;
  int Kt = R * S * C, P = N * Ho * Wo;
  if (splitk <= 0) {
    int tiles = ((Co + BM - 1) / BM) * ((Kt + BN - 1) / BN);
    splitk = max(1, min(min(1024, (P + 4 * BK - 1) / (4 * BK)), (148 * 4 + tiles - 1) / tiles));
  }
  DenseLoader a{dy, 1, Co, 0, 0, Co, P};          // A(co, pix) = dy[pix*Co + co]
  Im2colLoader<false> b{x, g};                    // B(tap, pix)
  Epilogue e{dw, Kt, 0, 0, nullptr, nullptr, nullptr, 1.f, 0, 2, 0.f, 0};
  return launch_gemm_simt(a, b, Co, Kt, P, e, 1, 1, splitk, stream);
}

MMFN_DEFINE_RNG_BINDER(core)
int mmfn_bind_rng_attn(const unsigned long long*);
int mmfn_bind_rng_gemm_tc(const unsigned long long*);
int mmfn_bind_rng_conv_tc(const unsigned long long*);
int mmfn_bind_rng_misc(const unsigned long long*);
int mmfn_bind_rng_pool(const unsigned long long*);
int mmfn_bind_rng_attn_tc(const unsigned long long*);
int mmfn_bind_rng_norm(const unsigned long long*);
int mmfn_bind_rng_attn_bf16(const unsigned long long*);
int mmfn_bind_rng_gpt_small(const unsigned long long*);

// Bind (or with null: unbind) a device-resident 64-bit offset that every dropout site adds to its
// seed at run time.  The training engine bumps it on the device once per step, so the dropout masks
// change between replays of a captured CUDA graph while forward/backward of one step stay consistent.
MMFN_API int mmfn_rng_bind(const unsigned long long* dev_offset) {
  int rc = mmfn_bind_rng_core(dev_offset);
  if (!rc) rc = mmfn_bind_rng_attn(dev_offset);
  if (!rc) rc = mmfn_bind_rng_gemm_tc(dev_offset);
  if (!rc) rc = mmfn_bind_rng_conv_tc(dev_offset);
  if (!rc) rc = mmfn_bind_rng_misc(dev_offset);
  if (!rc) rc = mmfn_bind_rng_pool(dev_offset);
  if (!rc) rc = mmfn_bind_rng_attn_tc(dev_offset);
  if (!rc) rc = mmfn_bind_rng_attn_bf16(dev_offset);
  if (!rc) rc = mmfn_bind_rng_gpt_small(dev_offset);
  if (!rc) rc = mmfn_bind_rng_norm(dev_offset);     // LayerNorm backward regenerates the residual-branch masks
  if (rc) mmfn_set_error("rng_bind: cudaMemcpyToSymbol failed (%d)", rc);
  return rc;
}
