// Generic fp32 SIMT GEMM with pluggable operand loaders (dense strided, im2col,
// transposed-conv gather) and a fused epilogue.  This is the exact-arithmetic
// path: it covers every shape on the MMFN step (odd K such as 7 or 162, strided
// head slices, implicit-GEMM convolutions) and is the numerical yardstick for the
// tcgen05 tensor-core kernels that take over the large aligned shapes.
#pragma once
#include "common.cuh"

namespace mmfn {

constexpr int BM = 64, BN = 64, BK = 16, GEMM_THREADS = 256;

struct Epilogue {
  float* C;            // output
  int64_t ldc;         // row stride of C (columns contiguous)
  int64_t c_b0, c_b1;  // batch strides
  const float* bias;   // [N] or null
  const float* res;    // same indexing as C, or null
  const float* mask;   // same indexing as C: out *= (mask > 0)
  float alpha;
  int act;             // 0 none, 1 relu
  int accum;           // 0 store, 1 C += v, 2 atomicAdd
  float drop_p;        // dropout applied to (alpha*acc + bias) before the residual
  uint64_t drop_seed;
};

// ---------------------------------------------------------------- loaders
struct DenseLoader {
  const float* p;
  int64_t s_r, s_k, s_b0, s_b1;
  int rows, ks;
  struct R { int64_t off; bool ok; };
  struct Kc { int64_t off; bool ok; };
  __device__ bool kfast() const { return s_k <= s_r; }
  __device__ const float* base(int b0, int b1) const { return p + b0 * s_b0 + b1 * s_b1; }
  __device__ R row(int r) const { return {r * s_r, r < rows}; }
  __device__ Kc kc(int k) const { return {k * s_k, k < ks}; }
  __device__ float load(const float* b, const R& r, const Kc& k) const {
    return (r.ok && k.ok) ? __ldg(b + r.off + k.off) : 0.f;
  }
};

struct ConvGeom {
  int N, H, W, C;      // input NHWC
  int R, S, stride, pad;
  int Ho, Wo, Co;      // output NHWC
};

// A(pix, tap) = X[n, ho*stride-pad+r, wo*stride-pad+s, c]; tap = (r*S+s)*C + c.
template <bool RowIsPix>
struct Im2colLoader {
  const float* x;
  ConvGeom g;
  struct P { int n, hb, wb; bool ok; };
  struct T { int r, s, c; bool ok; };
  using R = typename std::conditional<RowIsPix, P, T>::type;
  using Kc = typename std::conditional<RowIsPix, T, P>::type;
  __device__ bool kfast() const { return RowIsPix; }
  __device__ const float* base(int, int) const { return x; }
  __device__ P pix(int m) const {
    int hw = g.Ho * g.Wo;
    int n = m / hw, rem = m - n * hw;
    int ho = rem / g.Wo, wo = rem - ho * g.Wo;
    return {n, ho * g.stride - g.pad, wo * g.stride - g.pad, m < g.N * hw};
  }
  __device__ T tap(int k) const {
    int sc = g.S * g.C;
    int r = k / sc, rem = k - r * sc;
    int s = rem / g.C, c = rem - s * g.C;
    return {r, s, c, k < g.R * sc};
  }
  __device__ R row(int r) const { if constexpr (RowIsPix) return pix(r); else return tap(r); }
  __device__ Kc kc(int k) const { if constexpr (RowIsPix) return tap(k); else return pix(k); }
  __device__ float ld(const P& p, const T& t) const {
    int h = p.hb + t.r, w = p.wb + t.s;
    if (!(p.ok && t.ok) || h < 0 || h >= g.H || w < 0 || w >= g.W) return 0.f;
    return __ldg(x + (((int64_t)p.n * g.H + h) * g.W + w) * g.C + t.c);
  }
  __device__ float load(const float*, const R& r, const Kc& k) const {
    if constexpr (RowIsPix) return ld(r, k); else return ld(k, r);
  }
};

// Data-gradient gather: A(ipix, (r,s,co)) = dY[n, (h+pad-r)/stride, (w+pad-s)/stride, co]
// when the division is exact and in range, else 0.
struct DgradLoader {
  const float* dy;
  ConvGeom g;
  struct R { int n, hp, wp; bool ok; };
  struct Kc { int r, s, co; bool ok; };
  __device__ bool kfast() const { return true; }
  __device__ const float* base(int, int) const { return dy; }
  __device__ R row(int m) const {
    int hw = g.H * g.W;
    int n = m / hw, rem = m - n * hw;
    int h = rem / g.W, w = rem - h * g.W;
    return {n, h + g.pad, w + g.pad, m < g.N * hw};
  }
  __device__ Kc kc(int k) const {
    int sc = g.S * g.Co;
    int r = k / sc, rem = k - r * sc;
    int s = rem / g.Co, co = rem - s * g.Co;
    return {r, s, co, k < g.R * sc};
  }
  __device__ float load(const float*, const R& p, const Kc& t) const {
    if (!(p.ok && t.ok)) return 0.f;
    int hh = p.hp - t.r, ww = p.wp - t.s;
    if (hh < 0 || ww < 0) return 0.f;
    int ho = hh / g.stride, wo = ww / g.stride;
    if (ho * g.stride != hh || wo * g.stride != ww || ho >= g.Ho || wo >= g.Wo) return 0.f;
    return __ldg(dy + (((int64_t)p.n * g.Ho + ho) * g.Wo + wo) * g.Co + t.co);
  }
};

// ---------------------------------------------------------------- kernel
template <class AL, class BL>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_simt_kernel(AL A, BL B, int M, int N, int K, Epilogue e, int nb1, int splitk) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int z = blockIdx.z / splitk, ks = blockIdx.z - z * splitk;
  const int b0 = z / nb1, b1 = z - b0 * nb1;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const float* abase = A.base(b0, b1);
  const float* bbase = B.base(b0, b1);

  int kchunk = ((K + splitk - 1) / splitk + BK - 1) / BK * BK;
  int kbeg = ks * kchunk, kend = min(K, kbeg + kchunk);

  const bool akf = A.kfast(), bkf = B.kfast();
  // k-fast: thread owns k_local = tid%16 and rows tid/16 + 16*i; row-fast: row tid%64, k tid/64 + 4*i
  typename AL::R arow[4];
  typename BL::R brow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    arow[i] = A.row(m0 + (akf ? (tid >> 4) + 16 * i : (tid & 63)));
    brow[i] = B.row(n0 + (bkf ? (tid >> 4) + 16 * i : (tid & 63)));
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ty = tid >> 4, tx = tid & 15;
  // global -> register fetch of one BK-deep tile pair (issued one tile AHEAD of the math: the loads of tile k0 + BK
  // fly while tile k0 is multiplied out of shared memory, so the K loop no longer pays a full memory round trip per step)
  float av[4], bv[4];
  auto fetch = [&](int k0) {
    if (akf) {
      typename AL::Kc kc = A.kc(k0 + (tid & 15));
      kc.ok = kc.ok && (k0 + (tid & 15) < kend);
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = A.load(abase, arow[i], kc);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int k = k0 + (tid >> 6) + 4 * i;
        typename AL::Kc kc = A.kc(k);
        kc.ok = kc.ok && (k < kend);
        av[i] = A.load(abase, arow[0], kc);
      }
    }
    if (bkf) {
      typename BL::Kc kc = B.kc(k0 + (tid & 15));
      kc.ok = kc.ok && (k0 + (tid & 15) < kend);
#pragma unroll
      for (int i = 0; i < 4; ++i) bv[i] = B.load(bbase, brow[i], kc);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int k = k0 + (tid >> 6) + 4 * i;
        typename BL::Kc kc = B.kc(k);
        kc.ok = kc.ok && (k < kend);
        bv[i] = B.load(bbase, brow[0], kc);
      }
    }
  };
  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    __syncthreads();                                   // previous tile fully consumed
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (akf) As[tid & 15][(tid >> 4) + 16 * i] = av[i]; else As[(tid >> 6) + 4 * i][tid & 63] = av[i];
      if (bkf) Bs[tid & 15][(tid >> 4) + 16 * i] = bv[i]; else Bs[(tid >> 6) + 4 * i][tid & 63] = bv[i];
    }
    __syncthreads();
    if (k0 + BK < kend) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }

  const int64_t cb = (int64_t)b0 * e.c_b0 + (int64_t)b1 * e.c_b1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      int64_t idx = cb + (int64_t)m * e.ldc + n;
      float v = e.alpha * acc[i][j];
      if (e.bias && ks == 0) v += e.bias[n];
      if (e.act == 1) v = fmaxf(v, 0.f);
      if (e.mask) v = (e.mask[idx] > 0.f) ? v : 0.f;
      if (e.drop_p > 0.f) v *= mmfn_dropout_scale(e.drop_p, e.drop_seed, (uint64_t)idx);
      if (e.res && ks == 0) v += e.res[idx];
      if (e.accum == 0) e.C[idx] = v;
      else if (e.accum == 1) e.C[idx] += v;
      else atomicAdd(e.C + idx, v);
    }
  }
}

template <class AL, class BL>
static int launch_gemm_simt(const AL& A, const BL& B, int M, int N, int K, const Epilogue& e,
                            int nb0, int nb1, int splitk, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, nb0 * nb1 * splitk);
  gemm_simt_kernel<AL, BL><<<grid, GEMM_THREADS, 0, st>>>(A, B, M, N, K, e, nb1, splitk);
  return mmfn_launch_status("gemm_simt");
}

}  // namespace mmfn
