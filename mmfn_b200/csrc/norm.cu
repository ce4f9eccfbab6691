// Train-mode BatchNorm2d over NHWC row matrices (torchvision BasicBlock BNs; call sites
// model_rad.py:512-525, :542-544, :560-562, :577-579) and LayerNorm (+ReLU/GELU) rows
// (model_rad.py:117-118, :162, :253, :336, :346, :353).  HBM-bound kernels.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ BatchNorm
// Column (per-channel) reductions over the (M, C) NHWC row matrix, 4 channels per thread (float4 loads), fp64
// accumulation.  256 threads = QPR channel-quads x (256 / QPR) row lanes, QPR = min(C/4, 32); blockIdx.x selects
// the group of QPR quads, blockIdx.y a slab of `rows_per_block` rows.  BWD = false: sum x, sum x^2 (forward
// statistics).  BWD = true: sum dy', sum dy' * xhat with dy' = dy * (y > 0) when a ReLU followed.
//
// Scratch `ws` (doubles, zero on entry, left zero on exit -- no memset node per call):
//   [0, BN_SLOTS * 2C)  partial sums; CTA (.., y) adds into slot y % BN_SLOTS, so at most gridDim.y / BN_SLOTS
//                       fp64 atomics ever queue on one address (hundreds of CTAs hammering 2C addresses cost more
//                       than the reduction itself);
//   [BN_SLOTS * 2C]     arrival ticket: the LAST CTA to finish folds the slots, publishes the per-channel results
//                       as floats (forward: mean / rstd + running statistics; backward: mean(dy'), mean(dy' xhat)
//                       into ws_fin + dgamma / dbeta) and re-zeroes the scratch.
constexpr int BN_THREADS = 256;
constexpr int BN_SLOTS = 16;

struct BnFinal {
  float* mean; float* rstd; float* running_mean; float* running_var; float eps, momentum;   // forward
  float* fin; float* dgamma; float* dbeta;                                                  // backward
};

// ReLU mask of the BatchNorm backward: where the forward output was not positive the incoming gradient is dropped.
// The mask comes from the saved output y (fp32, or its bf16 twin: half the bytes, same sign), or -- when nothing was
// added between BatchNorm and ReLU -- is recomputed from z, which the backward reads anyway (no third stream at all);
// the expression is bn_apply_kernel's, term for term.
enum { BN_MASK_NONE = 0, BN_MASK_Y32 = 1, BN_MASK_Y16 = 2, BN_MASK_Z = 3 };
// The y value of the Y32 / Y16 modes is loaded by the caller next to its other loads (bn_mask_load) so that an unrolled
// loop keeps every load of its iterations in flight; MODE is a template parameter of the streaming kernels for the same
// reason (a run-time mode inside the loop serialised the loads: 16.5 -> 21.6 us for the layer-1 reduction).
template <int MODE>
__device__ __forceinline__ float4 bn_mask_load(const void* __restrict__ yout, int64_t i) {
  if (MODE == BN_MASK_Y32) return __ldg(reinterpret_cast<const float4*>(yout) + i);
  if (MODE == BN_MASK_Y16) return mmfn_unpack_bf16x4(__ldg(reinterpret_cast<const uint2*>(yout) + i));
  return make_float4(1.f, 1.f, 1.f, 1.f);
}
template <int MODE>
__device__ __forceinline__ void bn_mask_apply(float* ga, const float4 yv, const float* xa, const float* m4, const float* r4,
                                              const float* g4, const float* b4) {
  if (MODE == BN_MASK_Y32 || MODE == BN_MASK_Y16) {
    if (!(yv.x > 0.f)) ga[0] = 0.f;
    if (!(yv.y > 0.f)) ga[1] = 0.f;
    if (!(yv.z > 0.f)) ga[2] = 0.f;
    if (!(yv.w > 0.f)) ga[3] = 0.f;
  } else if (MODE == BN_MASK_Z) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (!((xa[k] - m4[k]) * r4[k] * g4[k] + b4[k] > 0.f)) ga[k] = 0.f;
  }
}
__device__ __forceinline__ void bn_relu_mask4(float* ga, int mode, const void* __restrict__ yout, int64_t i, const float* xa,
                                              const float* m4, const float* r4, const float* g4, const float* b4) {
  if (mode == BN_MASK_Y32) {
    const float4 yv = __ldg(reinterpret_cast<const float4*>(yout) + i);
    if (!(yv.x > 0.f)) ga[0] = 0.f;
    if (!(yv.y > 0.f)) ga[1] = 0.f;
    if (!(yv.z > 0.f)) ga[2] = 0.f;
    if (!(yv.w > 0.f)) ga[3] = 0.f;
  } else if (mode == BN_MASK_Y16) {
    const float4 yv = mmfn_unpack_bf16x4(__ldg(reinterpret_cast<const uint2*>(yout) + i));
    if (!(yv.x > 0.f)) ga[0] = 0.f;
    if (!(yv.y > 0.f)) ga[1] = 0.f;
    if (!(yv.z > 0.f)) ga[2] = 0.f;
    if (!(yv.w > 0.f)) ga[3] = 0.f;
  } else if (mode == BN_MASK_Z) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (!((xa[k] - m4[k]) * r4[k] * g4[k] + b4[k] > 0.f)) ga[k] = 0.f;
  }
}

// Block-level tail shared by the column-reduction kernels: fold the per-thread fp64 partials (thread = quad q of
// `qpr`, row lane threadIdx.x / qpr) over the row lanes, add them into scratch slot `slot`, and let the LAST CTA of the
// launch (`n_ctas` arrivals) fold the slots, publish and re-zero the scratch.
template <bool BWD>
__device__ __forceinline__ void bn_reduce_publish(const double (&acc)[8], int qpr, int quad0, int slot, unsigned n_ctas,
                                                  int64_t M, int C4, double* __restrict__ ws, const BnFinal& fz) {
  __shared__ double red[BN_THREADS][8];
  __shared__ bool last_cta;
  const int nrs = BN_THREADS / qpr;
  const int C = C4 * 4;
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x][k] = acc[k];
  __syncthreads();
  // thread t < qpr * 8 reduces one (quad, component) over the row lanes
  if (threadIdx.x < qpr * 8) {
    const int qq = threadIdx.x >> 3, k = threadIdx.x & 7;
    const int cqq = quad0 + qq;
    if (cqq < C4) {
      double t = 0.0;
      for (int rs = 0; rs < nrs; ++rs) t += red[rs * qpr + qq][k];
      double* sl = ws + (int64_t)slot * 2 * C;
      atomicAdd(sl + (k < 4 ? 0 : C) + cqq * 4 + (k & 3), t);
    }
  }
  // ---- last CTA: fold the slots, publish, re-zero
  unsigned int* ticket = reinterpret_cast<unsigned int*>(ws + (int64_t)BN_SLOTS * 2 * C);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last_cta = atomicAdd(ticket, 1u) == n_ctas - 1;
  __syncthreads();
  if (!last_cta) return;
  __threadfence();
  const double invM = 1.0 / (double)M;
  for (int c = threadIdx.x; c < C; c += BN_THREADS) {
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int sl = 0; sl < BN_SLOTS; ++sl) {
      double* slotp = ws + (int64_t)sl * 2 * C;
      s1 += __ldcg(slotp + c);
      s2 += __ldcg(slotp + C + c);
      slotp[c] = 0.0;
      slotp[C + c] = 0.0;
    }
    if (!BWD) {
      const double md = s1 * invM;
      double var = s2 * invM - md * md;
      if (var < 0.0) var = 0.0;
      fz.mean[c] = (float)md;
      fz.rstd[c] = (float)(1.0 / sqrt(var + (double)fz.eps));
      if (fz.running_mean) {
        const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
        fz.running_mean[c] = (float)((1.0 - fz.momentum) * fz.running_mean[c] + fz.momentum * md);
        fz.running_var[c] = (float)((1.0 - fz.momentum) * fz.running_var[c] + fz.momentum * unbiased);
      }
    } else {
      fz.fin[c] = (float)(s1 * invM);
      fz.fin[C + c] = (float)(s2 * invM);
      fz.dbeta[c] += (float)s1;
      fz.dgamma[c] += (float)s2;
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

template <bool BWD, int MODE>
__global__ void __launch_bounds__(BN_THREADS)
bn_colsum_kernel(const float4* __restrict__ x, const float4* __restrict__ dy, const void* __restrict__ yout,
                 const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int64_t M, int C4, int qpr,
                 int64_t rows_per_block, double* __restrict__ ws, BnFinal fz) {
  const int q = threadIdx.x % qpr, rsub = threadIdx.x / qpr, nrs = BN_THREADS / qpr;
  const int cq = blockIdx.x * qpr + q;                          // channel quad of this thread
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cq < C4) {
    float m4[4] = {0.f, 0.f, 0.f, 0.f}, r4[4] = {1.f, 1.f, 1.f, 1.f}, g4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
    if (BWD) {
      const float4 mm = __ldg(reinterpret_cast<const float4*>(mean) + cq), rr = __ldg(reinterpret_cast<const float4*>(rstd) + cq);
      m4[0] = mm.x; m4[1] = mm.y; m4[2] = mm.z; m4[3] = mm.w;
      r4[0] = rr.x; r4[1] = rr.y; r4[2] = rr.z; r4[3] = rr.w;
      if (MODE == BN_MASK_Z) {
        const float4 gg = __ldg(reinterpret_cast<const float4*>(gamma) + cq), bb = __ldg(reinterpret_cast<const float4*>(beta) + cq);
        g4[0] = gg.x; g4[1] = gg.y; g4[2] = gg.z; g4[3] = gg.w;
        b4[0] = bb.x; b4[1] = bb.y; b4[2] = bb.z; b4[3] = bb.w;
      }
    }
#pragma unroll 4
    for (int64_t r = r0 + rsub; r < r1; r += nrs) {
      const float4 xv = __ldg(x + r * C4 + cq);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
      if (!BWD) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[k] += xa[k]; acc[4 + k] += (double)xa[k] * xa[k]; }
      } else {
        const float4 gv = __ldg(dy + r * C4 + cq);
        const float4 yv = bn_mask_load<MODE>(yout, r * C4 + cq);
        float ga[4] = {gv.x, gv.y, gv.z, gv.w};
        bn_mask_apply<MODE>(ga, yv, xa, m4, r4, g4, b4);
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[k] += ga[k]; acc[4 + k] += (double)ga[k] * ((xa[k] - m4[k]) * r4[k]); }
      }
    }
  }
  bn_reduce_publish<BWD>(acc, qpr, blockIdx.x * qpr, blockIdx.y % BN_SLOTS, gridDim.x * gridDim.y, M, C4, ws, fz);
}

static inline void bn_colsum_grid(int64_t M, int C, int& qpr, dim3& grid, int64_t& rows_per_block) {
  const int C4 = C / 4;
  qpr = C4 < 32 ? C4 : 32;
  while (BN_THREADS % qpr) --qpr;                                // C4 = 16, 32, 64, 128 here; stay safe for odd widths
  const int groups = (C4 + qpr - 1) / qpr;
  const int nrs = BN_THREADS / qpr;
  int64_t slabs = ceil_div64(M, (int64_t)nrs * 8);               // at least 8 rows per thread
  const int64_t cap = (148 * 4 + groups - 1) / groups;
  if (slabs > cap) slabs = cap;
  if (slabs < 1) slabs = 1;
  rows_per_block = ceil_div64(M, slabs);
  grid = dim3((unsigned)groups, (unsigned)ceil_div64(M, rows_per_block));
}
static inline float* bn_ws_fin(double* ws, int C) { return reinterpret_cast<float*>(ws + (int64_t)BN_SLOTS * 2 * C + 2); }

// y = (x - mean) * rstd * gamma + beta (+res, ReLU) from the published per-channel mean / rstd.  The grid stride is a
// multiple of C/4 (power-of-two widths), so a thread keeps ONE channel quad: its coefficients live in registers and
// the loop body is pure streaming (no per-element index division).
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float4* __restrict__ x, float4* __restrict__ y, int64_t n4, int C4,
                const float* __restrict__ mean, const float* __restrict__ rstd,
                const float4* __restrict__ gamma, const float4* __restrict__ beta,
                const float4* __restrict__ res, int relu, uint2* __restrict__ y16) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool fixed = stride % C4 == 0;
  int c = (int)(i0 % C4);
  float4 m, r, g, b;
  auto coef = [&](int cq) {
    m = __ldg(reinterpret_cast<const float4*>(mean) + cq);
    r = __ldg(reinterpret_cast<const float4*>(rstd) + cq);
    g = __ldg(gamma + cq);
    b = __ldg(beta + cq);
  };
  coef(c);
#pragma unroll 4
  for (int64_t i = i0; i < n4; i += stride) {
    if (!fixed) { c = (int)(i % C4); coef(c); }
    const float4 v = __ldg(x + i);
    float4 o;
    o.x = (v.x - m.x) * r.x * g.x + b.x;
    o.y = (v.y - m.y) * r.y * g.y + b.y;
    o.z = (v.z - m.z) * r.z * g.z + b.z;
    o.w = (v.w - m.w) * r.w * g.w + b.w;
    if (res) { const float4 q = __ldg(res + i); o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    if (y) y[i] = o;                                              // null: only the bf16 twin is wanted
    if (y16) y16[i] = mmfn_pack_bf16x4(o.x, o.y, o.z, o.w);       // bf16 twin: the operand of the next convolution
  }
}

// dx = gamma * rstd * (dy' - mean(dy') - xhat * mean(dy' * xhat)); dres = dy' (residual branch).  float4 over channels;
// fin = [mean(dy') | mean(dy' xhat)] published by the reduction kernel.  Same fixed-channel-quad streaming loop.
template <int MODE>
__global__ void __launch_bounds__(256)
bn_bwd_dx_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                 const void* __restrict__ yout, const float* __restrict__ mean,
                 const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                 const float* __restrict__ fin, int64_t M, int C4,
                 float4* __restrict__ dx, float4* __restrict__ dres, uint2* __restrict__ dx16) {
  const int C = C4 * 4;
  const int64_t n4 = M * C4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool fixed = stride % C4 == 0;
  int c = (int)(i0 % C4);
  float4 m, r, k, a, b;
  float m4[4], r4[4], g4[4], b4[4];
  auto coef = [&](int cq) {
    m = __ldg(reinterpret_cast<const float4*>(mean) + cq);
    r = __ldg(reinterpret_cast<const float4*>(rstd) + cq);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + cq);
    k = make_float4(g.x * r.x, g.y * r.y, g.z * r.z, g.w * r.w);
    a = __ldg(reinterpret_cast<const float4*>(fin) + cq);
    b = __ldg(reinterpret_cast<const float4*>(fin + C) + cq);
    m4[0] = m.x; m4[1] = m.y; m4[2] = m.z; m4[3] = m.w;
    r4[0] = r.x; r4[1] = r.y; r4[2] = r.z; r4[3] = r.w;
    g4[0] = g.x; g4[1] = g.y; g4[2] = g.z; g4[3] = g.w;
    if (MODE == BN_MASK_Z) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(beta) + cq);
      b4[0] = bb.x; b4[1] = bb.y; b4[2] = bb.z; b4[3] = bb.w;
    } else {
      b4[0] = b4[1] = b4[2] = b4[3] = 0.f;
    }
  };
  coef(c);
#pragma unroll 4
  for (int64_t i = i0; i < n4; i += stride) {
    if (!fixed) { c = (int)(i % C4); coef(c); }
    const float4 gv = __ldg(dy + i);
    const float4 xv = __ldg(x + i);
    const float4 yv = bn_mask_load<MODE>(yout, i);
    float4 g;
    {
      float ga[4] = {gv.x, gv.y, gv.z, gv.w};
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
      bn_mask_apply<MODE>(ga, yv, xa, m4, r4, g4, b4);
      g = make_float4(ga[0], ga[1], ga[2], ga[3]);
    }
    float4 o;
    o.x = k.x * (g.x - a.x - (xv.x - m.x) * r.x * b.x);
    o.y = k.y * (g.y - a.y - (xv.y - m.y) * r.y * b.y);
    o.z = k.z * (g.z - a.z - (xv.z - m.z) * r.z * b.z);
    o.w = k.w * (g.w - a.w - (xv.w - m.w) * r.w * b.w);
    if (dx16) dx16[i] = mmfn_pack_bf16x4(o.x, o.y, o.z, o.w);     // dz only feeds the wgrad / dgrad MMAs
    else dx[i] = o;
    if (dres) dres[i] = g;
  }
}

// ---- small feature maps (M <= BN_SMALL_ROWS = 1024 rows: layer 4 at B=16): ONE launch.  Such tensors (<= 4 MB) live in L2, so
// the two-kernel scheme is pure launch/tail latency.  One CTA owns one channel quad: pass 1 reduces its column over
// all rows, the CTA publishes the statistics, pass 2 re-reads the column (L2/L1 hits) and writes the result.
template <bool BWD>
__global__ void __launch_bounds__(256)
bn_small_kernel(const float4* __restrict__ x, const float4* __restrict__ dy, const void* __restrict__ yout, int mask_mode,
                const float4* __restrict__ res, float4* __restrict__ out, float4* __restrict__ dres,
                float* __restrict__ mean, float* __restrict__ rstd, const float4* __restrict__ gamma,
                const float4* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var,
                float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int C4, float eps, float momentum, int relu,
                uint2* __restrict__ out16) {
  __shared__ double red[8][8];
  __shared__ float fin[8];
  const int cq = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m4[4] = {0.f, 0.f, 0.f, 0.f}, r4[4] = {1.f, 1.f, 1.f, 1.f}, gz4[4] = {0.f, 0.f, 0.f, 0.f}, bz4[4] = {0.f, 0.f, 0.f, 0.f};
  if (BWD) {
    const float4 mm = __ldg(reinterpret_cast<const float4*>(mean) + cq), rr = __ldg(reinterpret_cast<const float4*>(rstd) + cq);
    m4[0] = mm.x; m4[1] = mm.y; m4[2] = mm.z; m4[3] = mm.w;
    r4[0] = rr.x; r4[1] = rr.y; r4[2] = rr.z; r4[3] = rr.w;
    if (mask_mode == BN_MASK_Z) {
      const float4 gg = __ldg(gamma + cq), bb = __ldg(beta + cq);
      gz4[0] = gg.x; gz4[1] = gg.y; gz4[2] = gg.z; gz4[3] = gg.w;
      bz4[0] = bb.x; bz4[1] = bb.y; bz4[2] = bb.z; bz4[3] = bb.w;
    }
  }
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 8
  for (int r = threadIdx.x; r < M; r += 256) {
    const float4 xv = __ldg(x + (int64_t)r * C4 + cq);
    const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
    if (!BWD) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { acc[k] += xa[k]; acc[4 + k] += (double)xa[k] * xa[k]; }
    } else {
      const float4 gv = __ldg(dy + (int64_t)r * C4 + cq);
      float ga[4] = {gv.x, gv.y, gv.z, gv.w};
      bn_relu_mask4(ga, mask_mode, yout, (int64_t)r * C4 + cq, xa, m4, r4, gz4, bz4);
#pragma unroll
      for (int k = 0; k < 4; ++k) { acc[k] += ga[k]; acc[4 + k] += (double)ga[k] * ((xa[k] - m4[k]) * r4[k]); }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[warp][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const int k = threadIdx.x, c = cq * 4 + k;
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { s1 += red[w][k]; s2 += red[w][4 + k]; }
    const double invM = 1.0 / (double)M;
    if (!BWD) {
      const double md = s1 * invM;
      double var = s2 * invM - md * md;
      if (var < 0.0) var = 0.0;
      const float mf = (float)md, rf = (float)(1.0 / sqrt(var + (double)eps));
      mean[c] = mf; rstd[c] = rf;
      fin[k] = mf; fin[4 + k] = rf;
      if (running_mean) {
        const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * md);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
      }
    } else {
      fin[k] = (float)(s1 * invM);
      fin[4 + k] = (float)(s2 * invM);
      dbeta[c] += (float)s1;
      dgamma[c] += (float)s2;
    }
  }
  __syncthreads();
  const float4 g = __ldg(gamma + cq);
  const float ga4[4] = {g.x, g.y, g.z, g.w};
  if (!BWD) {
    const float4 b = __ldg(beta + cq);
    const float ba[4] = {b.x, b.y, b.z, b.w};
#pragma unroll 8
    for (int r = threadIdx.x; r < M; r += 256) {
      const int64_t i = (int64_t)r * C4 + cq;
      const float4 v = __ldg(x + i);
      float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = (o[k] - fin[k]) * fin[4 + k] * ga4[k] + ba[k];
      if (res) { const float4 q = __ldg(res + i); o[0] += q.x; o[1] += q.y; o[2] += q.z; o[3] += q.w; }
      if (relu) {
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = fmaxf(o[k], 0.f);
      }
      out[i] = make_float4(o[0], o[1], o[2], o[3]);
      if (out16) out16[i] = mmfn_pack_bf16x4(o[0], o[1], o[2], o[3]);
    }
  } else {
    float kk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) kk[k] = ga4[k] * r4[k];
#pragma unroll 8
    for (int r = threadIdx.x; r < M; r += 256) {
      const int64_t i = (int64_t)r * C4 + cq;
      const float4 gv = __ldg(dy + i);
      float gg[4] = {gv.x, gv.y, gv.z, gv.w};
      const float4 xv = __ldg(x + i);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
      bn_relu_mask4(gg, mask_mode, yout, i, xa, m4, r4, gz4, bz4);
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = kk[k] * (gg[k] - fin[k] - (xa[k] - m4[k]) * r4[k] * fin[4 + k]);
      if (out16) out16[i] = mmfn_pack_bf16x4(o[0], o[1], o[2], o[3]);
      else out[i] = make_float4(o[0], o[1], o[2], o[3]);
      if (dres) dres[i] = make_float4(gg[0], gg[1], gg[2], gg[3]);
    }
  }
}
// feature maps with at most this many rows take the single-launch kernel (mmfn_set_bn_small_rows; default 1024: at 2048 rows the two-kernel path with statistics from the convolution epilogue measured +1.7 % on the bf16 B = 32 step, 4096 rows on the single-launch kernel -4.5 % at TF32 B = 16)
static int BN_SMALL_ROWS = 1024;

// ------------------------------------------------------------------ stem tail: BatchNorm -> ReLU -> MaxPool 3x3 / 2
// The two ResNet stems end in bn1 -> relu -> maxpool (model_rad.py:512-521) on the largest activation of the network
// (B x 128 x 128 x 64).  Unfused that is three passes forward (apply writes y, the pool reads it back) and three
// backward (pool backward writes a 3/4-zero gradient map, the BatchNorm reduction and the dx pass each read it again).
// Fused: the forward reads z once and writes only the pooled map; y is never materialised.  The arg-max byte carries
// the ReLU mask (bit 7 set: the window maximum was not positive), so the backward needs neither y nor the pooled map.

// one thread = one pooled pixel x 4 channels; all nine taps are independent loads
__global__ void __launch_bounds__(256)
bn_relu_maxpool_fwd_kernel(const float* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ rstd,
                           const float4* __restrict__ gamma, const float4* __restrict__ beta, float* __restrict__ out,
                           uint2* __restrict__ out16, uint8_t* __restrict__ idx, float4* __restrict__ zmax,
                           int B, int H, int W, int C, int Ho, int Wo) {
  const int C4 = C >> 2;
  const int64_t n = (int64_t)B * Ho * Wo * C4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool fixed = stride % C4 == 0;
  int cq = (int)(i0 % C4);
  // (v - m) * r * g + b is evaluated in the order bn_apply_kernel uses: rstd and gamma stay separate factors
  float m[4], k[4], gk[4], bt[4];
  auto coef = [&](int c) {
    const float4 mm = __ldg(reinterpret_cast<const float4*>(mean) + c), rr = __ldg(reinterpret_cast<const float4*>(rstd) + c);
    const float4 g = __ldg(gamma + c), b = __ldg(beta + c);
    m[0] = mm.x; m[1] = mm.y; m[2] = mm.z; m[3] = mm.w;
    k[0] = rr.x; k[1] = rr.y; k[2] = rr.z; k[3] = rr.w;
    gk[0] = g.x; gk[1] = g.y; gk[2] = g.z; gk[3] = g.w;
    bt[0] = b.x; bt[1] = b.y; bt[2] = b.z; bt[3] = b.w;
  };
  coef(cq);
  for (int64_t i = i0; i < n; i += stride) {
    if (!fixed) { cq = (int)(i % C4); coef(cq); }
    const int t = (int)(i / C4);                  // (b, ho, wo) flattened: < 2^31 pixels
    const int wo = t % Wo, t2 = t / Wo;
    const int ho = t2 % Ho, b = t2 / Ho;
    float4 v[9];
    bool ok[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int sx = 0; sx < 3; ++sx) {
        const int h = ho * 2 - 1 + r, w = wo * 2 - 1 + sx;
        ok[r * 3 + sx] = h >= 0 && h < H && w >= 0 && w < W;
        const int hc = min(max(h, 0), H - 1), wc = min(max(w, 0), W - 1);
        v[r * 3 + sx] = __ldg(reinterpret_cast<const float4*>(z + (((int64_t)b * H + hc) * W + wc) * C) + cq);
      }
    }
    float best[4] = {0.f, 0.f, 0.f, 0.f}, zb[4] = {0.f, 0.f, 0.f, 0.f};
    int bi[4] = {-1, -1, -1, -1};
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      if (!ok[tp]) continue;
      const float a[4] = {v[tp].x, v[tp].y, v[tp].z, v[tp].w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float y = fmaxf((a[c] - m[c]) * k[c] * gk[c] + bt[c], 0.f);
        if (bi[c] < 0 || y > best[c] || y != y) { best[c] = y; bi[c] = tp; zb[c] = a[c]; }
      }
    }
    *reinterpret_cast<float4*>(out + (i << 2)) = make_float4(best[0], best[1], best[2], best[3]);
    if (zmax) zmax[i] = make_float4(zb[0], zb[1], zb[2], zb[3]);
    if (out16) out16[i] = mmfn_pack_bf16x4(best[0], best[1], best[2], best[3]);
    uint32_t code = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) code |= ((uint32_t)bi[c] | (best[c] > 0.f ? 0u : 0x80u)) << (8 * c);
    *reinterpret_cast<uint32_t*>(idx + (i << 2)) = code;
  }
}

// Reduction half of the backward.  The gradient of the BatchNorm output is non-zero only at the arg-max position of
// each pooling window, so sum(dy') and sum(dy' * xhat) run over the POOLED pixels: rows = (b, ho, wo), the thread layout
// of bn_colsum_kernel.  The forward saved z at the arg-max (`zmax`), so this is a pure stream over three pooled-size
// tensors (a gather of z at the arg-max tap measured 84 us at B = 32: 4-byte loads of 32-byte sectors, dependent on
// the arg-max load).
__global__ void __launch_bounds__(BN_THREADS)
stem_bwd_stats_kernel(const float4* __restrict__ dout, const uint8_t* __restrict__ idx, const float4* __restrict__ zmax,
                      const float* __restrict__ mean, const float* __restrict__ rstd, int C4,
                      int64_t Mp, int64_t M, int qpr, int64_t rows_per_block, double* __restrict__ ws, BnFinal fz) {
  const int q = threadIdx.x % qpr, rsub = threadIdx.x / qpr, nrs = BN_THREADS / qpr;
  const int cq = blockIdx.x * qpr + q;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(Mp, r0 + rows_per_block);
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cq < C4) {
    const float4 mm = __ldg(reinterpret_cast<const float4*>(mean) + cq), rr = __ldg(reinterpret_cast<const float4*>(rstd) + cq);
    const float m4[4] = {mm.x, mm.y, mm.z, mm.w}, r4[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll 4
    for (int64_t r = r0 + rsub; r < r1; r += nrs) {
      const uint32_t code = __ldg(reinterpret_cast<const uint32_t*>(idx) + r * C4 + cq);
      const float4 gv = __ldg(dout + r * C4 + cq);
      const float4 zv = __ldg(zmax + r * C4 + cq);
      const float ga[4] = {gv.x, gv.y, gv.z, gv.w}, za[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if ((code >> (8 * k)) & 0x80u) continue;                    // ReLU mask: the maximum was not positive
        acc[k] += ga[k];
        acc[4 + k] += (double)ga[k] * ((za[k] - m4[k]) * r4[k]);
      }
    }
  }
  bn_reduce_publish<true>(acc, qpr, blockIdx.x * qpr, blockIdx.y % BN_SLOTS, gridDim.x * gridDim.y, M, C4, ws, fz);
}

// dx half: one thread owns a 2 x 2 patch of input positions x 4 channels.  The patch at (2hp, 2wp) is covered by the
// four windows (hp + a, wp + c), a, c in {0, 1}: position (dh, dw) of the patch is tap (dh - 2a + 1, dw - 2c + 1) of
// window (a, c) when both are in 0..2.  4 arg-max words + 4 gradient quads + 4 z quads, all independent loads.
__global__ void __launch_bounds__(256)
stem_bwd_dx_kernel(const float* __restrict__ dout, const uint8_t* __restrict__ idx, const float* __restrict__ z,
                   const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                   const float* __restrict__ fin, int B, int H, int W, int C, int Ho, int Wo,
                   float* __restrict__ dz, uint2* __restrict__ dz16) {
  const int C4 = C >> 2;
  const int Hp = (H + 1) >> 1, Wp = (W + 1) >> 1;
  const int64_t n = (int64_t)B * Hp * Wp * C4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool fixed = stride % C4 == 0;
  int cq = (int)(i0 % C4);
  float m[4], r[4], kk[4], fa[4], fb[4];
  auto coef = [&](int c) {
    const float4 mm = __ldg(reinterpret_cast<const float4*>(mean) + c), rr = __ldg(reinterpret_cast<const float4*>(rstd) + c);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
    const float4 a = __ldg(reinterpret_cast<const float4*>(fin) + c), b = __ldg(reinterpret_cast<const float4*>(fin + C) + c);
    m[0] = mm.x; m[1] = mm.y; m[2] = mm.z; m[3] = mm.w;
    r[0] = rr.x; r[1] = rr.y; r[2] = rr.z; r[3] = rr.w;
    kk[0] = g.x * rr.x; kk[1] = g.y * rr.y; kk[2] = g.z * rr.z; kk[3] = g.w * rr.w;
    fa[0] = a.x; fa[1] = a.y; fa[2] = a.z; fa[3] = a.w;
    fb[0] = b.x; fb[1] = b.y; fb[2] = b.z; fb[3] = b.w;
  };
  coef(cq);
  for (int64_t i = i0; i < n; i += stride) {
    if (!fixed) { cq = (int)(i % C4); coef(cq); }
    const int t = (int)(i / C4);
    const int wp = t % Wp, t2 = t / Wp;
    const int hp = t2 % Hp, b = t2 / Hp;
    uint32_t code[4];
    float4 g4[4], z4[4];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int ho = hp + a, wo = wp + c;
        const bool ok = ho < Ho && wo < Wo;
        const int64_t o = (((int64_t)b * Ho + min(ho, Ho - 1)) * Wo + min(wo, Wo - 1)) * C4 + cq;
        const uint32_t cd = __ldg(reinterpret_cast<const uint32_t*>(idx) + o);
        code[a * 2 + c] = ok ? cd : 0xffffffffu;                       // 0xff never equals a tap
        g4[a * 2 + c] = __ldg(reinterpret_cast<const float4*>(dout) + o);
        const int h = min(2 * hp + a, H - 1), w = min(2 * wp + c, W - 1);
        z4[a * 2 + c] = __ldg(reinterpret_cast<const float4*>(z + (((int64_t)b * H + h) * W + w) * C) + cq);
      }
    }
#pragma unroll
    for (int dh = 0; dh < 2; ++dh) {
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        const int h = 2 * hp + dh, w = 2 * wp + dw;
        if (h >= H || w >= W) continue;
        float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int a = 0; a < 2; ++a) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int tr = dh - 2 * a + 1, tc = dw - 2 * c + 1;
            if (tr < 0 || tc < 0) continue;                          // compile-time: this window does not cover the position
            const uint32_t tap = (uint32_t)(tr * 3 + tc);
            const uint32_t cd = code[a * 2 + c];
            const float4 d = g4[a * 2 + c];
            if ((cd & 0xffu) == tap) g[0] += d.x;
            if (((cd >> 8) & 0xffu) == tap) g[1] += d.y;
            if (((cd >> 16) & 0xffu) == tap) g[2] += d.z;
            if ((cd >> 24) == tap) g[3] += d.w;
          }
        }
        const float4 zv = z4[dh * 2 + dw];
        const float za[4] = {zv.x, zv.y, zv.z, zv.w};
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) o[c] = kk[c] * (g[c] - fa[c] - (za[c] - m[c]) * r[c] * fb[c]);
        const int64_t e = (((int64_t)b * H + h) * W + w) * C4 + cq;
        if (dz16) dz16[e] = mmfn_pack_bf16x4(o[0], o[1], o[2], o[3]);
        else reinterpret_cast<float4*>(dz)[e] = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// ------------------------------------------------------------------ LayerNorm
__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  return v;
}
__device__ __forceinline__ float act_grad(float v, int act) {
  if (act == 1) return v > 0.f ? 1.f : 0.f;
  if (act == 2) {
    float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
    float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
    return cdf + v * pdf;
  }
  return 1.f;
}

// one warp per row
__global__ void ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float* __restrict__ y,
                              float* __restrict__ mean, float* __restrict__ rstd,
                              int64_t M, int C, float eps, int act, __nv_bfloat16* __restrict__ y16) {
  int lane = threadIdx.x & 31;
  int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* xr = x + row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  float mu = warp_sum(s) / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { float d = xr[c] - mu; v += d * d; }
  float rs = rsqrtf(warp_sum(v) / (float)C + eps);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  if (y16) {                                   // bf16 result: the LayerNorm output only feeds GEMMs
    __nv_bfloat16* yr = y16 + row * C;
    for (int c = lane; c < C; c += 32) yr[c] = __float2bfloat16_rn(act_fwd((xr[c] - mu) * rs * gamma[c] + beta[c], act));
    return;
  }
  float* yr = y + row * C;
  for (int c = lane; c < C; c += 32) yr[c] = act_fwd((xr[c] - mu) * rs * gamma[c] + beta[c], act);
}

// dx = LN'(dy * act'(ln)) (+ dres).  One warp per row, float4 columns (C % 4 == 0, C <= 128 * NV).
template <int NV>
__global__ void ln_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                 const float* __restrict__ dres, float* __restrict__ dx, int64_t M, int C, int act,
                                 float* __restrict__ dx_drop, float drop_p, uint64_t drop_seed,
                                 __nv_bfloat16* __restrict__ dx_drop16) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float mu = mean[row], rs = rstd[row];
  float gv[NV][4], xh[NV][4];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (lane + 32 * j) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) { gv[j][k] = 0.f; xh[j][k] = 0.f; }
    if (c < C) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + row * C + c));
      const float4 gy = __ldg(reinterpret_cast<const float4*>(dy + row * C + c));
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ga[4] = {gy.x, gy.y, gy.z, gy.w}, gma[4] = {gm.x, gm.y, gm.z, gm.w};
      float ba[4] = {0.f, 0.f, 0.f, 0.f};
      if (act) { const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + c)); ba[0] = bt.x; ba[1] = bt.y; ba[2] = bt.z; ba[3] = bt.w; }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        xh[j][k] = (xa[k] - mu) * rs;
        float g = ga[k];
        if (act) g *= act_grad(xh[j][k] * gma[k] + ba[k], act);
        gv[j][k] = g * gma[k];
        s1 += gv[j][k];
        s2 += gv[j][k] * xh[j][k];
      }
    }
  }
  s1 = warp_sum(s1) / (float)C;
  s2 = warp_sum(s2) / (float)C;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (lane + 32 * j) * 4;
    if (c < C) {
      float4 o;
      o.x = rs * (gv[j][0] - s1 - xh[j][0] * s2);
      o.y = rs * (gv[j][1] - s1 - xh[j][1] * s2);
      o.z = rs * (gv[j][2] - s1 - xh[j][2] * s2);
      o.w = rs * (gv[j][3] - s1 - xh[j][3] * s2);
      if (dres) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(dres + row * C + c));
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      *reinterpret_cast<float4*>(dx + row * C + c) = o;
      if (dx_drop || dx_drop16) {   // gradient entering the next residual branch's dropout (same hash as the forward mask)
        const uint64_t i0 = (uint64_t)(row * C + c);
        float ds[4];
        mmfn_dropout_scale4(drop_p, drop_seed, i0, ds);          // i0 % 4 == 0 (C % 4 == 0); all ones when drop_p == 0
        o.x *= ds[0]; o.y *= ds[1]; o.z *= ds[2]; o.w *= ds[3];
        if (dx_drop16) *reinterpret_cast<uint2*>(dx_drop16 + row * C + c) = mmfn_pack_bf16x4(o.x, o.y, o.z, o.w);
        else *reinterpret_cast<float4*>(dx_drop + row * C + c) = o;
      }
    }
  }
}

// dgamma[c] += sum_rows dy' * xhat, dbeta[c] += sum_rows dy'  (dy' = dy * act'(ln)).
// blockDim (32 columns, 8 row lanes); grid (C/32, row slabs).
__global__ void ln_bwd_param_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t M, int C, int act) {
  __shared__ float s1[8][33], s2[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int64_t rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float a = 0.f, b = 0.f;
  if (c < C) {
    const float gm = gamma[c], bt = beta[c];
#pragma unroll 4
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const float xh = (__ldg(x + r * C + c) - __ldg(mean + r)) * __ldg(rstd + r);
      float g = __ldg(dy + r * C + c);
      if (act) g *= act_grad(xh * gm + bt, act);
      a += g * xh;
      b += g;
    }
  }
  s1[threadIdx.y][threadIdx.x] = a;
  s2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; ++i) { a += s1[i][threadIdx.x]; b += s2[i][threadIdx.x]; }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
  }
}

// ---- C == 64 (VectorNet sub-graph rows: 622 592 of them in BASELINE configs[4]).  The warp-per-row kernels above leave
// half their lanes idle at 64 channels and keep two 256-byte loads in flight per warp; here a HALF-warp owns a row
// (16 lanes x float4) and every thread works on RPT rows at once, so 2 RPT 16-byte loads are in flight per thread.
// The half-warp butterfly (offsets 8, 4, 2, 1) adds in the order warp_sum does when lanes 16..31 hold zeros: results
// are bit-identical to ln_bwd_dx_kernel<1>.
template <int RPT>
__global__ void __launch_bounds__(256)
ln_bwd_dx_c64_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ dres, float* __restrict__ dx, int64_t M, int act) {
  constexpr int C = 64;
  const int q = threadIdx.x & 15;                            // channel quad
  // the loop bound is WARP-uniform (both half-warps take every trip; rows past the end are clamped on load and not
  // stored): the full-mask shuffles below need all 32 lanes
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int hsel = (threadIdx.x >> 4) & 1;
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + q);
  const float gma[4] = {gm.x, gm.y, gm.z, gm.w};
  float ba[4] = {0.f, 0.f, 0.f, 0.f};
  if (act) { const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + q); ba[0] = bt.x; ba[1] = bt.y; ba[2] = bt.z; ba[3] = bt.w; }
  for (int64_t base = warp * 2 * RPT; base < M; base += nwarp * 2 * RPT) {
    const int64_t r0 = base + hsel * RPT;
    float4 xv[RPT], gy[RPT];
    float mu[RPT], rs[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int64_t row = min(r0 + j, M - 1);
      xv[j] = __ldg(reinterpret_cast<const float4*>(x + row * C) + q);
      gy[j] = __ldg(reinterpret_cast<const float4*>(dy + row * C) + q);
      mu[j] = __ldg(mean + row);
      rs[j] = __ldg(rstd + row);
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const float xa[4] = {xv[j].x, xv[j].y, xv[j].z, xv[j].w}, ga[4] = {gy[j].x, gy[j].y, gy[j].z, gy[j].w};
      float gv[4], xh[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        xh[k] = (xa[k] - mu[j]) * rs[j];
        float g = ga[k];
        if (act) g *= act_grad(xh[k] * gma[k] + ba[k], act);
        gv[k] = g * gma[k];
        s1 += gv[k];
        s2 += gv[k] * xh[k];
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
      s1 /= (float)C;
      s2 /= (float)C;
      if (r0 + j < M) {
        float4 o4;
        o4.x = rs[j] * (gv[0] - s1 - xh[0] * s2);
        o4.y = rs[j] * (gv[1] - s1 - xh[1] * s2);
        o4.z = rs[j] * (gv[2] - s1 - xh[2] * s2);
        o4.w = rs[j] * (gv[3] - s1 - xh[3] * s2);
        if (dres) {
          const float4 r = __ldg(reinterpret_cast<const float4*>(dres + (r0 + j) * C) + q);
          o4.x += r.x; o4.y += r.y; o4.z += r.z; o4.w += r.w;
        }
        reinterpret_cast<float4*>(dx + (r0 + j) * C)[q] = o4;
      }
    }
  }
}

// dgamma / dbeta for C == 64: thread = (channel quad, one of 16 row lanes), 16-byte loads, four rows in flight.
__global__ void __launch_bounds__(256)
ln_bwd_param_c64_kernel(const float4* __restrict__ dy, const float4* __restrict__ x, const float4* __restrict__ gamma,
                        const float4* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ rstd,
                        float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t M, int act) {
  __shared__ float s1[16][65], s2[16][65];
  const int q = threadIdx.x & 15, rsub = threadIdx.x >> 4;
  const int64_t rows_per = (M + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = blockIdx.x * rows_per, r1 = min(M, r0 + rows_per);
  const float4 gm4 = __ldg(gamma + q), bt4 = __ldg(beta + q);
  const float gm[4] = {gm4.x, gm4.y, gm4.z, gm4.w}, bt[4] = {bt4.x, bt4.y, bt4.z, bt4.w};
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int64_t r = r0 + rsub; r < r1; r += 16) {
    const float4 xv = __ldg(x + r * 16 + q), gv = __ldg(dy + r * 16 + q);
    const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
    const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
    float g[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = (xa[k] - mu) * rs;
      if (act) g[k] *= act_grad(xh * gm[k] + bt[k], act);
      a[k] += g[k] * xh;
      b[k] += g[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { s1[rsub][q * 4 + k] = a[k]; s2[rsub][q * 4 + k] = b[k]; }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c = threadIdx.x & 63;
    const float (*src)[65] = threadIdx.x < 64 ? s1 : s2;
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += src[i][c];
    atomicAdd((threadIdx.x < 64 ? dgamma : dbeta) + c, t);
  }
}

// C % 128 == 0 (the transformer widths 128 / 256 / 512): 4 columns per thread as 16-byte loads -- a warp reads 512
// contiguous bytes of a row, 8 loads of 16 bytes in flight per thread (the scalar kernel above measured 19 us for the
// 8192 x 512 rows of transformer 4: one 128-byte line per warp-load, latency-bound).  grid (C/128, row slabs).
__global__ void __launch_bounds__(256)
ln_bwd_param_vec_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                        const float4* __restrict__ gamma, const float4* __restrict__ beta,
                        const float* __restrict__ mean, const float* __restrict__ rstd,
                        float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t M, int C4, int act) {
  __shared__ float4 s1[8][33], s2[8][33];
  const int cq = blockIdx.x * 32 + threadIdx.x;
  const int64_t rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  const float4 gm4 = __ldg(gamma + cq), bt4 = __ldg(beta + cq);
  const float gm[4] = {gm4.x, gm4.y, gm4.z, gm4.w}, bt[4] = {bt4.x, bt4.y, bt4.z, bt4.w};
#pragma unroll 4
  for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
    const float4 xv = __ldg(x + r * C4 + cq), gv = __ldg(dy + r * C4 + cq);
    const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
    const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
    float g[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = (xa[k] - mu) * rs;
      if (act) g[k] *= act_grad(xh * gm[k] + bt[k], act);
      a[k] += g[k] * xh;
      b[k] += g[k];
    }
  }
  s1[threadIdx.y][threadIdx.x] = make_float4(a[0], a[1], a[2], a[3]);
  s2[threadIdx.y][threadIdx.x] = make_float4(b[0], b[1], b[2], b[3]);
  __syncthreads();
  // 256 threads fold the 8 row lanes of the CTA's 128 columns x {dgamma, dbeta}
  const int t = threadIdx.y * 32 + threadIdx.x, col = t & 127, which = t >> 7;
  const float4 (*src)[33] = which ? s2 : s1;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += reinterpret_cast<const float*>(&src[i][col >> 2])[col & 3];
  atomicAdd((which ? dbeta : dgamma) + blockIdx.x * 128 + col, acc);
}

}  // namespace

// Row threshold below which train-mode BatchNorm forward / backward run as ONE launch (one CTA per channel quad, two
// passes over an L2-resident column) instead of reduction + streaming kernels.  Returns the previous value in *old
// (nullable).  Host-side setting, not stream-ordered.
MMFN_API int mmfn_set_bn_small_rows(int rows, int* old) {
  MMFN_CHECK_ARG(rows >= 0, "set_bn_small_rows: rows must be >= 0");
  if (old) *old = BN_SMALL_ROWS;
  BN_SMALL_ROWS = rows;
  return 0;
}

// y_bf16 (nullable): a bf16 twin of y written in the same pass -- the operand of the next bf16 convolution.
// x,y: (M,C) NHWC rows.  ws: per-stream scratch of at least 34*C + 8 doubles that is ZERO on entry (zero it once
// after allocation; every call leaves it zero again).  Writes mean/rstd (C each) and, when running_* are non-null,
// the momentum update with the unbiased variance.
MMFN_API int mmfn_bn_train_fwd(const float* x, float* y, int64_t M, int C,
                               const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float momentum, float eps,
                               float* mean, float* rstd, const float* res, int relu,
                               double* ws, void* y_bf16, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && y && gamma && beta && mean && rstd && ws, "bn_fwd: null pointer");
  MMFN_CHECK_ARG(((uintptr_t)y_bf16 & 7) == 0, "bn_fwd: y_bf16 must be 8-byte aligned");
  MMFN_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0, "bn_fwd: C must be a positive multiple of 4");
  MMFN_CHECK_ARG((((uintptr_t)x | (uintptr_t)y | (uintptr_t)res | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)mean | (uintptr_t)rstd) & 15) == 0,
                 "bn_fwd: 16-byte alignment");
  if (M <= BN_SMALL_ROWS) {
    bn_small_kernel<false><<<C / 4, 256, 0, stream>>>((const float4*)x, nullptr, nullptr, BN_MASK_NONE, (const float4*)res, (float4*)y, nullptr,
        mean, rstd, (const float4*)gamma, (const float4*)beta, running_mean, running_var, nullptr, nullptr, (int)M, C / 4, eps, momentum, relu,
        (uint2*)y_bf16);
    return mmfn_launch_status("bn_train_fwd");
  }
  int qpr; dim3 grid; int64_t rpb;
  bn_colsum_grid(M, C, qpr, grid, rpb);
  BnFinal fz = {mean, rstd, running_mean, running_var, eps, momentum, nullptr, nullptr, nullptr};
  bn_colsum_kernel<false, BN_MASK_NONE><<<grid, BN_THREADS, 0, stream>>>((const float4*)x, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, M, C / 4, qpr, rpb, ws, fz);
  int64_t n4 = M * C / 4;
  bn_apply_kernel<<<grid_1d(n4, 256), 256, 0, stream>>>((const float4*)x, (float4*)y, n4, C / 4,
      mean, rstd, (const float4*)gamma, (const float4*)beta, (const float4*)res, relu, (uint2*)y_bf16);
  return mmfn_launch_status("bn_train_fwd");
}

// dx_bf16 != 0: dx is a BF16 tensor (the gradient of the convolution output only feeds the wgrad / dgrad MMAs).
// Batch statistics only (internal: called by the convolution entry points when their epilogue could not accumulate
// them, i.e. for split-K launches): fills mean / rstd, updates the running statistics, leaves ws zero.
int mmfn_bn_stats_launch(const float* x, int64_t M, int C, float* mean, float* rstd, float* running_mean, float* running_var,
                         float momentum, float eps, double* ws, cudaStream_t stream) {
  int qpr; dim3 grid; int64_t rpb;
  bn_colsum_grid(M, C, qpr, grid, rpb);
  BnFinal fz = {mean, rstd, running_mean, running_var, eps, momentum, nullptr, nullptr, nullptr};
  bn_colsum_kernel<false, BN_MASK_NONE><<<grid, BN_THREADS, 0, stream>>>((const float4*)x, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, M, C / 4, qpr, rpb, ws, fz);
  return mmfn_launch_status("bn_stats");
}

// y = (x - mean) * rstd * gamma + beta (+ res, ReLU) from statistics that already exist (published by the epilogue of
// mmfn_conv2d_fwd_bn_*): the apply half of mmfn_bn_train_fwd.  y_bf16 (nullable): bf16 twin of y; y itself may be null
// when only the twin is wanted (the activation between the two convolutions of a BasicBlock in the bf16 configuration:
// its only readers are the next convolution's TMA loads).
MMFN_API int mmfn_bn_apply(const float* x, float* y, int64_t M, int C, const float* gamma, const float* beta,
                           const float* mean, const float* rstd, const float* res, int relu, void* y_bf16,
                           cudaStream_t stream) {
  MMFN_CHECK_ARG(x && (y || y_bf16) && gamma && beta && mean && rstd, "bn_apply: null pointer");
  MMFN_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0, "bn_apply: C must be a positive multiple of 4");
  MMFN_CHECK_ARG((((uintptr_t)x | (uintptr_t)y | (uintptr_t)res | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)mean | (uintptr_t)rstd) & 15) == 0 &&
                 ((uintptr_t)y_bf16 & 7) == 0, "bn_apply: alignment");
  const int64_t n4 = M * C / 4;
  bn_apply_kernel<<<grid_1d(n4, 256), 256, 0, stream>>>((const float4*)x, (float4*)y, n4, C / 4,
      mean, rstd, (const float4*)gamma, (const float4*)beta, (const float4*)res, relu, (uint2*)y_bf16);
  return mmfn_launch_status("bn_apply");
}

// yout: post-ReLU output of the forward (null when no ReLU followed), fp32 or -- yout_bf16 -- its bf16 twin (only the
// sign is used).  relu_from_z != 0 (with yout null and beta given): a ReLU followed the BatchNorm directly (nothing was
// added in between), so its mask is recomputed from x instead of being read.  dres (nullable) receives the ReLU-masked
// dy for the residual branch.  dgamma/dbeta are accumulated.  ws: as in mmfn_bn_train_fwd.
MMFN_API int mmfn_bn_train_bwd(const float* dy, const float* x, const void* yout, int yout_bf16, int relu_from_z,
                               const float* mean, const float* rstd, const float* gamma, const float* beta,
                               int64_t M, int C, void* dx, int dx_bf16, float* dres, float* dgamma, float* dbeta,
                               double* ws, cudaStream_t stream) {
  MMFN_CHECK_ARG(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && ws, "bn_bwd: null pointer");
  MMFN_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0, "bn_bwd: C must be a positive multiple of 4");
  MMFN_CHECK_ARG(!relu_from_z || (beta && !yout), "bn_bwd: relu_from_z needs beta and no yout");
  MMFN_CHECK_ARG((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)dx | (uintptr_t)dres | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)ws) & 15) == 0 &&
                 ((uintptr_t)yout & (yout_bf16 ? 7 : 15)) == 0, "bn_bwd: 16-byte alignment");
  const int mode = relu_from_z ? BN_MASK_Z : !yout ? BN_MASK_NONE : yout_bf16 ? BN_MASK_Y16 : BN_MASK_Y32;
  if (M <= BN_SMALL_ROWS) {
    bn_small_kernel<true><<<C / 4, 256, 0, stream>>>((const float4*)x, (const float4*)dy, yout, mode, nullptr, (float4*)dx, (float4*)dres,
        const_cast<float*>(mean), const_cast<float*>(rstd), (const float4*)gamma, (const float4*)beta, nullptr, nullptr, dgamma, dbeta, (int)M, C / 4, 0.f, 0.f, 0,
        dx_bf16 ? (uint2*)dx : nullptr);
    return mmfn_launch_status("bn_train_bwd");
  }
  int qpr; dim3 grid; int64_t rpb;
  bn_colsum_grid(M, C, qpr, grid, rpb);
  float* fin = bn_ws_fin(ws, C);
  BnFinal fz = {nullptr, nullptr, nullptr, nullptr, 0.f, 0.f, fin, dgamma, dbeta};
#define MMFN_BN_BWD(MODE)                                                                                                   \
  do {                                                                                                                      \
    bn_colsum_kernel<true, MODE><<<grid, BN_THREADS, 0, stream>>>((const float4*)x, (const float4*)dy, yout, mean, rstd,    \
                                                                  gamma, beta, M, C / 4, qpr, rpb, ws, fz);                 \
    bn_bwd_dx_kernel<MODE><<<grid_1d(M * C / 4, 256), 256, 0, stream>>>(                                                    \
        (const float4*)dy, (const float4*)x, yout, mean, rstd, gamma, beta, fin, M, C / 4, (float4*)dx, (float4*)dres,      \
        dx_bf16 ? (uint2*)dx : nullptr);                                                                                    \
  } while (0)
  if (mode == BN_MASK_Z) MMFN_BN_BWD(BN_MASK_Z);
  else if (mode == BN_MASK_Y16) MMFN_BN_BWD(BN_MASK_Y16);
  else if (mode == BN_MASK_Y32) MMFN_BN_BWD(BN_MASK_Y32);
  else MMFN_BN_BWD(BN_MASK_NONE);
#undef MMFN_BN_BWD
  return mmfn_launch_status("bn_train_bwd");
}

static inline int stem_grid(int64_t n) { const int64_t b = ceil_div64(n, 256); return (int)(b < 1 ? 1 : (b > 148 * 6 ? 148 * 6 : b)); }

// Stem tail forward: train-mode bn1 -> relu -> maxpool(3, 2, 1) of the two ResNet stems in two launches (batch
// statistics, then ONE pass that normalises, rectifies and pools; model_rad.py:513-515 image, :519-521 LiDAR).
// z (B, H, W, C) is the stem convolution's output; out (B, Ho, Wo, C) with Ho = (H - 1) / 2 + 1; out_bf16 (nullable): its
// bf16 twin; idx (B, Ho, Wo, C) bytes: arg-max tap 0..8 of each window, bit 7 set when the maximum is not positive
// (ReLU mask for mmfn_stem_bn_relu_maxpool_bwd -- NOT the plain tap mmfn_maxpool3x3s2_bwd expects); zmax (B, Ho, Wo, C):
// z at the arg-max, saved for the backward reduction.  mean / rstd (C) are written; running statistics updated with
// `momentum`.  ws: as in mmfn_bn_train_fwd.
MMFN_API int mmfn_stem_bn_relu_maxpool_fwd(const float* z, int B, int H, int W, int C, const float* gamma, const float* beta,
                                           float* running_mean, float* running_var, float momentum, float eps,
                                           float* mean, float* rstd, float* out, void* out_bf16, uint8_t* idx, float* zmax,
                                           double* ws, cudaStream_t stream) {
  MMFN_CHECK_ARG(z && gamma && beta && mean && rstd && out && idx && zmax && ws, "stem_fwd: null pointer");
  MMFN_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "stem_fwd: bad shape (C % 4 == 0)");
  MMFN_CHECK_ARG((((uintptr_t)z | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)out | (uintptr_t)zmax) & 15) == 0 &&
                 ((uintptr_t)out_bf16 & 7) == 0 && ((uintptr_t)idx & 3) == 0, "stem_fwd: alignment");
  const int64_t M = (int64_t)B * H * W;
  MMFN_CHECK_ARG(M < (int64_t)1 << 31, "stem_fwd: too many pixels");
  int rc = mmfn_bn_stats_launch(z, M, C, mean, rstd, running_mean, running_var, momentum, eps, ws, stream);
  if (rc) return rc;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  // 72 / 80 registers: three CTAs per SM -> grids of two full waves (148 x 3 x 2), not grid_1d's eight CTAs per SM
  bn_relu_maxpool_fwd_kernel<<<stem_grid((int64_t)B * Ho * Wo * (C / 4)), 256, 0, stream>>>(
      z, mean, rstd, (const float4*)gamma, (const float4*)beta, out, (uint2*)out_bf16, idx, (float4*)zmax, B, H, W, C, Ho, Wo);
  return mmfn_launch_status("stem_bn_relu_maxpool_fwd");
}

// Stem tail backward: dout (B, Ho, Wo, C) -> dz (B, H, W, C), the gradient of the stem convolution's output (bf16 when
// dz_bf16: it only feeds the weight-gradient MMA), through maxpool, ReLU and train-mode BatchNorm; dgamma / dbeta are
// accumulated.  idx / zmax: as written by the forward.  Two launches: a reduction over the pooled pixels (the only
// non-zero gradients) and one pass over z.
MMFN_API int mmfn_stem_bn_relu_maxpool_bwd(const float* dout, const uint8_t* idx, const float* z, const float* zmax, const float* mean,
                                           const float* rstd, const float* gamma, int B, int H, int W, int C, void* dz,
                                           int dz_bf16, float* dgamma, float* dbeta, double* ws, cudaStream_t stream) {
  MMFN_CHECK_ARG(dout && idx && z && zmax && mean && rstd && gamma && dz && dgamma && dbeta && ws, "stem_bwd: null pointer");
  MMFN_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "stem_bwd: bad shape (C % 4 == 0)");
  MMFN_CHECK_ARG((((uintptr_t)dout | (uintptr_t)z | (uintptr_t)zmax | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)gamma | (uintptr_t)dz | (uintptr_t)ws) & 15) == 0 &&
                 ((uintptr_t)idx & 3) == 0, "stem_bwd: alignment");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int64_t M = (int64_t)B * H * W, Mp = (int64_t)B * Ho * Wo;
  MMFN_CHECK_ARG(M < (int64_t)1 << 31, "stem_bwd: too many pixels");
  int qpr; dim3 grid; int64_t rpb;
  bn_colsum_grid(Mp, C, qpr, grid, rpb);
  float* fin = bn_ws_fin(ws, C);
  BnFinal fz = {nullptr, nullptr, nullptr, nullptr, 0.f, 0.f, fin, dgamma, dbeta};
  stem_bwd_stats_kernel<<<grid, BN_THREADS, 0, stream>>>((const float4*)dout, idx, (const float4*)zmax, mean, rstd, C / 4, Mp, M, qpr, rpb, ws, fz);
  const int Hp = (H + 1) / 2, Wp = (W + 1) / 2;
  stem_bwd_dx_kernel<<<stem_grid((int64_t)B * Hp * Wp * (C / 4)), 256, 0, stream>>>(
      dout, idx, z, mean, rstd, gamma, fin, B, H, W, C, Ho, Wo, dz_bf16 ? nullptr : (float*)dz, dz_bf16 ? (uint2*)dz : nullptr);
  return mmfn_launch_status("stem_bn_relu_maxpool_bwd");
}

namespace {
__global__ void bn_eval_stats_kernel(const float* rm, const float* rv, float eps, float* mean, float* rstd, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { mean[c] = rm[c]; rstd[c] = 1.0f / sqrtf(rv[c] + eps); }
}
}  // namespace

// Inference-mode BN: y = (x - running_mean) / sqrt(running_var + eps) * gamma + beta (+res, relu).
MMFN_API int mmfn_bn_eval_fwd(const float* x, float* y, int64_t M, int C, const float* gamma, const float* beta,
                              const float* running_mean, const float* running_var, float eps,
                              float* mean, float* rstd, const float* res, int relu, void* y_bf16, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && y && gamma && beta && running_mean && running_var && mean && rstd, "bn_eval: null pointer");
  MMFN_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0, "bn_eval: C must be a positive multiple of 4");
  bn_eval_stats_kernel<<<(C + 127) / 128, 128, 0, stream>>>(running_mean, running_var, eps, mean, rstd, C);
  int64_t n4 = M * C / 4;
  bn_apply_kernel<<<grid_1d(n4, 256), 256, 0, stream>>>((const float4*)x, (float4*)y, n4, C / 4,
      mean, rstd, (const float4*)gamma, (const float4*)beta, (const float4*)res, relu, (uint2*)y_bf16);
  return mmfn_launch_status("bn_eval_fwd");
}

// act: 0 none, 1 ReLU, 2 exact GELU applied to the LayerNorm output.
// y_bf16 != 0: y is a BF16 tensor (LayerNorm outputs that only feed bf16 GEMMs).
MMFN_API int mmfn_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_bf16,
                                float* mean, float* rstd, int64_t M, int C, float eps, int act,
                                cudaStream_t stream) {
  MMFN_CHECK_ARG(x && gamma && beta && y && mean && rstd, "ln_fwd: null pointer");
  MMFN_CHECK_ARG(M >= 0 && C > 0, "ln_fwd: bad sizes");
  if (M == 0) return 0;
  ln_fwd_kernel<<<(unsigned)ceil_div64(M, 8), 256, 0, stream>>>(x, gamma, beta, y_bf16 ? nullptr : (float*)y, mean, rstd, M, C, eps, act,
                                                               y_bf16 ? (__nv_bfloat16*)y : nullptr);
  return mmfn_launch_status("layernorm_fwd");
}

// drop_bf16 != 0: dx_drop is a BF16 tensor and is written even when drop_p == 0 (it feeds the bf16 GEMMs of the branch).
// parts: bit 0 = data gradient dx (+ dres; optionally also dx_drop = dx * dropout_mask(drop_p, drop_seed), the
// gradient entering the dropout of the next residual branch), bit 1 = parameter gradients (accumulated).  The two
// halves are independent kernels so that a caller can put the parameter reduction on a side stream.
MMFN_API int mmfn_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* beta,
                                const float* mean, const float* rstd, const float* dres, float* dx,
                                float* dgamma, float* dbeta, int64_t M, int C, int act, int parts,
                                void* dx_drop, int drop_bf16, float drop_p, uint64_t drop_seed, cudaStream_t stream) {
  MMFN_CHECK_ARG(dy && x && gamma && beta && mean && rstd, "ln_bwd: null pointer");
  MMFN_CHECK_ARG((parts & 3) != 0 && (!(parts & 1) || dx) && (!(parts & 2) || (dgamma && dbeta)), "ln_bwd: missing output for the requested parts");
  MMFN_CHECK_ARG(M >= 0 && C > 0 && C <= 512 && C % 4 == 0, "ln_bwd: C must be a multiple of 4 in (0, 512]");
  MMFN_CHECK_ARG((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)dx | (uintptr_t)dres | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)dx_drop) & 15) == 0,
                 "ln_bwd: 16-byte alignment");
  if (M == 0) return 0;
  if (parts & 1) {
    unsigned blocks = (unsigned)ceil_div64(M, 8);
    // fp32 copy: only needed when a mask applies (the caller reuses dx otherwise); bf16 copy: always (it is a conversion too)
    float* dd = (!drop_bf16 && drop_p > 0.f) ? (float*)dx_drop : nullptr;
    __nv_bfloat16* dd16 = drop_bf16 ? (__nv_bfloat16*)dx_drop : nullptr;
    if (C == 64 && !dd && !dd16 && M >= 4096) {
      constexpr int RPT = 4;
      ln_bwd_dx_c64_kernel<RPT><<<grid_1d(ceil_div64(M, RPT) * 16, 256), 256, 0, stream>>>(dy, x, gamma, beta, mean, rstd, dres, dx, M, act);
    } else if (C <= 128) ln_bwd_dx_kernel<1><<<blocks, 256, 0, stream>>>(dy, x, gamma, beta, mean, rstd, dres, dx, M, C, act, dd, drop_p, drop_seed, dd16);
    else if (C <= 256) ln_bwd_dx_kernel<2><<<blocks, 256, 0, stream>>>(dy, x, gamma, beta, mean, rstd, dres, dx, M, C, act, dd, drop_p, drop_seed, dd16);
    else ln_bwd_dx_kernel<4><<<blocks, 256, 0, stream>>>(dy, x, gamma, beta, mean, rstd, dres, dx, M, C, act, dd, drop_p, drop_seed, dd16);
  }
  if (parts & 2) {
    // 32 rows (4 row iterations per thread, all loads in flight) per CTA unless that exceeds ~8 waves of CTAs
    int64_t slabs = ceil_div64(M, 32);
    if (C == 64 && M >= 4096) {
      int64_t ctas = ceil_div64(M, 64);
      if (ctas > 148 * 4) ctas = 148 * 4;
      ln_bwd_param_c64_kernel<<<(unsigned)ctas, 256, 0, stream>>>((const float4*)dy, (const float4*)x, (const float4*)gamma, (const float4*)beta,
                                                                  mean, rstd, dgamma, dbeta, M, act);
      return mmfn_launch_status("layernorm_bwd");
    }
    if (C % 128 == 0 && M >= 1024) {
      const int64_t capv = ceil_div64(148 * 4, C / 128);
      if (slabs > capv) slabs = capv;
      ln_bwd_param_vec_kernel<<<dim3(C / 128, (unsigned)slabs), dim3(32, 8), 0, stream>>>(
          (const float4*)dy, (const float4*)x, (const float4*)gamma, (const float4*)beta, mean, rstd, dgamma, dbeta, M, C / 4, act);
      return mmfn_launch_status("layernorm_bwd");
    }
    const int64_t cap = ceil_div64(148 * 8, (C + 31) / 32);
    if (slabs > cap) slabs = cap;
    ln_bwd_param_kernel<<<dim3((C + 31) / 32, (unsigned)slabs), dim3(32, 8), 0, stream>>>(dy, x, gamma, beta, mean, rstd, dgamma, dbeta, M, C, act);
  }
  return mmfn_launch_status("layernorm_bwd");
}

MMFN_DEFINE_RNG_BINDER(norm)
