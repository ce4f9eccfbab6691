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
constexpr int BN_THREADS = 256;

template <bool BWD>
__global__ void __launch_bounds__(BN_THREADS)
bn_colsum_kernel(const float4* __restrict__ x, const float4* __restrict__ dy, const float4* __restrict__ yout,
                 const float* __restrict__ mean, const float* __restrict__ rstd, int64_t M, int C4, int qpr,
                 int64_t rows_per_block, double* __restrict__ ws) {
  __shared__ double red[BN_THREADS][8];
  const int q = threadIdx.x % qpr, rsub = threadIdx.x / qpr, nrs = BN_THREADS / qpr;
  const int cq = blockIdx.x * qpr + q;                          // channel quad of this thread
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cq < C4) {
    float m4[4] = {0.f, 0.f, 0.f, 0.f}, r4[4] = {1.f, 1.f, 1.f, 1.f};
    if (BWD) {
      const float4 mm = __ldg(reinterpret_cast<const float4*>(mean) + cq), rr = __ldg(reinterpret_cast<const float4*>(rstd) + cq);
      m4[0] = mm.x; m4[1] = mm.y; m4[2] = mm.z; m4[3] = mm.w;
      r4[0] = rr.x; r4[1] = rr.y; r4[2] = rr.z; r4[3] = rr.w;
    }
#pragma unroll 2
    for (int64_t r = r0 + rsub; r < r1; r += nrs) {
      const float4 xv = __ldg(x + r * C4 + cq);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
      if (!BWD) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[k] += xa[k]; acc[4 + k] += (double)xa[k] * xa[k]; }
      } else {
        const float4 gv = __ldg(dy + r * C4 + cq);
        float ga[4] = {gv.x, gv.y, gv.z, gv.w};
        if (yout) {
          const float4 yv = __ldg(yout + r * C4 + cq);
          if (!(yv.x > 0.f)) ga[0] = 0.f;
          if (!(yv.y > 0.f)) ga[1] = 0.f;
          if (!(yv.z > 0.f)) ga[2] = 0.f;
          if (!(yv.w > 0.f)) ga[3] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[k] += ga[k]; acc[4 + k] += (double)ga[k] * ((xa[k] - m4[k]) * r4[k]); }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x][k] = acc[k];
  __syncthreads();
  // thread t < qpr * 8 reduces one (quad, component) over the row lanes
  if (threadIdx.x < qpr * 8) {
    const int qq = threadIdx.x >> 3, k = threadIdx.x & 7;
    const int cqq = blockIdx.x * qpr + qq;
    if (cqq < C4) {
      double t = 0.0;
      for (int rs = 0; rs < nrs; ++rs) t += red[rs * qpr + qq][k];
      atomicAdd(ws + (k < 4 ? 0 : (int64_t)C4 * 4) + cqq * 4 + (k & 3), t);
    }
  }
}

static inline void bn_colsum_grid(int64_t M, int C, int& qpr, dim3& grid, int64_t& rows_per_block) {
  const int C4 = C / 4;
  qpr = C4 < 32 ? C4 : 32;
  while (BN_THREADS % qpr) --qpr;                                // C4 = 16, 32, 64, 128 here; stay safe for odd widths
  const int groups = (C4 + qpr - 1) / qpr;
  const int nrs = BN_THREADS / qpr;
  int64_t slabs = ceil_div64(M, (int64_t)nrs * 4);               // at least 4 rows per thread
  const int64_t cap = (148 * 4 + groups - 1) / groups;
  if (slabs > cap) slabs = cap;
  if (slabs < 1) slabs = 1;
  rows_per_block = ceil_div64(M, slabs);
  grid = dim3((unsigned)groups, (unsigned)ceil_div64(M, rows_per_block));
}

// y = (x - mean) * rstd * gamma + beta (+res, ReLU).  With ws != null the batch statistics are finalised here
// from the fp64 partial sums (every CTA recomputes the C means into shared memory; CTA 0 also publishes
// mean/rstd for backward and applies the running-statistics momentum update) -- no separate finalize launch.
__global__ void bn_apply_kernel(const float4* __restrict__ x, float4* __restrict__ y, int64_t n4, int C4,
                                float* __restrict__ mean, float* __restrict__ rstd,
                                const float4* __restrict__ gamma, const float4* __restrict__ beta,
                                const float4* __restrict__ res, int relu,
                                const double* __restrict__ ws, int64_t M, float eps, float momentum,
                                float* __restrict__ running_mean, float* __restrict__ running_var) {
  extern __shared__ float sm[];            // mean[C] | rstd[C]
  const int C = C4 * 4;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float m, r;
    if (ws) {
      double md = ws[c] / (double)M;
      double var = ws[C + c] / (double)M - md * md;
      if (var < 0.0) var = 0.0;
      m = (float)md;
      r = (float)(1.0 / sqrt(var + (double)eps));
      if (blockIdx.x == 0) {
        mean[c] = m;
        rstd[c] = r;
        if (running_mean) {
          double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
          running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * md);
          running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
        }
      }
    } else {
      m = mean[c];
      r = rstd[c];
    }
    sm[c] = m;
    sm[C + c] = r;
  }
  __syncthreads();
  const float4* sm_mean = reinterpret_cast<const float4*>(sm);
  const float4* sm_rstd = reinterpret_cast<const float4*>(sm + C);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C4);
    float4 v = __ldg(x + i), m = sm_mean[c], r = sm_rstd[c], g = __ldg(gamma + c), b = __ldg(beta + c);
    float4 o;
    o.x = (v.x - m.x) * r.x * g.x + b.x;
    o.y = (v.y - m.y) * r.y * g.y + b.y;
    o.z = (v.z - m.z) * r.z * g.z + b.z;
    o.w = (v.w - m.w) * r.w * g.w + b.w;
    if (res) { float4 q = __ldg(res + i); o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    y[i] = o;
  }
}

// dx = gamma * rstd * (dy' - mean(dy') - xhat * mean(dy' * xhat)); dres = dy' (residual branch).  float4 over channels;
// the per-channel coefficients are staged in shared memory once per CTA.  CTA 0 accumulates dgamma / dbeta.
__global__ void bn_bwd_dx_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                                 const float4* __restrict__ yout, const float* __restrict__ mean,
                                 const float* __restrict__ rstd, const float* __restrict__ gamma,
                                 const double* __restrict__ ws, int64_t M, int C4,
                                 float4* __restrict__ dx, float4* __restrict__ dres,
                                 float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float sm[];            // mean[C] | rstd[C] | gamma*rstd[C] | mean(dy')[C] | mean(dy' xhat)[C]
  const int C = C4 * 4;
  const double invM = 1.0 / (double)M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float rs = rstd[c];
    sm[c] = mean[c];
    sm[C + c] = rs;
    sm[2 * C + c] = gamma[c] * rs;
    sm[3 * C + c] = (float)(ws[c] * invM);
    sm[4 * C + c] = (float)(ws[C + c] * invM);
    if (blockIdx.x == 0) {
      dbeta[c] += (float)ws[c];
      dgamma[c] += (float)ws[C + c];
    }
  }
  __syncthreads();
  const float4* s_mean = reinterpret_cast<const float4*>(sm);
  const float4* s_rstd = reinterpret_cast<const float4*>(sm + C);
  const float4* s_k = reinterpret_cast<const float4*>(sm + 2 * C);
  const float4* s_dy = reinterpret_cast<const float4*>(sm + 3 * C);
  const float4* s_dx = reinterpret_cast<const float4*>(sm + 4 * C);
  const int64_t n4 = M * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    float4 g = __ldg(dy + i);
    if (yout) {
      const float4 yv = __ldg(yout + i);
      if (!(yv.x > 0.f)) g.x = 0.f;
      if (!(yv.y > 0.f)) g.y = 0.f;
      if (!(yv.z > 0.f)) g.z = 0.f;
      if (!(yv.w > 0.f)) g.w = 0.f;
    }
    const float4 xv = __ldg(x + i), m = s_mean[c], r = s_rstd[c], k = s_k[c], a = s_dy[c], b = s_dx[c];
    float4 o;
    o.x = k.x * (g.x - a.x - (xv.x - m.x) * r.x * b.x);
    o.y = k.y * (g.y - a.y - (xv.y - m.y) * r.y * b.y);
    o.z = k.z * (g.z - a.z - (xv.z - m.z) * r.z * b.z);
    o.w = k.w * (g.w - a.w - (xv.w - m.w) * r.w * b.w);
    dx[i] = o;
    if (dres) dres[i] = g;
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
                              int64_t M, int C, float eps, int act) {
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
  float* yr = y + row * C;
  for (int c = lane; c < C; c += 32) yr[c] = act_fwd((xr[c] - mu) * rs * gamma[c] + beta[c], act);
}

// dx = LN'(dy * act'(ln)) (+ dres).  One warp per row, float4 columns (C % 4 == 0, C <= 128 * NV).
template <int NV>
__global__ void ln_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                 const float* __restrict__ dres, float* __restrict__ dx, int64_t M, int C, int act,
                                 float* __restrict__ dx_drop, float drop_p, uint64_t drop_seed) {
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
      if (dx_drop) {        // gradient entering the next residual branch's dropout (same hash as the forward mask)
        const uint64_t i0 = (uint64_t)(row * C + c);
        o.x *= mmfn_dropout_scale(drop_p, drop_seed, i0);
        o.y *= mmfn_dropout_scale(drop_p, drop_seed, i0 + 1);
        o.z *= mmfn_dropout_scale(drop_p, drop_seed, i0 + 2);
        o.w *= mmfn_dropout_scale(drop_p, drop_seed, i0 + 3);
        *reinterpret_cast<float4*>(dx_drop + row * C + c) = o;
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
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const float xh = (__ldg(x + r * C + c) - mean[r]) * rstd[r];
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

}  // namespace

// x,y: (M,C) NHWC rows. ws: 2*C doubles of scratch. Writes mean/rstd (C each) and, when
// running_* are non-null, the momentum update with the unbiased variance.
MMFN_API int mmfn_bn_train_fwd(const float* x, float* y, int64_t M, int C,
                               const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float momentum, float eps,
                               float* mean, float* rstd, const float* res, int relu,
                               double* ws, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && y && gamma && beta && mean && rstd && ws, "bn_fwd: null pointer");
  MMFN_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0, "bn_fwd: C must be a positive multiple of 4");
  MMFN_CHECK_ARG((((uintptr_t)x | (uintptr_t)y | (uintptr_t)res | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0, "bn_fwd: 16-byte alignment");
  cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, stream);
  int qpr; dim3 grid; int64_t rpb;
  bn_colsum_grid(M, C, qpr, grid, rpb);
  bn_colsum_kernel<false><<<grid, BN_THREADS, 0, stream>>>((const float4*)x, nullptr, nullptr, nullptr, nullptr, M, C / 4, qpr, rpb, ws);
  int64_t n4 = M * C / 4;
  bn_apply_kernel<<<grid_1d(n4, 256), 256, 2 * C * sizeof(float), stream>>>((const float4*)x, (float4*)y, n4, C / 4,
      mean, rstd, (const float4*)gamma, (const float4*)beta, (const float4*)res, relu,
      ws, M, eps, momentum, running_mean, running_var);
  return mmfn_launch_status("bn_train_fwd");
}

// yout: post-ReLU output of the forward (null when no ReLU followed). dres (nullable)
// receives the ReLU-masked dy for the residual branch.  dgamma/dbeta are accumulated.
MMFN_API int mmfn_bn_train_bwd(const float* dy, const float* x, const float* yout,
                               const float* mean, const float* rstd, const float* gamma,
                               int64_t M, int C, float* dx, float* dres, float* dgamma, float* dbeta,
                               double* ws, cudaStream_t stream) {
  MMFN_CHECK_ARG(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && ws, "bn_bwd: null pointer");
  MMFN_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0, "bn_bwd: C must be a positive multiple of 4");
  MMFN_CHECK_ARG((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)yout | (uintptr_t)dx | (uintptr_t)dres | (uintptr_t)mean | (uintptr_t)rstd) & 15) == 0,
                 "bn_bwd: 16-byte alignment");
  cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, stream);
  int qpr; dim3 grid; int64_t rpb;
  bn_colsum_grid(M, C, qpr, grid, rpb);
  bn_colsum_kernel<true><<<grid, BN_THREADS, 0, stream>>>((const float4*)x, (const float4*)dy, (const float4*)yout, mean, rstd, M, C / 4, qpr, rpb, ws);
  bn_bwd_dx_kernel<<<grid_1d(M * C / 4, 256), 256, 5 * C * sizeof(float), stream>>>(
      (const float4*)dy, (const float4*)x, (const float4*)yout, mean, rstd, gamma, ws, M, C / 4, (float4*)dx, (float4*)dres, dgamma, dbeta);
  return mmfn_launch_status("bn_train_bwd");
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
                              float* mean, float* rstd, const float* res, int relu, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && y && gamma && beta && running_mean && running_var && mean && rstd, "bn_eval: null pointer");
  MMFN_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0, "bn_eval: C must be a positive multiple of 4");
  bn_eval_stats_kernel<<<(C + 127) / 128, 128, 0, stream>>>(running_mean, running_var, eps, mean, rstd, C);
  int64_t n4 = M * C / 4;
  bn_apply_kernel<<<grid_1d(n4, 256), 256, 2 * C * sizeof(float), stream>>>((const float4*)x, (float4*)y, n4, C / 4,
      mean, rstd, (const float4*)gamma, (const float4*)beta, (const float4*)res, relu,
      nullptr, M, eps, 0.f, nullptr, nullptr);
  return mmfn_launch_status("bn_eval_fwd");
}

// act: 0 none, 1 ReLU, 2 exact GELU applied to the LayerNorm output.
MMFN_API int mmfn_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y,
                                float* mean, float* rstd, int64_t M, int C, float eps, int act,
                                cudaStream_t stream) {
  MMFN_CHECK_ARG(x && gamma && beta && y && mean && rstd, "ln_fwd: null pointer");
  MMFN_CHECK_ARG(M >= 0 && C > 0, "ln_fwd: bad sizes");
  if (M == 0) return 0;
  ln_fwd_kernel<<<(unsigned)ceil_div64(M, 8), 256, 0, stream>>>(x, gamma, beta, y, mean, rstd, M, C, eps, act);
  return mmfn_launch_status("layernorm_fwd");
}

// parts: bit 0 = data gradient dx (+ dres; optionally also dx_drop = dx * dropout_mask(drop_p, drop_seed), the
// gradient entering the dropout of the next residual branch), bit 1 = parameter gradients (accumulated).  The two
// halves are independent kernels so that a caller can put the parameter reduction on a side stream.
MMFN_API int mmfn_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* beta,
                                const float* mean, const float* rstd, const float* dres, float* dx,
                                float* dgamma, float* dbeta, int64_t M, int C, int act, int parts,
                                float* dx_drop, float drop_p, uint64_t drop_seed, cudaStream_t stream) {
  MMFN_CHECK_ARG(dy && x && gamma && beta && mean && rstd, "ln_bwd: null pointer");
  MMFN_CHECK_ARG((parts & 3) != 0 && (!(parts & 1) || dx) && (!(parts & 2) || (dgamma && dbeta)), "ln_bwd: missing output for the requested parts");
  MMFN_CHECK_ARG(M >= 0 && C > 0 && C <= 512 && C % 4 == 0, "ln_bwd: C must be a multiple of 4 in (0, 512]");
  MMFN_CHECK_ARG((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)dx | (uintptr_t)dres | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)dx_drop) & 15) == 0,
                 "ln_bwd: 16-byte alignment");
  if (M == 0) return 0;
  if (parts & 1) {
    unsigned blocks = (unsigned)ceil_div64(M, 8);
    if (drop_p <= 0.f) dx_drop = nullptr;
    if (C <= 128) ln_bwd_dx_kernel<1><<<blocks, 256, 0, stream>>>(dy, x, gamma, beta, mean, rstd, dres, dx, M, C, act, dx_drop, drop_p, drop_seed);
    else if (C <= 256) ln_bwd_dx_kernel<2><<<blocks, 256, 0, stream>>>(dy, x, gamma, beta, mean, rstd, dres, dx, M, C, act, dx_drop, drop_p, drop_seed);
    else ln_bwd_dx_kernel<4><<<blocks, 256, 0, stream>>>(dy, x, gamma, beta, mean, rstd, dres, dx, M, C, act, dx_drop, drop_p, drop_seed);
  }
  if (parts & 2) {
    int64_t slabs = ceil_div64(M, 128);
    if (slabs > 256) slabs = 256;
    ln_bwd_param_kernel<<<dim3((C + 31) / 32, (unsigned)slabs), dim3(32, 8), 0, stream>>>(dy, x, gamma, beta, mean, rstd, dgamma, dbeta, M, C, act);
  }
  return mmfn_launch_status("layernorm_bwd");
}
