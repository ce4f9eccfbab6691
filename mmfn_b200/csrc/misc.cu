// Small HBM-bound helpers: bias-gradient column sums, ELU, dropout, VectorNet polyline
// vectorisation + segment max-pool (model_rad.py:270-283, :369-382), the radar
// log-softmax with the reference's (8,8,512)->transpose(1,3) relayout (:883-884).
#include "common.cuh"

namespace {

// out[n] (+)= sum_m x[m*ld + n];  blockDim (32, 8), grid.x over column groups, grid.y over row slabs
__global__ void colsum_kernel(const float* __restrict__ x, int64_t ld, int64_t M, int N, float* __restrict__ out) {
  __shared__ float s[8][33];
  int n = blockIdx.x * 32 + threadIdx.x;
  int64_t rows_per = (M + gridDim.y - 1) / gridDim.y;
  int64_t r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float a = 0.f;
  if (n < N) {
#pragma unroll 8
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) a += __ldg(x + r * ld + n);
  }
  s[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
#pragma unroll
    for (int i = 1; i < 8; ++i) a += s[i][threadIdx.x];
    atomicAdd(out + n, a);
  }
}

// bf16 input variant (gradients written in bf16 by the LayerNorm backward / GEMM epilogues of the bf16 configuration)
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, int64_t M, int N, float* __restrict__ out) {
  __shared__ float s[8][33];
  int n = blockIdx.x * 32 + threadIdx.x;
  int64_t rows_per = (M + gridDim.y - 1) / gridDim.y;
  int64_t r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float a = 0.f;
  if (n < N) {
#pragma unroll 8
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) a += __bfloat162float(x[r * ld + n]);
  }
  s[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
#pragma unroll
    for (int i = 1; i < 8; ++i) a += s[i][threadIdx.x];
    atomicAdd(out + n, a);
  }
}

// 16-byte loads: EPT = 4 fp32 / 8 bf16 columns per thread, a warp reads 512 contiguous bytes of a row (the scalar kernels
// fetch 128 / 64 bytes per warp-load: latency-bound on the 8192 x 2048 gradients of transformer 4).  N % EPT == 0,
// ld % EPT == 0, 16-byte aligned base.  blockDim (32, 8); grid (column groups of 32 EPT, row slabs).
template <typename T, int EPT>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* __restrict__ x, int64_t ld, int64_t M, int N, float* __restrict__ out) {
  __shared__ float s[8][32 * EPT + 1];
  const int n0 = (blockIdx.x * 32 + threadIdx.x) * EPT;
  const int64_t rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float a[EPT];
#pragma unroll
  for (int k = 0; k < EPT; ++k) a[k] = 0.f;
  if (n0 < N) {
#pragma unroll 4
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(x + r * ld + n0));
      if (EPT == 4) {
        a[0] += __uint_as_float(w.x); a[1] += __uint_as_float(w.y); a[2] += __uint_as_float(w.z); a[3] += __uint_as_float(w.w);
      } else {
        const float4 lo = mmfn_unpack_bf16x4(make_uint2(w.x, w.y)), hi = mmfn_unpack_bf16x4(make_uint2(w.z, w.w));
        a[0] += lo.x; a[1] += lo.y; a[2] += lo.z; a[3] += lo.w;
        a[4 % EPT] += hi.x; a[5 % EPT] += hi.y; a[6 % EPT] += hi.z; a[7 % EPT] += hi.w;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < EPT; ++k) s[threadIdx.y][threadIdx.x * EPT + k] = a[k];
  __syncthreads();
  for (int c = threadIdx.y * 32 + threadIdx.x; c < 32 * EPT; c += 256) {
    const int n = blockIdx.x * 32 * EPT + c;
    if (n < N) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += s[i][c];
      atomicAdd(out + n, t);
    }
  }
}

template <typename T, int EPT>
static bool colsum_vec_launch(const T* x, int64_t ld, int64_t M, int N, float* out, cudaStream_t stream) {
  if (N % EPT != 0 || ld % EPT != 0 || ((uintptr_t)x & 15) != 0 || M < 512) return false;
  const int64_t colgroups = (N + 32 * EPT - 1) / (32 * EPT);
  int64_t slabs = ceil_div64(M, 32);
  const int64_t cap = ceil_div64(148 * 4, colgroups);
  if (slabs > cap) slabs = cap;
  colsum_vec_kernel<T, EPT><<<dim3((unsigned)colgroups, (unsigned)slabs), dim3(32, 8), 0, stream>>>(x, ld, M, N, out);
  return true;
}

// y = bf16(x), 4 elements per thread
__global__ void f32_to_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ y, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    y[i] = mmfn_pack_bf16x4(v.x, v.y, v.z, v.w);
  }
}

__global__ void elu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i];
    y[i] = v > 0.f ? v : expm1f(v);
  }
}
// dx = dy * (y > 0 ? 1 : y + 1)  (alpha = 1; uses the forward OUTPUT y)
__global__ void elu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = y[i];
    dx[i] = dy[i] * (v > 0.f ? 1.f : v + 1.f);
  }
}

// dx = dy * (y > 0)
__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float p, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = x[i] * mmfn_dropout_scale(p, seed, (uint64_t)i);
}

__global__ void axpy_kernel(const float* __restrict__ x, float* __restrict__ y, float a, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] += a * x[i];
}

// im2col for the two network stems (7x7 stride-2 convolutions on 3 / 2 input channels, model_rad.py:22, :58-62):
// col[pixel, k] with k = (r*S + s)*C + c, rows padded with zeros to Kp (a multiple of 32 = one TMA k-block), so the
// stem runs as ONE dense tensor-core GEMM (forward) and one split-K GEMM (weight gradient) over this matrix instead of
// a SIMT gather-GEMM.  One thread writes 4 consecutive k (16-byte store); the gathers hit L1/L2 (the input is small).
__global__ void im2col_kernel(const float* __restrict__ x, float* __restrict__ col, int N, int H, int W, int C, int R, int S,
                              int stride, int pad, int Ho, int Wo, int K, int Kp) {
  const int kq = Kp >> 2;
  const int64_t n4 = (int64_t)N * Ho * Wo * kq;
  const int SC = S * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / kq;
    const int k0 = (int)(i - pix * kq) * 4;
    const int wo = (int)(pix % Wo);
    const int64_t t = pix / Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const int h0 = ho * stride - pad, w0 = wo * stride - pad;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + j;
      float val = 0.f;
      if (k < K) {
        const int r = k / SC, rem = k - r * SC;
        const int sx = rem / C, c = rem - sx * C;
        const int h = h0 + r, w = w0 + sx;
        if (h >= 0 && h < H && w >= 0 && w < W) val = __ldg(x + (((int64_t)n * H + h) * W + w) * C + c);
      }
      v[j] = val;
    }
    *reinterpret_cast<float4*>(col + i * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// Fast path for the stems (compile-time S, C): one warp per output pixel, lane l writes columns l, l+32, ... of the
// pixel's row (fully coalesced 128-byte stores); a filter row is a contiguous run of S*C input floats, so the loads of
// neighbouring lanes are contiguous too, and the only divisions are by compile-time constants.
__device__ __forceinline__ void im2col_store(float* p, float v) { *p = v; }
__device__ __forceinline__ void im2col_store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <int S, int C, int KP, typename TOUT = float>
__global__ void __launch_bounds__(256)
im2col_rows_kernel(const float* __restrict__ x, TOUT* __restrict__ col, int N, int H, int W, int R,
                   int stride, int pad, int Ho, int Wo) {
  constexpr int SC = S * C;
  constexpr int NK = KP / 32;                // columns per lane
  const int K = R * SC;
  const int lane = threadIdx.x & 31;
  const int npix = N * Ho * Wo;
  const int wpg = (gridDim.x * blockDim.x) >> 5;
  // per-lane column decomposition is the same for every pixel: hoisted out of the pixel loop
  int rr[NK], rem[NK], dw[NK];
  bool kin[NK];
#pragma unroll
  for (int j = 0; j < NK; ++j) {
    const int k = lane + 32 * j;
    kin[j] = k < K;
    rr[j] = k / SC;
    rem[j] = k - rr[j] * SC;
    dw[j] = rem[j] / C;
  }
  for (int pix = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pix < npix; pix += wpg) {
    const int wo = pix % Wo, t = pix / Wo;
    const int ho = t % Ho, n = t / Ho;
    const int h0 = ho * stride - pad, w0 = wo * stride - pad;
    const float* xn = x + ((int64_t)n * H * W + w0) * C;
    float v[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) {             // all loads of the pixel row in flight before the first store
      const int h = h0 + rr[j], w = w0 + dw[j];
      v[j] = 0.f;
      if (kin[j] && h >= 0 && h < H && w >= 0 && w < W) v[j] = __ldg(xn + (int64_t)h * W * C + rem[j]);
    }
    TOUT* dst = col + (int64_t)pix * KP + lane;
#pragma unroll
    for (int j = 0; j < NK; ++j) im2col_store(dst + 32 * j, v[j]);
  }
}

// dst[r, c] (+)= src[r, c] for c < cols, between two row pitches (padded <-> packed filter matrices of the stems)
__global__ void copy2d_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd,
                              int rows, int cols, int accum) {
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    const float v = src[r * lds + c];
    if (accum) dst[r * ldd + c] += v; else dst[r * ldd + c] = v;
  }
}

// lane (G, P, F>=5) -> vec (G*(P-1), 7) = [x_i, y_i, x_{i+1}, y_{i+1}, attr_{i+1}[0..2]]
__global__ void lane_to_vector_kernel(const float* __restrict__ lane, float* __restrict__ vec, int64_t G, int P) {
  int V = P - 1;
  int64_t n = G * V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t g = i / V;
    int v = (int)(i - g * V);
    const float* a = lane + (g * P + v) * 5;
    const float* b = a + 5;
    float* o = vec + i * 7;
    o[0] = a[0]; o[1] = a[1]; o[2] = b[0]; o[3] = b[1]; o[4] = b[2]; o[5] = b[3]; o[6] = b[4];
  }
}

// y[g,v,:C] = x[g,v,:], y[g,v,C:] = max_v x[g,v,:]; arg[g,c] = first argmax
__global__ void subgraph_pool_fwd_kernel(const float* __restrict__ x, int64_t G, int V, int C,
                                         float* __restrict__ y, int* __restrict__ arg) {
  int64_t n = G * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t g = i / C;
    int c = (int)(i - g * C);
    const float* xr = x + g * V * C + c;
    float best = xr[0];
    int bi = 0;
    for (int v = 1; v < V; ++v) { float t = xr[(int64_t)v * C]; if (t > best || t != t) { best = t; bi = v; } }
    arg[i] = bi;
    float* yr = y + g * V * 2 * C + c;
    for (int v = 0; v < V; ++v) { yr[(int64_t)v * 2 * C] = xr[(int64_t)v * C]; yr[(int64_t)v * 2 * C + C] = best; }
  }
}
__global__ void subgraph_pool_bwd_kernel(const float* __restrict__ dy, const int* __restrict__ arg,
                                         int64_t G, int V, int C, float* __restrict__ dx) {
  int64_t n = G * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t g = i / C;
    int c = (int)(i - g * C);
    const float* gr = dy + g * V * 2 * C + c;
    float s = 0.f;
    for (int v = 0; v < V; ++v) s += gr[(int64_t)v * 2 * C + C];
    int a = arg[i];
    float* dr = dx + g * V * C + c;
    for (int v = 0; v < V; ++v) dr[(int64_t)v * C] = gr[(int64_t)v * 2 * C] + (v == a ? s : 0.f);
  }
}
// C % 4 == 0: four channels per thread, 16-byte loads / stores, 32-bit index arithmetic (G * C / 4 < 2^31)
__global__ void __launch_bounds__(256)
subgraph_pool_bwd_vec_kernel(const float4* __restrict__ dy, const int4* __restrict__ arg, int n4, int V, int C4,
                             float4* __restrict__ dx) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const int g = i / C4, cq = i - g * C4;
    const float4* gr = dy + (int64_t)g * V * 2 * C4 + cq;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int v = 0; v < V; ++v) { const float4 t = __ldg(gr + (int64_t)v * 2 * C4 + C4); s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
    const int4 a = __ldg(arg + i);
    float4* dr = dx + (int64_t)g * V * C4 + cq;
#pragma unroll 4
    for (int v = 0; v < V; ++v) {
      float4 t = __ldg(gr + (int64_t)v * 2 * C4);
      if (v == a.x) t.x += s.x;
      if (v == a.y) t.y += s.y;
      if (v == a.z) t.z += s.z;
      if (v == a.w) t.w += s.w;
      dr[(int64_t)v * C4] = t;
    }
  }
}

// dw (64, 7) += dy^T x for the polyline input layer (lane_subgraph.layers.mlp_0.mlp.0, model_rad.py:253: 7 -> 64): dy (M, 64),
// x (M, 7), M = polylines x vectors (622 592 rows in BASELINE configs[4]).  Thread = (channel quad, one of 16 row lanes): a
// 16-byte gradient load + the row's 7 inputs (the same address for the 16 quads of a row: one broadcast transaction) feed
// 28 FMAs; the row lanes are folded through shared memory, one atomic per element and CTA.  Exact fp32.
__global__ void __launch_bounds__(256)
wgrad_n64_k7_kernel(const float4* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw, int64_t M) {
  __shared__ float red[16][449];
  const int q = threadIdx.x & 15, rsub = threadIdx.x >> 4;
  const int64_t rows_per = (M + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = blockIdx.x * rows_per, r1 = min(M, r0 + rows_per);
  float acc[4][7];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[c][k] = 0.f;
#pragma unroll 4
  for (int64_t r = r0 + rsub; r < r1; r += 16) {
    const float4 g = __ldg(dy + r * 16 + q);
    const float* xr = x + r * 7;
    float xv[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) xv[k] = __ldg(xr + k);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      acc[0][k] = fmaf(g.x, xv[k], acc[0][k]); acc[1][k] = fmaf(g.y, xv[k], acc[1][k]);
      acc[2][k] = fmaf(g.z, xv[k], acc[2][k]); acc[3][k] = fmaf(g.w, xv[k], acc[3][k]);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < 7; ++k) red[rsub][(q * 4 + c) * 7 + k] = acc[c][k];
  __syncthreads();
  for (int e = threadIdx.x; e < 448; e += 256) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += red[i][e];
    atomicAdd(dw + e, t);
  }
}

__global__ void segmax_fwd_kernel(const float* __restrict__ x, int64_t G, int V, int C,
                                  float* __restrict__ out, int* __restrict__ arg) {
  int64_t n = G * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t g = i / C;
    int c = (int)(i - g * C);
    const float* xr = x + g * V * C + c;
    float best = xr[0];
    int bi = 0;
    for (int v = 1; v < V; ++v) { float t = xr[(int64_t)v * C]; if (t > best || t != t) { best = t; bi = v; } }
    arg[i] = bi;
    out[i] = best;
  }
}
__global__ void segmax_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ arg,
                                  int64_t G, int V, int C, float* __restrict__ dx) {
  int64_t n = G * V * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t = i / C;
    int v = (int)(t % V);
    int64_t g = t / V;
    dx[i] = arg[g * C + c] == v ? dout[g * C + c] : 0.f;
  }
}
// C % 4 == 0, G * V * C / 4 < 2^31: four channels per thread (16-byte stores, 32-bit index arithmetic).  The scalar
// kernel above spent its time in 64-bit divisions and 4-byte stores: 277 us for the 319 MB of BASELINE configs[4].
__global__ void __launch_bounds__(256)
segmax_bwd_vec_kernel(const float4* __restrict__ dout, const int4* __restrict__ arg, int V, int C4, int n4, float4* __restrict__ dx) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const int cq = i % C4, t = i / C4;
    const int v = t % V, g = t / V;
    const int4 a = __ldg(arg + g * C4 + cq);
    const float4 d = __ldg(dout + g * C4 + cq);
    dx[i] = make_float4(a.x == v ? d.x : 0.f, a.y == v ? d.y : 0.f, a.z == v ? d.z : 0.f, a.w == v ? d.w : 0.f);
  }
}

// y[b, j*8+i, :] = log_softmax(v[b, i*8+j, :]) over C channels; block per (b, row)
__global__ void radar_logsoftmax_fwd_kernel(const float* __restrict__ v, float* __restrict__ y, int C) {
  __shared__ float red[32];
  int b = blockIdx.x >> 6, r = blockIdx.x & 63;
  int i = r >> 3, j = r & 7;
  const float* src = v + ((int64_t)b * 64 + r) * C;
  float* dst = y + ((int64_t)b * 64 + j * 8 + i) * C;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, src[c]);
  mx = block_max(mx, red);
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s += expf(src[c] - mx);
  s = block_sum(s, red);
  float lse = mx + logf(s);
  for (int c = threadIdx.x; c < C; c += blockDim.x) dst[c] = src[c] - lse;
}
// dv[b, i*8+j, c] = dy[b, j*8+i, c] - exp(y[b, j*8+i, c]) * sum_c dy
__global__ void radar_logsoftmax_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                            float* __restrict__ dv, int C) {
  __shared__ float red[32];
  int b = blockIdx.x >> 6, r = blockIdx.x & 63;
  int i = r >> 3, j = r & 7;
  const float* g = dy + ((int64_t)b * 64 + j * 8 + i) * C;
  const float* yy = y + ((int64_t)b * 64 + j * 8 + i) * C;
  float* dst = dv + ((int64_t)b * 64 + r) * C;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s += g[c];
  s = block_sum(s, red);
  for (int c = threadIdx.x; c < C; c += blockDim.x) dst[c] = g[c] - expf(yy[c]) * s;
}

}  // namespace

// out[n] += sum_m x[m*ld + n]   (out must be initialised by the caller)
MMFN_API int mmfn_colsum_f32(const float* x, int64_t ld, int64_t M, int N, float* out, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && out && M >= 0 && N > 0 && ld >= N, "colsum: bad args");
  if (M == 0) return 0;
  if (colsum_vec_launch<float, 4>(x, ld, M, N, out, stream)) return mmfn_launch_status("colsum");
  // 64 rows (8 loads in flight per thread) per CTA unless that exceeds ~8 waves of CTAs
  const int64_t colgroups = (N + 31) / 32;
  int64_t slabs = ceil_div64(M, 64);
  const int64_t cap = ceil_div64(148 * 8, colgroups);
  if (slabs > cap) slabs = cap;
  if (slabs > 65535) slabs = 65535;
  dim3 grid((unsigned)colgroups, (unsigned)slabs);
  colsum_kernel<<<grid, dim3(32, 8), 0, stream>>>(x, ld, M, N, out);
  return mmfn_launch_status("colsum");
}

// out[n] += sum_m x[m*ld + n] for a BF16 matrix (bias gradients of the bf16 configuration)
MMFN_API int mmfn_colsum_bf16(const void* x, int64_t ld, int64_t M, int N, float* out, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && out && M >= 0 && N > 0 && ld >= N, "colsum_bf16: bad args");
  if (M == 0) return 0;
  if (colsum_vec_launch<__nv_bfloat16, 8>((const __nv_bfloat16*)x, ld, M, N, out, stream)) return mmfn_launch_status("colsum_bf16");
  const int64_t colgroups = (N + 31) / 32;
  int64_t slabs = ceil_div64(M, 64);
  const int64_t cap = ceil_div64(148 * 8, colgroups);
  if (slabs > cap) slabs = cap;
  if (slabs > 65535) slabs = 65535;
  colsum_bf16_kernel<<<dim3((unsigned)colgroups, (unsigned)slabs), dim3(32, 8), 0, stream>>>((const __nv_bfloat16*)x, ld, M, N, out);
  return mmfn_launch_status("colsum_bf16");
}

// y (bf16) = x (fp32), n % 4 == 0, 16 / 8-byte aligned: the bf16 shadow of the master weights after a checkpoint load,
// and tensors produced by fp32-only kernels that feed a bf16 GEMM.
MMFN_API int mmfn_f32_to_bf16(const float* x, void* y, int64_t n, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && y && n >= 0 && n % 4 == 0, "f32_to_bf16: n must be a multiple of 4");
  MMFN_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 7) == 0, "f32_to_bf16: alignment");
  if (n == 0) return 0;
  f32_to_bf16_kernel<<<grid_1d(n / 4, 256), 256, 0, stream>>>((const float4*)x, (uint2*)y, n / 4);
  return mmfn_launch_status("f32_to_bf16");
}

MMFN_API int mmfn_elu_fwd(const float* x, float* y, int64_t n, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && y && n >= 0, "elu_fwd: bad args");
  if (n == 0) return 0;
  elu_fwd_kernel<<<grid_1d(n, 256), 256, 0, stream>>>(x, y, n);
  return mmfn_launch_status("elu_fwd");
}
MMFN_API int mmfn_elu_bwd(const float* dy, const float* y, float* dx, int64_t n, cudaStream_t stream) {
  MMFN_CHECK_ARG(dy && y && dx && n >= 0, "elu_bwd: bad args");
  if (n == 0) return 0;
  elu_bwd_kernel<<<grid_1d(n, 256), 256, 0, stream>>>(dy, y, dx, n);
  return mmfn_launch_status("elu_bwd");
}
MMFN_API int mmfn_relu_bwd(const float* dy, const float* y, float* dx, int64_t n, cudaStream_t stream) {
  MMFN_CHECK_ARG(dy && y && dx && n >= 0, "relu_bwd: bad args");
  if (n == 0) return 0;
  relu_bwd_kernel<<<grid_1d(n, 256), 256, 0, stream>>>(dy, y, dx, n);
  return mmfn_launch_status("relu_bwd");
}
// y = x * keep_scale(p, seed, index); same call regenerates the mask in backward.
MMFN_API int mmfn_dropout_f32(const float* x, float* y, int64_t n, float p, uint64_t seed, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && y && n >= 0 && p >= 0.f && p < 1.f, "dropout: bad args");
  if (n == 0) return 0;
  dropout_kernel<<<grid_1d(n, 256), 256, 0, stream>>>(x, y, n, p, seed);
  return mmfn_launch_status("dropout");
}
MMFN_API int mmfn_axpy_f32(const float* x, float* y, float a, int64_t n, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && y && n >= 0, "axpy: bad args");
  if (n == 0) return 0;
  axpy_kernel<<<grid_1d(n, 256), 256, 0, stream>>>(x, y, a, n);
  return mmfn_launch_status("axpy");
}
MMFN_API int mmfn_lane_to_vector(const float* lane, float* vec, int64_t G, int P, cudaStream_t stream) {
  MMFN_CHECK_ARG(lane && vec && G >= 0 && P >= 2, "lane_to_vector: bad args");
  if (G == 0) return 0;
  lane_to_vector_kernel<<<grid_1d(G * (P - 1), 256), 256, 0, stream>>>(lane, vec, G, P);
  return mmfn_launch_status("lane_to_vector");
}
MMFN_API int mmfn_subgraph_pool_fwd(const float* x, int64_t G, int V, int C, float* y, int* arg, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && y && arg && G >= 0 && V > 0 && C > 0, "subgraph_pool_fwd: bad args");
  if (G == 0) return 0;
  subgraph_pool_fwd_kernel<<<grid_1d(G * C, 256), 256, 0, stream>>>(x, G, V, C, y, arg);
  return mmfn_launch_status("subgraph_pool_fwd");
}
MMFN_API int mmfn_subgraph_pool_bwd(const float* dy, const int* arg, int64_t G, int V, int C, float* dx, cudaStream_t stream) {
  MMFN_CHECK_ARG(dy && dx && arg && G >= 0 && V > 0 && C > 0, "subgraph_pool_bwd: bad args");
  if (G == 0) return 0;
  if (C % 4 == 0 && G * C / 4 < ((int64_t)1 << 31) && (((uintptr_t)dy | (uintptr_t)arg | (uintptr_t)dx) & 15) == 0) {
    const int n4 = (int)(G * C / 4);
    subgraph_pool_bwd_vec_kernel<<<grid_1d(n4, 256), 256, 0, stream>>>((const float4*)dy, (const int4*)arg, n4, V, C / 4, (float4*)dx);
    return mmfn_launch_status("subgraph_pool_bwd");
  }
  subgraph_pool_bwd_kernel<<<grid_1d(G * C, 256), 256, 0, stream>>>(dy, arg, G, V, C, dx);
  return mmfn_launch_status("subgraph_pool_bwd");
}
// dw (64, 7) += dy^T x: weight gradient of the polyline input layer (7 -> 64, model_rad.py:253 via :263) from dy (M, 64)
// and x (M, 7), both contiguous; exact fp32, atomic accumulation into dw.
MMFN_API int mmfn_wgrad_n64_k7(const float* dy, const float* x, float* dw, int64_t M, cudaStream_t stream) {
  MMFN_CHECK_ARG(dy && x && dw && M >= 0, "wgrad_n64_k7: bad args");
  MMFN_CHECK_ARG(((uintptr_t)dy & 15) == 0, "wgrad_n64_k7: dy must be 16-byte aligned");
  if (M == 0) return 0;
  int64_t ctas = ceil_div64(M, 256);
  if (ctas > 148 * 4) ctas = 148 * 4;
  wgrad_n64_k7_kernel<<<(unsigned)ctas, 256, 0, stream>>>((const float4*)dy, x, dw, M);
  return mmfn_launch_status("wgrad_n64_k7");
}

MMFN_API int mmfn_segmax_fwd(const float* x, int64_t G, int V, int C, float* out, int* arg, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && out && arg && G >= 0 && V > 0 && C > 0, "segmax_fwd: bad args");
  if (G == 0) return 0;
  segmax_fwd_kernel<<<grid_1d(G * C, 256), 256, 0, stream>>>(x, G, V, C, out, arg);
  return mmfn_launch_status("segmax_fwd");
}
MMFN_API int mmfn_segmax_bwd(const float* dout, const int* arg, int64_t G, int V, int C, float* dx, cudaStream_t stream) {
  MMFN_CHECK_ARG(dout && dx && arg && G >= 0 && V > 0 && C > 0, "segmax_bwd: bad args");
  if (G == 0) return 0;
  if (C % 4 == 0 && G * V * C / 4 < ((int64_t)1 << 31) && (((uintptr_t)dout | (uintptr_t)arg | (uintptr_t)dx) & 15) == 0) {
    const int n4 = (int)(G * V * C / 4);
    segmax_bwd_vec_kernel<<<grid_1d(n4, 256), 256, 0, stream>>>((const float4*)dout, (const int4*)arg, V, C / 4, n4, (float4*)dx);
    return mmfn_launch_status("segmax_bwd");
  }
  segmax_bwd_kernel<<<grid_1d(G * V * C, 256), 256, 0, stream>>>(dout, arg, G, V, C, dx);
  return mmfn_launch_status("segmax_bwd");
}
// v: (B, 64, C) rows r = i*8+j ; y: NHWC (B, 8, 8, C) with y[b, j, i, :] = log_softmax(v[b, r, :])
MMFN_API int mmfn_radar_logsoftmax_fwd(const float* v, float* y, int B, int C, cudaStream_t stream) {
  MMFN_CHECK_ARG(v && y && B > 0 && C > 0, "radar_logsoftmax_fwd: bad args");
  radar_logsoftmax_fwd_kernel<<<B * 64, 128, 0, stream>>>(v, y, C);
  return mmfn_launch_status("radar_logsoftmax_fwd");
}
MMFN_API int mmfn_radar_logsoftmax_bwd(const float* dy, const float* y, float* dv, int B, int C, cudaStream_t stream) {
  MMFN_CHECK_ARG(dy && y && dv && B > 0 && C > 0, "radar_logsoftmax_bwd: bad args");
  radar_logsoftmax_bwd_kernel<<<B * 64, 128, 0, stream>>>(dy, y, dv, C);
  return mmfn_launch_status("radar_logsoftmax_bwd");
}

MMFN_DEFINE_RNG_BINDER(misc)

// col: (N*Ho*Wo, Kp) row-major, Kp % 4 == 0, Kp >= R*S*C; columns [R*S*C, Kp) are written as zeros.
MMFN_API int mmfn_im2col_nhwc(const float* x, float* col, int N, int H, int W, int C, int R, int S, int stride, int pad,
                              int Ho, int Wo, int Kp, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && col && N > 0 && H > 0 && W > 0 && C > 0 && R > 0 && S > 0 && stride > 0, "im2col: bad args");
  MMFN_CHECK_ARG(Kp % 4 == 0 && Kp >= R * S * C && (((uintptr_t)col) & 15) == 0, "im2col: Kp must be a multiple of 4, >= R*S*C; col 16-byte aligned");
  const int64_t n4 = (int64_t)N * Ho * Wo * (Kp / 4);
  const int64_t npix = (int64_t)N * Ho * Wo;
  if (S == 7 && R == 7 && ((C == 3 && Kp == 160) || (C == 2 && Kp == 128)) && npix < (1ll << 31)) {
    const int grid = grid_1d(npix * 32, 256);
    if (C == 3) im2col_rows_kernel<7, 3, 160><<<grid, 256, 0, stream>>>(x, col, N, H, W, R, stride, pad, Ho, Wo);
    else im2col_rows_kernel<7, 2, 128><<<grid, 256, 0, stream>>>(x, col, N, H, W, R, stride, pad, Ho, Wo);
    return mmfn_launch_status("im2col_nhwc");
  }
  im2col_kernel<<<grid_1d(n4, 256), 256, 0, stream>>>(x, col, N, H, W, C, R, S, stride, pad, Ho, Wo, R * S * C, Kp);
  return mmfn_launch_status("im2col_nhwc");
}

// im2col of the two stems with a BF16 column matrix (bf16 configuration: halves the 168 MB-per-16-frames matrix the
// stem GEMMs stream): 7x7 filters on 3 / 2 channels only, Kp = 192 / 128 (multiples of the 64-element bf16 k-block).
MMFN_API int mmfn_im2col_stem_bf16(const float* x, void* col, int N, int H, int W, int C, int stride, int pad,
                                   int Ho, int Wo, int Kp, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && col && N > 0 && H > 0 && W > 0 && stride > 0, "im2col_stem_bf16: bad args");
  MMFN_CHECK_ARG((C == 3 && Kp == 192) || (C == 2 && Kp == 128), "im2col_stem_bf16: C = 3 (Kp 192) or C = 2 (Kp 128) only");
  MMFN_CHECK_ARG((((uintptr_t)col) & 15) == 0, "im2col_stem_bf16: col must be 16-byte aligned");
  const int grid = grid_1d((int64_t)N * Ho * Wo * 32, 256);
  if (C == 3) im2col_rows_kernel<7, 3, 192, __nv_bfloat16><<<grid, 256, 0, stream>>>(x, (__nv_bfloat16*)col, N, H, W, 7, stride, pad, Ho, Wo);
  else im2col_rows_kernel<7, 2, 128, __nv_bfloat16><<<grid, 256, 0, stream>>>(x, (__nv_bfloat16*)col, N, H, W, 7, stride, pad, Ho, Wo);
  return mmfn_launch_status("im2col_stem_bf16");
}

// dst[r*ldd + c] (+)= src[r*lds + c], r < rows, c < cols
MMFN_API int mmfn_copy2d_f32(const float* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, int accum,
                             cudaStream_t stream) {
  MMFN_CHECK_ARG(src && dst && rows > 0 && cols > 0 && lds >= cols && ldd >= cols, "copy2d: bad args");
  copy2d_kernel<<<grid_1d((int64_t)rows * cols, 256), 256, 0, stream>>>(src, lds, dst, ldd, rows, cols, accum);
  return mmfn_launch_status("copy2d_f32");
}
