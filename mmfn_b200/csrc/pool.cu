// Layout / pooling / resampling kernels around the fusion transformers (HBM-bound):
//   NCHW->NHWC (+ImageNet normalise, model_rad.py:33-44), MaxPool 3x3/2 (:515,:521),
//   AdaptiveAvgPool(8,8) + token assembly (+pos_emb +vel_emb +dropout, :224-236, :530-533),
//   bilinear align_corners upsample + residual add (:534-539 ...), final pool+sum (:590-609).
#include "common.cuh"

namespace {

template <typename TIn>
__global__ void nchw_to_nhwc_kernel(const TIn* __restrict__ x, float* __restrict__ y,
                                    int B, int C, int64_t HW, const float* __restrict__ mean,
                                    const float* __restrict__ stdv) {
  int64_t n = (int64_t)B * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / HW, p = i - b * HW;
    for (int c = 0; c < C; ++c) {
      float v = (float)__ldg(x + (b * C + c) * HW + p);
      if (mean) v = (v - mean[c]) / stdv[c];
      y[i * C + c] = v;
    }
  }
}

// tiled batched transpose: out[b][c][r] = in[b][r][c]
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc) {
  __shared__ float t[32][33];
  const float* ib = in + (int64_t)blockIdx.z * R * Cc;
  float* ob = out + (int64_t)blockIdx.z * R * Cc;
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < Cc) t[j][threadIdx.x] = ib[(int64_t)r * Cc + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < Cc) ob[(int64_t)c * R + r] = t[threadIdx.x][j];
  }
}

// 4 channels per thread (C % 4 == 0): float4 loads, the four argmax taps packed into one 32-bit store
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ idx,
                                   int B, int H, int W, int C, int Ho, int Wo, uint2* __restrict__ y16) {
  const int C4 = C >> 2;
  const int64_t n = (int64_t)B * Ho * Wo * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int cq = (int)(i % C4);
    const int t = (int)(i / C4);                  // (b, ho, wo) flattened: < 2^31 pixels
    const int wo = t % Wo, t2 = t / Wo;
    const int ho = t2 % Ho, b = t2 / Ho;
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int bi[4] = {-1, -1, -1, -1};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = ho * 2 - 1 + r;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int sx = 0; sx < 3; ++sx) {
        const int w = wo * 2 - 1 + sx;
        if (w < 0 || w >= W) continue;
        const float4 v4 = __ldg(reinterpret_cast<const float4*>(x + (((int64_t)b * H + h) * W + w) * C) + cq);
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (bi[k] < 0 || v[k] > best[k] || v[k] != v[k]) { best[k] = v[k]; bi[k] = r * 3 + sx; }
      }
    }
    *reinterpret_cast<float4*>(y + (i << 2)) = make_float4(best[0], best[1], best[2], best[3]);
    if (y16) y16[i] = mmfn_pack_bf16x4(best[0], best[1], best[2], best[3]);
    *reinterpret_cast<uint32_t*>(idx + (i << 2)) = (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
  }
}

// 4 channels per thread (C % 4 == 0): float4 gradients, the four argmax bytes as one 32-bit load
__global__ void maxpool_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ idx,
                                   float* __restrict__ dx, int B, int H, int W, int C, int Ho, int Wo) {
  const int C4 = C >> 2;
  int64_t n = (int64_t)B * H * W * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const int t = (int)(i / C4);                  // (b, h, w) flattened: < 2^31 pixels
    const int w = t % W, t2 = t / W;
    const int h = t2 % H, b = t2 / H;
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    int ho_lo = h / 2, ho_hi = (h + 1) / 2;     // windows ho with ho*2-1 <= h <= ho*2+1
    int wo_lo = w / 2, wo_hi = (w + 1) / 2;
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
      if (ho >= Ho) continue;
      for (int wo = wo_lo; wo <= wo_hi; ++wo) {
        if (wo >= Wo) continue;
        uint32_t tap = (uint32_t)((h - (ho * 2 - 1)) * 3 + (w - (wo * 2 - 1)));
        int64_t o = (((int64_t)b * Ho + ho) * Wo + wo) * C + c;
        uint32_t am = __ldg(reinterpret_cast<const uint32_t*>(idx + o));
        if (((am & 0xffu) != tap) && (((am >> 8) & 0xffu) != tap) && (((am >> 16) & 0xffu) != tap) && ((am >> 24) != tap)) continue;
        float4 d = __ldg(reinterpret_cast<const float4*>(dy + o));
        if ((am & 0xffu) == tap) g[0] += d.x;
        if (((am >> 8) & 0xffu) == tap) g[1] += d.y;
        if (((am >> 16) & 0xffu) == tap) g[2] += d.z;
        if ((am >> 24) == tap) g[3] += d.w;
      }
    }
    *reinterpret_cast<float4*>(dx + (i << 2)) = make_float4(g[0], g[1], g[2], g[3]);
  }
}

struct FeatPtrs { const float* p[4]; };
struct FeatPtrsW { float* p[4]; };

// tokens[b, m*64 + ph*8 + pw, c] = drop(mean_{k x k} feat_m + pos_emb + vel_w*v[b] + vel_b)
// One CTA per (sample, token): the kh x kw window is spread over 256 / (C/4) pixel lanes, 4 channels per thread
// (float4 loads), reduced through shared memory; C % 4 == 0.
__global__ void __launch_bounds__(256)
tokens_fwd_kernel(FeatPtrs f, int nmod, int B, int H, int W, int C,
                  const float* __restrict__ pos, const float* __restrict__ vel_w,
                  const float* __restrict__ vel_b, const float* __restrict__ vel,
                  float* __restrict__ tok, float drop_p, uint64_t seed) {
  extern __shared__ float4 red4[];                           // [lanes][C4]
  const int T = nmod * 64, kh = H / 8, kw = W / 8, C4 = C >> 2;
  const int b = blockIdx.x / T, t = blockIdx.x - b * T;
  const int m = t >> 6, ph = (t >> 3) & 7, pw = t & 7;
  const int lanes = blockDim.x / C4;                         // host guarantees >= 1
  const int cq = threadIdx.x % C4, lane = threadIdx.x / C4;
  const float* src = f.p[m] + (((int64_t)b * H + ph * kh) * W + pw * kw) * C;
  const int npx = kh * kw;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < lanes) {
    for (int i = lane; i < npx; i += lanes) {
      const int r = i / kw, q = i - r * kw;
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + ((int64_t)r * W + q) * C) + cq);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    red4[lane * C4 + cq] = acc;
  }
  __syncthreads();
  if (threadIdx.x < C4) {
    for (int l = 1; l < lanes; ++l) {
      const float4 u = red4[l * C4 + cq];
      acc.x += u.x; acc.y += u.y; acc.z += u.z; acc.w += u.w;
    }
    const float inv = 1.0f / (float)npx, vb = vel[b];
    const int c = cq * 4;
    const float4 pe = __ldg(reinterpret_cast<const float4*>(pos + (int64_t)t * C + c));
    const float4 vw = __ldg(reinterpret_cast<const float4*>(vel_w + c)), vbias = __ldg(reinterpret_cast<const float4*>(vel_b + c));
    const int64_t i0 = ((int64_t)b * T + t) * C + c;
    float ds[4];
    mmfn_dropout_scale4(drop_p, seed, (uint64_t)i0, ds);     // i0 % 4 == 0
    float4 o;
    o.x = (acc.x * inv + pe.x + vw.x * vb + vbias.x) * ds[0];
    o.y = (acc.y * inv + pe.y + vw.y * vb + vbias.y) * ds[1];
    o.z = (acc.z * inv + pe.z + vw.z * vb + vbias.z) * ds[2];
    o.w = (acc.w * inv + pe.w + vw.w * vb + vbias.w) * ds[3];
    *reinterpret_cast<float4*>(tok + i0) = o;
  }
}

// dfeat_m[b,h,w,c] += dtok'[b, m*64 + (h/kh)*8 + w/kw, c] / (kh*kw)   (dtok' = dropout-masked dtok)
// One launch for all modalities (blockIdx.y), 4 channels per thread.
__global__ void tokens_bwd_feat_kernel(FeatPtrsW df, int B, int H, int W, int C, int T,
                                       const float* __restrict__ dtok, float drop_p, uint64_t seed) {
  const int m = blockIdx.y;
  float* dst = df.p[m];
  if (!dst) return;                              // modality without a gradient consumer
  int kh = H / 8, kw = W / 8, C4 = C >> 2;
  float inv = 1.0f / (float)(kh * kw);
  int64_t n = (int64_t)B * H * W * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C4) * 4;
    int64_t t2 = i / C4;
    int w = (int)(t2 % W); t2 /= W;
    int h = (int)(t2 % H);
    int b = (int)(t2 / H);
    int64_t ti = ((int64_t)b * T + m * 64 + (h / kh) * 8 + (w / kw)) * C + c;
    float4 g = __ldg(reinterpret_cast<const float4*>(dtok + ti));
    float4 d = *reinterpret_cast<const float4*>(dst + (i << 2));
    float ds[4];
    mmfn_dropout_scale4(drop_p, seed, (uint64_t)ti, ds);        // ti % 4 == 0 (C % 4 == 0)
    d.x += g.x * ds[0] * inv;
    d.y += g.y * ds[1] * inv;
    d.z += g.z * ds[2] * inv;
    d.w += g.w * ds[3] * inv;
    *reinterpret_cast<float4*>(dst + (i << 2)) = d;
  }
}

__global__ void tokens_bwd_param_kernel(const float* __restrict__ dtok, const float* __restrict__ vel,
                                        int B, int T, int C, float* __restrict__ dpos,
                                        float* __restrict__ dvel_w, float* __restrict__ dvel_b,
                                        float drop_p, uint64_t seed) {
  int n = T * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int c = i % C;
    float s = 0.f, sv = 0.f;
    for (int b = 0; b < B; ++b) {
      int64_t ti = (int64_t)b * n + i;
      float g = __ldg(dtok + ti) * mmfn_dropout_scale(drop_p, seed, (uint64_t)ti);
      s += g;
      sv += g * vel[b];
    }
    dpos[i] += s;
    atomicAdd(dvel_b + c, s);
    atomicAdd(dvel_w + c, sv);
  }
}

// align: src = dst * (in-1)/(out-1)  (F.interpolate(align_corners=True), model_rad.py:530-539);
// otherwise the half-pixel rule src = max(0, (dst + 0.5) * in/out - 0.5)  (align_corners=False, the default used by
// benchmarks/transfuser/model.py:338-339)
__device__ __forceinline__ float bilinear_scale(int in_size, int out_size, int align) {
  if (align) return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
  return (float)in_size / (float)out_size;
}
__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1, int align = 1) {
  float src = align ? scale * (float)dst : fmaxf(0.f, ((float)dst + 0.5f) * scale - 0.5f);
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.0f - l1;
}

// out = feat + bilinear_up(tok[:, m*64:(m+1)*64, :] as 8x8xC -> HxW); 4 channels per thread
__global__ void upsample_add_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ tok,
                                        float* __restrict__ out, int m, int T, int B, int H, int W, int C, int align,
                                        uint2* __restrict__ out16) {
  float sh = bilinear_scale(8, H, align), sw = bilinear_scale(8, W, align);
  const int C4 = C >> 2;
  int64_t n = (int64_t)B * H * W * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C4) * 4;
    int64_t t2 = i / C4;
    int w = (int)(t2 % W); t2 /= W;
    int h = (int)(t2 % H);
    int b = (int)(t2 / H);
    int h0, h1, w0, w1; float a0, a1, b0, b1;
    bilinear_src(h, sh, 8, h0, h1, a0, a1, align);
    bilinear_src(w, sw, 8, w0, w1, b0, b1, align);
    const float* tb = tok + ((int64_t)b * T + m * 64) * C + c;
    const float4 t00 = __ldg(reinterpret_cast<const float4*>(tb + (h0 * 8 + w0) * C));
    const float4 t01 = __ldg(reinterpret_cast<const float4*>(tb + (h0 * 8 + w1) * C));
    const float4 t10 = __ldg(reinterpret_cast<const float4*>(tb + (h1 * 8 + w0) * C));
    const float4 t11 = __ldg(reinterpret_cast<const float4*>(tb + (h1 * 8 + w1) * C));
    float4 f = __ldg(reinterpret_cast<const float4*>(feat + (i << 2)));
    f.x += a0 * (b0 * t00.x + b1 * t01.x) + a1 * (b0 * t10.x + b1 * t11.x);
    f.y += a0 * (b0 * t00.y + b1 * t01.y) + a1 * (b0 * t10.y + b1 * t11.y);
    f.z += a0 * (b0 * t00.z + b1 * t01.z) + a1 * (b0 * t10.z + b1 * t11.z);
    f.w += a0 * (b0 * t00.w + b1 * t01.w) + a1 * (b0 * t10.w + b1 * t11.w);
    *reinterpret_cast<float4*>(out + (i << 2)) = f;
    if (out16) out16[i] = mmfn_pack_bf16x4(f.x, f.y, f.z, f.w);
  }
}

// dtok[b, m*64+p, c] = sum_{h,w} wy(p|h) wx(p|w) dA[b,h,w,c].
// One CTA per (sample, anchor p): the (<= 2/scale + 3)^2 output pixels that can touch the anchor are spread over
// 256 / (C/4) pixel lanes, 4 channels per thread (float4 loads), then reduced through shared memory.
__global__ void __launch_bounds__(256)
upsample_add_bwd_kernel(const float* __restrict__ dA, float* __restrict__ dtok,
                        int m, int T, int B, int H, int W, int C, int align) {
  extern __shared__ float4 red4[];                           // [lanes][C4]
  const float sh = bilinear_scale(8, H, align), sw = bilinear_scale(8, W, align);
  const int C4 = C >> 2;
  const int b = blockIdx.x >> 6, p = blockIdx.x & 63;
  const int py = p >> 3, px = p & 7;
  const int lanes = blockDim.x / C4;                         // host guarantees >= 1
  const int cq = threadIdx.x % C4, lane = threadIdx.x / C4;
  int hlo = 0, hhi = H - 1, wlo = 0, whi = W - 1;
  // conservative window of output pixels whose two source taps can include anchor (py, px); the half-pixel rule
  // shifts the window by up to 0.5 / scale pixels.  Exact weights come from bilinear_src below.
  if (sh > 0.f) { int mg = align ? 1 : (int)ceilf(0.5f / sh) + 1; hlo = max(0, (int)floorf((py - 1) / sh) - mg); hhi = min(H - 1, (int)ceilf((py + 1) / sh) + mg); }
  if (sw > 0.f) { int mg = align ? 1 : (int)ceilf(0.5f / sw) + 1; wlo = max(0, (int)floorf((px - 1) / sw) - mg); whi = min(W - 1, (int)ceilf((px + 1) / sw) + mg); }
  const int nw = whi - wlo + 1, npx = (hhi - hlo + 1) * nw;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < lanes) {
    // four pixels per trip: weights first, then the (predicated) loads of all four in flight together
    for (int i0 = lane; i0 < npx; i0 += 4 * lanes) {
      float wt[4];
      const float4* src[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * lanes;
        wt[u] = 0.f;
        src[u] = nullptr;
        if (i < npx) {
          const int h = hlo + i / nw, w = wlo + i % nw;
          int h0, h1, w0, w1; float a0, a1, b0, b1;
          bilinear_src(h, sh, 8, h0, h1, a0, a1, align);
          bilinear_src(w, sw, 8, w0, w1, b0, b1, align);
          const float wy = (h0 == py ? a0 : 0.f) + (h1 == py ? a1 : 0.f);
          const float wx = (w0 == px ? b0 : 0.f) + (w1 == px ? b1 : 0.f);
          wt[u] = wy * wx;
          src[u] = reinterpret_cast<const float4*>(dA + (((int64_t)b * H + h) * W + w) * C) + cq;
        }
      }
      float4 g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) g[u] = wt[u] != 0.f ? __ldg(src[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) { acc.x += wt[u] * g[u].x; acc.y += wt[u] * g[u].y; acc.z += wt[u] * g[u].z; acc.w += wt[u] * g[u].w; }
    }
    red4[lane * C4 + cq] = acc;
  }
  __syncthreads();
  if (threadIdx.x < C4) {
    for (int l = 1; l < lanes; ++l) {
      const float4 t = red4[l * C4 + cq];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    *reinterpret_cast<float4*>(dtok + ((int64_t)b * T + m * 64 + p) * C + cq * 4) = acc;
  }
}

// fused[b,c] = sum_m mean_p (feat_m[b,p,c] + tok[b, m*64+p, c]),  P = 64 positions.
// One CTA per (sample, 32-channel group): 32 channel lanes x 8 position lanes, 8 loads in flight per thread.
__global__ void __launch_bounds__(256)
pool_sum_fwd_kernel(FeatPtrs f, int nmod, const float* __restrict__ tok, int B, int C,
                    float* __restrict__ fused) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int groups = (C + 31) / 32;
  const int b = blockIdx.x / groups, c = (blockIdx.x % groups) * 32 + cx;
  const int T = nmod * 64;
  float s = 0.f;
  if (c < C) {
    for (int m = 0; m < nmod; ++m) {
      const float* fm = f.p[m] + (int64_t)b * 64 * C + c;
      const float* tm = tok + ((int64_t)b * T + m * 64) * C + c;
#pragma unroll
      for (int p = py; p < 64; p += 8) s += __ldg(fm + (int64_t)p * C) + __ldg(tm + (int64_t)p * C);
    }
  }
  red[py][cx] = s;
  __syncthreads();
  if (py == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; ++i) s += red[i][cx];
    fused[(int64_t)b * C + c] = s * (1.0f / 64.0f);
  }
}

__global__ void pool_sum_bwd_kernel(const float* __restrict__ dfused, FeatPtrsW df, int nmod,
                                    float* __restrict__ dtok, int B, int C) {
  int T = nmod * 64;
  int64_t n = (int64_t)B * T * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t2 = i / C;
    int t = (int)(t2 % T);
    int b = (int)(t2 / T);
    float g = __ldg(dfused + b * C + c) * (1.0f / 64.0f);
    dtok[i] = g;
    df.p[t >> 6][((int64_t)b * 64 + (t & 63)) * C + c] = g;
  }
}

}  // namespace

// y(B,H,W,C) = transpose(x(B,C,H,W)); with mean/std (C floats each, device) also (x-mean)/std.
MMFN_API int mmfn_nchw_to_nhwc_f32(const float* x, float* y, int B, int C, int H, int W,
                                   const float* mean, const float* stdv, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && y && B > 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: bad args");
  MMFN_CHECK_ARG((mean == nullptr) == (stdv == nullptr), "nchw_to_nhwc: mean/std must come together");
  int64_t n = (int64_t)B * H * W;
  nchw_to_nhwc_kernel<float><<<grid_1d(n, 256), 256, 0, stream>>>(x, y, B, C, (int64_t)H * W, mean, stdv);
  return mmfn_launch_status("nchw_to_nhwc");
}

// Same for uint8 camera frames (the collated `fronts`, phase2_train_net.py:80 casts them to float).
MMFN_API int mmfn_nchw_u8_to_nhwc_f32(const uint8_t* x, float* y, int B, int C, int H, int W,
                                      const float* mean, const float* stdv, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && y && B > 0 && C > 0 && H > 0 && W > 0, "nchw_u8_to_nhwc: bad args");
  MMFN_CHECK_ARG((mean == nullptr) == (stdv == nullptr), "nchw_u8_to_nhwc: mean/std must come together");
  int64_t n = (int64_t)B * H * W;
  nchw_to_nhwc_kernel<uint8_t><<<grid_1d(n, 256), 256, 0, stream>>>(x, y, B, C, (int64_t)H * W, mean, stdv);
  return mmfn_launch_status("nchw_u8_to_nhwc");
}

// out[b][c][r] = in[b][r][c]
MMFN_API int mmfn_transpose_f32(const float* in, float* out, int nb, int R, int Cc, cudaStream_t stream) {
  MMFN_CHECK_ARG(in && out && nb > 0 && R > 0 && Cc > 0 && nb <= 65535, "transpose: bad args");
  dim3 grid((Cc + 31) / 32, (R + 31) / 32, nb);
  MMFN_CHECK_ARG(grid.y <= 65535, "transpose: too many rows");
  transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(in, out, R, Cc);
  return mmfn_launch_status("transpose");
}

// y_bf16 (nullable): bf16 twin of y (operand of the first layer-1 convolution in the bf16 configuration).
MMFN_API int mmfn_maxpool3x3s2_fwd(const float* x, float* y, uint8_t* idx, int B, int H, int W, int C,
                                   void* y_bf16, cudaStream_t stream) {
  MMFN_CHECK_ARG(x && y && idx && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "maxpool_fwd: bad args (C % 4 == 0)");
  MMFN_CHECK_ARG((((uintptr_t)x | (uintptr_t)y) & 15) == 0 && ((uintptr_t)idx & 3) == 0, "maxpool_fwd: alignment");
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  maxpool_fwd_kernel<<<grid_1d((int64_t)B * Ho * Wo * (C / 4), 256), 256, 0, stream>>>(x, y, idx, B, H, W, C, Ho, Wo, (uint2*)y_bf16);
  return mmfn_launch_status("maxpool_fwd");
}

MMFN_API int mmfn_maxpool3x3s2_bwd(const float* dy, const uint8_t* idx, float* dx, int B, int H, int W, int C,
                                   cudaStream_t stream) {
  MMFN_CHECK_ARG(dy && dx && idx && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "maxpool_bwd: bad args (C % 4 == 0)");
  MMFN_CHECK_ARG((((uintptr_t)dy | (uintptr_t)dx) & 15) == 0 && ((uintptr_t)idx & 3) == 0, "maxpool_bwd: alignment");
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  maxpool_bwd_kernel<<<grid_1d((int64_t)B * H * W * (C / 4), 256), 256, 0, stream>>>(dy, idx, dx, B, H, W, C, Ho, Wo);
  return mmfn_launch_status("maxpool_bwd");
}

// feats: nmod device pointers (each (B,H,W,C) NHWC); tokens: (B, nmod*64, C).
MMFN_API int mmfn_tokens_fwd(const float* f0, const float* f1, const float* f2, const float* f3, int nmod,
                             int B, int H, int W, int C, const float* pos_emb, const float* vel_w,
                             const float* vel_b, const float* velocity, float* tokens,
                             float drop_p, uint64_t seed, cudaStream_t stream) {
  MMFN_CHECK_ARG(nmod >= 1 && nmod <= 4 && f0 && (nmod < 2 || f1) && (nmod < 3 || f2) && (nmod < 4 || f3),
                 "tokens_fwd: bad modality pointers");
  MMFN_CHECK_ARG(pos_emb && vel_w && vel_b && velocity && tokens, "tokens_fwd: null pointer");
  MMFN_CHECK_ARG(B > 0 && C > 0 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0, "tokens_fwd: H,W must be multiples of 8");
  MMFN_CHECK_ARG(C % 4 == 0 && C <= 1024 && (((uintptr_t)f0 | (uintptr_t)f1 | (uintptr_t)f2 | (uintptr_t)f3 | (uintptr_t)pos_emb |
                                               (uintptr_t)vel_w | (uintptr_t)vel_b | (uintptr_t)tokens) & 15) == 0,
                 "tokens_fwd: C % 4 == 0, C <= 1024, 16-byte aligned buffers");
  FeatPtrs f{{f0, f1, f2, f3}};
  tokens_fwd_kernel<<<B * nmod * 64, 256, 256 * sizeof(float4), stream>>>(f, nmod, B, H, W, C, pos_emb, vel_w, vel_b,
                                                                          velocity, tokens, drop_p, seed);
  return mmfn_launch_status("tokens_fwd");
}

// df*: feature-map gradients, accumulated in place.  dpos/dvel_w/dvel_b accumulated.
MMFN_API int mmfn_tokens_bwd(const float* dtokens, float* df0, float* df1, float* df2, float* df3, int nmod,
                             int B, int H, int W, int C, const float* velocity,
                             float* dpos_emb, float* dvel_w, float* dvel_b,
                             float drop_p, uint64_t seed, cudaStream_t stream) {
  MMFN_CHECK_ARG(nmod >= 1 && nmod <= 4 && dtokens && velocity && dpos_emb && dvel_w && dvel_b, "tokens_bwd: null pointer");
  MMFN_CHECK_ARG(B > 0 && C > 0 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0, "tokens_bwd: H,W must be multiples of 8");
  FeatPtrsW df{{df0, df1, df2, df3}};
  int T = nmod * 64;
  MMFN_CHECK_ARG(C % 4 == 0 && (((uintptr_t)dtokens | (uintptr_t)df0 | (uintptr_t)df1 | (uintptr_t)df2 | (uintptr_t)df3) & 15) == 0,
                 "tokens_bwd: C % 4 == 0 and 16-byte aligned buffers");
  {
    int gx = grid_1d((int64_t)B * H * W * (C / 4), 256);
    if (gx > 148 * 4) gx = 148 * 4;
    tokens_bwd_feat_kernel<<<dim3(gx, nmod), 256, 0, stream>>>(df, B, H, W, C, T, dtokens, drop_p, seed);
  }
  tokens_bwd_param_kernel<<<grid_1d((int64_t)T * C, 128), 128, 0, stream>>>(dtokens, velocity, B, T, C, dpos_emb, dvel_w, dvel_b, drop_p, seed);
  return mmfn_launch_status("tokens_bwd");
}

// out_bf16 (nullable): bf16 twin of out (operand of the next layer's first convolution in the bf16 configuration).
MMFN_API int mmfn_upsample_add_fwd(const float* feat, const float* tokens, float* out, int m, int T,
                                   int B, int H, int W, int C, int align_corners, void* out_bf16, cudaStream_t stream) {
  MMFN_CHECK_ARG(feat && tokens && out && B > 0 && H >= 8 && W >= 8 && C > 0 && m >= 0 && (m + 1) * 64 <= T, "upsample_add_fwd: bad args");
  MMFN_CHECK_ARG(C % 4 == 0 && (((uintptr_t)feat | (uintptr_t)tokens | (uintptr_t)out) & 15) == 0, "upsample_add_fwd: C % 4 == 0, 16-byte aligned");
  upsample_add_fwd_kernel<<<grid_1d((int64_t)B * H * W * (C / 4), 256), 256, 0, stream>>>(feat, tokens, out, m, T, B, H, W, C, align_corners,
                                                                                     (uint2*)out_bf16);
  return mmfn_launch_status("upsample_add_fwd");
}

MMFN_API int mmfn_upsample_add_bwd(const float* dA, float* dtokens, int m, int T, int B, int H, int W, int C,
                                   int align_corners, cudaStream_t stream) {
  MMFN_CHECK_ARG(dA && dtokens && B > 0 && H >= 8 && W >= 8 && C > 0 && m >= 0 && (m + 1) * 64 <= T, "upsample_add_bwd: bad args");
  MMFN_CHECK_ARG(C % 4 == 0 && C <= 1024 && (((uintptr_t)dA | (uintptr_t)dtokens) & 15) == 0, "upsample_add_bwd: C % 4 == 0, C <= 1024, 16-byte aligned");
  upsample_add_bwd_kernel<<<B * 64, 256, 256 * sizeof(float4), stream>>>(dA, dtokens, m, T, B, H, W, C, align_corners);
  return mmfn_launch_status("upsample_add_bwd");
}

MMFN_API int mmfn_pool_sum_fwd(const float* f0, const float* f1, const float* f2, const float* f3, int nmod,
                               const float* tokens, int B, int C, float* fused, cudaStream_t stream) {
  MMFN_CHECK_ARG(nmod >= 1 && nmod <= 4 && f0 && tokens && fused && B > 0 && C > 0, "pool_sum_fwd: bad args");
  FeatPtrs f{{f0, f1, f2, f3}};
  pool_sum_fwd_kernel<<<B * ((C + 31) / 32), 256, 0, stream>>>(f, nmod, tokens, B, C, fused);
  return mmfn_launch_status("pool_sum_fwd");
}

MMFN_API int mmfn_pool_sum_bwd(const float* dfused, float* df0, float* df1, float* df2, float* df3, int nmod,
                               float* dtokens, int B, int C, cudaStream_t stream) {
  MMFN_CHECK_ARG(nmod >= 1 && nmod <= 4 && dfused && df0 && dtokens && B > 0 && C > 0, "pool_sum_bwd: bad args");
  MMFN_CHECK_ARG((nmod < 2 || df1) && (nmod < 3 || df2) && (nmod < 4 || df3), "pool_sum_bwd: null modality gradient");
  FeatPtrsW df{{df0, df1, df2, df3}};
  pool_sum_bwd_kernel<<<grid_1d((int64_t)B * nmod * 64 * C, 256), 256, 0, stream>>>(dfused, df, nmod, dtokens, B, C);
  return mmfn_launch_status("pool_sum_bwd");
}

MMFN_DEFINE_RNG_BINDER(pool)
