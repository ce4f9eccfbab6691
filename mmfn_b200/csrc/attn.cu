// Softmax-family kernels: fusion-transformer attention probabilities (model_rad.py:101-103),
// radar GAT masked attention (:808-821) and the VectorNet lane-to-lane attention restricted
// to query row 0, the only row consumed downstream (:405-413).
#include "common.cuh"

namespace {

// one warp per row; cols <= 32*J (J = 6 / 8 for the 192- / 256-token fusion transformers)
template <int J>
__global__ void softmax_fwd_kernel(const float* __restrict__ s, float* __restrict__ p, float* __restrict__ pd,
                                   int64_t rows, int cols, float scale, float drop_p, uint64_t seed) {
  int lane = threadIdx.x & 31;
  int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* sr = s + row * cols;
  float v[J];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    int c = lane + 32 * j;
    v[j] = c < cols ? sr[c] * scale : -INFINITY;
    mx = fmaxf(mx, v[j]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    int c = lane + 32 * j;
    v[j] = c < cols ? expf(v[j] - mx) : 0.f;
    sum += v[j];
  }
  float inv = 1.0f / warp_sum(sum);
#pragma unroll
  for (int j = 0; j < J; ++j) {
    int c = lane + 32 * j;
    if (c < cols) {
      float q = v[j] * inv;
      p[row * cols + c] = q;
      if (pd) pd[row * cols + c] = q * mmfn_dropout_scale(drop_p, seed, (uint64_t)(row * cols + c));
    }
  }
}

// ds = scale * p * (dp - sum(dp*p)),  dp = dpd * dropout_scale
template <int J>
__global__ void softmax_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dpd, float* __restrict__ ds,
                                   int64_t rows, int cols, float scale, float drop_p, uint64_t seed) {
  int lane = threadIdx.x & 31;
  int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float pv[J], g[J];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    int c = lane + 32 * j;
    pv[j] = 0.f; g[j] = 0.f;
    if (c < cols) {
      int64_t i = row * cols + c;
      pv[j] = p[i];
      g[j] = dpd[i] * mmfn_dropout_scale(drop_p, seed, (uint64_t)i);
      dot += pv[j] * g[j];
    }
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int j = 0; j < J; ++j) {
    int c = lane + 32 * j;
    if (c < cols) ds[row * cols + c] = scale * pv[j] * (g[j] - dot);
  }
}

// float4 variant (cols % 4 == 0, cols <= 128 * J4): lane handles 4 consecutive keys per step, one dropout hash per vector
template <int J4>
__global__ void softmax_bwd_v4_kernel(const float4* __restrict__ p, const float4* __restrict__ dpd, float4* __restrict__ ds,
                                      int64_t rows, int cols4, float scale, float drop_p, uint64_t seed) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4 pv[J4], g[J4];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < J4; ++j) {
    const int c = lane + 32 * j;
    pv[j] = make_float4(0.f, 0.f, 0.f, 0.f); g[j] = pv[j];
    if (c < cols4) {
      const int64_t i = row * cols4 + c;
      pv[j] = __ldg(p + i);
      g[j] = __ldg(dpd + i);
    }
  }
#pragma unroll
  for (int j = 0; j < J4; ++j) {
    const int c = lane + 32 * j;
    if (c < cols4) {
      float dsc[4];
      mmfn_dropout_scale4(drop_p, seed, (uint64_t)(row * cols4 + c) << 2, dsc);
      g[j].x *= dsc[0]; g[j].y *= dsc[1]; g[j].z *= dsc[2]; g[j].w *= dsc[3];
      dot += pv[j].x * g[j].x + pv[j].y * g[j].y + pv[j].z * g[j].z + pv[j].w * g[j].w;
    }
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int j = 0; j < J4; ++j) {
    const int c = lane + 32 * j;
    if (c < cols4)
      ds[row * cols4 + c] = make_float4(scale * pv[j].x * (g[j].x - dot), scale * pv[j].y * (g[j].y - dot),
                                        scale * pv[j].z * (g[j].z - dot), scale * pv[j].w * (g[j].w - dot));
  }
}

// GAT: att = softmax(where(adj > 0, leakyrelu(z), -9e15)); one warp per row, cols <= 128
__global__ void gat_softmax_fwd_kernel(const float* __restrict__ z, const float* __restrict__ adj,
                                       float* __restrict__ att, float* __restrict__ attd,
                                       int64_t rows, int cols, float alpha, float drop_p, uint64_t seed) {
  int lane = threadIdx.x & 31;
  int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[4];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = lane + 32 * j;
    v[j] = -INFINITY;
    if (c < cols) {
      int64_t i = row * cols + c;
      float e = z[i];
      e = e > 0.f ? e : alpha * e;
      v[j] = adj[i] > 0.f ? e : -9e15f;
    }
    mx = fmaxf(mx, v[j]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = lane + 32 * j;
    v[j] = c < cols ? expf(v[j] - mx) : 0.f;
    sum += v[j];
  }
  float inv = 1.0f / warp_sum(sum);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = lane + 32 * j;
    if (c < cols) {
      int64_t i = row * cols + c;
      float q = v[j] * inv;
      att[i] = q;
      attd[i] = q * mmfn_dropout_scale(drop_p, seed, (uint64_t)i);
    }
  }
}

__global__ void gat_softmax_bwd_kernel(const float* __restrict__ z, const float* __restrict__ adj,
                                       const float* __restrict__ att, const float* __restrict__ dattd,
                                       float* __restrict__ dz, int64_t rows, int cols, float alpha,
                                       float drop_p, uint64_t seed) {
  int lane = threadIdx.x & 31;
  int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float pv[4], g[4];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = lane + 32 * j;
    pv[j] = 0.f; g[j] = 0.f;
    if (c < cols) {
      int64_t i = row * cols + c;
      pv[j] = att[i];
      g[j] = dattd[i] * mmfn_dropout_scale(drop_p, seed, (uint64_t)i);
      dot += pv[j] * g[j];
    }
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = lane + 32 * j;
    if (c < cols) {
      int64_t i = row * cols + c;
      float de = pv[j] * (g[j] - dot);
      float zz = z[i];
      dz[i] = adj[i] > 0.f ? de * (zz > 0.f ? 1.f : alpha) : 0.f;
    }
  }
}

// VectorNet L2L attention, query row 0 only.  qkv: (B, L, 3*H*D) = [q | k | v].  block per (b, head).
template <int D>
__global__ void l2l_row0_fwd_kernel(const float* __restrict__ qkv, const int* __restrict__ lane_num,
                                    int L, int H, float scale, float* __restrict__ prob, float* __restrict__ out) {
  extern __shared__ float sh[];   // L scores + 32 reduce
  float* sc = sh;
  float* red = sh + L;
  int b = blockIdx.x / H, h = blockIdx.x % H;
  int HD = H * D, ld = 3 * HD;
  const float* base = qkv + (int64_t)b * L * ld;
  const float* q0 = base + h * D;
  int nvalid = lane_num[b];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = warp; j < L; j += nw) {
    const float* kj = base + (int64_t)j * ld + HD + h * D;
    float d = 0.f;
    for (int t = lane; t < D; t += 32) d += q0[t] * kj[t];
    d = warp_sum(d) * scale;
    if (lane == 0) sc[j] = j < nvalid ? d : -1e9f;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < L; j += blockDim.x) mx = fmaxf(mx, sc[j]);
  mx = block_max(mx, red);
  float sum = 0.f;
  for (int j = threadIdx.x; j < L; j += blockDim.x) { float e = expf(sc[j] - mx); sc[j] = e; sum += e; }
  sum = block_sum(sum, red);
  float inv = 1.0f / sum;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    float p = sc[j] * inv;
    sc[j] = p;
    prob[((int64_t)b * H + h) * L + j] = p;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < D; t += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < L; ++j) acc += sc[j] * base[(int64_t)j * ld + 2 * HD + h * D + t];
    out[(int64_t)b * HD + h * D + t] = acc;
  }
}

template <int D>
__global__ void l2l_row0_bwd_kernel(const float* __restrict__ qkv, const int* __restrict__ lane_num,
                                    const float* __restrict__ prob, const float* __restrict__ dout,
                                    int L, int H, float scale, float* __restrict__ dqkv) {
  extern __shared__ float sh[];   // L ds + 32 reduce
  float* ds = sh;
  float* red = sh + L;
  int b = blockIdx.x / H, h = blockIdx.x % H;
  int HD = H * D, ld = 3 * HD;
  const float* base = qkv + (int64_t)b * L * ld;
  float* dbase = dqkv + (int64_t)b * L * ld;
  const float* q0 = base + h * D;
  const float* go = dout + (int64_t)b * HD + h * D;
  const float* pr = prob + ((int64_t)b * H + h) * L;
  int nvalid = lane_num[b];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // dp_j = dout . v_j ; dv_j = p_j * dout
  for (int j = warp; j < L; j += nw) {
    const float* vj = base + (int64_t)j * ld + 2 * HD + h * D;
    float* dvj = dbase + (int64_t)j * ld + 2 * HD + h * D;
    float p = pr[j], d = 0.f;
    for (int t = lane; t < D; t += 32) { d += go[t] * vj[t]; dvj[t] = p * go[t]; }
    d = warp_sum(d);
    if (lane == 0) ds[j] = d;
  }
  __syncthreads();
  float dot = 0.f;
  for (int j = threadIdx.x; j < L; j += blockDim.x) dot += pr[j] * ds[j];
  dot = block_sum(dot, red);
  for (int j = threadIdx.x; j < L; j += blockDim.x) ds[j] = j < nvalid ? pr[j] * (ds[j] - dot) * scale : 0.f;
  __syncthreads();
  // dk_j = ds_j * q0 ; dq_j = 0 for j > 0
  for (int j = warp; j < L; j += nw) {
    float* dkj = dbase + (int64_t)j * ld + HD + h * D;
    float* dqj = dbase + (int64_t)j * ld + h * D;
    float g = ds[j];
    for (int t = lane; t < D; t += 32) { dkj[t] = g * q0[t]; if (j > 0) dqj[t] = 0.f; }
  }
  for (int t = threadIdx.x; t < D; t += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < L; ++j) acc += ds[j] * base[(int64_t)j * ld + HD + h * D + t];
    dbase[h * D + t] = acc;
  }
}

}  // namespace

// p = softmax(scale * s) row-wise; pd (nullable) = dropout(p).  p may alias s.
MMFN_API int mmfn_softmax_fwd(const float* s, float* p, float* pd, int64_t rows, int cols, float scale,
                              float drop_p, uint64_t seed, cudaStream_t stream) {
  MMFN_CHECK_ARG(s && p && rows >= 0 && cols > 0 && cols <= 1024, "softmax_fwd: bad args (cols <= 1024)");
  if (rows == 0) return 0;
  unsigned grid = (unsigned)ceil_div64(rows, 8);
  if (cols <= 192) softmax_fwd_kernel<6><<<grid, 256, 0, stream>>>(s, p, pd, rows, cols, scale, drop_p, seed);
  else if (cols <= 256) softmax_fwd_kernel<8><<<grid, 256, 0, stream>>>(s, p, pd, rows, cols, scale, drop_p, seed);
  else softmax_fwd_kernel<32><<<grid, 256, 0, stream>>>(s, p, pd, rows, cols, scale, drop_p, seed);
  return mmfn_launch_status("softmax_fwd");
}

MMFN_API int mmfn_softmax_bwd(const float* p, const float* dpd, float* ds, int64_t rows, int cols, float scale,
                              float drop_p, uint64_t seed, cudaStream_t stream) {
  MMFN_CHECK_ARG(p && dpd && ds && rows >= 0 && cols > 0 && cols <= 1024, "softmax_bwd: bad args (cols <= 1024)");
  if (rows == 0) return 0;
  unsigned grid = (unsigned)ceil_div64(rows, 8);
  if (cols % 4 == 0 && cols <= 256 && (((uintptr_t)p | (uintptr_t)dpd | (uintptr_t)ds) & 15) == 0)
    softmax_bwd_v4_kernel<2><<<grid, 256, 0, stream>>>((const float4*)p, (const float4*)dpd, (float4*)ds, rows, cols / 4, scale, drop_p, seed);
  else if (cols <= 192) softmax_bwd_kernel<6><<<grid, 256, 0, stream>>>(p, dpd, ds, rows, cols, scale, drop_p, seed);
  else if (cols <= 256) softmax_bwd_kernel<8><<<grid, 256, 0, stream>>>(p, dpd, ds, rows, cols, scale, drop_p, seed);
  else softmax_bwd_kernel<32><<<grid, 256, 0, stream>>>(p, dpd, ds, rows, cols, scale, drop_p, seed);
  return mmfn_launch_status("softmax_bwd");
}

MMFN_API int mmfn_gat_softmax_fwd(const float* z, const float* adj, float* att, float* att_drop,
                                  int64_t rows, int cols, float alpha, float drop_p, uint64_t seed,
                                  cudaStream_t stream) {
  MMFN_CHECK_ARG(z && adj && att && att_drop && rows >= 0 && cols > 0 && cols <= 128, "gat_softmax_fwd: bad args (cols <= 128)");
  if (rows == 0) return 0;
  gat_softmax_fwd_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, stream>>>(z, adj, att, att_drop, rows, cols, alpha, drop_p, seed);
  return mmfn_launch_status("gat_softmax_fwd");
}

MMFN_API int mmfn_gat_softmax_bwd(const float* z, const float* adj, const float* att, const float* datt_drop,
                                  float* dz, int64_t rows, int cols, float alpha, float drop_p, uint64_t seed,
                                  cudaStream_t stream) {
  MMFN_CHECK_ARG(z && adj && att && datt_drop && dz && rows >= 0 && cols > 0 && cols <= 128, "gat_softmax_bwd: bad args (cols <= 128)");
  if (rows == 0) return 0;
  gat_softmax_bwd_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, stream>>>(z, adj, att, datt_drop, dz, rows, cols, alpha, drop_p, seed);
  return mmfn_launch_status("gat_softmax_bwd");
}

// qkv (B,L,3*heads*64); lane_num (B) int32; prob (B,heads,L); out (B,heads*64).
MMFN_API int mmfn_l2l_row0_fwd(const float* qkv, const int* lane_num, int B, int L, int heads, int dim_head,
                               float* prob, float* out, cudaStream_t stream) {
  MMFN_CHECK_ARG(qkv && lane_num && prob && out && B > 0 && L > 0 && heads > 0, "l2l_fwd: bad args");
  MMFN_CHECK_ARG(dim_head == 64, "l2l_fwd: dim_head must be 64");
  MMFN_CHECK_ARG(L <= 8192, "l2l_fwd: too many lanes");
  size_t smem = sizeof(float) * (L + 32);
  l2l_row0_fwd_kernel<64><<<B * heads, 256, smem, stream>>>(qkv, lane_num, L, heads, 1.0f / sqrtf((float)dim_head), prob, out);
  return mmfn_launch_status("l2l_row0_fwd");
}

MMFN_API int mmfn_l2l_row0_bwd(const float* qkv, const int* lane_num, const float* prob, const float* dout,
                               int B, int L, int heads, int dim_head, float* dqkv, cudaStream_t stream) {
  MMFN_CHECK_ARG(qkv && lane_num && prob && dout && dqkv && B > 0 && L > 0 && heads > 0, "l2l_bwd: bad args");
  MMFN_CHECK_ARG(dim_head == 64, "l2l_bwd: dim_head must be 64");
  MMFN_CHECK_ARG(L <= 8192, "l2l_bwd: too many lanes");
  size_t smem = sizeof(float) * (L + 32);
  l2l_row0_bwd_kernel<64><<<B * heads, 256, smem, stream>>>(qkv, lane_num, prob, dout, L, heads, 1.0f / sqrtf((float)dim_head), dqkv);
  return mmfn_launch_status("l2l_row0_bwd");
}

MMFN_DEFINE_RNG_BINDER(attn)
