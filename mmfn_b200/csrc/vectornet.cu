// Fused VectorNet polyline Subgraph forward (model_rad.py:260-283 Subgraph, :249-258 MLP, :369-382 _lane_to_vector):
//
//   vec[v] = [lane[v].xy, lane[v+1].xy, lane[v+1].attr3]                            (V = P - 1 vectors of 7 features)
//   for i in 0..2:  y_i = x_i W_i^T + b_i ;  h_i = ReLU(LayerNorm(y_i)) ;  m_i = max_v h_i ;  x_{i+1} = [h_i | m_i]
//   tok = max_v x_3 = [m_2 | m_2]
//
// ONE launch instead of lane_to_vector + 3 x (GEMM + LayerNorm + max-pool/concat) + segment-max, the design north_star
// names: a polyline lives in (part of) one warp -- lane = one vector node -- so
//   * the small GEMMs (7 -> 64, 128 -> 64) are per-lane register GEMMs: 64 fp32 accumulators per lane, the weights
//     (k-major, shared by the CTA) broadcast from shared memory as 16-byte vectors;
//   * LayerNorm is lane-local (a lane owns the 64 channels of its vector);
//   * the segment max-pool over the polyline's nodes is a warp reduction: REDUX.MAX over the lanes of the polyline on the
//     (non-negative, hence order-preserving) float bit patterns, the arg-max (first node on ties, as the unfused kernels)
//     from a ballot;
//   * the broadcast half of the next layer's input ([h | m]: the same m for every node) enters the next GEMM ONCE per
//     polyline: its 64 x 64 product is split over the polyline's lanes and handed round with shuffles.
// V <= 32; floor(32 / V) polylines share a warp (3 for the 10-node lanes of the reference, 1 for the 20-node lanes of
// BASELINE configs[4]).  Persistent CTAs: the 67 KB of weights are staged into shared memory once per CTA.
// The kernel writes exactly what the (unfused) backward consumes: the inputs of the three linears (vec, x_1, x_2), their
// pre-LayerNorm outputs y_i with the row statistics, the arg-max routing tables and the polyline tokens.
#include "common.cuh"

namespace {

constexpr int SG_THREADS = 384;                 // 12 warps, one CTA per SM (64 + 64 accumulators per lane)
constexpr int HID = 64;

struct SubgraphParams {
  const float* lane;                            // (G, P, 5)
  const float* w[3]; const float* b[3]; const float* gamma[3]; const float* beta[3];
  float* vec;                                   // (G*V, 7)
  float* y[3];                                  // (G*V, 64) pre-LayerNorm
  float* mean[3]; float* rstd[3];               // (G*V)
  float* x1; float* x2;                         // (G*V, 128) = [h | max]
  int* arg[3];                                  // (G, 64)
  float* tok; int* argf;                        // (G, 128)
  long long G;
  float eps;
};

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <int V>
__global__ void __launch_bounds__(SG_THREADS, 1)
subgraph_fused_fwd_kernel(SubgraphParams p) {
  constexpr int PPW = 32 / V;                   // polylines per warp
  constexpr int NCM = (HID + V - 1) / V;        // outputs of the broadcast-half product per lane
  constexpr int P = V + 1;
  extern __shared__ float sm[];
  float* w0t = sm;                              // [7][64]    k-major copies of the weights
  float* w1t = w0t + 7 * HID;                   // [128][64]
  float* w2t = w1t + 128 * HID;                 // [128][64]
  float* par = w2t + 128 * HID;                 // [3][3][64]  bias, gamma, beta per layer
  float* mst = par + 9 * HID;                   // [warps][PPW][64]  pooled maxima of the current layer (per polyline)
  float* hst = mst + (SG_THREADS / 32) * PPW * HID;   // [64][threads]  this thread's activations, k-major (bank = thread)
  for (int i = threadIdx.x; i < 7 * HID; i += SG_THREADS) { const int k = i / HID, c = i - k * HID; w0t[i] = p.w[0][c * 7 + k]; }
  // (64,128) row-major weights -> k-major copies: coalesced global reads into a padded staging tile (the activation
  // buffer, not yet in use), transposed on the way out of it (stride 129: conflict-free both ways)
  for (int m = 1; m <= 2; ++m) {
    float* dst = m == 1 ? w1t : w2t;
    for (int i = threadIdx.x; i < 128 * HID; i += SG_THREADS) { const int c = i >> 7, k = i & 127; hst[c * 129 + k] = p.w[m][i]; }
    __syncthreads();
    for (int i = threadIdx.x; i < 128 * HID; i += SG_THREADS) { const int k = i / HID, c = i - k * HID; dst[i] = hst[c * 129 + k]; }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 3 * HID; i += SG_THREADS) {
    const int l = i / HID, c = i - l * HID;
    par[(l * 3 + 0) * HID + c] = p.b[l][c];
    par[(l * 3 + 1) * HID + c] = p.gamma[l][c];
    par[(l * 3 + 2) * HID + c] = p.beta[l][c];
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = lane / V, v = lane - slot * V;
  const bool in_slot = slot < PPW;
  const unsigned gmask = in_slot ? (((V == 32) ? 0xffffffffu : ((1u << V) - 1u)) << (slot * V)) : 0u;
  const int gstart = slot * V;
  float* mrow = mst + (warp * PPW + (in_slot ? slot : 0)) * HID;
  const long long ngroups = (p.G + PPW - 1) / PPW;
  const long long wstride = (long long)gridDim.x * (SG_THREADS / 32);

  for (long long grp = (long long)blockIdx.x * (SG_THREADS / 32) + warp; grp < ngroups; grp += wstride) {
    const long long g = grp * PPW + slot;
    const bool on = in_slot && g < p.G;                     // uniform within a polyline's lanes
    if (!on) continue;                                      // (whole slots drop out together: group ops stay consistent)
    const long long row = g * V + v;
    float y[HID];
    float* hcol = hst + threadIdx.x;
    // ---- polyline vectorisation + layer 0 (7 -> 64)
    {
      const float* a = p.lane + (g * P + v) * 5;
      const float xin[7] = {a[0], a[1], a[5], a[6], a[7], a[8], a[9]};
      float* vo = p.vec + row * 7;
#pragma unroll
      for (int k = 0; k < 7; ++k) vo[k] = xin[k];
#pragma unroll
      for (int c = 0; c < HID; ++c) y[c] = par[0 * HID + c];
#pragma unroll
      for (int k = 0; k < 7; ++k) {
#pragma unroll
        for (int c = 0; c < HID; c += 4) {
          const float4 w = lds4(w0t + k * HID + c);
          y[c] = fmaf(w.x, xin[k], y[c]); y[c + 1] = fmaf(w.y, xin[k], y[c + 1]);
          y[c + 2] = fmaf(w.z, xin[k], y[c + 2]); y[c + 3] = fmaf(w.w, xin[k], y[c + 3]);
        }
      }
    }
#pragma unroll
    for (int layer = 0; layer < 3; ++layer) {
      const float* gam = par + (layer * 3 + 1) * HID;
      const float* bet = par + (layer * 3 + 2) * HID;
      // ---- y_layer is complete: save it, LayerNorm + ReLU lane-locally (two-pass variance, as ln_fwd_kernel)
      float* yo = p.y[layer] + row * HID;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < HID; c += 4) {
        *reinterpret_cast<float4*>(yo + c) = make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]);
        s += (y[c] + y[c + 1]) + (y[c + 2] + y[c + 3]);
      }
      const float mu = s / (float)HID;
      float q = 0.f;
#pragma unroll
      for (int c = 0; c < HID; ++c) { const float d = y[c] - mu; q = fmaf(d, d, q); }
      const float rs = rsqrtf(q / (float)HID + p.eps);
      p.mean[layer][row] = mu;
      p.rstd[layer][row] = rs;
      // ---- ReLU(LN(y)); segment max over the polyline's nodes + arg-max; the broadcast half of the NEXT layer's GEMM on
      //      the fly; x_{layer+1}[:, :64] = h for the weight gradient of the next linear
      const float* wn = layer == 0 ? w1t : w2t;             // next layer's weights (unused after layer 2)
      float cm[NCM];
#pragma unroll
      for (int j = 0; j < NCM; ++j) cm[j] = 0.f;
      int* argo = p.arg[layer] + g * HID;
      float* xo = (layer == 0 ? p.x1 : p.x2) + row * 2 * HID;
#pragma unroll
      for (int c4 = 0; c4 < HID; c4 += 4) {
        float h4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c4 + u;
          const float hc = fmaxf(fmaf((y[c] - mu) * rs, gam[c], bet[c]), 0.f);
          h4[u] = hc;
          const int hb = __float_as_int(hc);                // h >= 0: the bit pattern orders like the value (-0 lowest)
          const int mb = __reduce_max_sync(gmask, hb);
          const unsigned who = __ballot_sync(gmask, hb == mb) & gmask;
          const float m = __int_as_float(mb);
          if (v == 0) { argo[c] = __ffs(who) - 1 - gstart; mrow[c] = m; }
          if (layer < 2) {
            hcol[c * SG_THREADS] = hc;
#pragma unroll
            for (int j = 0; j < NCM; ++j) {
              const int co = v + j * V;                     // this lane's outputs of the broadcast-half product
              if (co < HID) cm[j] = fmaf(wn[(HID + c) * HID + co], m, cm[j]);
            }
          }
        }
        if (layer < 2) *reinterpret_cast<float4*>(xo + c4) = make_float4(h4[0], h4[1], h4[2], h4[3]);
      }
      __syncwarp(gmask);                                    // mrow visible to the polyline's lanes
      if (layer == 2) {                                     // tok = [m_2 | m_2]; arg-max of the broadcast half is node 0
        float* to = p.tok + g * 2 * HID;
        int* af = p.argf + g * 2 * HID;
        for (int c = v; c < HID; c += V) {
          const float m = mrow[c];
          to[c] = m; to[HID + c] = m;
          af[c] = argo[c]; af[HID + c] = 0;
        }
        break;
      }
#pragma unroll
      for (int c = 0; c < HID; c += 4) *reinterpret_cast<float4*>(xo + HID + c) = lds4(mrow + c);
      // ---- next layer: y = b + W[:, 64:] m (once per polyline, handed round by shuffles) + W[:, :64] h (per lane)
      const float* bn = par + ((layer + 1) * 3 + 0) * HID;
#pragma unroll
      for (int c = 0; c < HID; ++c) {
        const float part = __shfl_sync(gmask, cm[c / V], gstart + (c % V));
        y[c] = bn[c] + part;
      }
#pragma unroll 4
      for (int k = 0; k < HID; ++k) {
        const float hk = hcol[k * SG_THREADS];
#pragma unroll
        for (int c = 0; c < HID; c += 4) {
          const float4 w = lds4(wn + k * HID + c);
          y[c] = fmaf(w.x, hk, y[c]); y[c + 1] = fmaf(w.y, hk, y[c + 1]);
          y[c + 2] = fmaf(w.z, hk, y[c + 2]); y[c + 3] = fmaf(w.w, hk, y[c + 3]);
        }
      }
      __syncwarp(gmask);                                    // everyone has read mrow before the next layer overwrites it
    }
  }
}

template <int V>
int launch_subgraph(const SubgraphParams& p, cudaStream_t stream) {
  constexpr int PPW = 32 / V;
  const int smem = (7 * HID + 2 * 128 * HID + 9 * HID + (SG_THREADS / 32) * PPW * HID + HID * SG_THREADS) * (int)sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(subgraph_fused_fwd_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (ce != cudaSuccess) { mmfn_set_error("subgraph_fused_fwd: smem attribute: %s", cudaGetErrorString(ce)); return (int)ce; }
    attr_set = true;
  }
  const long long ngroups = (p.G + PPW - 1) / PPW;
  const long long need = (ngroups + SG_THREADS / 32 - 1) / (SG_THREADS / 32);
  const int grid = (int)(need < 148 ? need : 148);
  subgraph_fused_fwd_kernel<V><<<grid, SG_THREADS, smem, stream>>>(p);
  return mmfn_launch_status("subgraph_fused_fwd");
}

}  // namespace

// lane: (G, V + 1, 5) polyline nodes; w0 (64,7), w1 / w2 (64,128), biases and LayerNorm parameters (64) of the three
// Subgraph layers (reference keys lane_subgraph.layers.mlp_{0,1,2}.mlp.{0,1}).  Outputs (caller-allocated): vec (G*V,7);
// y0..y2 (G*V,64) pre-LayerNorm linear outputs; mean*/rstd* (G*V); x1, x2 (G*V,128) inputs of layers 1 / 2; arg0..arg2
// (G,64) node index of each pooled maximum; tok (G,128) polyline tokens; argf (G,128) arg-max of the final pool.
// V in {9, 19} (10- / 20-node lanes): other lane lengths take the unfused kernels.
MMFN_API int mmfn_subgraph_fused_fwd(const float* lane, int64_t G, int V,
                                     const float* w0, const float* b0, const float* g0, const float* be0,
                                     const float* w1, const float* b1, const float* g1, const float* be1,
                                     const float* w2, const float* b2, const float* g2, const float* be2,
                                     float* vec, float* y0, float* y1, float* y2,
                                     float* mean0, float* rstd0, float* mean1, float* rstd1, float* mean2, float* rstd2,
                                     float* x1, float* x2, int* arg0, int* arg1, int* arg2, float* tok, int* argf,
                                     float eps, cudaStream_t stream) {
  MMFN_CHECK_ARG(lane && w0 && b0 && g0 && be0 && w1 && b1 && g1 && be1 && w2 && b2 && g2 && be2, "subgraph_fused_fwd: null parameter");
  MMFN_CHECK_ARG(vec && y0 && y1 && y2 && mean0 && rstd0 && mean1 && rstd1 && mean2 && rstd2 && x1 && x2 && arg0 && arg1 && arg2 && tok && argf,
                 "subgraph_fused_fwd: null output");
  MMFN_CHECK_ARG(G >= 0 && (V == 9 || V == 19), "subgraph_fused_fwd: V must be 9 or 19 (10- / 20-node lanes)");
  MMFN_CHECK_ARG((((uintptr_t)y0 | (uintptr_t)y1 | (uintptr_t)y2 | (uintptr_t)x1 | (uintptr_t)x2) & 15) == 0, "subgraph_fused_fwd: 16-byte alignment");
  if (G == 0) return 0;
  SubgraphParams p;
  p.lane = lane;
  p.w[0] = w0; p.w[1] = w1; p.w[2] = w2; p.b[0] = b0; p.b[1] = b1; p.b[2] = b2;
  p.gamma[0] = g0; p.gamma[1] = g1; p.gamma[2] = g2; p.beta[0] = be0; p.beta[1] = be1; p.beta[2] = be2;
  p.vec = vec; p.y[0] = y0; p.y[1] = y1; p.y[2] = y2;
  p.mean[0] = mean0; p.mean[1] = mean1; p.mean[2] = mean2; p.rstd[0] = rstd0; p.rstd[1] = rstd1; p.rstd[2] = rstd2;
  p.x1 = x1; p.x2 = x2; p.arg[0] = arg0; p.arg[1] = arg1; p.arg[2] = arg2; p.tok = tok; p.argf = argf;
  p.G = G; p.eps = eps;
  if (V == 9) return launch_subgraph<9>(p, stream);
  return launch_subgraph<19>(p, stream);
}
