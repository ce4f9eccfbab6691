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

// ------------------------------------------------------------------------------------------------------------------
// Tensor-core formulation (TF32 configuration): the same sub-graph, the 64-wide linears as warp-level mma.sync m16n8k8
// TF32 tiles.  A persistent CTA holds two groups of 8 warps, each walking tiles of floor(128 / V) whole polylines (126 / 114
// node rows); warp =
// one 16-row m-tile x all 64 channels, so a row's LayerNorm is a reduction over the four lanes of a quad; activations
// stay in shared memory between layers (fp32, rounded to TF32 as the next layer's A fragments are loaded); the
// segment max-pool scans the tile's rows in shared memory, and the broadcast half of the next layer's input ([h | m]:
// the same m for every node of a polyline) is multiplied ONCE per polyline as a (polylines x 64) x (64 x 64) product
// whose result seeds the accumulators.  The lane-per-node SIMT kernel above spends a 16-byte shared-memory broadcast per
// four FMAs (1.2 ms for the 622 592 rows of BASELINE configs[4]); here the tile's writes of the saved-for-backward
// tensors (1.8 KB per row) are what remains.  Same outputs, same arg-max rule (first node on ties).
// Two GROUPS of 8 warps per CTA, each walking its own 128-row tiles and synchronising on its own named barrier: the
// groups drift apart, so one group's MMA phase overlaps the other's store-heavy epilogue (one 16-warp group in lockstep
// measured 403 us at configs[4] with the DRAM pipe at 35 %).
constexpr int SGM_THREADS = 512, SGM_GROUP = 256, SGM_TILE = 128, SGM_LD = 68, SGM_LD0 = 12;
__device__ __forceinline__ void sg_group_sync(int grp) { asm volatile("bar.sync %0, %1;" :: "r"(grp + 1), "n"(SGM_GROUP) : "memory"); }

__device__ __forceinline__ uint32_t sg_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void sg_mma(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int sgm_group_words() { return SGM_TILE * SGM_LD + 2 * 16 * SGM_LD + SGM_TILE * SGM_LD0; }   // Hs, Ms, MBs, Vs of one group
constexpr int sgm_smem_words() { return 64 * SGM_LD0 + 4 * 64 * SGM_LD + 9 * HID + 2 * sgm_group_words(); }

template <int V>
__global__ void __launch_bounds__(SGM_THREADS, 1)
subgraph_mma_fwd_kernel(SubgraphParams p) {
  constexpr int PP = SGM_TILE / V;              // polylines per tile (14 / 6)
  static_assert(PP * V <= SGM_TILE && PP <= 16, "tile geometry");   // 126 / 114 node rows per tile, one m-tile of polylines
  constexpr int P = V + 1;
  extern __shared__ __align__(16) float smf[];
  uint32_t* W0s = reinterpret_cast<uint32_t*>(smf);          // [64][12]  layer 0 weights, k padded to 8
  uint32_t* Wab = W0s + 64 * SGM_LD0;                        // [4][64][68]: W1[:, :64], W1[:, 64:], W2[:, :64], W2[:, 64:]
  float* par = reinterpret_cast<float*>(Wab + 4 * 64 * SGM_LD);   // [3][3][64] bias, gamma, beta
  const int grp = threadIdx.x / SGM_GROUP, gtid = threadIdx.x - grp * SGM_GROUP;
  float* Hs = par + 9 * HID + grp * sgm_group_words();       // [128][68] activations of the group's tile
  float* Ms = Hs + SGM_TILE * SGM_LD;                        // [16][68]  pooled maxima per polyline (rows >= PP stay zero)
  float* MBs = Ms + 16 * SGM_LD;                             // [16][68]  W[:, 64:] m per polyline
  uint32_t* Vs = reinterpret_cast<uint32_t*>(MBs + 16 * SGM_LD);  // [128][12] polyline vectors (TF32), k padded to 8
  for (int i = threadIdx.x; i < 64 * 8; i += SGM_THREADS) { const int n = i >> 3, k = i & 7; W0s[n * SGM_LD0 + k] = k < 7 ? sg_tf32(p.w[0][n * 7 + k]) : 0u; }
  for (int i = threadIdx.x; i < 2 * 64 * 128; i += SGM_THREADS) {
    const int m = i >> 13, r = i & 8191, n = r >> 7, k = r & 127;
    Wab[((m * 2 + (k >> 6)) * 64 + n) * SGM_LD + (k & 63)] = sg_tf32(p.w[m + 1][r]);
  }
  for (int i = threadIdx.x; i < 3 * HID; i += SGM_THREADS) {
    const int l = i / HID, c = i - l * HID;
    par[(l * 3 + 0) * HID + c] = p.b[l][c];
    par[(l * 3 + 1) * HID + c] = p.gamma[l][c];
    par[(l * 3 + 2) * HID + c] = p.beta[l][c];
  }
  for (int i = gtid; i < 2 * 16 * SGM_LD; i += SGM_GROUP) Ms[i] = 0.f;                 // Ms and MBs of this group
  __syncthreads();                                            // weights, parameters staged; from here on the groups run apart
  const int warp = gtid >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int row0 = warp * 16 + g, row1 = row0 + 8;
  const int pl0 = min(row0 / V, PP - 1), pl1 = min(row1 / V, PP - 1);
  const long long ntiles = (p.G + PP - 1) / PP;
  for (long long tile = (long long)blockIdx.x * 2 + grp; tile < ntiles; tile += (long long)gridDim.x * 2) {
    const long long g0 = tile * PP;
    const long long grow0 = g0 * V;                           // global row of tile row 0 (polylines are consecutive)
    const int npl = (int)min((long long)PP, p.G - g0);        // valid polylines of this tile
    const int nrows = npl * V;
    const bool ok0 = row0 < nrows, ok1 = row1 < nrows;
    sg_group_sync(grp);                                       // the previous tile is done with Vs / Hs / Ms
    // ---- polyline vectorisation (model_rad.py:369-382): node v and v + 1 of the lane
    if (gtid < SGM_TILE) {
      const int r = gtid;
      uint32_t* vr = Vs + r * SGM_LD0;
      if (r < nrows) {
        const int pl = r / V, v = r - pl * V;
        const float* a = p.lane + ((g0 + pl) * P + v) * 5;
        const float xin[7] = {a[0], a[1], a[5], a[6], a[7], a[8], a[9]};
        float* vo = p.vec + (grow0 + r) * 7;
#pragma unroll
        for (int k = 0; k < 7; ++k) { vo[k] = xin[k]; vr[k] = sg_tf32(xin[k]); }
      } else {
#pragma unroll
        for (int k = 0; k < 7; ++k) vr[k] = 0u;
      }
      vr[7] = 0u;
    }
    sg_group_sync(grp);
#pragma unroll 1
    for (int layer = 0; layer < 3; ++layer) {
      const float* bias = par + (layer * 3 + 0) * HID;
      const float* gam = par + (layer * 3 + 1) * HID;
      const float* bet = par + (layer * 3 + 2) * HID;
      float acc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c = nt * 8 + 2 * t;
        float b0 = bias[c], b1 = bias[c + 1];
        acc[nt][0] = b0; acc[nt][1] = b1; acc[nt][2] = b0; acc[nt][3] = b1;
        if (layer > 0) {                                      // + W[:, 64:] m of the row's polyline
          acc[nt][0] += MBs[pl0 * SGM_LD + c]; acc[nt][1] += MBs[pl0 * SGM_LD + c + 1];
          acc[nt][2] += MBs[pl1 * SGM_LD + c]; acc[nt][3] += MBs[pl1 * SGM_LD + c + 1];
        }
      }
      if (layer == 0) {
        uint32_t a[4] = {Vs[row0 * SGM_LD0 + t], Vs[row1 * SGM_LD0 + t], Vs[row0 * SGM_LD0 + t + 4], Vs[row1 * SGM_LD0 + t + 4]};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) sg_mma(acc[nt], a, W0s[(nt * 8 + g) * SGM_LD0 + t], W0s[(nt * 8 + g) * SGM_LD0 + t + 4]);
      } else {
        const uint32_t* Wa = Wab + (layer - 1) * 2 * 64 * SGM_LD;
#pragma unroll 2
        for (int ks = 0; ks < 8; ++ks) {
          const int k = ks * 8 + t;
          uint32_t a[4] = {sg_tf32(Hs[row0 * SGM_LD + k]), sg_tf32(Hs[row1 * SGM_LD + k]),
                           sg_tf32(Hs[row0 * SGM_LD + k + 4]), sg_tf32(Hs[row1 * SGM_LD + k + 4])};
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) sg_mma(acc[nt], a, Wa[(nt * 8 + g) * SGM_LD + k], Wa[(nt * 8 + g) * SGM_LD + k + 4]);
        }
      }
      // ---- epilogue: save y, LayerNorm over the quad's 64 channels (two-pass variance, as ln_fwd_kernel), ReLU,
      //      h -> shared memory (next layer's operand, max-pool) and -> x_{layer+1}[:, :64]
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = h ? row1 : row0;
        const bool ok = h ? ok1 : ok0;
        float s = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) s += acc[nt][2 * h] + acc[nt][2 * h + 1];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const float mu = s / (float)HID;
        float q = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { const float d0 = acc[nt][2 * h] - mu, d1 = acc[nt][2 * h + 1] - mu; q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); }
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        const float rs = rsqrtf(q / (float)HID + p.eps);
        const long long gr = grow0 + row;
        if (ok && t == 0) { p.mean[layer][gr] = mu; p.rstd[layer][gr] = rs; }
        float* yo = p.y[layer] + gr * HID;
        float* xo = (layer == 0 ? p.x1 : p.x2) + gr * 2 * HID;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int c = nt * 8 + 2 * t;
          const float y0 = acc[nt][2 * h], y1 = acc[nt][2 * h + 1];
          const float h0 = fmaxf(fmaf((y0 - mu) * rs, gam[c], bet[c]), 0.f), h1 = fmaxf(fmaf((y1 - mu) * rs, gam[c + 1], bet[c + 1]), 0.f);
          *reinterpret_cast<float2*>(Hs + row * SGM_LD + c) = make_float2(h0, h1);
          if (ok) {
            *reinterpret_cast<float2*>(yo + c) = make_float2(y0, y1);
            if (layer < 2) *reinterpret_cast<float2*>(xo + c) = make_float2(h0, h1);
          }
        }
      }
      sg_group_sync(grp);
      // ---- segment max-pool over each polyline's nodes (first node on ties, NaN propagates like the unfused kernels)
      for (int i = gtid; i < PP * HID; i += SGM_GROUP) {
        const int pl = i >> 6, c = i & 63;
        const float* hr = Hs + pl * V * SGM_LD + c;
        float best = hr[0];
        int bi = 0;
#pragma unroll
        for (int v = 1; v < V; ++v) { const float tv = hr[v * SGM_LD]; if (tv > best || tv != tv) { best = tv; bi = v; } }
        Ms[pl * SGM_LD + c] = best;
        if (pl < npl) {
          const long long gg = g0 + pl;
          p.arg[layer][gg * HID + c] = bi;
          if (layer == 2) {                                   // tok = [m_2 | m_2]; arg-max of the broadcast half is node 0
            p.tok[gg * 2 * HID + c] = best; p.tok[gg * 2 * HID + HID + c] = best;
            p.argf[gg * 2 * HID + c] = bi; p.argf[gg * 2 * HID + HID + c] = 0;
          }
        }
      }
      sg_group_sync(grp);
      if (layer == 2) break;
      // ---- x_{layer+1}[:, 64:] = m of the row's polyline
      {
        float* xo = (layer == 0 ? p.x1 : p.x2) + grow0 * 2 * HID;
        for (int i = gtid; i < nrows * 16; i += SGM_GROUP) {
          const int r = i >> 4, q4 = i & 15;
          *reinterpret_cast<float4*>(xo + (long long)r * 2 * HID + HID + q4 * 4) = *reinterpret_cast<const float4*>(Ms + (r / V) * SGM_LD + q4 * 4);
        }
      }
      // ---- per-polyline product with the broadcast half of the next layer's weights: MBs = Ms W[:, 64:]^T
      {
        const int nt = warp, mt = 0;
        const uint32_t* Wb = Wab + (layer * 2 + 1) * 64 * SGM_LD;
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        const float* m0 = Ms + (mt * 16 + g) * SGM_LD;
        const float* m1 = m0 + 8 * SGM_LD;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const int k = ks * 8 + t;
          uint32_t a[4] = {sg_tf32(m0[k]), sg_tf32(m1[k]), sg_tf32(m0[k + 4]), sg_tf32(m1[k + 4])};
          sg_mma(d, a, Wb[(nt * 8 + g) * SGM_LD + k], Wb[(nt * 8 + g) * SGM_LD + k + 4]);
        }
        const int c = nt * 8 + 2 * t;
        *reinterpret_cast<float2*>(MBs + (mt * 16 + g) * SGM_LD + c) = make_float2(d[0], d[1]);
        *reinterpret_cast<float2*>(MBs + (mt * 16 + g + 8) * SGM_LD + c) = make_float2(d[2], d[3]);
      }
      sg_group_sync(grp);
    }
  }
}

template <int V>
int launch_subgraph_mma(const SubgraphParams& p, cudaStream_t stream) {
  constexpr int PP = SGM_TILE / V;
  const int smem = sgm_smem_words() * (int)sizeof(float);
  cudaError_t ce = cudaFuncSetAttribute(subgraph_mma_fwd_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (ce != cudaSuccess) { mmfn_set_error("subgraph_fused_fwd: smem attribute: %s", cudaGetErrorString(ce)); return (int)ce; }
  const long long ntiles = (p.G + PP - 1) / PP, need = (ntiles + 1) / 2;     // two tile-walking groups per CTA
  const int grid = (int)(need < 148 ? need : 148);
  subgraph_mma_fwd_kernel<V><<<grid, SGM_THREADS, smem, stream>>>(p);
  return mmfn_launch_status("subgraph_fused_fwd");
}

}  // namespace

// lane: (G, V + 1, 5) polyline nodes; w0 (64,7), w1 / w2 (64,128), biases and LayerNorm parameters (64) of the three
// Subgraph layers (reference keys lane_subgraph.layers.mlp_{0,1,2}.mlp.{0,1}).  Outputs (caller-allocated): vec (G*V,7);
// y0..y2 (G*V,64) pre-LayerNorm linear outputs; mean*/rstd* (G*V); x1, x2 (G*V,128) inputs of layers 1 / 2; arg0..arg2
// (G,64) node index of each pooled maximum; tok (G,128) polyline tokens; argf (G,128) arg-max of the final pool.
// V in {9, 19} (10- / 20-node lanes): other lane lengths take the unfused kernels.  tf32 != 0: the linears multiply on the
// tensor cores (mma.sync TF32, fp32 accumulate) instead of exact fp32 FMAs.
MMFN_API int mmfn_subgraph_fused_fwd(const float* lane, int64_t G, int V, int tf32,
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
  if (tf32) return V == 9 ? launch_subgraph_mma<9>(p, stream) : launch_subgraph_mma<19>(p, stream);
  if (V == 9) return launch_subgraph<9>(p, stream);
  return launch_subgraph<19>(p, stream);
}
