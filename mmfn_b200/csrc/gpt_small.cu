// Whole-GPT forward for the two narrow fusion transformers (reference model_rad.py:112-133 Block, :211-247 GPT.forward
// body; n_embd 64 / 128, 4 heads, T = 128 / 192 tokens) in ONE launch.
//
// Why: at these widths a transformer block is 7 dependent kernels of a few MFLOP each; every one of them costs its fixed
// latency (launch, barrier / TMEM set-up, first TMA round trip, epilogue drain: 5-10 us measured) whatever its size, and
// 8 layers x 7 kernels put ~0.5 ms on the critical path of the step for 0.02 ms worth of arithmetic
// (profiles/r02_ablate_substeps_*.json).  Here a sample's tokens never leave the SM between layers:
//   * one thread-block CLUSTER per sample, one CTA per 64-token slab; the residual stream (fp32), the LayerNorm output,
//     Q, K and V^T of the slab live in shared memory for all layers;
//   * keys / values of the sample's other slabs are read straight from the peer CTAs' shared memory (DSMEM) -- two
//     cluster barriers per layer are the only cross-CTA synchronisation;
//   * every linear is a sequence of [64 x C] x [C x C] blocks (3 for QKV, 1 projection, 4 + 4 for the MLP whose hidden
//     activations are consumed chunk by chunk, never materialised on chip as a whole); weight tiles stream from L2
//     through a cp.async double buffer that keeps prefetching across the attention and LayerNorm phases;
//   * softmax runs on the score fragments in registers; the probabilities feed the PV product without leaving them.
// Tensor cores: warp-level mma.sync (bf16 m16n8k16 / tf32 m16n8k8, fp32 accumulate).  The tcgen05 path (128-row tiles,
// accumulators in TMEM, TMA-fed) is the right tool for the big GEMMs of this model (gemm_tc.cu / conv_tc.cu); here the
// tiles are 16 x 64 per warp, operands are produced by the previous stage in the same CTA, and a TMEM round trip per
// stage (7 per layer) would put back the latency this kernel removes.
// Everything the backward needs is written out as the layers go (the same tensors the per-op path saves):
// layer inputs, LayerNorm statistics and outputs, qkv, P / dropout(P), attention output, MLP hidden.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int GS_THREADS = 256;
constexpr int GS_ROWS = 64;                    // token rows per CTA
constexpr int GS_MAXL = 12;

struct GptFwdParams {
  const float* x0;                             // (M, C) tokens entering block 0
  float* xout;                                 // [L][M][C] block outputs (block l+1's input)
  float* x1;                                   // [L][M][C] x + drop(proj(attention))
  void *h1, *qkv, *y, *h2, *a;                 // [L][M][C | 3C | C | C | 4C] operand-typed (bf16 / fp32)
  void *P, *Pd;                                // [L][B][nh][T][T]; Pd null without attention dropout
  float *mean1, *rstd1, *mean2, *rstd2;        // [L][M]
  const void* w[GS_MAXL][4];                   // qkv (3C,C) [key|query|value], proj (C,C), fc1 (4C,C), fc2 (C,4C)
  const float* bias[GS_MAXL][4];
  const float* ln[GS_MAXL][4];                 // ln1 gamma, beta, ln2 gamma, beta
  int B, T, L;
  float attn_p, resid_p, eps;
  unsigned long long seed;                     // block l: attention seed + 3l + 1, projection + 3l + 2, MLP + 3l + 3
  unsigned long long* trace;                   // optional: %globaltimer stamps of CTA (0,0), 10 per block (tools/gpt_bench.py)
};

__device__ __forceinline__ void gs_stamp(unsigned long long* trace, int i) {
  if (trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[i] = t;
  }
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Operand precision.  In units of 32-bit words both have the same fragment addressing: one MMA k-step spans 8 words of
// a row; a thread (g = lane / 4, t = lane % 4) reads words t and t + 4 of rows g (and g + 8 for A).
struct PrecBF {
  static constexpr int EPW = 2;                // elements per word
  using elem = __nv_bfloat16;
  __device__ static __forceinline__ void mma(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  // two adjacent columns (col even) of one row
  __device__ static __forceinline__ void st_smem(uint32_t* base, int ld, int row, int col, float v0, float v1) {
    base[row * ld + (col >> 1)] = pack_bf16(v0, v1);
  }
  __device__ static __forceinline__ void st_smem_t(uint32_t* base, int ld, int row, int col, float v) {   // transposed element
    reinterpret_cast<__nv_bfloat16*>(base + row * ld)[col] = __float2bfloat16_rn(v);
  }
  __device__ static __forceinline__ void st_global(void* base, long long idx, float v0, float v1) {
    *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = pack_bf16(v0, v1);
  }
};
struct PrecTF {
  static constexpr int EPW = 1;
  using elem = float;
  __device__ static __forceinline__ void mma(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  __device__ static __forceinline__ void st_smem(uint32_t* base, int ld, int row, int col, float v0, float v1) {
    *reinterpret_cast<float2*>(base + row * ld + col) = make_float2(v0, v1);
  }
  __device__ static __forceinline__ void st_smem_t(uint32_t* base, int ld, int row, int col, float v) {
    base[row * ld + col] = __float_as_uint(v);
  }
  __device__ static __forceinline__ void st_global(void* base, long long idx, float v0, float v1) {
    *reinterpret_cast<float2*>(reinterpret_cast<float*>(base) + idx) = make_float2(v0, v1);
  }
};

template <int C, class Pr>
struct Lay {
  static constexpr int EPW = Pr::EPW;
  static constexpr int XS_LD = C + 8;                    // fp32 residual stream, words per row (float2 fragment stores)
  static constexpr int A_LD = C / EPW + 4;               // MMA operand rows (word stride = 4 mod 32: conflict-free)
  static constexpr int V_LD = GS_ROWS / EPW + (EPW == 1 ? 8 : 4);   // V^T rows: 64 keys of the slab
  static constexpr int W_LD = 36;                        // weight tile rows: 32 words = one 128-byte line of k
  static constexpr int KT = 32 * EPW;                    // elements of k per staged weight tile
  static constexpr int TILES = C / KT;                   // tiles per C x C block
  static constexpr int NT = C / 16;                      // 8-column MMA tiles per warp (a warp owns C / 2 columns)
  static constexpr int XS = 0;
  static constexpr int HS = XS + GS_ROWS * XS_LD;
  static constexpr int QS = HS + GS_ROWS * A_LD;
  static constexpr int KS = QS + GS_ROWS * A_LD;
  static constexpr int VT = KS + GS_ROWS * A_LD;
  static constexpr int WS = VT + C * V_LD;
  // weight-tile ring: the stream is latency-bound (one tile = 8-18 KB, L2 round trip ~1.5 us), so as many tiles as
  // shared memory allows are kept in flight
  static constexpr int STAGES = (C == 64) ? 8 : (EPW == 2 ? 5 : 3);
  static constexpr int WTILE = C * W_LD;
  static constexpr int WORDS = WS + STAGES * WTILE;
};

// Weight-tile stream: tile index ti -> (layer, block, k-tile); block 0..2 key / query / value, 3 projection,
// 4 + 2c fc1 chunk c, 5 + 2c fc2 chunk c.
template <int C, class Pr>
__device__ __forceinline__ void load_tile(const GptFwdParams& p, uint32_t* wbuf, int ti) {
  using L = Lay<C, Pr>;
  using E = typename Pr::elem;
  const int kt = ti % L::TILES, bl = ti / L::TILES, blk = bl % 12, layer = bl / 12;
  const E* W;
  int n0 = 0, k0 = 0, ldw = C;
  if (blk < 3) { W = (const E*)p.w[layer][0]; n0 = blk * C; }
  else if (blk == 3) { W = (const E*)p.w[layer][1]; }
  else if ((blk & 1) == 0) { W = (const E*)p.w[layer][2]; n0 = ((blk - 4) >> 1) * C; }
  else { W = (const E*)p.w[layer][3]; k0 = ((blk - 5) >> 1) * C; ldw = 4 * C; }
  const E* src = W + (long long)n0 * ldw + k0 + kt * L::KT;
  for (int i = threadIdx.x; i < C * 8; i += GS_THREADS) {
    const int n = i >> 3, ch = i & 7;
    cp_async16(wbuf + n * L::W_LD + ch * 4, reinterpret_cast<const char*>(src + (long long)n * ldw) + ch * 16);
  }
}

// acc[NT][4] += A[64 x C] (shared, operand-typed) . Wblk[C x C]^T for this warp's 16 rows x C/2 columns.
// `ti` is the stream index of the block's first tile; tiles ti .. ti + STAGES - 2 are already in flight.
template <int C, class Pr>
__device__ __forceinline__ void gemm_block(const GptFwdParams& p, uint32_t* smem, const uint32_t* A, float (*acc)[4], int& ti) {
  using L = Lay<C, Pr>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int rb = warp & 3, chalf = warp >> 2;
  const int total = p.L * 12 * L::TILES;
  for (int kt = 0; kt < L::TILES; ++kt, ++ti) {
    cp_async_wait<L::STAGES - 2>();                      // this thread's part of tile ti has landed
    __syncthreads();                                     // ... everyone's; and everyone is done with tile ti - 1's slot
    if (ti + L::STAGES - 1 < total) load_tile<C, Pr>(p, smem + L::WS + ((ti + L::STAGES - 1) % L::STAGES) * L::WTILE, ti + L::STAGES - 1);
    cp_async_commit();                                   // (an empty group keeps the group count uniform)
    const uint32_t* Wt = smem + L::WS + (ti % L::STAGES) * L::WTILE + (chalf * (C / 2) + g) * L::W_LD + t;
    const uint32_t* Ar = A + (rb * 16 + g) * L::A_LD + kt * 32 + t;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[4];
      a[0] = Ar[ks * 8]; a[1] = Ar[8 * L::A_LD + ks * 8]; a[2] = Ar[ks * 8 + 4]; a[3] = Ar[8 * L::A_LD + ks * 8 + 4];
#pragma unroll
      for (int j = 0; j < L::NT; ++j) {
        const uint32_t b0 = Wt[j * 8 * L::W_LD + ks * 8], b1 = Wt[j * 8 * L::W_LD + ks * 8 + 4];
        Pr::mma(acc[j], a, b0, b1);
      }
    }
  }
}

// LayerNorm of the slab's 64 rows: xs (fp32) -> hs (operand-typed) + global copy + row statistics.
// Two-pass variance, rsqrtf(var + eps): the arithmetic of norm.cu's ln_fwd_kernel.
template <int C, class Pr>
__device__ __forceinline__ void layernorm_slab(const float* xs, uint32_t* hs, const float* gamma, const float* beta,
                                               void* hout, float* mean, float* rstd, long long row0, float eps) {
  using L = Lay<C, Pr>;
  constexpr int CPL = C / 32;                           // columns per lane: 2 or 4
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float gm[CPL], bt[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) { gm[i] = gamma[lane * CPL + i]; bt[i] = beta[lane * CPL + i]; }
#pragma unroll 2
  for (int rr = 0; rr < 8; ++rr) {
    const int row = warp * 8 + rr;
    float v[CPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) { v[i] = xs[row * L::XS_LD + lane * CPL + i]; s += v[i]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) { v[i] -= mu; q = fmaf(v[i], v[i], q); }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = rsqrtf(q / (float)C + eps);
#pragma unroll
    for (int i = 0; i < CPL; ++i) v[i] = fmaf(v[i] * rs, gm[i], bt[i]);
#pragma unroll
    for (int i = 0; i < CPL; i += 2) {
      Pr::st_smem(hs, L::A_LD, row, lane * CPL + i, v[i], v[i + 1]);
      Pr::st_global(hout, (row0 + row) * C + lane * CPL + i, v[i], v[i + 1]);
    }
    if (lane == 0) { mean[row0 + row] = mu; rstd[row0 + row] = rs; }
  }
}

template <int C, int NS, class Pr>
__global__ void __launch_bounds__(GS_THREADS, 1) gpt_small_fwd_kernel(const __grid_constant__ GptFwdParams p) {
  using L = Lay<C, Pr>;
  constexpr int NH = 4, HS = C / NH, EPW = Pr::EPW;
  constexpr int HW = HS / EPW;                          // words per head row
  constexpr int QK_STEPS = HW / 8;                      // MMA k-steps of one score
  constexpr int NT = L::NT;
  extern __shared__ __align__(16) uint32_t smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int slab = (int)cluster.block_rank();           // == blockIdx.x
  const int b = blockIdx.y;
  const int T = NS * GS_ROWS;
  const long long M = (long long)p.B * T;
  const long long row0 = (long long)b * T + slab * GS_ROWS;        // first global token row of this slab
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int rb = warp & 3, chalf = warp >> 2;
  float* xs = reinterpret_cast<float*>(smem + L::XS);
  uint32_t* hs = smem + L::HS;
  uint32_t* qs = smem + L::QS;
  uint32_t* ks = smem + L::KS;
  uint32_t* vt = smem + L::VT;
  const uint32_t* ks_of[NS];
  const uint32_t* vt_of[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) { ks_of[s] = cluster.map_shared_rank(ks, s); vt_of[s] = cluster.map_shared_rank(vt, s); }

  int ti = 0;                                            // weight-tile stream position
  {
    const int total = p.L * 12 * L::TILES;
#pragma unroll 1
    for (int s = 0; s < L::STAGES - 1; ++s) {
      if (s < total) load_tile<C, Pr>(p, smem + L::WS + s * L::WTILE, s);
      cp_async_commit();
    }
  }
  // residual stream of the slab
  for (int i = threadIdx.x; i < GS_ROWS * (C / 4); i += GS_THREADS) {
    const int r = i / (C / 4), c4 = i - r * (C / 4);
    *reinterpret_cast<float4*>(xs + r * L::XS_LD + c4 * 4) = *reinterpret_cast<const float4*>(p.x0 + (row0 + r) * C + c4 * 4);
  }
  __syncthreads();

  const float keep_r = 1.0f / (1.0f - p.resid_p), keep_a = 1.0f / (1.0f - p.attn_p);
  const uint32_t thr_r = mmfn_drop_threshold(p.resid_p), thr_a = mmfn_drop_threshold(p.attn_p);
  const float sc_log2 = rsqrtf((float)HS) * 1.4426950408889634f;
  const int r0 = rb * 16 + g;                            // this thread's fragment rows: r0, r0 + 8

  for (int layer = 0; layer < p.L; ++layer) {
    const long long lM = (long long)layer * M;
    const uint64_t seed_a = mmfn_drop_seed(p.seed + 3 * layer + 1), seed_p = mmfn_drop_seed(p.seed + 3 * layer + 2),
                   seed_m = mmfn_drop_seed(p.seed + 3 * layer + 3);
    gs_stamp(p.trace, layer * 10 + 0);
    // ---- ln1
    layernorm_slab<C, Pr>(xs, hs, p.ln[layer][0], p.ln[layer][1], reinterpret_cast<typename Pr::elem*>(p.h1) + lM * C,
                          p.mean1 + lM, p.rstd1 + lM, row0, p.eps);
    gs_stamp(p.trace, layer * 10 + 1);
    // ---- key / query / value
    typename Pr::elem* qkv_g = reinterpret_cast<typename Pr::elem*>(p.qkv) + lM * 3 * C;
#pragma unroll 1
    for (int part = 0; part < 3; ++part) {
      float acc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      gemm_block<C, Pr>(p, smem, hs, acc, ti);
      const float* bias = p.bias[layer][0] + part * C;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int col = chalf * (C / 2) + j * 8 + 2 * t;
        const float b0 = bias[col], b1 = bias[col + 1];
        const float v00 = acc[j][0] + b0, v01 = acc[j][1] + b1, v10 = acc[j][2] + b0, v11 = acc[j][3] + b1;
        Pr::st_global(qkv_g, (row0 + r0) * 3 * C + part * C + col, v00, v01);
        Pr::st_global(qkv_g, (row0 + r0 + 8) * 3 * C + part * C + col, v10, v11);
        if (part < 2) {
          uint32_t* dst = part == 0 ? ks : qs;
          Pr::st_smem(dst, L::A_LD, r0, col, v00, v01);
          Pr::st_smem(dst, L::A_LD, r0 + 8, col, v10, v11);
        } else {                                          // V^T: [dim][key of the slab]
          Pr::st_smem_t(vt, L::V_LD, col, r0, v00); Pr::st_smem_t(vt, L::V_LD, col + 1, r0, v01);
          Pr::st_smem_t(vt, L::V_LD, col, r0 + 8, v10); Pr::st_smem_t(vt, L::V_LD, col + 1, r0 + 8, v11);
        }
      }
    }
    gs_stamp(p.trace, layer * 10 + 2);
    cluster.sync();                                       // K, V^T of every slab of the sample are in place
    gs_stamp(p.trace, layer * 10 + 3);
    // ---- attention: warp = (16 query rows, 2 heads); scores, softmax and P V in registers
#pragma unroll 1
    for (int u = 0; u < 2; ++u) {
      const int h = chalf * 2 + u;
      uint32_t qf[QK_STEPS][4];
      {
        const uint32_t* qr = qs + r0 * L::A_LD + h * HW + t;
#pragma unroll
        for (int s = 0; s < QK_STEPS; ++s) {
          qf[s][0] = qr[s * 8]; qf[s][1] = qr[8 * L::A_LD + s * 8]; qf[s][2] = qr[s * 8 + 4]; qf[s][3] = qr[8 * L::A_LD + s * 8 + 4];
        }
      }
      float sa[NS * 8][4];
#pragma unroll
      for (int j = 0; j < NS * 8; ++j) sa[j][0] = sa[j][1] = sa[j][2] = sa[j][3] = 0.f;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const uint32_t* kr = ks_of[s] + g * L::A_LD + h * HW + t;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int st = 0; st < QK_STEPS; ++st)
            Pr::mma(sa[s * 8 + j], qf[st], kr[j * 8 * L::A_LD + st * 8], kr[j * 8 * L::A_LD + st * 8 + 4]);
        }
      }
      float m0 = -3.0e38f, m1 = -3.0e38f;
#pragma unroll
      for (int j = 0; j < NS * 8; ++j) { m0 = fmaxf(m0, fmaxf(sa[j][0], sa[j][1])); m1 = fmaxf(m1, fmaxf(sa[j][2], sa[j][3])); }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      const float o0 = m0 * sc_log2, o1 = m1 * sc_log2;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int j = 0; j < NS * 8; ++j) {
        sa[j][0] = ex2f(fmaf(sa[j][0], sc_log2, -o0)); sa[j][1] = ex2f(fmaf(sa[j][1], sc_log2, -o0));
        sa[j][2] = ex2f(fmaf(sa[j][2], sc_log2, -o1)); sa[j][3] = ex2f(fmaf(sa[j][3], sc_log2, -o1));
        s0 += sa[j][0] + sa[j][1]; s1 += sa[j][2] + sa[j][3];
      }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      const float i0 = 1.0f / s0, i1 = 1.0f / s1;
      // P[b, h, i, :] (and dropout(P)) to HBM for the backward; index = dropout hash key, as attn_bf16.cu
      const long long pl = (long long)layer * p.B * NH * T * T;
      const long long prow0 = pl + (((long long)b * NH + h) * T + slab * GS_ROWS + r0) * T, prow1 = prow0 + 8LL * T;
      const bool drop = p.attn_p > 0.f;
#pragma unroll
      for (int j = 0; j < NS * 8; ++j) {
        const int col = j * 8 + 2 * t;
        sa[j][0] *= i0; sa[j][1] *= i0; sa[j][2] *= i1; sa[j][3] *= i1;
        Pr::st_global(p.P, prow0 + col, sa[j][0], sa[j][1]);
        Pr::st_global(p.P, prow1 + col, sa[j][2], sa[j][3]);
        if (drop) {
          const int sh = (col & 2) * 16;                  // 16-bit field of element col in its group of four
          const uint64_t h0 = mmfn_hash64(seed_a, (uint64_t)((prow0 - pl) + col) >> 2) >> sh;
          const uint64_t h1 = mmfn_hash64(seed_a, (uint64_t)((prow1 - pl) + col) >> 2) >> sh;
          sa[j][0] = ((uint32_t)h0 & 0xFFFFu) >= thr_a ? sa[j][0] * keep_a : 0.f;
          sa[j][1] = ((uint32_t)(h0 >> 16) & 0xFFFFu) >= thr_a ? sa[j][1] * keep_a : 0.f;
          sa[j][2] = ((uint32_t)h1 & 0xFFFFu) >= thr_a ? sa[j][2] * keep_a : 0.f;
          sa[j][3] = ((uint32_t)(h1 >> 16) & 0xFFFFu) >= thr_a ? sa[j][3] * keep_a : 0.f;
          Pr::st_global(p.Pd, prow0 + col, sa[j][0], sa[j][1]);
          Pr::st_global(p.Pd, prow1 + col, sa[j][2], sa[j][3]);
        }
      }
      // y = dropout(P) V: the score fragments ARE the A operand
      float oa[HS / 8][4];
#pragma unroll
      for (int j = 0; j < HS / 8; ++j) oa[j][0] = oa[j][1] = oa[j][2] = oa[j][3] = 0.f;
      if constexpr (EPW == 2) {
#pragma unroll
        for (int kk = 0; kk < NS * 4; ++kk) {             // 16 keys per step = two score tiles
          uint32_t a[4];
          a[0] = pack_bf16(sa[2 * kk][0], sa[2 * kk][1]); a[1] = pack_bf16(sa[2 * kk][2], sa[2 * kk][3]);
          a[2] = pack_bf16(sa[2 * kk + 1][0], sa[2 * kk + 1][1]); a[3] = pack_bf16(sa[2 * kk + 1][2], sa[2 * kk + 1][3]);
          const uint32_t* vr = vt_of[kk >> 2] + (h * HS + g) * L::V_LD + (kk & 3) * 8 + t;
#pragma unroll
          for (int jn = 0; jn < HS / 8; ++jn) Pr::mma(oa[jn], a, vr[jn * 8 * L::V_LD], vr[jn * 8 * L::V_LD + 4]);
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < NS * 8; ++kk) {             // 8 keys per step; MMA k index t <-> key 2t, t + 4 <-> key 2t + 1
          uint32_t a[4];
          a[0] = __float_as_uint(sa[kk][0]); a[1] = __float_as_uint(sa[kk][2]);
          a[2] = __float_as_uint(sa[kk][1]); a[3] = __float_as_uint(sa[kk][3]);
          const uint32_t* vr = vt_of[kk >> 3] + (h * HS + g) * L::V_LD + (kk & 7) * 8 + 2 * t;
#pragma unroll
          for (int jn = 0; jn < HS / 8; ++jn) {
            const uint2 bb = *reinterpret_cast<const uint2*>(vr + jn * 8 * L::V_LD);
            Pr::mma(oa[jn], a, bb.x, bb.y);
          }
        }
      }
      typename Pr::elem* y_g = reinterpret_cast<typename Pr::elem*>(p.y) + lM * C;
#pragma unroll
      for (int jn = 0; jn < HS / 8; ++jn) {
        const int col = h * HS + jn * 8 + 2 * t;
        Pr::st_smem(hs, L::A_LD, r0, col, oa[jn][0], oa[jn][1]);
        Pr::st_smem(hs, L::A_LD, r0 + 8, col, oa[jn][2], oa[jn][3]);
        Pr::st_global(y_g, (row0 + r0) * C + col, oa[jn][0], oa[jn][1]);
        Pr::st_global(y_g, (row0 + r0 + 8) * C + col, oa[jn][2], oa[jn][3]);
      }
    }
    gs_stamp(p.trace, layer * 10 + 4);
    cluster.sync();                                       // peers are done with this slab's K / V^T
    gs_stamp(p.trace, layer * 10 + 5);
    // ---- projection + residual
    {
      float acc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      gemm_block<C, Pr>(p, smem, hs, acc, ti);
      const float* bias = p.bias[layer][1];
      float* x1_g = p.x1 + lM * C;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int col = chalf * (C / 2) + j * 8 + 2 * t;
        const float b0 = bias[col], b1 = bias[col + 1];
        float v[4] = {acc[j][0] + b0, acc[j][1] + b1, acc[j][2] + b0, acc[j][3] + b1};
        if (p.resid_p > 0.f) {
          const int sh = (col & 2) * 16;
          const uint64_t h0 = mmfn_hash64(seed_p, (uint64_t)((row0 + r0) * C + col) >> 2) >> sh;
          const uint64_t h1 = mmfn_hash64(seed_p, (uint64_t)((row0 + r0 + 8) * C + col) >> 2) >> sh;
          v[0] = ((uint32_t)h0 & 0xFFFFu) >= thr_r ? v[0] * keep_r : 0.f;
          v[1] = ((uint32_t)(h0 >> 16) & 0xFFFFu) >= thr_r ? v[1] * keep_r : 0.f;
          v[2] = ((uint32_t)h1 & 0xFFFFu) >= thr_r ? v[2] * keep_r : 0.f;
          v[3] = ((uint32_t)(h1 >> 16) & 0xFFFFu) >= thr_r ? v[3] * keep_r : 0.f;
        }
        float2* xa = reinterpret_cast<float2*>(xs + r0 * L::XS_LD + col);
        float2* xb = reinterpret_cast<float2*>(xs + (r0 + 8) * L::XS_LD + col);
        float2 ua = *xa, ub = *xb;
        ua.x += v[0]; ua.y += v[1]; ub.x += v[2]; ub.y += v[3];
        *xa = ua; *xb = ub;
        *reinterpret_cast<float2*>(x1_g + (row0 + r0) * C + col) = ua;
        *reinterpret_cast<float2*>(x1_g + (row0 + r0 + 8) * C + col) = ub;
      }
    }
    __syncthreads();
    gs_stamp(p.trace, layer * 10 + 6);
    // ---- ln2
    layernorm_slab<C, Pr>(xs, hs, p.ln[layer][2], p.ln[layer][3], reinterpret_cast<typename Pr::elem*>(p.h2) + lM * C,
                          p.mean2 + lM, p.rstd2 + lM, row0, p.eps);
    gs_stamp(p.trace, layer * 10 + 7);
    // ---- MLP: hidden chunk c = ReLU(h2 W1[c]^T + b1[c]) -> shared (Q's buffer) -> acc2 += chunk . W2[:, c]^T
    {
      float acc2[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) acc2[j][0] = acc2[j][1] = acc2[j][2] = acc2[j][3] = 0.f;
      typename Pr::elem* a_g = reinterpret_cast<typename Pr::elem*>(p.a) + lM * 4 * C;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float acc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        gemm_block<C, Pr>(p, smem, hs, acc, ti);
        const float* bias = p.bias[layer][2] + c * C;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const int col = chalf * (C / 2) + j * 8 + 2 * t;
          const float b0 = bias[col], b1 = bias[col + 1];
          const float v00 = fmaxf(acc[j][0] + b0, 0.f), v01 = fmaxf(acc[j][1] + b1, 0.f);
          const float v10 = fmaxf(acc[j][2] + b0, 0.f), v11 = fmaxf(acc[j][3] + b1, 0.f);
          Pr::st_smem(qs, L::A_LD, r0, col, v00, v01);
          Pr::st_smem(qs, L::A_LD, r0 + 8, col, v10, v11);
          Pr::st_global(a_g, (row0 + r0) * 4 * C + c * C + col, v00, v01);
          Pr::st_global(a_g, (row0 + r0 + 8) * 4 * C + c * C + col, v10, v11);
        }
        gemm_block<C, Pr>(p, smem, qs, acc2, ti);   // (its first __syncthreads publishes the chunk)
      }
      const float* bias = p.bias[layer][3];
      float* xo_g = p.xout + lM * C;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int col = chalf * (C / 2) + j * 8 + 2 * t;
        const float b0 = bias[col], b1 = bias[col + 1];
        float v[4] = {acc2[j][0] + b0, acc2[j][1] + b1, acc2[j][2] + b0, acc2[j][3] + b1};
        if (p.resid_p > 0.f) {
          const int sh = (col & 2) * 16;
          const uint64_t h0 = mmfn_hash64(seed_m, (uint64_t)((row0 + r0) * C + col) >> 2) >> sh;
          const uint64_t h1 = mmfn_hash64(seed_m, (uint64_t)((row0 + r0 + 8) * C + col) >> 2) >> sh;
          v[0] = ((uint32_t)h0 & 0xFFFFu) >= thr_r ? v[0] * keep_r : 0.f;
          v[1] = ((uint32_t)(h0 >> 16) & 0xFFFFu) >= thr_r ? v[1] * keep_r : 0.f;
          v[2] = ((uint32_t)h1 & 0xFFFFu) >= thr_r ? v[2] * keep_r : 0.f;
          v[3] = ((uint32_t)(h1 >> 16) & 0xFFFFu) >= thr_r ? v[3] * keep_r : 0.f;
        }
        float2* xa = reinterpret_cast<float2*>(xs + r0 * L::XS_LD + col);
        float2* xb = reinterpret_cast<float2*>(xs + (r0 + 8) * L::XS_LD + col);
        float2 ua = *xa, ub = *xb;
        ua.x += v[0]; ua.y += v[1]; ub.x += v[2]; ub.y += v[3];
        *xa = ua; *xb = ub;
        *reinterpret_cast<float2*>(xo_g + (row0 + r0) * C + col) = ua;
        *reinterpret_cast<float2*>(xo_g + (row0 + r0 + 8) * C + col) = ub;
      }
    }
    __syncthreads();
    gs_stamp(p.trace, layer * 10 + 8);
  }
  cp_async_wait_all();
}

template <int C, int NS, class Pr>
int launch_gpt_fwd(const GptFwdParams& p, cudaStream_t stream) {
  using L = Lay<C, Pr>;
  const int smem = L::WORDS * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(gpt_small_fwd_kernel<C, NS, Pr>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (ce != cudaSuccess) { mmfn_set_error("gpt_small_fwd: shared memory attribute (%d B): %s", smem, cudaGetErrorString(ce)); return (int)ce; }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(NS, p.B, 1);
  cfg.blockDim = dim3(GS_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = NS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t ce = cudaLaunchKernelEx(&cfg, gpt_small_fwd_kernel<C, NS, Pr>, p);
  if (ce != cudaSuccess) { mmfn_set_error("gpt_small_fwd: launch: %s", cudaGetErrorString(ce)); return (int)ce; }
  return mmfn_launch_status("gpt_small_fwd");
}

}  // namespace

MMFN_DEFINE_RNG_BINDER(gpt_small)

static unsigned long long* g_gpt_trace = nullptr;
// Debug aid for tools/gpt_bench.py: device buffer of >= 10 * n_layer uint64 that receives %globaltimer stamps of the
// first CTA at the phase boundaries of every block (null switches tracing off).
MMFN_API int mmfn_gpt_small_trace(void* buf) { g_gpt_trace = (unsigned long long*)buf; return 0; }

// All n_layer pre-LN transformer blocks of one fusion GPT (model_rad.py:112-133 x n_layer, called from :236) in one
// launch, for n_embd C in {64, 128}, 4 heads, T in {128, 192} tokens.  dtype MMFN_BF16: bf16 operands (weights = the
// bf16 shadow, saved operand tensors bf16); MMFN_TF32: fp32 storage multiplied as TF32.  x0 (B*T, C) fp32 tokens.
// tab: HOST array of n_layer x 12 device pointers per block: weights qkv (3C,C) rows [key|query|value], proj (C,C),
// fc1 (4C,C), fc2 (C,4C) in the operand type; their fp32 biases; ln1 gamma, beta, ln2 gamma, beta.
// Outputs, each stacked over the blocks [n_layer][...]: xout (M,C) block outputs, x1 (M,C) post-attention residual,
// h1 / y / h2 (M,C), qkv (M,3C), a (M,4C) in the operand type, P and Pd (B,4,T,T) in the operand type (Pd may be null
// when attn_p == 0), LayerNorm statistics mean1 / rstd1 / mean2 / rstd2 (M).  Dropout streams of block l: attention
// seed + 3l + 1, projection seed + 3l + 2, MLP seed + 3l + 3 (the per-op path's numbering).
MMFN_API int mmfn_gpt_small_fwd(const float* x0, int B, int T, int C, int nh, int n_layer, int dtype, const void* const* tab,
                                float* xout, float* x1, void* h1, void* qkv, void* y, void* h2, void* a, void* P, void* Pd,
                                float* mean1, float* rstd1, float* mean2, float* rstd2,
                                float attn_p, float resid_p, uint64_t seed, float eps, cudaStream_t stream) {
  MMFN_CHECK_ARG(x0 && tab && xout && x1 && h1 && qkv && y && h2 && a && P && mean1 && rstd1 && mean2 && rstd2, "gpt_small_fwd: null pointer");
  MMFN_CHECK_ARG((C == 64 || C == 128) && nh == 4 && (T == 128 || T == 192), "gpt_small_fwd: needs C in {64,128}, 4 heads, T in {128,192}");
  MMFN_CHECK_ARG(n_layer >= 1 && n_layer <= GS_MAXL && B >= 1 && B <= 65535, "gpt_small_fwd: 1 <= n_layer <= 12, 1 <= B <= 65535");
  MMFN_CHECK_ARG(dtype == 1 || dtype == 2, "gpt_small_fwd: dtype must be MMFN_TF32 or MMFN_BF16");
  MMFN_CHECK_ARG(attn_p >= 0.f && attn_p < 1.f && resid_p >= 0.f && resid_p < 1.f && (attn_p == 0.f || Pd), "gpt_small_fwd: bad dropout arguments");
  GptFwdParams p;
  p.x0 = x0; p.xout = xout; p.x1 = x1; p.h1 = h1; p.qkv = qkv; p.y = y; p.h2 = h2; p.a = a; p.P = P; p.Pd = Pd;
  p.mean1 = mean1; p.rstd1 = rstd1; p.mean2 = mean2; p.rstd2 = rstd2;
  for (int l = 0; l < n_layer; ++l) {
    for (int i = 0; i < 4; ++i) {
      p.w[l][i] = tab[l * 12 + i];
      p.bias[l][i] = (const float*)tab[l * 12 + 4 + i];
      p.ln[l][i] = (const float*)tab[l * 12 + 8 + i];
      MMFN_CHECK_ARG(p.w[l][i] && p.bias[l][i] && p.ln[l][i], "gpt_small_fwd: null parameter pointer in the table");
      MMFN_CHECK_ARG(((uintptr_t)p.w[l][i] & 15) == 0, "gpt_small_fwd: weights must be 16-byte aligned");
    }
  }
  p.B = B; p.T = T; p.L = n_layer; p.attn_p = attn_p; p.resid_p = resid_p; p.eps = eps; p.seed = seed; p.trace = g_gpt_trace;
  const bool bf = dtype == 2;
  if (C == 64 && T == 128) return bf ? launch_gpt_fwd<64, 2, PrecBF>(p, stream) : launch_gpt_fwd<64, 2, PrecTF>(p, stream);
  if (C == 64 && T == 192) return bf ? launch_gpt_fwd<64, 3, PrecBF>(p, stream) : launch_gpt_fwd<64, 3, PrecTF>(p, stream);
  if (C == 128 && T == 128) return bf ? launch_gpt_fwd<128, 2, PrecBF>(p, stream) : launch_gpt_fwd<128, 2, PrecTF>(p, stream);
  return bf ? launch_gpt_fwd<128, 3, PrecBF>(p, stream) : launch_gpt_fwd<128, 3, PrecTF>(p, stream);
}
