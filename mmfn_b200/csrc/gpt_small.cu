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
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
// Distributed shared memory: address of the same variable in CTA `rank` of the cluster, and loads through the SHARED
// pipe.  (Generic loads of cluster.map_shared_rank() pointers go through the local/global queue instead: 19 % of this
// kernel's stall samples were lg_throttle on exactly those loads, profiles/r02_ncu_gpt_small_v1.txt.)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r; asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ uint2 ldsc2(uint32_t addr) {
  uint2 v; asm volatile("ld.shared::cluster.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory"); return v;
}
__device__ __forceinline__ uint4 ldsc4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared::cluster.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// n words (2 or 4) of a peer CTA's shared memory in one instruction
template <int N>
__device__ __forceinline__ void ldsc_words(uint32_t addr, uint32_t* out) {
  if constexpr (N == 2) { const uint2 v = ldsc2(addr); out[0] = v.x; out[1] = v.y; }
  else { const uint4 v = ldsc4(addr); out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w; }
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t saddr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
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
  // K, elements col / col + 1 of a key row, words of a head permuted (see the attention phase): word w of the head
  // (HW words) is stored at (w & 3) * (HW / 4) + (w >> 2)
  template <int HS>
  __device__ static __forceinline__ void st_k(uint32_t* base, int ld, int row, int col, float v0, float v1) {
    constexpr int HW = HS / 2;
    const int w = (col % HS) >> 1;
    base[row * ld + (col / HS) * HW + (w & 3) * (HW / 4) + (w >> 2)] = pack_bf16(v0, v1);
  }
  // V^T element (dim `row`, key `key` of the slab): key word w = key / 2 is stored at (w & 3) * 8 + (w >> 2)
  __device__ static __forceinline__ void st_vt(uint32_t* base, int ld, int row, int key, float v) {
    const int w = key >> 1;
    reinterpret_cast<__nv_bfloat16*>(base + row * ld + (w & 3) * 8 + (w >> 2))[key & 1] = __float2bfloat16_rn(v);
  }
  __device__ static __forceinline__ void st_global(void* base, long long idx, float v0, float v1) {
    *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = pack_bf16(v0, v1);
  }
  // staging tile (16 rows x 32 words, 16-byte chunks XOR-swizzled by the row): elements col, col + 1 (col even, < 64)
  __device__ static __forceinline__ void st_stage(uint32_t* stg, int r, int col, float v0, float v1) {
    stg[r * 32 + ((((col >> 3) ^ (r & 7)) << 2) | ((col >> 1) & 3))] = pack_bf16(v0, v1);
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
  template <int HS>
  __device__ static __forceinline__ void st_k(uint32_t* base, int ld, int row, int col, float v0, float v1) {
    constexpr int HW = HS;
    const int w = col % HS;                              // even: w and w + 1 share w >> 2
    uint32_t* r = base + row * ld + (col / HS) * HW + (w >> 2);
    r[(w & 3) * (HW / 4)] = __float_as_uint(v0);
    r[((w + 1) & 3) * (HW / 4)] = __float_as_uint(v1);
  }
  // V^T element (dim `row`, key): the PV product pairs MMA k index t with key 8 k8 + 2t and t + 4 with 8 k8 + 2t + 1;
  // key is stored at ((key & 7) >> 1) * 16 + (key >> 3) * 2 + (key & 1): a thread's 16 words of the slab are contiguous
  __device__ static __forceinline__ void st_vt(uint32_t* base, int ld, int row, int key, float v) {
    base[row * ld + ((key & 7) >> 1) * 16 + (key >> 3) * 2 + (key & 1)] = __float_as_uint(v);
  }
  __device__ static __forceinline__ void st_global(void* base, long long idx, float v0, float v1) {
    *reinterpret_cast<float2*>(reinterpret_cast<float*>(base) + idx) = make_float2(v0, v1);
  }
  __device__ static __forceinline__ void st_stage(uint32_t* stg, int r, int col, float v0, float v1) {   // col even, < 32
    *reinterpret_cast<float2*>(stg + r * 32 + ((((col >> 2) ^ (r & 7)) << 2) | (col & 3))) = make_float2(v0, v1);
  }
};

// Coalesced write-out of a warp's 16-row tile that sits in shared memory (NW words per row, row stride ldw words;
// SWZ: the XOR-swizzled staging tile): 16-byte chunks, a row's chunks on consecutive lanes -> full 128-byte lines.
// (Storing the MMA fragments directly costs one 4/8-byte store per thread = 8 partial lines per instruction: those
// stores were ~half of the first version's attention phase.)
template <int NW, bool SWZ>
__device__ __forceinline__ void flush_rows(const uint32_t* src, int ldw, void* dst, long long ld_bytes) {
  constexpr int CH = NW / 4, RPI = 32 / CH;              // chunks per row, rows per instruction
  const int lane = threadIdx.x & 31;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 16 / RPI; ++i) {
    const int r = i * RPI + lane / CH, c = lane % CH;
    const uint4 v = *reinterpret_cast<const uint4*>(src + r * ldw + ((SWZ ? (c ^ (r & 7)) : c) << 2));
    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(dst) + r * ld_bytes + c * 16) = v;
  }
  __syncwarp();
}

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
  static constexpr int PR = VT + C * V_LD;              // 2 x 13C floats: biases and LayerNorm parameters of this / the next block
  static constexpr int SG = PR + 2 * 13 * C;            // per-warp staging tiles (16 x 32 words) for coalesced write-out
  static constexpr int WS = SG + (GS_THREADS / 32) * 512;
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

// acc[NT][4] += A[64 x C] (shared, operand-typed, row stride lda words) . Wblk[C x C]^T for this warp's 16 rows x C/2
// columns.  `ti` is the stream index of the block's first tile; tiles ti .. ti + STAGES - 2 are already in flight;
// load(wbuf, i) issues the cp.async copies of stream tile i, `total` is the length of the stream.
template <int C, class Pr, class Loader>
__device__ __forceinline__ void gemm_block(const Loader& load, int total, uint32_t* smem_w, const uint32_t* A, int lda,
                                           float (*acc)[4], int& ti) {
  using L = Lay<C, Pr>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = warp & 3, chalf = warp >> 2;
  for (int kt = 0; kt < L::TILES; ++kt, ++ti) {
    cp_async_wait<L::STAGES - 2>();                      // this thread's part of tile ti has landed
    __syncthreads();                                     // ... everyone's; and everyone is done with tile ti - 1's slot
    if (ti + L::STAGES - 1 < total) load(smem_w + ((ti + L::STAGES - 1) % L::STAGES) * L::WTILE, ti + L::STAGES - 1);
    cp_async_commit();                                   // (an empty group keeps the group count uniform)
    // fragments by ldmatrix: in 32-bit words an 8 x 8 b16 matrix is 8 rows x 4 words and thread (g, t) receives word t of
    // row g -- the A / B fragment pieces of both operand types.  A: rows 0-7 / 8-15 x words 0-3 / 4-7 of the k-step;
    // B: two column tiles x words 0-3 / 4-7.  (One instruction instead of four / four 32-bit loads.)
    const uint32_t wt_s = (uint32_t)__cvta_generic_to_shared(smem_w + (ti % L::STAGES) * L::WTILE) +
                          (uint32_t)((chalf * (C / 2) + (lane >> 4) * 8 + (lane & 7)) * L::W_LD + ((lane >> 3) & 1) * 4) * 4u;
    const uint32_t a_s = (uint32_t)__cvta_generic_to_shared(A) +
                         (uint32_t)((rb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * lda + kt * 32 + (lane >> 4) * 4) * 4u;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[4];
      ldmatrix_x4(a_s + ks * 32, a);
#pragma unroll
      for (int j = 0; j < L::NT; j += 2) {
        uint32_t bb[4];
        ldmatrix_x4(wt_s + (uint32_t)(j * 8 * L::W_LD + ks * 8) * 4u, bb);
        Pr::mma(acc[j], a, bb[0], bb[1]);
        Pr::mma(acc[j + 1], a, bb[2], bb[3]);
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
#pragma unroll 1
  for (int r4 = 0; r4 < 8; r4 += 4) {                   // four rows in flight: the shuffle chains overlap
    float v[4][CPL], s[4], q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s[u] = 0.f;
#pragma unroll
      for (int i = 0; i < CPL; ++i) { v[u][i] = xs[(warp * 8 + r4 + u) * L::XS_LD + lane * CPL + i]; s[u] += v[u][i]; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s[u] = s[u] / (float)C;
      q[u] = 0.f;
#pragma unroll
      for (int i = 0; i < CPL; ++i) { v[u][i] -= s[u]; q[u] = fmaf(v[u][i], v[u][i], q[u]); }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] += __shfl_xor_sync(0xffffffffu, q[u], o);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int row = warp * 8 + r4 + u;
      const float rs = rsqrtf(q[u] / (float)C + eps);
#pragma unroll
      for (int i = 0; i < CPL; ++i) v[u][i] = fmaf(v[u][i] * rs, gm[i], bt[i]);
#pragma unroll
      for (int i = 0; i < CPL; i += 2) {
        Pr::st_smem(hs, L::A_LD, row, lane * CPL + i, v[u][i], v[u][i + 1]);
        Pr::st_global(hout, (row0 + row) * C + lane * CPL + i, v[u][i], v[u][i + 1]);
      }
      if (lane == 0) { mean[row0 + row] = s[u]; rstd[row0 + row] = rs; }
    }
  }
}

// biases and LayerNorm parameters of one block -> shared: [bqkv 3C | bproj C | b1 4C | b2 C | ln1 g, b | ln2 g, b]
template <int C>
__device__ __forceinline__ void load_params(const GptFwdParams& p, float* dst, int layer) {
  for (int i = threadIdx.x; i < 13 * C; i += GS_THREADS) {
    const float* src;
    if (i < 3 * C) src = p.bias[layer][0] + i;
    else if (i < 4 * C) src = p.bias[layer][1] + (i - 3 * C);
    else if (i < 8 * C) src = p.bias[layer][2] + (i - 4 * C);
    else if (i < 9 * C) src = p.bias[layer][3] + (i - 8 * C);
    else src = p.ln[layer][(i - 9 * C) / C] + (i - 9 * C) % C;
    cp_async4(dst + i, src);
  }
}

template <int C, int NS, class Pr>
__global__ void __launch_bounds__(GS_THREADS, 1) gpt_small_fwd_kernel(const __grid_constant__ GptFwdParams p) {
  using L = Lay<C, Pr>;
  constexpr int NH = 4, HS = C / NH, EPW = Pr::EPW;
  constexpr int HW = HS / EPW;                          // words per head row
  constexpr int QK_STEPS = HW / 8;                      // MMA k-steps of one score
  constexpr int NT = L::NT;
  constexpr int NWE = (C / 2) / EPW;                    // words per row of a warp's operand-typed output tile (16 or 32)
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
  float* prm = reinterpret_cast<float*>(smem + L::PR);
  uint32_t* stg = smem + L::SG + warp * 512;            // this warp's staging tile

  int ti = 0;                                            // weight-tile stream position
  const int total_tiles = p.L * 12 * L::TILES;
  const auto loader = [&p](uint32_t* wbuf, int i) { load_tile<C, Pr>(p, wbuf, i); };
  load_params<C>(p, prm, 0);
  cp_async_commit();
#pragma unroll 1
  for (int s = 0; s < L::STAGES - 1; ++s) {
    if (s < total_tiles) loader(smem + L::WS + s * L::WTILE, s);
    cp_async_commit();
  }
  // residual stream of the slab
  for (int i = threadIdx.x; i < GS_ROWS * (C / 4); i += GS_THREADS) {
    const int r = i / (C / 4), c4 = i - r * (C / 4);
    *reinterpret_cast<float4*>(xs + r * L::XS_LD + c4 * 4) = *reinterpret_cast<const float4*>(p.x0 + (row0 + r) * C + c4 * 4);
  }
  cp_async_wait<L::STAGES - 1>();                        // the oldest group = block 0's parameters
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
    const float* pl = prm + (layer & 1) * 13 * C;         // this block's biases / LayerNorm parameters (shared)
    if (layer + 1 < p.L) load_params<C>(p, prm + ((layer + 1) & 1) * 13 * C, layer + 1);   // rides in the next tile's group
    // ---- ln1
    layernorm_slab<C, Pr>(xs, hs, pl + 9 * C, pl + 10 * C, reinterpret_cast<typename Pr::elem*>(p.h1) + lM * C,
                          p.mean1 + lM, p.rstd1 + lM, row0, p.eps);
    gs_stamp(p.trace, layer * 10 + 1);
    // ---- key / query / value
    typename Pr::elem* qkv_g = reinterpret_cast<typename Pr::elem*>(p.qkv) + lM * 3 * C;
#pragma unroll 1
    for (int part = 0; part < 3; ++part) {
      float acc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      gemm_block<C, Pr>(loader, total_tiles, smem + L::WS, hs, L::A_LD, acc, ti);
      const float* bias = pl + part * C;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int col = chalf * (C / 2) + j * 8 + 2 * t;
        const float b0 = bias[col], b1 = bias[col + 1];
        const float v00 = acc[j][0] + b0, v01 = acc[j][1] + b1, v10 = acc[j][2] + b0, v11 = acc[j][3] + b1;
        if (part == 1) {
          Pr::st_smem(qs, L::A_LD, r0, col, v00, v01);
          Pr::st_smem(qs, L::A_LD, r0 + 8, col, v10, v11);
        } else {
          if (part == 0) {                                // K: head words permuted for wide DSMEM loads
            Pr::template st_k<HS>(ks, L::A_LD, r0, col, v00, v01);
            Pr::template st_k<HS>(ks, L::A_LD, r0 + 8, col, v10, v11);
          } else {                                        // V^T: [dim][key of the slab], key words permuted
            Pr::st_vt(vt, L::V_LD, col, r0, v00); Pr::st_vt(vt, L::V_LD, col + 1, r0, v01);
            Pr::st_vt(vt, L::V_LD, col, r0 + 8, v10); Pr::st_vt(vt, L::V_LD, col + 1, r0 + 8, v11);
          }
          Pr::st_stage(stg, g, j * 8 + 2 * t, v00, v01);  // row-major copy for the write-out
          Pr::st_stage(stg, g + 8, j * 8 + 2 * t, v10, v11);
        }
      }
      {
        char* dst = reinterpret_cast<char*>(qkv_g + (row0 + rb * 16) * 3 * C + part * C + chalf * (C / 2));
        if (part == 1) flush_rows<NWE, false>(qs + rb * 16 * L::A_LD + chalf * NWE, L::A_LD, dst, 3LL * C * sizeof(typename Pr::elem));
        else flush_rows<NWE, true>(stg, 32, dst, 3LL * C * sizeof(typename Pr::elem));
      }
    }
    gs_stamp(p.trace, layer * 10 + 2);
    cluster.sync();                                       // K, V^T of every slab of the sample are in place
    gs_stamp(p.trace, layer * 10 + 3);
    // ---- attention: warp = (16 query rows, 2 heads); scores, softmax and P V in registers.
    // Slabs are visited in rotated order -- the CTA's own first, peers after -- and a peer's K / V^T fragments are
    // requested one slab ahead of their MMAs: a DSMEM round trip is ~1 us, six exposed ones per head were most of
    // this phase in the first version.
    int rk[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) rk[i] = slab + i >= NS ? slab + i - NS : slab + i;
    constexpr int VPT = 64 / EPW / 4;                     // V^T words per thread, dim row and slab: 8 (bf16) / 16 (fp32)
    uint32_t kb[2][8][HW / 4];
    // a thread's words of a key's head row (MMA words st * 8 + t and st * 8 + t + 4) are contiguous: one load per key
    auto load_k = [&](uint32_t (*dst)[HW / 4], int i, int h) {
      const uint32_t base = mapa_u32((uint32_t)__cvta_generic_to_shared(ks), rk[i]) + (uint32_t)(g * L::A_LD + h * HW + t * (HW / 4)) * 4u;
#pragma unroll
      for (int j = 0; j < 8; ++j) ldsc_words<HW / 4>(base + (uint32_t)(j * 8 * L::A_LD) * 4u, dst[j]);
    };
    // 16-dim heads (8 registers per slab): the next head's first K slabs are requested during this head's P V phase;
    // the 32-dim variants have no registers to carry them across the softmax
    constexpr bool K_EARLY = HW / 4 == 2;
    if constexpr (K_EARLY) { load_k(kb[0], 0, chalf * 2); load_k(kb[1], 1, chalf * 2); }
#pragma unroll 1
    for (int u = 0; u < 2; ++u) {
      const int h = chalf * 2 + u;
      if constexpr (!K_EARLY) { load_k(kb[0], 0, h); load_k(kb[1], 1, h); }
      const uint32_t voff = (uint32_t)((h * HS + g) * L::V_LD + VPT * t) * 4u;
      uint32_t vb[2][HS / 8][VPT];
      auto load_v = [&](uint32_t (*dst)[VPT], int i) {
        const uint32_t base = mapa_u32((uint32_t)__cvta_generic_to_shared(vt), rk[i]) + voff;
#pragma unroll
        for (int jn = 0; jn < HS / 8; ++jn) {
#pragma unroll
          for (int q4 = 0; q4 < VPT / 4; ++q4) ldsc_words<4>(base + (uint32_t)(jn * 8 * L::V_LD + 4 * q4) * 4u, dst[jn] + 4 * q4);
        }
      };
      uint32_t qf[QK_STEPS][4];
      {
        const uint32_t* qr = qs + r0 * L::A_LD + h * HW + t;
#pragma unroll
        for (int s = 0; s < QK_STEPS; ++s) {
          qf[s][0] = qr[s * 8]; qf[s][1] = qr[8 * L::A_LD + s * 8]; qf[s][2] = qr[s * 8 + 4]; qf[s][3] = qr[8 * L::A_LD + s * 8 + 4];
        }
      }
      float sa[NS * 8][4];                                // sa[8 i + j]: keys 64 rk[i] + 8 j .. + 7
#pragma unroll
      for (int j = 0; j < NS * 8; ++j) sa[j][0] = sa[j][1] = sa[j][2] = sa[j][3] = 0.f;
      auto mma_k = [&](int i, uint32_t (*src)[HW / 4]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int st = 0; st < QK_STEPS; ++st) Pr::mma(sa[i * 8 + j], qf[st], src[j][2 * st], src[j][2 * st + 1]);
        }
      };
      mma_k(0, kb[0]);                                    // (this head's first two slabs were requested a phase ago)
      if constexpr (NS > 2) load_k(kb[0], 2, h);
      mma_k(1, kb[1]);
      if constexpr (NS > 2) mma_k(2, kb[0]);
      load_v(vb[0], 0);                                   // in flight during the softmax
      load_v(vb[1], 1);
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
      const long long pofs = (long long)layer * p.B * NH * T * T;
      const long long prow0 = (((long long)b * NH + h) * T + slab * GS_ROWS + r0) * T, prow1 = prow0 + 8LL * T;   // hash keys
      const long long ptile = pofs + (((long long)b * NH + h) * T + slab * GS_ROWS + rb * 16) * T;   // P[b, h, first row of the warp, 0]
      const bool drop = p.attn_p > 0.f;
      constexpr int TPS = 4 * EPW;                        // score tiles per staging pass (64 keys bf16, 32 keys fp32)
#pragma unroll
      for (int j = 0; j < NS * 8; ++j) { sa[j][0] *= i0; sa[j][1] *= i0; sa[j][2] *= i1; sa[j][3] *= i1; }
      auto store_probs = [&](void* dstp) {
#pragma unroll
        for (int ps = 0; ps < NS * 8 / TPS; ++ps) {
#pragma unroll
          for (int jj = 0; jj < TPS; ++jj) {
            Pr::st_stage(stg, g, jj * 8 + 2 * t, sa[ps * TPS + jj][0], sa[ps * TPS + jj][1]);
            Pr::st_stage(stg, g + 8, jj * 8 + 2 * t, sa[ps * TPS + jj][2], sa[ps * TPS + jj][3]);
          }
          const int key0 = rk[ps * TPS / 8] * 64 + (ps * TPS % 8) * 8;      // first key of this pass
          flush_rows<32, true>(stg, 32, reinterpret_cast<typename Pr::elem*>(dstp) + ptile + key0, (long long)T * sizeof(typename Pr::elem));
        }
      };
      store_probs(p.P);
      if (drop) {
        // one 64-bit hash covers four consecutive keys of a row = this thread's pair and its quad neighbour's: the even
        // thread hashes row g, the odd one row g + 8, and they swap the halves the other needs
        const long long prow_mine = (t & 1) ? prow1 : prow0;
#pragma unroll
        for (int j = 0; j < NS * 8; ++j) {
          const int col = rk[j / 8] * 64 + (j % 8) * 8 + 2 * t;
          const uint64_t hh = mmfn_hash64(seed_a, (uint64_t)(prow_mine + col) >> 2);
          const uint32_t mine = (t & 1) ? (uint32_t)(hh >> 32) : (uint32_t)hh;      // this thread's columns, the row it hashed
          const uint32_t give = (t & 1) ? (uint32_t)hh : (uint32_t)(hh >> 32);      // the neighbour's columns of that row
          const uint32_t got = __shfl_xor_sync(0xffffffffu, give, 1);               // this thread's columns, the other row
          const uint32_t f0 = (t & 1) ? got : mine, f1 = (t & 1) ? mine : got;      // rows g, g + 8
          sa[j][0] = (f0 & 0xFFFFu) >= thr_a ? sa[j][0] * keep_a : 0.f;
          sa[j][1] = (f0 >> 16) >= thr_a ? sa[j][1] * keep_a : 0.f;
          sa[j][2] = (f1 & 0xFFFFu) >= thr_a ? sa[j][2] * keep_a : 0.f;
          sa[j][3] = (f1 >> 16) >= thr_a ? sa[j][3] * keep_a : 0.f;
        }
      }
      // y = dropout(P) V: the score fragments ARE the A operand
      float oa[HS / 8][4];
#pragma unroll
      for (int j = 0; j < HS / 8; ++j) oa[j][0] = oa[j][1] = oa[j][2] = oa[j][3] = 0.f;
      auto mma_v = [&](int i, uint32_t (*src)[VPT]) {
        if constexpr (EPW == 2) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {                // 16 keys per step = two score tiles; src[.][2 k4 + b] = the step's b0 / b1
            const int kk = i * 4 + k4;
            uint32_t a[4];
            a[0] = pack_bf16(sa[2 * kk][0], sa[2 * kk][1]); a[1] = pack_bf16(sa[2 * kk][2], sa[2 * kk][3]);
            a[2] = pack_bf16(sa[2 * kk + 1][0], sa[2 * kk + 1][1]); a[3] = pack_bf16(sa[2 * kk + 1][2], sa[2 * kk + 1][3]);
#pragma unroll
            for (int jn = 0; jn < HS / 8; ++jn) Pr::mma(oa[jn], a, src[jn][2 * k4], src[jn][2 * k4 + 1]);
          }
        } else {
#pragma unroll
          for (int k8 = 0; k8 < 8; ++k8) {                // 8 keys per step; MMA k index t <-> key 2t, t + 4 <-> key 2t + 1
            const int kk = i * 8 + k8;
            uint32_t a[4];
            a[0] = __float_as_uint(sa[kk][0]); a[1] = __float_as_uint(sa[kk][2]);
            a[2] = __float_as_uint(sa[kk][1]); a[3] = __float_as_uint(sa[kk][3]);
#pragma unroll
            for (int jn = 0; jn < HS / 8; ++jn) Pr::mma(oa[jn], a, src[jn][2 * k8], src[jn][2 * k8 + 1]);
          }
        }
      };
      mma_v(0, vb[0]);
      if constexpr (NS > 2) load_v(vb[0], 2);             // the last slab's V^T and the next head's first K slabs travel
      if constexpr (K_EARLY) {                            // while dropout(P) is written out
        if (u == 0) { load_k(kb[0], 0, h + 1); load_k(kb[1], 1, h + 1); }
      }
      if (drop) store_probs(p.Pd);
      mma_v(1, vb[1]);
      if constexpr (NS > 2) mma_v(2, vb[0]);
#pragma unroll
      for (int jn = 0; jn < HS / 8; ++jn) {
        const int col = h * HS + jn * 8 + 2 * t;
        Pr::st_smem(hs, L::A_LD, r0, col, oa[jn][0], oa[jn][1]);
        Pr::st_smem(hs, L::A_LD, r0 + 8, col, oa[jn][2], oa[jn][3]);
      }
    }
    // the warp's two heads are its column half of y: write it out from the operand buffer
    flush_rows<NWE, false>(hs + rb * 16 * L::A_LD + chalf * NWE, L::A_LD,
                           reinterpret_cast<typename Pr::elem*>(p.y) + lM * C + (row0 + rb * 16) * C + chalf * (C / 2),
                           (long long)C * sizeof(typename Pr::elem));
    gs_stamp(p.trace, layer * 10 + 4);
    cluster.sync();                                       // peers are done with this slab's K / V^T
    gs_stamp(p.trace, layer * 10 + 5);
    // ---- projection + residual
    {
      float acc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      gemm_block<C, Pr>(loader, total_tiles, smem + L::WS, hs, L::A_LD, acc, ti);
      const float* bias = pl + 3 * C;
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
      }
      flush_rows<C / 2, false>(reinterpret_cast<const uint32_t*>(xs) + rb * 16 * L::XS_LD + chalf * (C / 2), L::XS_LD,
                               x1_g + (row0 + rb * 16) * C + chalf * (C / 2), (long long)C * 4);
    }
    __syncthreads();
    gs_stamp(p.trace, layer * 10 + 6);
    // ---- ln2
    layernorm_slab<C, Pr>(xs, hs, pl + 11 * C, pl + 12 * C, reinterpret_cast<typename Pr::elem*>(p.h2) + lM * C,
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
        gemm_block<C, Pr>(loader, total_tiles, smem + L::WS, hs, L::A_LD, acc, ti);
        const float* bias = pl + 4 * C + c * C;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const int col = chalf * (C / 2) + j * 8 + 2 * t;
          const float b0 = bias[col], b1 = bias[col + 1];
          const float v00 = fmaxf(acc[j][0] + b0, 0.f), v01 = fmaxf(acc[j][1] + b1, 0.f);
          const float v10 = fmaxf(acc[j][2] + b0, 0.f), v11 = fmaxf(acc[j][3] + b1, 0.f);
          Pr::st_smem(qs, L::A_LD, r0, col, v00, v01);
          Pr::st_smem(qs, L::A_LD, r0 + 8, col, v10, v11);
        }
        flush_rows<NWE, false>(qs + rb * 16 * L::A_LD + chalf * NWE, L::A_LD, a_g + (row0 + rb * 16) * 4 * C + c * C + chalf * (C / 2),
                               4LL * C * sizeof(typename Pr::elem));
        gemm_block<C, Pr>(loader, total_tiles, smem + L::WS, qs, L::A_LD, acc2, ti);   // (its first __syncthreads publishes the chunk)
      }
      const float* bias = pl + 8 * C;
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
      }
      flush_rows<C / 2, false>(reinterpret_cast<const uint32_t*>(xs) + rb * 16 * L::XS_LD + chalf * (C / 2), L::XS_LD,
                               xo_g + (row0 + rb * 16) * C + chalf * (C / 2), (long long)C * 4);
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


// ------------------------------------------------------------------------------------------------------------------
// Backward, row-local part.  Between two attention backwards everything is row-local (no token mixes with another):
//   part A (block `hi`):  dh1 = dqkv . Wqkv ;  dx = dx1 + LayerNorm1'(dh1)                                   [3 blocks]
//   part B (block `lo` = hi - 1):  dz = dx o mask_mlp ;  da = (dz . W2) o [a > 0] ;  dh2 = da . W1 ;
//                                  dx1 = dx + LayerNorm2'(dh2) ;  dzp = dx1 o mask_proj ;  dy = dzp . Wproj  [9 blocks]
// = 6 dependent launches of the per-op chain (4 dgrad GEMMs + 2 LayerNorm backward) in ONE, any 64 rows per CTA, no
// cluster.  The products are written out for the weight-gradient GEMMs / LayerNorm parameter reductions, which are leaves
// of the backward graph and stay on the side streams (dz, da, dzp, dqkv are their `dy` operands; dh1, dh2 the LayerNorm
// ones); dy feeds the attention backward.  Same GEMM primitive as the forward; B operands come from TRANSPOSED weight
// copies (mmfn_gpt_small_transpose) so that the contraction index is contiguous, as in the forward.
struct GptBwdParams {
  int has_a, has_b;
  // part A
  const void* dqkv;                            // (M, 3C) operand-typed
  const float *dx1_in, *xa, *mean_a, *rstd_a, *gamma_a;
  const void* wqkvT;                           // (C, 3C)
  float *dh1, *dx_out;                         // dx_out: written when there is no part B (gradient leaving the GPT)
  // part B
  const float* dx2_in;                         // read when there is no part A (gradient entering the GPT's last block)
  const void* a;                               // (M, 4C) saved MLP hidden (ReLU mask)
  const float *x1, *mean_b, *rstd_b, *gamma_b;
  const void *w2T, *w1T, *wpT;                 // (4C, C), (C, 4C), (C, C)
  void *dz, *da, *dzp, *dy;                    // operand-typed (M,C), (M,4C), (M,C), (M,C)
  float *dh2, *dx1_out;
  float resid_p;
  unsigned long long seed_m, seed_p;
};

template <int C, class Pr>
struct BLay {
  using L = Lay<C, Pr>;
  static constexpr int EPW = Pr::EPW;
  static constexpr int AQ_LD = 3 * C / EPW + 4;
  static constexpr int DXS = 0;                          // fp32 [64][C + 8] gradient of the residual stream
  static constexpr int FB = DXS + GS_ROWS * L::XS_LD;    // fp32 [64][C + 8] GEMM result entering a LayerNorm backward
  static constexpr int AQ = FB + GS_ROWS * L::XS_LD;     // dqkv slab [64][3C]; part B: two [64][C] operand buffers
  static constexpr int GM = AQ + GS_ROWS * AQ_LD;        // gamma of the two LayerNorms
  static constexpr int WS = GM + 2 * C;
  static constexpr int WORDS = WS + L::STAGES * L::WTILE;
};

// dx = dres + LayerNorm'(dh) for the slab: dh in fb (shared fp32), x / mean / rstd from global, gamma shared.
// dres_g != null: residual gradient from global, else from dxs (in place).  Result -> dxs and (out_g != null) global.
template <int C, class Pr>
__device__ __forceinline__ void ln_bwd_slab(const float* fb, float* dxs, const float* x_g, const float* mean, const float* rstd,
                                            const float* gamma, const float* dres_g, float* out_g, long long row0) {
  using L = Lay<C, Pr>;
  constexpr int CPL = C / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float gm[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) gm[i] = gamma[lane * CPL + i];
#pragma unroll 1
  for (int r4 = 0; r4 < 8; r4 += 4) {
    float gv[4][CPL], xh[4][CPL], s1[4], s2[4], rs[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int row = warp * 8 + r4 + u;
      const float mu = mean[row0 + row];
      rs[u] = rstd[row0 + row];
      s1[u] = s2[u] = 0.f;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const float xv = x_g[(row0 + row) * C + lane * CPL + i];
        xh[u][i] = (xv - mu) * rs[u];
        gv[u][i] = fb[row * L::XS_LD + lane * CPL + i] * gm[i];
        s1[u] += gv[u][i];
        s2[u] = fmaf(gv[u][i], xh[u][i], s2[u]);
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) { s1[u] += __shfl_xor_sync(0xffffffffu, s1[u], o); s2[u] += __shfl_xor_sync(0xffffffffu, s2[u], o); }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int row = warp * 8 + r4 + u;
      const float m1 = s1[u] / (float)C, m2 = s2[u] / (float)C;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const int col = lane * CPL + i;
        float o = rs[u] * (gv[u][i] - m1 - xh[u][i] * m2);
        o += dres_g ? dres_g[(row0 + row) * C + col] : dxs[row * L::XS_LD + col];
        dxs[row * L::XS_LD + col] = o;
        if (out_g) out_g[(row0 + row) * C + col] = o;
      }
    }
  }
}

// abuf (operand-typed [64][C]) and out_g = dxs o dropout mask (seed, index row * C + col): the gradient entering a
// residual branch's dropout
template <int C, class Pr>
__device__ __forceinline__ void masked_copy_slab(const float* dxs, uint32_t* abuf, void* out_g, long long row0, float p, uint64_t seed) {
  using L = Lay<C, Pr>;
  const uint32_t thr = mmfn_drop_threshold(p);
  const float keep = 1.0f / (1.0f - p);
  for (int i = threadIdx.x; i < GS_ROWS * (C / 4); i += GS_THREADS) {
    const int r = i / (C / 4), c = (i - r * (C / 4)) * 4;
    float4 v = *reinterpret_cast<const float4*>(dxs + r * L::XS_LD + c);
    if (p > 0.f) {
      const uint64_t h = mmfn_hash64(seed, (uint64_t)((row0 + r) * C + c) >> 2);
      const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
      v.x = (lo & 0xFFFFu) >= thr ? v.x * keep : 0.f;
      v.y = (lo >> 16) >= thr ? v.y * keep : 0.f;
      v.z = (hi & 0xFFFFu) >= thr ? v.z * keep : 0.f;
      v.w = (hi >> 16) >= thr ? v.w * keep : 0.f;
    }
    Pr::st_smem(abuf, L::A_LD, r, c, v.x, v.y);
    Pr::st_smem(abuf, L::A_LD, r, c + 2, v.z, v.w);
    Pr::st_global(out_g, (row0 + r) * C + c, v.x, v.y);
    Pr::st_global(out_g, (row0 + r) * C + c + 2, v.z, v.w);
  }
}

__device__ __forceinline__ float2 ld_pair(const __nv_bfloat16* p) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); }
__device__ __forceinline__ float2 ld_pair(const float* p) { return *reinterpret_cast<const float2*>(p); }

template <int C, class Pr>
__global__ void __launch_bounds__(GS_THREADS, 1) gpt_small_bwd_rows_kernel(const __grid_constant__ GptBwdParams p) {
  using L = Lay<C, Pr>;
  using BL = BLay<C, Pr>;
  using E = typename Pr::elem;
  constexpr int NT = L::NT;
  constexpr int NWE = (C / 2) / L::EPW;
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ const E* tab_ptr[12];
  __shared__ int tab_ld[12];
  const long long row0 = (long long)blockIdx.x * GS_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int rb = warp & 3, chalf = warp >> 2;
  const int r0 = rb * 16 + g;
  float* dxs = reinterpret_cast<float*>(smem + BL::DXS);
  float* fb = reinterpret_cast<float*>(smem + BL::FB);
  uint32_t* aq = smem + BL::AQ;
  uint32_t* abuf0 = aq;
  uint32_t* abuf1 = aq + GS_ROWS * L::A_LD;
  float* gam = reinterpret_cast<float*>(smem + BL::GM);
  if (threadIdx.x == 0) {
    int n = 0;
    if (p.has_a) for (int j = 0; j < 3; ++j) { tab_ptr[n] = (const E*)p.wqkvT + j * C; tab_ld[n++] = 3 * C; }
    if (p.has_b) {
      for (int c = 0; c < 4; ++c) {
        tab_ptr[n] = (const E*)p.w2T + (long long)c * C * C; tab_ld[n++] = C;
        tab_ptr[n] = (const E*)p.w1T + c * C; tab_ld[n++] = 4 * C;
      }
      tab_ptr[n] = (const E*)p.wpT; tab_ld[n++] = C;
    }
  }
  for (int i = threadIdx.x; i < 2 * C; i += GS_THREADS)
    gam[i] = i < C ? (p.has_a ? p.gamma_a[i] : 0.f) : (p.has_b ? p.gamma_b[i - C] : 0.f);
  __syncthreads();
  const int total_tiles = (3 * p.has_a + 9 * p.has_b) * L::TILES;
  const auto loader = [&](uint32_t* wbuf, int ti) {
    const int blk = ti / L::TILES, kt = ti - blk * L::TILES;
    const E* src = tab_ptr[blk] + kt * L::KT;
    const int ldw = tab_ld[blk];
    for (int i = threadIdx.x; i < C * 8; i += GS_THREADS) {
      const int n = i >> 3, ch = i & 7;
      cp_async16(wbuf + n * L::W_LD + ch * 4, reinterpret_cast<const char*>(src + (long long)n * ldw) + ch * 16);
    }
  };
  int ti = 0;
#pragma unroll 1
  for (int s = 0; s < L::STAGES - 1; ++s) {
    if (s < total_tiles) loader(smem + BL::WS + s * L::WTILE, s);
    cp_async_commit();
  }

  if (p.has_a) {
    // ---- dqkv slab -> shared (rows of 3C operand-typed elements = 3C / EPW words)
    constexpr int RW = 3 * C / L::EPW;
    for (int i = threadIdx.x; i < GS_ROWS * (RW / 4); i += GS_THREADS) {
      const int r = i / (RW / 4), c4 = i - r * (RW / 4);
      *reinterpret_cast<uint4*>(aq + r * BL::AQ_LD + c4 * 4) =
          *reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(p.dqkv) + (row0 + r) * RW + c4 * 4);
    }
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll 1
    for (int j = 0; j < 3; ++j)                          // (the first tile's __syncthreads publishes the slab)
      gemm_block<C, Pr>(loader, total_tiles, smem + BL::WS, aq + j * (C / L::EPW), BL::AQ_LD, acc, ti);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int col = chalf * (C / 2) + j * 8 + 2 * t;
      *reinterpret_cast<float2*>(fb + r0 * L::XS_LD + col) = make_float2(acc[j][0], acc[j][1]);
      *reinterpret_cast<float2*>(fb + (r0 + 8) * L::XS_LD + col) = make_float2(acc[j][2], acc[j][3]);
    }
    flush_rows<C / 2, false>(reinterpret_cast<const uint32_t*>(fb) + rb * 16 * L::XS_LD + chalf * (C / 2), L::XS_LD,
                             p.dh1 + (row0 + rb * 16) * C + chalf * (C / 2), (long long)C * 4);
    __syncthreads();
    ln_bwd_slab<C, Pr>(fb, dxs, p.xa, p.mean_a, p.rstd_a, gam, p.dx1_in, p.has_b ? nullptr : p.dx_out, row0);
    __syncthreads();
  } else {
    for (int i = threadIdx.x; i < GS_ROWS * (C / 4); i += GS_THREADS) {
      const int r = i / (C / 4), c4 = i - r * (C / 4);
      *reinterpret_cast<float4*>(dxs + r * L::XS_LD + c4 * 4) = *reinterpret_cast<const float4*>(p.dx2_in + (row0 + r) * C + c4 * 4);
    }
    __syncthreads();
  }
  if (p.has_b) {
    // ---- dz = dx o mask of the MLP branch's dropout
    masked_copy_slab<C, Pr>(dxs, abuf0, p.dz, row0, p.resid_p, mmfn_drop_seed(p.seed_m));
    float acc2[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc2[j][0] = acc2[j][1] = acc2[j][2] = acc2[j][3] = 0.f;
    const E* a_g = reinterpret_cast<const E*>(p.a);
    E* da_g = reinterpret_cast<E*>(p.da);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float acc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      // the ReLU mask of this thread's fragment: issue the loads before the GEMM so they are back when it ends
      float2 m0[NT], m1[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int col = c * C + chalf * (C / 2) + j * 8 + 2 * t;
        m0[j] = ld_pair(a_g + (row0 + r0) * 4 * C + col);
        m1[j] = ld_pair(a_g + (row0 + r0 + 8) * 4 * C + col);
      }
      gemm_block<C, Pr>(loader, total_tiles, smem + BL::WS, abuf0, L::A_LD, acc, ti);     // dz . W2[:, chunk c]
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int col = chalf * (C / 2) + j * 8 + 2 * t;
        const float v00 = m0[j].x > 0.f ? acc[j][0] : 0.f, v01 = m0[j].y > 0.f ? acc[j][1] : 0.f;
        const float v10 = m1[j].x > 0.f ? acc[j][2] : 0.f, v11 = m1[j].y > 0.f ? acc[j][3] : 0.f;
        Pr::st_smem(abuf1, L::A_LD, r0, col, v00, v01);
        Pr::st_smem(abuf1, L::A_LD, r0 + 8, col, v10, v11);
      }
      flush_rows<NWE, false>(abuf1 + rb * 16 * L::A_LD + chalf * NWE, L::A_LD, da_g + (row0 + rb * 16) * 4 * C + c * C + chalf * (C / 2),
                             4LL * C * sizeof(E));
      gemm_block<C, Pr>(loader, total_tiles, smem + BL::WS, abuf1, L::A_LD, acc2, ti);    // += da_c . W1[chunk c, :]
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int col = chalf * (C / 2) + j * 8 + 2 * t;
      *reinterpret_cast<float2*>(fb + r0 * L::XS_LD + col) = make_float2(acc2[j][0], acc2[j][1]);
      *reinterpret_cast<float2*>(fb + (r0 + 8) * L::XS_LD + col) = make_float2(acc2[j][2], acc2[j][3]);
    }
    flush_rows<C / 2, false>(reinterpret_cast<const uint32_t*>(fb) + rb * 16 * L::XS_LD + chalf * (C / 2), L::XS_LD,
                             p.dh2 + (row0 + rb * 16) * C + chalf * (C / 2), (long long)C * 4);
    __syncthreads();
    ln_bwd_slab<C, Pr>(fb, dxs, p.x1, p.mean_b, p.rstd_b, gam + C, nullptr, p.dx1_out, row0);
    __syncthreads();
    // ---- dzp = dx1 o mask of the projection's dropout; dy = dzp . Wproj
    masked_copy_slab<C, Pr>(dxs, abuf0, p.dzp, row0, p.resid_p, mmfn_drop_seed(p.seed_p));
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    gemm_block<C, Pr>(loader, total_tiles, smem + BL::WS, abuf0, L::A_LD, acc, ti);
    E* dy_g = reinterpret_cast<E*>(p.dy);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int col = chalf * (C / 2) + j * 8 + 2 * t;
      Pr::st_smem(abuf1, L::A_LD, r0, col, acc[j][0], acc[j][1]);      // (the da buffer is free: staging for the write-out)
      Pr::st_smem(abuf1, L::A_LD, r0 + 8, col, acc[j][2], acc[j][3]);
    }
    flush_rows<NWE, false>(abuf1 + rb * 16 * L::A_LD + chalf * NWE, L::A_LD, dy_g + (row0 + rb * 16) * C + chalf * (C / 2), (long long)C * sizeof(E));
  }
  cp_async_wait_all();
}

template <int C, class Pr>
int launch_gpt_bwd_rows(const GptBwdParams& p, long long M, cudaStream_t stream) {
  const int smem = BLay<C, Pr>::WORDS * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(gpt_small_bwd_rows_kernel<C, Pr>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (ce != cudaSuccess) { mmfn_set_error("gpt_small_bwd_rows: shared memory attribute (%d B): %s", smem, cudaGetErrorString(ce)); return (int)ce; }
    attr_set = true;
  }
  gpt_small_bwd_rows_kernel<C, Pr><<<(unsigned)(M / GS_ROWS), GS_THREADS, smem, stream>>>(p);
  return mmfn_launch_status("gpt_small_bwd_rows");
}

// (R, Cc) row-major -> (Cc, R), batched over blockIdx.z = (block, matrix); 32 x 32 tiles through shared memory
template <class E>
__global__ void gpt_transpose_kernel(const void* const* __restrict__ tab, E* __restrict__ out, int C) {
  __shared__ E tile[32][33];
  const int layer = blockIdx.z >> 2, m = blockIdx.z & 3;
  const int R = m == 0 ? 3 * C : (m == 2 ? 4 * C : C), Cc = m == 3 ? 4 * C : C;
  const long long off = (long long)layer * 12 * C * C + (m == 0 ? 0 : m == 1 ? 3 : m == 2 ? 4 : 8) * (long long)C * C;
  const int tiles_c = Cc / 32, ntiles = (R / 32) * tiles_c;
  const E* src = reinterpret_cast<const E*>(tab[layer * 12 + m]);
  for (int tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
    const int tr = tl / tiles_c, tc = tl - tr * tiles_c;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) tile[i][threadIdx.x] = src[(long long)(tr * 32 + i) * Cc + tc * 32 + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) out[off + (long long)(tc * 32 + i) * R + tr * 32 + threadIdx.x] = tile[threadIdx.x][i];
    __syncthreads();
  }
}

}  // namespace

MMFN_DEFINE_RNG_BINDER(gpt_small)

static unsigned long long* g_gpt_trace = nullptr;
// Debug aid for tools/gpt_bench.py: device buffer of >= 10 * n_layer uint64 that receives %globaltimer stamps of the
// first CTA at the phase boundaries of every block (null switches tracing off).
MMFN_API int mmfn_gpt_small_trace(void* buf) { g_gpt_trace = (unsigned long long*)buf; return 0; }

// All n_layer pre-LN transformer blocks of one fusion GPT (model_rad.py:112-133 x n_layer, called from :236) in one
// launch, for n_embd C in {64, 128} (128: bf16 only), 4 heads, T in {128, 192} tokens.  dtype MMFN_BF16: bf16 operands (weights = the
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
  // n_embd 128 as TF32 does not fit: fp32 operands double every operand buffer and leave room for only two weight tiles
  // in flight (measured 2.3x SLOWER than the per-op chain, profiles/r02_gpt_bench_v1.json)
  MMFN_CHECK_ARG(bf || C == 64, "gpt_small_fwd: n_embd 128 needs dtype MMFN_BF16");
  if (C == 64 && T == 128) return bf ? launch_gpt_fwd<64, 2, PrecBF>(p, stream) : launch_gpt_fwd<64, 2, PrecTF>(p, stream);
  if (C == 64 && T == 192) return bf ? launch_gpt_fwd<64, 3, PrecBF>(p, stream) : launch_gpt_fwd<64, 3, PrecTF>(p, stream);
  if (T == 128) return launch_gpt_fwd<128, 2, PrecBF>(p, stream);
  return launch_gpt_fwd<128, 3, PrecBF>(p, stream);
}

// Transposed operand-typed copies of one fusion GPT's linear weights for mmfn_gpt_small_bwd_rows: per block
// [Wqkv^T (C,3C) | Wproj^T (C,C) | Wfc1^T (C,4C) | Wfc2^T (4C,C)] = 12 C^2 elements, blocks back to back in `out`.
// tab: DEVICE array of n_layer x 12 pointers laid out as mmfn_gpt_small_fwd's table (only the four weights are read).
MMFN_API int mmfn_gpt_small_transpose(const void* const* tab_dev, int n_layer, int C, int dtype, void* out, cudaStream_t stream) {
  MMFN_CHECK_ARG(tab_dev && out && n_layer >= 1 && n_layer <= GS_MAXL && (C == 64 || C == 128) && (dtype == 1 || dtype == 2),
                 "gpt_small_transpose: bad arguments");
  const dim3 grid(16, 1, n_layer * 4), block(32, 8);
  if (dtype == 2) gpt_transpose_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(tab_dev, (__nv_bfloat16*)out, C);
  else gpt_transpose_kernel<float><<<grid, block, 0, stream>>>(tab_dev, (float*)out, C);
  return mmfn_launch_status("gpt_small_transpose");
}

// Row-local backward between two attention backwards of a narrow fusion GPT (see the kernel comment): part A finishes
// block `hi` (dqkv -> dh1 -> dx through LayerNorm1), part B starts block `lo` = hi - 1 (MLP and LayerNorm2 backward,
// projection data gradient).  has_a = 0: the chain starts at the GPT's last block from dx2_in; has_b = 0: it ends at
// block 0 and writes dx_out.  wT_hi / wT_lo: that block's slice of mmfn_gpt_small_transpose's output.  M = B*T rows
// (multiple of 64).  Outputs (caller-allocated): dh1 (M,C) fp32, dz / dzp / dy (M,C) and da (M,4C) operand-typed,
// dh2 / dx1_out (M,C) fp32, dx_out (M,C) fp32.  seed_mlp / seed_proj: block lo's dropout streams (attention seed + 2, + 1).
MMFN_API int mmfn_gpt_small_bwd_rows(int64_t M, int C, int dtype, int has_a, int has_b,
                                     const void* dqkv, const float* dx1_in, const float* x_hi, const float* mean1, const float* rstd1,
                                     const float* gamma1, const void* wT_hi, float* dh1, float* dx_out,
                                     const float* dx2_in, const void* a, const float* x1, const float* mean2, const float* rstd2,
                                     const float* gamma2, const void* wT_lo, void* dz, void* da, float* dh2, float* dx1_out,
                                     void* dzp, void* dy, float resid_p, uint64_t seed_mlp, uint64_t seed_proj, cudaStream_t stream) {
  MMFN_CHECK_ARG((C == 64 || C == 128) && (dtype == 2 || (dtype == 1 && C == 64)) && M > 0 && M % GS_ROWS == 0,
                 "gpt_small_bwd_rows: needs C in {64,128} (128: bf16 only), M a positive multiple of 64");
  MMFN_CHECK_ARG(has_a || has_b, "gpt_small_bwd_rows: nothing to do");
  MMFN_CHECK_ARG(!has_a || (dqkv && dx1_in && x_hi && mean1 && rstd1 && gamma1 && wT_hi && dh1 && (has_b || dx_out)), "gpt_small_bwd_rows: null pointer (part A)");
  MMFN_CHECK_ARG(!has_b || ((has_a || dx2_in) && a && x1 && mean2 && rstd2 && gamma2 && wT_lo && dz && da && dh2 && dx1_out && dzp && dy),
                 "gpt_small_bwd_rows: null pointer (part B)");
  MMFN_CHECK_ARG(resid_p >= 0.f && resid_p < 1.f, "gpt_small_bwd_rows: bad dropout probability");
  const size_t es = dtype == 2 ? 2 : 4;
  GptBwdParams p = {};
  p.has_a = has_a; p.has_b = has_b;
  p.dqkv = dqkv; p.dx1_in = dx1_in; p.xa = x_hi; p.mean_a = mean1; p.rstd_a = rstd1; p.gamma_a = gamma1; p.dh1 = dh1; p.dx_out = dx_out;
  p.wqkvT = wT_hi;
  p.dx2_in = dx2_in; p.a = a; p.x1 = x1; p.mean_b = mean2; p.rstd_b = rstd2; p.gamma_b = gamma2;
  if (has_b) {
    const char* w = (const char*)wT_lo;
    p.wpT = w + (size_t)3 * C * C * es; p.w1T = w + (size_t)4 * C * C * es; p.w2T = w + (size_t)8 * C * C * es;
  }
  p.dz = dz; p.da = da; p.dzp = dzp; p.dy = dy; p.dh2 = dh2; p.dx1_out = dx1_out;
  p.resid_p = resid_p; p.seed_m = seed_mlp; p.seed_p = seed_proj;
  if (C == 64) return dtype == 2 ? launch_gpt_bwd_rows<64, PrecBF>(p, M, stream) : launch_gpt_bwd_rows<64, PrecTF>(p, M, stream);
  return launch_gpt_bwd_rows<128, PrecBF>(p, M, stream);
}
