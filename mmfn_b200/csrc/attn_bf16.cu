// Fused bf16 attention forward for the fusion transformers (SelfAttention.forward, model_rad.py:96-105, under the
// bf16 configuration -- BASELINE configs[2]):
//   y[b, t, h*hs:(h+1)*hs] = dropout(softmax(q k^T / sqrt(hs))) v          for every (batch, head)
// one CTA per (128 query rows, head, batch), tcgen05 kind::f16 with fp32 accumulation in TMEM.
//
// What changed against the TF32 kernel (attn_tc.cu), and why (profiles/r01_ncu_full_kernels.json: 10 % tensor pipe):
//   * Q tile, ALL keys and ALL values of the (batch, head) are requested by TMA at kernel start and stay resident in
//     shared memory (bf16: 32 + 64 + 64 KB at T = 256, hs = 128) -- the TF32 kernel streamed 32-key V boxes through a
//     two-stage ring and exposed one L2 round trip per stage on its serial chain;
//   * the softmax runs on 16 warps, FOUR threads per query row (interleaved 32-key chunks), so every SM sub-partition
//     holds four softmax warps and MUFU / TMEM-load / integer-hash latencies overlap (two threads per row measured
//     21 us at B = 32: 34 % issue utilisation, stalled on latency -- profiles/r02_attn_ncu_v2.txt);
//   * the scores are read from TMEM ONCE and exp2 is evaluated ONCE per score (chunk-local maxima, rescaled when the
//     row maximum is known): the unnormalised probabilities wait in registers as packed bf16 until the row sum is
//     known, then are normalised, written as the swizzled K-major A operand of the PV MMA (one 64-key tile per
//     mbarrier, so the MMA of tile j overlaps the normalisation of tile j + 1) and -- only when the caller wants them
//     for the unfused backward -- stored to HBM as bf16 straight from registers (64 contiguous bytes per thread);
//   * the row statistics (max, sum) are saved so that a backward pass can recompute P instead of loading it.
// T in {128, 192, 256} (multiples of 64), hs in {16, 32, 64, 128}: ragged heads rely on TMA zero-fill.
#include "tc_kernel.cuh"

namespace {

constexpr int AB_THREADS = 576;                 // warp 0 TMA, warp 1 MMA, warps 2-17 softmax / epilogue (4 threads per row)

struct AttnBf16Params {
  int T, nh, C;
  float scale_log2;                             // log2(e) / sqrt(hs)
  float drop_p;
  uint64_t seed;
  int o_col, tmem_cols;
  __nv_bfloat16* y;                             // (B*T, C)
  __nv_bfloat16* P;                             // (B, nh, T, T) softmax probabilities, or null
  __nv_bfloat16* Pd;                            // after dropout (null when drop_p == 0 or P is null)
  float2* stats;                                // (B, nh, T) {row max of scale_log2 * s, row sum of exp2}, or null
  unsigned long long* trace;                    // developer aid: %globaltimer stamps of CTA 0, or null
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float ex2(float x) {               // MUFU.EX2, flush-to-zero: one instruction
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#define AT_STAMP(i) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.trace[i] = tc::gtimer(); } while (0)

// Thread (row, part), part = 0..3, owns the 32-key chunks c = part, part + 4 of its query row (NCH = chunks per thread).
// One pass over TMEM: per chunk a LOCAL maximum m_c, e = exp2(s - m_c) packed to bf16, local sum; the four threads of
// a row exchange (max, sum) once through shared memory; chunk c is then rescaled by exp2(m_c - M) / total while it is
// written out -- the scores are read from TMEM once (TMEM reads run at 64 B/clk per SM: a second pass over the
// 128 x 256 fp32 tile costs 2048 cycles, as much as both MMAs of the tile).
template <int HS, int NJB>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_fwd_bf16_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, AttnBf16Params p) {
  constexpr int NKB = (HS + 63) / 64;           // 64-element (128-byte) blocks of the head dimension
  constexpr int T = NJB * 64;
  constexpr int NC = T / 32;                    // 32-key chunks of a row
  constexpr int NCH = (NC + 3) / 4;             // chunks per thread
  constexpr int Q_BYTES = NKB * 16384, K_BYTES = NKB * T * 128, V_BYTES = NJB * NKB * 8192, P_BYTES = NJB * 16384;
  constexpr int NO = HS < 16 ? 16 : HS;         // N of the PV MMA
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                           // [NKB][128 rows][128 B]; after S is complete: row-statistics exchange
  uint8_t* sK = sQ + Q_BYTES;                   // [NKB][T rows][128 B]
  uint8_t* sV = sK + K_BYTES;                   // [NJB][NKB] boxes of [64 keys][64 d]
  uint8_t* sP = sV + V_BYTES;                   // [NJB][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* qk_full = bars;
  uint64_t* v_full = bars + 1;
  uint64_t* s_full = bars + 2;
  uint64_t* o_full = bars + 3;
  uint64_t* p_full = bars + 4;                  // [NJB]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + NJB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  if (threadIdx.x == 0) AT_STAMP(0);

  if (warp == 0 && lane == 0) { tc::prefetch_tmap(&tmQ); tc::prefetch_tmap(&tmK); tc::prefetch_tmap(&tmV); }
  if (warp == 1 && lane == 0) {
    tc::mbar_init(qk_full, 1); tc::mbar_init(v_full, 1); tc::mbar_init(s_full, 1); tc::mbar_init(o_full, 1);
    for (int j = 0; j < NJB; ++j) tc::mbar_init(&p_full[j], 256);     // two 32-key chunks x 128 rows per 64-key tile
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_o = tmem + (uint32_t)p.o_col;
  if (threadIdx.x == 0) AT_STAMP(1);

  if (warp == 0) {
    if (tc::elect_one()) {                       // ===== TMA producer: everything up front =====
      tc::mbar_expect_tx(qk_full, Q_BYTES + K_BYTES);
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        tc::tma_load_4d(sQ + kb * 16384, &tmQ, qk_full, kb * 64, h, m0, b);
        tc::tma_load_4d(sK + kb * T * 128, &tmK, qk_full, kb * 64, h, 0, b);
      }
      tc::mbar_expect_tx(v_full, V_BYTES);
#pragma unroll
      for (int jb = 0; jb < NJB; ++jb)
#pragma unroll
        for (int db = 0; db < NKB; ++db)
          tc::tma_load_4d(sV + (jb * NKB + db) * 8192, &tmV, v_full, db * 64, h, jb * 64, b);
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {                       // ===== MMA issuer =====
      tc::mbar_wait(qk_full, 0);
      tc::tc_fence_after();
      AT_STAMP(2);
      const uint32_t idesc_s = tc::idesc_bf16(128, T, false, false);
      const uint32_t q0 = tc::smem_u32(sQ), k0 = tc::smem_u32(sK);
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        const int nk = (HS - kb * 64) >= 64 ? 4 : (HS - kb * 64 + 15) / 16;
        for (int k = 0; k < nk; ++k)
          tc::mma_bf16(tmem, tc::smem_desc_kmajor(q0 + kb * 16384 + k * 32), tc::smem_desc_kmajor(k0 + kb * T * 128 + k * 32),
                       idesc_s, (kb | k) ? 1u : 0u);
      }
      tc::mma_commit(s_full);
      tc::mbar_wait(v_full, 0);
      const uint32_t idesc_o = tc::idesc_bf16(128, NO, false, true);
      const uint32_t p0 = tc::smem_u32(sP), v0 = tc::smem_u32(sV);
#pragma unroll 1
      for (int jb = 0; jb < NJB; ++jb) {
        tc::mbar_wait(&p_full[jb], 0);
        tc::tc_fence_after();
        if (jb == 0) AT_STAMP(6);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::mma_bf16(tmem_o, tc::smem_desc_kmajor(p0 + jb * 16384 + k * 32),
                       tc::smem_desc_mnmajor16(v0 + jb * NKB * 8192 + k * 2048, 8192), idesc_o, (jb | k) ? 1u : 0u);
      }
      tc::mma_commit(o_full);
      AT_STAMP(7);
    }
  } else {
    // ===== softmax: four threads per query row =====
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;                       // 0..3
    const int row = q * 32 + lane;                          // row inside the tile == TMEM lane
    const bool row_ok = m0 + row < T;
    const uint32_t t_row = tmem + ((uint32_t)(q * 32) << 16);
    float2* xch = reinterpret_cast<float2*>(sQ);            // [4][128] {max, sum}  (Q is dead once s_full fired)
    tc::mbar_wait(s_full, 0);
    tc::tc_fence_after();
    if (threadIdx.x == 64) AT_STAMP(3);
    float e[NCH][32];                                       // exp2(s - chunk maximum), fp32 until the row sum is known
    float mc[NCH];
    float mt = -INFINITY, st = 0.f;                         // this thread's running (max, sum)
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = part + 4 * i;
      mc[i] = -INFINITY;
      if (c < NC) {                                         // warp-uniform
        tc::tmem_ld32(t_row + (uint32_t)(c * 32), e[i]);
        float m = e[i][0];
#pragma unroll
        for (int j = 1; j < 32; ++j) m = fmaxf(m, e[i][j]);
        m *= p.scale_log2;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          e[i][j] = ex2(fmaf(e[i][j], p.scale_log2, -m));
          e[i][j + 1] = ex2(fmaf(e[i][j + 1], p.scale_log2, -m));
          s0 += e[i][j]; s1 += e[i][j + 1];
        }
        mc[i] = m;
        const float nm = fmaxf(mt, m);
        st = st * ex2(mt - nm) + (s0 + s1) * ex2(m - nm);
        mt = nm;
      }
    }
    xch[part * 128 + row] = make_float2(mt, st);
    asm volatile("bar.sync 1, 512;" ::: "memory");
    float M = mt;
#pragma unroll
    for (int o = 1; o < 4; ++o) M = fmaxf(M, xch[((part + o) & 3) * 128 + row].x);
    float total = 0.f;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const float2 t = xch[o * 128 + row];
      total += t.y * ex2(t.x - M);
    }
    const float inv = 1.0f / total;
    if (p.stats && part == 0 && row_ok) p.stats[((int64_t)b * p.nh + h) * T + m0 + row] = make_float2(M, total);
    if (threadIdx.x == 64) AT_STAMP(4);
    // normalise, (store), dropout, A-operand tiles of the PV MMA.  One fp32 -> bf16 pack per stored / consumed value
    // (F2FP is a quarter-rate instruction: with the MUFU it is what bounds this phase).
    const int64_t prow = (((int64_t)b * p.nh + h) * T + (m0 + row)) * T;   // linear index of P[b,h,i,0] (dropout hash key)
    const bool drop = p.drop_p > 0.f;
    const uint32_t thr = mmfn_drop_threshold(p.drop_p);
    const float keep = 1.0f / (1.0f - p.drop_p);
    const uint64_t dseed = mmfn_drop_seed(p.seed);
    const uint32_t tile_row = tc::smem_u32(sP) + row * 128;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = part + 4 * i;
      if (c < NC) {
        const int col0 = c * 32, jb = c >> 1, hf = c & 1;
        const float f = ex2(mc[i] - M) * inv;               // chunk-local maximum -> row maximum, and 1 / sum
#pragma unroll
        for (int j = 0; j < 32; ++j) e[i][j] *= f;
        uint32_t w[16];
        if (p.P && row_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = pack2(e[i][2 * j], e[i][2 * j + 1]);
          uint4* dst = reinterpret_cast<uint4*>(p.P + prow + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        }
        if (drop) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {                 // one hash per four keys (prow % 4 == 0, col0 % 32 == 0)
            const uint64_t hsh = mmfn_hash64(dseed, (uint64_t)(prow + col0 + j) >> 2);
            const uint32_t lo = (uint32_t)hsh, hi = (uint32_t)(hsh >> 32);
            e[i][j] = (lo & 0xFFFFu) >= thr ? e[i][j] * keep : 0.f;
            e[i][j + 1] = (lo >> 16) >= thr ? e[i][j + 1] * keep : 0.f;
            e[i][j + 2] = (hi & 0xFFFFu) >= thr ? e[i][j + 2] * keep : 0.f;
            e[i][j + 3] = (hi >> 16) >= thr ? e[i][j + 3] * keep : 0.f;
          }
        }
        if (drop || !(p.P && row_ok)) {                     // (without dropout the stored P tile is the MMA operand)
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = pack2(e[i][2 * j], e[i][2 * j + 1]);
        }
        if (drop && p.Pd && row_ok) {
          uint4* dst = reinterpret_cast<uint4*>(p.Pd + prow + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        }
        // K-major SWIZZLE_128B tile: row r at r*128 bytes, 16-byte chunk k stored at position k ^ (r & 7);
        // this thread owns chunks hf*4 .. hf*4 + 3 (its 32 keys) of the 64-key tile jb
#pragma unroll
        for (int k = 0; k < 4; ++k)
          sts128u(tile_row + jb * 16384 + ((((hf << 2) + k) ^ (row & 7)) << 4), w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
        tc::fence_async_smem();                             // generic-proxy writes -> visible to the UMMA
        tc::mbar_arrive(&p_full[jb]);
      }
    }
    if (threadIdx.x == 64) AT_STAMP(5);
    // O: TMEM -> bf16 -> HBM straight from registers: the four threads of a row take the 32-column chunks d = part, ...
    tc::mbar_wait(o_full, 0);
    tc::tc_fence_after();
    if (threadIdx.x == 64) AT_STAMP(8);
#pragma unroll
    for (int d = 0; d < (HS + 31) / 32; ++d) {
      if ((d & 3) == part) {                                // warp-uniform
        float v[32];
        tc::tmem_ld32(tmem_o + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * 32), v);
        if (row_ok) {
          __nv_bfloat16* dst = p.y + ((int64_t)b * T + m0 + row) * p.C + h * HS + d * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            if (d * 32 + j < HS)
              *reinterpret_cast<uint4*>(dst + j) = make_uint4(pack2(v[j], v[j + 1]), pack2(v[j + 2], v[j + 3]),
                                                              pack2(v[j + 4], v[j + 5]), pack2(v[j + 6], v[j + 7]));
        }
      }
    }
    if (threadIdx.x == 64) AT_STAMP(9);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem, p.tmem_cols);
}

// rank-4 view (inner = head dim, heads, tokens, batch) of a bf16 (B*T, pitch) activation matrix
int head_tmap16(CUtensorMap* m, const void* base, int B, int T, int nh, int hs, int64_t pitch, int rows_box) {
  uint64_t dims[4] = {(uint64_t)hs, (uint64_t)nh, (uint64_t)T, (uint64_t)B};
  uint64_t strides[4] = {1, (uint64_t)hs, (uint64_t)pitch, (uint64_t)pitch * T};
  uint32_t box[4] = {64, 1, (uint32_t)rows_box, 1};
  return mmfn_make_tmap_bf16(m, base, 4, dims, strides, box, nullptr);
}

template <int HS, int NJB>
int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnBf16Params& p, int B,
                cudaStream_t stream) {
  constexpr int NKB = (HS + 63) / 64, T = NJB * 64;
  constexpr int SMEM = NKB * 16384 + NKB * T * 128 + NJB * NKB * 8192 + NJB * 16384 + 256 + 1024;
  static_assert(SMEM <= 232448, "attention tile does not fit shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(attn_fwd_bf16_kernel<HS, NJB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (ce != cudaSuccess) { mmfn_set_error("attention_fwd_bf16: smem attribute: %s", cudaGetErrorString(ce)); return (int)ce; }
    attr_set = true;
  }
  dim3 grid((T + 127) / 128, p.nh, B);
  attn_fwd_bf16_kernel<HS, NJB><<<grid, AB_THREADS, SMEM, stream>>>(tq, tk, tv, p);
  return mmfn_launch_status("attention_fwd_bf16");
}

// ds = scale * p * (dp - sum(dp * p)), dp = dpd * dropout_scale; p bf16 (4 per lane and step), dpd fp32, ds bf16
template <int J4>
__global__ void softmax_bwd_bf16_kernel(const uint2* __restrict__ p, const float4* __restrict__ dpd, uint2* __restrict__ ds,
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
      pv[j] = mmfn_unpack_bf16x4(__ldg(p + i));
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
      ds[row * cols4 + c] = mmfn_pack_bf16x4(scale * pv[j].x * (g[j].x - dot), scale * pv[j].y * (g[j].y - dot),
                                             scale * pv[j].z * (g[j].z - dot), scale * pv[j].w * (g[j].w - dot));
  }
}

}  // namespace

// qkv: (B*T, 3C) BF16 fused projections, columns [key | query | value] (model_rad.py:96-98); y: (B*T, C) BF16;
// prob / prob_drop: (B, nh, T, T) BF16 softmax probabilities before / after dropout, saved for the unfused backward --
// both may be null (inference, or a backward that recomputes them from `stats`); prob_drop is required iff prob is given
// and drop_p > 0.  stats (nullable): (B, nh, T) float2 {row max of log2(e)/sqrt(hs) * s, row sum of exp2}.
// T in {128, 192, 256}, head size C / nh in {16, 32, 64, 128}.
MMFN_API int mmfn_attention_fwd_bf16(const void* qkv, void* y, void* prob, void* prob_drop, float* stats,
                                     int B, int T, int C, int nh, float drop_p, uint64_t seed, cudaStream_t stream) {
  MMFN_CHECK_ARG(qkv && y, "attention_fwd_bf16: null pointer");
  MMFN_CHECK_ARG(!prob || drop_p <= 0.f || prob_drop, "attention_fwd_bf16: prob_drop is required with dropout");
  MMFN_CHECK_ARG(B > 0 && nh > 0 && C % nh == 0, "attention_fwd_bf16: bad sizes");
  const int hs = C / nh;
  MMFN_CHECK_ARG(T == 128 || T == 192 || T == 256, "attention_fwd_bf16: T must be 128, 192 or 256");
  MMFN_CHECK_ARG(hs == 16 || hs == 32 || hs == 64 || hs == 128, "attention_fwd_bf16: head size must be 16, 32, 64 or 128");
  MMFN_CHECK_ARG((((uintptr_t)qkv | (uintptr_t)y | (uintptr_t)prob | (uintptr_t)prob_drop) & 15) == 0 && ((uintptr_t)stats & 7) == 0,
                 "attention_fwd_bf16: 16-byte alignment");
  const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(qkv);
  CUtensorMap tq, tk, tv;
  if (int rc = head_tmap16(&tk, base, B, T, nh, hs, 3 * C, T)) return rc;
  if (int rc = head_tmap16(&tq, base + C, B, T, nh, hs, 3 * C, 128)) return rc;
  if (int rc = head_tmap16(&tv, base + 2 * C, B, T, nh, hs, 3 * C, 64)) return rc;
  AttnBf16Params p;
  p.T = T; p.nh = nh; p.C = C;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)hs);
  p.drop_p = drop_p; p.seed = seed;
  p.o_col = T;
  const int need = T + (hs < 32 ? 32 : hs);
  p.tmem_cols = need <= 256 ? 256 : 512;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.P = static_cast<__nv_bfloat16*>(prob);
  p.Pd = (prob && drop_p > 0.f) ? static_cast<__nv_bfloat16*>(prob_drop) : nullptr;
  p.stats = reinterpret_cast<float2*>(stats);
  p.trace = mmfn_tc_trace_ptr();
#define MMFN_ATTN_CASE(HS_)                                                                     \
  case HS_:                                                                                     \
    if (T == 128) return launch_attn<HS_, 2>(tq, tk, tv, p, B, stream);                         \
    if (T == 192) return launch_attn<HS_, 3>(tq, tk, tv, p, B, stream);                         \
    return launch_attn<HS_, 4>(tq, tk, tv, p, B, stream);
  switch (hs) {
    MMFN_ATTN_CASE(16) MMFN_ATTN_CASE(32) MMFN_ATTN_CASE(64) MMFN_ATTN_CASE(128)
  }
#undef MMFN_ATTN_CASE
  return MMFN_BAD_ARG;
}

// Softmax backward of the bf16 configuration: p (BF16 probabilities saved by mmfn_attention_fwd_bf16), dpd (fp32
// gradient of the dropped probabilities), ds (BF16, operand of the dQ / dK GEMMs).  cols % 4 == 0, cols <= 256.
MMFN_API int mmfn_softmax_bwd_bf16(const void* p, const float* dpd, void* ds, int64_t rows, int cols, float scale,
                                   float drop_p, uint64_t seed, cudaStream_t stream) {
  MMFN_CHECK_ARG(p && dpd && ds && rows >= 0 && cols > 0 && cols <= 256 && cols % 4 == 0, "softmax_bwd_bf16: cols must be a multiple of 4, <= 256");
  MMFN_CHECK_ARG((((uintptr_t)p | (uintptr_t)ds) & 7) == 0 && ((uintptr_t)dpd & 15) == 0, "softmax_bwd_bf16: alignment");
  if (rows == 0) return 0;
  softmax_bwd_bf16_kernel<2><<<(unsigned)ceil_div64(rows, 8), 256, 0, stream>>>((const uint2*)p, (const float4*)dpd, (uint2*)ds,
                                                                              rows, cols / 4, scale, drop_p, seed);
  return mmfn_launch_status("softmax_bwd_bf16");
}

MMFN_DEFINE_RNG_BINDER(attn_bf16)
