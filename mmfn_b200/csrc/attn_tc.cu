// Fused attention forward for the fusion transformers (SelfAttention.forward, model_rad.py:96-105):
//   y[b, t, h*hs:(h+1)*hs] = dropout(softmax(q k^T / sqrt(hs))) v     for every (batch, head)
// on the tcgen05 tensor cores, one CTA per (128 query rows, head, batch):
//
//   phase 1  S = Q K^T      : TMA streams 128-byte k-blocks of Q (128 rows) and K (all T keys) straight out of the
//                             fused qkv buffer; tcgen05.mma (N = T <= 256) accumulates the 128 x T score tile in TMEM.
//   softmax                 : four warps (one thread per query row) read S from TMEM twice: online max/sum, then
//                             normalised probabilities, written 32 keys at a time as a swizzled K-major smem tile
//                             (and, with dropout, a second masked tile).
//   phase 2  O = P V        : that tile is the A operand of the second MMA (B = V, MN-major TMA boxes), O accumulates
//                             in TMEM next to S.  The same tiles are TMA-stored to HBM as the saved probabilities the
//                             backward pass needs -- S itself never leaves the SM.
//   epilogue                : O from TMEM -> smem -> TMA store into the (B*T, C) head slice.
//
// T <= 256 keys fit one tile (T = 192 / 256 here, SURVEY.md section 5 "no sequence parallelism needed");
// hs in {16, 32, 64, 128}: ragged 16-wide heads rely on TMA zero-fill / store clipping.
#include "tc_kernel.cuh"

namespace {

constexpr int AT_THREADS = 192;                 // warp 0 TMA, warp 1 MMA, warps 2-5 softmax/epilogue
constexpr int STAGE = 48 * 1024;                // phase 1: Q 16K + K 32K; phase 2: Pd 16K + V 16K + P 16K
constexpr int AT_SMEM = 2 * STAGE + 256 + 1024;

struct AttnParams {
  int T, hs, nh;
  float scale_log2;                             // log2(e) / sqrt(hs)
  float drop_p;
  uint64_t seed;
  int o_col;                                    // TMEM column of the O accumulator
  int tmem_cols;
};

__device__ __forceinline__ void sts_swizzled_row(uint8_t* tile, int row, const float* v) {
  // K-major SWIZZLE_128B tile: row r at r*128 bytes, 16-byte chunk c stored at position c ^ (r & 7)
#pragma unroll
  for (int c = 0; c < 8; ++c)
    tc::sts128(tc::smem_u32(tile) + row * 128 + ((c ^ (row & 7)) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

__global__ void __launch_bounds__(AT_THREADS)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmP,
                const __grid_constant__ CUtensorMap tmPd, const __grid_constant__ CUtensorMap tmY, AttnParams p,
                __nv_bfloat16* __restrict__ y16, int C) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGE);
  uint64_t* qk_full = bars;            // [2]
  uint64_t* qk_empty = bars + 2;       // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* v_full = bars + 5;         // [2]
  uint64_t* pv_empty = bars + 7;       // [2]
  uint64_t* p_full = bars + 9;         // [2]
  uint64_t* o_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int T = p.T, hs = p.hs;
  const int nkc = (hs + 31) / 32;               // k-blocks of the score GEMM
  const int njb = T / 32;                       // key blocks of the PV GEMM
  const bool drop = p.drop_p > 0.f;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmQ); tc::prefetch_tmap(&tmK); tc::prefetch_tmap(&tmV);
    tc::prefetch_tmap(&tmP); tc::prefetch_tmap(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&qk_full[s], 1); tc::mbar_init(&qk_empty[s], 1);
      tc::mbar_init(&v_full[s], 1); tc::mbar_init(&pv_empty[s], 1); tc::mbar_init(&p_full[s], 128);
    }
    tc::mbar_init(s_full, 1); tc::mbar_init(o_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_o = tmem + (uint32_t)p.o_col;

  if (warp == 0) {
    if (tc::elect_one()) {                       // ===== TMA producer =====
      for (int kc = 0; kc < nkc; ++kc) {
        int s = kc & 1; uint32_t ph = (kc >> 1) & 1;
        tc::mbar_wait(&qk_empty[s], ph ^ 1);
        uint8_t* st = smem + s * STAGE;
        tc::mbar_expect_tx(&qk_full[s], 128 * 128 + T * 128);
        tc::tma_load_4d(st, &tmQ, &qk_full[s], kc * 32, h, m0, b);
        tc::tma_load_4d(st + 16384, &tmK, &qk_full[s], kc * 32, h, 0, b);
      }
      tc::mbar_wait(s_full, 0);                  // phase-1 buffers are free once S is complete
      for (int jb = 0; jb < njb; ++jb) {
        int s = jb & 1; uint32_t ph = (jb >> 1) & 1;
        tc::mbar_wait(&pv_empty[s], ph ^ 1);
        uint8_t* vt = smem + s * STAGE + 16384;
        tc::mbar_expect_tx(&v_full[s], nkc * tc::BOX_BYTES);
        for (int d = 0; d < nkc; ++d) tc::tma_load_4d(vt + d * tc::BOX_BYTES, &tmV, &v_full[s], d * 32, h, jb * 32, b);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {                       // ===== MMA issuer =====
      const uint32_t idesc_s = tc::idesc_tf32(128, T, false, false);
      for (int kc = 0; kc < nkc; ++kc) {
        int s = kc & 1; uint32_t ph = (kc >> 1) & 1;
        tc::mbar_wait(&qk_full[s], ph);
        tc::tc_fence_after();
        const uint32_t sq = tc::smem_u32(smem + s * STAGE), sk = sq + 16384;
        const int nk = min(4, (hs - kc * 32 + 7) / 8);
        for (int k = 0; k < nk; ++k)
          tc::mma_tf32(tmem, tc::smem_desc_kmajor(sq + k * 32), tc::smem_desc_kmajor(sk + k * 32), idesc_s, (kc | k) ? 1u : 0u);
        tc::mma_commit(&qk_empty[s]);
      }
      tc::mma_commit(s_full);
      const int n_o = hs < 16 ? 16 : hs;
      const uint32_t idesc_o = tc::idesc_tf32(128, n_o, false, true);
      for (int jb = 0; jb < njb; ++jb) {
        int s = jb & 1; uint32_t ph = (jb >> 1) & 1;
        tc::mbar_wait(&v_full[s], ph);
        tc::mbar_wait(&p_full[s], ph);
        tc::tc_fence_after();
        const uint32_t sp = tc::smem_u32(smem + s * STAGE), sv = sp + 16384;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::mma_tf32(tmem_o, tc::smem_desc_kmajor(sp + k * 32), tc::smem_desc_mnmajor(sv + k * 1024, tc::BOX_BYTES), idesc_o,
                       (jb | k) ? 1u : 0u);
        tc::mma_commit(&pv_empty[s]);
      }
      tc::mma_commit(o_full);
    }
  } else {
    // ===== softmax + epilogue: thread = query row =====
    const int q = warp & 3;
    const int row = q * 32 + lane;                        // row inside the tile == TMEM lane
    const uint32_t t_row = tmem + ((uint32_t)(q * 32) << 16);
    const int64_t prow = (((int64_t)b * p.nh + h) * T + (m0 + row)) * T;   // linear index of P[b,h,i,0] (dropout hash key)
    tc::mbar_wait(s_full, 0);
    tc::tc_fence_after();
    float mx = -INFINITY, sum = 0.f;
    for (int c = 0; c < njb; ++c) {
      float v[32];
      tc::tmem_ld32(t_row + (uint32_t)(c * 32), v);
      float cm = v[0];
#pragma unroll
      for (int j = 1; j < 32; ++j) cm = fmaxf(cm, v[j]);
      cm *= p.scale_log2;
      float nm = fmaxf(mx, cm);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += exp2f(v[j] * p.scale_log2 - nm);
      sum = sum * exp2f(mx - nm) + acc;
      mx = nm;
    }
    const float inv = 1.0f / sum;
    for (int jb = 0; jb < njb; ++jb) {
      int s = jb & 1; uint32_t ph = (jb >> 1) & 1;
      float v[32];
      tc::tmem_ld32(t_row + (uint32_t)(jb * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = exp2f(v[j] * p.scale_log2 - mx) * inv;
      // the stage's tiles are free once the MMA that read them two blocks ago retired and their TMA stores were read
      tc::mbar_wait(&pv_empty[s], ph ^ 1);
      if (threadIdx.x == 64) tc::tma_store_wait_read<1>();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      uint8_t* pd_tile = smem + s * STAGE;
      uint8_t* p_tile = pd_tile + 32768;
      if (drop) {
        sts_swizzled_row(p_tile, row, v);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {                     // prow % 32 == 0: one hash per four keys
          float ds[4];
          mmfn_dropout_scale4(p.drop_p, p.seed, (uint64_t)(prow + jb * 32 + j), ds);
          v[j] *= ds[0]; v[j + 1] *= ds[1]; v[j + 2] *= ds[2]; v[j + 3] *= ds[3];
        }
      }
      sts_swizzled_row(pd_tile, row, v);
      tc::fence_async_smem();                              // generic-proxy writes -> visible to UMMA / TMA
      tc::mbar_arrive(&p_full[s]);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        if (drop) {
          tc::tma_store_4d(p_tile, &tmP, jb * 32, m0, h, b);
          tc::tma_store_4d(pd_tile, &tmPd, jb * 32, m0, h, b);
        } else {
          tc::tma_store_4d(pd_tile, &tmP, jb * 32, m0, h, b);
        }
        tc::tma_store_commit();
      }
    }
    // O: TMEM -> swizzled smem tiles (stage 0 is idle: its last reader was the MMA of block njb-2) -> TMA store
    tc::mbar_wait(o_full, 0);
    tc::tc_fence_after();
    if (threadIdx.x == 64) tc::tma_store_wait_read<0>();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (y16) {
      // bf16 attention output (it only feeds the projection GEMM): each thread owns one row -- 32 columns = 64
      // contiguous bytes = two full 32-byte sectors per chunk, written straight from registers
      for (int d = 0; d < nkc; ++d) {
        float v[32];
        tc::tmem_ld32(tmem_o + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * 32), v);
        if (m0 + row < T) {
          __nv_bfloat16* dst = y16 + ((int64_t)b * T + m0 + row) * C + h * hs + d * 32;
          const int nv = min(32, hs - d * 32);
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            if (j < nv) {
              const uint2 lo = mmfn_pack_bf16x4(v[j], v[j + 1], v[j + 2], v[j + 3]), hi = mmfn_pack_bf16x4(v[j + 4], v[j + 5], v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(dst + j) = make_uint4(lo.x, lo.y, hi.x, hi.y);
            }
        }
      }
    } else {
      for (int d = 0; d < nkc; ++d) {
        float v[32];
        tc::tmem_ld32(tmem_o + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * 32), v);
        sts_swizzled_row(smem + d * 16384, row, v);
      }
      tc::fence_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        for (int d = 0; d < nkc; ++d) tc::tma_store_4d(smem + d * 16384, &tmY, d * 32, h, m0, b);
        tc::tma_store_commit();
        tc::tma_store_wait_read<0>();
      }
    }
    if (threadIdx.x == 64) tc::tma_store_wait_read<0>();      // outstanding P tile stores must have read their smem
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem, p.tmem_cols);
}

__device__ __forceinline__ void lds_swizzled_row(const uint8_t* tile, int row, float* v) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 t = tc::lds128(tc::smem_u32(tile) + row * 128 + ((c ^ (row & 7)) << 4));
    v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
  }
}

// Fused attention BACKWARD, critical half: for one (128 query rows, head, batch) tile
//   dPd = dY V^T                 (tcgen05, 128 x T tile in TMEM -- never written to HBM)
//   dS  = scale * P o (dPd o mask/(1-p) - delta),   delta_i = sum_d dY[i,d] Y[i,d]  (= sum_j dP_ij P_ij)
//   dQ  = dS K                   (tcgen05; dS tiles are the A operand AND are TMA-stored for the dK GEMM)
// The same skeleton as attn_fwd_kernel with (Q, K, softmax, V) -> (dY, V, dS, K): P tiles are TMA-loaded next to the
// K boxes, the dS tile is written by the row threads into the slot the forward kernel uses for the dropped
// probabilities.  dK = dS^T Q and dV = Pd^T dY stay batched GEMMs on the fork streams (they are not on the chain).
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmdY, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmP,
                   const __grid_constant__ CUtensorMap tmdS, const __grid_constant__ CUtensorMap tmdQ,
                   const float* __restrict__ dy, const float* __restrict__ y, int C, float scale, AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGE);
  uint64_t* qk_full = bars;            // [2]  phase 1: dY / V k-blocks landed
  uint64_t* qk_empty = bars + 2;       // [2]
  uint64_t* s_full = bars + 4;         //      dPd complete in TMEM
  uint64_t* v_full = bars + 5;         // [2]  phase 2: P tile + K boxes landed
  uint64_t* pv_empty = bars + 7;       // [2]  MMA of the stage retired
  uint64_t* p_full = bars + 9;         // [2]  dS tile written by the 128 row threads
  uint64_t* o_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int T = p.T, hs = p.hs;
  const int nkc = (hs + 31) / 32;
  const int njb = T / 32;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmdY); tc::prefetch_tmap(&tmV); tc::prefetch_tmap(&tmK);
    tc::prefetch_tmap(&tmP); tc::prefetch_tmap(&tmdS); tc::prefetch_tmap(&tmdQ);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&qk_full[s], 1); tc::mbar_init(&qk_empty[s], 1);
      tc::mbar_init(&v_full[s], 1); tc::mbar_init(&pv_empty[s], 1); tc::mbar_init(&p_full[s], 128);
    }
    tc::mbar_init(s_full, 1); tc::mbar_init(o_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_o = tmem + (uint32_t)p.o_col;

  if (warp == 0) {
    if (tc::elect_one()) {                       // ===== TMA producer =====
      for (int kc = 0; kc < nkc; ++kc) {
        int s = kc & 1; uint32_t ph = (kc >> 1) & 1;
        tc::mbar_wait(&qk_empty[s], ph ^ 1);
        uint8_t* st = smem + s * STAGE;
        tc::mbar_expect_tx(&qk_full[s], 128 * 128 + T * 128);
        tc::tma_load_4d(st, &tmdY, &qk_full[s], kc * 32, h, m0, b);
        tc::tma_load_4d(st + 16384, &tmV, &qk_full[s], kc * 32, h, 0, b);
      }
      tc::mbar_wait(s_full, 0);                  // phase-1 buffers are free once dPd is complete
      for (int jb = 0; jb < njb; ++jb) {
        int s = jb & 1; uint32_t ph = (jb >> 1) & 1;
        tc::mbar_wait(&pv_empty[s], ph ^ 1);     // K boxes and P tile of this stage were consumed two blocks ago
        uint8_t* kt = smem + s * STAGE + 16384;
        tc::mbar_expect_tx(&v_full[s], 16384 + nkc * tc::BOX_BYTES);
        tc::tma_load_4d(smem + s * STAGE + 32768, &tmP, &v_full[s], jb * 32, m0, h, b);
        for (int d = 0; d < nkc; ++d) tc::tma_load_4d(kt + d * tc::BOX_BYTES, &tmK, &v_full[s], d * 32, h, jb * 32, b);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {                       // ===== MMA issuer =====
      const uint32_t idesc_s = tc::idesc_tf32(128, T, false, false);
      for (int kc = 0; kc < nkc; ++kc) {
        int s = kc & 1; uint32_t ph = (kc >> 1) & 1;
        tc::mbar_wait(&qk_full[s], ph);
        tc::tc_fence_after();
        const uint32_t sq = tc::smem_u32(smem + s * STAGE), sk = sq + 16384;
        const int nk = min(4, (hs - kc * 32 + 7) / 8);
        for (int k = 0; k < nk; ++k)
          tc::mma_tf32(tmem, tc::smem_desc_kmajor(sq + k * 32), tc::smem_desc_kmajor(sk + k * 32), idesc_s, (kc | k) ? 1u : 0u);
        tc::mma_commit(&qk_empty[s]);
      }
      tc::mma_commit(s_full);
      const int n_o = hs < 16 ? 16 : hs;
      const uint32_t idesc_o = tc::idesc_tf32(128, n_o, false, true);
      for (int jb = 0; jb < njb; ++jb) {
        int s = jb & 1; uint32_t ph = (jb >> 1) & 1;
        tc::mbar_wait(&v_full[s], ph);
        tc::mbar_wait(&p_full[s], ph);
        tc::tc_fence_after();
        const uint32_t sp = tc::smem_u32(smem + s * STAGE), sv = sp + 16384;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::mma_tf32(tmem_o, tc::smem_desc_kmajor(sp + k * 32), tc::smem_desc_mnmajor(sv + k * 1024, tc::BOX_BYTES), idesc_o,
                       (jb | k) ? 1u : 0u);
        tc::mma_commit(&pv_empty[s]);
      }
      tc::mma_commit(o_full);
    }
  } else {
    // ===== dS + epilogue: thread = query row =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t t_row = tmem + ((uint32_t)(q * 32) << 16);
    const int64_t prow = (((int64_t)b * p.nh + h) * T + (m0 + row)) * T;   // linear index of P[b,h,i,0] (dropout hash key)
    float delta = 0.f;
    if (m0 + row < T) {
      const float4* a = reinterpret_cast<const float4*>(dy + ((int64_t)b * T + m0 + row) * C + h * hs);
      const float4* c = reinterpret_cast<const float4*>(y + ((int64_t)b * T + m0 + row) * C + h * hs);
      for (int d = 0; d < hs / 4; ++d) {
        const float4 u = __ldg(a + d), w = __ldg(c + d);
        delta += u.x * w.x + u.y * w.y + u.z * w.z + u.w * w.w;
      }
    }
    tc::mbar_wait(s_full, 0);
    tc::tc_fence_after();
    for (int jb = 0; jb < njb; ++jb) {
      int s = jb & 1; uint32_t ph = (jb >> 1) & 1;
      float g[32], pv[32];
      tc::tmem_ld32(t_row + (uint32_t)(jb * 32), g);
      tc::mbar_wait(&v_full[s], ph);                       // P tile of this key block landed
      lds_swizzled_row(smem + s * STAGE + 32768, row, pv);
      if (p.drop_p > 0.f) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float ds[4];
          mmfn_dropout_scale4(p.drop_p, p.seed, (uint64_t)(prow + jb * 32 + j), ds);
          g[j] *= ds[0]; g[j + 1] *= ds[1]; g[j + 2] *= ds[2]; g[j + 3] *= ds[3];
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) g[j] = scale * pv[j] * (g[j] - delta);
      // the dS tile of this stage is free once the MMA that read it two blocks ago retired and its TMA store was read
      tc::mbar_wait(&pv_empty[s], ph ^ 1);
      if (threadIdx.x == 64) tc::tma_store_wait_read<1>();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      uint8_t* ds_tile = smem + s * STAGE;
      sts_swizzled_row(ds_tile, row, g);
      tc::fence_async_smem();
      tc::mbar_arrive(&p_full[s]);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        tc::tma_store_4d(ds_tile, &tmdS, jb * 32, m0, h, b);
        tc::tma_store_commit();
      }
    }
    // dQ: TMEM -> swizzled smem tiles -> TMA store into the query slice of dqkv
    tc::mbar_wait(o_full, 0);
    tc::tc_fence_after();
    if (threadIdx.x == 64) tc::tma_store_wait_read<0>();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int d = 0; d < nkc; ++d) {
      float v[32];
      tc::tmem_ld32(tmem_o + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * 32), v);
      sts_swizzled_row(smem + d * 16384, row, v);
    }
    tc::fence_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 64) {
      for (int d = 0; d < nkc; ++d) tc::tma_store_4d(smem + d * 16384, &tmdQ, d * 32, h, m0, b);
      tc::tma_store_commit();
      tc::tma_store_wait_read<0>();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem, p.tmem_cols);
}

// rank-4 view (inner = head dim, heads, tokens, batch) of a (B*T, pitch) activation matrix
int head_tmap(CUtensorMap* m, const float* base, int B, int T, int nh, int hs, int64_t pitch, int rows_box, bool swz32, bool tf32) {
  uint64_t dims[4] = {(uint64_t)hs, (uint64_t)nh, (uint64_t)T, (uint64_t)B};
  uint64_t strides[4] = {1, (uint64_t)hs, (uint64_t)pitch, (uint64_t)pitch * T};
  uint32_t box[4] = {32, 1, (uint32_t)rows_box, 1};
  return mmfn_make_tmap_f32(m, base, 4, dims, strides, box, nullptr, swz32, tf32);
}

}  // namespace

// qkv: (B*T, 3C) fused projections, columns [key | query | value] (model_rad.py:96-98); y: (B*T, C);
// prob: (B, nh, T, T) softmax probabilities saved for backward; prob_drop: same after dropout (required iff
// drop_p > 0, else may be null).  T in {64, 128, 192, 256}, head size C/nh in {16, 32, 64, 128}.
// y_bf16 != 0: y is a BF16 tensor (BASELINE configs[2]: the attention output only feeds the bf16 projection GEMM).
MMFN_API int mmfn_attention_fwd_tf32(const float* qkv, void* y, int y_bf16, float* prob, float* prob_drop,
                                     int B, int T, int C, int nh, float drop_p, uint64_t seed, cudaStream_t stream) {
  MMFN_CHECK_ARG(qkv && y && prob, "attention_fwd: null pointer");
  MMFN_CHECK_ARG(drop_p <= 0.f || prob_drop, "attention_fwd: prob_drop is required with dropout");
  MMFN_CHECK_ARG(B > 0 && nh > 0 && C % nh == 0, "attention_fwd: bad sizes");
  const int hs = C / nh;
  MMFN_CHECK_ARG(T % 32 == 0 && T >= 32 && T <= 256 && T % 16 == 0, "attention_fwd: T must be a multiple of 32, <= 256");
  MMFN_CHECK_ARG(hs == 16 || hs == 32 || hs == 64 || hs == 128, "attention_fwd: head size must be 16, 32, 64 or 128");
  MMFN_CHECK_ARG((((uintptr_t)qkv | (uintptr_t)y | (uintptr_t)prob | (uintptr_t)prob_drop) & 15) == 0, "attention_fwd: 16-byte alignment");
  CUtensorMap tq, tk, tv, tp, tpd, ty;
  if (int rc = head_tmap(&tk, qkv, B, T, nh, hs, 3 * C, T, false, true)) return rc;
  if (int rc = head_tmap(&tq, qkv + C, B, T, nh, hs, 3 * C, 128, false, true)) return rc;
  if (int rc = head_tmap(&tv, qkv + 2 * C, B, T, nh, hs, 3 * C, 32, true, true)) return rc;
  if (int rc = head_tmap(&ty, y_bf16 ? qkv : static_cast<const float*>(y), B, T, nh, hs, y_bf16 ? 3 * C : C, 128, false, false)) return rc;   // unused when y_bf16
  {
    uint64_t dims[4] = {(uint64_t)T, (uint64_t)T, (uint64_t)nh, (uint64_t)B};
    uint64_t strides[4] = {1, (uint64_t)T, (uint64_t)T * T, (uint64_t)T * T * nh};
    uint32_t box[4] = {32, 128, 1, 1};
    if (int rc = mmfn_make_tmap_f32(&tp, prob, 4, dims, strides, box, nullptr, false, false)) return rc;
    if (int rc = mmfn_make_tmap_f32(&tpd, prob_drop ? prob_drop : prob, 4, dims, strides, box, nullptr, false, false)) return rc;
  }
  AttnParams p;
  p.T = T; p.hs = hs; p.nh = nh;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)hs);
  p.drop_p = drop_p; p.seed = seed;
  p.o_col = T <= 128 ? 128 : (T <= 192 ? 192 : 256);
  int need = p.o_col + (hs < 32 ? 32 : hs);
  p.tmem_cols = need <= 256 ? 256 : 512;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (ce != cudaSuccess) { mmfn_set_error("attention_fwd: smem attribute: %s", cudaGetErrorString(ce)); return (int)ce; }
    attr_set = true;
  }
  dim3 grid((T + 127) / 128, nh, B);
  attn_fwd_kernel<<<grid, AT_THREADS, AT_SMEM, stream>>>(tq, tk, tv, tp, tpd, ty, p, y_bf16 ? static_cast<__nv_bfloat16*>(y) : nullptr, C);
  return mmfn_launch_status("attention_fwd");
}

// Critical half of the attention backward (see attn_bwd_dq_kernel).  qkv: (B*T, 3C) [key|query|value] saved by the
// forward; dy, y: (B*T, C) gradient / value of the attention output (before the projection); prob: (B,nh,T,T) saved
// softmax probabilities (pre-dropout).  Writes ds (B,nh,T,T) and the QUERY slice of dqkv (B*T, 3C).
MMFN_API int mmfn_attention_bwd_dq_tf32(const float* qkv, const float* dy, const float* y, const float* prob,
                                        float* ds, float* dqkv, int B, int T, int C, int nh,
                                        float drop_p, uint64_t seed, cudaStream_t stream) {
  MMFN_CHECK_ARG(qkv && dy && y && prob && ds && dqkv, "attention_bwd_dq: null pointer");
  MMFN_CHECK_ARG(B > 0 && nh > 0 && C % nh == 0, "attention_bwd_dq: bad sizes");
  const int hs = C / nh;
  MMFN_CHECK_ARG(T % 32 == 0 && T >= 32 && T <= 256, "attention_bwd_dq: T must be a multiple of 32, <= 256");
  MMFN_CHECK_ARG(hs == 16 || hs == 32 || hs == 64 || hs == 128, "attention_bwd_dq: head size must be 16, 32, 64 or 128");
  MMFN_CHECK_ARG((((uintptr_t)qkv | (uintptr_t)dy | (uintptr_t)y | (uintptr_t)prob | (uintptr_t)ds | (uintptr_t)dqkv) & 15) == 0,
                 "attention_bwd_dq: 16-byte alignment");
  CUtensorMap tdy, tv, tk, tp, tds, tdq;
  if (int rc = head_tmap(&tdy, dy, B, T, nh, hs, C, 128, false, true)) return rc;
  if (int rc = head_tmap(&tv, qkv + 2 * C, B, T, nh, hs, 3 * C, T, false, true)) return rc;
  if (int rc = head_tmap(&tk, qkv, B, T, nh, hs, 3 * C, 32, true, true)) return rc;
  if (int rc = head_tmap(&tdq, dqkv + C, B, T, nh, hs, 3 * C, 128, false, false)) return rc;
  {
    uint64_t dims[4] = {(uint64_t)T, (uint64_t)T, (uint64_t)nh, (uint64_t)B};
    uint64_t strides[4] = {1, (uint64_t)T, (uint64_t)T * T, (uint64_t)T * T * nh};
    uint32_t box[4] = {32, 128, 1, 1};
    if (int rc = mmfn_make_tmap_f32(&tp, prob, 4, dims, strides, box, nullptr, false, false)) return rc;
    if (int rc = mmfn_make_tmap_f32(&tds, ds, 4, dims, strides, box, nullptr, false, false)) return rc;
  }
  AttnParams p;
  p.T = T; p.hs = hs; p.nh = nh;
  p.scale_log2 = 0.f;
  p.drop_p = drop_p; p.seed = seed;
  p.o_col = T <= 128 ? 128 : (T <= 192 ? 192 : 256);
  int need = p.o_col + (hs < 32 ? 32 : hs);
  p.tmem_cols = need <= 256 ? 256 : 512;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (ce != cudaSuccess) { mmfn_set_error("attention_bwd_dq: smem attribute: %s", cudaGetErrorString(ce)); return (int)ce; }
    attr_set = true;
  }
  dim3 grid((T + 127) / 128, nh, B);
  attn_bwd_dq_kernel<<<grid, AT_THREADS, AT_SMEM, stream>>>(tdy, tv, tk, tp, tds, tdq, dy, y, C, 1.0f / sqrtf((float)hs), p);
  return mmfn_launch_status("attention_bwd_dq");
}

MMFN_DEFINE_RNG_BINDER(attn_tc)
