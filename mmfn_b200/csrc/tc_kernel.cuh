// Generic warp-specialised tcgen05 pipeline shared by the TF32 GEMM and the implicit-GEMM
// convolutions:
//
//   warp 0   TMA producer    : Op::load() issues cp.async.bulk.tensor boxes for k-block kb
//   warp 1   MMA issuer      : 4 x tcgen05.mma (K = 8 tf32 each) per 128-byte k-block, accumulator in TMEM
//   warps 2-9 epilogue       : tcgen05.ld -> smem re-tile -> fused epilogue -> HBM (coalesced float4),
//                              row addressing by Op::out_row()
//
// The shared-memory ring holds STAGES x (A tile 128 rows + B tile TBN rows) x 128 bytes.
// K-major tiles use SWIZZLE_128B; MN-major tiles are stacks of [32 k-rows][32 fp32] boxes in the
// SWIZZLE_128B_ATOM_32B pattern (the only MN-major layout the tensor core accepts for TF32).
#pragma once
#include "tc_common.cuh"

namespace tc {

constexpr int TBM = 128;          // tile rows (UMMA M)
constexpr int TBK = 32;           // fp32 elements per k-block = one 128-byte swizzle span (bf16 pipelines: Op::EB = 64)
constexpr int UMMA_K = 8;         // tf32 (bf16: 16); either way a k-block is FOUR MMAs, 32 bytes apart in a K-major row
constexpr int TC_THREADS = 320;        // TMA warp + MMA warp + 8 epilogue warps
constexpr int BOX_BYTES = 32 * 128;   // one MN-major fp32 box: 32 k-rows x 128 B (bf16: ElemTraits<64>::BOX_BYTES)

struct Epilogue {
  float* C;
  const float* bias;   // per output column, or null
  const float* res;    // same indexing as C, or null
  const float* mask;   // same indexing as C: out *= (mask > 0)
  float alpha;
  int act;             // 0 none, 1 relu
  int accum;           // 0 store, != 0 atomic accumulate (split-K / weight gradients)
  float drop_p;
  uint64_t drop_seed;
  unsigned long long* trace;   // optional: 8 x %globaltimer stamps of CTA 0 (mmfn_tc_set_trace), else null
  __nv_bfloat16* C16;                // OUT16 instantiations: the result is written here as bf16 (same indexing as C)
  const __nv_bfloat16* mask16;       // like mask, for a bf16 tensor
  // BatchNorm batch statistics of the OUTPUT (train-mode BN after a convolution), accumulated by the epilogue instead
  // of a separate pass over the tensor: per-channel sum / sum of squares into the fp64 slot scratch of norm.cu; the
  // last CTA folds the slots and publishes mean / rstd (+ running statistics).  Null ws = off.
  double* bn_ws;
  float* bn_mean; float* bn_rstd; float* bn_rmean; float* bn_rvar;
  float bn_eps, bn_momentum;
  long long bn_rows;                 // rows of the (rows, C) output = BatchNorm sample count per channel
};

constexpr int BN_SLOTS = 16;         // must match norm.cu

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP(i) do { if (e.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) e.trace[i] = gtimer(); } while (0)

template <int TBN, int STAGES>
struct Smem {
  static constexpr int A_BYTES = TBM * 128;
  static constexpr int B_BYTES = TBN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 128 + 1024;                          // barriers + alignment slack
  // the epilogue re-uses the (by then idle) operand stages: 8 warps x [32][36] floats + 128 row offsets
  static_assert(8 * 32 * 36 * 4 + 128 * 8 <= STAGES * STAGE_BYTES, "epilogue staging does not fit");
};

// Epilogue shared by every tcgen05 kernel of the library (called by warps 2..9 = 8 epilogue warps; `smem` is the
// 1024-byte aligned dynamic shared memory whose first 8*32*36*4 + 128*8 bytes are idle once tmem_full has fired).
// Persistent kernels (tc_gemm_persist_kernel) run TWO epilogue groups of 8 warps: `ew_` = the warp's index inside its group,
// `parity` = phase of the accumulator barrier for this tile, `bar_id` = the group's named barrier.
template <class Op, int TBN, bool FULL, bool OUT16 = false, bool WIDE = false>
__device__ __forceinline__ void tc_epilogue(const Op& op, const Epilogue& e, uint8_t* smem, uint64_t* tmem_full,
                                            uint32_t tmem_base, int kb0, int kb1, int ew_ = -1, uint32_t parity = 0, int bar_id = 1) {
  // WIDE: 16 epilogue warps (persistent kernel, TBN = 128): every warp owns ONE 32-column chunk of its lane quarter, so
  // the per-tile read-out chain is half as long; staging = 16 x [32][36] floats
  constexpr int NEW = WIDE ? 16 : 8;                    // epilogue warps
  constexpr int CSTEP = NEW / 4;                        // column-chunk stride of a warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // ===== epilogue: 8 warps; warp w owns TMEM lane quarter (w % 4) and every other 32-column chunk =====
    // tcgen05.ld hands each thread one accumulator ROW (32 consecutive columns).  The rows are
    // re-tiled through shared memory (the idle pipeline stages, 36-float pitch: conflict-free for
    // 128-bit accesses) so that each warp instruction touches four fully coalesced 128-byte row
    // segments: result stores / atomics, residual and mask loads all run as 16-byte vectors.
    const int ew = ew_ >= 0 ? ew_ : warp - 2;             // 0..NEW-1
    const int q = warp & 3;
    const uint32_t stg = smem_u32(smem) + ew * (32 * 36 * 4);                      // this warp's [32][36] staging tile
    const uint32_t row_off = smem_u32(smem) + NEW * 32 * 36 * 4;                   // int64 [128]
    int64_t my_off = 0;
    const bool my_ok = op.out_row(q * 32 + lane, my_off);
    const int N = op.n_cols(), n0 = op.col0();
    const bool have_k = kb1 > kb0;
    const float* bias = op.first_split() ? e.bias : nullptr;    // bias / residual are added by the first K split only
    const float* res = op.first_split() ? e.res : nullptr;
    // 16-byte path: every row of this warp's quarter starts on a float4 boundary and the tile is full in N
    const bool vec = (N % 4 == 0) && (OUT16 ? (reinterpret_cast<uintptr_t>(e.C16) & 7) == 0 : (reinterpret_cast<uintptr_t>(e.C) & 15) == 0) &&
                     (n0 + TBN <= N) && __all_sync(0xffffffffu, !my_ok || (my_off & 3) == 0);
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
    // bias vectors of this warp's (up to two) column chunks: requested while the main loop still runs -- read inside the
    // chunk loop, the L2 round trip (~0.7 us) sat in front of every chunk's stores (profiles/r02_tc_trace_gemm.txt)
    float4 b_lo = make_float4(0.f, 0.f, 0.f, 0.f), b_hi = b_lo;
    if (vec && bias) {
      b_lo = __ldg(reinterpret_cast<const float4*>(bias + n0 + (ew >> 2) * 32 + c4));
      if constexpr (TBN / 32 > CSTEP) b_hi = __ldg(reinterpret_cast<const float4*>(bias + n0 + ((ew >> 2) + CSTEP) * 32 + c4));
    }
    mbar_wait(tmem_full, parity);                         // all MMAs retired: TMEM valid, smem stages idle
    tc_fence_after();
    if (ew < 4) sts64(row_off + (q * 32 + lane) * 8, my_ok ? my_off : (int64_t)-1);
    if constexpr (WIDE) asm volatile("bar.sync 1, 512;" ::: "memory");    // row_off visible to the epilogue warps
    else if (bar_id == 1) asm volatile("bar.sync 1, 256;" ::: "memory");
    else asm volatile("bar.sync 2, 256;" ::: "memory");
    if (threadIdx.x == 64) TC_STAMP(4);                   // accumulator visible to the epilogue
#pragma unroll 1
    for (int c = ew >> 2; c < TBN / 32; c += CSTEP) {
      const int col0 = n0 + c * 32;
      if (col0 >= N) break;                               // warp-uniform
      {
        float v[32];
        __syncwarp();                                     // tcgen05.ld is warp-collective (.sync.aligned)
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
        if (threadIdx.x == 64 && c == 0) TC_STAMP(7);       // first accumulator chunk in registers
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          sts128(stg + (lane * 36 + j) * 4, v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      __syncwarp();
      if (threadIdx.x == 64 && c == 0) TC_STAMP(8);         // re-tiled in shared memory
      const int col = col0 + c4;
      if (vec) {
        // ---- fast path: float4 everywhere; kept small on purpose (this code runs once per CTA, from a cold
        // instruction cache -- a fully unrolled, branchy epilogue costs more in fetch stalls than it saves)
        const float4 b4 = c < CSTEP ? b_lo : b_hi;        // (chunks c, c + CSTEP of this warp)
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
        float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};     // BatchNorm partials of this thread's 8 rows
        // per half (4 rows): the rows' offsets, then all residual / mask vectors, are requested before the first use --
        // four independent 16-byte loads in flight per operand instead of one per (dependent) loop iteration.
        // (All eight at once pushes the FULL variant past 102 registers = one CTA per SM.)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          int64_t offs[4];
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) offs[rr] = lds64(row_off + (q * 32 + (hh * 4 + rr) * 4 + rsub) * 8);
          float4 r4[4], m4v[4];
          if (res) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr)
              r4[rr] = offs[rr] >= 0 ? __ldg(reinterpret_cast<const float4*>(res + offs[rr] + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if constexpr (FULL) {
            if (e.mask) {
#pragma unroll
              for (int rr = 0; rr < 4; ++rr)
                m4v[rr] = offs[rr] >= 0 ? __ldg(reinterpret_cast<const float4*>(e.mask + offs[rr] + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else if (e.mask16) {
#pragma unroll
              for (int rr = 0; rr < 4; ++rr) {
                m4v[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (offs[rr] >= 0) {
                  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(e.mask16 + offs[rr] + col));
                  const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
                  const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
                  m4v[rr] = make_float4(lo.x, lo.y, hi.x, hi.y);
                }
              }
            }
          }
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const int r = (hh * 4 + rr) * 4 + rsub;
            if (offs[rr] < 0) continue;
            const int64_t idx = offs[rr] + col;
            const float4 a4 = lds128(stg + (r * 36 + c4) * 4);
            float x[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = (have_k ? e.alpha * x[j] : 0.f) + bb[j];
            if constexpr (FULL) {
              if (e.act == 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = fmaxf(x[j], 0.f);
              }
              if (e.mask || e.mask16) {
                const float mk[4] = {m4v[rr].x, m4v[rr].y, m4v[rr].z, m4v[rr].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = mk[j] > 0.f ? x[j] : 0.f;
              }
              if (e.drop_p > 0.f) {
                float ds[4];
                mmfn_dropout_scale4(e.drop_p, e.drop_seed, (uint64_t)idx, ds);   // idx % 4 == 0 on the vector path
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] *= ds[j];
              }
            }
            if (res) { x[0] += r4[rr].x; x[1] += r4[rr].y; x[2] += r4[rr].z; x[3] += r4[rr].w; }
            if constexpr (!FULL && !OUT16) {
              if (e.bn_ws) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { st1[j] += x[j]; st2[j] += x[j] * x[j]; }
              }
            }
            if constexpr (OUT16) {          // bf16 result (tensors that only feed further MMAs): one 8-byte store
              const __nv_bfloat162 lo = __floats2bfloat162_rn(x[0], x[1]), hi = __floats2bfloat162_rn(x[2], x[3]);
              uint2 pk;
              pk.x = *reinterpret_cast<const uint32_t*>(&lo);
              pk.y = *reinterpret_cast<const uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(e.C16 + idx) = pk;
            } else if (e.accum == 0) *reinterpret_cast<float4*>(e.C + idx) = make_float4(x[0], x[1], x[2], x[3]);
            else  // split-K / weight-gradient accumulation: one 16-byte vector reduction instead of four scalar atomics
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(e.C + idx), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]) : "memory");
          }
        }
        if constexpr (!FULL && !OUT16) {
          if (e.bn_ws) {
            // the four row sub-groups of the warp (lanes l, l+8, l+16, l+24) hold the same 4 columns: fold them, then
            // lanes 0-7 add this warp's 32 rows x 32 columns into the slot of this CTA (fp64 atomics, <= gridDim/16 per address)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              st1[j] += __shfl_xor_sync(0xffffffffu, st1[j], 8);  st2[j] += __shfl_xor_sync(0xffffffffu, st2[j], 8);
              st1[j] += __shfl_xor_sync(0xffffffffu, st1[j], 16); st2[j] += __shfl_xor_sync(0xffffffffu, st2[j], 16);
            }
            if (lane < 8) {
              double* slot = e.bn_ws + (int64_t)((blockIdx.x + blockIdx.y) % BN_SLOTS) * 2 * N;
#pragma unroll
              for (int j = 0; j < 4; ++j) { atomicAdd(slot + col + j, (double)st1[j]); atomicAdd(slot + N + col + j, (double)st2[j]); }
            }
          }
        }
      } else {
        // ---- ragged / unaligned tiles: scalar, not unrolled
#pragma unroll 1
        for (int rr = 0; rr < 8; ++rr) {
          const int r = rr * 4 + rsub;
          const int64_t off_r = lds64(row_off + (q * 32 + r) * 8);
          if (off_r < 0) continue;
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            if (col + j >= N) break;
            const int64_t idx = off_r + col + j;
            float t = (have_k ? e.alpha * lds32(stg + (r * 36 + c4 + j) * 4) : 0.f) + (bias ? __ldg(bias + col + j) : 0.f);
            if constexpr (FULL) {
              if (e.act == 1) t = fmaxf(t, 0.f);
              if (e.mask) t = __ldg(e.mask + idx) > 0.f ? t : 0.f;
              else if (e.mask16) t = __bfloat162float(e.mask16[idx]) > 0.f ? t : 0.f;
              if (e.drop_p > 0.f) t *= mmfn_dropout_scale(e.drop_p, e.drop_seed, (uint64_t)idx);
            }
            if (res) t += __ldg(res + idx);
            if constexpr (OUT16) e.C16[idx] = __float2bfloat16_rn(t);
            else if (e.accum == 0) e.C[idx] = t;
            else atomicAdd(e.C + idx, t);
          }
        }
      }
      if (threadIdx.x == 64 && c == 0) TC_STAMP(9);         // first chunk's stores issued
    }
}

// Last-CTA finalize of the epilogue-accumulated BatchNorm statistics (called by EVERY thread of the CTA after the
// epilogue's stores were issued).  Same arithmetic as bn_colsum_kernel<false> of norm.cu: fp64 fold of the slots, biased
// variance for the normalisation, unbiased for the running estimate; the scratch is left zero for the next call.
__device__ __forceinline__ void tc_bn_finalize(const Epilogue& e, int C, bool* last_flag) {
  if (!e.bn_ws) return;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(e.bn_ws + (int64_t)BN_SLOTS * 2 * C);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *last_flag = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y * gridDim.z - 1;
  __syncthreads();
  if (!*last_flag) return;
  __threadfence();
  const double M = (double)e.bn_rows, invM = 1.0 / M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int sl = 0; sl < BN_SLOTS; ++sl) {
      double* slot = e.bn_ws + (int64_t)sl * 2 * C;
      s1 += __ldcg(slot + c);
      s2 += __ldcg(slot + C + c);
      slot[c] = 0.0;
      slot[C + c] = 0.0;
    }
    const double md = s1 * invM;
    double var = s2 * invM - md * md;
    if (var < 0.0) var = 0.0;
    e.bn_mean[c] = (float)md;
    e.bn_rstd[c] = (float)(1.0 / sqrt(var + (double)e.bn_eps));
    if (e.bn_rmean) {
      const double unbiased = e.bn_rows > 1 ? var * M / (M - 1.0) : var;
      e.bn_rmean[c] = (float)((1.0 - e.bn_momentum) * e.bn_rmean[c] + e.bn_momentum * md);
      e.bn_rvar[c] = (float)((1.0 - e.bn_momentum) * e.bn_rvar[c] + e.bn_momentum * unbiased);
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// Op contract (all __device__):
//   static constexpr bool A_MN, B_MN;
//   void setup();                                 per-CTA decode of blockIdx (called by every thread)
//   int kb_begin(), kb_end();                     this CTA's k-block range
//   void load(kb, sa, sb, bar, &tmA, &tmB);       issue the TMA boxes of k-block kb (one thread)
//   bool out_row(r, int64_t& off);                element offset of tile row r, column 0 of the OUTPUT row
//   int n_cols();  int col0();                    valid output columns, first column of this tile
//   bool first_split();                           bias / residual are added by the first split only
// FULL = false compiles the epilogue down to alpha*acc + bias + residual (most launches); the
// ReLU / mask / dropout variant is a separate instantiation so its hash arithmetic is never if-converted in.
template <class Op, int TBN, int STAGES, bool FULL, bool OUT16 = false>
__global__ void __launch_bounds__(TC_THREADS, 2)
tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Op op, Epilogue e) {
  using L = Smem<TBN, STAGES>;
  using ET = ElemTraits<Op::EB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  op.setup();
  const int kb0 = op.kb_begin(), kb1 = op.kb_end();
  if (threadIdx.x == 0) TC_STAMP(0);                      // kernel entry

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TBN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TC_STAMP(1);                      // barriers + TMEM ready

  if (warp == 0) {
    if (elect_one()) {                                   // ===== TMA producer =====
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * L::STAGE_BYTES;
        mbar_expect_tx(&full[stage], L::STAGE_BYTES);
        op.load(kb, sa, sa + L::A_BYTES, &full[stage], &tmA, &tmB);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {                                   // ===== MMA issuer =====
      const uint32_t idesc = ET::idesc(TBM, TBN, Op::A_MN, Op::B_MN);
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (kb == kb0) TC_STAMP(2);                      // first operands landed
        const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
        const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < TBK / UMMA_K; ++k) {
          // K-major: 8 tf32 / 16 bf16 = 32 bytes along the swizzled row; MN-major: the next 8 / 16 k-rows = 1024 / 2048 bytes
          uint64_t ad = Op::A_MN ? ET::mn_desc(sa + k * ET::MN_STEP) : smem_desc_kmajor(sa + k * 32);
          uint64_t bd = Op::B_MN ? ET::mn_desc(sb + k * ET::MN_STEP) : smem_desc_kmajor(sb + k * 32);
          ET::mma(tmem_base, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        mma_commit(&empty[stage]);                       // frees the smem slot once these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      mma_commit(tmem_full);                             // accumulator complete
      TC_STAMP(3);                                       // all MMAs issued
    }
  } else {
    tc_epilogue<Op, TBN, FULL, OUT16>(op, e, smem, tmem_full, tmem_base, kb0, kb1);
  }
  if (threadIdx.x == 64) TC_STAMP(5);                    // epilogue stores issued
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TBN);
  if (threadIdx.x == 64) TC_STAMP(6);                    // TMEM released
  if constexpr (!FULL && !OUT16) {
    __shared__ bool bn_last;
    tc_bn_finalize(e, op.n_cols(), &bn_last);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent variant for large plain GEMMs (no split-K, no batch): ONE CTA per SM walks the output tiles; the TMA ring
// runs ahead across tile boundaries, the accumulator is double-buffered in TMEM (2 x 128 columns) and SIXTEEN epilogue
// warps (one 32 x 32 chunk of the tile each) read one accumulator out while the MMA warp fills the other, so operand
// streaming never stops for an epilogue.  Why: the one-tile-per-CTA
// kernel keeps the SM's L2 ingest port busy about a third of the time on the K = 512 MLP GEMMs -- set-up, first-TMA
// latency and a 3.5 us epilogue per tile, only partly hidden by the second resident CTA (profiles/r02_tc_trace_gemm.txt).
// Op needs set_tile(int) (tile index -> m0 / n0, n fastest so that neighbouring CTAs share the A rows in L2).
constexpr int TCP_THREADS = 64 + 2 * 256;            // TMA warp + MMA warp + two epilogue groups
#ifndef MMFN_TCP_WIDE
#define MMFN_TCP_WIDE 1
#endif
constexpr bool TCP_WIDE = MMFN_TCP_WIDE != 0;        // 16 warps on every tile | two groups of 8 alternating tiles
constexpr int TCP_GRP_BYTES = (8 * 32 * 36 * 4 + 128 * 8 + 1023) / 1024 * 1024;
constexpr int TCP_STG_BYTES = TCP_WIDE ? 16 * 32 * 36 * 4 + 128 * 8 : 2 * TCP_GRP_BYTES;     // epilogue staging + row offsets

// TBN = 128: 4 stages x 32 KB; TBN = 256 (N % 256 == 0): 3 stages x 48 KB -- 1.33x the MMA work per operand byte, which is
// what bounds these GEMMs (the bytes one CTA can keep in flight against the ~1.5 us L2 round trip under load)
template <int TBN>
struct SmemP {
  static constexpr int STAGES = TBN == 256 ? 3 : 4;
  static constexpr int STAGE_BYTES = TBM * 128 + TBN * 128;
  static constexpr int STG_OFF = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = STG_OFF + (TCP_STG_BYTES + 1023) / 1024 * 1024;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

// developer aid (tools/gemm_persist_bench.py --trace): %globaltimer stamps of CTA 0, 8 per tile --
// 0 producer issues the tile's first k-block, 1 its last; 2 MMA warp has the accumulator, 3 first operands landed, 4 tile committed;
// 5 epilogue sees the accumulator, 6 its stores are issued
static __device__ unsigned long long* g_tcp_trace = nullptr;
#define TCP_STAMP(it, i) do { if (g_tcp_trace && blockIdx.x == 0 && (it) < 16) g_tcp_trace[(it) * 8 + (i)] = gtimer(); } while (0)

template <class Op, int TCP_TBN, bool FULL, bool OUT16>
__global__ void __launch_bounds__(TCP_THREADS, 1)
tc_gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Op op, Epilogue e, int ntiles) {
  using ET = ElemTraits<Op::EB>;
  using SmemP = tc::SmemP<TCP_TBN>;
  constexpr int TCP_STAGES = SmemP::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SmemP::BAR_OFF);
  uint64_t* empty = full + TCP_STAGES;
  uint64_t* acc_full = empty + TCP_STAGES;               // [2]
  uint64_t* acc_empty = acc_full + 2;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TCP_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], TCP_WIDE ? 16 : 8); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 2 * TCP_TBN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nkb = op.kb_total();

  if (warp == 0) {
    if (elect_one()) {                                   // ===== TMA producer: one ring across all of this CTA's tiles =====
      int stage = 0; uint32_t phase = 0;
      int pit = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++pit) {
        op.set_tile(tile);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (kb == 0) TCP_STAMP(pit, 0);
          uint8_t* sa = smem + stage * SmemP::STAGE_BYTES;
          mbar_expect_tx(&full[stage], SmemP::STAGE_BYTES);
          op.load(kb, sa, sa + TBM * 128, &full[stage], &tmA, &tmB);
          if (++stage == TCP_STAGES) { stage = 0; phase ^= 1; }
        }
        TCP_STAMP(pit, 1);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {                                   // ===== MMA issuer =====
      const uint32_t idesc = ET::idesc(TBM, TCP_TBN, Op::A_MN, Op::B_MN);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait(&acc_empty[b], ((it >> 1) & 1) ^ 1);   // this accumulator's previous tile has been read out
        tc_fence_after();
        TCP_STAMP(it, 2);
        const uint32_t acc = tmem_base + (uint32_t)(b * TCP_TBN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (kb == 0) TCP_STAMP(it, 3);
          const uint32_t sa = smem_u32(smem + stage * SmemP::STAGE_BYTES);
          const uint32_t sb = sa + TBM * 128;
#pragma unroll
          for (int k = 0; k < TBK / UMMA_K; ++k) {
            uint64_t ad = Op::A_MN ? ET::mn_desc(sa + k * ET::MN_STEP) : smem_desc_kmajor(sa + k * 32);
            uint64_t bd = Op::B_MN ? ET::mn_desc(sb + k * ET::MN_STEP) : smem_desc_kmajor(sb + k * 32);
            ET::mma(acc, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          mma_commit(&empty[stage]);
          if (++stage == TCP_STAGES) { stage = 0; phase ^= 1; }
        }
        mma_commit(&acc_full[b]);
        TCP_STAMP(it, 4);
      }
    }
  } else {
    if constexpr (TCP_WIDE) {
      // ===== 16 epilogue warps read out every tile (one 32 x 32 chunk each) while the MMA warp fills the other accumulator =====
      const int ew = warp - 2;
      uint8_t* stg = smem + SmemP::STG_OFF;
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        op.set_tile(tile);
        if (threadIdx.x == 64) { mbar_wait(&acc_full[b], (uint32_t)((it >> 1) & 1)); TCP_STAMP(it, 5); }
        tc_epilogue<Op, TCP_TBN, FULL, OUT16, true>(op, e, stg, &acc_full[b], tmem_base + (uint32_t)(b * TCP_TBN), 0, nkb, ew,
                                                    (uint32_t)((it >> 1) & 1), 1);
        tc_fence_before();                               // this warp's tcgen05.ld of the tile are complete (wait::ld inside)
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[b]);
        if (threadIdx.x == 64) TCP_STAMP(it, 6);
      }
    } else {
      // ===== two epilogue groups: group g takes this CTA's tiles g, g + 2, ... (= accumulator g) =====
      const int grp = (warp - 2) >> 3, ew = (warp - 2) & 7;
      uint8_t* stg = smem + SmemP::STG_OFF + grp * TCP_GRP_BYTES;
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        if ((it & 1) != grp) continue;
        op.set_tile(tile);
        tc_epilogue<Op, TCP_TBN, FULL, OUT16, false>(op, e, stg, &acc_full[grp], tmem_base + (uint32_t)(grp * TCP_TBN), 0, nkb, ew,
                                                     (uint32_t)((it >> 1) & 1), 1 + grp);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[grp]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * TCP_TBN);
}

template <class Op, int TBN, bool FULL, bool OUT16>
static int launch_persist_impl(const CUtensorMap& ta, const CUtensorMap& tb, const Op& op, const Epilogue& e, int ntiles, int nsm,
                               cudaStream_t stream, const char* what) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(tc_gemm_persist_kernel<Op, TBN, FULL, OUT16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemP<TBN>::TOTAL);
    if (ce != cudaSuccess) { mmfn_set_error("%s: smem attribute (persistent): %s", what, cudaGetErrorString(ce)); return (int)ce; }
    attr_set = true;
  }
  const int grid = ntiles < nsm ? ntiles : nsm;
  tc_gemm_persist_kernel<Op, TBN, FULL, OUT16><<<grid, TCP_THREADS, SmemP<TBN>::TOTAL, stream>>>(ta, tb, op, e, ntiles);
  return mmfn_launch_status(what);
}

template <class Op, int TBN>
static int launch_persist(const CUtensorMap& ta, const CUtensorMap& tb, const Op& op, const Epilogue& e, int ntiles, int nsm,
                          cudaStream_t stream, const char* what) {
  const bool full = e.act != 0 || e.mask != nullptr || e.mask16 != nullptr || e.drop_p > 0.f;
  if (e.C16 != nullptr) {
    if (e.accum != 0) { mmfn_set_error("%s: a bf16 result cannot be accumulated atomically", what); return MMFN_BAD_ARG; }
    if (full) return launch_persist_impl<Op, TBN, true, true>(ta, tb, op, e, ntiles, nsm, stream, what);
    return launch_persist_impl<Op, TBN, false, true>(ta, tb, op, e, ntiles, nsm, stream, what);
  }
  if (full) return launch_persist_impl<Op, TBN, true, false>(ta, tb, op, e, ntiles, nsm, stream, what);
  return launch_persist_impl<Op, TBN, false, false>(ta, tb, op, e, ntiles, nsm, stream, what);
}

template <class Op, int TBN, int STAGES, bool FULL, bool OUT16 = false>
static int launch_impl(const CUtensorMap& ta, const CUtensorMap& tb, const Op& op, const Epilogue& e, dim3 grid,
                       cudaStream_t stream, const char* what) {
  using L = Smem<TBN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(tc_kernel<Op, TBN, STAGES, FULL, OUT16>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (ce != cudaSuccess) { mmfn_set_error("%s: smem attribute: %s", what, cudaGetErrorString(ce)); return (int)ce; }
    attr_set = true;
  }
  tc_kernel<Op, TBN, STAGES, FULL, OUT16><<<grid, TC_THREADS, L::TOTAL, stream>>>(ta, tb, op, e);
  return mmfn_launch_status(what);
}

// ALLOW16: instantiate the bf16-output epilogues for this Op (GEMMs; the convolutions always write fp32)
template <class Op, int TBN, int STAGES, bool ALLOW16 = false>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const Op& op, const Epilogue& e, dim3 grid,
                  cudaStream_t stream, const char* what) {
  const bool full = e.act != 0 || e.mask != nullptr || e.mask16 != nullptr || e.drop_p > 0.f;
  if (e.C16 != nullptr) {
    if constexpr (ALLOW16) {
      if (e.accum != 0) { mmfn_set_error("%s: a bf16 result cannot be accumulated atomically", what); return MMFN_BAD_ARG; }
      if (full) return launch_impl<Op, TBN, STAGES, true, true>(ta, tb, op, e, grid, stream, what);
      return launch_impl<Op, TBN, STAGES, false, true>(ta, tb, op, e, grid, stream, what);
    } else {
      mmfn_set_error("%s: bf16 output is not available for this kernel", what);
      return MMFN_BAD_ARG;
    }
  }
  if (full) return launch_impl<Op, TBN, STAGES, true>(ta, tb, op, e, grid, stream, what);
  return launch_impl<Op, TBN, STAGES, false>(ta, tb, op, e, grid, stream, what);
}

}  // namespace tc
