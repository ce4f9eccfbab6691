// Generic warp-specialised tcgen05 pipeline shared by the TF32 GEMM and the implicit-GEMM
// convolutions:
//
//   warp 0   TMA producer    : Op::load() issues cp.async.bulk.tensor boxes for k-block kb
//   warp 1   MMA issuer      : 4 x tcgen05.mma (K = 8 tf32 each) per 128-byte k-block, accumulator in TMEM
//   warps 2-5 epilogue       : tcgen05.ld -> fused epilogue -> HBM, row addressing by Op::out_row()
//
// The shared-memory ring holds STAGES x (A tile 128 rows + B tile TBN rows) x 128 bytes.
// K-major tiles use SWIZZLE_128B; MN-major tiles are stacks of [32 k-rows][32 fp32] boxes in the
// SWIZZLE_128B_ATOM_32B pattern (the only MN-major layout the tensor core accepts for TF32).
#pragma once
#include "tc_common.cuh"

namespace tc {

constexpr int TBM = 128;          // tile rows (UMMA M)
constexpr int TBK = 32;           // fp32 elements per k-block = one 128-byte swizzle span
constexpr int UMMA_K = 8;         // tf32
constexpr int TC_THREADS = 192;
constexpr int BOX_BYTES = 32 * 128;   // one MN-major box: 32 k-rows x 128 B

struct Epilogue {
  float* C;
  const float* bias;   // per output column, or null
  const float* res;    // same indexing as C, or null
  const float* mask;   // same indexing as C: out *= (mask > 0)
  float alpha;
  int act;             // 0 none, 1 relu
  int accum;           // 0 store, 1 +=, 2 atomicAdd
  float drop_p;
  uint64_t drop_seed;
};

template <int TBN, int STAGES>
struct Smem {
  static constexpr int A_BYTES = TBM * 128;
  static constexpr int B_BYTES = TBN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;   // + alignment slack
};

// Op contract (all __device__):
//   static constexpr bool A_MN, B_MN;
//   void setup();                                 per-CTA decode of blockIdx (called by every thread)
//   int kb_begin(), kb_end();                     this CTA's k-block range
//   void load(kb, sa, sb, bar, &tmA, &tmB);       issue the TMA boxes of k-block kb (one thread)
//   bool out_row(r, int64_t& off);                element offset of tile row r, column 0 of the OUTPUT row
//   int n_cols();  int col0();                    valid output columns, first column of this tile
//   bool first_split();                           bias / residual are added by the first split only
template <class Op, int TBN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS)
tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Op op, Epilogue e) {
  using L = Smem<TBN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  op.setup();
  const int kb0 = op.kb_begin(), kb1 = op.kb_end();

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
      const uint32_t idesc = idesc_tf32(TBM, TBN, Op::A_MN, Op::B_MN);
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
        const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < TBK / UMMA_K; ++k) {
          // K-major: 8 tf32 = 32 bytes along the swizzled row; MN-major: the next 8 k-rows = 1024 bytes
          uint64_t ad = Op::A_MN ? smem_desc_mnmajor(sa + k * 1024, BOX_BYTES) : smem_desc_kmajor(sa + k * 32);
          uint64_t bd = Op::B_MN ? smem_desc_mnmajor(sb + k * 1024, BOX_BYTES) : smem_desc_kmajor(sb + k * 32);
          mma_tf32(tmem_base, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        mma_commit(&empty[stage]);                       // frees the smem slot once these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      mma_commit(tmem_full);                             // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5 own TMEM lane quarters (warp % 4) =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int64_t off = 0;
    const bool row_ok = op.out_row(r, off);
    const int N = op.n_cols(), n0 = op.col0();
    const bool first = op.first_split(), have_k = kb1 > kb0;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < TBN / 32; ++c) {
      float v[32];
      __syncwarp();                                      // tcgen05.ld is warp-collective (.sync.aligned)
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      const int col0 = n0 + c * 32;
      if (!row_ok || col0 >= N) continue;                // re-converges at the __syncwarp above
      const int64_t base = off + col0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = have_k ? e.alpha * v[j] : 0.f;
        if (col0 + j < N) {
          if (e.bias && first) x += __ldg(e.bias + col0 + j);
          if (e.act == 1) x = fmaxf(x, 0.f);
          if (e.mask) x = (__ldg(e.mask + base + j) > 0.f) ? x : 0.f;
          if (e.drop_p > 0.f) x *= mmfn_dropout_scale(e.drop_p, e.drop_seed, (uint64_t)(base + j));
          if (e.res && first) x += __ldg(e.res + base + j);
        }
        v[j] = x;
      }
      float* dst = e.C + base;
      if (e.accum == 0 && col0 + 32 <= N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
        for (int j = 0; j < 32 && col0 + j < N; ++j) {
          if (e.accum == 0) dst[j] = v[j];
          else if (e.accum == 1) dst[j] += v[j];
          else atomicAdd(dst + j, v[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TBN);
}

template <class Op, int TBN, int STAGES>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const Op& op, const Epilogue& e, dim3 grid,
                  cudaStream_t stream, const char* what) {
  using L = Smem<TBN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(tc_kernel<Op, TBN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (ce != cudaSuccess) { mmfn_set_error("%s: smem attribute: %s", what, cudaGetErrorString(ce)); return (int)ce; }
    attr_set = true;
  }
  tc_kernel<Op, TBN, STAGES><<<grid, TC_THREADS, L::TOTAL, stream>>>(ta, tb, op, e);
  return mmfn_launch_status(what);
}

}  // namespace tc
