// Waypoint head: 4-step GRUCell roll-out + Linear(64,2) + cumulative waypoints
// (model_rad.py:679-693), fused L1 loss (run_steps/phase2_train_net.py:104) and AdamW
// (phase2_train_net.py:256,:110; torch.optim.AdamW defaults).
#include "common.cuh"

namespace {

constexpr int HID = 64;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// Block per sample, 192 threads = one per gate row of W_hh, whose 64 weights stay in registers for the whole roll-out
// (the 4 steps are strictly sequential, so the kernel is pure latency: no global weight reads inside the time loop).
// saved: (B, steps, 5, 64) = h_prev, r, z, n, gh_n ; xin: (B, steps, 2)
constexpr int GRU_THREADS = 3 * HID;

__global__ void __launch_bounds__(GRU_THREADS)
gru_head_fwd_kernel(const float* __restrict__ z0, const float* __restrict__ target,
                    const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                    const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                    const float* __restrict__ w_out, const float* __restrict__ b_out,
                    int steps, float* __restrict__ pred, float* __restrict__ saved,
                    float* __restrict__ xin_saved, float* __restrict__ hlast) {
  __shared__ float h[HID], hn[HID], gi_s[3 * HID], gh_s[3 * HID];
  __shared__ float x[2], xin[2];
  const int b = blockIdx.x, row = threadIdx.x;
  float w[HID];
  if ((reinterpret_cast<uintptr_t>(w_hh) & 15) == 0) {
#pragma unroll
    for (int q = 0; q < HID / 4; ++q) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(w_hh + row * HID) + q);
      w[4 * q] = t4.x; w[4 * q + 1] = t4.y; w[4 * q + 2] = t4.z; w[4 * q + 3] = t4.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < HID; ++k) w[k] = __ldg(w_hh + row * HID + k);
  }
  const float wi0 = w_ih[row * 2], wi1 = w_ih[row * 2 + 1], bi = b_ih[row], bh = b_hh[row];
  if (row < HID) h[row] = z0[b * HID + row];
  if (row < 2) x[row] = 0.f;
  __syncthreads();
  for (int t = 0; t < steps; ++t) {
    if (row < 2) { xin[row] = x[row] + target[b * 2 + row]; xin_saved[(b * steps + t) * 2 + row] = xin[row]; }
    __syncthreads();
    float a = bh;
#pragma unroll
    for (int k = 0; k < HID; ++k) a += w[k] * h[k];
    gi_s[row] = wi0 * xin[0] + wi1 * xin[1] + bi;
    gh_s[row] = a;
    __syncthreads();
    if (row < HID) {
      const int j = row;
      const float r = sigmoidf_(gi_s[j] + gh_s[j]);
      const float zg = sigmoidf_(gi_s[HID + j] + gh_s[HID + j]);
      const float n = tanhf(gi_s[2 * HID + j] + r * gh_s[2 * HID + j]);
      const float hnew = (1.f - zg) * n + zg * h[j];
      float* sv = saved + ((int64_t)(b * steps + t) * 5) * HID;
      sv[j] = h[j]; sv[HID + j] = r; sv[2 * HID + j] = zg; sv[3 * HID + j] = n; sv[4 * HID + j] = gh_s[2 * HID + j];
      hn[j] = hnew;
    }
    __syncthreads();
    if (row < HID) h[row] = hn[row];
    if (row >= HID && row < HID + 2) {                     // output layer: two threads of the second warp pair
      const int c = row - HID;
      float d = b_out[c];
      for (int k = 0; k < HID; ++k) d += w_out[c * HID + k] * hn[k];
      x[c] += d;
      pred[(b * steps + t) * 2 + c] = x[c];
    }
    __syncthreads();
  }
  if (row < HID) hlast[b * HID + row] = h[row];
}

// Backward through the roll-out.  Thread = gate row: its row of dW_hh (64 values) and its dW_ih / bias gradients are
// accumulated in registers over the time steps and leave with ONE atomic pass at the end; the W_hh^T dgh product is
// split over the three gates (thread (g, j) sums gate g's 64 rows of column j, coalesced reads).
__global__ void __launch_bounds__(GRU_THREADS)
gru_head_bwd_kernel(const float* __restrict__ dpred, const float* __restrict__ saved,
                    const float* __restrict__ xin_saved, const float* __restrict__ hlast,
                    const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                    const float* __restrict__ w_out, int steps,
                    float* __restrict__ dz0, float* __restrict__ dw_ih, float* __restrict__ dw_hh,
                    float* __restrict__ db_ih, float* __restrict__ db_hh,
                    float* __restrict__ dw_out, float* __restrict__ db_out) {
  __shared__ float dgi[3 * HID], dgh[3 * HID], hprev[HID], part[3][HID];
  __shared__ float gx[2], dxc[2], px[2][GRU_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x, j = tid & (HID - 1), g = tid / HID;
  const int lane = tid & 31, warp = tid >> 5;
  const float wi0 = w_ih[tid * 2], wi1 = w_ih[tid * 2 + 1];
  float dwhh[HID];
#pragma unroll
  for (int k = 0; k < HID; ++k) dwhh[k] = 0.f;
  float dwih0 = 0.f, dwih1 = 0.f, dbih = 0.f, dbhh = 0.f, dwo0 = 0.f, dwo1 = 0.f, dbo = 0.f;
  float dh = 0.f, dhp = 0.f;
  if (tid < 2) dxc[tid] = 0.f;
  __syncthreads();
  for (int t = steps - 1; t >= 0; --t) {
    const float* sv = saved + ((int64_t)(b * steps + t) * 5) * HID;
    float hp = 0.f, r = 0.f, zg = 0.f, n = 0.f, ghn = 0.f, hnew = 0.f;
    if (tid < HID) {
      hp = sv[j]; r = sv[HID + j]; zg = sv[2 * HID + j]; n = sv[3 * HID + j]; ghn = sv[4 * HID + j];
      // h' of this step = h_prev of the next step (or hlast)
      hnew = (t == steps - 1) ? hlast[b * HID + j] : saved[((int64_t)(b * steps + t + 1) * 5) * HID + j];
      hprev[j] = hp;
    }
    if (tid < 2) gx[tid] = dpred[(b * steps + t) * 2 + tid] + dxc[tid];
    __syncthreads();
    if (tid < HID) {
      dwo0 += gx[0] * hnew;
      dwo1 += gx[1] * hnew;
      const float dhn = dh + w_out[j] * gx[0] + w_out[HID + j] * gx[1];
      const float dn = dhn * (1.f - zg), dzg = dhn * (hp - n);
      dhp = dhn * zg;
      const float dan = dn * (1.f - n * n);
      const float dr = dan * ghn;
      const float daz = dzg * zg * (1.f - zg);
      const float dar = dr * r * (1.f - r);
      dgi[j] = dar; dgi[HID + j] = daz; dgi[2 * HID + j] = dan;
      dgh[j] = dar; dgh[HID + j] = daz; dgh[2 * HID + j] = dan * r;
    }
    if (tid < 2) dbo += gx[tid];
    __syncthreads();
    // parameter gradients of row `tid`
    const float xi0 = xin_saved[(b * steps + t) * 2], xi1 = xin_saved[(b * steps + t) * 2 + 1];
    const float gi = dgi[tid], gh = dgh[tid];
    dwih0 += gi * xi0; dwih1 += gi * xi1; dbih += gi; dbhh += gh;
#pragma unroll
    for (int k = 0; k < HID; ++k) dwhh[k] += gh * hprev[k];
    // dh_prev += W_hh^T dgh: gate g's share of column j
    float acc = 0.f;
#pragma unroll 8
    for (int rr = 0; rr < HID; ++rr) acc += __ldg(w_hh + (g * HID + rr) * HID + j) * dgh[g * HID + rr];
    part[g][j] = acc;
    // dx_in = W_ih^T dgi: block reduction over the 192 rows
    float p0 = warp_sum(wi0 * gi), p1 = warp_sum(wi1 * gi);
    if (lane == 0) { px[0][warp] = p0; px[1][warp] = p1; }
    __syncthreads();
    if (tid < HID) dh = dhp + part[0][j] + part[1][j] + part[2][j];
    if (tid < 2) {
      float sacc = 0.f;
      for (int wv = 0; wv < GRU_THREADS / 32; ++wv) sacc += px[tid][wv];
      dxc[tid] = gx[tid] + sacc;
    }
    __syncthreads();
  }
  if (tid < HID) {
    dz0[b * HID + j] = dh;
    atomicAdd(dw_out + j, dwo0);
    atomicAdd(dw_out + HID + j, dwo1);
  }
  if (tid < 2) atomicAdd(db_out + tid, dbo);
  atomicAdd(dw_ih + tid * 2, dwih0);
  atomicAdd(dw_ih + tid * 2 + 1, dwih1);
  atomicAdd(db_ih + tid, dbih);
  atomicAdd(db_hh + tid, dbhh);
#pragma unroll
  for (int k = 0; k < HID; ++k) atomicAdd(dw_hh + tid * HID + k, dwhh[k]);
}

// loss = mean |pred - gt| ; dpred = sign(pred - gt) * gscale / n
__global__ void l1_loss_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int n,
                               float* __restrict__ loss, float* __restrict__ dpred, float gscale) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float d = pred[i] - gt[i];
    s += fabsf(d);
    if (dpred) dpred[i] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * gscale / (float)n;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) *loss = s / (float)n;
}

// state: [0] = step count (as float), [1] = 1-beta1^t, [2] = 1-beta2^t
__global__ void adamw_advance_kernel(float* state, float beta1, float beta2) {
  float t = state[0] + 1.f;
  state[0] = t;
  state[1] = (float)(1.0 - pow((double)beta1, (double)t));
  state[2] = (float)(1.0 - pow((double)beta2, (double)t));
}

__global__ void adamw_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                             float4* __restrict__ v, int64_t n4, float lr, float beta1, float beta2, float eps,
                             float wd, const float* __restrict__ state, float gscale, uint2* __restrict__ p16) {
  float bc1 = state[1], bc2 = state[2];
  float step_size = lr / bc1, rbc2 = 1.0f / sqrtf(bc2), decay = 1.0f - lr * wd;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr = ga[k] * gscale;
      float pk = pa[k] * decay;
      ma[k] = beta1 * ma[k] + (1.f - beta1) * gr;
      va[k] = beta2 * va[k] + (1.f - beta2) * gr * gr;
      float denom = sqrtf(va[k]) * rbc2 + eps;
      pa[k] = pk - step_size * (ma[k] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (p16) p16[i] = mmfn_pack_bf16x4(pp.x, pp.y, pp.z, pp.w);      // bf16 shadow read by the tensor-core kernels
  }
}

}  // namespace

MMFN_API int mmfn_gru_head_fwd(const float* z0, const float* target, const float* w_ih, const float* w_hh,
                               const float* b_ih, const float* b_hh, const float* w_out, const float* b_out,
                               int B, int steps, float* pred, float* saved, float* xin_saved, float* hlast,
                               cudaStream_t stream) {
  MMFN_CHECK_ARG(z0 && target && w_ih && w_hh && b_ih && b_hh && w_out && b_out && pred && saved && xin_saved && hlast,
                 "gru_head_fwd: null pointer");
  MMFN_CHECK_ARG(B > 0 && steps > 0, "gru_head_fwd: bad sizes");
  gru_head_fwd_kernel<<<B, GRU_THREADS, 0, stream>>>(z0, target, w_ih, w_hh, b_ih, b_hh, w_out, b_out, steps, pred, saved, xin_saved, hlast);
  return mmfn_launch_status("gru_head_fwd");
}

// All parameter gradients are accumulated (atomicAdd).
MMFN_API int mmfn_gru_head_bwd(const float* dpred, const float* saved, const float* xin_saved, const float* hlast,
                               const float* w_ih, const float* w_hh, const float* w_out, int B, int steps,
                               float* dz0, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh,
                               float* dw_out, float* db_out, cudaStream_t stream) {
  MMFN_CHECK_ARG(dpred && saved && xin_saved && hlast && w_ih && w_hh && w_out && dz0 && dw_ih && dw_hh && db_ih && db_hh &&
                 dw_out && db_out, "gru_head_bwd: null pointer");
  MMFN_CHECK_ARG(B > 0 && steps > 0, "gru_head_bwd: bad sizes");
  gru_head_bwd_kernel<<<B, GRU_THREADS, 0, stream>>>(dpred, saved, xin_saved, hlast, w_ih, w_hh, w_out, steps, dz0, dw_ih, dw_hh,
                                             db_ih, db_hh, dw_out, db_out);
  return mmfn_launch_status("gru_head_bwd");
}

// loss (1 float, device) = mean|pred-gt| over n; dpred (nullable) = d(loss*gscale)/dpred.
MMFN_API int mmfn_l1_loss(const float* pred, const float* gt, int n, float* loss, float* dpred, float gscale,
                          cudaStream_t stream) {
  MMFN_CHECK_ARG(pred && gt && loss && n > 0, "l1_loss: bad args");
  l1_loss_kernel<<<1, 256, 0, stream>>>(pred, gt, n, loss, dpred, gscale);
  return mmfn_launch_status("l1_loss");
}

// state: 3 device floats {t, 1-beta1^t, 1-beta2^t}; advanced on device so a captured graph replays correctly.
// p_bf16 (nullable): bf16 shadow of p refreshed in the same pass (BASELINE configs[2]).
MMFN_API int mmfn_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, float* state, float grad_scale,
                             void* p_bf16, cudaStream_t stream) {
  MMFN_CHECK_ARG(p && g && m && v && state && n >= 0 && n % 4 == 0, "adamw: n must be a multiple of 4");
  MMFN_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adamw: buffers must be 16B aligned");
  adamw_advance_kernel<<<1, 1, 0, stream>>>(state, beta1, beta2);
  if (n > 0)
    adamw_kernel<<<grid_1d(n / 4, 256), 256, 0, stream>>>((float4*)p, (const float4*)g, (float4*)m, (float4*)v, n / 4, lr,
                                                           beta1, beta2, eps, weight_decay, state, grad_scale, (uint2*)p_bf16);
  return mmfn_launch_status("adamw");
}

// The two halves of mmfn_adamw_step for a bucketed optimizer: advance the step count / bias corrections ONCE per
// step, then apply the update to any number of disjoint parameter ranges (each as soon as its gradients are final).
MMFN_API int mmfn_adamw_advance(float* state, float beta1, float beta2, cudaStream_t stream) {
  MMFN_CHECK_ARG(state, "adamw_advance: null state");
  adamw_advance_kernel<<<1, 1, 0, stream>>>(state, beta1, beta2);
  return mmfn_launch_status("adamw_advance");
}

MMFN_API int mmfn_adamw_apply(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, const float* state, float grad_scale,
                              void* p_bf16, cudaStream_t stream) {
  MMFN_CHECK_ARG(p && g && m && v && state && n >= 0 && n % 4 == 0, "adamw_apply: n must be a multiple of 4");
  MMFN_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adamw_apply: buffers must be 16B aligned");
  if (n > 0)
    adamw_kernel<<<grid_1d(n / 4, 256), 256, 0, stream>>>((float4*)p, (const float4*)g, (float4*)m, (float4*)v, n / 4, lr,
                                                           beta1, beta2, eps, weight_decay, state, grad_scale, (uint2*)p_bf16);
  return mmfn_launch_status("adamw_apply");
}
