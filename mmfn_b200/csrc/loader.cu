// Loader-side device kernels: the parts of the reference's CPU data path that touch
// every LiDAR point / histogram cell / radar pair, moved to the GPU so that a training batch crosses PCIe in its
// packed form (uint8 histogram counts, no adjacency matrix) and phase-1 preprocessing needs no per-point CPU work.
//   mmfn_lidar_ego_transform_f64   CARLA_Data.__getitem__ (team_code/mmfn_utils/datasets/dataloader.py:229-239): y flip +
//                                  transform_2d_points (:311-334) of a raw sweep, float64 like the reference
//   mmfn_bev_pack_u8 / _unpack_u8  the 2 x 256 x 256 histogram as the per-cell point count min(k, 5) (1 byte) <-> the
//                                  float32 value k / 5 the reference stores (dataloader.py:283-285)
//   mmfn_radar_adjacency_f64       PRE_Data.__getitem__ (dataloader.py:379-384): adj[i, j] = az[j] - az[i] in float64
// All are HBM-bound streaming kernels.
#include "common.cuh"

namespace {

// pose (per frame, 6 doubles): cos r1, sin r1, cos r2, sin r2, t1_x - t2_x, t1_y - t2_y -- evaluated on the host with the
// same numpy calls as mmfn_b200/preprocess.py:transform_points_2d, whose operation order the kernel follows with
// explicitly rounded (non-contracted) float64 operations: the result is bit-identical to that mirror.
__global__ void lidar_ego_transform_kernel(const float* __restrict__ pts, int in_stride, const double* __restrict__ pose,
                                           float* __restrict__ out, int N, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = i / N;
    const double* p = pose + f * 6;
    const double c1 = p[0], s1 = p[1], c2 = p[2], s2 = p[3], dx = p[4], dy = p[5];
    const float* q = pts + i * in_stride;
    const double x = (double)q[0], y = -(double)q[1];                 // dataloader.py:232: the y axis is flipped first
    const double wx = __dadd_rn(__dadd_rn(__dmul_rn(c1, x), __dmul_rn(s1, y)), dx);
    const double wy = __dadd_rn(__dadd_rn(__dmul_rn(-s1, x), __dmul_rn(c1, y)), dy);
    float* o = out + i * 3;
    o[0] = __double2float_rn(__dsub_rn(__dmul_rn(c2, wx), __dmul_rn(s2, wy)));
    o[1] = __double2float_rn(__dadd_rn(__dmul_rn(s2, wx), __dmul_rn(c2, wy)));
    o[2] = q[2];
  }
}

// four cells per thread
__global__ void bev_pack_u8_kernel(const float4* __restrict__ hist, uint32_t* __restrict__ counts, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(hist + i);
    const uint32_t a = (uint32_t)__float2int_rn(v.x * 5.0f), b = (uint32_t)__float2int_rn(v.y * 5.0f),
                   c = (uint32_t)__float2int_rn(v.z * 5.0f), d = (uint32_t)__float2int_rn(v.w * 5.0f);
    counts[i] = (a & 0xffu) | ((b & 0xffu) << 8) | ((c & 0xffu) << 16) | ((d & 0xffu) << 24);
  }
}
__global__ void bev_unpack_u8_kernel(const uint32_t* __restrict__ counts, float4* __restrict__ hist, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t w = __ldg(counts + i);
    // float32(k) * 0.2f == float32(k / 5.0) bit for bit on k = 0..5 (the expression bev.cu writes the histogram with)
    hist[i] = make_float4(__fmul_rn((float)(w & 0xffu), 0.2f), __fmul_rn((float)((w >> 8) & 0xffu), 0.2f),
                          __fmul_rn((float)((w >> 16) & 0xffu), 0.2f), __fmul_rn((float)(w >> 24), 0.2f));
  }
}

__global__ void radar_adjacency_kernel(const double* __restrict__ az, float* __restrict__ adj, int R, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / ((int64_t)R * R);
    const int rem = (int)(i - b * R * R);
    const int r = rem / R, c = rem - r * R;
    adj[i] = __double2float_rn(__dsub_rn(az[b * R + c], az[b * R + r]));
  }
}

}  // namespace

// out (F, N, 3) fp32 = the raw sweeps pts (F, N, in_stride >= 3; x y z [intensity]) with y flipped and moved from frame 1
// to frame 2 (dataloader.py:229-239, :311-334), computed in float64 and rounded once.  pose (F, 6) doubles on the device:
// cos r1, sin r1, cos r2, sin r2, t1_x - t2_x, t1_y - t2_y.  Feeds mmfn_bev_scatter directly.
MMFN_API int mmfn_lidar_ego_transform_f64(const float* pts, int in_stride, const double* pose, float* out, int F, int N,
                                          cudaStream_t stream) {
  MMFN_CHECK_ARG(pts && pose && out, "lidar_ego_transform: null pointer");
  MMFN_CHECK_ARG(F > 0 && N > 0 && in_stride >= 3, "lidar_ego_transform: bad sizes");
  MMFN_CHECK_ARG(((uintptr_t)pose & 7) == 0, "lidar_ego_transform: pose must be 8-byte aligned");
  const int64_t total = (int64_t)F * N;
  lidar_ego_transform_kernel<<<grid_1d(total, 256), 256, 0, stream>>>(pts, in_stride, pose, out, N, total);
  return mmfn_launch_status("lidar_ego_transform_f64");
}

// counts[i] = round(hist[i] * 5): the histogram value k / 5 (k = min(points in the cell, 5)) back to its count, one byte
// per cell (n % 4 == 0).  The packed form of the LiDAR input on disk and across PCIe (4x smaller).
MMFN_API int mmfn_bev_pack_u8(const float* hist, uint8_t* counts, int64_t n, cudaStream_t stream) {
  MMFN_CHECK_ARG(hist && counts && n > 0 && n % 4 == 0, "bev_pack_u8: bad args (n % 4 == 0)");
  MMFN_CHECK_ARG(((uintptr_t)hist & 15) == 0 && ((uintptr_t)counts & 3) == 0, "bev_pack_u8: alignment");
  bev_pack_u8_kernel<<<grid_1d(n / 4, 256), 256, 0, stream>>>((const float4*)hist, (uint32_t*)counts, n / 4);
  return mmfn_launch_status("bev_pack_u8");
}

// hist[i] = float32(counts[i]) * 0.2f: bit-identical to the float32 histogram mmfn_bev_scatter and the reference's
// lidar_to_histogram_features (dataloader.py:271-293) produce for the same counts (n % 4 == 0).
MMFN_API int mmfn_bev_unpack_u8(const uint8_t* counts, float* hist, int64_t n, cudaStream_t stream) {
  MMFN_CHECK_ARG(hist && counts && n > 0 && n % 4 == 0, "bev_unpack_u8: bad args (n % 4 == 0)");
  MMFN_CHECK_ARG(((uintptr_t)hist & 15) == 0 && ((uintptr_t)counts & 3) == 0, "bev_unpack_u8: alignment");
  bev_unpack_u8_kernel<<<grid_1d(n / 4, 256), 256, 0, stream>>>((const uint32_t*)counts, (float4*)hist, n / 4);
  return mmfn_launch_status("bev_unpack_u8");
}

// adj (B, R, R) fp32, adj[b, i, j] = float32(az[b, j] - az[b, i]) with az (B, R) the float64 azimuth column of the radar
// returns: PRE_Data.__getitem__'s "adjacency" (dataloader.py:379-384) followed by Engine.train's float32 cast.
MMFN_API int mmfn_radar_adjacency_f64(const double* az, float* adj, int B, int R, cudaStream_t stream) {
  MMFN_CHECK_ARG(az && adj && B > 0 && R > 0, "radar_adjacency: bad args");
  MMFN_CHECK_ARG(((uintptr_t)az & 7) == 0, "radar_adjacency: az must be 8-byte aligned");
  const int64_t total = (int64_t)B * R * R;
  radar_adjacency_kernel<<<grid_1d(total, 256), 256, 0, stream>>>(az, adj, R, total);
  return mmfn_launch_status("radar_adjacency_f64");
}
