// Plane regularisers of the training loop, each as ONE streaming pass that produces the loss term
// and its gradient ("next" rank 3):
//
//   k_tv_plane   TVLoss.forward (utils/tensorf_utils.py:139-158) of one NCHW plane (1, C, H, W) as
//                used by TV_loss_density / TV_loss_app (models/tensorf_keyframe.py:205-231):
//                  loss = weight 2 (tfac sum_h (x[y+1] - x[y])^2 / count_h + sum_w (x[.,x+1] - x[.,x])^2 / count_w)
//                with count_h = C (H-1) W, count_w = C H (W-1), tfac = 3 for time planes (t=True)
//   k_l1_plane   density_L1 (models/tensorf_keyframe.py:188-203): mean |x| (space planes) or
//                mean |1 - x| (time planes)
//
// Both are HBM-bound: one read of the plane (the +-1 neighbours come from L1/L2), one write of
// the gradient; the loss is reduced per block in FP64 and added to a device accumulator.  The
// reference runs ~10 elementwise torch kernels plus their autograd twins per plane.
#include "nvfi_common.cuh"

namespace nvfi {

__device__ __forceinline__ double block_sum_f64(double v, double* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (warp == 0) {
    s = (lane < (int)(blockDim.x >> 5)) ? sm[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  return s;   // valid in thread 0
}

// grid: (ceil(W / 32 / XPT), ceil(H / ROWS), C); block: 32 x 8.  Each thread walks ROWS rows of
// one column: the vertical neighbours stay in registers, the horizontal ones are re-read (L1).
template <int ROWS>
__global__ void __launch_bounds__(256)
    k_tv_plane(const float* __restrict__ x, int H, int W, float a_h, float a_w, double* __restrict__ loss,
               float* __restrict__ grad) {
  __shared__ double sm[8];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int y0 = (blockIdx.y * 8 + threadIdx.y) * ROWS;
  const size_t plane = (size_t)blockIdx.z * H * W;
  const float* xp = x + plane;
  float* gp = grad ? grad + plane : nullptr;
  double acc = 0.0;
  if (col < W && y0 < H) {
    float up = (y0 > 0) ? xp[(size_t)(y0 - 1) * W + col] : 0.f;
    float cur = xp[(size_t)y0 * W + col];
    float sh = 0.f, sw = 0.f;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int y = y0 + r;
      if (y >= H) break;
      const bool has_dn = y + 1 < H;
      const float dn = has_dn ? xp[(size_t)(y + 1) * W + col] : 0.f;
      const float lf = (col > 0) ? xp[(size_t)y * W + col - 1] : 0.f;
      const float rt = (col + 1 < W) ? xp[(size_t)y * W + col + 1] : 0.f;
      const float dh = has_dn ? dn - cur : 0.f;                 // owned by (y, col)
      const float dh_up = (y > 0) ? cur - up : 0.f;
      const float dw = (col + 1 < W) ? rt - cur : 0.f;          // owned by (y, col)
      const float dw_lf = (col > 0) ? cur - lf : 0.f;
      sh = fmaf(dh, dh, sh);
      sw = fmaf(dw, dw, sw);
      if (gp) gp[(size_t)y * W + col] = 2.f * a_h * (dh_up - dh) + 2.f * a_w * (dw_lf - dw);
      up = cur;
      cur = dn;
    }
    acc = (double)a_h * (double)sh + (double)a_w * (double)sw;
  }
  const int tid = threadIdx.y * 32 + threadIdx.x;
  // block_sum_f64 indexes by threadIdx.x only: flatten
  double v = acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((tid & 31) == 0) sm[tid >> 5] = v;
  __syncthreads();
  if (tid < 32) {
    double s = (tid < 8) ? sm[tid] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (tid == 0 && s != 0.0) atomicAdd(loss, s);
  }
}

// mean |x - off| over n elements, gradient sign(x - off) / n (0 at the kink, as torch.abs)
__global__ void __launch_bounds__(256)
    k_l1_plane(const float* __restrict__ x, long long n, float off, float inv_n, double* __restrict__ loss,
               float* __restrict__ grad) {
  __shared__ double sm[8];
  double acc = 0.0;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float d[4] = {v.x - off, v.y - off, v.z - off, v.w - off};
    float g[4];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s += fabsf(d[k]);
      g[k] = (d[k] > 0.f) ? inv_n : ((d[k] < 0.f) ? -inv_n : 0.f);
    }
    acc += (double)s;
    if (grad) reinterpret_cast<float4*>(grad)[i] = make_float4(g[0], g[1], g[2], g[3]);
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = x[i] - off;
    acc += (double)fabsf(d);
    if (grad) grad[i] = (d > 0.f) ? inv_n : ((d < 0.f) ? -inv_n : 0.f);
  }
  const double s = block_sum_f64(acc * (double)inv_n, sm);
  if (threadIdx.x == 0 && s != 0.0) atomicAdd(loss, s);
}

}  // namespace nvfi

using namespace nvfi;

extern "C" int nvfi_tv_loss(const float* plane, int C, int H, int W, int time_plane, float scale,
                            double* loss_accum, float* grad, void* stream) {
  if (!plane || !loss_accum || C <= 0 || H <= 0 || W <= 0) return NVFI_EINVAL;
  if (H < 2 || W < 2) return NVFI_EUNSUPPORTED;   // the reference divides by count_h / count_w = 0
  const double count_h = (double)C * (H - 1) * W, count_w = (double)C * H * (W - 1);
  const float a_h = (float)(2.0 * scale * (time_plane ? 3.0 : 1.0) / count_h);
  const float a_w = (float)(2.0 * scale / count_w);
  constexpr int ROWS = 8;
  const dim3 grid((W + 31) / 32, (H + 8 * ROWS - 1) / (8 * ROWS), C), block(32, 8);
  NVFI_LAUNCH(k_tv_plane<ROWS>, grid, block, 0, (cudaStream_t)stream, plane, H, W, a_h, a_w, loss_accum, grad);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_l1_loss(const float* plane, int64_t n, float offset, float scale, double* loss_accum,
                            float* grad, void* stream) {
  if (!plane || !loss_accum || n <= 0) return NVFI_EINVAL;
  if ((reinterpret_cast<uintptr_t>(plane) & 15) || (grad && (reinterpret_cast<uintptr_t>(grad) & 15)))
    return NVFI_EINVAL;
  const int blocks = (int)((n / 4 + 255) / 256 < 148 * 8 ? (n / 4 + 255) / 256 + 1 : 148 * 8);
  NVFI_LAUNCH(k_l1_plane, blocks, 256, 0, (cudaStream_t)stream, plane, (long long)n, offset,
              (float)((double)scale / (double)n), loss_accum, grad);
  return (int)cudaGetLastError();
}
