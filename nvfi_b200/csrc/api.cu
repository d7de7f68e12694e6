// C-ABI entry points (include/nvfi_b200.h): layout packing, ray generation and the
// forward-render orchestration.  Every function validates its arguments, enqueues
// kernels on the caller's stream and returns an error code; nothing here touches torch.
#include <cstring>

#include <cuda_fp16.h>

#include "nvfi_common.cuh"

extern "C" int nvfi_launch_sample_advect(const NvfiField*, const NvfiRenderArgs*,
                                         const NvfiRenderBuffers*, cudaStream_t);
extern "C" int nvfi_launch_march(const NvfiField*, const NvfiRenderArgs*, const NvfiRenderBuffers*,
                                 cudaStream_t);
extern "C" int nvfi_launch_chunk_inside(const NvfiField*, const NvfiRenderArgs*, const NvfiRenderBuffers*, cudaStream_t);
extern "C" int nvfi_launch_sample_advect_wave(const NvfiField*, const NvfiRenderArgs*, const NvfiRenderBuffers*, int,
                                              int, cudaStream_t);
extern "C" int nvfi_launch_march_wave(const NvfiField*, const NvfiRenderArgs*, const NvfiRenderBuffers*, int, int,
                                      cudaStream_t);
extern "C" int nvfi_launch_appearance(const NvfiField*, const NvfiRenderArgs*,
                                      const NvfiRenderBuffers*, cudaStream_t);
extern "C" int nvfi_launch_composite(const NvfiField*, const NvfiRenderArgs*,
                                     const NvfiRenderBuffers*, cudaStream_t);

namespace nvfi {

// (R, H*W) -> (H*W, R) tiled transpose through shared memory (both sides coalesced).
__global__ void k_transpose(const float* __restrict__ src, float* __restrict__ dst, int rows,
                            long long cols) {
  __shared__ float tile[32][33];
  const long long c0 = (long long)blockIdx.x * 32;
  const int r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j;
    const long long c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[(long long)r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long c = c0 + j;
    const int r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[c * rows + r] = tile[threadIdx.x][j];
  }
}

// Up to 12 (rows, cols) -> (cols, rows) transposes in one launch: blockIdx.z selects the job; blocks
// outside a job's extent exit.
struct TransposeJob {
  const float* src;
  float* dst;
  int rows;
  long long cols;
};
struct TransposeBatch {
  TransposeJob job[12];
};
__global__ void k_transpose_batch(const __grid_constant__ TransposeBatch tb) {
  const TransposeJob& J = tb.job[blockIdx.z];
  if (J.src == nullptr) return;
  __shared__ float tile[32][33];
  const long long c0 = (long long)blockIdx.x * 32;
  const int r0 = blockIdx.y * 32;
  if (c0 >= J.cols || r0 >= J.rows) return;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j;
    const long long c = c0 + threadIdx.x;
    if (r < J.rows && c < J.cols) tile[j][threadIdx.x] = J.src[(long long)r * J.cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long c = c0 + j;
    const int r = r0 + threadIdx.x;
    if (r < J.rows && c < J.cols) J.dst[c * J.rows + r] = tile[threadIdx.x][j];
  }
}

// Up to 12 packed (k_pad, n_pad) [+ (n_pad)] -> nn.Linear (out, in) [+ (out)] in one launch.
struct UnpackJob {
  const float* wt;
  const float* bias_in;
  float* w;
  float* b;
  int out_dim, in_dim, n_pad;
  int v4;   // packed layout [k / 4][n][k % 4] (grad_layout_v4) instead of [k][n]
};
struct UnpackBatch {
  UnpackJob job[12];
};
__global__ void k_unpack_linear_batch(const __grid_constant__ UnpackBatch ub) {
  const UnpackJob& J = ub.job[blockIdx.y];
  if (J.wt == nullptr) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < J.out_dim * J.in_dim) {
    const int n = i / J.in_dim, k = i - n * J.in_dim;
    J.w[i] = J.v4 ? J.wt[((k >> 2) * J.n_pad + n) * 4 + (k & 3)] : J.wt[k * J.n_pad + n];
  }
  if (J.b && J.bias_in && i < J.out_dim) J.b[i] = J.bias_in[i];
}

// nn.Linear (out,in) -> W^T zero-padded (k_pad, n_pad)
__global__ void k_pack_linear(const float* __restrict__ w, const float* __restrict__ b,
                              float* __restrict__ wt, float* __restrict__ bo, int out_dim,
                              int in_dim, int k_pad, int n_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k_pad * n_pad) {
    const int k = i / n_pad, n = i - k * n_pad;
    wt[i] = (k < in_dim && n < out_dim) ? w[n * in_dim + k] : 0.f;
  }
  if (bo && i < n_pad) bo[i] = (b && i < out_dim) ? b[i] : 0.f;
}

__global__ void k_unpack_linear(const float* __restrict__ wt, const float* __restrict__ bi,
                                float* __restrict__ w, float* __restrict__ b, int out_dim,
                                int in_dim, int k_pad, int n_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < out_dim * in_dim) {
    const int n = i / in_dim, k = i - n * in_dim;
    w[i] = wt[k * n_pad + n];
  }
  if (b && bi && i < out_dim) b[i] = bi[i];
}

// nn.Linear (out,in) -> tensor-core image: per 32-wide K block the TF32-rounded "hi" slab
// [n_rows][32] then the residual "lo" slab, rows of 128 bytes whose 16-byte chunks are
// XOR-swizzled with (row & 7) (canonical K-major SWIZZLE_128B operand layout of tcgen05.mma).
__global__ void k_pack_linear_umma(const float* __restrict__ w, float* __restrict__ dst, int out_dim,
                                   int in_dim, int n_rows, int k_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_kb = n_rows * 32;
  if (i >= (k_pad / 32) * per_kb) return;
  const int kb = i / per_kb, r = i - kb * per_kb;
  const int n = r >> 5, kk = r & 31;
  const int k = kb * 32 + kk;
  const float v = (n < out_dim && k < in_dim) ? w[n * in_dim + k] : 0.f;
  uint32_t hb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
  const float hi = __uint_as_float(hb);
  const float lo = v - hi;
  const int off = n * 32 + (((kk >> 2) ^ (n & 7)) << 2) + (kk & 3);
  float* blk = dst + (size_t)kb * 2 * per_kb;
  blk[off] = hi;
  blk[per_kb + off] = lo;
}

// nn.Linear (out,in) -> FP16-split tensor-core image (NvfiLinear.himg / himgT): per 64-wide K block the
// FP16 "hi" slab [n_rows][64] then the "lo" slab (v - hi, rounded to FP16), rows of 128 bytes whose
// 16-byte chunks are XOR-swizzled with (row & 7) (canonical K-major SWIZZLE_128B layout).  Element
// (row r, K index c) is w[r * sr + c * sc]: (sr, sc) = (in_dim, 1) for W, (1, in_dim) for W^T.
__global__ void k_pack_linear_h(const float* __restrict__ w, unsigned char* __restrict__ dst, int r_valid,
                                int c_valid, int sr, int sc, int n_rows, int k_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * k_pad) return;
  const int r = i / k_pad, c = i - r * k_pad;
  const float v = (r < r_valid && c < c_valid) ? w[(size_t)r * sr + (size_t)c * sc] : 0.f;
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
  const size_t blk = (size_t)(c >> 6) * ((size_t)n_rows * 256);
  const size_t off = (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 + (size_t)((((c & 63) >> 3) ^ (r & 7)) << 4) +
                     (size_t)(c & 7) * 2;
  *reinterpret_cast<__half*>(dst + blk + off) = hi;
  *reinterpret_cast<__half*>(dst + blk + (size_t)n_rows * 128 + off) = lo;
}

// Camera.get_ray_bundle (models/camera.py:112-138), per selected pixel.
__global__ void k_raygen(const float* __restrict__ pose, int H, int W, float focal,
                         const long long* __restrict__ pix, long long n, float* __restrict__ ro,
                         float* __restrict__ rd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long p = pix ? pix[i] : i;
  const int row = (int)(p / W), col = (int)(p - (long long)row * W);
  const float dx = __fdiv_rn(__fsub_rn((float)col, (float)W * 0.5f), focal);
  const float dy = -__fdiv_rn(__fsub_rn((float)row, (float)H * 0.5f), focal);
  const float dz = -1.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    // torch.sum(dirs[..., None, :] * pose[:3, :3], dim=-1): sequential sum of 3 products
    const float s0 = __fmul_rn(dx, pose[a * 4 + 0]);
    const float s1 = __fmul_rn(dy, pose[a * 4 + 1]);
    const float s2 = __fmul_rn(dz, pose[a * 4 + 2]);
    rd[i * 3 + a] = __fadd_rn(__fadd_rn(s0, s1), s2);
    ro[i * 3 + a] = pose[a * 4 + 3];
  }
}

}  // namespace nvfi

using namespace nvfi;

extern "C" int nvfi_abi_version(void) { return NVFI_ABI_VERSION; }

static int launch_transpose(const float* src, float* dst, int rows, long long cols,
                            cudaStream_t st) {
  if (!src || !dst || rows <= 0 || cols <= 0) return NVFI_EINVAL;
  dim3 block(32, 8);
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  NVFI_LAUNCH(k_transpose, grid, block, 0, st, src, dst, rows, cols);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_pack_plane(const float* src_nchw, float* dst_hwc, int r, int h, int w,
                               void* stream) {
  return launch_transpose(src_nchw, dst_hwc, r, (long long)h * w, (cudaStream_t)stream);
}

extern "C" int nvfi_unpack_plane(const float* src_hwc, float* dst_nchw, int r, int h, int w,
                                 void* stream) {
  // (H*W, R) -> (R, H*W): the same transpose with the roles of rows / cols exchanged
  if (!src_hwc || !dst_nchw || r <= 0 || h <= 0 || w <= 0) return NVFI_EINVAL;
  const long long hw = (long long)h * w;
  if (hw > 0x7fffffffLL) return NVFI_EUNSUPPORTED;
  dim3 block(32, 8);
  dim3 grid((unsigned)((r + 31) / 32), (unsigned)((hw + 31) / 32));
  NVFI_LAUNCH(k_transpose, grid, block, 0, (cudaStream_t)stream, src_hwc, dst_nchw, (int)hw, r);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_unpack_render_grads(const NvfiField* F, const NvfiRenderGrads* D, const NvfiParamGrads* P,
                                        void* stream) {
  if (!F || !D || !P) return NVFI_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  // ---- planes: packed (H*W, R) -> (R, H*W)
  TransposeBatch tb;
  memset(&tb, 0, sizeof(tb));
  long long max_hw = 0;
  int max_r = 0, nj = 0;
  const int K = F->num_keyframes;
  for (int k = 0; k < 3; ++k) {
    // space plane k: H = grid[m1], W = grid[m0]; time plane k: H = K, W = grid[n0]  (include/nvfi_b200.h)
    const int m0 = (k == 2) ? 1 : 0, m1 = (k == 0) ? 1 : 2, n0 = 2 - k;
    const long long hw_s = (long long)F->grid[m1] * F->grid[m0], hw_t = (long long)K * F->grid[n0];
    const struct {
      const float* src;
      float* dst;
      long long hw;
      int r;
    } e[4] = {{D->g_dplane_space[k], P->dplane_space[k], hw_s, F->rd},
              {D->g_dplane_time[k], P->dplane_time[k], hw_t, F->rd},
              {D->g_aplane_space[k], P->aplane_space[k], hw_s, F->ra},
              {D->g_aplane_time[k], P->aplane_time[k], hw_t, F->ra}};
    for (int j = 0; j < 4; ++j) {
      if (!e[j].src || !e[j].dst) continue;
      if (e[j].hw > 0x7fffffffLL) return NVFI_EUNSUPPORTED;
      tb.job[nj++] = TransposeJob{e[j].src, e[j].dst, (int)e[j].hw, (long long)e[j].r};
      if (e[j].hw > max_hw) max_hw = e[j].hw;
      if (e[j].r > max_r) max_r = e[j].r;
    }
  }
  if (nj > 0) {
    dim3 block(32, 8);
    dim3 grid((unsigned)((max_r + 31) / 32), (unsigned)((max_hw + 31) / 32), 12);
    if (grid.y > 65535u) return NVFI_EUNSUPPORTED;
    NVFI_LAUNCH(k_transpose_batch, grid, block, 0, st, tb);
    NVFI_CUDA_OK(cudaGetLastError());
  }
  // ---- linear layers
  UnpackBatch ub;
  memset(&ub, 0, sizeof(ub));
  int nl = 0, max_n = 0;
  auto add = [&](const NvfiLinear& L, const float* wt, const float* bi, float* w, float* b, int v4 = 0) {
    if (!wt || !w) return;
    ub.job[nl++] = UnpackJob{wt, bi, w, b, L.out_dim, L.in_dim, L.n_pad, v4};
    if (L.out_dim * L.in_dim > max_n) max_n = L.out_dim * L.in_dim;
  };
  add(F->basis_mat, D->g_basis_mat, nullptr, P->basis_mat, nullptr);
  if (F->shading_mode == NVFI_SHADING_MLP_PE)
    for (int i = 0; i < 3; ++i) add(F->render_mlp[i], D->g_render_w[i], D->g_render_b[i], P->render_w[i], P->render_b[i]);
  if (F->use_vel) {
    const bool v4 = grad_layout_v4(F, F->vel_net);
    for (int l = 0; l < NVFI_VEL_LAYERS; ++l)
      add(F->vel_net[l], D->g_vel_w[l], D->g_vel_b[l], P->vel_w[l], P->vel_b[l], (v4 && l < NVFI_VEL_LAYERS - 1) ? 1 : 0);
  }
  if (nl > 0) {
    dim3 grid((unsigned)((max_n + 255) / 256), 12);
    NVFI_LAUNCH(k_unpack_linear_batch, grid, 256, 0, st, ub);
    NVFI_CUDA_OK(cudaGetLastError());
  }
  return NVFI_OK;
}

extern "C" int nvfi_unpack_pde_grads(const NvfiField* F, const NvfiPdeGrads* G, const NvfiPdeParamGrads* P,
                                     void* stream) {
  if (!F || !G || !P) return NVFI_EINVAL;
  UnpackBatch ub;
  memset(&ub, 0, sizeof(ub));
  int nl = 0, max_n = 0;
  const int v4[2] = {grad_layout_v4(F, F->vel_net) ? 1 : 0, grad_layout_v4(F, F->acc_net) ? 1 : 0};
  for (int l = 0; l < NVFI_VEL_LAYERS; ++l) {
    const struct {
      const NvfiLinear& L;
      const float* wt;
      const float* bi;
      float* w;
      float* b;
    } e[2] = {{F->vel_net[l], G->g_vel_w[l], G->g_vel_b[l], P->vel_w[l], P->vel_b[l]},
              {F->acc_net[l], G->g_acc_w[l], G->g_acc_b[l], P->acc_w[l], P->acc_b[l]}};
    for (int j = 0; j < 2; ++j) {
      if (!e[j].wt || !e[j].w) continue;
      ub.job[nl++] = UnpackJob{e[j].wt, e[j].bi, e[j].w, e[j].b, e[j].L.out_dim, e[j].L.in_dim, e[j].L.n_pad,
                               (v4[j] && l < NVFI_VEL_LAYERS - 1) ? 1 : 0};
      if (e[j].L.out_dim * e[j].L.in_dim > max_n) max_n = e[j].L.out_dim * e[j].L.in_dim;
    }
  }
  if (nl > 0) {
    dim3 grid((unsigned)((max_n + 255) / 256), 12);
    NVFI_LAUNCH(k_unpack_linear_batch, grid, 256, 0, (cudaStream_t)stream, ub);
    NVFI_CUDA_OK(cudaGetLastError());
  }
  return NVFI_OK;
}

extern "C" int nvfi_pack_linear(const float* w, const float* b, float* wt, float* bias_out,
                                int out_dim, int in_dim, int k_pad, int n_pad, void* stream) {
  if (!w || !wt || out_dim <= 0 || in_dim <= 0 || k_pad < in_dim || n_pad < out_dim)
    return NVFI_EINVAL;
  const int n = k_pad * n_pad > n_pad ? k_pad * n_pad : n_pad;
  NVFI_LAUNCH(k_pack_linear, (n + 255) / 256, 256, 0, (cudaStream_t)stream, w, b, wt, bias_out, out_dim, in_dim, k_pad, n_pad);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_unpack_linear(const float* wt, const float* bias_in, float* w, float* b,
                                  int out_dim, int in_dim, int k_pad, int n_pad, void* stream) {
  if (!w || !wt || out_dim <= 0 || in_dim <= 0 || k_pad < in_dim || n_pad < out_dim)
    return NVFI_EINVAL;
  const int n = out_dim * in_dim;
  NVFI_LAUNCH(k_unpack_linear, (n + 255) / 256, 256, 0, (cudaStream_t)stream, wt, bias_in, w, b, out_dim, in_dim, k_pad, n_pad);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_pack_linear_umma(const float* w, float* dst, int out_dim, int in_dim, int n_rows,
                                     int k_pad, void* stream) {
  if (!w || !dst || out_dim <= 0 || in_dim <= 0 || n_rows < out_dim || k_pad < in_dim ||
      (k_pad & 31) || (n_rows & 7))
    return NVFI_EINVAL;
  const int n = (k_pad / 32) * n_rows * 32;
  NVFI_LAUNCH(k_pack_linear_umma, (n + 255) / 256, 256, 0, (cudaStream_t)stream, w, dst, out_dim,
              in_dim, n_rows, k_pad);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_pack_linear_h(const float* w, void* dst, int out_dim, int in_dim, int n_rows, int k_pad,
                                  int transposed, void* stream) {
  const int r_valid = transposed ? in_dim : out_dim, c_valid = transposed ? out_dim : in_dim;
  if (!w || !dst || out_dim <= 0 || in_dim <= 0 || n_rows < r_valid || k_pad < c_valid || (k_pad & 63) ||
      (n_rows & 7))
    return NVFI_EINVAL;
  const int n = n_rows * k_pad;
  NVFI_LAUNCH(k_pack_linear_h, (n + 255) / 256, 256, 0, (cudaStream_t)stream, w,
              reinterpret_cast<unsigned char*>(dst), r_valid, c_valid, transposed ? 1 : in_dim,
              transposed ? in_dim : 1, n_rows, k_pad);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_raygen(const float* pose4x4, int h, int w, float focal,
                           const int64_t* pixel_ids, int64_t n, float* rays_o, float* rays_d,
                           void* stream) {
  if (!pose4x4 || !rays_o || !rays_d || h <= 0 || w <= 0 || n < 0) return NVFI_EINVAL;
  if (n == 0) return NVFI_OK;
  NVFI_LAUNCH(k_raygen, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, pose4x4, h, w, focal, reinterpret_cast<const long long*>(pixel_ids), n, rays_o, rays_d);
  return (int)cudaGetLastError();
}

static int check_render_args(const NvfiField* F, const NvfiRenderArgs* A,
                             const NvfiRenderBuffers* B) {
  if (!F || !A || !B) return NVFI_EINVAL;
  if (A->n_rays < 0 || A->ray_chunk <= 0) return NVFI_EINVAL;
  if (A->n_rays == 0) return NVFI_OK;
  if (!A->rays_o || !A->rays_d || !B->rgb_map || !B->depth_map || !B->acc_map || !B->weights ||
      !B->x_adv || !B->valid || !B->rgb || !B->chunk_inside || !B->counters)
    return NVFI_EINVAL;
  if (A->training && !A->jitter) return NVFI_EINVAL;
  if (F->n_samples <= 0 || F->n_samples > 4096) return NVFI_EUNSUPPORTED;
  if (F->rd % 4 != 0 || F->rd <= 0 || F->rd > 64) return NVFI_EUNSUPPORTED;
  if (F->mask_layers > 0 && !B->mask_map) return NVFI_EINVAL;
  if (A->advect) {
    if (!F->use_vel) return NVFI_EINVAL;
    for (int l = 0; l < NVFI_VEL_LAYERS; ++l) {
      const NvfiLinear& L = F->vel_net[l];
      if (!L.wt || !L.bias) return NVFI_EINVAL;
      if (l == 0 && (L.in_dim != NVFI_VEL_IN || L.k_pad != 32)) return NVFI_EUNSUPPORTED;
      if (l > 0 && L.k_pad != 128) return NVFI_EUNSUPPORTED;
      if (l < NVFI_VEL_LAYERS - 1 && L.n_pad != 128) return NVFI_EUNSUPPORTED;
      if (l == NVFI_VEL_LAYERS - 1 && L.n_pad != 8) return NVFI_EUNSUPPORTED;
    }
  }
  return NVFI_OK;
}

extern "C" int nvfi_render_forward(const NvfiField* F, const NvfiRenderArgs* A,
                                   const NvfiRenderBuffers* B, void* stream) {
  int rc = check_render_args(F, A, B);
  if (rc != NVFI_OK) return rc;
  if (A->n_rays == 0) return NVFI_OK;
  cudaStream_t st = (cudaStream_t)stream;
  NVFI_CUDA_OK(cudaMemsetAsync(B->counters, 0, 16 * sizeof(int32_t), st));
  if (B->stats) NVFI_CUDA_OK(cudaMemsetAsync(B->stats, 0, 4 * sizeof(int64_t), st));
  if ((B->ray_T == nullptr) != (B->ray_term == nullptr)) return NVFI_EINVAL;
  const int S = F->n_samples;
  // Early ray termination (include/nvfi_b200.h, NvfiRenderBuffers.ray_T): depth waves of 32 samples (64 for
  // long rays); each wave advects the samples of the rays that are still alive and continues their march.
  const int wave = (S <= 256) ? 32 : 64;
  if (B->ray_T && A->advect && mlp_mode_of(F) == NVFI_MLP_F16X3 && S > wave) {
    rc = nvfi_launch_chunk_inside(F, A, B, st);
    if (rc != NVFI_OK) return rc;
    for (int s0 = 0; s0 < S; s0 += wave) {
      const int sw = (S - s0 < wave) ? S - s0 : wave;
      rc = nvfi_launch_sample_advect_wave(F, A, B, s0, sw, st);
      if (rc != NVFI_OK) return rc;
      rc = nvfi_launch_march_wave(F, A, B, s0, sw, st);
      if (rc != NVFI_OK) return rc;
    }
  } else {
    rc = nvfi_launch_sample_advect(F, A, B, st);
    if (rc != NVFI_OK) return rc;
    rc = nvfi_launch_march(F, A, B, st);
    if (rc != NVFI_OK) return rc;
    if (B->ray_term)   // not a wave render: every sample was evaluated
      NVFI_CUDA_OK(cudaMemsetAsync(B->ray_term, 0x7f, (size_t)A->n_rays * sizeof(int32_t), st));
  }
  rc = nvfi_launch_appearance(F, A, B, st);
  if (rc != NVFI_OK) return rc;
  return nvfi_launch_composite(F, A, B, st);
}

extern "C" int nvfi_render_forward_host(const NvfiField* F, const NvfiRenderArgs* A_in,
                                        const float* rays_o_host, const float* rays_d_host,
                                        const float* jitter_host, float* dev_rays_o,
                                        float* dev_rays_d, float* dev_jitter,
                                        const NvfiRenderBuffers* B, float* rgb_host,
                                        float* depth_host, float* acc_host, void* stream) {
  if (!F || !A_in || !B || !rays_o_host || !rays_d_host || !dev_rays_o || !dev_rays_d)
    return NVFI_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  NvfiRenderArgs A = *A_in;
  const size_t n = (size_t)A.n_rays;
  NVFI_CUDA_OK(cudaMemcpyAsync(dev_rays_o, rays_o_host, n * 3 * sizeof(float),
                               cudaMemcpyHostToDevice, st));
  NVFI_CUDA_OK(cudaMemcpyAsync(dev_rays_d, rays_d_host, n * 3 * sizeof(float),
                               cudaMemcpyHostToDevice, st));
  A.rays_o = dev_rays_o;
  A.rays_d = dev_rays_d;
  A.jitter = nullptr;
  if (jitter_host) {
    if (!dev_jitter) return NVFI_EINVAL;
    NVFI_CUDA_OK(
        cudaMemcpyAsync(dev_jitter, jitter_host, n * sizeof(float), cudaMemcpyHostToDevice, st));
    A.jitter = dev_jitter;
  }
  int rc = nvfi_render_forward(F, &A, B, st);
  if (rc != NVFI_OK) return rc;
  if (rgb_host)
    NVFI_CUDA_OK(cudaMemcpyAsync(rgb_host, B->rgb_map, n * 3 * sizeof(float),
                                 cudaMemcpyDeviceToHost, st));
  if (depth_host)
    NVFI_CUDA_OK(
        cudaMemcpyAsync(depth_host, B->depth_map, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (acc_host)
    NVFI_CUDA_OK(
        cudaMemcpyAsync(acc_host, B->acc_map, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  return (int)cudaStreamSynchronize(st);
}
