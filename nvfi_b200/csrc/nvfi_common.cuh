// Shared device code of the nvfi_b200 kernels (sm_100a).
//
// Numerics notes (parity with the reference's PyTorch FP32 path):
//  * Everything that decides an INTEGER quantity (sample validity, keyframe snap, RK2
//    step count, gate masks, texel indices) is written with explicit round-to-nearest
//    intrinsics (__fadd_rn / __fmul_rn / __fsub_rn / __fdiv_rn) so that nvcc cannot
//    contract a*b+c into an FMA: PyTorch executes those as separate rounded ops
//    (models/tensorf_base.py:290-314, models/tensorf_keyframe.py:575-611).
//  * Value arithmetic (bilinear blends, MLP dot products, compositing sums) may use FMA
//    and a different summation order; it is held to 1e-4 relative, not bit-exactness.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

#include "nvfi_b200.h"

#define NVFI_TM 128        // samples per MLP tile
#define NVFI_THREADS 256   // threads per CTA of the tile kernels
#define NVFI_KC 32         // k rows of W^T staged in shared memory per pipeline stage
#define NVFI_ACT_ROWS 160  // rows of the activation tile (max padded input width)
#define NVFI_QCAP 384      // compaction queue capacity (128 carried + 256 new)
#define NVFI_SUBS 8        // sub-batches of 256 raw samples per grabbed batch

#define ACT_NONE 0
#define ACT_RELU 1
#define ACT_SILU 2

#define NVFI_CUDA_OK(call)                         \
  do {                                             \
    cudaError_t _e = (call);                       \
    if (_e != cudaSuccess) return (int)_e;         \
  } while (0)

#define NVFI_LAUNCH(kernel, grid, block, smem, st, ...)       \
  do {                                                         \
    nvfi::ProfScope ps_(#kernel, (st));                        \
    kernel<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__);    \
  } while (0)

namespace nvfi {

// Arithmetic of the velocity-MLP GEMMs of a call (NvfiField.mlp_mode; 0 / out of range = the product path).
inline int mlp_mode_of(const NvfiField* F) {
  const int m = F->mlp_mode;
  return (m < NVFI_MLP_FP32_SIMT || m > NVFI_MLP_F16X3) ? NVFI_MLP_F16X3 : m;
}

// The FP16-split tensor-core kernels (backward_h.cu) accumulate the weight gradients of the hidden layers
// in the layout [k / 4][n][k % 4] (a thread adds 4 consecutive k of its unit n with ONE vector reduction,
// a warp 512 contiguous bytes) instead of the [k][n] of NvfiLinear.wt; they run when the call selects
// NVFI_MLP_F16X3 and the net carries both FP16 images.  The unpack entry points ask the same question.
inline bool grad_layout_v4(const NvfiField* F, const NvfiLinear* net) {
  if (mlp_mode_of(F) != NVFI_MLP_F16X3) return false;
  for (int l = 0; l < NVFI_VEL_LAYERS; ++l)
    if (!net[l].himg || !net[l].himgT) return false;
  return true;
}

constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = kMaxDevices - 1;
  return dev;
}
// Opt-in to more than 48 KB of dynamic shared memory.  cudaFuncSetAttribute is per DEVICE, so the
// "already done" cache is per (kernel, device): one instantiation of this template per kernel.
template <auto Kernel>
inline int ensure_smem(size_t smem) {
  static size_t cached[kMaxDevices] = {};
  const int dev = current_device();
  if (smem > cached[dev] || dev == kMaxDevices - 1) {
    const cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cached[dev] = smem;
  }
  return 0;
}
// SM count of the current device (per-device cache).
inline int device_sms() {
  static int cached[kMaxDevices] = {};
  const int dev = current_device();
  if (cached[dev] <= 0) {
    int n = 0, d = 0;
    cudaGetDevice(&d);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

// Sub-batches (of `threads` raw samples) per atomically grabbed batch of the persistent
// compaction kernels: NVFI_SUBS when there is plenty of work, fewer when the launch would
// otherwise have fewer than ~6 batches per SM.
inline int grab_subs(long long total, int threads, int sms) {
  const long long subs = total / ((long long)threads * sms * 6);
  return (int)(subs < 1 ? 1 : (subs > NVFI_SUBS ? NVFI_SUBS : subs));
}

// Host side: counts a kernel launch and, when profiling is on, times it (prof.cu).
// Used through NVFI_LAUNCH.
struct ProfScope {
  ProfScope(const char* name, cudaStream_t st);
  ~ProfScope();
  cudaStream_t st_;
  int idx_;
};

// ------------------------------------------------------------------ scalar helpers
__device__ __forceinline__ float softplus_f(float x) {  // F.softplus(beta=1, threshold=20)
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }

template <int ACT>
__device__ __forceinline__ float activate(float x) {
  if (ACT == ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == ACT_SILU) return silu_f(x);
  return x;
}

__device__ __forceinline__ float feature2density(const NvfiField& F, float feat) {
  // models/tensorf_keyframe.py:312-325 (densityMode == "Density")
  if (F.fea2dense_act == NVFI_ACT_SOFTPLUS) return softplus_f(__fadd_rn(feat, F.density_shift));
  if (F.fea2dense_act == NVFI_ACT_RELU) return fmaxf(feat, 0.f);
  return fmaxf(fabsf(feat), 0.f);
}

// ------------------------------------------------------------------ ray sampling
// models/tensorf_base.py:290-314.  `inside` is the chunk-global predicate of line 294.
__device__ __forceinline__ float ray_tmin(const NvfiField& F, const float o[3], const float d[3],
                                          bool inside) {
  if (inside) return F.near;
  float tm = -INFINITY;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float vec = (d[a] == 0.f) ? 1e-6f : d[a];
    float ra = __fdiv_rn(__fsub_rn(F.aabb_max[a], o[a]), vec);
    float rb = __fdiv_rn(__fsub_rn(F.aabb_min[a], o[a]), vec);
    tm = fmaxf(tm, fminf(ra, rb));
  }
  return fminf(fmaxf(tm, F.near), F.far);
}

__device__ __forceinline__ float sample_z(float tmin, float step, int s, float u, bool train) {
  float r = (float)s;
  if (train) r = __fadd_rn(r, u);
  return __fadd_rn(tmin, __fmul_rn(step, r));
}

// Position, validity and normalised coordinate of one sample.
__device__ __forceinline__ bool sample_point(const NvfiField& F, const float o[3], const float d[3],
                                             float z, float xn[3]) {
  bool ok = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float p = __fadd_rn(o[a], __fmul_rn(d[a], z));
    ok = ok && !(F.aabb_min[a] > p) && !(p > F.aabb_max[a]);
    xn[a] = __fsub_rn(__fmul_rn(__fsub_rn(p, F.aabb_min[a]), F.inv_aabb[a]), 1.f);
  }
  return ok;
}

// AlphaGridMask.sample_alpha(...) > 0  (models/tensorf_model_utils.py:433-439):
// trilinear grid_sample, align_corners=True, zero padding, on a 0/1 volume.
__device__ __forceinline__ bool alpha_mask_keep(const NvfiField& F, const float xn[3]) {
  const int gx = F.alpha_grid[0], gy = F.alpha_grid[1], gz = F.alpha_grid[2];
  float fx = __fmul_rn(__fadd_rn(xn[0], 1.f), 0.5f * (float)(gx - 1));
  float fy = __fmul_rn(__fadd_rn(xn[1], 1.f), 0.5f * (float)(gy - 1));
  float fz = __fmul_rn(__fadd_rn(xn[2], 1.f), 0.5f * (float)(gz - 1));
  float x0f = floorf(fx), y0f = floorf(fy), z0f = floorf(fz);
  int x0 = (int)x0f, y0 = (int)y0f, z0 = (int)z0f;
  float wx1 = fx - x0f, wy1 = fy - y0f, wz1 = fz - z0f;
  float wx0 = (x0f + 1.f) - fx, wy0 = (y0f + 1.f) - fy, wz0 = (z0f + 1.f) - fz;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    int xi = x0 + (c & 1), yi = y0 + ((c >> 1) & 1), zi = z0 + (c >> 2);
    if (xi < 0 || xi >= gx || yi < 0 || yi >= gy || zi < 0 || zi >= gz) continue;
    float w = ((c & 1) ? wx1 : wx0) * (((c >> 1) & 1) ? wy1 : wy0) * ((c >> 2) ? wz1 : wz0);
    if (F.alpha_volume[((size_t)zi * gy + yi) * gx + xi]) acc += w;
  }
  return acc > 0.f;
}

// ------------------------------------------------------------------ velocity basis
// models/velocity_field.py:77-98: v = sum_i w_i b_i(x)
__device__ __forceinline__ void basis_velocity(const float w[6], float x, float y, float z,
                                               float v[3]) {
  v[0] = w[0] - w[4] * z + w[5] * y;
  v[1] = w[1] + w[3] * z - w[5] * x;
  v[2] = w[2] - w[3] * y + w[4] * x;
}
// models/velocity_field.py:69-75, 94-97
__device__ __forceinline__ void basis_acceleration(const float a[6], float x, float y, float z,
                                                   float o[3]) {
  o[0] = a[0] - a[4] * x - a[5] * x;
  o[1] = a[1] - a[3] * y - a[5] * y;
  o[2] = a[2] - a[3] * z - a[4] * z;
}

__device__ __forceinline__ bool gate_outside(const NvfiField& F, float x, float y, float z) {
  return (x < F.gate_lo[0]) | (x > F.gate_hi[0]) | (y < F.gate_lo[1]) | (y > F.gate_hi[1]) |
         (z < F.gate_lo[2]) | (z > F.gate_hi[2]);
}

// ------------------------------------------------------------------ bilinear plane access
// grid_sample(bilinear, zeros, align_corners=True) on a packed (H, W, R) plane
// (SURVEY.md Appendix A item 8).  gx indexes W, gy indexes H.
struct Bilerp {
  int off[4];    // float offsets of the 4 corner vectors (clamped in range)
  float w[4];    // nw, ne, sw, se weights; 0 for out-of-range corners
  float dwx[4];  // d w / d fx (for coordinate gradients)
  float dwy[4];  // d w / d fy
};

__device__ __forceinline__ void bilerp_setup(float gx, float gy, int H, int W, int R, Bilerp& b) {
  float fx = __fmul_rn(__fadd_rn(gx, 1.f), 0.5f * (float)(W - 1));
  float fy = __fmul_rn(__fadd_rn(gy, 1.f), 0.5f * (float)(H - 1));
  float x0f = floorf(fx), y0f = floorf(fy);
  // keep the int conversion well defined for wild coordinates
  x0f = fminf(fmaxf(x0f, -2.f), (float)W);
  y0f = fminf(fmaxf(y0f, -2.f), (float)H);
  int x0 = (int)x0f, y0 = (int)y0f;
  float wx1 = fx - x0f, wy1 = fy - y0f;
  float wx0 = (x0f + 1.f) - fx, wy0 = (y0f + 1.f) - fy;
  bool vx0 = (x0 >= 0) & (x0 < W), vx1 = (x0 + 1 >= 0) & (x0 + 1 < W);
  bool vy0 = (y0 >= 0) & (y0 < H), vy1 = (y0 + 1 >= 0) & (y0 + 1 < H);
  int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x0 + 1, 0), W - 1);
  int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y0 + 1, 0), H - 1);
  b.off[0] = (cy0 * W + cx0) * R;
  b.off[1] = (cy0 * W + cx1) * R;
  b.off[2] = (cy1 * W + cx0) * R;
  b.off[3] = (cy1 * W + cx1) * R;
  float m0 = (vx0 & vy0) ? 1.f : 0.f, m1 = (vx1 & vy0) ? 1.f : 0.f;
  float m2 = (vx0 & vy1) ? 1.f : 0.f, m3 = (vx1 & vy1) ? 1.f : 0.f;
  b.w[0] = wx0 * wy0 * m0;
  b.w[1] = wx1 * wy0 * m1;
  b.w[2] = wx0 * wy1 * m2;
  b.w[3] = wx1 * wy1 * m3;
  b.dwx[0] = -wy0 * m0;
  b.dwx[1] = wy0 * m1;
  b.dwx[2] = -wy1 * m2;
  b.dwx[3] = wy1 * m3;
  b.dwy[0] = -wx0 * m0;
  b.dwy[1] = -wx1 * m1;
  b.dwy[2] = wx0 * m2;
  b.dwy[3] = wx1 * m3;
}

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float4 f4_blend(const float4& a, const float4& b, const float4& c,
                                           const float4& d, const float w[4]) {
  float4 r;
  r.x = a.x * w[0] + b.x * w[1] + c.x * w[2] + d.x * w[3];
  r.y = a.y * w[0] + b.y * w[1] + c.y * w[2] + d.y * w[3];
  r.z = a.z * w[0] + b.z * w[1] + c.z * w[2] + d.z * w[3];
  r.w = a.w * w[0] + b.w * w[1] + c.w * w[2] + d.w * w[3];
  return r;
}
__device__ __forceinline__ float4 f4_mul(const float4& a, const float4& b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float f4_dot(const float4& a, const float4& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
__device__ __forceinline__ float f4_sum(const float4& a) { return (a.x + a.y) + (a.z + a.w); }

// Plane geometry of the k-planes factorisation (models/tensorf_keyframe.py:39-40,143-149).
//   plane p in 0..2 : space plane p, coords (x[m0], x[m1]),   H = G[m1], W = G[m0]
//   plane p in 3..5 : time plane p-3, coords (x[n0], t),       H = K,     W = G[n0]
__device__ __forceinline__ void plane_geom(const NvfiField& F, int p, int& cx, int& cy, int& H,
                                           int& W) {
  // matModeSpace = [[0,1],[0,2],[1,2]]   matModeTime = [[2,3],[1,3],[0,3]]
  const int m0[6] = {0, 0, 1, 2, 1, 0};
  const int m1[6] = {1, 2, 2, 3, 3, 3};
  cx = m0[p];
  cy = m1[p];
  W = F.grid[cx];
  H = (p < 3) ? F.grid[cy] : F.num_keyframes;
}

// Hadamard-product features of one point, cooperatively by the 8 lanes of a group.
// Lane l8 owns float4 slots q = l8 + 8*j < R4.  On return prod[j] holds, for its 4
// channels, prod over the 6 planes (space factors multiplied first, then time factors,
// then the two — the reference's order, models/tensorf_keyframe.py:266-272).
template <int NSLOT>
__device__ __forceinline__ void kplanes_features(const NvfiField& F, const float* const sp[3],
                                                 const float* const tp[3], int R, const float xt[4],
                                                 int l8, float4 prod[NSLOT]) {
  const int R4 = R >> 2;
  float4 ps[NSLOT], pt[NSLOT];
#pragma unroll
  for (int j = 0; j < NSLOT; ++j) {
    ps[j] = make_float4(1.f, 1.f, 1.f, 1.f);
    pt[j] = make_float4(1.f, 1.f, 1.f, 1.f);
  }
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    int cx, cy, H, W;
    plane_geom(F, p, cx, cy, H, W);
    Bilerp b;
    bilerp_setup(xt[cx], xt[cy], H, W, R, b);
    const float* base = (p < 3) ? sp[p] : tp[p - 3];
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) {
      int q = l8 + 8 * j;
      if (q < R4) {
        float4 v0 = ldg4(base + b.off[0] + 4 * q);
        float4 v1 = ldg4(base + b.off[1] + 4 * q);
        float4 v2 = ldg4(base + b.off[2] + 4 * q);
        float4 v3 = ldg4(base + b.off[3] + 4 * q);
        float4 v = f4_blend(v0, v1, v2, v3, b.w);
        if (p < 3)
          ps[j] = f4_mul(ps[j], v);
        else
          pt[j] = f4_mul(pt[j], v);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NSLOT; ++j) prod[j] = f4_mul(ps[j], pt[j]);
}

// Sum over the 8 lanes of a group (lanes g*8 .. g*8+7).  Only the group's own lanes
// are named in the mask, so different groups of a warp may diverge.
__device__ __forceinline__ unsigned group8_mask() {
  return 0xffu << ((threadIdx.x & 31u) & ~7u);
}
__device__ __forceinline__ float group8_sum(float v) {
  const unsigned gm = group8_mask();
  v += __shfl_xor_sync(gm, v, 4);
  v += __shfl_xor_sync(gm, v, 2);
  v += __shfl_xor_sync(gm, v, 1);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Density feature of one point by an 8-lane group; every lane returns the full sum.
__device__ __forceinline__ float density_feature_group(const NvfiField& F, const float xt[4],
                                                       int l8) {
  float part = 0.f;
  const int R4 = F.rd >> 2;
  if (R4 <= 8) {
    float4 pr[1];
    kplanes_features<1>(F, F.dplane_space, F.dplane_time, F.rd, xt, l8, pr);
    if (l8 < R4) part = f4_sum(pr[0]);
  } else {
    float4 pr[2];
    kplanes_features<2>(F, F.dplane_space, F.dplane_time, F.rd, xt, l8, pr);
    if (l8 < R4) part = f4_sum(pr[0]);
    if (l8 + 8 < R4) part += f4_sum(pr[1]);
  }
  return group8_sum(part);
}

// ------------------------------------------------------------------ cp.async helpers
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ------------------------------------------------------------------ tile GEMM (FP32 SIMT)
// One dense layer on a tile of 128 samples held k-major in shared memory:
//   actT[k][m] (k < L.k_pad, m < 128)  ->  actT[n][m] = act(bias[n] + sum_k actT[k][m] W^T[k][n])
// for n < 128 (in place).  256 threads; thread (tm, tn) owns rows {4tm..4tm+3, 64+4tm..+3}
// x columns {8tn..8tn+7}.  W^T streams through a 2-stage cp.async ring of 32 k-rows
// (16 KB) per stage.  Rows of actT in [in_dim, k_pad) must be zero (finite).
template <int ACT, int RS = NVFI_TM, bool STASH = false>
__device__ void tile_linear128(float* __restrict__ actT, float* __restrict__ wS,
                               const NvfiLinear& L, float* __restrict__ stash = nullptr) {
  const int tid = threadIdx.x;
  const int tm = tid & 15, tn = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nchunk = L.k_pad / NVFI_KC;
  const float* wt = L.wt;
  // stage 0
#pragma unroll
  for (int i = 0; i < 4; ++i) cp_async16(wS + (tid + i * 256) * 4, wt + (tid + i * 256) * 4);
  cp_async_commit();
  for (int c = 0; c < nchunk; ++c) {
    if (c + 1 < nchunk) {
      float* dst = wS + ((c + 1) & 1) * (NVFI_KC * 128);
      const float* src = wt + (size_t)(c + 1) * (NVFI_KC * 128);
#pragma unroll
      for (int i = 0; i < 4; ++i) cp_async16(dst + (tid + i * 256) * 4, src + (tid + i * 256) * 4);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* w = wS + (c & 1) * (NVFI_KC * 128) + tn * 8;
    const float* a = actT + c * (NVFI_KC * RS) + tm * 4;
#pragma unroll 4
    for (int kk = 0; kk < NVFI_KC; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(a + kk * RS);
      const float4 a1 = *reinterpret_cast<const float4*>(a + kk * RS + 64);
      const float4 w0 = *reinterpret_cast<const float4*>(w + kk * 128);
      const float4 w1 = *reinterpret_cast<const float4*>(w + kk * 128 + 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // epilogue (all reads of actT are behind the last barrier)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = tn * 8 + j;
    const float bv = L.bias ? __ldg(L.bias + n) : 0.f;
    float4 h0 = make_float4(acc[0][j] + bv, acc[1][j] + bv, acc[2][j] + bv, acc[3][j] + bv);
    float4 h1 = make_float4(acc[4][j] + bv, acc[5][j] + bv, acc[6][j] + bv, acc[7][j] + bv);
    if (STASH) {  // pre-activations, (128, 128) row-major per layer, for the backward pass
      *reinterpret_cast<float4*>(stash + n * NVFI_TM + tm * 4) = h0;
      *reinterpret_cast<float4*>(stash + n * NVFI_TM + 64 + tm * 4) = h1;
    }
    float4 o0, o1;
    o0.x = activate<ACT>(h0.x);
    o0.y = activate<ACT>(h0.y);
    o0.z = activate<ACT>(h0.z);
    o0.w = activate<ACT>(h0.w);
    o1.x = activate<ACT>(h1.x);
    o1.y = activate<ACT>(h1.y);
    o1.z = activate<ACT>(h1.z);
    o1.w = activate<ACT>(h1.w);
    *reinterpret_cast<float4*>(actT + n * RS + tm * 4) = o0;
    *reinterpret_cast<float4*>(actT + n * RS + 64 + tm * 4) = o1;
  }
  __syncthreads();
}

// Narrow output layer (n_pad <= 2*NH): thread handles sample m = tid & 127 and output
// columns [half*nh, half*nh + nh) with half = tid >> 7, nh = n_pad / 2.  Results go to
// outS[n][m] (a separate shared buffer, n < n_pad), then a barrier.
template <int NH, int RS = NVFI_TM>
__device__ void tile_linear_small(const float* __restrict__ actT, float* __restrict__ outS,
                                  const NvfiLinear& L) {
  const int tid = threadIdx.x;
  const int m = tid & 127, half = tid >> 7;
  const int nh = L.n_pad >> 1;
  const int n0 = half * nh;
  float acc[NH];
#pragma unroll
  for (int j = 0; j < NH; ++j) acc[j] = 0.f;
  const float* wt = L.wt + n0;
  for (int k = 0; k < L.in_dim; ++k) {
    const float a = actT[k * RS + m];
    const float* wr = wt + (size_t)k * L.n_pad;
#pragma unroll
    for (int j = 0; j < NH; ++j)
      if (j < nh) acc[j] = fmaf(a, __ldg(wr + j), acc[j]);
  }
#pragma unroll
  for (int j = 0; j < NH; ++j)
    if (j < nh) outS[(n0 + j) * NVFI_TM + m] = acc[j] + (L.bias ? __ldg(L.bias + n0 + j) : 0.f);
  __syncthreads();
}

// PositionEncoder(3) of (x, y, z, t) into rows 0..31 of actT (28 values + 4 zero rows)
// (models/base_network.py:42-54).  256 threads: two threads per sample.
template <int RS = NVFI_TM>
__device__ __forceinline__ void vel_encode_tile(float* __restrict__ actT, const float* xs,
                                                const float* ys, const float* zs, const float* ts) {
  const int tid = threadIdx.x;
  const int m = tid & 127, part = tid >> 7;
  const float q[4] = {xs[m], ys[m], zs[m], ts[m]};
  if (part == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      actT[i * RS + m] = q[i];
      float s, c;
      sincosf(q[i], &s, &c);
      actT[(4 + i) * RS + m] = s;
      actT[(8 + i) * RS + m] = c;
      actT[(12 + i) * RS + m] = sinf(q[i] * 2.f);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      actT[(16 + i) * RS + m] = cosf(q[i] * 2.f);
      float s, c;
      sincosf(q[i] * 4.f, &s, &c);
      actT[(20 + i) * RS + m] = s;
      actT[(24 + i) * RS + m] = c;
      actT[(28 + i) * RS + m] = 0.f;
    }
  }
  __syncthreads();
}

// Weight net of VelBasis on a tile: inputs (x,y,z,t)[m] -> outS[0..5][m] (6 basis weights).
// With STASH the pre-activations of the 5 hidden layers go to stash[l] (128x128 each).
template <int ACT, int RS = NVFI_TM, bool STASH = false>
__device__ void vel_net_tile(const NvfiLinear* net, float* actT, float* wS, float* outS,
                             const float* xs, const float* ys, const float* zs, const float* ts,
                             float* stash = nullptr) {
  vel_encode_tile<RS>(actT, xs, ys, zs, ts);
#pragma unroll 1
  for (int l = 0; l < NVFI_VEL_LAYERS - 1; ++l)
    tile_linear128<ACT, RS, STASH>(actT, wS, net[l],
                                   STASH ? stash + (size_t)l * NVFI_TM * NVFI_TM : nullptr);
  tile_linear_small<4, RS>(actT, outS, net[NVFI_VEL_LAYERS - 1]);
}

// Shared memory of one RK2 advection tile.
struct AdvectTile {
  float x[3][NVFI_TM];    // current position (normalised)
  float xm[3][NVFI_TM];   // midpoint position
  float tcur[NVFI_TM];    // current time
  float tmid[NVFI_TM];
  float off[NVFI_TM];     // remaining time offset (0 = finished / padding row)
  float dt[NVFI_TM];
  float wout[8][NVFI_TM]; // basis weights
};

// integrate_pos on one tile (models/tensorf_keyframe.py:575-611): RK2 midpoint steps
// until every row's remaining offset is exactly zero.
// `net(xs, ys, zs, ts)` evaluates the weight net on the tile and leaves the 6 basis weights in
// T.wout (FP32 SIMT tile GEMM or the tcgen05 path of mlp_tc.cuh).
template <class Net>
__device__ inline void advect_tile_with(const NvfiField& F, AdvectTile& T, Net net) {
  const int tid = threadIdx.x;
  for (;;) {
    int active = (tid < NVFI_TM) ? (fabsf(T.off[tid]) > 0.f) : 0;
    if (!__syncthreads_or(active)) break;
    net(T.x[0], T.x[1], T.x[2], T.tcur);
    if (tid < NVFI_TM) {
      const int m = tid;
      const float off = T.off[m];
      const float a = fabsf(off);
      float dt = fminf(a, F.dt_max);
      dt = (off > 0.f) ? dt : ((off < 0.f) ? -dt : 0.f);
      const float x = T.x[0][m], y = T.x[1][m], z = T.x[2][m];
      float v[3] = {0.f, 0.f, 0.f};
      if (!gate_outside(F, x, y, z)) {
        const float w[6] = {T.wout[0][m], T.wout[1][m], T.wout[2][m],
                            T.wout[3][m], T.wout[4][m], T.wout[5][m]};
        basis_velocity(w, x, y, z, v);
      }
      const float hdt = 0.5f * dt;
      T.xm[0][m] = __fsub_rn(x, __fmul_rn(hdt, v[0]));
      T.xm[1][m] = __fsub_rn(y, __fmul_rn(hdt, v[1]));
      T.xm[2][m] = __fsub_rn(z, __fmul_rn(hdt, v[2]));
      T.tmid[m] = __fsub_rn(T.tcur[m], hdt);
      T.dt[m] = dt;
    }
    __syncthreads();
    net(T.xm[0], T.xm[1], T.xm[2], T.tmid);
    if (tid < NVFI_TM) {
      const int m = tid;
      const float off = T.off[m];
      if (fabsf(off) > 0.f) {
        const float dt = T.dt[m];
        const float xm = T.xm[0][m], ym = T.xm[1][m], zm = T.xm[2][m];
        float v[3] = {0.f, 0.f, 0.f};
        if (!gate_outside(F, xm, ym, zm)) {
          const float w[6] = {T.wout[0][m], T.wout[1][m], T.wout[2][m],
                              T.wout[3][m], T.wout[4][m], T.wout[5][m]};
          basis_velocity(w, xm, ym, zm, v);
        }
        const float x = T.x[0][m], y = T.x[1][m], z = T.x[2][m];
        float nx = __fsub_rn(x, __fmul_rn(dt, v[0]));
        float ny = __fsub_rn(y, __fmul_rn(dt, v[1]));
        float nz = __fsub_rn(z, __fmul_rn(dt, v[2]));
        if (F.vel_gate == NVFI_GATE_SUR && gate_outside(F, nx, ny, nz)) {  // :603-605
          nx = x;
          ny = y;
          nz = z;
        }
        T.x[0][m] = nx;
        T.x[1][m] = ny;
        T.x[2][m] = nz;
        T.off[m] = __fsub_rn(off, dt);
        T.tcur[m] = __fsub_rn(T.tcur[m], dt);
      }
    }
    __syncthreads();
  }
}

__device__ inline void advect_tile(const NvfiField& F, AdvectTile& T, float* actT, float* wS) {
  advect_tile_with(F, T, [&](const float* xs, const float* ys, const float* zs, const float* ts) {
    vel_net_tile<ACT_SILU>(F.vel_net, actT, wS, &T.wout[0][0], xs, ys, zs, ts);
  });
}

}  // namespace nvfi
