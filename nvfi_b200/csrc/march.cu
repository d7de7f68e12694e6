// Kernel 2 of the render path: density gather over the k-planes factors, density
// activation, alpha / transmittance scan and the per-ray sums that do not need colour.
// Kernel 4: colour / mask composite.
//
// Replaces:
//   compute_densityfeature   models/tensorf_keyframe.py:233-272  (6 x F.grid_sample)
//   feature2density          models/tensorf_keyframe.py:312-325
//   raw2alpha                models/tensorf_model_utils.py:186-197
//   acc / depth / rgb / mask sums, background, clamp
//                            models/tensorf_keyframe.py:737-753
//
// One warp marches one ray.  The gather is done by 8-lane groups: each group owns one
// valid sample, each lane a float4 (4 components) of every corner vector of the packed
// (H, W, R) planes, so a corner is one 96-byte (R=24) contiguous read and the 24 corner
// reads of a sample are independent loads in flight.  sigma of the ray is staged in
// shared memory, then alpha/T/weights are a warp-shuffle product scan along the ray.
#include "nvfi_common.cuh"

namespace nvfi {

#define MARCH_WARPS 8

// Depth wave [s0, s0 + sw) of every ray (s0 a multiple of 32; the whole ray when s0 == 0 and sw >= S).  With
// early ray termination (B.ray_T / B.ray_term, include/nvfi_b200.h) the transmittance and the partial sums
// are carried from wave to wave; a ray whose carried transmittance is exactly 0 is finished: its later waves
// only write zeros.
__global__ void __launch_bounds__(MARCH_WARPS * 32, 4)
    k_march(const NvfiField F, const NvfiRenderArgs A, const NvfiRenderBuffers B, int S,
            int s_pad, int s0, int sw) {
  extern __shared__ __align__(16) float sig_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ray = (long long)blockIdx.x * MARCH_WARPS + warp;
  if (ray >= A.n_rays) return;
  float* sig = sig_all + warp * s_pad;
  const int g = lane >> 3, l8 = lane & 7;
  const long long row = ray * S;
  const int s_end = min(S, s0 + sw);
  const int c_begin = s0 >> 5, n_it = (s_end + 31) / 32;
  const bool waves = B.ray_T != nullptr;
  if (waves && s0 > 0 && B.ray_term[ray] != S) {   // terminated in an earlier wave
    for (int s = s0 + lane; s < s_end; s += 32) {
      B.weights[row + s] = 0.f;
      if (B.sigma) B.sigma[row + s] = 0.f;
    }
    return;
  }

  // ---- density gather for the valid samples of this ray
  for (int c = c_begin; c < n_it; ++c) {
    const int s = c * 32 + lane;
    const bool v = (s < S) && (B.valid[row + s] != 0);
    if (s < s_pad) sig[s] = 0.f;
    unsigned m = __ballot_sync(0xffffffffu, v);
    __syncwarp();
    while (m) {
      // the g-th set bit of m (groups 0..3 take the 4 lowest set bits)
      unsigned mm = m;
      int mine = -1;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int b = mm ? (__ffs(mm) - 1) : -1;
        if (k == g) mine = b;
        mm &= mm - 1;
      }
      m = mm;
      if (mine >= 0) {
        const long long gi = row + c * 32 + mine;
        float xt[4];
        xt[0] = __ldg(B.x_adv + gi * 3 + 0);
        xt[1] = __ldg(B.x_adv + gi * 3 + 1);
        xt[2] = __ldg(B.x_adv + gi * 3 + 2);
        xt[3] = A.t_norm_base;
        const float feat = density_feature_group(F, xt, l8);
        if (l8 == 0) sig[c * 32 + mine] = feature2density(F, feat);
      }
    }
  }
  __syncwarp();

  // ---- per-ray constants of the sampler (for z and dists)
  const float o[3] = {__ldg(A.rays_o + ray * 3), __ldg(A.rays_o + ray * 3 + 1),
                      __ldg(A.rays_o + ray * 3 + 2)};
  const float d[3] = {__ldg(A.rays_d + ray * 3), __ldg(A.rays_d + ray * 3 + 1),
                      __ldg(A.rays_d + ray * 3 + 2)};
  const bool inside = B.chunk_inside[ray / A.ray_chunk] != 0;
  const float tmin = ray_tmin(F, o, d, inside);
  const bool train = A.jitter != nullptr;
  const float u = train ? __ldg(A.jitter + ray) : 0.f;

  // ---- alpha, transmittance (exclusive product scan), weights, acc, depth
  float carry = (waves && s0 > 0) ? B.ray_T[ray] : 1.f, acc = 0.f, dep = 0.f;
  unsigned n_app = 0;
  for (int c = c_begin; c < n_it; ++c) {
    const int s = c * 32 + lane;
    float alpha = 0.f, z = 0.f, sg = 0.f;
    if (s < S) {
      sg = sig[s];
      z = sample_z(tmin, F.step_size, s, u, train);
      float dist = 0.f;
      if (s + 1 < S) dist = __fsub_rn(sample_z(tmin, F.step_size, s + 1, u, train), z);
      const float dd = __fmul_rn(dist, F.distance_scale);
      // 1 - exp(-x) (models/tensorf_model_utils.py:188) evaluated as -expm1(-x): the same value, without the
      // cancellation that leaves the literal FP32 expression with only ~1e-3 relative accuracy on the tiny
      // alphas of near-empty rays (x ~ 4e-5); measured against a float64 evaluation in
      // tests/test_gpu_headline_parity.py
      alpha = -expm1f(__fmul_rn(-sg, dd));
    }
    const float f = (s < S) ? __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f) : 1.f;
    float p = f;
#pragma unroll
    for (int o2 = 1; o2 < 32; o2 <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, p, o2);
      if (lane >= o2) p *= t;
    }
    float excl = __shfl_up_sync(0xffffffffu, p, 1);
    if (lane == 0) excl = 1.f;
    const float T = carry * excl;
    const float w = alpha * T;
    carry *= __shfl_sync(0xffffffffu, p, 31);
    if (s < S) {
      B.weights[row + s] = w;
      if (B.sigma) B.sigma[row + s] = sg;
      acc += w;
      dep = fmaf(w, z, dep);
      n_app += (w > F.weight_thres) ? 1u : 0u;
    }
  }
  acc = warp_sum(acc);
  dep = warp_sum(dep);
  if (lane == 0) {
    if (!waves) {
      B.acc_map[ray] = acc;
      B.depth_map[ray] = dep + (1.f - acc) * F.far;
    } else {
      // running sums in acc_map / depth_map (the depth without its background term until the ray is finished)
      if (s0 > 0) {
        acc += B.acc_map[ray];
        dep += B.depth_map[ray];
      }
      const bool last = s_end >= S, dead = carry == 0.f;
      B.acc_map[ray] = acc;
      B.depth_map[ray] = (last || dead) ? dep + (1.f - acc) * F.far : dep;
      B.ray_T[ray] = carry;
      B.ray_term[ray] = (dead && !last) ? s_end : S;
    }
  }
  if (B.stats) {
    const float c = warp_sum((float)n_app);
    if (lane == 0 && c > 0.f)
      atomicAdd(reinterpret_cast<unsigned long long*>(B.stats) + 2, (unsigned long long)c);
  }
}

// Colour composite: rgb_map = clamp(sum_s w rgb + bg (1 - acc), 0, 1)
// (models/tensorf_keyframe.py:738-743).  One warp per ray.
__global__ void __launch_bounds__(256)
    k_composite(const NvfiField F, const NvfiRenderArgs A, const NvfiRenderBuffers B, int S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ray = (long long)blockIdx.x * 8 + warp;
  if (ray >= A.n_rays) return;
  const long long row = ray * S;
  float r = 0.f, gch = 0.f, b = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float w = B.weights[row + s];
    if (w > F.weight_thres) {
      const float* c = B.rgb + (row + s) * 3;
      r = fmaf(w, c[0], r);
      gch = fmaf(w, c[1], gch);
      b = fmaf(w, c[2], b);
    }
  }
  r = warp_sum(r);
  gch = warp_sum(gch);
  b = warp_sum(b);
  if (lane == 0) {
    const bool white = A.chunk_bg ? (A.chunk_bg[ray / A.ray_chunk] != 0) : (A.white_bg != 0);
    const float bg = white ? (1.f - B.acc_map[ray]) : 0.f;
    B.rgb_map[ray * 3 + 0] = fminf(fmaxf(r + bg, 0.f), 1.f);
    B.rgb_map[ray * 3 + 1] = fminf(fmaxf(gch + bg, 0.f), 1.f);
    B.rgb_map[ray * 3 + 2] = fminf(fmaxf(b + bg, 0.f), 1.f);
  }
}

// Stand-alone compute_densityfeature (+ optional feature2density) on arbitrary points.
__global__ void __launch_bounds__(256)
    k_density_points(const NvfiField F, const float* __restrict__ xyzt, long long n,
                     float* __restrict__ feat_out, float* __restrict__ sigma_out) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 3, l8 = lane & 7;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long base = gw * 4; base < n; base += nw * 4) {
    const long long i = base + g;
    if (i < n) {
      const float xt[4] = {__ldg(xyzt + i * 4), __ldg(xyzt + i * 4 + 1), __ldg(xyzt + i * 4 + 2),
                           __ldg(xyzt + i * 4 + 3)};
      const float feat = density_feature_group(F, xt, l8);
      if (l8 == 0) {
        if (feat_out) feat_out[i] = feat;
        if (sigma_out) sigma_out[i] = feature2density(F, feat);
      }
    }
  }
}

__global__ void k_feature2density(const NvfiField F, const float* __restrict__ feat, long long n,
                                  float* __restrict__ sigma) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    sigma[i] = feature2density(F, feat[i]);
}

}  // namespace nvfi

using namespace nvfi;

extern "C" int nvfi_launch_march_wave(const NvfiField* F, const NvfiRenderArgs* A, const NvfiRenderBuffers* B,
                                      int s0, int sw, cudaStream_t st);
extern "C" int nvfi_launch_march(const NvfiField* F, const NvfiRenderArgs* A,
                                 const NvfiRenderBuffers* B, cudaStream_t st) {
  NvfiRenderBuffers b = *B;   // the whole ray in one pass: no state is carried
  b.ray_T = nullptr;
  b.ray_term = nullptr;
  return nvfi_launch_march_wave(F, A, &b, 0, F->n_samples, st);
}

extern "C" int nvfi_launch_march_wave(const NvfiField* F, const NvfiRenderArgs* A, const NvfiRenderBuffers* B,
                                      int s0, int sw, cudaStream_t st) {
  const int S = F->n_samples;
  if (A->n_rays <= 0) return NVFI_OK;
  const int s_pad = ((S + 31) / 32) * 32;
  const size_t smem = (size_t)MARCH_WARPS * s_pad * sizeof(float);
  if (smem > 200 * 1024) return NVFI_EUNSUPPORTED;
  if (smem > 48 * 1024) {
    const int rc = ensure_smem<k_march>(smem);
    if (rc != NVFI_OK) return rc;
  }
  const long long grid = (A->n_rays + MARCH_WARPS - 1) / MARCH_WARPS;
  NVFI_LAUNCH(k_march, (unsigned)grid, MARCH_WARPS * 32, smem, st, *F, *A, *B, S, s_pad, s0, sw);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_launch_composite(const NvfiField* F, const NvfiRenderArgs* A,
                                     const NvfiRenderBuffers* B, cudaStream_t st) {
  if (A->n_rays <= 0) return NVFI_OK;
  const long long grid = (A->n_rays + 7) / 8;
  NVFI_LAUNCH(k_composite, (unsigned)grid, 256, 0, st, *F, *A, *B, F->n_samples);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_density_feature(const NvfiField* F, const float* xyzt, int64_t n, float* feat,
                                    void* stream) {
  if (!F || !xyzt || !feat || n < 0) return NVFI_EINVAL;
  if (n == 0) return NVFI_OK;
  const long long groups = (n + 3) / 4;  // warps needed
  long long grid = (groups + 7) / 8;
  if (grid > 148 * 64) grid = 148 * 64;
  NVFI_LAUNCH(k_density_points, (unsigned)grid, 256, 0, (cudaStream_t)stream, *F, xyzt, n, feat, nullptr);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_density_sigma(const NvfiField* F, const float* xyzt, int64_t n, float* sigma,
                                  void* stream) {
  if (!F || !xyzt || !sigma || n < 0) return NVFI_EINVAL;
  if (n == 0) return NVFI_OK;
  const long long groups = (n + 3) / 4;
  long long grid = (groups + 7) / 8;
  if (grid > 148 * 64) grid = 148 * 64;
  NVFI_LAUNCH(k_density_points, (unsigned)grid, 256, 0, (cudaStream_t)stream, *F, xyzt, n, nullptr, sigma);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_feature2density(const NvfiField* F, const float* feat, int64_t n, float* sigma,
                                    void* stream) {
  if (!F || !feat || !sigma || n < 0) return NVFI_EINVAL;
  if (n == 0) return NVFI_OK;
  long long grid = (n + 255) / 256;
  if (grid > 148 * 32) grid = 148 * 32;
  NVFI_LAUNCH(k_feature2density, (unsigned)grid, 256, 0, (cudaStream_t)stream, *F, feat, n, sigma);
  return (int)cudaGetLastError();
}
