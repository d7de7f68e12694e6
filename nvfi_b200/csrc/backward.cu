// Backward pass of the fused render (autograd of models/tensorf_keyframe.py:613-755 as
// derived in SURVEY.md Appendix E), hand written:
//
//   k_march_bwd    per ray: clamp mask, dL/dw_i, reverse (suffix) scan -> dL/dsigma_i
//   k_app_bwd      appearance samples: recompute gather + MLP_PE (or SH), backprop to the
//                  render MLP, basis_mat, the 48-component planes and the sample position
//   k_density_bwd  valid samples: dL/dsigma -> 24-component planes + sample position
//   k_advect_bwd   advected samples: RK2 adjoint through the velocity MLP -> vel_net grads
//   k_reduce_*     sum the per-CTA weight-gradient partials into the packed gradients
#include "backward_common.cuh"

namespace nvfi {

// ---------------------------------------------------------------------------------------
// k_march_bwd
// ---------------------------------------------------------------------------------------
// The reverse scan runs in FLOAT64: dL/dsigma_i = (G_i T_i - R_i / (1 - a_i + eps)) d_i (1 - a_i) is a
// difference of nearly equal terms once a ray saturates, and the rounding of an FP32 scan is
// common to all samples in front of a surface, so it does not average out in the plane
// gradients (measured: 1e-4 relative there, against 2e-5 for the reference's own FP32 autograd;
// profiles/r01a_diag_grad_chess.txt).  The kernel is <0.1 % of a step either way.
__global__ void __launch_bounds__(256)
    k_march_bwd(const NvfiField F, const NvfiRenderArgs A, const NvfiRenderBuffers B,
                const NvfiRenderGrads D, int S, int s_pad) {
  extern __shared__ __align__(16) double smd_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ray = (long long)blockIdx.x * 8 + warp;
  if (ray >= A.n_rays) return;
  double* om = smd_all + (size_t)warp * 3 * s_pad;   // 1 - alpha_i
  double* GT = om + s_pad;
  double* Gw = GT + s_pad;
  const long long row = ray * S;
  const int n_it = (S + 31) / 32;

  const float o[3] = {__ldg(A.rays_o + ray * 3), __ldg(A.rays_o + ray * 3 + 1),
                      __ldg(A.rays_o + ray * 3 + 2)};
  const float d[3] = {__ldg(A.rays_d + ray * 3), __ldg(A.rays_d + ray * 3 + 1),
                      __ldg(A.rays_d + ray * 3 + 2)};
  const bool inside = B.chunk_inside[ray / A.ray_chunk] != 0;
  const float tmin = ray_tmin(F, o, d, inside);
  const bool train = A.jitter != nullptr;
  const float u = train ? __ldg(A.jitter + ray) : 0.f;
  const bool white = A.chunk_bg ? (A.chunk_bg[ray / A.ray_chunk] != 0) : (A.white_bg != 0);

  // pre-clamp colour -> clamp mask (models/tensorf_keyframe.py:738-743), from the forward's values
  float r = 0.f, g = 0.f, bl = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float w = B.weights[row + s];
    if (w > F.weight_thres) {
      const float* c = B.rgb + (row + s) * 3;
      r = fmaf(w, c[0], r);
      g = fmaf(w, c[1], g);
      bl = fmaf(w, c[2], bl);
    }
  }
  r = warp_sum(r);
  g = warp_sum(g);
  bl = warp_sum(bl);
  const float bg = white ? (1.f - B.acc_map[ray]) : 0.f;
  float e[3] = {0.f, 0.f, 0.f};
  if (D.g_rgb) {
    const float pre[3] = {r + bg, g + bg, bl + bg};
#pragma unroll
    for (int c = 0; c < 3; ++c)
      e[c] = (pre[c] >= 0.f && pre[c] <= 1.f) ? D.g_rgb[ray * 3 + c] : 0.f;
  }
  if (lane == 0) {
    D.g_rgb_eff[ray * 3 + 0] = e[0];
    D.g_rgb_eff[ray * 3 + 1] = e[1];
    D.g_rgb_eff[ray * 3 + 2] = e[2];
  }
  const double gsum = white ? ((double)e[0] + (double)e[1] + (double)e[2]) : 0.0;
  const double gD = D.g_depth ? (double)D.g_depth[ray] : 0.0;
  const double gA = D.g_acc ? (double)D.g_acc[ray] : 0.0;

  // forward sweep: alpha, T, dL/dw
  double carry = 1.0;
  for (int c = 0; c < n_it; ++c) {
    const int s = c * 32 + lane;
    double one_m = 1.0;   // exp(-sigma d) = 1 - alpha
    float z = 0.f;
    if (s < S) {
      const float sg = B.sigma[row + s];
      z = sample_z(tmin, F.step_size, s, u, train);
      float dist = 0.f;
      if (s + 1 < S) dist = __fsub_rn(sample_z(tmin, F.step_size, s + 1, u, train), z);
      one_m = exp(-(double)sg * (double)__fmul_rn(dist, F.distance_scale));
    }
    const double f = (s < S) ? one_m + 1e-10 : 1.0;
    double p = f;
#pragma unroll
    for (int o2 = 1; o2 < 32; o2 <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, p, o2);
      if (lane >= o2) p *= t;
    }
    double excl = __shfl_up_sync(0xffffffffu, p, 1);
    if (lane == 0) excl = 1.0;
    const double T = carry * excl;
    carry *= __shfl_sync(0xffffffffu, p, 31);
    if (s < S) {
      const double w = (1.0 - one_m) * T;
      double G = -gsum + gD * ((double)z - (double)F.far) + gA;
      if (D.g_weights) G += (double)D.g_weights[row + s];
      if (B.weights[row + s] > F.weight_thres) {   // appearance membership as decided in the forward
        const float* cc = B.rgb + (row + s) * 3;
        G += (double)e[0] * cc[0] + (double)e[1] * cc[1] + (double)e[2] * cc[2];
      }
      om[s] = one_m;
      GT[s] = G * T;
      Gw[s] = G * w;
    }
  }
  __syncwarp();
  // reverse sweep: R_i = sum_{j>i} G_j w_j
  double tail = 0.0;
  for (int c = n_it - 1; c >= 0; --c) {
    const int s = c * 32 + lane;
    const double v = (s < S) ? Gw[s] : 0.0;
    double q = v;  // inclusive suffix sum within the chunk
#pragma unroll
    for (int o2 = 1; o2 < 32; o2 <<= 1) {
      const double t = __shfl_down_sync(0xffffffffu, q, o2);
      if (lane + o2 < 32) q += t;
    }
    const double R = tail + (q - v);
    tail += __shfl_sync(0xffffffffu, q, 0);
    if (s < S) {
      float gs = 0.f;
      // samples behind the point where the forward pass stopped evaluating the ray (early termination: the
      // transmittance is exactly 0 there) have no gradient
      if (B.valid[row + s] && (B.ray_term == nullptr || s < B.ray_term[ray])) {
        const double one_m = om[s];
        const double galpha = GT[s] - R / (one_m + 1e-10);
        float dist = 0.f;
        if (s + 1 < S)
          dist = __fsub_rn(sample_z(tmin, F.step_size, s + 1, u, train),
                           sample_z(tmin, F.step_size, s, u, train));
        gs = (float)(galpha * (double)__fmul_rn(dist, F.distance_scale) * one_m);
      }
      D.g_sigma[row + s] = gs;
    }
  }
}

// d sigma / d feature from the saved density (softplus: sigmoid(x) = 1 - exp(-softplus(x)))
__device__ __forceinline__ float dsigma_dfeat(const NvfiField& F, float sigma, float feat) {
  if (F.fea2dense_act == NVFI_ACT_SOFTPLUS) return -expm1f(-sigma);
  if (F.fea2dense_act == NVFI_ACT_RELU) return feat > 0.f ? 1.f : 0.f;
  return feat > 0.f ? 1.f : (feat < 0.f ? -1.f : 0.f);
}

// ---------------------------------------------------------------------------------------
// k_density_bwd: one warp per ray, 8-lane groups per valid sample.
// g_x_adv[sample] = (appearance part, already written for w > thres) + density part.
// ---------------------------------------------------------------------------------------
#ifndef NVFI_DBWD_BLOCKS
#define NVFI_DBWD_BLOCKS 3   // measured against 2 and 4 resident blocks per SM
#endif
__global__ void __launch_bounds__(256, NVFI_DBWD_BLOCKS)
    k_density_bwd(const NvfiField F, const NvfiRenderArgs A, const NvfiRenderBuffers B,
                  const NvfiRenderGrads D, int S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ray = (long long)blockIdx.x * 8 + warp;
  if (ray >= A.n_rays) return;
  const int g = lane >> 3, l8 = lane & 7;
  const long long row = ray * S;
  const int n_it = (S + 31) / 32;
  const int R4 = F.rd >> 2;
  float* gsp[3] = {D.g_dplane_space[0], D.g_dplane_space[1], D.g_dplane_space[2]};
  float* gtp[3] = {D.g_dplane_time[0], D.g_dplane_time[1], D.g_dplane_time[2]};
  for (int c = 0; c < n_it; ++c) {
    const int s = c * 32 + lane;
    const bool v = (s < S) && (B.valid[row + s] != 0);
    // Only samples with a non-zero dL/dsigma get one of the warp's 4 gather slots: behind a saturated surface
    // the gradient is exactly zero (44 % of the valid samples of the bench frame), and such a sample would
    // keep an 8-lane group idle for a whole round.  Its dL/dx is written here (coalesced loads of g_sigma).
    const float gs_own = v ? D.g_sigma[row + s] : 0.f;
    const bool p = v && gs_own != 0.f;
    if (v && !p) {
      const bool is_app = (D.g_rgb != nullptr) && (B.weights[row + s] > F.weight_thres);
      if (!is_app) {   // an appearance sample keeps the gradient k_app_bwd wrote
        D.g_x_adv[(row + s) * 3 + 0] = 0.f;
        D.g_x_adv[(row + s) * 3 + 1] = 0.f;
        D.g_x_adv[(row + s) * 3 + 2] = 0.f;
      }
    }
    unsigned m = __ballot_sync(0xffffffffu, p);
    while (m) {
      unsigned mm = m;
      int mine = -1;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int b = mm ? (__ffs(mm) - 1) : -1;
        if (k == g) mine = b;
        mm &= mm - 1;
      }
      m = mm;
      const float gsig = __shfl_sync(0xffffffffu, gs_own, mine >= 0 ? mine : 0);
      if (mine >= 0) {
        const long long gi = row + c * 32 + mine;
        const float xt[4] = {__ldg(B.x_adv + gi * 3 + 0), __ldg(B.x_adv + gi * 3 + 1),
                             __ldg(B.x_adv + gi * 3 + 2), A.t_norm_base};
        const float sigma = B.sigma[gi];
        float feat = 0.f;
        if (F.fea2dense_act != NVFI_ACT_SOFTPLUS) feat = density_feature_group(F, xt, l8);
        const float gF = gsig * dsigma_dfeat(F, sigma, feat);
        float gx[3] = {0.f, 0.f, 0.f};
        if (gF != 0.f) {
          if (R4 <= 8) {
            float4 gch[1] = {make_float4(gF, gF, gF, gF)};
            kplanes_backward<1>(F, F.dplane_space, F.dplane_time, gsp, gtp, F.rd, xt, l8, gch, gx);
          } else {
            float4 gch[2] = {make_float4(gF, gF, gF, gF), make_float4(gF, gF, gF, gF)};
            kplanes_backward<2>(F, F.dplane_space, F.dplane_time, gsp, gtp, F.rd, xt, l8, gch, gx);
          }
        }
        gx[0] = group8_sum(gx[0]);
        gx[1] = group8_sum(gx[1]);
        gx[2] = group8_sum(gx[2]);
        if (l8 < 3) {
          const float mine_g = (l8 == 0) ? gx[0] : ((l8 == 1) ? gx[1] : gx[2]);
          const bool is_app = (D.g_rgb != nullptr) && (B.weights[gi] > F.weight_thres);
          const float prev = is_app ? D.g_x_adv[gi * 3 + l8] : 0.f;
          D.g_x_adv[gi * 3 + l8] = prev + mine_g;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// k_app_bwd
// ---------------------------------------------------------------------------------------
struct AppBwdTile {
  float x[3][NVFI_TM];
  float d[3][NVFI_TM];
  float gc[3][NVFI_TM];    // dL/d rgb_i = w_i * g_rgb_eff[ray]
  float gpts[3][NVFI_TM];  // dL/dx through the MLP_PE position inputs
  float gout[8][NVFI_TM];
  int gidx[NVFI_TM];
  int q_idx[NVFI_QCAP];
  int warp_cnt[2][NVFI_THREADS / 32];
  int batch;
};

__device__ __forceinline__ void sh_bases_deg2_b(float x, float y, float z, float sh[9]) {
  const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  sh[0] = C0;
  sh[1] = -C1 * y;
  sh[2] = C1 * z;
  sh[3] = -C1 * x;
  sh[4] = 1.0925484305920792f * xy;
  sh[5] = -1.0925484305920792f * yz;
  sh[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
  sh[7] = -1.0925484305920792f * xz;
  sh[8] = 0.5462742152960396f * (xx - yy);
}

// Rows [app_dim, k_pad) of the MLP_PE input, BRS-strided variant of mlp_pe_inputs_tile.
__device__ void mlp_pe_inputs_tile_b(const NvfiField& F, const AppBwdTile& T,
                                     float* __restrict__ At) {
  const int tid = threadIdx.x;
  const int m = tid & 127, part = tid >> 7;
  const int b0 = F.app_dim, pp = F.pos_pe, vp = F.view_pe;
  const int row_pe_p = b0 + 6, row_pe_v = row_pe_p + 6 * pp, row_end = row_pe_v + 6 * vp;
  if (part == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float p = T.x[a][m];
      At[(b0 + 3 + a) * BRS + m] = p;
      float f = 1.f;
      for (int k = 0; k < pp; ++k) {
        float s, c;
        sincosf(p * f, &s, &c);
        At[(row_pe_p + a * pp + k) * BRS + m] = s;
        At[(row_pe_p + 3 * pp + a * pp + k) * BRS + m] = c;
        f *= 2.f;
      }
    }
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float p = T.d[a][m];
      At[(b0 + a) * BRS + m] = p;
      float f = 1.f;
      for (int k = 0; k < vp; ++k) {
        float s, c;
        sincosf(p * f, &s, &c);
        At[(row_pe_v + a * vp + k) * BRS + m] = s;
        At[(row_pe_v + 3 * vp + a * vp + k) * BRS + m] = c;
        f *= 2.f;
      }
    }
    for (int r = row_end; r < NVFI_TM; ++r) At[r * BRS + m] = 0.f;
  }
  __syncthreads();
}

__device__ void app_bwd_tile(const NvfiField& F, const NvfiRenderArgs& A,
                             const NvfiRenderBuffers& B, const NvfiRenderGrads& D, AppBwdTile& T,
                             int n, float* __restrict__ At, float* __restrict__ Gt,
                             float* __restrict__ wS, float* __restrict__ ws) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 3, l8 = lane & 7;
  const int R4 = F.ra >> 2;
  float* outS = wS;
  // 1. gather the Ra-component features -> At rows [0, Ra); stash them
#pragma unroll 1
  for (int pass = 0; pass < NVFI_TM / 32; ++pass) {
    const int m = pass * 32 + warp * 4 + g;
    float4 pr[2];
    pr[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    pr[1] = pr[0];
    if (m < n) {
      const float xt[4] = {T.x[0][m], T.x[1][m], T.x[2][m], A.t_norm_base};
      kplanes_features<2>(F, F.aplane_space, F.aplane_time, F.ra, xt, l8, pr);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int q = l8 + 8 * j;
      if (q < R4) {
        At[(4 * q + 0) * BRS + m] = pr[j].x;
        At[(4 * q + 1) * BRS + m] = pr[j].y;
        At[(4 * q + 2) * BRS + m] = pr[j].z;
        At[(4 * q + 3) * BRS + m] = pr[j].w;
      }
    }
  }
  __syncthreads();
  store_stash(At, ws + AW_STASH_F48, F.ra);
  // 2. basis_mat
  tile_linear_small<16, BRS>(At, outS, F.basis_mat);
  if (F.shading_mode == NVFI_SHADING_SH) {
    // rgb = relu(sum_b Y_b feat[ch*9+b] + .5): g_feat -> Gt rows [0, 27)
    if (tid < NVFI_TM) {
      const int m = tid;
      float sh[9];
      sh_bases_deg2_b(T.d[0][m], T.d[1][m], T.d[2][m], sh);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float acc = 0.f;
#pragma unroll
        for (int b = 0; b < 9; ++b) acc = fmaf(sh[b], outS[(ch * 9 + b) * NVFI_TM + m], acc);
        const float gch = (m < n && acc + 0.5f > 0.f) ? T.gc[ch][m] : 0.f;
#pragma unroll
        for (int b = 0; b < 9; ++b) Gt[(ch * 9 + b) * BRS + m] = gch * sh[b];
      }
      T.gpts[0][m] = T.gpts[1][m] = T.gpts[2][m] = 0.f;
    }
    __syncthreads();
  } else {
    for (int i = tid; i < F.app_dim * NVFI_TM; i += NVFI_THREADS) {
      const int r = i >> 7, m = i & 127;
      At[r * BRS + m] = outS[i];
    }
    mlp_pe_inputs_tile_b(F, T, At);
    store_stash(At, ws + AW_STASH_X);
    tile_linear128<ACT_RELU, BRS, true>(At, wS, F.render_mlp[0], ws + AW_STASH_H0);
    tile_linear128<ACT_RELU, BRS, true>(At, wS, F.render_mlp[1], ws + AW_STASH_H1);
    tile_linear_small<2, BRS>(At, outS, F.render_mlp[2]);
    if (tid < NVFI_TM) {
      const int m = tid;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float sg = sigmoid_f(outS[ch * NVFI_TM + m]);
        T.gout[ch][m] = (m < n) ? T.gc[ch][m] * sg * (1.f - sg) : 0.f;
      }
      T.gout[3][m] = 0.f;
    }
    __syncthreads();
    // At still holds a1 = relu(h1): last layer
    small_layer_bwd(At, &T.gout[0][0], F.render_mlp[2], Gt, ws + AW_W2, ws + AW_B(2));
    apply_act_grad<ACT_RELU>(ws + AW_STASH_H1, Gt);
    // layer 1
    load_stash_act<ACT_RELU>(ws + AW_STASH_H0, At);
    tile_rowsum_acc(Gt, ws + AW_B(1));
    tile_outer_acc<8>(At, Gt, ws + AW_W1);
    {
      NvfiLinear Lb = F.render_mlp[1];
      Lb.wt = F.render_mlp[1].w_rows;
      Lb.bias = nullptr;
      Lb.k_pad = 128;
      tile_linear128<ACT_NONE, BRS>(Gt, wS, Lb);
    }
    apply_act_grad<ACT_RELU>(ws + AW_STASH_H0, Gt);
    // layer 0
    load_stash_act<ACT_NONE>(ws + AW_STASH_X, At);
    tile_rowsum_acc(Gt, ws + AW_B(0));
    tile_outer_acc<8>(At, Gt, ws + AW_W0);
    {
      NvfiLinear Lb = F.render_mlp[0];
      Lb.wt = F.render_mlp[0].w_rows;
      Lb.bias = nullptr;
      Lb.k_pad = 128;
      tile_linear128<ACT_NONE, BRS>(Gt, wS, Lb);
    }
    // Gt rows = dL/d(mlp input).  Position gradient incl. the PE chain rule.
    if (tid < NVFI_TM) {
      const int m = tid;
      const int b0 = F.app_dim, pp = F.pos_pe;
      const int row_pe_p = b0 + 6;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float p = T.x[a][m];
        float acc = Gt[(b0 + 3 + a) * BRS + m];
        float f = 1.f;
        for (int k = 0; k < pp; ++k) {
          float s, c;
          sincosf(p * f, &s, &c);
          acc += f * (Gt[(row_pe_p + a * pp + k) * BRS + m] * c -
                      Gt[(row_pe_p + 3 * pp + a * pp + k) * BRS + m] * s);
          f *= 2.f;
        }
        T.gpts[a][m] = (m < n) ? acc : 0.f;
      }
    }
    __syncthreads();
  }
  // 4. basis_mat backward: dB and g_F48 -> Gt rows [64, 64 + Ra)
  load_stash_act<ACT_NONE>(ws + AW_STASH_F48, At, F.ra);
  {
    const int E = F.app_dim * F.ra;  // <= 2048
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid * 8 + i;
      if (e < E) {
        const int nn = e / F.ra, c = e - nn * F.ra;
        float s = 0.f;
        for (int m = 0; m < NVFI_TM; ++m) s = fmaf(Gt[nn * BRS + m], At[c * BRS + m], s);
        acc[i] = s;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) ws[AW_BASIS + tid * 8 + i] += acc[i];
  }
  {
    const int m = tid & 127, half = tid >> 7;
    const int ch = F.ra >> 1;
    const int np = F.basis_mat.n_pad;
    for (int c = half * ch; c < (half + 1) * ch; ++c) {
      const float* wr = F.basis_mat.wt + (size_t)c * np;
      float s = 0.f;
      for (int nn = 0; nn < F.app_dim; ++nn) s = fmaf(__ldg(wr + nn), Gt[nn * BRS + m], s);
      Gt[(64 + c) * BRS + m] = s;
    }
  }
  __syncthreads();
  // 5. planes + position
  float* gsp[3] = {D.g_aplane_space[0], D.g_aplane_space[1], D.g_aplane_space[2]};
  float* gtp[3] = {D.g_aplane_time[0], D.g_aplane_time[1], D.g_aplane_time[2]};
#pragma unroll 1
  for (int pass = 0; pass < NVFI_TM / 32; ++pass) {
    const int m = pass * 32 + warp * 4 + g;
    if (m < n) {
      const float xt[4] = {T.x[0][m], T.x[1][m], T.x[2][m], A.t_norm_base};
      float4 gch[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int q = l8 + 8 * j;
        gch[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < R4) {
          gch[j].x = Gt[(64 + 4 * q + 0) * BRS + m];
          gch[j].y = Gt[(64 + 4 * q + 1) * BRS + m];
          gch[j].z = Gt[(64 + 4 * q + 2) * BRS + m];
          gch[j].w = Gt[(64 + 4 * q + 3) * BRS + m];
        }
      }
      float gx[3];
      kplanes_backward<2>(F, F.aplane_space, F.aplane_time, gsp, gtp, F.ra, xt, l8, gch, gx);
      gx[0] = group8_sum(gx[0]);
      gx[1] = group8_sum(gx[1]);
      gx[2] = group8_sum(gx[2]);
      if (l8 < 3) {
        const float v = (l8 == 0) ? gx[0] : ((l8 == 1) ? gx[1] : gx[2]);
        D.g_x_adv[(long long)T.gidx[m] * 3 + l8] = v + T.gpts[l8][m];
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(NVFI_THREADS, 1)
    k_app_bwd(const NvfiField F, const NvfiRenderArgs A, const NvfiRenderBuffers B,
              const NvfiRenderGrads D, int S, long long total, int n_batches) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* At = reinterpret_cast<float*>(smem_raw);
  float* Gt = At + TILE_F;
  float* wS = Gt + TILE_F;
  AppBwdTile& T = *reinterpret_cast<AppBwdTile*>(wS + 2 * NVFI_KC * 128);
  float* ws = D.workspace + (size_t)blockIdx.x * WS_CTA_F;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < AW_PART_F; i += NVFI_THREADS) ws[i] = 0.f;
  __syncthreads();

  int sub = NVFI_SUBS;
  long long batch_base = 0;
  bool exhausted = false;
  int qc = 0, par = 0;
  unsigned long long n_done = 0;
  for (;;) {
    while (qc < NVFI_TM && !exhausted) {
      if (sub == NVFI_SUBS) {
        if (tid == 0) T.batch = atomicAdd(&B.counters[2], 1);
        __syncthreads();
        const int b = T.batch;
        __syncthreads();
        if (b >= n_batches) {
          exhausted = true;
          break;
        }
        batch_base = (long long)b * (NVFI_SUBS * NVFI_THREADS);
        sub = 0;
      }
      const long long idx = batch_base + (long long)sub * NVFI_THREADS + tid;
      ++sub;
      const bool push = (idx < total) && (B.weights[idx] > F.weight_thres);
      const unsigned bal = __ballot_sync(0xffffffffu, push);
      if (lane == 0) T.warp_cnt[par][warp] = __popc(bal);
      const int tot = __syncthreads_count(push);
      if (push) {
        int pos = qc + __popc(bal & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) pos += T.warp_cnt[par][w];
        T.q_idx[pos] = (int)idx;
      }
      qc += tot;
      par ^= 1;
    }
    if (qc == 0) break;
    __syncthreads();
    const int n = min(NVFI_TM, qc);
    const int start = qc - n;
    qc = start;
    if (tid < NVFI_TM) {
      const bool live = tid < n;
      const long long gi = live ? T.q_idx[start + tid] : 0;
      const long long ray = gi / S;
      T.gidx[tid] = (int)gi;
      const float w = live ? B.weights[gi] : 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        T.x[a][tid] = live ? B.x_adv[gi * 3 + a] : 0.f;
        T.d[a][tid] = live ? __ldg(A.rays_d + ray * 3 + a) : 0.f;
        T.gc[a][tid] = live ? w * D.g_rgb_eff[ray * 3 + a] : 0.f;
      }
    }
    __syncthreads();
    app_bwd_tile(F, A, B, D, T, n, At, Gt, wS, ws);
    n_done += n;
  }
  if (tid == 0 && n_done)  // samples back-propagated (bench.py's algorithmic flop count)
    atomicAdd(reinterpret_cast<unsigned long long*>(B.counters + 8), n_done);
}

// ---------------------------------------------------------------------------------------
// k_advect_bwd: RK2 adjoint through the velocity MLP
// ---------------------------------------------------------------------------------------
struct AdvBwdTile {
  float x0[3][NVFI_TM];
  float xm[3][NVFI_TM];
  float gbar[3][NVFI_TM];
  float gm[3][NVFI_TM];
  float w0[8][NVFI_TM];
  float w1[8][NVFI_TM];
  float gout[8][NVFI_TM];
  unsigned char gate0[NVFI_TM], gate1[NVFI_TM], reverted[NVFI_TM];
  int gidx[NVFI_TM];
  int q_idx[NVFI_QCAP];
  int warp_cnt[2][NVFI_THREADS / 32];
  int batch;
};

// backward through the 6-layer weight net for one evaluation whose pre-activations are in
// `stash` (5 layers).  On entry T.gout holds dL/d(basis weights).  On exit outS[r][m]
// (stride 128, r < 32) holds dL/d(encoded input).
__device__ void vel_net_bwd_tile(const NvfiField& F, AdvBwdTile& T, float* __restrict__ At,
                                 float* __restrict__ Gt, float* __restrict__ wS,
                                 float* __restrict__ ws, const float* __restrict__ stash,
                                 const float* xs, const float* ys, const float* zs, float tval) {
  // last layer: input a4 = silu(h4)
  load_stash_act<ACT_SILU>(stash + 4 * STASH_F, At);
  small_layer_bwd(At, &T.gout[0][0], F.vel_net[5], Gt, ws + VW_W5, ws + VW_B(5));
  apply_act_grad<ACT_SILU>(stash + 4 * STASH_F, Gt);
#pragma unroll 1
  for (int l = 4; l >= 1; --l) {
    load_stash_act<ACT_SILU>(stash + (size_t)(l - 1) * STASH_F, At);
    tile_rowsum_acc(Gt, ws + VW_B(l));
    tile_outer_acc<8>(At, Gt, ws + VW_W(l));
    NvfiLinear Lb = F.vel_net[l];
    Lb.wt = F.vel_net[l].w_rows;
    Lb.bias = nullptr;
    Lb.k_pad = 128;
    tile_linear128<ACT_NONE, BRS>(Gt, wS, Lb);
    apply_act_grad<ACT_SILU>(stash + (size_t)(l - 1) * STASH_F, Gt);
  }
  // layer 0: input = encoding (recomputed)
  {
    // a per-thread constant time array is not available: encode with a broadcast time
    const int tid = threadIdx.x;
    const int m = tid & 127, part = tid >> 7;
    const float q[4] = {xs[m], ys[m], zs[m], tval};
    if (part == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        At[i * BRS + m] = q[i];
        float s, c;
        sincosf(q[i], &s, &c);
        At[(4 + i) * BRS + m] = s;
        At[(8 + i) * BRS + m] = c;
        At[(12 + i) * BRS + m] = sinf(q[i] * 2.f);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        At[(16 + i) * BRS + m] = cosf(q[i] * 2.f);
        float s, c;
        sincosf(q[i] * 4.f, &s, &c);
        At[(20 + i) * BRS + m] = s;
        At[(24 + i) * BRS + m] = c;
        At[(28 + i) * BRS + m] = 0.f;
      }
    }
    __syncthreads();
  }
  tile_rowsum_acc(Gt, ws + VW_B(0));
  tile_outer_acc<2>(At, Gt, ws + VW_W0);
  {
    NvfiLinear Lb;
    Lb.wt = F.vel_net[0].w_rows;  // (128, 32)
    Lb.bias = nullptr;
    Lb.w_rows = nullptr;
    Lb.in_dim = 128;
    Lb.out_dim = 32;
    Lb.k_pad = 128;
    Lb.n_pad = 32;
    tile_linear_small<16, BRS>(Gt, wS, Lb);  // -> wS[r][m], r < 32
  }
}

// dL/dq_i from dL/d(encoding) (encoder tangents)
__device__ __forceinline__ float pe_chain(const float* __restrict__ genc, int i, int m, float q) {
  float s1, c1, s2, c2, s4, c4;
  sincosf(q, &s1, &c1);
  sincosf(q * 2.f, &s2, &c2);
  sincosf(q * 4.f, &s4, &c4);
  return genc[i * NVFI_TM + m] + genc[(4 + i) * NVFI_TM + m] * c1 - genc[(8 + i) * NVFI_TM + m] * s1 +
         2.f * (genc[(12 + i) * NVFI_TM + m] * c2 - genc[(16 + i) * NVFI_TM + m] * s2) +
         4.f * (genc[(20 + i) * NVFI_TM + m] * c4 - genc[(24 + i) * NVFI_TM + m] * s4);
}

// v = basis(w, x): dL/dw and the explicit dL/dx from dL/dv
__device__ __forceinline__ void basis_velocity_bwd(const float w[6], float x, float y, float z,
                                                   const float gv[3], float gw[6], float gxe[3]) {
  gw[0] = gv[0];
  gw[1] = gv[1];
  gw[2] = gv[2];
  gw[3] = gv[1] * z - gv[2] * y;
  gw[4] = -gv[0] * z + gv[2] * x;
  gw[5] = gv[0] * y - gv[1] * x;
  gxe[0] = -w[5] * gv[1] + w[4] * gv[2];
  gxe[1] = w[5] * gv[0] - w[3] * gv[2];
  gxe[2] = -w[4] * gv[0] + w[3] * gv[1];
}

__global__ void __launch_bounds__(NVFI_THREADS, 1)
    k_advect_bwd(const NvfiField F, const NvfiRenderArgs A, const NvfiRenderBuffers B,
                 const NvfiRenderGrads D, int S, long long total, int n_batches) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* At = reinterpret_cast<float*>(smem_raw);
  float* Gt = At + TILE_F;
  float* wS = Gt + TILE_F;
  AdvBwdTile& T = *reinterpret_cast<AdvBwdTile*>(wS + 2 * NVFI_KC * 128);
  float* ws = D.workspace + (size_t)blockIdx.x * WS_CTA_F;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < VW_PART_F; i += NVFI_THREADS) ws[i] = 0.f;
  __syncthreads();

  // uniform RK2 schedule of this render call (models/tensorf_keyframe.py:577-609)
  float sched_dt[MAX_RK2_STEPS], sched_t[MAX_RK2_STEPS];
  int n_steps = 0;
  {
    float off = __fsub_rn(A.t, A.base_time), tc = A.t;
    while (fabsf(off) > 0.f && n_steps < MAX_RK2_STEPS) {
      float dt = fminf(fabsf(off), F.dt_max);
      dt = (off > 0.f) ? dt : -dt;
      sched_dt[n_steps] = dt;
      sched_t[n_steps] = tc;
      off = __fsub_rn(off, dt);
      tc = __fsub_rn(tc, dt);
      ++n_steps;
    }
  }

  int sub = NVFI_SUBS;
  long long batch_base = 0;
  bool exhausted = false;
  int qc = 0, par = 0;
  unsigned long long n_done = 0;
  for (;;) {
    while (qc < NVFI_TM && !exhausted) {
      if (sub == NVFI_SUBS) {
        if (tid == 0) T.batch = atomicAdd(&B.counters[3], 1);
        __syncthreads();
        const int b = T.batch;
        __syncthreads();
        if (b >= n_batches) {
          exhausted = true;
          break;
        }
        batch_base = (long long)b * (NVFI_SUBS * NVFI_THREADS);
        sub = 0;
      }
      const long long idx = batch_base + (long long)sub * NVFI_THREADS + tid;
      ++sub;
      bool push = false;
      if (idx < total && B.valid[idx]) {
        const float g0 = D.g_x_adv[idx * 3], g1 = D.g_x_adv[idx * 3 + 1],
                    g2 = D.g_x_adv[idx * 3 + 2];
        push = (g0 != 0.f) | (g1 != 0.f) | (g2 != 0.f);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, push);
      if (lane == 0) T.warp_cnt[par][warp] = __popc(bal);
      const int tot = __syncthreads_count(push);
      if (push) {
        int pos = qc + __popc(bal & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) pos += T.warp_cnt[par][w];
        T.q_idx[pos] = (int)idx;
      }
      qc += tot;
      par ^= 1;
    }
    if (qc == 0) break;
    __syncthreads();
    const int n = min(NVFI_TM, qc);
    const int start = qc - n;
    qc = start;
    n_done += n;
    // ---- load the tile: start position (sampler recompute) and upstream gradient
    if (tid < NVFI_TM) {
      const bool live = tid < n;
      const long long gi = live ? T.q_idx[start + tid] : 0;
      T.gidx[tid] = (int)gi;
      float xn[3] = {0.f, 0.f, 0.f};
      if (live) {
        const long long ray = gi / S;
        const int s = (int)(gi - ray * S);
        const float o[3] = {__ldg(A.rays_o + ray * 3), __ldg(A.rays_o + ray * 3 + 1),
                            __ldg(A.rays_o + ray * 3 + 2)};
        const float d[3] = {__ldg(A.rays_d + ray * 3), __ldg(A.rays_d + ray * 3 + 1),
                            __ldg(A.rays_d + ray * 3 + 2)};
        const bool inside = B.chunk_inside[ray / A.ray_chunk] != 0;
        const float tmin = ray_tmin(F, o, d, inside);
        const bool train = A.jitter != nullptr;
        const float u = train ? __ldg(A.jitter + ray) : 0.f;
        sample_point(F, o, d, sample_z(tmin, F.step_size, s, u, train), xn);
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        T.x0[a][tid] = xn[a];
        T.gbar[a][tid] = live ? D.g_x_adv[gi * 3 + a] : 0.f;
      }
    }
    __syncthreads();
    // ---- forward sweep over all but the last step, remembering the step start positions
    float* xsteps = ws + VW_XSTEPS;
    for (int k = 0; k + 1 < n_steps; ++k) {
      if (tid < NVFI_TM) {
#pragma unroll
        for (int a = 0; a < 3; ++a) xsteps[(k * 3 + a) * NVFI_TM + tid] = T.x0[a][tid];
      }
      const float dt = sched_dt[k], tc = sched_t[k], hdt = 0.5f * dt;
      // eval 1 (the time row is broadcast: reuse gm[0] as a scratch time vector)
      if (tid < NVFI_TM) T.gm[0][tid] = tc;
      __syncthreads();
      vel_net_tile<ACT_SILU, BRS>(F.vel_net, At, wS, &T.w0[0][0], T.x0[0], T.x0[1], T.x0[2],
                                  T.gm[0]);
      if (tid < NVFI_TM) {
        const int m = tid;
        const float x = T.x0[0][m], y = T.x0[1][m], z = T.x0[2][m];
        float v[3] = {0.f, 0.f, 0.f};
        if (!gate_outside(F, x, y, z)) {
          const float w[6] = {T.w0[0][m], T.w0[1][m], T.w0[2][m], T.w0[3][m], T.w0[4][m], T.w0[5][m]};
          basis_velocity(w, x, y, z, v);
        }
        T.xm[0][m] = __fsub_rn(x, __fmul_rn(hdt, v[0]));
        T.xm[1][m] = __fsub_rn(y, __fmul_rn(hdt, v[1]));
        T.xm[2][m] = __fsub_rn(z, __fmul_rn(hdt, v[2]));
        T.gm[0][m] = __fsub_rn(tc, hdt);
      }
      __syncthreads();
      vel_net_tile<ACT_SILU, BRS>(F.vel_net, At, wS, &T.w1[0][0], T.xm[0], T.xm[1], T.xm[2],
                                  T.gm[0]);
      if (tid < NVFI_TM) {
        const int m = tid;
        const float xm = T.xm[0][m], ym = T.xm[1][m], zm = T.xm[2][m];
        float v[3] = {0.f, 0.f, 0.f};
        if (!gate_outside(F, xm, ym, zm)) {
          const float w[6] = {T.w1[0][m], T.w1[1][m], T.w1[2][m], T.w1[3][m], T.w1[4][m], T.w1[5][m]};
          basis_velocity(w, xm, ym, zm, v);
        }
        const float x = T.x0[0][m], y = T.x0[1][m], z = T.x0[2][m];
        float nx = __fsub_rn(x, __fmul_rn(dt, v[0]));
        float ny = __fsub_rn(y, __fmul_rn(dt, v[1]));
        float nz = __fsub_rn(z, __fmul_rn(dt, v[2]));
        if (F.vel_gate == NVFI_GATE_SUR && gate_outside(F, nx, ny, nz)) {
          nx = x;
          ny = y;
          nz = z;
        }
        T.x0[0][m] = nx;
        T.x0[1][m] = ny;
        T.x0[2][m] = nz;
      }
      __syncthreads();
    }
    // ---- reverse sweep
    for (int k = n_steps - 1; k >= 0; --k) {
      if (k < n_steps - 1) {
        if (tid < NVFI_TM) {
#pragma unroll
          for (int a = 0; a < 3; ++a) T.x0[a][tid] = xsteps[(k * 3 + a) * NVFI_TM + tid];
        }
        __syncthreads();
      }
      const float dt = sched_dt[k], tc = sched_t[k], hdt = 0.5f * dt;
      const float tmid = __fsub_rn(tc, hdt);
      float* stash0 = ws + VW_STASH;
      float* stash1 = ws + VW_STASH + 5 * STASH_F;
      // recompute eval 1 with stash
      if (tid < NVFI_TM) T.gm[0][tid] = tc;
      __syncthreads();
      vel_net_tile<ACT_SILU, BRS, true>(F.vel_net, At, wS, &T.w0[0][0], T.x0[0], T.x0[1], T.x0[2],
                                        T.gm[0], stash0);
      if (tid < NVFI_TM) {
        const int m = tid;
        const float x = T.x0[0][m], y = T.x0[1][m], z = T.x0[2][m];
        float v[3] = {0.f, 0.f, 0.f};
        const bool out0 = gate_outside(F, x, y, z);
        T.gate0[m] = out0;
        if (!out0) {
          const float w[6] = {T.w0[0][m], T.w0[1][m], T.w0[2][m], T.w0[3][m], T.w0[4][m], T.w0[5][m]};
          basis_velocity(w, x, y, z, v);
        }
        T.xm[0][m] = __fsub_rn(x, __fmul_rn(hdt, v[0]));
        T.xm[1][m] = __fsub_rn(y, __fmul_rn(hdt, v[1]));
        T.xm[2][m] = __fsub_rn(z, __fmul_rn(hdt, v[2]));
        T.gm[0][m] = tmid;
      }
      __syncthreads();
      vel_net_tile<ACT_SILU, BRS, true>(F.vel_net, At, wS, &T.w1[0][0], T.xm[0], T.xm[1], T.xm[2],
                                        T.gm[0], stash1);
      // adjoint of x1 = x0 - dt v1(m)
      if (tid < NVFI_TM) {
        const int m = tid;
        const float xm = T.xm[0][m], ym = T.xm[1][m], zm = T.xm[2][m];
        const bool out1 = gate_outside(F, xm, ym, zm);
        const float w[6] = {T.w1[0][m], T.w1[1][m], T.w1[2][m], T.w1[3][m], T.w1[4][m], T.w1[5][m]};
        float v[3] = {0.f, 0.f, 0.f};
        if (!out1) basis_velocity(w, xm, ym, zm, v);
        const float x = T.x0[0][m], y = T.x0[1][m], z = T.x0[2][m];
        const float nx = __fsub_rn(x, __fmul_rn(dt, v[0]));
        const float ny = __fsub_rn(y, __fmul_rn(dt, v[1]));
        const float nz = __fsub_rn(z, __fmul_rn(dt, v[2]));
        const bool rev = (F.vel_gate == NVFI_GATE_SUR) && gate_outside(F, nx, ny, nz);
        T.gate1[m] = out1;
        T.reverted[m] = rev;
        float gw[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gxe[3] = {0.f, 0.f, 0.f};
        if (!rev && !out1) {
          const float gv[3] = {-dt * T.gbar[0][m], -dt * T.gbar[1][m], -dt * T.gbar[2][m]};
          basis_velocity_bwd(w, xm, ym, zm, gv, gw, gxe);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) T.gout[i][m] = gw[i];
        T.gout[6][m] = T.gout[7][m] = 0.f;
        T.gm[0][m] = gxe[0];
        T.gm[1][m] = gxe[1];
        T.gm[2][m] = gxe[2];
      }
      __syncthreads();
      vel_net_bwd_tile(F, T, At, Gt, wS, ws, stash1, T.xm[0], T.xm[1], T.xm[2], tmid);
      // adjoint of m = x0 - dt/2 v0(x0)
      if (tid < NVFI_TM) {
        const int m = tid;
        float gmv[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) gmv[a] = T.gm[a][m] + pe_chain(wS, a, m, T.xm[a][m]);
        float gx0[3] = {T.gbar[0][m] + gmv[0], T.gbar[1][m] + gmv[1], T.gbar[2][m] + gmv[2]};
        float gw[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (!T.gate0[m]) {
          const float w[6] = {T.w0[0][m], T.w0[1][m], T.w0[2][m], T.w0[3][m], T.w0[4][m], T.w0[5][m]};
          const float gv[3] = {-hdt * gmv[0], -hdt * gmv[1], -hdt * gmv[2]};
          float gxe[3];
          basis_velocity_bwd(w, T.x0[0][m], T.x0[1][m], T.x0[2][m], gv, gw, gxe);
          gx0[0] += gxe[0];
          gx0[1] += gxe[1];
          gx0[2] += gxe[2];
        }
        T.gbar[0][m] = gx0[0];
        T.gbar[1][m] = gx0[1];
        T.gbar[2][m] = gx0[2];
#pragma unroll
        for (int i = 0; i < 6; ++i) T.gout[i][m] = gw[i];
      }
      __syncthreads();
      vel_net_bwd_tile(F, T, At, Gt, wS, ws, stash0, T.x0[0], T.x0[1], T.x0[2], tc);
      if (tid < NVFI_TM) {
        const int m = tid;
#pragma unroll
        for (int a = 0; a < 3; ++a) T.gbar[a][m] += pe_chain(wS, a, m, T.x0[a][m]);
      }
      __syncthreads();
    }
  }
  if (tid == 0 && n_done)
    atomicAdd(reinterpret_cast<unsigned long long*>(B.counters + 10), n_done);
}

// ---------------------------------------------------------------------------------------
// reductions of the per-CTA partials into the packed gradients
// ---------------------------------------------------------------------------------------
// tile_outer layout -> packed W^T gradient (k_pad, 128)
__global__ void k_reduce_outer(const float* __restrict__ ws, int n_cta, int off, int ki,
                               float* __restrict__ g_wt) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ki * 2048) return;
  float s = 0.f;
  for (int c = 0; c < n_cta; ++c) s += ws[(size_t)c * WS_CTA_F + off + e];
  const int q = e >> 10, tid = (e & 1023) >> 2, jl = e & 3;
  const int i = q >> 1, j = (q & 1) * 4 + jl;
  const int k = (tid & 15) + 16 * i, n = (tid >> 4) + 16 * j;
  g_wt[k * 128 + n] += s;
}
// small-layer layout -> packed (128, n_pad)
__global__ void k_reduce_small(const float* __restrict__ ws, int n_cta, int off, int n_pad,
                               float* __restrict__ g_wt) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 1024) return;
  const int tid = e >> 2, j = e & 3;
  const int nq = n_pad >> 1;
  if (j >= nq) return;
  float s = 0.f;
  for (int c = 0; c < n_cta; ++c) s += ws[(size_t)c * WS_CTA_F + off + e];
  const int k = tid >> 1, n = (tid & 1) * nq + j;
  g_wt[k * n_pad + n] += s;
}
__global__ void k_reduce_vec(const float* __restrict__ ws, int n_cta, int off, int n,
                             float* __restrict__ g) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float s = 0.f;
  for (int c = 0; c < n_cta; ++c) s += ws[(size_t)c * WS_CTA_F + off + e];
  g[e] += s;
}
// basis_mat thread-owned layout e = n * Ra + c -> packed (k_pad = Ra_pad rows c, n_pad cols n)
__global__ void k_reduce_basis(const float* __restrict__ ws, int n_cta, int off, int app_dim, int ra,
                               int n_pad, float* __restrict__ g_wt) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= app_dim * ra) return;
  float s = 0.f;
  for (int c = 0; c < n_cta; ++c) s += ws[(size_t)c * WS_CTA_F + off + e];
  const int nn = e / ra, cc = e - nn * ra;
  g_wt[cc * n_pad + nn] += s;
}

}  // namespace nvfi

using namespace nvfi;

extern "C" int nvfi_launch_advect_bwd_tc(const NvfiField*, const NvfiRenderArgs*, const NvfiRenderBuffers*,
                                         const NvfiRenderGrads*, int, long long, int, int, cudaStream_t);
extern "C" int nvfi_launch_advect_bwd_h(const NvfiField*, const NvfiRenderArgs*, const NvfiRenderBuffers*,
                                        const NvfiRenderGrads*, int, long long, int, cudaStream_t);

static int bwd_num_sms() { return device_sms(); }

extern "C" int64_t nvfi_backward_workspace_bytes(void) {
  return (int64_t)bwd_num_sms() * WS_CTA_F * (int64_t)sizeof(float);
}

extern "C" int nvfi_render_backward(const NvfiField* F, const NvfiRenderArgs* A,
                                    const NvfiRenderBuffers* B, const NvfiRenderGrads* D,
                                    void* stream) {
  if (!F || !A || !B || !D) return NVFI_EINVAL;
  if (A->n_rays <= 0) return NVFI_OK;
  if (!B->sigma || !B->weights || !B->rgb || !B->x_adv || !B->valid || !B->counters ||
      !D->g_x_adv || !D->g_sigma || !D->g_rgb_eff || !D->workspace)
    return NVFI_EINVAL;
  if (D->workspace_bytes < nvfi_backward_workspace_bytes()) return NVFI_EINVAL;
  for (int k = 0; k < 3; ++k)
    if (!D->g_dplane_space[k] || !D->g_dplane_time[k] || !D->g_aplane_space[k] ||
        !D->g_aplane_time[k])
      return NVFI_EINVAL;
  if (F->ra > 64 || F->app_dim * F->ra > 2048 || F->app_dim > 32) return NVFI_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int S = F->n_samples;
  const long long total = (long long)A->n_rays * S;
  const int sms = bwd_num_sms();
  const int n_batches = (int)((total + NVFI_SUBS * NVFI_THREADS - 1) / (NVFI_SUBS * NVFI_THREADS));
  NVFI_CUDA_OK(cudaMemsetAsync(B->counters, 0, 16 * sizeof(int32_t), st));

  // 1. per-ray reverse scan
  {
    const int s_pad = ((S + 31) / 32) * 32;
    const size_t smem = (size_t)8 * 3 * s_pad * sizeof(double);
    if (smem > 200 * 1024) return NVFI_EUNSUPPORTED;
    if (smem > 48 * 1024) {
      const int rc = ensure_smem<k_march_bwd>(smem);
      if (rc != NVFI_OK) return rc;
    }
    NVFI_LAUNCH(k_march_bwd, (unsigned)((A->n_rays + 7) / 8), 256, smem, st, *F, *A, *B, *D, S, s_pad);
    NVFI_CUDA_OK(cudaGetLastError());
  }
  // 2. appearance
  const size_t tile_smem = (size_t)(2 * TILE_F + 2 * NVFI_KC * 128) * sizeof(float);
  if (D->g_rgb) {
    if (F->shading_mode == NVFI_SHADING_MLP_PE) {
      if (!F->render_mlp[0].w_rows || !F->render_mlp[1].w_rows || F->render_mlp[0].k_pad != 128 ||
          !D->g_render_w[0] || !D->g_render_w[1] || !D->g_render_w[2] || !D->g_render_b[0] ||
          !D->g_render_b[1] || !D->g_render_b[2])
        return NVFI_EUNSUPPORTED;
    }
    if (!D->g_basis_mat) return NVFI_EINVAL;
    const size_t smem = tile_smem + sizeof(AppBwdTile);
    {
      const int rc = ensure_smem<k_app_bwd>(smem);
      if (rc != NVFI_OK) return rc;
    }
    const int grid = n_batches < sms ? n_batches : sms;
    NVFI_LAUNCH(k_app_bwd, grid, NVFI_THREADS, smem, st, *F, *A, *B, *D, S, total, n_batches);
    NVFI_CUDA_OK(cudaGetLastError());
    if (F->shading_mode == NVFI_SHADING_MLP_PE) {
      NVFI_LAUNCH(k_reduce_outer, (8 * 2048 + 255) / 256, 256, 0, st, D->workspace, grid, AW_W0, 8, D->g_render_w[0]);
      NVFI_LAUNCH(k_reduce_outer, (8 * 2048 + 255) / 256, 256, 0, st, D->workspace, grid, AW_W1, 8, D->g_render_w[1]);
      NVFI_LAUNCH(k_reduce_small, 4, 256, 0, st, D->workspace, grid, AW_W2, F->render_mlp[2].n_pad, D->g_render_w[2]);
      NVFI_LAUNCH(k_reduce_vec, 1, 128, 0, st, D->workspace, grid, AW_B(0), 128, D->g_render_b[0]);
      NVFI_LAUNCH(k_reduce_vec, 1, 128, 0, st, D->workspace, grid, AW_B(1), 128, D->g_render_b[1]);
      NVFI_LAUNCH(k_reduce_vec, 1, 128, 0, st, D->workspace, grid, AW_B(2), F->render_mlp[2].n_pad, D->g_render_b[2]);
    }
    NVFI_LAUNCH(k_reduce_basis, (F->app_dim * F->ra + 255) / 256, 256, 0, st, D->workspace, grid, AW_BASIS, F->app_dim, F->ra, F->basis_mat.n_pad, D->g_basis_mat);
    NVFI_CUDA_OK(cudaGetLastError());
  }
  // 3. density planes + positions
  NVFI_LAUNCH(k_density_bwd, (unsigned)((A->n_rays + 7) / 8), 256, 0, st, *F, *A, *B, *D, S);
  NVFI_CUDA_OK(cudaGetLastError());
  // 4. velocity net
  if (A->advect) {
    {  // the kernel keeps the uniform RK2 schedule in a fixed-size array
      float off = A->t - A->base_time;
      int steps = 0;
      while (fabsf(off) > 0.f && steps <= MAX_RK2_STEPS) {
        float dt = fminf(fabsf(off), F->dt_max);
        off = off - ((off > 0.f) ? dt : -dt);
        ++steps;
      }
      if (steps > MAX_RK2_STEPS) return NVFI_EUNSUPPORTED;
    }
    for (int l = 0; l < NVFI_VEL_LAYERS; ++l)
      if (!D->g_vel_w[l] || !D->g_vel_b[l] || (l < 5 && !F->vel_net[l].w_rows))
        return NVFI_EINVAL;
    const int mlp_mode = mlp_mode_of(F);
    if (mlp_mode == NVFI_MLP_F16X3)       // product path: FP16-split tcgen05 (backward_h.cu)
      return nvfi_launch_advect_bwd_h(F, A, B, D, S, total, sms, st);
    if (mlp_mode != NVFI_MLP_FP32_SIMT)   // round-1 path: 3xTF32 tcgen05 (backward_tc.cu)
      return nvfi_launch_advect_bwd_tc(F, A, B, D, S, total, sms, mlp_mode, st);
    const size_t smem = tile_smem + sizeof(AdvBwdTile);
    {
      const int rc = ensure_smem<k_advect_bwd>(smem);
      if (rc != NVFI_OK) return rc;
    }
    const int grid = n_batches < sms ? n_batches : sms;
    NVFI_LAUNCH(k_advect_bwd, grid, NVFI_THREADS, smem, st, *F, *A, *B, *D, S, total, n_batches);
    NVFI_CUDA_OK(cudaGetLastError());
    NVFI_LAUNCH(k_reduce_outer, (2 * 2048 + 255) / 256, 256, 0, st, D->workspace, grid, VW_W0, 2, D->g_vel_w[0]);
    for (int l = 1; l <= 4; ++l)
      NVFI_LAUNCH(k_reduce_outer, (8 * 2048 + 255) / 256, 256, 0, st, D->workspace, grid, VW_W(l), 8, D->g_vel_w[l]);
    NVFI_LAUNCH(k_reduce_small, 4, 256, 0, st, D->workspace, grid, VW_W5, 8, D->g_vel_w[5]);
    for (int l = 0; l < 5; ++l)
      NVFI_LAUNCH(k_reduce_vec, 1, 128, 0, st, D->workspace, grid, VW_B(l), 128, D->g_vel_b[l]);
    NVFI_LAUNCH(k_reduce_vec, 1, 128, 0, st, D->workspace, grid, VW_B(5), 8, D->g_vel_b[5]);
    NVFI_CUDA_OK(cudaGetLastError());
  }
  return NVFI_OK;
}
