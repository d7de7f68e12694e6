// Kernel 3 of the render path: appearance of the samples whose weight passes
// rayMarch_weight_thres.
//
// Replaces:
//   app_mask = weight > thres                     models/tensorf_keyframe.py:719
//   compute_appfeature (6 x grid_sample + basis)  models/tensorf_keyframe.py:274-310
//   MLPRender_PE.forward / positional_encoding    models/tensorf_base.py:88-98,
//                                                 models/tensorf_model_utils.py:176-183
//   SHRender / eval_sh_bases(deg 2)               models/tensorf_model_utils.py:292-296,
//                                                 models/sh.py:87-116
//   MaskField.forward + weighted mask sum         models/mask_field.py:68-83,
//                                                 models/tensorf_keyframe.py:749-753
//
// Persistent CTAs scan the (n_rays, S) weights, compact the passing slots into a
// shared-memory queue and decode them 128 at a time: 8-lane-group gather of the
// 48-component planes into a k-major activation tile, then the FP32 tile-GEMM MLP.
// rgb goes to a dense (n_rays, S, 3) buffer (only passing slots are written/read), so
// the colour composite (k_composite) is deterministic.
#include "nvfi_common.cuh"

namespace nvfi {

struct AppTile {
  float x[3][NVFI_TM];
  float d[3][NVFI_TM];
  int gidx[NVFI_TM];
  int q_idx[NVFI_QCAP];
  int warp_cnt[2][NVFI_THREADS / 32];
  int batch;
};

// models/sh.py:87-116, deg 2
__device__ __forceinline__ void sh_bases_deg2(float x, float y, float z, float sh[9]) {
  const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
  const float C20 = 1.0925484305920792f, C21 = -1.0925484305920792f, C22 = 0.31539156525252005f,
              C23 = -1.0925484305920792f, C24 = 0.5462742152960396f;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  sh[0] = C0;
  sh[1] = -C1 * y;
  sh[2] = C1 * z;
  sh[3] = -C1 * x;
  sh[4] = C20 * xy;
  sh[5] = C21 * yz;
  sh[6] = C22 * (2.0f * zz - xx - yy);
  sh[7] = C23 * xz;
  sh[8] = C24 * (xx - yy);
}

// Gather the Ra-component appearance features of the tile's n points into actT rows
// [0, Ra): actT[c][m].  Rows of dead columns (m >= n) are zeroed.
__device__ void app_gather_tile(const NvfiField& F, float tnb, const AppTile& T, int n,
                                float* __restrict__ actT) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 3, l8 = lane & 7;
  const int R4 = F.ra >> 2;
#pragma unroll 1
  for (int pass = 0; pass < NVFI_TM / 32; ++pass) {
    const int m = pass * 32 + warp * 4 + g;
    float4 pr[2];
    pr[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    pr[1] = pr[0];
    if (m < n) {
      const float xt[4] = {T.x[0][m], T.x[1][m], T.x[2][m], tnb};
      kplanes_features<2>(F, F.aplane_space, F.aplane_time, F.ra, xt, l8, pr);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int q = l8 + 8 * j;
      if (q < R4) {
        actT[(4 * q + 0) * NVFI_TM + m] = pr[j].x;
        actT[(4 * q + 1) * NVFI_TM + m] = pr[j].y;
        actT[(4 * q + 2) * NVFI_TM + m] = pr[j].z;
        actT[(4 * q + 3) * NVFI_TM + m] = pr[j].w;
      }
    }
  }
  __syncthreads();
}

// Rows [app_dim, k_pad) of the MLP_PE input (models/tensorf_base.py:88-96):
// [viewdirs(3), pts(3), PE(pts, pos_pe), PE(viewdirs, view_pe)], zero padding after.
__device__ void mlp_pe_inputs_tile(const NvfiField& F, const AppTile& T, float* __restrict__ actT) {
  const int tid = threadIdx.x;
  const int m = tid & 127, part = tid >> 7;
  const int b0 = F.app_dim;
  const int pp = F.pos_pe, vp = F.view_pe;
  const int row_pe_p = b0 + 6;
  const int row_pe_v = row_pe_p + 6 * pp;
  const int row_end = row_pe_v + 6 * vp;
  if (part == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float p = T.x[a][m];
      actT[(b0 + 3 + a) * NVFI_TM + m] = p;
      float f = 1.f;
      for (int k = 0; k < pp; ++k) {
        float s, c;
        sincosf(p * f, &s, &c);
        actT[(row_pe_p + a * pp + k) * NVFI_TM + m] = s;
        actT[(row_pe_p + 3 * pp + a * pp + k) * NVFI_TM + m] = c;
        f *= 2.f;
      }
    }
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float p = T.d[a][m];
      actT[(b0 + a) * NVFI_TM + m] = p;
      float f = 1.f;
      for (int k = 0; k < vp; ++k) {
        float s, c;
        sincosf(p * f, &s, &c);
        actT[(row_pe_v + a * vp + k) * NVFI_TM + m] = s;
        actT[(row_pe_v + 3 * vp + a * vp + k) * NVFI_TM + m] = c;
        f *= 2.f;
      }
    }
    for (int r = row_end; r < F.render_mlp[0].k_pad; ++r) actT[r * NVFI_TM + m] = 0.f;
  }
  __syncthreads();
}

// Decode one tile of n appearance samples.  outS aliases the W staging buffer.
__device__ void appearance_tile(const NvfiField& F, const NvfiRenderArgs& A,
                                const NvfiRenderBuffers& B, const AppTile& T, int n, int S,
                                float* __restrict__ actT, float* __restrict__ wS) {
  const int tid = threadIdx.x;
  float* outS = wS;
  app_gather_tile(F, A.t_norm_base, T, n, actT);
  tile_linear_small<16>(actT, outS, F.basis_mat);
  if (F.shading_mode == NVFI_SHADING_SH) {
    if (tid < n) {
      const int m = tid;
      float sh[9];
      sh_bases_deg2(T.d[0][m], T.d[1][m], T.d[2][m], sh);
      const long long gi = T.gidx[m];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float acc = 0.f;
#pragma unroll
        for (int b = 0; b < 9; ++b) acc = fmaf(sh[b], outS[(ch * 9 + b) * NVFI_TM + m], acc);
        B.rgb[gi * 3 + ch] = fmaxf(acc + 0.5f, 0.f);
      }
    }
    __syncthreads();
  } else {
    // features -> rows [0, app_dim)
    for (int i = tid; i < F.app_dim * NVFI_TM; i += NVFI_THREADS) actT[i] = outS[i];
    mlp_pe_inputs_tile(F, T, actT);  // ends with a barrier
    tile_linear128<ACT_RELU>(actT, wS, F.render_mlp[0]);
    tile_linear128<ACT_RELU>(actT, wS, F.render_mlp[1]);
    tile_linear_small<2>(actT, outS, F.render_mlp[2]);
    if (tid < n) {
      const long long gi = T.gidx[tid];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) B.rgb[gi * 3 + ch] = sigmoid_f(outS[ch * NVFI_TM + tid]);
    }
    __syncthreads();
  }
  if (F.mask_layers > 0) {
    // MaskField on the advected normalised position (models/tensorf_keyframe.py:749-753)
    for (int i = tid; i < 32 * NVFI_TM; i += NVFI_THREADS) {
      const int r = i >> 7, m = i & 127;
      actT[i] = (r < 3) ? T.x[r][m] : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int l = 0; l < F.mask_layers - 1; ++l) tile_linear128<ACT_RELU>(actT, wS, F.mask_net[l]);
    tile_linear_small<16>(actT, outS, F.mask_net[F.mask_layers - 1]);
    if (tid < n) {
      const int m = tid, md = F.mask_dim;
      float mx = -INFINITY;
      for (int j = 0; j < md; ++j) mx = fmaxf(mx, outS[j * NVFI_TM + m]);
      float den = 0.f;
      for (int j = 0; j < md; ++j) den += expf(outS[j * NVFI_TM + m] - mx);
      const long long gi = T.gidx[m];
      const long long ray = gi / S;
      const float w = B.weights[gi];
      for (int j = 0; j < md; ++j)
        atomicAdd(B.mask_map + ray * md + j, w * (expf(outS[j * NVFI_TM + m] - mx) / den));
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(NVFI_THREADS, 2)
    k_appearance(const NvfiField F, const NvfiRenderArgs A, const NvfiRenderBuffers B, int S,
                 long long total, int n_batches, int act_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* actT = reinterpret_cast<float*>(smem_raw);
  float* wS = actT + (size_t)act_rows * NVFI_TM;
  AppTile& T = *reinterpret_cast<AppTile*>(wS + 2 * NVFI_KC * 128);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  int sub = NVFI_SUBS;
  long long batch_base = 0;
  bool exhausted = false;
  int qc = 0, par = 0;

  for (;;) {
    while (qc < NVFI_TM && !exhausted) {
      if (sub == NVFI_SUBS) {
        if (tid == 0) T.batch = atomicAdd(&B.counters[1], 1);
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
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        T.x[a][tid] = live ? B.x_adv[gi * 3 + a] : 0.f;
        T.d[a][tid] = live ? __ldg(A.rays_d + ray * 3 + a) : 0.f;
      }
    }
    __syncthreads();
    appearance_tile(F, A, B, T, n, S, actT, wS);
  }
}

// Stand-alone compute_appfeature on arbitrary points (n, 4) -> (n, app_dim).
__global__ void __launch_bounds__(NVFI_THREADS, 2)
    k_app_feature_points(const NvfiField F, const float* __restrict__ xyzt, long long n,
                         float* __restrict__ out, int* counter, int act_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* actT = reinterpret_cast<float*>(smem_raw);
  float* wS = actT + (size_t)act_rows * NVFI_TM;
  AppTile& T = *reinterpret_cast<AppTile*>(wS + 2 * NVFI_KC * 128);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, g = lane >> 3, l8 = lane & 7;
  const long long n_tiles = (n + NVFI_TM - 1) / NVFI_TM;
  const int R4 = F.ra >> 2;
  for (;;) {
    if (tid == 0) T.batch = atomicAdd(counter, 1);
    __syncthreads();
    const long long tile = T.batch;
    __syncthreads();
    if (tile >= n_tiles) break;
    const long long i0 = tile * NVFI_TM;
    // per-point time coordinate: gather directly (cannot use the shared t of a render)
#pragma unroll 1
    for (int pass = 0; pass < NVFI_TM / 32; ++pass) {
      const int m = pass * 32 + warp * 4 + g;
      const long long i = i0 + m;
      float4 pr[2];
      pr[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      pr[1] = pr[0];
      if (i < n) {
        const float xt[4] = {__ldg(xyzt + i * 4), __ldg(xyzt + i * 4 + 1), __ldg(xyzt + i * 4 + 2),
                             __ldg(xyzt + i * 4 + 3)};
        kplanes_features<2>(F, F.aplane_space, F.aplane_time, F.ra, xt, l8, pr);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int q = l8 + 8 * j;
        if (q < R4) {
          actT[(4 * q + 0) * NVFI_TM + m] = pr[j].x;
          actT[(4 * q + 1) * NVFI_TM + m] = pr[j].y;
          actT[(4 * q + 2) * NVFI_TM + m] = pr[j].z;
          actT[(4 * q + 3) * NVFI_TM + m] = pr[j].w;
        }
      }
    }
    __syncthreads();
    tile_linear_small<16>(actT, wS, F.basis_mat);
    for (int idx = tid; idx < F.app_dim * NVFI_TM; idx += NVFI_THREADS) {
      const int m = idx / F.app_dim, c = idx - m * F.app_dim;
      if (i0 + m < n) out[(i0 + m) * F.app_dim + c] = wS[c * NVFI_TM + m];
    }
    __syncthreads();
  }
}

}  // namespace nvfi

using namespace nvfi;

static int app_act_rows(const NvfiField* F) {
  int rows = NVFI_TM;
  if (F->shading_mode == NVFI_SHADING_MLP_PE && F->render_mlp[0].k_pad > rows)
    rows = F->render_mlp[0].k_pad;
  if (F->ra > rows) rows = F->ra;
  return rows;
}

static int check_app_config(const NvfiField* F) {
  if (F->ra % 4 != 0 || F->ra > 64 || F->ra <= 0) return NVFI_EUNSUPPORTED;
  if (F->basis_mat.n_pad > 32 || F->basis_mat.in_dim != F->ra) return NVFI_EUNSUPPORTED;
  if (F->shading_mode == NVFI_SHADING_MLP_PE) {
    const int in_dim = F->app_dim + 6 + 6 * F->pos_pe + 6 * F->view_pe;
    if (F->render_mlp[0].in_dim != in_dim || F->render_mlp[0].k_pad > NVFI_ACT_ROWS)
      return NVFI_EUNSUPPORTED;
    if (F->render_mlp[0].n_pad != 128 || F->render_mlp[1].k_pad != 128 ||
        F->render_mlp[1].n_pad != 128 || F->render_mlp[2].n_pad != 4)
      return NVFI_EUNSUPPORTED;
  } else if (F->shading_mode == NVFI_SHADING_SH) {
    if (F->app_dim != 27) return NVFI_EUNSUPPORTED;
  } else {
    return NVFI_EUNSUPPORTED;
  }
  if (F->mask_layers > 0) {
    if (F->mask_layers > NVFI_MAX_MASK_LAYERS || F->mask_dim > NVFI_MAX_MASK_DIM ||
        F->mask_layers < 2)
      return NVFI_EUNSUPPORTED;
    for (int l = 0; l < F->mask_layers - 1; ++l)
      if (F->mask_net[l].n_pad != 128 || F->mask_net[l].k_pad > 128) return NVFI_EUNSUPPORTED;
    if (F->mask_net[0].in_dim != 3 || F->mask_net[F->mask_layers - 1].n_pad > 32)
      return NVFI_EUNSUPPORTED;
  }
  return NVFI_OK;
}

extern "C" int nvfi_launch_appearance(const NvfiField* F, const NvfiRenderArgs* A,
                                      const NvfiRenderBuffers* B, cudaStream_t st) {
  const int S = F->n_samples;
  const long long total = (long long)A->n_rays * S;
  if (total <= 0) return NVFI_OK;
  int rc = check_app_config(F);
  if (rc != NVFI_OK) return rc;
  const int rows = app_act_rows(F);
  const size_t smem = (size_t)rows * NVFI_TM * 4 + 2 * NVFI_KC * 128 * 4 + sizeof(AppTile);
  rc = ensure_smem<k_appearance>(smem);
  if (rc != NVFI_OK) return rc;
  const int sms = device_sms();
  const int n_batches = (int)((total + NVFI_SUBS * NVFI_THREADS - 1) / (NVFI_SUBS * NVFI_THREADS));
  const int grid = n_batches < sms * 2 ? n_batches : sms * 2;
  NVFI_LAUNCH(k_appearance, grid, NVFI_THREADS, smem, st, *F, *A, *B, S, total, n_batches, rows);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_app_feature(const NvfiField* F, const float* xyzt, int64_t n, float* feat,
                                int32_t* counters, void* stream) {
  if (!F || !xyzt || !feat || !counters || n < 0) return NVFI_EINVAL;
  if (n == 0) return NVFI_OK;
  int rc = check_app_config(F);
  if (rc != NVFI_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = app_act_rows(F);
  const size_t smem = (size_t)rows * NVFI_TM * 4 + 2 * NVFI_KC * 128 * 4 + sizeof(AppTile);
  rc = ensure_smem<k_app_feature_points>(smem);
  if (rc != NVFI_OK) return rc;
  NVFI_CUDA_OK(cudaMemsetAsync(counters, 0, sizeof(int32_t), st));
  const int sms = device_sms();
  const long long n_tiles = (n + NVFI_TM - 1) / NVFI_TM;
  const int grid = (int)(n_tiles < (long long)sms * 2 ? n_tiles : (long long)sms * 2);
  NVFI_LAUNCH(k_app_feature_points, grid, NVFI_THREADS, smem, st, *F, xyzt, n, feat, counters, rows);
  return (int)cudaGetLastError();
}
