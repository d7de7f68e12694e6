// PDE loss of the velocity field (NVFi.get_vel_loss, models/nvfi.py:42-84): divergence and
// transport residuals from the Jacobian of VelBasis.forward, hand-written in forward mode
// with a hand-written reverse pass (the reference uses functorch vmap(jacrev) + autograd).
//
//   k_pde_jac      tiles of 25 points x 5 rows (value + 4 tangents d/dx, d/dy, d/dz, d/dt)
//                  through the SiLU weight net; per point: v, J_v, div, transport residual,
//                  loss sums, output adjoints; then the reverse pass of the forward-mode
//                  computation (second order through SiLU) -> weight_net gradients and
//                  dL/da per point
//   k_accnet_bwd   plain backward of the ReLU twin net (a enters the residual by value)
//
// Derivation: SURVEY.md Appendix E "Velocity Jacobian for the PDE loss".  For one layer with
// value h = W a + b and tangents h_k = W a_k, the activation maps (h, h_k) to
// (s(h), s'(h) h_k).  Stacking the five rows of a point turns both the forward pass and the
// reverse pass of the LINEAR parts into ordinary tile GEMMs over 5 x 25 = 125 rows; only the
// activation step couples the rows of a point:
//   reverse:  gh   = ga s'(h) + sum_k ga_k s''(h) h_k,     gh_k = ga_k s'(h)
// with SiLU' = sg (1 + h (1 - sg)), SiLU'' = sg (1 - sg)(2 + h (1 - 2 sg)), sg = sigmoid(h).
#include "backward_common.cuh"

namespace nvfi {

#define PDE_PTS 25   // points per tile (5 rows each)

struct PdeTile {
  float q[4][32];        // x, y, z, t of the tile's points
  float a[3][32];        // acceleration (value only) from the twin net
  float gout[8][NVFI_TM];
  double red[2][8];
  int next;
};

// Rows of the encoded input (models/base_network.py:42-54) and its tangents, BRS-strided:
// row m = 5 p + j; j = 0 value, j = 1..4 derivative w.r.t. coordinate j - 1.
__device__ void pde_encode_tile(const PdeTile& T, int n_pts, float* __restrict__ At) {
  const int tid = threadIdx.x;
  if (tid < NVFI_TM) {
    const int m = tid, p = m / 5, j = m - 5 * p;
    const bool live = p < n_pts && m < 5 * PDE_PTS;
#pragma unroll
    for (int r = 0; r < 32; ++r) At[r * BRS + m] = 0.f;
    if (live) {
      if (j == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float q = T.q[i][p];
          At[i * BRS + m] = q;
          float s, c;
          sincosf(q, &s, &c);
          At[(4 + i) * BRS + m] = s;
          At[(8 + i) * BRS + m] = c;
          sincosf(q * 2.f, &s, &c);
          At[(12 + i) * BRS + m] = s;
          At[(16 + i) * BRS + m] = c;
          sincosf(q * 4.f, &s, &c);
          At[(20 + i) * BRS + m] = s;
          At[(24 + i) * BRS + m] = c;
        }
      } else {
        const int i = j - 1;
        const float q = T.q[i][p];
        At[i * BRS + m] = 1.f;
        float s, c;
        sincosf(q, &s, &c);
        At[(4 + i) * BRS + m] = c;
        At[(8 + i) * BRS + m] = -s;
        sincosf(q * 2.f, &s, &c);
        At[(12 + i) * BRS + m] = 2.f * c;
        At[(16 + i) * BRS + m] = -2.f * s;
        sincosf(q * 4.f, &s, &c);
        At[(20 + i) * BRS + m] = 4.f * c;
        At[(24 + i) * BRS + m] = -4.f * s;
      }
    }
  }
  __syncthreads();
}

// (h + b, h_k) -> stash, then (silu(h), silu'(h) h_k) in place.  Dead points are zeroed.
__device__ void pde_act_fwd(float* __restrict__ At, const float* __restrict__ bias,
                            float* __restrict__ stash, int n_pts) {
  for (int idx = threadIdx.x; idx < NVFI_TM * PDE_PTS; idx += NVFI_THREADS) {
    const int n = idx / PDE_PTS, p = idx - n * PDE_PTS;
    float* a = At + n * BRS + 5 * p;
    float* st = stash + n * NVFI_TM + 5 * p;
    if (p < n_pts) {
      const float h = a[0] + __ldg(bias + n);
      const float sg = sigmoid_f(h);
      const float d1 = sg * (1.f + h * (1.f - sg));
      st[0] = h;
      a[0] = h * sg;
#pragma unroll
      for (int k = 1; k < 5; ++k) {
        const float hk = a[k];
        st[k] = hk;
        a[k] = d1 * hk;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        st[k] = 0.f;
        a[k] = 0.f;
      }
    }
  }
  __syncthreads();
}

// At <- (silu(h), silu'(h) h_k) recomputed from a stash (rows 125..127 zero).
__device__ void pde_load_act(const float* __restrict__ stash, float* __restrict__ At) {
  for (int idx = threadIdx.x; idx < NVFI_TM * PDE_PTS; idx += NVFI_THREADS) {
    const int n = idx / PDE_PTS, p = idx - n * PDE_PTS;
    const float* st = stash + n * NVFI_TM + 5 * p;
    float* a = At + n * BRS + 5 * p;
    const float h = st[0];
    const float sg = sigmoid_f(h);
    const float d1 = sg * (1.f + h * (1.f - sg));
    a[0] = h * sg;
#pragma unroll
    for (int k = 1; k < 5; ++k) a[k] = d1 * st[k];
  }
  if (threadIdx.x < NVFI_TM) {
    At[threadIdx.x * BRS + 125] = 0.f;
    At[threadIdx.x * BRS + 126] = 0.f;
    At[threadIdx.x * BRS + 127] = 0.f;
  }
  __syncthreads();
}

// Reverse of the activation step, in place on G (adjoint of the layer outputs).
__device__ void pde_act_bwd(const float* __restrict__ stash, float* __restrict__ G) {
  for (int idx = threadIdx.x; idx < NVFI_TM * PDE_PTS; idx += NVFI_THREADS) {
    const int n = idx / PDE_PTS, p = idx - n * PDE_PTS;
    const float* st = stash + n * NVFI_TM + 5 * p;
    float* g = G + n * BRS + 5 * p;
    const float h = st[0];
    const float sg = sigmoid_f(h);
    const float d1 = sg * (1.f + h * (1.f - sg));
    const float d2 = sg * (1.f - sg) * (2.f + h * (1.f - 2.f * sg));
    float g0 = g[0] * d1;
#pragma unroll
    for (int k = 1; k < 5; ++k) {
      const float gk = g[k];
      g0 = fmaf(gk * d2, st[k], g0);
      g[k] = gk * d1;
    }
    g[0] = g0;
  }
  __syncthreads();
}

// partial_b[n] += sum over VALUE rows (m = 5 p) of G[n][m].  No trailing barrier.
__device__ inline void pde_bias_acc(const float* __restrict__ G, float* __restrict__ partial_b) {
  const int tid = threadIdx.x;
  if (tid < NVFI_TM) {
    float s = 0.f;
#pragma unroll 5
    for (int p = 0; p < PDE_PTS; ++p) s += G[tid * BRS + 5 * p];
    partial_b[tid] += s;
  }
}

__device__ __forceinline__ void bv(const float u[6], float x, float y, float z, float o[3]) {
  basis_velocity(u, x, y, z, o);
}
// transpose of bv: gu = B gv
__device__ __forceinline__ void bv_t(const float gv[3], float x, float y, float z, float gu[6]) {
  gu[0] = gv[0];
  gu[1] = gv[1];
  gu[2] = gv[2];
  gu[3] = gv[1] * z - gv[2] * y;
  gu[4] = -gv[0] * z + gv[2] * x;
  gu[5] = gv[0] * y - gv[1] * x;
}

__global__ void __launch_bounds__(NVFI_THREADS, 1)
    k_pde_jac(const __grid_constant__ NvfiField F, const float* __restrict__ xyzt,
              const float* __restrict__ va, long long n_total, float c_div, float c_tr,
              int want_grad, float* __restrict__ g_acc_out, double* __restrict__ loss_sums,
              float* __restrict__ workspace, int* counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* At = reinterpret_cast<float*>(smem_raw);
  float* Gt = At + TILE_F;
  float* wS = Gt + TILE_F;
  PdeTile& T = *reinterpret_cast<PdeTile*>(wS + 2 * NVFI_KC * 128);
  float* ws = workspace + (size_t)blockIdx.x * WS_CTA_F;
  float* stash = ws + VW_STASH;
  const int tid = threadIdx.x;
  for (int i = tid; i < VW_PART_F; i += NVFI_THREADS) ws[i] = 0.f;
  __syncthreads();
  const long long n_tiles = (n_total + PDE_PTS - 1) / PDE_PTS;
  double s_div = 0.0, s_tr = 0.0;
  float* outS = wS;

  for (;;) {
    if (tid == 0) T.next = atomicAdd(counter, 1);
    __syncthreads();
    const long long tile = T.next;
    __syncthreads();
    if (tile >= n_tiles) break;
    const long long p0 = tile * PDE_PTS;
    const int n_pts = (int)min((long long)PDE_PTS, n_total - p0);
    if (tid < 32) {
      const bool live = tid < n_pts;
#pragma unroll
      for (int i = 0; i < 4; ++i) T.q[i][tid] = live ? __ldg(xyzt + (p0 + tid) * 4 + i) : 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) T.a[i][tid] = live ? __ldg(va + (p0 + tid) * 6 + 3 + i) : 0.f;
    }
    __syncthreads();
    // ---- forward mode through the 5 hidden layers
    pde_encode_tile(T, n_pts, At);
#pragma unroll 1
    for (int l = 0; l < NVFI_VEL_LAYERS - 1; ++l) {
      NvfiLinear L = F.vel_net[l];
      const float* bias = L.bias;
      L.bias = nullptr;
      tile_linear128<ACT_NONE, BRS>(At, wS, L);
      pde_act_fwd(At, bias, stash + (size_t)l * STASH_F, n_pts);
    }
    {
      NvfiLinear L = F.vel_net[NVFI_VEL_LAYERS - 1];
      L.bias = nullptr;
      tile_linear_small<4, BRS>(At, outS, L);   // outS[i][m], i < 8
    }
    // ---- per point: v, J_v, residuals, output adjoints
    for (int i = tid; i < 8 * NVFI_TM; i += NVFI_THREADS) T.gout[0][i] = 0.f;
    __syncthreads();
    if (tid < n_pts) {
      const int p = tid, m0 = 5 * p;
      const float x = T.q[0][p], y = T.q[1][p], z = T.q[2][p];
      const float* b5 = F.vel_net[NVFI_VEL_LAYERS - 1].bias;
      float w[6], dw[4][6];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        w[i] = outS[i * NVFI_TM + m0] + __ldg(b5 + i);
#pragma unroll
        for (int k = 0; k < 4; ++k) dw[k][i] = outS[i * NVFI_TM + m0 + 1 + k];
      }
      float v[3], J[3][4];
      bv(w, x, y, z, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float c[3];
        bv(dw[k], x, y, z, c);
        J[0][k] = c[0];
        J[1][k] = c[1];
        J[2][k] = c[2];
      }
      // explicit dependence of the basis on the position (models/velocity_field.py:77-98)
      J[0][1] += w[5];
      J[0][2] -= w[4];
      J[1][0] -= w[5];
      J[1][2] += w[3];
      J[2][0] += w[4];
      J[2][1] -= w[3];
      const float div = J[0][0] + J[1][1] + J[2][2];
      float tr[3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
        tr[i] = J[i][0] * v[0] + J[i][1] * v[1] + J[i][2] * v[2] + J[i][3] - T.a[i][p];
      s_div += (double)div * div;
      s_tr += (double)tr[0] * tr[0] + (double)tr[1] * tr[1] + (double)tr[2] * tr[2];
      if (want_grad) {
        const float gdiv = 2.f * c_div * div;
        float gtr[3], gJ[3][4], gv[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 3; ++i) gtr[i] = 2.f * c_tr * tr[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            gJ[i][k] = gtr[i] * v[k] + ((i == k) ? gdiv : 0.f);
            gv[k] += gtr[i] * J[i][k];
          }
          gJ[i][3] = gtr[i];
          g_acc_out[(p0 + p) * 3 + i] = -gtr[i];
        }
        float gw[6];
        bv_t(gv, x, y, z, gw);
        gw[5] += gJ[0][1] - gJ[1][0];
        gw[4] += gJ[2][0] - gJ[0][2];
        gw[3] += gJ[1][2] - gJ[2][1];
#pragma unroll
        for (int i = 0; i < 6; ++i) T.gout[i][m0] = gw[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float gc[3] = {gJ[0][k], gJ[1][k], gJ[2][k]};
          float gd[6];
          bv_t(gc, x, y, z, gd);
#pragma unroll
          for (int i = 0; i < 6; ++i) T.gout[i][m0 + 1 + k] = gd[i];
        }
      }
    }
    __syncthreads();
    if (!want_grad) continue;
    // ---- reverse pass
    small_layer_bwd(At, &T.gout[0][0], F.vel_net[5], Gt, ws + VW_W5, ws + VW_B(5), 5);
#pragma unroll 1
    for (int l = NVFI_VEL_LAYERS - 2; l >= 0; --l) {
      pde_act_bwd(stash + (size_t)l * STASH_F, Gt);
      if (l > 0)
        pde_load_act(stash + (size_t)(l - 1) * STASH_F, At);
      else
        pde_encode_tile(T, n_pts, At);
      pde_bias_acc(Gt, ws + VW_B(l));
      if (l > 0) {
        tile_outer_acc<8>(At, Gt, ws + VW_W(l));
        NvfiLinear Lb = F.vel_net[l];
        Lb.wt = F.vel_net[l].w_rows;
        Lb.bias = nullptr;
        Lb.k_pad = 128;
        tile_linear128<ACT_NONE, BRS>(Gt, wS, Lb);
      } else {
        tile_outer_acc<2>(At, Gt, ws + VW_W0);
      }
    }
  }
  // ---- loss sums: block reduction, one atomic per CTA
  {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s_div += __shfl_xor_sync(0xffffffffu, s_div, o);
      s_tr += __shfl_xor_sync(0xffffffffu, s_tr, o);
    }
    if (lane == 0) {
      T.red[0][warp] = s_div;
      T.red[1][warp] = s_tr;
    }
    __syncthreads();
    if (tid == 0) {
      double a = 0.0, b = 0.0;
      for (int w = 0; w < NVFI_THREADS / 32; ++w) {
        a += T.red[0][w];
        b += T.red[1][w];
      }
      atomicAdd(loss_sums, a);
      atomicAdd(loss_sums + 1, b);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Backward of the ReLU twin net for per-point upstream gradients ga (n,3) of the
// acceleration a = basis_acceleration(net(enc(q)), x)  (models/velocity_field.py:69-75).
// ---------------------------------------------------------------------------------------
struct AccBwdTile {
  float q[4][NVFI_TM];
  float wout[8][NVFI_TM];
  float gout[8][NVFI_TM];
  int next;
};

__global__ void __launch_bounds__(NVFI_THREADS, 1)
    k_accnet_bwd(const __grid_constant__ NvfiField F, const float* __restrict__ xyzt,
                 const float* __restrict__ ga, long long n_total, float* __restrict__ workspace,
                 int* counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* At = reinterpret_cast<float*>(smem_raw);
  float* Gt = At + TILE_F;
  float* wS = Gt + TILE_F;
  AccBwdTile& T = *reinterpret_cast<AccBwdTile*>(wS + 2 * NVFI_KC * 128);
  float* ws = workspace + (size_t)blockIdx.x * WS_CTA_F;
  float* stash = ws + VW_STASH;
  const int tid = threadIdx.x;
  for (int i = tid; i < VW_PART_F; i += NVFI_THREADS) ws[i] = 0.f;
  __syncthreads();
  const long long n_tiles = (n_total + NVFI_TM - 1) / NVFI_TM;
  for (;;) {
    if (tid == 0) T.next = atomicAdd(counter, 1);
    __syncthreads();
    const long long tile = T.next;
    __syncthreads();
    if (tile >= n_tiles) break;
    const long long i0 = tile * NVFI_TM;
    if (tid < NVFI_TM) {
      const bool live = i0 + tid < n_total;
#pragma unroll
      for (int i = 0; i < 4; ++i) T.q[i][tid] = live ? __ldg(xyzt + (i0 + tid) * 4 + i) : 0.f;
    }
    __syncthreads();
    vel_net_tile<ACT_RELU, BRS, true>(F.acc_net, At, wS, &T.wout[0][0], T.q[0], T.q[1], T.q[2],
                                      T.q[3], stash);
    if (tid < NVFI_TM) {
      const int m = tid;
      const bool live = i0 + m < n_total;
      float g[3] = {0.f, 0.f, 0.f};
      if (live) {
        g[0] = __ldg(ga + (i0 + m) * 3);
        g[1] = __ldg(ga + (i0 + m) * 3 + 1);
        g[2] = __ldg(ga + (i0 + m) * 3 + 2);
      }
      const float x = T.q[0][m], y = T.q[1][m], z = T.q[2][m];
      T.gout[0][m] = g[0];
      T.gout[1][m] = g[1];
      T.gout[2][m] = g[2];
      T.gout[3][m] = -y * g[1] - z * g[2];
      T.gout[4][m] = -x * g[0] - z * g[2];
      T.gout[5][m] = -x * g[0] - y * g[1];
      T.gout[6][m] = T.gout[7][m] = 0.f;
    }
    __syncthreads();
    // last layer: input a4 = relu(h4)
    load_stash_act<ACT_RELU>(stash + 4 * STASH_F, At);
    small_layer_bwd(At, &T.gout[0][0], F.acc_net[5], Gt, ws + VW_W5, ws + VW_B(5));
    apply_act_grad<ACT_RELU>(stash + 4 * STASH_F, Gt);
#pragma unroll 1
    for (int l = 4; l >= 1; --l) {
      load_stash_act<ACT_RELU>(stash + (size_t)(l - 1) * STASH_F, At);
      tile_rowsum_acc(Gt, ws + VW_B(l));
      tile_outer_acc<8>(At, Gt, ws + VW_W(l));
      NvfiLinear Lb = F.acc_net[l];
      Lb.wt = F.acc_net[l].w_rows;
      Lb.bias = nullptr;
      Lb.k_pad = 128;
      tile_linear128<ACT_NONE, BRS>(Gt, wS, Lb);
      apply_act_grad<ACT_RELU>(stash + (size_t)(l - 1) * STASH_F, Gt);
    }
    vel_encode_tile<BRS>(At, T.q[0], T.q[1], T.q[2], T.q[3]);
    tile_rowsum_acc(Gt, ws + VW_B(0));
    tile_outer_acc<2>(At, Gt, ws + VW_W0);
  }
}

}  // namespace nvfi

using namespace nvfi;

static int pde_sms() { return device_sms(); }

extern "C" int nvfi_launch_pde_jac_h(const NvfiField*, const float*, const float*, long long, float, float, int,
                                     const NvfiPdeGrads*, double*, int*, int, cudaStream_t);
extern "C" int nvfi_launch_accnet_bwd_h(const NvfiField*, const float*, const float*, long long, const NvfiPdeGrads*,
                                        int*, int, cudaStream_t);

static void reduce_net(const float* ws, int grid, float* const gw[NVFI_VEL_LAYERS],
                       float* const gb[NVFI_VEL_LAYERS], cudaStream_t st) {
  NVFI_LAUNCH(k_reduce_outer, (2 * 2048 + 255) / 256, 256, 0, st, ws, grid, VW_W0, 2, gw[0]);
  for (int l = 1; l <= 4; ++l)
    NVFI_LAUNCH(k_reduce_outer, (8 * 2048 + 255) / 256, 256, 0, st, ws, grid, VW_W(l), 8, gw[l]);
  NVFI_LAUNCH(k_reduce_small, 4, 256, 0, st, ws, grid, VW_W5, 8, gw[5]);
  for (int l = 0; l < 5; ++l) NVFI_LAUNCH(k_reduce_vec, 1, 128, 0, st, ws, grid, VW_B(l), 128, gb[l]);
  NVFI_LAUNCH(k_reduce_vec, 1, 128, 0, st, ws, grid, VW_B(5), 8, gb[5]);
}

extern "C" int nvfi_pde_loss(const NvfiField* F, const float* xyzt, const float* va, int64_t n,
                             double* loss_sums, const NvfiPdeGrads* G, int32_t want_grad,
                             int32_t* counters, void* stream) {
  if (!F || !loss_sums || !counters || !G || n < 0) return NVFI_EINVAL;
  if (!F->use_vel) return NVFI_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  NVFI_CUDA_OK(cudaMemsetAsync(loss_sums, 0, 2 * sizeof(double), st));
  if (n == 0) return NVFI_OK;
  if (!xyzt || !va) return NVFI_EINVAL;
  for (int l = 0; l < NVFI_VEL_LAYERS; ++l) {
    const NvfiLinear& L = F->vel_net[l];
    if (!L.wt || !L.bias) return NVFI_EINVAL;
    if (l == 0 && (L.in_dim != NVFI_VEL_IN || L.k_pad != 32)) return NVFI_EUNSUPPORTED;
    if (l > 0 && L.k_pad != 128) return NVFI_EUNSUPPORTED;
    if (l < NVFI_VEL_LAYERS - 1 && L.n_pad != 128) return NVFI_EUNSUPPORTED;
    if (l == NVFI_VEL_LAYERS - 1 && L.n_pad != 8) return NVFI_EUNSUPPORTED;
    if (want_grad && l >= 1 && l <= 4 && (!L.w_rows || !F->acc_net[l].w_rows)) return NVFI_EINVAL;
  }
  if (!G->workspace || G->workspace_bytes < nvfi_backward_workspace_bytes()) return NVFI_EINVAL;
  if (want_grad) {
    if (!G->g_acc_pts) return NVFI_EINVAL;
    for (int l = 0; l < NVFI_VEL_LAYERS; ++l)
      if (!G->g_vel_w[l] || !G->g_vel_b[l] || !G->g_acc_w[l] || !G->g_acc_b[l]) return NVFI_EINVAL;
  }
  NVFI_CUDA_OK(cudaMemsetAsync(counters, 0, 4 * sizeof(int32_t), st));
  const int sms = pde_sms();
  // NVFi.get_vel_loss: 5 * mean(div^2) + 0.1 * mean(transport^2) over (n) and (n, 3) entries
  const float c_div = (float)(5.0 / (double)n), c_tr = (float)(0.1 / (3.0 * (double)n));
  const size_t tile_smem = (size_t)(2 * TILE_F + 2 * NVFI_KC * 128) * sizeof(float);
  const bool h16 = grad_layout_v4(F, F->vel_net);
  if (h16) {   // product path: forward-mode rows on the tensor cores (backward_h.cu: k_pde_jac_h)
    const int rc = nvfi_launch_pde_jac_h(F, xyzt, va, (long long)n, c_div, c_tr, (int)want_grad, G, loss_sums,
                                         counters, sms, st);
    if (rc != NVFI_OK) return rc;
  } else {
    const size_t smem = tile_smem + sizeof(PdeTile);
    {
      const int rc = ensure_smem<k_pde_jac>(smem);
      if (rc != NVFI_OK) return rc;
    }
    const long long n_tiles = (n + PDE_PTS - 1) / PDE_PTS;
    const int grid = (int)(n_tiles < sms ? n_tiles : sms);
    NVFI_LAUNCH(k_pde_jac, grid, NVFI_THREADS, smem, st, *F, xyzt, va, (long long)n, c_div, c_tr,
                (int)want_grad, G->g_acc_pts, loss_sums, G->workspace, counters);
    NVFI_CUDA_OK(cudaGetLastError());
    if (want_grad) reduce_net(G->workspace, grid, G->g_vel_w, G->g_vel_b, st);
  }
  const bool h16a = want_grad && grad_layout_v4(F, F->acc_net);
  if (h16a) {
    const int rc = nvfi_launch_accnet_bwd_h(F, xyzt, G->g_acc_pts, (long long)n, G, counters + 1, sms, st);
    if (rc != NVFI_OK) return rc;
  } else if (want_grad) {
    const size_t smem = tile_smem + sizeof(AccBwdTile);
    {
      const int rc = ensure_smem<k_accnet_bwd>(smem);
      if (rc != NVFI_OK) return rc;
    }
    const long long n_tiles = (n + NVFI_TM - 1) / NVFI_TM;
    const int grid = (int)(n_tiles < sms ? n_tiles : sms);
    NVFI_LAUNCH(k_accnet_bwd, grid, NVFI_THREADS, smem, st, *F, xyzt, G->g_acc_pts, (long long)n,
                G->workspace, counters + 1);
    NVFI_CUDA_OK(cudaGetLastError());
    reduce_net(G->workspace, grid, G->g_acc_w, G->g_acc_b, st);
  }
  return NVFI_OK;
}
