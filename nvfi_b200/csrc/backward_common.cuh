// Tile primitives of the backward pass (FP32 SIMT).  All tiles are feature-major in
// shared memory with a row stride of BRS = 132 floats: tile[f][m], f = feature / unit,
// m = sample within the 128-sample tile.  The odd-multiple-of-4 stride makes both access
// patterns conflict free: rows read along m (forward / input-gradient GEMMs) and 16
// different rows read at the same m (weight-gradient GEMM).
//
// Activations are not kept between forward and backward (SURVEY.md Appendix E:
// "recompute, do not store"): the backward kernels re-run the forward layers, stash the
// pre-activations of the current tile in a per-CTA global scratch (L2 resident) and walk
// the layers in reverse.  Weight gradients accumulate in thread-owned per-CTA partial
// buffers (plain read-modify-write, no atomics) that a small reduce kernel sums.
#pragma once

#include "nvfi_common.cuh"

#define BRS 132
#define TILE_F (NVFI_TM * BRS)
#define STASH_F (NVFI_TM * NVFI_TM)

namespace nvfi {

// ---------------------------------------------------------------------------------------
// per-CTA workspace layouts (float offsets)
// ---------------------------------------------------------------------------------------
// velocity (k_advect_bwd)
#define VW_W0 0                                 // 32 x 128 (tile_outer layout, KI = 2)
#define VW_W(l) (4096 + ((l) - 1) * 16384)      // l = 1..4, 128 x 128 (KI = 8)
#define VW_W5 (4096 + 4 * 16384)                // small layout, 1024
#define VW_B(l) (VW_W5 + 1024 + (l) * 128)      // l = 0..5 (last uses 8)
#define VW_PART_F (VW_W5 + 1024 + 6 * 128)      // floats of partials
#define VW_STASH (VW_PART_F)                    // [2 evals][5 layers][STASH_F]
#define VW_XSTEPS (VW_STASH + 10 * STASH_F)     // [32 steps][3][128]
#define VW_TOTAL (VW_XSTEPS + 32 * 3 * NVFI_TM)
// appearance (k_app_bwd)
#define AW_W0 0
#define AW_W1 16384
#define AW_W2 32768                             // small layout, 1024
#define AW_B(l) (AW_W2 + 1024 + (l) * 128)      // l = 0..2
#define AW_BASIS (AW_W2 + 1024 + 3 * 128)       // thread-owned, 256 * 8
#define AW_PART_F (AW_BASIS + 2048)
#define AW_STASH_X (AW_PART_F)
#define AW_STASH_H0 (AW_STASH_X + STASH_F)
#define AW_STASH_H1 (AW_STASH_H0 + STASH_F)
#define AW_STASH_F48 (AW_STASH_H1 + STASH_F)    // 64 x 128
#define AW_TOTAL (AW_STASH_F48 + 64 * NVFI_TM)
// the tensor-core backward (backward_tc.cu) additionally ping-pongs G between two 128x128 scratch tiles
#define TCW_TOTAL (VW_TOTAL + 2 * STASH_F)
#define WS_CTA_F (TCW_TOTAL > AW_TOTAL ? TCW_TOTAL : AW_TOTAL)
#define MAX_RK2_STEPS 32


// reductions of the per-CTA partials into the packed gradients (backward.cu)
__global__ void k_reduce_outer(const float* __restrict__ ws, int n_cta, int off, int ki,
                               float* __restrict__ g_wt);
__global__ void k_reduce_small(const float* __restrict__ ws, int n_cta, int off, int n_pad,
                               float* __restrict__ g_wt);
__global__ void k_reduce_vec(const float* __restrict__ ws, int n_cta, int off, int n,
                             float* __restrict__ g);

template <int ACT>
__device__ __forceinline__ float act_grad(float h) {
  if (ACT == ACT_RELU) return h > 0.f ? 1.f : 0.f;
  if (ACT == ACT_SILU) {
    const float s = sigmoid_f(h);
    return s * (1.f + h * (1.f - s));
  }
  return 1.f;
}

__device__ __forceinline__ void red_add_v4(float* addr, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// tile[f][m] = act(stash[f][m]) for f < rows
template <int ACT>
__device__ void load_stash_act(const float* __restrict__ stash, float* __restrict__ tile,
                               int rows = NVFI_TM) {
  for (int i = threadIdx.x; i < rows * 32; i += NVFI_THREADS) {
    const int f = i >> 5, m4 = i & 31;
    float4 h = *reinterpret_cast<const float4*>(stash + f * NVFI_TM + m4 * 4);
    h.x = activate<ACT>(h.x);
    h.y = activate<ACT>(h.y);
    h.z = activate<ACT>(h.z);
    h.w = activate<ACT>(h.w);
    *reinterpret_cast<float4*>(tile + f * BRS + m4 * 4) = h;
  }
  __syncthreads();
}

// tile[f][m] -> stash[f][m]
__device__ inline void store_stash(const float* __restrict__ tile, float* __restrict__ stash,
                                   int rows = NVFI_TM) {
  for (int i = threadIdx.x; i < rows * 32; i += NVFI_THREADS) {
    const int f = i >> 5, m4 = i & 31;
    *reinterpret_cast<float4*>(stash + f * NVFI_TM + m4 * 4) =
        *reinterpret_cast<const float4*>(tile + f * BRS + m4 * 4);
  }
  __syncthreads();
}

// G[f][m] *= act'(stash[f][m])
template <int ACT>
__device__ void apply_act_grad(const float* __restrict__ stash, float* __restrict__ G) {
  for (int i = threadIdx.x; i < NVFI_TM * 32; i += NVFI_THREADS) {
    const int f = i >> 5, m4 = i & 31;
    const float4 h = *reinterpret_cast<const float4*>(stash + f * NVFI_TM + m4 * 4);
    float4 g = *reinterpret_cast<float4*>(G + f * BRS + m4 * 4);
    g.x *= act_grad<ACT>(h.x);
    g.y *= act_grad<ACT>(h.y);
    g.z *= act_grad<ACT>(h.z);
    g.w *= act_grad<ACT>(h.w);
    *reinterpret_cast<float4*>(G + f * BRS + m4 * 4) = g;
  }
  __syncthreads();
}

// Weight gradient of one layer on one tile: dW^T[k][n] += sum_m A[k][m] G[n][m] for
// k < 16*KI, n < 128.  Thread (tk, tn) = (tid & 15, tid >> 4) owns k = tk + 16 i,
// n = tn + 16 j and keeps its 8*KI sums in the per-CTA partial buffer at
// float4 index q * 256 + tid, q = 2 i + (j >> 2), component j & 3.
template <int KI>
__device__ void tile_outer_acc(const float* __restrict__ A, const float* __restrict__ G,
                               float* __restrict__ partial) {
  const int tid = threadIdx.x;
  const int tk = tid & 15, tn = tid >> 4;
  float acc[KI][8];
#pragma unroll
  for (int i = 0; i < KI; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 2
  for (int m = 0; m < NVFI_TM; m += 4) {
    float4 g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = *reinterpret_cast<const float4*>(G + (tn + 16 * j) * BRS + m);
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const float4 a = *reinterpret_cast<const float4*>(A + (tk + 16 * i) * BRS + m);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] += f4_dot(a, g[j]);
    }
  }
  float4* p4 = reinterpret_cast<float4*>(partial);
#pragma unroll
  for (int q = 0; q < 2 * KI; ++q) {
    float4 v = p4[q * NVFI_THREADS + tid];
    v.x += acc[q >> 1][(q & 1) * 4 + 0];
    v.y += acc[q >> 1][(q & 1) * 4 + 1];
    v.z += acc[q >> 1][(q & 1) * 4 + 2];
    v.w += acc[q >> 1][(q & 1) * 4 + 3];
    p4[q * NVFI_THREADS + tid] = v;
  }
  __syncthreads();
}

// Bias gradient: partial_b[n] += sum_m G[n][m].  No trailing barrier (read only).
__device__ inline void tile_rowsum_acc(const float* __restrict__ G, float* __restrict__ partial_b) {
  const int tid = threadIdx.x;
  const int n = tid >> 1, half = tid & 1;
  float s = 0.f;
#pragma unroll 4
  for (int m4 = 0; m4 < 16; ++m4)
    s += f4_sum(*reinterpret_cast<const float4*>(G + n * BRS + half * 64 + m4 * 4));
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (!half) partial_b[n] += s;
}

// Backward of a narrow output layer (n_pad <= 8): gout[n][m] (stride 128) is the gradient
// of the layer output.  A holds the layer input a[k][m].
//   partial_w[tid * 4 + j] += dW^T[k][n],  k = tid >> 1, n = (tid & 1) * (n_pad / 2) + j
//   partial_b[n]           += sum_m gout[n][m]
//   G[k][m]                 = sum_n W^T[k][n] gout[n][m]          (input gradient)
__device__ inline void small_layer_bwd(const float* __restrict__ A, const float* __restrict__ gout,
                                       const NvfiLinear& L, float* __restrict__ G,
                                       float* __restrict__ partial_w,
                                       float* __restrict__ partial_b, int bias_stride = 1) {
  const int tid = threadIdx.x;
  const int nq = L.n_pad >> 1;  // <= 4
  {
    const int k = tid >> 1, n0 = (tid & 1) * nq;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int m = 0; m < NVFI_TM; ++m) {
      const float a = A[k * BRS + m];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nq) acc[j] = fmaf(a, gout[(n0 + j) * NVFI_TM + m], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nq) partial_w[tid * 4 + j] += acc[j];
  }
  if (tid < L.n_pad) {
    float s = 0.f;
    for (int m = 0; m < NVFI_TM; m += bias_stride) s += gout[tid * NVFI_TM + m];
    partial_b[tid] += s;
  }
  {
    const int m = tid & 127, k0 = (tid >> 7) * 64;
    float go[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) go[n] = (n < L.n_pad) ? gout[n * NVFI_TM + m] : 0.f;
    for (int k = k0; k < k0 + 64; ++k) {
      const float* wr = L.wt + (size_t)k * L.n_pad;
      float s = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n)
        if (n < L.n_pad) s = fmaf(__ldg(wr + n), go[n], s);
      G[k * BRS + m] = s;
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------
// k-planes gather backward for one point by an 8-lane group.
//   gch[j]   : dL/dF_c for the lane's channels (slot j)
//   gsp/gtp  : packed plane-gradient accumulators (red.add)
//   gx[3]    : this lane's share of dL/dx (caller sums over the group)
// Recomputes the six bilinear factors, forms for every plane the product of the other
// five (prefix / suffix products), scatters w_corner * g into the four corners and
// accumulates the coordinate gradient  d b / d gx = (W-1)/2 * sum_c dwx_c v_c
// (SURVEY.md Appendix E "Gather").
// ---------------------------------------------------------------------------------------
template <int NSLOT>
__device__ __forceinline__ void kplanes_backward(const NvfiField& F, const float* const sp[3],
                                                 const float* const tp[3], float* const gsp[3],
                                                 float* const gtp[3], int R, const float xt[4],
                                                 int l8, const float4 gch[NSLOT], float gx[3]) {
  const int R4 = R >> 2;
  float4 b[6][NSLOT];
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    int cx, cy, H, W;
    plane_geom(F, p, cx, cy, H, W);
    Bilerp bl;
    bilerp_setup(xt[cx], xt[cy], H, W, R, bl);
    const float* base = (p < 3) ? sp[p] : tp[p - 3];
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) {
      const int q = l8 + 8 * j;
      b[p][j] = make_float4(1.f, 1.f, 1.f, 1.f);
      if (q < R4) {
        const float4 v0 = ldg4(base + bl.off[0] + 4 * q);
        const float4 v1 = ldg4(base + bl.off[1] + 4 * q);
        const float4 v2 = ldg4(base + bl.off[2] + 4 * q);
        const float4 v3 = ldg4(base + bl.off[3] + 4 * q);
        b[p][j] = f4_blend(v0, v1, v2, v3, bl.w);
      }
    }
  }
  // go[p] = g * prod_{n != p} b[n]
  float4 go[6][NSLOT];
#pragma unroll
  for (int j = 0; j < NSLOT; ++j) {
    float4 pre = gch[j];
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      go[p][j] = pre;
      pre = f4_mul(pre, b[p][j]);
    }
    float4 suf = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
    for (int p = 5; p >= 0; --p) {
      go[p][j] = f4_mul(go[p][j], suf);
      suf = f4_mul(suf, b[p][j]);
    }
  }
  gx[0] = gx[1] = gx[2] = 0.f;
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    int cx, cy, H, W;
    plane_geom(F, p, cx, cy, H, W);
    Bilerp bl;
    bilerp_setup(xt[cx], xt[cy], H, W, R, bl);
    const float* base = (p < 3) ? sp[p] : tp[p - 3];
    float* gbase = (p < 3) ? gsp[p] : gtp[p - 3];
    float dgx = 0.f, dgy = 0.f;
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) {
      const int q = l8 + 8 * j;
      if (q < R4) {
        const float4 g = go[p][j];
        float4 dfx = make_float4(0.f, 0.f, 0.f, 0.f), dfy = dfx;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 v = ldg4(base + bl.off[c] + 4 * q);
          dfx.x = fmaf(bl.dwx[c], v.x, dfx.x);
          dfx.y = fmaf(bl.dwx[c], v.y, dfx.y);
          dfx.z = fmaf(bl.dwx[c], v.z, dfx.z);
          dfx.w = fmaf(bl.dwx[c], v.w, dfx.w);
          dfy.x = fmaf(bl.dwy[c], v.x, dfy.x);
          dfy.y = fmaf(bl.dwy[c], v.y, dfy.y);
          dfy.z = fmaf(bl.dwy[c], v.z, dfy.z);
          dfy.w = fmaf(bl.dwy[c], v.w, dfy.w);
          if (gbase != nullptr && bl.w[c] != 0.f) {
            const float wc = bl.w[c];
            red_add_v4(gbase + bl.off[c] + 4 * q,
                       make_float4(wc * g.x, wc * g.y, wc * g.z, wc * g.w));
          }
        }
        dgx += f4_dot(g, dfx);
        dgy += f4_dot(g, dfy);
      }
    }
    gx[cx] += dgx * (0.5f * (float)(W - 1));
    if (p < 3) gx[cy] += dgy * (0.5f * (float)(H - 1));
  }
}

}  // namespace nvfi
