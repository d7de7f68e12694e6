// RK2 adjoint through the velocity MLP on the tensor cores with FP16-split operands (mlp_h.cuh): the
// product path of nvfi_render_backward step 4 (backward_tc.cu = the round-1 3xTF32 path, k_advect_bwd in
// backward.cu = the FP32 SIMT verification path).  Math: SURVEY.md Appendix E "RK2 step"; reference:
// autograd of models/tensorf_keyframe.py:575-611 + models/velocity_field.py:54-98.
//
// Per tile of 128 advected samples and per RK2 step (reverse order) the kernel runs the program
//     F1' (x0, no stash: gives the midpoint)  F2 (midpoint, stashed)  B2  F1 (x0, stashed)  B1
// so that only ONE evaluation's stash is live (per CTA 608 KB: it stays in L2).  A stashed forward
// evaluation leaves, per hidden layer l,
//     A_l = silu(h_l)   as the 64 KB FP16 hi|lo tile image the next layer's MMAs read anyway, copied to
//                       global memory by the TMA engine (cp.async.bulk, no thread instructions), and
//     S_l = silu'(h_l)  FP32, unit-major, stored by the epilogue threads (coalesced),
// so the backward pass evaluates no activation function at all.  With G_l = dL/dh_l (tile T_G) and
// A_{l-1} (tile T_A, TMA-loaded from the stash), per layer l = 4..0:
//     dX:  D0[m][k] = sum_n G_l[m][n] W_l[n][k]        A = T_G K-major,  B = W_l^T image (ring)
//     dW:  D1[n][k] = sum_m G_l[m][n] A_{l-1}[m][k]    A = T_G MN-major, B = T_A MN-major  (same bytes!)
//     G_{l-1} = D0 (.) S_{l-1} -> T_G (FP16 split)
//     dW flush: the accumulator lane IS the output unit n, and the packed gradient is [k][n], so the
//     32 threads of a warp add 128 contiguous bytes per red.global.add instruction (coalesced,
//     fire-and-forget; D1 ping-pongs in tensor memory so the flush runs under the next layer's MMAs)
//     db: column sums of G_{l-1} by a register butterfly in the epilogue (FP32, before the split)
// No operand is ever transposed by a thread: FP16 tiles can be read K-major and MN-major.
// The head (128 -> 6) is two small MMAs against a [128][16] tile of the upstream gradient.
// Tried and rejected on the bench workload (all parity-green): TMA bulk reduce-add of D1 staged in T_A
// (353 ms: the staging tile, the reduction's read and the next A_{l-2} load serialise), per-CTA private
// partials with plain read-modify-write (291 ms: two exposed L2 round trips per layer), constant-ones
// MMAs for the bias sums (16 tiny MMAs per layer cost ~650 cycles of tensor time), 4 dedicated flush
// warps that read D1 out under the next layer's epilogue (259 ms against 222: 21 warps cap the kernel at
// 80 registers, and 4 warps push the 64 KB of reductions per layer slower than 16 do), vector reductions
// (red.global.add.v4.f32, kept: same time — the flush is bound by RED bytes per SM, ~12 B/clk), the flush
// of layer l + 1 interleaved, 8 columns per group, with the epilogue of layer l (249 ms: the reductions
// stall the in-order warps just the same, and they now sit BEFORE the barrier that releases dX(l - 1)
// instead of running under the next layer's MMAs), the two stashed evaluations of a single-step call as the two
// tiles of ONE paired evaluation with two stash slots (273 ms: the pair took 77 K cycles against 2 x 41 K —
// the in-place tile images must be read out by the TMA engine before the next epilogue may overwrite them,
// and 2 x 672 KB of stash per CTA no longer fit the L2 at all).  Timing experiments with parts switched
// off: without the S stash stores 229 ms (L2 write bandwidth is not the limiter), without the flush's
// reductions 210 ms (the flush costs 8 % on the critical path, not the 20 % its phase length suggests).
//
// FP16 range: the upstream gradient of a tile is scaled by a power of two so that its largest
// component is in [8, 16) (12 binades of headroom for growth through the layers, 2^-29 of the tile
// maximum as absolute resolution); every result is unscaled when it leaves the tensor cores.
#include <cstring>

#include "backward_common.cuh"

#ifdef NVFI_TIMELINE
// (tag, clock64) pairs of CTA 0: who = 0 -> worker thread 0 (first half of the buffer), who = 1 -> lane 0 of
// the issuer warp (second half).  Buffer pointer and counters live in shared memory, so that a mark is one
// shared-memory round trip and one fire-and-forget 16-byte store (with everything in global memory a mark
// cost ~400 cycles).  Recording starts with the 20th tile (warm caches, steady state).
__device__ long long* g_tlh_buf = nullptr;
__device__ int g_tlh_cap = 0;
__device__ int g_tlh_n[2] = {0, 0};
__shared__ long long* tl_ptr[2];
__shared__ int tl_cnt[2], tl_max, tl_tiles;
__device__ __forceinline__ void tlh_mark_at(int tag, int who, long long clk) {
  if (blockIdx.x == 0 && threadIdx.x == (who ? 512 : 0)) {
    if (tag == 0) ++tl_tiles;
    const int i = tl_cnt[who];
    if (i < tl_max && tl_tiles >= 20) {
      *reinterpret_cast<longlong2*>(tl_ptr[who] + 2 * i) = make_longlong2((long long)tag, clk);
      tl_cnt[who] = i + 1;
    }
  }
}
#define NVFI_TLH(tag, who) tlh_mark_at((tag), (who), clock64())
// pseudo-event with a given time stamp (e.g. the latest arrival of the 16 worker warps on a barrier)
#define NVFI_TLH_AT(tag, who, clk) tlh_mark_at((tag), (who), (clk))
__device__ __forceinline__ void tlh_init() {
  if (threadIdx.x == 0) {
    const int half = g_tlh_cap / 2;
    tl_ptr[0] = g_tlh_buf;
    tl_ptr[1] = g_tlh_buf + half;
    tl_cnt[0] = tl_cnt[1] = 0;
    tl_max = (blockIdx.x == 0 && g_tlh_buf != nullptr) ? half / 2 : 0;
    tl_tiles = 0;
  }
}
#endif
#include "mlp_h.cuh"

namespace nvfi {
namespace thb {

constexpr int NT = th::kThreads;            // 512 worker threads (+ the issuer warp)
constexpr uint32_t kStages = 2;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColD0 = 0;              // input-gradient accumulator (forward evaluations: 0 / 128)
constexpr uint32_t kColD1 = 256;            // weight-gradient accumulators: layer l -> 256 + 128 (l & 1)
constexpr uint32_t kColDh = 384;            // head weight gradient (16 columns; read before layer 3 writes there)
constexpr uint32_t kColGhi = 128, kColGlo = 192;   // G_l as the TS-form A operand of dX (two FP16 per column)

// per-CTA global scratch (byte offsets)
constexpr size_t kWsStashA = 0;
constexpr size_t kWsStashS = th::kStashABytes;
constexpr size_t kWsXsteps = kWsStashS + th::kStashSFloats * sizeof(float);
constexpr size_t kWsBytes = kWsXsteps + (size_t)MAX_RK2_STEPS * 3 * NVFI_TM * sizeof(float);
static_assert(kWsBytes <= (size_t)WS_CTA_F * sizeof(float), "per-CTA workspace of k_advect_bwd_h exceeds WS_CTA_F");

struct BwdTile {
  alignas(128) unsigned char gw_hi[4096];   // upstream gradient of the head, [128 samples][16] FP16 (no swizzle)
  alignas(128) unsigned char gw_lo[4096];
  float x0[3][NVFI_TM];
  float xm[3][NVFI_TM];
  float gbar[3][NVFI_TM];
  float gm[3][NVFI_TM];
  float w0[6][NVFI_TM];
  float w1[6][NVFI_TM];
  float gout[6][NVFI_TM];   // in: dL/d(basis weights); out (rows 0..2): dL/d(x, y, z) of the eval input
  float tvec[NVFI_TM];
  unsigned char gate0[NVFI_TM], gate1[NVFI_TM], reverted[NVFI_TM];
  int gidx[NVFI_TM];
  int q_idx[NVFI_TM + NT];
  int warp_cnt[2][NT / 32];
  int batch;
  unsigned gmax;            // bits of the largest |upstream gradient| of the tile
  unsigned gmax_l[2];       // JVP: the same per layer (ping-pong)
  unsigned gmax3[3];        // largest |G_{L-1}| recorded by layer L (slot L % 3), read by layer L - 1
  th::Issuer iss;           // weight-ring state of the issuer warp between calls
};

// ld.global.cg as a volatile asm: consecutive calls are issued back to back
__device__ __forceinline__ float4 ldcg4_now(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// The stash of an evaluation is dead once its backward pass has read it: tell the L2 so, or every dirty
// line is written back to HBM when the next evaluation's stash displaces it (148 CTAs x 608 KB = 90 MB
// cycle through the cache: ncu showed 9.2 KB of DRAM writes per sample, the whole stash).
__device__ __forceinline__ void discard_l2(const void* p) {
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}
// Sum over the 32 rows (lanes) of a warp of 8 per-lane values: lane j returns the sum of v[j & 7].
// Butterfly: 7 exchanges that halve the value count, then 2 plain ones.
__device__ __forceinline__ float colsum8(const float v[8], int lane) {
  float w4[4], w2[2];
  const bool u4 = lane & 4, u2 = lane & 2, u1 = lane & 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = u4 ? v[i] : v[i + 4], keep = u4 ? v[i + 4] : v[i];
    w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = u2 ? w4[i] : w4[i + 2], keep = u2 ? w4[i + 2] : w4[i];
    w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  float w = (u1 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, u1 ? w2[0] : w2[1], 1);
  w += __shfl_xor_sync(0xffffffffu, w, 8);
  w += __shfl_xor_sync(0xffffffffu, w, 16);
  return w;
}

// ---- backward through one weight-net evaluation -------------------------------------------
// In: T.gout[n][m] = dL/d(basis weights); tile tA holds A_4 of the stashed evaluation that has just run;
// stash_a / stash_s = that evaluation's stash; (xs, ys, zs)[m] = its input.  Out: T.gout[0..2][m] =
// dL/d(x, y, z) through the network input; weight gradients added to the packed gradient buffers, bias
// and head gradients to the register accumulators.  Whole CTA (13 block barriers).
// JVP = 1: the reverse pass of the forward-mode evaluation of the PDE loss (mlp_h.cuh, k_pde_jac_h): rows
// are 6 points x (value + 4 tangents) per lane quadrant; the activation step couples the rows of a point,
//     g_h0 = g_a0 S + sum_j g_aj S2_j,   g_hj = g_aj S        (S = silu'(h_0), S2_j = silu''(h_0) h_j),
// bias gradients come from the value rows only and no input gradient is produced.
template <int JVP = 0>
__device__ void bwd_eval_h(th::Ctl1& c, th::Issuer& is_shared, BwdTile& T, uint32_t tA, uint32_t tG,
                           const NvfiRenderGrads& D, const unsigned char* __restrict__ stash_a,
                           const float* __restrict__ stash_s, const float* xs, const float* ys, const float* zs,
                           uint32_t& dphase, uint32_t& wphase, uint32_t& aphase, float (&acc_head)[6],
                           float (&acc_bias)[6], const float* __restrict__ stash_s2 = nullptr) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool value_row = !JVP || (lane < 30 && lane % 5 == 0);
  NVFI_TLH(98, 0);
  const uint32_t gw_hi = tc::smem_u32(T.gw_hi), gw_lo = tc::smem_u32(T.gw_lo);

  // ---- scale of the tile: largest |dL/dw| -> [8, 16)
  if (tid < NVFI_TM) {
    float mx = 0.f;
#pragma unroll
    for (int n = 0; n < 6; ++n) mx = fmaxf(mx, fabsf(T.gout[n][tid]));
    const unsigned b = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
    if (lane == 0 && b) atomicMax(&T.gmax, b);
  }
  __syncthreads();   // (S0)
  float scale = 1.f, inv_scale = 1.f;
  {
    const unsigned b = T.gmax;
    if (b) {
      unsigned ex = b >> 23;
      ex = ex < 30u ? 30u : (ex > 250u ? 250u : ex);
      scale = __uint_as_float((257u - ex) << 23);
      inv_scale = __uint_as_float((ex - 3u) << 23);
    }
  }

  if (warp == th::kIssuerWarp) {   // ==================================================== issuer warp
    th::Issuer is = is_shared;
    NVFI_TLH(1100, 1);
    th::ring_top_up(c, is, false);
    __syncthreads();   // (A) upstream-gradient tile written
    tc::tc_fence_after();
    NVFI_TLH(1101, 1);
    const uint32_t tb = is.tb;
    const uint32_t id128 = th::idesc_f16(128), id32 = th::idesc_f16(32);
    const uint32_t hs = th::kDescHiSw128;
    if (tc::elect_one()) th::bulk_wait0();   // the stash copies of the forward evaluation have landed
    __syncwarp();
    // ---- head, dX: D0[m][u] = sum_n gw[m][n] W5[n][u]   (K = 16: one step)
    {
      const uint32_t st = th::ring_acquire(c, is);
      if (tc::elect_one()) {
        const uint32_t a_h = th::desc_lo(gw_hi, 128), a_l = th::desc_lo(gw_lo, 128);
        const uint32_t b_h = th::desc_lo(st), b_l = th::desc_lo(st + 128u * 128u);
        th::mma_f16_ss(tb + kColD0, a_h, th::kDescHiSmallK, b_h, hs, id128, 0u);
        th::mma_f16_ss(tb + kColD0, a_l, th::kDescHiSmallK, b_h, hs, id128, 1u);
        th::mma_f16_ss(tb + kColD0, a_h, th::kDescHiSmallK, b_l, hs, id128, 1u);
        tc::tc_commit(&c.empty[is.c_stage]);
        tc::tc_commit(&c.dbar);
      }
      __syncwarp();
      th::ring_advance(is);
    }
    // ---- head, dW: Dh[k][n] = sum_m A_4[m][k] gw[m][n]   (A = tA MN-major, B = gw MN-major, N = 16)
    if (tc::elect_one()) {
      const uint32_t idh = th::idesc_f16(16, 1, 1);
#pragma unroll 2
      for (uint32_t ks = 0; ks < 8; ++ks) {
        const uint32_t a_h = th::desc_lo(tA + ks * 2048u, 16384), a_l = th::desc_lo(tA + th::kLoOff + ks * 2048u, 16384);
        const uint32_t b_h = th::desc_lo(gw_hi + ks * 512u, 256), b_l = th::desc_lo(gw_lo + ks * 512u, 256);
        th::mma_f16_ss(tb + kColDh, a_h, hs, b_h, th::kDescHiSmallMN, idh, ks ? 1u : 0u);
        th::mma_f16_ss(tb + kColDh, a_l, hs, b_h, th::kDescHiSmallMN, idh, 1u);
        th::mma_f16_ss(tb + kColDh, a_h, hs, b_l, th::kDescHiSmallMN, idh, 1u);
      }
      tc::tc_commit(&c.wbar);
    }
    __syncwarp();
    NVFI_TLH(1102, 1);
    tc::mbar_wait(&c.wbar, wphase & 1);   // A_4 consumed: tile tA is free
    ++wphase;
    NVFI_TLH(1103, 1);
    if (tc::elect_one()) {
      tc::mbar_expect_tx(&c.abar, th::kTileBytes);
      tc::bulk_g2s_u32(tA, stash_a + 32768 + (size_t)3 * th::kTileBytes, th::kTileBytes, &c.abar);   // A_3
    }
    // dX of layer l: D0[m][k] = sum_n G_l[m][n] W_l[n][k]
    // Both K blocks of W_l^T are awaited BEFORE the block barrier that publishes G_l (dx_ready), so that after
    // the barrier the MMAs are issued without a single load/store-unit instruction: the 16 worker warps start
    // the dW flush right behind that barrier, and an mbarrier probe of this warp then queues behind their
    // reductions (phase timeline: the issue of dX(l - 1) used to last exactly as long as the flush of layer l).
    auto dx_ready = [&]() {
#ifndef NVFI_BWD_NO_PREACQUIRE
      th::ring_wait(c, is, 0);
      th::ring_wait(c, is, 1);
#endif
    };
    auto issue_dx = [&](int l) {
      const uint32_t nx = (l == 0) ? 32u : 128u;
      const uint32_t idx = (l == 0) ? id32 : id128;
#pragma unroll 1
      for (uint32_t kb = 0; kb < 2; ++kb) {
#ifndef NVFI_BWD_NO_PREACQUIRE
        const uint32_t st = is.ring_u32 + is.c_stage * th::kStageBytes;   // awaited by dx_ready()
#else
        const uint32_t st = th::ring_acquire(c, is);
#endif
        if (tc::elect_one()) {
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks) {
            // A = G_l from tensor memory (TS form: the MMAs only read the weights from shared memory)
            const uint32_t ta = tb + (kb * 4u + ks) * 8u;
            const uint32_t b_h = th::desc_lo(st) + ks * 2u, b_l = th::desc_lo(st + nx * 128u) + ks * 2u;
            th::mma_f16_ts(tb + kColD0, ta + kColGhi, b_h, hs, idx, (kb | ks) ? 1u : 0u);
            th::mma_f16_ts(tb + kColD0, ta + kColGlo, b_h, hs, idx, 1u);
            th::mma_f16_ts(tb + kColD0, ta + kColGhi, b_l, hs, idx, 1u);
          }
          tc::tc_commit(&c.empty[is.c_stage]);
          if (kb == 1) tc::tc_commit(&c.dbar);
        }
        __syncwarp();
        th::ring_advance(is);
      }
    };
    __syncwarp();
    th::ring_top_up(c, is, false);
    dx_ready();
    __syncthreads();   // (B) G_4 in tile tG
    tc::tc_fence_after();
    NVFI_TLH(1104, 1);
    issue_dx(4);
#pragma unroll 1
    for (int l = 4; l >= 0; --l) {
      NVFI_TLH(1105, 1);
      tc::mbar_wait(&c.abar, aphase & 1);   // A_{l-1} (l = 0: the encoding) is in tile tA
      ++aphase;
      tc::tc_fence_after();
      NVFI_TLH(1110 + l, 1);
      if (tc::elect_one()) {
        // dW: D1[n][k] = sum_m G_l[m][n] A_{l-1}[m][k]
        const uint32_t idw = th::idesc_f16(l == 0 ? 32 : 128, 1, 1);
        const uint32_t d1 = tb + kColD1 + 128u * (uint32_t)(l & 1);
#pragma unroll 2
        for (uint32_t ks = 0; ks < 8; ++ks) {
          const uint32_t g_h = th::desc_lo(tG + ks * 2048u, 16384), g_l = th::desc_lo(tG + th::kLoOff + ks * 2048u, 16384);
          const uint32_t a_h = th::desc_lo(tA + ks * 2048u, 16384), a_l = th::desc_lo(tA + th::kLoOff + ks * 2048u, 16384);
          th::mma_f16_ss(d1, g_h, hs, a_h, hs, idw, ks ? 1u : 0u);
          th::mma_f16_ss(d1, g_l, hs, a_h, hs, idw, 1u);
          th::mma_f16_ss(d1, g_h, hs, a_l, hs, idw, 1u);
        }
        tc::tc_commit(&c.wbar);
      }
      __syncwarp();
      NVFI_TLH(1120 + l, 1);
#ifndef NVFI_BWD_NO_PREACQUIRE
      th::ring_top_up(c, is, true);         // as soon as dX(l) has completed its stages take the next W^T blocks
      NVFI_TLH(1160 + l, 1);
      tc::mbar_wait(&c.wbar, wphase & 1);   // dW(l) done: tile tA is free for A_{l-2}
      ++wphase;
      NVFI_TLH(1170 + l, 1);
#else
      th::ring_top_up(c, is, false);
      tc::mbar_wait(&c.wbar, wphase & 1);
      ++wphase;
      th::ring_top_up(c, is, false);
#endif
      if (l > 0 && tc::elect_one()) {
        if (l >= 2) {
          tc::mbar_expect_tx(&c.abar, th::kTileBytes);
          tc::bulk_g2s_u32(tA, stash_a + 32768 + (size_t)(l - 2) * th::kTileBytes, th::kTileBytes, &c.abar);
        } else {   // the encoding: columns 0..31 of slab 0, hi and lo
          tc::mbar_expect_tx(&c.abar, 2u * th::kSlab);
          tc::bulk_g2s_u32(tA, stash_a, th::kSlab, &c.abar);
          tc::bulk_g2s_u32(tA + th::kLoOff, stash_a + th::kSlab, th::kSlab, &c.abar);
        }
      }
      __syncwarp();
      if (l > 0) dx_ready();                // W_{l-1}^T has landed: nothing but MMAs behind the barrier
      NVFI_TLH(1130 + l, 1);
      __syncthreads();   // (C1) G_{l-1} in tile tG
      tc::tc_fence_after();
      NVFI_TLH(1140 + l, 1);
      if (l > 0) issue_dx(l - 1);
      NVFI_TLH(1150 + l, 1);
    }
    dphase += 6;
    is_shared = is;
    return;
  }

  // ================================================================================ worker warps
  const int q = warp & 3, h = warp >> 2;
  const int m = q * 32 + lane;                 // sample (sample-major steps) / unit n (flush steps) of this thread
  const uint32_t tb = c.tmem_base;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const uint32_t row_u32 = tG + (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
  const uint32_t x7 = (uint32_t)(m & 7);

  if (tid < NVFI_TM) {   // upstream gradient of the head -> [128][16] FP16 tile (columns 6..15 zero)
    float v[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) v[n] = (n < 6) ? T.gout[n][m] * scale : 0.f;
    uint4 hi, lo;
    th::split8(v, hi, lo);
    const uint32_t off = (uint32_t)((m >> 3) * 256 + (m & 7) * 16);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    th::st_shared_v4(gw_hi + off, hi);
    th::st_shared_v4(gw_hi + off + 128u, z);
    th::st_shared_v4(gw_lo + off, lo);
    th::st_shared_v4(gw_lo + off + 128u, z);
    th::fence_async_smem();
  }
  if (warp < 6 && value_row)   // db5[n] = sum_m gout[n][m]: warp n, each lane 4 samples (reduced at kernel end)
    acc_bias[5] += T.gout[warp][lane] + T.gout[warp][lane + 32] + T.gout[warp][lane + 64] + T.gout[warp][lane + 96];
  if (tid == 0) T.gmax3[5 % 3] = 0u;   // the head layer records first; the other slots are cleared one layer ahead
  tc::tc_fence_before();
  NVFI_TLH(100, 0);
  __syncthreads();   // (A)
  NVFI_TLH(101, 0);
  if (tid == 0) T.gmax = 0u;

  // S_l[unit][sample] of this thread's 32 columns: coalesced 16-byte loads, issued one phase ahead
  float4 sv[4][2];
  auto load_s = [&](int l) {
    const float4* sp = reinterpret_cast<const float4*>(stash_s) + ((size_t)l * 32 + h * 2) * NVFI_TM + m;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      sv[g][0] = ldcg4_now(sp + (size_t)(g * 8) * NVFI_TM);
      sv[g][1] = ldcg4_now(sp + (size_t)(g * 8 + 1) * NVFI_TM);
    }
  };
  load_s(4);
  float inv_cum = inv_scale;   // 1 / (cumulative scale of G_L), updated layer by layer
#pragma unroll 1
  for (int L = 5; L >= 0; --L) {
    const float inv_L = inv_cum;   // un-scaling of this layer's weight gradient (G_L carries the cumulative scale)
    NVFI_TLH(110 + L, 0);
    tc::mbar_wait(&c.dbar, dphase & 1);   // D0 = dX(L)
    ++dphase;
    tc::tc_fence_after();
    NVFI_TLH(120 + L, 0);
    uint4 ghi[4], glo[4];
    float bsum = 0.f;
    if (L > 0) {
      // G_{L-1} = D0 (.) S_{L-1}, FP16 split, held in registers until tile tG is free
      const uint32_t dcol = tb + lane_base + kColD0 + (uint32_t)(h * 8);
      uint32_t raw[4][8];
#pragma unroll
      for (int g = 0; g < 4; ++g) tc::tmem_ld8_nowait(dcol + 32u * g, raw[g]);
      tc::tmem_ld_wait();
      if (!JVP) {
        // Re-scaling without a barrier: G shrinks by ~0.3 per layer and the per-sample gradients of a tile
        // span orders of magnitude, so by the first layers most entries would sit in FP16's subnormal range
        // (absolute resolution 2^-25 of the scaled value: the first layer's weight gradient was 7e-5 off).
        // Every layer records the largest |G_{L-1}| of the tile (gmax3[L % 3]); the NEXT layer — one block
        // barrier later — multiplies its result by the power of two that would have put that maximum into
        // [8, 16).  The information is one layer stale, which the 12 binades of headroom absorb.
        float f = 1.f;
        if (L < 5) {
          const unsigned bm = T.gmax3[(L + 1) % 3];
          if (bm) {
            unsigned ex = bm >> 23;
            ex = ex < 30u ? 30u : (ex > 250u ? 250u : ex);
            f = __uint_as_float((257u - ex) << 23);
            inv_cum *= __uint_as_float((ex - 3u) << 23);
          }
        }
        if (tid == 0) T.gmax3[(L + 2) % 3] = 0u;   // the slot layer L - 1 will record into
        float mx = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float s8[8] = {sv[g][0].x, sv[g][0].y, sv[g][0].z, sv[g][0].w, sv[g][1].x, sv[g][1].y, sv[g][1].z, sv[g][1].w};
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[i] = __uint_as_float(raw[g][i]) * (s8[i] * f);
            mx = fmaxf(mx, fabsf(v[i]));
          }
          th::split8(v, ghi[g], glo[g]);
          th::tmem_st4(tb + lane_base + kColGhi + (uint32_t)(16 * g + 4 * h), ghi[g]);   // A operand of dX(L - 1)
          th::tmem_st4(tb + lane_base + kColGlo + (uint32_t)(16 * g + 4 * h), glo[g]);
          // bias gradient of layer L - 1: column sums of G_{L-1} over the warp's 32 samples; lane j keeps
          // column 32 (j >> 3) + 8 h + (j & 7)
          const float cs = colsum8(v, lane);
          if ((lane >> 3) == g) bsum = cs;
        }
        tc::tmem_st_wait();
        {
          const unsigned bw = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
          if (lane == 0 && bw) atomicMax(&T.gmax3[L % 3], bw);
        }
      } else {
        // forward-mode rows: the value row collects the second-order terms of its 4 tangent rows, and the
        // tile is RE-SCALED per layer (power of two, largest |G_{L-1}| -> [8, 16)): the adjoints of a PDE
        // tile span many binades (one point with a large residual sets the head's scale) and shrink by
        // ~0.3 per layer, and FP16 hi + lo only resolves 2^-25 of the scaled values in absolute terms
        float vv[4][8];
        float mx = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float s8[8] = {sv[g][0].x, sv[g][0].y, sv[g][0].z, sv[g][0].w, sv[g][1].x, sv[g][1].y, sv[g][1].z, sv[g][1].w};
          const float4* sp2 = reinterpret_cast<const float4*>(stash_s2) + ((size_t)(L - 1) * 32 + h * 2 + g * 8) * NVFI_TM + m;
          const float4 q0 = ldcg4_now(sp2), q1 = ldcg4_now(sp2 + NVFI_TM);
          const float t8[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float ga = __uint_as_float(raw[g][i]);
            const float cj = ga * t8[i];
            const float u = cj + __shfl_down_sync(0xffffffffu, cj, 1);
            const float cross = __shfl_down_sync(0xffffffffu, u, 1) + __shfl_down_sync(0xffffffffu, u, 3);
            vv[g][i] = fmaf(ga, s8[i], value_row ? cross : 0.f);
            mx = fmaxf(mx, fabsf(vv[g][i]));
          }
        }
        if (tid == 0) T.gmax_l[(L & 1) ^ 1] = 0u;   // next layer's slot (its last readers passed barrier (C1))
        const unsigned bw = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
        if (lane == 0 && bw) atomicMax(&T.gmax_l[L & 1], bw);
        asm volatile("bar.sync 1, 512;" ::: "memory");   // the 16 worker warps
        float f = 1.f, finv = 1.f;
        {
          const unsigned bm = T.gmax_l[L & 1];
          if (bm) {
            unsigned ex = bm >> 23;
            ex = ex < 30u ? 30u : (ex > 250u ? 250u : ex);
            f = __uint_as_float((257u - ex) << 23);
            finv = __uint_as_float((ex - 3u) << 23);
          }
        }
        inv_cum *= finv;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = vv[g][i] * f;
          th::split8(v, ghi[g], glo[g]);
          th::tmem_st4(tb + lane_base + kColGhi + (uint32_t)(16 * g + 4 * h), ghi[g]);
          th::tmem_st4(tb + lane_base + kColGlo + (uint32_t)(16 * g + 4 * h), glo[g]);
          if (!value_row) {   // bias gradients come from the value rows only
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
          }
          const float cs = colsum8(v, lane);
          if ((lane >> 3) == g) bsum = cs;
        }
        tc::tmem_st_wait();
      }
      acc_bias[L - 1] = fmaf(bsum, inv_cum, acc_bias[L - 1]);
    } else if (!JVP && h == 0) {
      // dL/d(encoding) -> dL/d(x, y, z)  (encoder tangents)
      float ge[32];
      tc::tmem_ld32(tb + lane_base + kColD0, ge);
      const float qv[3] = {xs[m], ys[m], zs[m]};
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float s1, c1, s2, c2, s4, c4;
        tc::sincos_bounded(qv[i], s1, c1);
        tc::sincos_bounded(qv[i] * 2.f, s2, c2);
        tc::sincos_bounded(qv[i] * 4.f, s4, c4);
        T.gout[i][m] = inv_cum * (ge[i] + ge[4 + i] * c1 - ge[8 + i] * s1 +
                                    2.f * (ge[12 + i] * c2 - ge[16 + i] * s2) +
                                    4.f * (ge[20 + i] * c4 - ge[24 + i] * s4));
      }
    }
    NVFI_TLH(130 + L, 0);
    tc::mbar_wait(&c.wbar, wphase & 1);   // dW(L) done: tiles tG and tA are free
    ++wphase;
    tc::tc_fence_after();
    NVFI_TLH(140 + L, 0);
#ifndef NVFI_NO_STASH_DISCARD
    // The stash of layer L - 1 is dead: S_{L-1} is in this thread's G_{L-1}, A_{L-1} went through tile tA
    // into dW(L).  Drop the lines from the L2 now, layer by layer, so that the dirty footprint of the 148
    // CTAs (which run out of phase) stays near half of 148 x 608 KB.
    if (L >= 1) {   // 64 KB = 512 lines each: one line per worker thread
      discard_l2(reinterpret_cast<const unsigned char*>(stash_s) + (size_t)(L - 1) * 65536 + (size_t)tid * 128);
      if (JVP) discard_l2(reinterpret_cast<const unsigned char*>(stash_s2) + (size_t)(L - 1) * 65536 + (size_t)tid * 128);
      if (L <= 4) discard_l2(stash_a + 32768 + (size_t)(L - 1) * th::kTileBytes + (size_t)tid * 128);
    } else if (tid < 256) {
      discard_l2(stash_a + (size_t)tid * 128);   // the encoding (hi | lo slab 0)
    }
#endif
    if (L > 0) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t a = row_u32 + (uint32_t)(g >> 1) * th::kSlab + ((((uint32_t)(4 * g + h) & 7u) ^ x7) << 4);
        th::st_shared_v4(a, ghi[g]);
        th::st_shared_v4(a + th::kLoOff, glo[g]);
      }
      th::fence_async_smem();
    }
    tc::tc_fence_before();
    __syncthreads();   // (B) for the head, (C1) for the hidden layers
    NVFI_TLH(150 + L, 0);
    if (L >= 2) load_s(L - 2);   // for the next layer's epilogue, under this layer's flush
    if (L == 5) {
      if (h == 0) {   // dW5^T[k][n]: unit k = m
        uint32_t r[8];
        tc::tmem_ld8_nowait(tb + lane_base + kColDh, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int n = 0; n < 6; ++n) acc_head[n] = fmaf(__uint_as_float(r[n]), inv_L, acc_head[n]);
      }
      continue;
    }
    // ---- D1[n][k] (this thread: n = m, k = 32 h + i) is added to the packed gradient [k][n]; the flush
    //      overlaps the next layer's MMAs (D1 ping-pongs in tensor memory)
    if (L > 0 || h == 0) {
      // coalesced fire-and-forget VECTOR reductions into the [k / 4][n][k % 4] gradient image
      // (grad_layout_v4): 4 consecutive k per thread and instruction, 512 contiguous bytes per warp
      float* pp = D.g_vel_w[L] + ((size_t)(h * 8) * NVFI_TM + m) * 4;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float dv[16];
        tc::tmem_ld16(tb + lane_base + kColD1 + 128u * (uint32_t)(L & 1) + (uint32_t)(h * 32 + half * 16), dv);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          red_add_v4(pp + (size_t)(half * 4 + q4) * NVFI_TM * 4,
                     make_float4(dv[4 * q4] * inv_L, dv[4 * q4 + 1] * inv_L, dv[4 * q4 + 2] * inv_L, dv[4 * q4 + 3] * inv_L));
      }
    }
    tc::tc_fence_before();   // the D1 reads are ordered before the next barrier (the issuer reuses D1 two layers on)
    NVFI_TLH(160 + L, 0);
  }
  NVFI_TLH(170, 0);
  asm volatile("fence.proxy.async.global;" ::: "memory");   // the next evaluation's bulk copies rewrite the discarded lines
  NVFI_TLH(171, 0);
}

// v = basis(w, x): dL/dw and the explicit dL/dx from dL/dv (as in backward.cu)
__device__ __forceinline__ void basis_bwd(const float w[6], float x, float y, float z, const float gv[3],
                                          float gw[6], float gxe[3]) {
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

// 17 warps are allocated as 20: 96 registers per thread is the cap.
__global__ void __launch_bounds__(th::kLaunchThreads, 1)
    k_advect_bwd_h(const __grid_constant__ NvfiField F, const NvfiRenderArgs A, const NvfiRenderBuffers B,
                   const NvfiRenderGrads D, int S, long long total, int n_batches, int subs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* p = smem_raw;
  {
    const uint32_t a = tc::smem_u32(p);
    p += (1024u - (a & 1023u)) & 1023u;
  }
  const uint32_t tA = tc::smem_u32(p);                    // activation tile
  const uint32_t tG = tA + th::kTileBytes;                // gradient tile
  const uint32_t ring = tG + th::kTileBytes;              // kStages x 32 KB
  th::Ctl1& ctl = *reinterpret_cast<th::Ctl1*>(p + 2 * th::kTileBytes + kStages * th::kStageBytes);
  BwdTile& T = *reinterpret_cast<BwdTile*>(reinterpret_cast<unsigned char*>(&ctl) + ((sizeof(th::Ctl1) + 127) & ~(size_t)127));
  unsigned char* ws = reinterpret_cast<unsigned char*>(D.workspace) + (size_t)blockIdx.x * WS_CTA_F * sizeof(float);
  unsigned char* stash_a = ws + kWsStashA;
  float* stash_s = reinterpret_cast<float*>(ws + kWsStashS);
  float* xsteps = reinterpret_cast<float*>(ws + kWsXsteps);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // uniform RK2 schedule of this render call (models/tensorf_keyframe.py:577-609)
  float sched_dt[MAX_RK2_STEPS], sched_t[MAX_RK2_STEPS];
  int n_steps = 0;
  {
    float off = __fsub_rn(A.t, A.base_time), tc_ = A.t;
    while (fabsf(off) > 0.f && n_steps < MAX_RK2_STEPS) {
      float dt = fminf(fabsf(off), F.dt_max);
      dt = (off > 0.f) ? dt : -dt;
      sched_dt[n_steps] = dt;
      sched_t[n_steps] = tc_;
      off = __fsub_rn(off, dt);
      tc_ = __fsub_rn(tc_, dt);
      ++n_steps;
    }
  }

  // with the midpoints saved by the forward pass (single-step calls) the evaluation that only finds them is skipped
  const bool use_mid = (B.x_mid != nullptr) && n_steps == 1;
#ifdef NVFI_TIMELINE
  tlh_init();
#endif
  th::setup(ctl, F.vel_net, nullptr, kTmemCols);
  if (tid == 0) {   // weight segments in the order one tile consumes them
    int n = 0;
    for (int k = 0; k + 1 < n_steps; ++k) {
      ctl.prog[n++] = th::SEG_FWD0;
      ctl.prog[n++] = th::SEG_FWD0;
    }
    for (int k = 0; k < n_steps; ++k) {   // [F1',] F2 (stashed), B2, F1 (stashed), B1
      if (!use_mid) ctl.prog[n++] = th::SEG_FWD0;
      ctl.prog[n++] = th::SEG_FWD0;
      ctl.prog[n++] = th::SEG_BWD0;
      ctl.prog[n++] = th::SEG_FWD0;
      ctl.prog[n++] = th::SEG_BWD0;
    }
    ctl.prog_len = (uint32_t)n;
    T.gmax = 0u;
  }
  __syncthreads();
  if (warp == th::kIssuerWarp) T.iss.init(ctl, ring, kStages);
  __syncthreads();
  uint32_t dphase = 0, kphase = 0, wphase = 0, aphase = 0;

  int sub = subs;
  long long batch_base = 0;
  bool exhausted = false;
  int qc = 0, par = 0;
  unsigned long long n_done = 0;
  // per-thread partial sums of the head-layer weight gradient (unit k = tid < 128) and of the bias
  // gradients (unit n = tid < 128; the head's: warps 0..5), carried in registers across all tiles
  float acc_head[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, acc_bias[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

  // One tile = a short program of evaluations.  There is exactly ONE call site of the forward tile
  // evaluation and one of the backward tile evaluation (instruction-cache footprint); the small
  // per-sample glue around them is selected by `kind`.
  enum { K_FWD_A = 0, K_FWD_B, K_REV_A0, K_REV_B, K_BWD2, K_REV_A, K_BWD1 };
  const int rk = use_mid ? 4 : 5;   // evaluations per reverse step
  const int n_ops = 2 * (n_steps - 1) + rk * n_steps;

  for (;;) {
    while (qc < NVFI_TM && !exhausted) {
      if (sub == subs) {
        if (tid == 0) T.batch = atomicAdd(&B.counters[3], 1);
        __syncthreads();
        const int b = T.batch;
        __syncthreads();
        if (b >= n_batches) {
          exhausted = true;
          break;
        }
        batch_base = (long long)b * ((long long)subs * NT);
        sub = 0;
      }
      const long long idx = batch_base + (long long)sub * NT + tid;
      ++sub;
      bool push = false;
#ifndef NVFI_BWD_SERIAL_SCAN
      if (tid < NT && idx < total) {
        // all three streams are read at once (one memory latency instead of three dependent ones; the values
        // of samples that are not valid are never used: both buffers are fully allocated)
        const unsigned char v = B.valid[idx];
        const float g0 = D.g_x_adv[idx * 3], g1 = D.g_x_adv[idx * 3 + 1], g2 = D.g_x_adv[idx * 3 + 2];
        float m0 = 0.f, m1 = 0.f, m2 = 0.f;
        if (use_mid) {
          m0 = B.x_mid[idx * 3];
          m1 = B.x_mid[idx * 3 + 1];
          m2 = B.x_mid[idx * 3 + 2];
        }
        push = v && ((g0 != 0.f) | (g1 != 0.f) | (g2 != 0.f));
        // A sample whose RK2 midpoint lies outside the velocity gate moved with v1 = 0: x1 = x0 whatever the
        // network said at x0 or at the midpoint, so no gradient reaches the network through it (the adjoint
        // seeds of both evaluations are exactly zero).  With the saved midpoints that is known here.
        if (push && use_mid && gate_outside(F, m0, m1, m2)) push = false;
      }
#else
      if (tid < NT && idx < total && B.valid[idx]) {
        const float g0 = D.g_x_adv[idx * 3], g1 = D.g_x_adv[idx * 3 + 1], g2 = D.g_x_adv[idx * 3 + 2];
        push = (g0 != 0.f) | (g1 != 0.f) | (g2 != 0.f);
        if (push && use_mid &&
            gate_outside(F, B.x_mid[idx * 3], B.x_mid[idx * 3 + 1], B.x_mid[idx * 3 + 2]))
          push = false;
      }
#endif
      const unsigned bal = __ballot_sync(0xffffffffu, push);
      if (lane == 0 && warp < NT / 32) T.warp_cnt[par][warp] = __popc(bal);
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
    NVFI_TLH(0, 0);
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
      float gb[3] = {0.f, 0.f, 0.f}, xmv[3] = {0.f, 0.f, 0.f};   // loaded first: independent of the ray's loads
      if (live) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          gb[a] = D.g_x_adv[gi * 3 + a];
          if (use_mid) xmv[a] = B.x_mid[gi * 3 + a];
        }
      }
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
        T.gbar[a][tid] = gb[a];
        if (use_mid) T.xm[a][tid] = xmv[a];
      }
      T.gate0[tid] = gate_outside(F, xn[0], xn[1], xn[2]);
    }
    __syncthreads();
    NVFI_TLH(2, 0);

#pragma unroll 1
    for (int op = 0; op < n_ops; ++op) {
      int k, kind;
      if (op < 2 * (n_steps - 1)) {   // forward sweep over all but the last step
        k = op >> 1;
        kind = K_FWD_A + (op & 1);
      } else {                        // reverse sweep
        const int r = op - 2 * (n_steps - 1);
        k = n_steps - 1 - r / rk;
        kind = K_REV_A0 + (5 - rk) + r % rk;
      }
      const float dt = sched_dt[k], tcur = sched_t[k], hdt = 0.5f * dt;
      const float tmid = __fsub_rn(tcur, hdt);
      const bool at_mid = (kind == K_FWD_B || kind == K_REV_B || kind == K_BWD2);
      // ---- glue before the evaluation
      if (tid < NVFI_TM) {
        if (kind == K_FWD_A) {
#pragma unroll
          for (int a = 0; a < 3; ++a) xsteps[(k * 3 + a) * NVFI_TM + tid] = T.x0[a][tid];
        }
        if (kind == K_REV_A0 && k < n_steps - 1) {
#pragma unroll
          for (int a = 0; a < 3; ++a) T.x0[a][tid] = xsteps[(k * 3 + a) * NVFI_TM + tid];
        }
        if (kind != K_BWD2 && kind != K_BWD1) T.tvec[tid] = at_mid ? tmid : tcur;
      }
      __syncthreads();
      const float* xs = at_mid ? T.xm[0] : T.x0[0];
      const float* ys = at_mid ? T.xm[1] : T.x0[1];
      const float* zs = at_mid ? T.xm[2] : T.x0[2];
      if (kind == K_BWD2 || kind == K_BWD1) {
        // a stashed evaluation leaves A_4 in tile 1 (ping-pong, mlp_h.cuh): tile 1 is the activation tile of
        // the backward evaluation, tile 0 its gradient tile
        bwd_eval_h<0>(ctl, T.iss, T, tG, tA, D, stash_a, stash_s, xs, ys, zs, dphase, wphase, aphase, acc_head,
                      acc_bias);
      } else {
        float* wout = at_mid ? &T.w1[0][0] : &T.w0[0][0];
        const bool st = (kind == K_REV_B || kind == K_REV_A);
        // TS form (activations also in tensor memory, columns 384..511: free between backward evaluations):
        // SS-form MMAs of this shape are bound by the shared-memory operand reads (~120 cycles against 68)
        th::vel_net_tile_h<ACT_SILU, true>(ctl, T.iss, 0, wout, xs, ys, zs, T.tvec, tA, dphase, kphase, st ? tG : 0u,
                                    st ? stash_a : nullptr, st ? stash_s : nullptr);
      }
      // ---- glue after the evaluation
      NVFI_TLH(4, 0);
      if (tid < NVFI_TM) {
        const int m = tid;
        if (kind == K_FWD_A || kind == K_REV_A0) {   // midpoint m = x0 - dt/2 v0(x0)
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
        } else if (kind == K_FWD_B || kind == K_REV_B) {   // x1 = x0 - dt v1(m)
          const float xm = T.xm[0][m], ym = T.xm[1][m], zm = T.xm[2][m];
          const bool out1 = gate_outside(F, xm, ym, zm);
          const float w[6] = {T.w1[0][m], T.w1[1][m], T.w1[2][m], T.w1[3][m], T.w1[4][m], T.w1[5][m]};
          float v[3] = {0.f, 0.f, 0.f};
          if (!out1) basis_velocity(w, xm, ym, zm, v);
          const float x = T.x0[0][m], y = T.x0[1][m], z = T.x0[2][m];
          float nx = __fsub_rn(x, __fmul_rn(dt, v[0]));
          float ny = __fsub_rn(y, __fmul_rn(dt, v[1]));
          float nz = __fsub_rn(z, __fmul_rn(dt, v[2]));
          const bool rev = (F.vel_gate == NVFI_GATE_SUR) && gate_outside(F, nx, ny, nz);
          if (kind == K_FWD_B) {      // advance to the next step
            T.x0[0][m] = rev ? x : nx;
            T.x0[1][m] = rev ? y : ny;
            T.x0[2][m] = rev ? z : nz;
          } else {                    // adjoint of x1 = x0 - dt v1(m)
            T.gate1[m] = out1;
            T.reverted[m] = rev;
            float gw[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gxe[3] = {0.f, 0.f, 0.f};
            if (!rev && !out1) {
              const float gv[3] = {-dt * T.gbar[0][m], -dt * T.gbar[1][m], -dt * T.gbar[2][m]};
              basis_bwd(w, xm, ym, zm, gv, gw, gxe);
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) T.gout[i][m] = gw[i];
            T.gm[0][m] = gxe[0];
            T.gm[1][m] = gxe[1];
            T.gm[2][m] = gxe[2];
          }
        } else if (kind == K_BWD2) {   // adjoint of m = x0 - dt/2 v0(x0): the upstream gradient of evaluation 1
          float gmv[3];                // (it does not depend on w0, which the stashed evaluation that follows recomputes)
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            gmv[a] = T.gm[a][m] + T.gout[a][m];
            T.gm[a][m] = gmv[a];
          }
          float gw[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (!T.gate0[m]) {
            const float w[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const float gv[3] = {-hdt * gmv[0], -hdt * gmv[1], -hdt * gmv[2]};
            float gxe[3];
            basis_bwd(w, T.x0[0][m], T.x0[1][m], T.x0[2][m], gv, gw, gxe);
          }
#pragma unroll
          for (int i = 0; i < 6; ++i) T.gout[i][m] = gw[i];
        } else if (kind == K_BWD1) {   // dL/dx0 = g + g_m + explicit part of v0's basis + network-input part
          float gx0[3];
#pragma unroll
          for (int a = 0; a < 3; ++a) gx0[a] = T.gbar[a][m] + T.gm[a][m] + T.gout[a][m];
          if (!T.gate0[m]) {
            const float w[6] = {T.w0[0][m], T.w0[1][m], T.w0[2][m], T.w0[3][m], T.w0[4][m], T.w0[5][m]};
            const float gv[3] = {-hdt * T.gm[0][m], -hdt * T.gm[1][m], -hdt * T.gm[2][m]};
            float gw[6], gxe[3];
            basis_bwd(w, T.x0[0][m], T.x0[1][m], T.x0[2][m], gv, gw, gxe);
            gx0[0] += gxe[0];
            gx0[1] += gxe[1];
            gx0[2] += gxe[2];
          }
          T.gbar[0][m] = gx0[0];
          T.gbar[1][m] = gx0[1];
          T.gbar[2][m] = gx0[2];
        }
      }
      __syncthreads();
    }
  }
  if (tid < NVFI_TM) {
#pragma unroll
    for (int n2 = 0; n2 < 6; ++n2) red_add(D.g_vel_w[5] + tid * 8 + n2, acc_head[n2]);
  }
  if (tid < NT) {   // bias partials: this thread's column over its warp's rows, summed over all tiles
    const int col = 32 * (lane >> 3) + 8 * (warp >> 2) + (lane & 7);
#pragma unroll
    for (int l = 0; l < 5; ++l) red_add(D.g_vel_b[l] + col, acc_bias[l]);
  }
  if (warp < 6) {
    const float s5 = warp_sum(acc_bias[5]);
    if (lane == 0) red_add(D.g_vel_b[5] + warp, s5);
  }
  if (warp == th::kIssuerWarp && tc::elect_one()) th::bulk_wait0();   // outstanding bulk copies
  th::teardown(ctl, T.iss, kTmemCols);
  if (tid == 0 && n_done)
    atomicAdd(reinterpret_cast<unsigned long long*>(B.counters + 10), n_done);
}


// ---------------------------------------------------------------------------------------------------
// PDE loss of the velocity field on the tensor cores (NVFi.get_vel_loss, models/nvfi.py:69-84; math and
// the FP32 SIMT twin: pde.cu).  A tile = 24 points x 5 forward-mode rows (value, d/dx, d/dy, d/dz, d/dt)
// = 120 of the 128 rows; per tile: the stashed forward-mode evaluation (vel_net_tile_h<JVP>), the
// per-point residuals and output adjoints (one thread per point), the reverse pass (bwd_eval_h<JVP>).
// ---------------------------------------------------------------------------------------------------
constexpr int kPdePts = 24;
constexpr size_t kWsStashS2 = kWsStashS + th::kStashSFloats * sizeof(float);
static_assert(kWsStashS2 + th::kStashSFloats * sizeof(float) <= (size_t)WS_CTA_F * sizeof(float),
              "per-CTA workspace of k_pde_jac_h exceeds WS_CTA_F");

__global__ void __launch_bounds__(th::kLaunchThreads, 1)
    k_pde_jac_h(const __grid_constant__ NvfiField F, const float* __restrict__ xyzt, const float* __restrict__ va,
                long long n_total, float c_div, float c_tr, int want_grad, float* __restrict__ g_acc_out,
                double* __restrict__ loss_sums, const NvfiRenderGrads D, int* counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* p = smem_raw;
  {
    const uint32_t a = tc::smem_u32(p);
    p += (1024u - (a & 1023u)) & 1023u;
  }
  const uint32_t tA = tc::smem_u32(p);
  const uint32_t tG = tA + th::kTileBytes;
  const uint32_t ring = tG + th::kTileBytes;
  th::Ctl1& ctl = *reinterpret_cast<th::Ctl1*>(p + 2 * th::kTileBytes + kStages * th::kStageBytes);
  BwdTile& T = *reinterpret_cast<BwdTile*>(reinterpret_cast<unsigned char*>(&ctl) + ((sizeof(th::Ctl1) + 127) & ~(size_t)127));
  unsigned char* ws = reinterpret_cast<unsigned char*>(D.workspace) + (size_t)blockIdx.x * WS_CTA_F * sizeof(float);
  unsigned char* stash_a = ws + kWsStashA;
  float* stash_s = reinterpret_cast<float*>(ws + kWsStashS);
  float* stash_s2 = reinterpret_cast<float*>(ws + kWsStashS2);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  th::setup(ctl, F.vel_net, nullptr, kTmemCols);
  if (tid == 0) {
    ctl.prog[0] = th::SEG_FWD0;
    ctl.prog[1] = th::SEG_BWD0;
    ctl.prog_len = want_grad ? 2u : 1u;
    T.gmax = 0u;
    T.gmax_l[0] = T.gmax_l[1] = 0u;
  }
  __syncthreads();
  if (warp == th::kIssuerWarp) T.iss.init(ctl, ring, kStages);
  __syncthreads();
  uint32_t dphase = 0, kphase = 0, wphase = 0, aphase = 0;
  float acc_head[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, acc_bias[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long n_tiles = (n_total + kPdePts - 1) / kPdePts;
  double s_div = 0.0, s_tr = 0.0;
  // row m = 32 q + lane: point 6 q + lane / 5 of the tile, row type lane % 5 (lanes 30, 31 dead)
  const int jt = (lane < 30) ? lane % 5 : 5;
  const int pt = (tid >> 5) * 6 + lane / 5;

  for (;;) {
    if (tid == 0) T.batch = atomicAdd(counter, 1);
    __syncthreads();
    const long long tile = T.batch;
    __syncthreads();
    if (tile >= n_tiles) break;
    const long long p0 = tile * kPdePts;
    if (tid < NVFI_TM) {
      const bool live = jt < 5 && p0 + pt < n_total;
      const float* q = xyzt + (p0 + pt) * 4;
#pragma unroll
      for (int a = 0; a < 3; ++a) T.x0[a][tid] = live ? __ldg(q + a) : 0.f;
      T.tvec[tid] = live ? __ldg(q + 3) : 0.f;
    }
    __syncthreads();
    th::vel_net_tile_h<ACT_SILU, false, 1>(ctl, T.iss, 0, &T.w0[0][0], T.x0[0], T.x0[1], T.x0[2], T.tvec, tA, dphase,
                                           kphase, tG, stash_a, stash_s, stash_s2);
    // ---- per point: v, J_v, residuals, output adjoints (models/nvfi.py:69-84; same arithmetic as k_pde_jac)
    if (tid < NVFI_TM) {
#pragma unroll
      for (int i = 0; i < 6; ++i) T.gout[i][tid] = 0.f;
    }
    __syncthreads();
    if (tid < NVFI_TM && jt == 0 && p0 + pt < n_total) {
      const int m0 = tid;
      const float x = T.x0[0][m0], y = T.x0[1][m0], z = T.x0[2][m0];
      float w[6], dw[4][6];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        w[i] = T.w0[i][m0];
#pragma unroll
        for (int k = 0; k < 4; ++k) dw[k][i] = T.w0[i][m0 + 1 + k];
      }
      float v[3], J[3][4];
      basis_velocity(w, x, y, z, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float cc[3];
        basis_velocity(dw[k], x, y, z, cc);
        J[0][k] = cc[0];
        J[1][k] = cc[1];
        J[2][k] = cc[2];
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
        tr[i] = J[i][0] * v[0] + J[i][1] * v[1] + J[i][2] * v[2] + J[i][3] - __ldg(va + (p0 + pt) * 6 + 3 + i);
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
          g_acc_out[(p0 + pt) * 3 + i] = -gtr[i];
        }
        float gw[6], gxe[3];
        basis_bwd(w, x, y, z, gv, gw, gxe);   // gw = B^T gv (the position part gxe is not needed)
        gw[5] += gJ[0][1] - gJ[1][0];
        gw[4] += gJ[2][0] - gJ[0][2];
        gw[3] += gJ[1][2] - gJ[2][1];
#pragma unroll
        for (int i = 0; i < 6; ++i) T.gout[i][m0] = gw[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float gc[3] = {gJ[0][k], gJ[1][k], gJ[2][k]};
          float gd[6];
          basis_bwd(w, x, y, z, gc, gd, gxe);
#pragma unroll
          for (int i = 0; i < 6; ++i) T.gout[i][m0 + 1 + k] = gd[i];
        }
      }
    }
    __syncthreads();
    if (want_grad) {
      bwd_eval_h<1>(ctl, T.iss, T, tG, tA, D, stash_a, stash_s, T.x0[0], T.x0[1], T.x0[2], dphase, wphase, aphase,
                    acc_head, acc_bias, stash_s2);
      __syncthreads();
    }
  }
  if (want_grad) {
    if (tid < NVFI_TM) {
#pragma unroll
      for (int n2 = 0; n2 < 6; ++n2) red_add(D.g_vel_w[5] + tid * 8 + n2, acc_head[n2]);
    }
    if (tid < NT) {
      const int col = 32 * (lane >> 3) + 8 * (warp >> 2) + (lane & 7);
#pragma unroll
      for (int l = 0; l < 5; ++l) red_add(D.g_vel_b[l] + col, acc_bias[l]);
    }
    if (warp < 6) {
      const float s5 = warp_sum(acc_bias[5]);
      if (lane == 0) red_add(D.g_vel_b[5] + warp, s5);
    }
  }
  // ---- loss sums: warp shuffles, then one FP64 atomic per warp that has points
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s_div += __shfl_xor_sync(0xffffffffu, s_div, o);
      s_tr += __shfl_xor_sync(0xffffffffu, s_tr, o);
    }
    if (lane == 0 && warp < NVFI_TM / 32) {
      atomicAdd(loss_sums, s_div);
      atomicAdd(loss_sums + 1, s_tr);
    }
  }
  if (warp == th::kIssuerWarp && tc::elect_one()) th::bulk_wait0();
  th::teardown(ctl, T.iss, kTmemCols);
}


// Backward of the ReLU twin net (a_weight_net) for per-point upstream gradients ga (n, 3) of the
// acceleration a = basis_acceleration(net(enc(q)), x) (models/velocity_field.py:69-75), on the tensor
// cores: a stashed forward evaluation and the reverse pass of 128 points per tile (the FP32 SIMT twin is
// k_accnet_bwd in pde.cu).  D.g_vel_w / g_vel_b point at the a_weight_net accumulators.
__global__ void __launch_bounds__(th::kLaunchThreads, 1)
    k_accnet_bwd_h(const __grid_constant__ NvfiField F, const float* __restrict__ xyzt, const float* __restrict__ ga,
                   long long n_total, const NvfiRenderGrads D, int* counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* p = smem_raw;
  {
    const uint32_t a = tc::smem_u32(p);
    p += (1024u - (a & 1023u)) & 1023u;
  }
  const uint32_t tA = tc::smem_u32(p);
  const uint32_t tG = tA + th::kTileBytes;
  const uint32_t ring = tG + th::kTileBytes;
  th::Ctl1& ctl = *reinterpret_cast<th::Ctl1*>(p + 2 * th::kTileBytes + kStages * th::kStageBytes);
  BwdTile& T = *reinterpret_cast<BwdTile*>(reinterpret_cast<unsigned char*>(&ctl) + ((sizeof(th::Ctl1) + 127) & ~(size_t)127));
  unsigned char* ws = reinterpret_cast<unsigned char*>(D.workspace) + (size_t)blockIdx.x * WS_CTA_F * sizeof(float);
  unsigned char* stash_a = ws + kWsStashA;
  float* stash_s = reinterpret_cast<float*>(ws + kWsStashS);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  th::setup(ctl, F.acc_net, nullptr, kTmemCols);
  if (tid == 0) {
    ctl.prog[0] = th::SEG_FWD0;
    ctl.prog[1] = th::SEG_BWD0;
    ctl.prog_len = 2u;
    T.gmax = 0u;
  }
  __syncthreads();
  if (warp == th::kIssuerWarp) T.iss.init(ctl, ring, kStages);
  __syncthreads();
  uint32_t dphase = 0, kphase = 0, wphase = 0, aphase = 0;
  float acc_head[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, acc_bias[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long n_tiles = (n_total + NVFI_TM - 1) / NVFI_TM;
  for (;;) {
    if (tid == 0) T.batch = atomicAdd(counter, 1);
    __syncthreads();
    const long long tile = T.batch;
    __syncthreads();
    if (tile >= n_tiles) break;
    const long long i0 = tile * NVFI_TM;
    if (tid < NVFI_TM) {
      const bool live = i0 + tid < n_total;
#pragma unroll
      for (int a = 0; a < 3; ++a) T.x0[a][tid] = live ? __ldg(xyzt + (i0 + tid) * 4 + a) : 0.f;
      T.tvec[tid] = live ? __ldg(xyzt + (i0 + tid) * 4 + 3) : 0.f;
    }
    __syncthreads();
    th::vel_net_tile_h<ACT_RELU, false>(ctl, T.iss, 0, &T.w0[0][0], T.x0[0], T.x0[1], T.x0[2], T.tvec, tA, dphase,
                                        kphase, tG, stash_a, stash_s);
    if (tid < NVFI_TM) {
      const int m = tid;
      float g[3] = {0.f, 0.f, 0.f};
      if (i0 + m < n_total) {
        g[0] = __ldg(ga + (i0 + m) * 3);
        g[1] = __ldg(ga + (i0 + m) * 3 + 1);
        g[2] = __ldg(ga + (i0 + m) * 3 + 2);
      }
      const float x = T.x0[0][m], y = T.x0[1][m], z = T.x0[2][m];
      T.gout[0][m] = g[0];
      T.gout[1][m] = g[1];
      T.gout[2][m] = g[2];
      T.gout[3][m] = -y * g[1] - z * g[2];
      T.gout[4][m] = -x * g[0] - z * g[2];
      T.gout[5][m] = -x * g[0] - y * g[1];
    }
    __syncthreads();
    bwd_eval_h<0>(ctl, T.iss, T, tG, tA, D, stash_a, stash_s, T.x0[0], T.x0[1], T.x0[2], dphase, wphase, aphase,
                  acc_head, acc_bias);
    __syncthreads();
  }
  if (tid < NVFI_TM) {
#pragma unroll
    for (int n2 = 0; n2 < 6; ++n2) red_add(D.g_vel_w[5] + tid * 8 + n2, acc_head[n2]);
  }
  if (tid < NT) {
    const int col = 32 * (lane >> 3) + 8 * (warp >> 2) + (lane & 7);
#pragma unroll
    for (int l = 0; l < 5; ++l) red_add(D.g_vel_b[l] + col, acc_bias[l]);
  }
  if (warp < 6) {
    const float s5 = warp_sum(acc_bias[5]);
    if (lane == 0) red_add(D.g_vel_b[5] + warp, s5);
  }
  if (warp == th::kIssuerWarp && tc::elect_one()) th::bulk_wait0();
  th::teardown(ctl, T.iss, kTmemCols);
}

}  // namespace thb
}  // namespace nvfi

using namespace nvfi;

// development aid (called by nvfi_debug_timeline)
extern "C" int nvfi_debug_timeline_h(long long* dev_buf, int cap) {
#ifdef NVFI_TIMELINE
  const int zero[2] = {0, 0};
  NVFI_CUDA_OK(cudaMemcpyToSymbol(g_tlh_buf, &dev_buf, sizeof(dev_buf)));
  NVFI_CUDA_OK(cudaMemcpyToSymbol(g_tlh_cap, &cap, sizeof(cap)));
  NVFI_CUDA_OK(cudaMemcpyToSymbol(g_tlh_n, zero, sizeof(zero)));
#else
  (void)dev_buf;
  (void)cap;
#endif
  return NVFI_OK;
}

extern "C" int nvfi_launch_advect_bwd_h(const NvfiField* F, const NvfiRenderArgs* A, const NvfiRenderBuffers* B,
                                        const NvfiRenderGrads* D, int S, long long total, int sms, cudaStream_t st) {
  for (int l = 0; l < NVFI_VEL_LAYERS; ++l)
    if (!F->vel_net[l].himg || !F->vel_net[l].himgT) return NVFI_EINVAL;
  if (F->vel_net[5].n_pad != 8) return NVFI_EUNSUPPORTED;
  const size_t smem = 1024 + 2 * (size_t)th::kTileBytes + (size_t)thb::kStages * th::kStageBytes +
                      ((sizeof(th::Ctl1) + 127) & ~(size_t)127) + sizeof(thb::BwdTile);
  {
    const int rc = ensure_smem<thb::k_advect_bwd_h>(smem);
    if (rc != NVFI_OK) return rc;
  }
  const int subs = grab_subs(total, thb::NT, sms);
  const int per_batch = subs * thb::NT;
  const int n_batches = (int)((total + per_batch - 1) / per_batch);
  const int grid = n_batches < sms ? n_batches : sms;
  NVFI_LAUNCH(thb::k_advect_bwd_h, grid, th::kLaunchThreads, smem, st, *F, *A, *B, *D, S, total, n_batches, subs);
  NVFI_CUDA_OK(cudaGetLastError());
  return (int)cudaGetLastError();
}

// The forward-mode Jacobian + reverse pass of nvfi_pde_loss on the tensor cores (pde.cu dispatches here
// for NVFI_MLP_F16X3): gradients are added straight into the packed accumulators of `G`.
extern "C" int nvfi_launch_pde_jac_h(const NvfiField* F, const float* xyzt, const float* va, long long n,
                                     float c_div, float c_tr, int want_grad, const NvfiPdeGrads* G,
                                     double* loss_sums, int* counter, int sms, cudaStream_t st) {
  for (int l = 0; l < NVFI_VEL_LAYERS; ++l)
    if (!F->vel_net[l].himg || !F->vel_net[l].himgT) return NVFI_EINVAL;
  if (F->vel_net[5].n_pad != 8) return NVFI_EUNSUPPORTED;
  NvfiRenderGrads D;
  memset(&D, 0, sizeof(D));
  for (int l = 0; l < NVFI_VEL_LAYERS; ++l) {
    D.g_vel_w[l] = G->g_vel_w[l];
    D.g_vel_b[l] = G->g_vel_b[l];
  }
  D.workspace = G->workspace;
  D.workspace_bytes = G->workspace_bytes;
  const size_t smem = 1024 + 2 * (size_t)th::kTileBytes + (size_t)thb::kStages * th::kStageBytes +
                      ((sizeof(th::Ctl1) + 127) & ~(size_t)127) + sizeof(thb::BwdTile);
  {
    const int rc = ensure_smem<thb::k_pde_jac_h>(smem);
    if (rc != NVFI_OK) return rc;
  }
  const long long n_tiles = (n + thb::kPdePts - 1) / thb::kPdePts;
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  NVFI_LAUNCH(thb::k_pde_jac_h, grid, th::kLaunchThreads, smem, st, *F, xyzt, va, n, c_div, c_tr, want_grad,
              G->g_acc_pts, loss_sums, D, counter);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_launch_accnet_bwd_h(const NvfiField* F, const float* xyzt, const float* ga, long long n,
                                        const NvfiPdeGrads* G, int* counter, int sms, cudaStream_t st) {
  for (int l = 0; l < NVFI_VEL_LAYERS; ++l)
    if (!F->acc_net[l].himg || !F->acc_net[l].himgT) return NVFI_EINVAL;
  if (F->acc_net[5].n_pad != 8) return NVFI_EUNSUPPORTED;
  NvfiRenderGrads D;
  memset(&D, 0, sizeof(D));
  for (int l = 0; l < NVFI_VEL_LAYERS; ++l) {
    D.g_vel_w[l] = G->g_acc_w[l];
    D.g_vel_b[l] = G->g_acc_b[l];
  }
  D.workspace = G->workspace;
  D.workspace_bytes = G->workspace_bytes;
  const size_t smem = 1024 + 2 * (size_t)th::kTileBytes + (size_t)thb::kStages * th::kStageBytes +
                      ((sizeof(th::Ctl1) + 127) & ~(size_t)127) + sizeof(thb::BwdTile);
  {
    const int rc = ensure_smem<thb::k_accnet_bwd_h>(smem);
    if (rc != NVFI_OK) return rc;
  }
  const long long n_tiles = (n + NVFI_TM - 1) / NVFI_TM;
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  NVFI_LAUNCH(thb::k_accnet_bwd_h, grid, th::kLaunchThreads, smem, st, *F, xyzt, ga, n, *(&D), counter);
  return (int)cudaGetLastError();
}
