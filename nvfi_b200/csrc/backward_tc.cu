// RK2 adjoint through the velocity MLP on the tensor cores (tcgen05): the product path of
// nvfi_render_backward step 4 (k_advect_bwd in backward.cu is the FP32 SIMT verification
// path).  Math: SURVEY.md Appendix E "RK2 step"; reference: autograd of
// models/tensorf_keyframe.py:575-611 + models/velocity_field.py:54-98.
//
// Per tile of 128 advected samples and per RK2 step (reverse order):
//   F   recompute both weight-net evaluations on the tensor cores (mlp_tc.cuh), stashing the
//       pre-activations h_l[m][n] of the 5 hidden layers in a per-CTA L2-resident scratch
//   B   for each evaluation, walk the layers in reverse.  With G_l = dL/dh_l (128 samples x
//       128 units), A_{l-1} = silu(h_{l-1}):
//         G_{l-1}[m][k] = (sum_n G_l[m][n] W_l[n][k]) silu'(h_{l-1}[m][k])   reduction over UNITS
//         dW_l^T [k][n] = sum_m A_{l-1}[m][k] G_l[m][n]                      reduction over SAMPLES
//       Every GEMM is a K-major tcgen05.mma (MN-major TF32 operands need a second swizzle of
//       the same data, which shared memory has no room for):
//         dX: A = G_l in TENSOR MEMORY (lane = sample, column = unit — written there straight
//             from the previous epilogue's registers), B = W_l^T image block from the ring:
//             the forward machinery with other weights and another epilogue;
//         dW: A = A_{l-1}^T in tensor memory (lane = input unit, column = sample), built by
//             reading the stash TRANSPOSED — a coalesced global read —, B = G_l^T in shared
//             memory (row = unit, 128-byte rows of 32 samples, 128B swizzle), built the same
//             way from a per-CTA global copy of G_l.  The L2 round trip IS the transpose.
//       Both accumulate in TMEM (D0 = dX, D1 = dW^T); the epilogue turns D0 into G_{l-1}
//       (registers -> global copy -> TMEM) and adds D1 to the CTA's partial weight gradient.
//   The 6-wide head layer and the position-encoding chain rule are FP32 SIMT (tiny).
// All GEMMs use the 3-term TF32 split (FP32-grade) unless the single-pass mode is selected.
#include "backward_common.cuh"
#include "mlp_tc.cuh"

namespace nvfi {
namespace tcb {

constexpr int NT = tc::kThreads;            // 512 worker threads (+ the issuer warp)
constexpr uint32_t kBwdStages = 2;          // ring depth: shared memory is needed for G
constexpr int kLayerF = NVFI_TM * NVFI_TM;  // 16384 floats

// Per-CTA global scratch (float offsets): the stashes of the two evaluations of a step, two
// copies of G, the step positions.  Weight and bias gradients go straight to the packed
// gradient buffers with red.global.add (no per-CTA partials, no reduce kernels).
constexpr int kStashF = 5 * kLayerF + 32 * NVFI_TM;   // [5 layers][m][n] + enc[m][32]
constexpr int TW_STASH = 0;                     // [2 evals][kStashF]
constexpr int TW_XSTEPS = TW_STASH + 2 * kStashF;
constexpr int TW_TOTAL = TW_XSTEPS + MAX_RK2_STEPS * 3 * NVFI_TM;
// The per-tile dW^T accumulator (128 x 128 FP32 in TMEM) reaches the packed gradient through the
// TMA engine: staged in shared memory, then cp.reduce.async.bulk .add.f32 (the reduction runs in
// L2), one 512-byte row per operation.  red.global.add from the threads costs the LSU ~1 cycle per
// ELEMENT (16 K cycles per layer and tile, measured); private per-CTA partials, atomic or plain
// read-modify-write, were slower still.
constexpr uint32_t kStagePitch = 528;   // staging row pitch (512 B + 16 B): conflict-free float4 stores

struct BwdTile {
  float x0[3][NVFI_TM];
  float xm[3][NVFI_TM];
  float gbar[3][NVFI_TM];
  float gm[3][NVFI_TM];
  float w0[6][NVFI_TM];
  float w1[6][NVFI_TM];
  float gout[6][NVFI_TM];   // in: dL/d(basis weights); out (rows 0..2): dL/d(x, y, z) of the eval input
  float tvec[NVFI_TM];
  float w5s[6][NVFI_TM];    // head weights W5[n][k]
  unsigned char gate0[NVFI_TM], gate1[NVFI_TM], reverted[NVFI_TM];
  int gidx[NVFI_TM];
  int q_idx[NVFI_TM + NT];
  int warp_cnt[2][NT / 32];
  int batch;
  // weight-ring state of the issuer warp (kept out of the workers' registers).  Two rings: W^T
  // blocks of the backward evaluations in the 2 dedicated stages, W blocks of the forward
  // recomputation in 4 stages that alias the G^T tile (idle during forward evaluations).
  tc::Issuer iss, iss_fwd;
};

// Development aid: when a buffer is registered with nvfi_debug_timeline, thread 0 of CTA 0
// records (tag, clock64) pairs at the phase boundaries below (tools/probe_timeline.py).
__device__ long long* g_tl_buf = nullptr;
__device__ int g_tl_cap = 0;
__device__ int g_tl_n = 0;
#ifdef NVFI_TIMELINE
__device__ __forceinline__ void TL(int tag) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && g_tl_buf != nullptr) {
    const int i = g_tl_n;
    if (i + 2 <= g_tl_cap) {
      g_tl_buf[i] = tag;
      g_tl_buf[i + 1] = clock64();
      g_tl_n = i + 2;
    }
  }
}
#else
// compiled out of the product build: the pointer test alone costs every warp an L2 round trip
// per call (measured: ~5 % of the kernel).  Build with -DNVFI_TIMELINE for tools/probe_timeline.py.
__device__ __forceinline__ void TL(int) {}
#endif

// ld.global.cg as a volatile asm: consecutive calls are issued back to back (the compiler may not
// sink one below the use of another, which it otherwise does under register pressure and turns
// eight independent L2 round trips into a serial chain).
__device__ __forceinline__ float4 ldcg4_now(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ float sigmoid_fast(float h) {   // ex2.approx + rcp.approx, no fix-up code
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(h * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}
__device__ __forceinline__ float silu_d(float h) {   // SiLU'(h) = s + h s (1 - s)
  const float s = sigmoid_fast(h);
  return fmaf(h * s, 1.f - s, s);
}
__device__ __forceinline__ float silu_v(float h) { return h * sigmoid_fast(h); }

// ---- issuer warp ----------------------------------------------------------------------------
// dX of layer l:  D0[m][k] = sum_n G[m][n] Wt[k][n]   A = G in TMEM, B = W^T image blocks (ring)
__device__ __forceinline__ void issue_dx(tc::Ctl& c, tc::Issuer& is, int layer, int mode3) {
  const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t nx = (layer == 0) ? 32u : 128u;
  const uint32_t idesc = tc::instr_desc_tf32((int)nx);
  const uint32_t d0 = is.tb + tc::kColD;
  for (uint32_t kb = 0; kb < 4; ++kb) {
    tc::ring_top_up(c, is, mode3);
    tc::mbar_wait(&c.full[is.bar_base + is.c_stage], is.c_round & 1);
    tc::tc_fence_after();
    const uint32_t st = is.ring_u32 + is.c_stage * (uint32_t)tc::kStageBytes;
    const uint32_t w_hi = ((st >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t w_lo = (((st + nx * 128u) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t a_hi = is.tb + tc::kColAhi + kb * 32u, a_lo = is.tb + tc::kColAlo + kb * 32u;
    if (tc::elect_one()) {
#pragma unroll
      for (uint32_t ks = 0; ks < 4; ++ks) {
        const uint64_t wh = ((uint64_t)desc_hi << 32) | (uint64_t)(w_hi + ks * 2u);
        tc::mma_tf32_ts(d0, a_hi + ks * 8u, wh, idesc, (kb | ks) ? 1u : 0u);
        if (mode3) {
          const uint64_t wl = ((uint64_t)desc_hi << 32) | (uint64_t)(w_lo + ks * 2u);
          tc::mma_tf32_ts(d0, a_hi + ks * 8u, wl, idesc, 1u);
          tc::mma_tf32_ts(d0, a_lo + ks * 8u, wh, idesc, 1u);
        }
      }
      tc::tc_commit(&c.empty[is.bar_base + is.c_stage]);
      if (kb == 3) tc::tc_commit(&c.dbar);
    }
    __syncwarp();
    if (++is.c_stage == is.n_stages) {
      is.c_stage = 0;
      ++is.c_round;
    }
    --is.in_flight;
  }
}
// dW of layer l:  D1[k][n] = sum_m A^T[k][m] G^T[n][m]   A = A^T in TMEM, B = G^T in shared memory
__device__ __forceinline__ void issue_dw(tc::Ctl& c, tc::Issuer& is, uint32_t gt_hi, uint32_t gt_lo,
                                         int mode3) {
  const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t idesc = tc::instr_desc_tf32(128);
  const uint32_t d1 = is.tb + tc::kColD + 128u;
  const uint32_t gh = ((gt_hi >> 4) & 0x3FFFu) | (1u << 16), gl = ((gt_lo >> 4) & 0x3FFFu) | (1u << 16);
  if (tc::elect_one()) {
#pragma unroll 4
    for (uint32_t ks = 0; ks < 16; ++ks) {   // 4 K blocks (16 KB apart) x 4 steps of 8 samples (32 B)
      const uint32_t step = (ks >> 2) * 1024u + (ks & 3u) * 2u;
      const uint64_t bh = ((uint64_t)desc_hi << 32) | (uint64_t)(gh + step);
      const uint32_t a_hi = is.tb + tc::kColAhi + ks * 8u, a_lo = is.tb + tc::kColAlo + ks * 8u;
      tc::mma_tf32_ts(d1, a_hi, bh, idesc, ks ? 1u : 0u);
      if (mode3) {
        const uint64_t bl = ((uint64_t)desc_hi << 32) | (uint64_t)(gl + step);
        tc::mma_tf32_ts(d1, a_hi, bl, idesc, 1u);
        tc::mma_tf32_ts(d1, a_lo, bh, idesc, 1u);
      }
    }
    tc::tc_commit(&c.dbar);
  }
  __syncwarp();
}

// 16-column variant: units [32 h + 16 half, +16)
__device__ __forceinline__ void gt_store_col16(unsigned char* t_hi, unsigned char* t_lo, int q, int lane,
                                               int h, int half, const float v[16], int mode3) {
  const uint32_t base = (uint32_t)((q << 14) + (h << 12) + (half << 11) + ((lane & 3) << 2));
  const uint32_t lc = (uint32_t)(lane >> 2);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t off = base + (uint32_t)(((i >> 3) << 10) + ((i & 7) << 7)) + ((lc ^ (uint32_t)(i & 7)) << 4);
    const float hi = __uint_as_float(tc::to_tf32(v[i]));
    *reinterpret_cast<float*>(t_hi + off) = hi;
    if (mode3) *reinterpret_cast<float*>(t_lo + off) = v[i] - hi;
  }
}

// 16 already-split values (one of the hi / lo tiles) -> column m of rows 32 h + 16 half + i
__device__ __forceinline__ void gt_store_raw16(unsigned char* tile, int q, int lane, int h, int half,
                                               const float v[16]) {
  const uint32_t base = (uint32_t)((q << 14) + (h << 12) + (half << 11) + ((lane & 3) << 2));
  const uint32_t lc = (uint32_t)(lane >> 2);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t off = base + (uint32_t)(((i >> 3) << 10) + ((i & 7) << 7)) + ((lc ^ (uint32_t)(i & 7)) << 4);
    *reinterpret_cast<float*>(tile + off) = v[i];
  }
}

// Thread (unit k, sample block h): sum over the 32 samples of block h of row k of the G^T tile.
__device__ __forceinline__ float gt_row_sum(const unsigned char* t_hi, const unsigned char* t_lo, int k, int h,
                                            int mode3) {
  const uint32_t row = (uint32_t)((h << 14) + ((k >> 3) << 10) + ((k & 7) << 7));
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t ch = (uint32_t)((j ^ (k & 7)) << 4);   // lane-rotated chunk order: conflict-free
    const float4 a = *reinterpret_cast<const float4*>(t_hi + row + ch);
    s0 += (a.x + a.y) + (a.z + a.w);
    if (mode3) {
      const float4 b = *reinterpret_cast<const float4*>(t_lo + row + ch);
      s1 += (b.x + b.y) + (b.z + b.w);
    }
  }
  return s0 + s1;
}

__device__ __forceinline__ float ldcg_now(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// 16 values -> TMEM operand region columns [col, col + 16), hi/lo
__device__ __forceinline__ void tm_store16(uint32_t tb, uint32_t lane_base, uint32_t col, const float v[16],
                                           int mode3) {
  uint32_t hi[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) hi[i] = tc::to_tf32(v[i]);
  tc::tmem_st16(tb + lane_base + tc::kColAhi + col, hi);
  if (mode3) {
    uint32_t lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) lo[i] = __float_as_uint(v[i] - __uint_as_float(hi[i]));
    tc::tmem_st16(tb + lane_base + tc::kColAlo + col, lo);
  }
}

// TMA bulk reduction shared memory -> global (f32 add performed in L2), bulk async-group completion
__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst),
               "r"(tc::smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// ---- backward through one weight-net evaluation -------------------------------------------
// In: T.gout[n][m] = dL/d(basis weights), stash = the evaluation's pre-activations, (xs,ys,zs)[m]
// and tval = its input.  Out: T.gout[0..2][m] = dL/d(x, y, z) through the network input; weight
// and bias gradients added to the packed gradient buffers.  Whole CTA (11 block barriers).
//
// Per layer l (4..0), with G_l in the TMEM operand region and G_l^T in shared memory:
//   issuer: dX(l) MMAs -> D0                      workers: flush dW(l+1) from D1, db_l, prefetch h_{l-1}
//   workers: G_{l-1} = D0 * silu'(h_{l-1}) parked in D0 (frees 32 registers);
//            A_{l-1}^T (transposed stash read) -> operand region                       -- barrier B
//   issuer: dW(l) MMAs -> D1
//   workers: G_{l-1}: D0 -> operand region (lane = sample) and, transposed by 32 conflict-free
//            scalar stores per thread, -> shared memory (row = unit)                   -- barrier C
__device__ void bwd_eval_tc(tc::Ctl& c, tc::Issuer& is_shared, BwdTile& T, unsigned char* gt_hi,
                            unsigned char* gt_lo, float* __restrict__ ws, const NvfiRenderGrads& D,
                            const float* stash, const float* xs, const float* ys, const float* zs,
                            float tval, uint32_t& dphase, int mode3, float (&acc_head)[6],
                            float (&acc_bias)[6]) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == tc::kIssuerWarp) {
    tc::Issuer is = is_shared;   // ring state: registers while issuing, shared memory between calls
    const uint32_t gh = tc::uniform(tc::smem_u32(gt_hi)), gl = tc::uniform(tc::smem_u32(gt_lo));
    tc::ring_top_up(c, is, mode3);   // W^T blocks stream in under the head phase
    __syncthreads();   // (A) head done: G_4 in TMEM and G_4^T in shared memory
#pragma unroll 1
    for (int l = 4; l >= 0; --l) {
      tc::tc_fence_after();
      issue_dx(c, is, l, mode3);
      __syncthreads();   // (B) A_{l-1}^T in TMEM, D1 flushed
      tc::tc_fence_after();
      issue_dw(c, is, gh, gl, mode3);
      tc::ring_top_up(c, is, mode3);   // both stages full again before the next dX
      __syncthreads();   // (C) G_{l-1} in TMEM and G_{l-1}^T in shared memory
    }
    dphase += 10;
    is_shared = is;
    return;
  }
  const int q = warp & 3, h = warp >> 2;
  const int m = q * 32 + lane;                 // sample of this thread in the sample-major steps
  const int k = m;                             // unit   of this thread in the unit-major steps
  const uint32_t tb = c.tmem_base;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;

  TL(100);
  // ---- head layer (128 -> 6), FP32 SIMT: G_4 -> TMEM + shared memory; dW5, db5
  {
    // stash[l][unit][sample]: the sample-major reads (this thread = sample m) are coalesced scalar
    // loads; the unit-major reads (this thread = unit k) are 16-byte loads of its own row
    const float* hp = stash + ((size_t)4 * NVFI_TM + h * 32) * NVFI_TM + m;                    // h4[m][32 h + i]
    const float4* tp = reinterpret_cast<const float4*>(stash + ((size_t)4 * NVFI_TM + k) * NVFI_TM + h * 32);
    float gw[6];
#pragma unroll
    for (int n = 0; n < 6; ++n) gw[n] = T.gout[n][m];
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // two passes of 16 columns / 16 samples (not unrolled: bounded register use; each pass has all
    // of its 20 loads in flight at once)
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      float hv[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) hv[i] = ldcg_now(hp + (size_t)(half * 16 + i) * NVFI_TM);
      float4 at4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) at4[j] = ldcg4_now(tp + half * 4 + j);
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int kk = h * 32 + half * 16 + i;
        float s = 0.f;
#pragma unroll
        for (int n = 0; n < 6; ++n) s = fmaf(gw[n], T.w5s[n][kk], s);
        v[i] = s * silu_d(hv[i]);
      }
      tm_store16(tb, lane_base, (uint32_t)(h * 32 + half * 16), v, mode3);
      gt_store_col16(gt_hi, gt_lo, q, lane, h, half, v, mode3);
      // dW5^T[k][n] = sum_m silu(h4[m][k]) gout[n][m], this thread: unit k, samples [32 h + 16 half, +16)
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int mm = h * 32 + half * 16 + i;
        const float4 a4 = at4[i >> 2];
        const float a = silu_v((i & 3) == 0 ? a4.x : ((i & 3) == 1 ? a4.y : ((i & 3) == 2 ? a4.z : a4.w)));
#pragma unroll
        for (int n = 0; n < 6; ++n) acc[n] = fmaf(a, T.gout[n][mm], acc[n]);
      }
    }
    fence_async_smem();
#pragma unroll
    for (int n = 0; n < 6; ++n) acc_head[n] += acc[n];   // flushed once, at kernel end
    if (warp < 6)     // db5[n] = sum_m gout[n][m]: warp n, each lane 4 samples (reduced at kernel end)
      acc_bias[5] += T.gout[warp][lane] + T.gout[warp][lane + 32] + T.gout[warp][lane + 64] + T.gout[warp][lane + 96];
    tc::tmem_st_wait();
  }
  float4 rn[4];   // rows of the unit-major stash for the next A^T operand, loaded one phase ahead
  {
    const float4* sp = reinterpret_cast<const float4*>(stash + ((size_t)3 * NVFI_TM + k) * NVFI_TM + h * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) rn[j] = ldcg4_now(sp + j);
  }
  tc::tc_fence_before();
  __syncthreads();   // (A)
  TL(101);

#pragma unroll 1
  for (int l = 4; l >= 0; --l) {
    // ---- under the dX MMAs: prefetch row k of the unit-major stash (A_{l-1}^T: unit k, samples
    //      [32 h, +32) = 128 contiguous bytes per thread, 32 lines per warp instruction: the slow
    //      access goes where it is hidden), bias gradient
    if (l < 4) {
      // dW^T of layer l + 1: D1 -> staging rows in the dead G^T tile -> one TMA bulk reduce-add per
      // 512-byte row into the packed gradient (all hidden layers above 0 have 128 rows)
      {
        unsigned char* srow = gt_hi + (size_t)k * kStagePitch + h * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float dwv[16];
          tc::tmem_ld16(tb + lane_base + tc::kColD + 128u + (uint32_t)(h * 32 + half * 16), dwv);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(srow + half * 64 + j * 16) =
                make_float4(dwv[4 * j + 0], dwv[4 * j + 1], dwv[4 * j + 2], dwv[4 * j + 3]);
        }
      }
      fence_async_smem();
      workers_sync();
      if (h == 0) {
        bulk_reduce_add_f32(D.g_vel_w[l + 1] + k * NVFI_TM, gt_hi + (size_t)k * kStagePitch, 512u);
        bulk_commit();
        bulk_wait_read0();
      }
      workers_sync();
      // G_l^T tile for the dW MMAs: the operand region already holds G_l split hi | lo (the dX
      // MMAs only read it), transposed by conflict-free scalar stores
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float g[16];
        tc::tmem_ld16(tb + lane_base + tc::kColAhi + (uint32_t)(h * 32 + half * 16), g);
        gt_store_raw16(gt_hi, q, lane, h, half, g);
        if (mode3) {
          tc::tmem_ld16(tb + lane_base + tc::kColAlo + (uint32_t)(h * 32 + half * 16), g);
          gt_store_raw16(gt_lo, q, lane, h, half, g);
        }
      }
      fence_async_smem();   // visible to the dW MMAs (async proxy) and, after barrier B, to the bias sums
    }
    float4 r[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) r[j] = rn[j];   // first half: issued in the previous dW window
    if (l > 0) {   // second half: last, so that nothing forces these registers out before they are used
      const float4* sp = reinterpret_cast<const float4*>(stash + ((size_t)(l - 1) * NVFI_TM + k) * NVFI_TM + h * 32);
#pragma unroll
      for (int j = 4; j < 8; ++j) r[j] = ldcg4_now(sp + j);
    }
    TL(110 + l);
    tc::mbar_wait(&c.dbar, dphase & 1);   // dX accumulator
    ++dphase;
    tc::tc_fence_after();
    TL(120 + l);
    // ---- A_{l-1}^T into the TMEM operand region (the dX MMAs are done with G_l): all the dW MMAs
    //      wait for; the dX epilogue follows under them
    if (l > 0) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float a[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 rr = r[half * 4 + j];
          a[4 * j + 0] = silu_v(rr.x); a[4 * j + 1] = silu_v(rr.y);
          a[4 * j + 2] = silu_v(rr.z); a[4 * j + 3] = silu_v(rr.w);
        }
        tm_store16(tb, lane_base, (uint32_t)(h * 32 + half * 16), a, mode3);
      }
    } else {
      // encoding^T from the copy the forward recompute stashed (enc[m][32]): unit k < 32
      const float* sp = stash + (size_t)5 * kLayerF + (size_t)(h * 32) * 32 + (k & 31);
      const bool live = k < 32;
      float a0[16], a1[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) a0[i] = ldcg_now(sp + (size_t)i * 32);
#pragma unroll
      for (int i = 0; i < 16; ++i) a1[i] = ldcg_now(sp + (size_t)(16 + i) * 32);
#pragma unroll
      for (int i = 0; i < 16; ++i) a0[i] = live ? a0[i] : 0.f;
      tm_store16(tb, lane_base, (uint32_t)(h * 32), a0, mode3);
#pragma unroll
      for (int i = 0; i < 16; ++i) a1[i] = live ? a1[i] : 0.f;
      tm_store16(tb, lane_base, (uint32_t)(h * 32 + 16), a1, mode3);
    }
    tc::tmem_st_wait();
    TL(140 + l);
    tc::tc_fence_before();
    __syncthreads();   // (B)
    TL(150 + l);
    // ---- under the dW MMAs: the dX epilogue (sample-major) G_{l-1} = D0 * silu'(h_{l-1}), parked in
    //      D0 (the dW MMAs accumulate in D1), or the encoder chain rule
    if (l > 0) {
      // h_{l-1}[m][32 h + i]: coalesced scalar loads of the unit-major stash
      float hv[32];
      const float* hp = stash + ((size_t)(l - 1) * NVFI_TM + h * 32) * NVFI_TM + m;
#pragma unroll
      for (int i = 0; i < 32; ++i) hv[i] = ldcg_now(hp + (size_t)i * NVFI_TM);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t dcol = tb + lane_base + tc::kColD + (uint32_t)(h * 32 + half * 16);
        float part[16];
        tc::tmem_ld16(dcol, part);
        uint32_t gq[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) gq[i] = __float_as_uint(part[i] * silu_d(hv[half * 16 + i]));
        tc::tmem_st16(dcol, gq);
      }
      tc::tmem_st_wait();
    } else if (h == 0) {
      // dL/d(encoding) -> dL/d(x, y, z)  (encoder tangents)
      float ge[32];
      tc::tmem_ld32(tb + lane_base + tc::kColD, ge);
      const float qv[3] = {xs[m], ys[m], zs[m]};
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float s1, c1, s2, c2, s4, c4;
        tc::sincos_bounded(qv[i], s1, c1);
        tc::sincos_bounded(qv[i] * 2.f, s2, c2);
        tc::sincos_bounded(qv[i] * 4.f, s4, c4);
        T.gout[i][m] = ge[i] + ge[4 + i] * c1 - ge[8 + i] * s1 +
                       2.f * (ge[12 + i] * c2 - ge[16 + i] * s2) +
                       4.f * (ge[20 + i] * c4 - ge[24 + i] * s4);
      }
    }
    // ---- under the dW MMAs: bias gradient from the G_l^T tile (the MMAs only read it), first half
    //      of the next layer's transposing loads
    acc_bias[l] += gt_row_sum(gt_hi, gt_lo, k, h, mode3);   // unit k, samples [32 h, +32)
    if (l > 1) {
      const float4* sp = reinterpret_cast<const float4*>(stash + ((size_t)(l - 2) * NVFI_TM + k) * NVFI_TM + h * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) rn[j] = ldcg4_now(sp + j);
    }
    tc::mbar_wait(&c.dbar, dphase & 1);   // dW accumulator
    ++dphase;
    tc::tc_fence_after();
    TL(160 + l);
    // ---- G_{l-1}: D0 -> operand region, split hi | lo; the next dX starts right after barrier C
    //      (the dW flush and the G^T tile follow under its MMAs)
    if (l > 0) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float g[16];
        tc::tmem_ld16(tb + lane_base + tc::kColD + (uint32_t)(h * 32 + half * 16), g);
        tm_store16(tb, lane_base, (uint32_t)(h * 32 + half * 16), g, mode3);
      }
      tc::tmem_st_wait();
    }
    TL(170 + l);
    tc::tc_fence_before();
    __syncthreads();   // (C)
    TL(180 + l);
  }
  // dW^T of layer 0 (32 rows: lane quadrant 0)
  if (q == 0) {
    unsigned char* srow = gt_hi + (size_t)k * kStagePitch + h * 128;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float dwv[16];
      tc::tmem_ld16(tb + lane_base + tc::kColD + 128u + (uint32_t)(h * 32 + half * 16), dwv);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(srow + half * 64 + j * 16) =
            make_float4(dwv[4 * j + 0], dwv[4 * j + 1], dwv[4 * j + 2], dwv[4 * j + 3]);
    }
  }
  fence_async_smem();
  workers_sync();
  if (h == 0 && q == 0) {
    bulk_reduce_add_f32(D.g_vel_w[0] + k * NVFI_TM, gt_hi + (size_t)k * kStagePitch, 512u);
    bulk_commit();
    bulk_wait_read0();   // the caller's next phase may overwrite the staging rows
  }
  workers_sync();
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

// 17 warps are allocated as 20 (groups of 4 per scheduler): 96 registers per thread is the cap
// (a 112-register build fails to launch).
__global__ void __launch_bounds__(tc::kLaunchThreads, 1)
    k_advect_bwd_tc(const __grid_constant__ NvfiField F, const NvfiRenderArgs A, const NvfiRenderBuffers B,
                    const NvfiRenderGrads D, int S, long long total, int n_batches, int mode, int subs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* p = smem_raw;
  {
    const uint32_t a = tc::smem_u32(p);
    p += (1024u - (a & 1023u)) & 1023u;
  }
  unsigned char* ring = p;                                   // kBwdStages x 32 KB
  unsigned char* g_hi = ring + kBwdStages * tc::kStageBytes; // G^T hi: 64 KB, 1024-aligned
  unsigned char* g_lo = g_hi + 65536;                        // G^T lo
  tc::Ctl& ctl = *reinterpret_cast<tc::Ctl*>(g_lo + 65536);
  BwdTile& T = *reinterpret_cast<BwdTile*>(reinterpret_cast<unsigned char*>(&ctl) + sizeof(tc::Ctl));
  float* ws = D.workspace + (size_t)blockIdx.x * WS_CTA_F;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int mode3 = (mode == NVFI_MLP_TF32X3) ? 1 : 0;
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

  tc::setup(ctl, F.vel_net, nullptr);
  if (tid == 0) {   // weight segments in the order one tile consumes them
    int n = 0;
    for (int k = 0; k + 1 < n_steps; ++k) {
      ctl.prog[n++] = tc::SEG_FWD0;
      ctl.prog[n++] = tc::SEG_FWD0;
    }
    for (int k = 0; k < n_steps; ++k) {   // F1, F2 (both stashed), B2, B1
      ctl.prog[n++] = tc::SEG_FWD0;
      ctl.prog[n++] = tc::SEG_FWD0;
      ctl.prog[n++] = tc::SEG_BWD0;
      ctl.prog[n++] = tc::SEG_BWD0;
    }
    ctl.prog_len = (uint32_t)n;
  }
  for (int i = tid; i < 6 * NVFI_TM; i += blockDim.x) {
    const int n = i / NVFI_TM, kk = i - n * NVFI_TM;
    T.w5s[n][kk] = __ldg(F.vel_net[5].wt + (size_t)kk * F.vel_net[5].n_pad + n);
  }
  __syncthreads();
  tc::Issuer& is = T.iss;
  tc::Issuer& is_fwd = T.iss_fwd;
  if (warp == tc::kIssuerWarp) {
    is.init(ctl, tc::smem_u32(ring), kBwdStages, 0u, tc::RING_BWD);
    is_fwd.init(ctl, tc::smem_u32(g_hi), 4u, kBwdStages, tc::RING_FWD);
  }
  __syncthreads();
  uint32_t dphase = 0, kphase = 0;

  int sub = subs;
  long long batch_base = 0;
  bool exhausted = false;
  int qc = 0, par = 0;
  unsigned long long n_done = 0;
  float* stash = ws + TW_STASH;
  float* xsteps = ws + TW_XSTEPS;
  // per-thread partial sums of the head-layer weight gradient (unit k = tid & 127) and of the
  // bias gradients, carried in registers across all tiles of this CTA
  float acc_head[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, acc_bias[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

  // One tile = a short program of evaluations.  There is exactly ONE call site of the forward
  // tile evaluation and one of the backward tile evaluation (the kernel would not fit the
  // instruction cache otherwise); the small per-sample glue around them is selected by `kind`.
  enum { K_FWD_A = 0, K_FWD_B, K_REV_A, K_REV_B, K_BWD2, K_BWD1 };
  const int n_ops = 2 * (n_steps - 1) + 4 * n_steps;

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
      if (tid < NT && idx < total && B.valid[idx]) {
        const float g0 = D.g_x_adv[idx * 3], g1 = D.g_x_adv[idx * 3 + 1], g2 = D.g_x_adv[idx * 3 + 2];
        push = (g0 != 0.f) | (g1 != 0.f) | (g2 != 0.f);
      }
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
    TL(0);
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

#pragma unroll 1
    for (int op = 0; op < n_ops; ++op) {
      int k, kind;
      if (op < 2 * (n_steps - 1)) {   // forward sweep over all but the last step
        k = op >> 1;
        kind = K_FWD_A + (op & 1);
      } else {                        // reverse sweep
        const int r = op - 2 * (n_steps - 1);
        k = n_steps - 1 - (r >> 2);
        kind = K_REV_A + (r & 3);
      }
      const float dt = sched_dt[k], tcur = sched_t[k], hdt = 0.5f * dt;
      const float tmid = __fsub_rn(tcur, hdt);
      // ---- glue before the evaluation
      if (tid < NVFI_TM) {
        if (kind == K_FWD_A) {
#pragma unroll
          for (int a = 0; a < 3; ++a) xsteps[(k * 3 + a) * NVFI_TM + tid] = T.x0[a][tid];
        }
        if (kind == K_REV_A && k < n_steps - 1) {
#pragma unroll
          for (int a = 0; a < 3; ++a) T.x0[a][tid] = xsteps[(k * 3 + a) * NVFI_TM + tid];
        }
        if (kind == K_FWD_A || kind == K_REV_A) T.tvec[tid] = tcur;
        if (kind == K_FWD_B || kind == K_REV_B) T.tvec[tid] = tmid;
      }
      __syncthreads();
      const bool at_mid = (kind == K_FWD_B || kind == K_REV_B || kind == K_BWD2);
      const float* xs = at_mid ? T.xm[0] : T.x0[0];
      const float* ys = at_mid ? T.xm[1] : T.x0[1];
      const float* zs = at_mid ? T.xm[2] : T.x0[2];
      if (kind == K_BWD2 || kind == K_BWD1) {
        bwd_eval_tc(ctl, is, T, g_hi, g_lo, ws, D, at_mid ? stash + kStashF : stash, xs, ys, zs,
                    at_mid ? tmid : tcur, dphase, mode3, acc_head, acc_bias);
      } else {
        TL(1);
        float* wout = (kind == K_FWD_A || kind == K_REV_A) ? &T.w0[0][0] : &T.w1[0][0];
        float* st = (kind == K_REV_A) ? stash : ((kind == K_REV_B) ? stash + kStashF : nullptr);
        tc::vel_net_tile_tc<ACT_SILU>(ctl, is_fwd, 0, wout, xs, ys, zs, T.tvec, dphase, kphase, mode3, st);
      }
      // ---- glue after the evaluation
      if (tid < NVFI_TM) {
        const int m = tid;
        if (kind == K_FWD_A || kind == K_REV_A) {   // midpoint m = x0 - dt/2 v0(x0)
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
        } else if (kind == K_BWD2) {   // adjoint of m = x0 - dt/2 v0(x0)
          float gmv[3];
#pragma unroll
          for (int a = 0; a < 3; ++a) gmv[a] = T.gm[a][m] + T.gout[a][m];
          float gx0[3] = {T.gbar[0][m] + gmv[0], T.gbar[1][m] + gmv[1], T.gbar[2][m] + gmv[2]};
          float gw[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (!T.gate0[m]) {
            const float w[6] = {T.w0[0][m], T.w0[1][m], T.w0[2][m], T.w0[3][m], T.w0[4][m], T.w0[5][m]};
            const float gv[3] = {-hdt * gmv[0], -hdt * gmv[1], -hdt * gmv[2]};
            float gxe[3];
            basis_bwd(w, T.x0[0][m], T.x0[1][m], T.x0[2][m], gv, gw, gxe);
            gx0[0] += gxe[0];
            gx0[1] += gxe[1];
            gx0[2] += gxe[2];
          }
          T.gbar[0][m] = gx0[0];
          T.gbar[1][m] = gx0[1];
          T.gbar[2][m] = gx0[2];
#pragma unroll
          for (int i = 0; i < 6; ++i) T.gout[i][m] = gw[i];
        } else if (kind == K_BWD1) {
#pragma unroll
          for (int a = 0; a < 3; ++a) T.gbar[a][m] += T.gout[a][m];
        }
      }
      __syncthreads();
    }
  }
  if (tid < NT) {
    const int kk = tid & 127;
    bulk_wait0();   // outstanding TMA reductions of this thread are complete
#pragma unroll
    for (int n2 = 0; n2 < 6; ++n2) red_add(D.g_vel_w[5] + kk * 8 + n2, acc_head[n2]);
#pragma unroll
    for (int l = 0; l < 5; ++l) red_add(D.g_vel_b[l] + kk, acc_bias[l]);
    if (warp < 6) {
      const float s5 = warp_sum(acc_bias[5]);
      if (lane == 0) red_add(D.g_vel_b[5] + warp, s5);
    }
  }
  tc::teardown(ctl, is, &is_fwd);
  if (tid == 0 && n_done)
    atomicAdd(reinterpret_cast<unsigned long long*>(B.counters + 10), n_done);
}

}  // namespace tcb
}  // namespace nvfi

using namespace nvfi;

static_assert(tcb::TW_TOTAL <= WS_CTA_F, "per-CTA workspace of the tensor-core backward exceeds WS_CTA_F");

extern "C" int nvfi_debug_timeline_h(long long* dev_buf, int cap);
extern "C" int nvfi_debug_timeline(long long* dev_buf, int cap) {
  const int zero = 0;
  const int rc = nvfi_debug_timeline_h(dev_buf, cap);   // both tensor-core backward kernels get the buffer
  if (rc != NVFI_OK) return rc;
  NVFI_CUDA_OK(cudaMemcpyToSymbol(tcb::g_tl_buf, &dev_buf, sizeof(dev_buf)));
  NVFI_CUDA_OK(cudaMemcpyToSymbol(tcb::g_tl_cap, &cap, sizeof(cap)));
  NVFI_CUDA_OK(cudaMemcpyToSymbol(tcb::g_tl_n, &zero, sizeof(zero)));
  return NVFI_OK;
}

extern "C" int nvfi_launch_advect_bwd_tc(const NvfiField* F, const NvfiRenderArgs* A,
                                         const NvfiRenderBuffers* B, const NvfiRenderGrads* D, int S,
                                         long long total, int sms, int mode, cudaStream_t st) {
  for (int l = 0; l < NVFI_VEL_LAYERS; ++l) {
    if (!F->vel_net[l].umma) return NVFI_EINVAL;
    if (l < NVFI_VEL_LAYERS - 1 && (!F->vel_net[l].ummaT || F->vel_net[l].ummaT_rows != (l == 0 ? 32 : 128)))
      return NVFI_EINVAL;
  }
  if (F->vel_net[5].n_pad != 8) return NVFI_EUNSUPPORTED;
  const size_t smem = 1024 + (size_t)tcb::kBwdStages * tc::kStageBytes + 2 * 65536 + sizeof(tc::Ctl) +
                      sizeof(tcb::BwdTile);
  {
    const int rc = ensure_smem<tcb::k_advect_bwd_tc>(smem);
    if (rc != NVFI_OK) return rc;
  }
  const int subs = grab_subs(total, tcb::NT, sms);
  const int per_batch = subs * tcb::NT;
  const int n_batches = (int)((total + per_batch - 1) / per_batch);
  const int grid = n_batches < sms ? n_batches : sms;
  NVFI_LAUNCH(tcb::k_advect_bwd_tc, grid, tc::kLaunchThreads, smem, st, *F, *A, *B, *D, S, total, n_batches, mode, subs);
  return (int)cudaGetLastError();
}
