// Velocity-MLP tile evaluation on the 5th-generation tensor cores with FP16-SPLIT operands
// (tcgen05.mma kind::f16, sm_100a) — the product path of round 2.
//
// Every operand is split in two FP16 numbers, x = hi + lo (22 mantissa bits), and a GEMM is three MMAs
//      D = A_hi B_hi + A_lo B_hi + A_hi B_lo                (the dropped A_lo B_lo term is 2^-22 relative)
// with FP32 accumulation in tensor memory.  Against the 3xTF32 path of mlp_tc.cuh this halves the
// tensor-pipe cycles (K = 16 per instruction instead of 8) and, because FP16 operands may be read
// MN-major from shared memory, removes every transpose from the backward pass:
//
//   universal tile  [128 samples][128 units] FP16, 64 KB: two "hi" slabs (units 0-63, 64-127) then two
//   "lo" slabs; a slab is 128 rows of 128 bytes, 8-row groups of 1 KB, 16-byte chunks XOR-swizzled
//   with (row & 7) — the canonical SWIZZLE_128B layout.  The same bytes are
//     * the K-major A operand of a forward layer / input-gradient GEMM (M = sample, K = unit), and
//     * the MN-major A or B operand of a weight-gradient GEMM (M or N = unit, K = sample)
//   (tools/probe_h16.cu pins both descriptor forms on the hardware).
//
// One CTA owns a tile of 128 samples.  16 worker warps run the epilogues (warp w: TMEM lane quadrant
// w & 3 = 32 samples, columns 8 (w >> 2) + 32 g of every 32-column group g), a 17th warp issues the
// weight copies (cp.async.bulk into a ring of 32 KB K blocks) and all MMAs.  The epilogue of layer l
// writes the A operand of layer l + 1 IN PLACE into the tile (layer l's MMAs are complete by then),
// one 32-column group at a time; the issuer starts the MMAs of a group as soon as all 16 warps have
// stored it (kready[g]), into the other accumulator (D ping-pong).
//
// Weights are pre-packed (nvfi_pack_linear_h) into images of the same slab layout: per 64-wide K
// block the hi slab [rows][64] then the lo slab, so a K block is one contiguous bulk copy.
#pragma once

#include <cuda_fp16.h>

#include "mlp_tc.cuh"

// Development aid (tools/probe_timeline.py): a translation unit may define NVFI_TLH(tag, who) to record
// (tag, clock64) pairs at the phase boundaries; compiled out otherwise.
#ifndef NVFI_TLH
#define NVFI_TLH(tag, who)
#endif

namespace nvfi {
namespace th {

constexpr int kThreads = 512;                 // 16 worker warps
constexpr int kIssuerWarp = kThreads / 32;    // + 1 issuer warp
constexpr int kLaunchThreads = kThreads + 32;
constexpr uint32_t kTileBytes = 65536;
constexpr uint32_t kSlab = 16384;             // [128 rows][64 fp16]
constexpr uint32_t kLoOff = 32768;            // lo slabs follow the two hi slabs
constexpr uint32_t kStageBytes = 32768;       // one K block of a 128-row image: hi slab + lo slab
constexpr int kMaxStages = 4;
constexpr uint32_t kColD = 0;                 // accumulators D[0] / D[1] at TMEM columns 0 / 128
// TS form: the A operand (activations / upstream gradient) lives in tensor memory, two FP16 per 32-bit
// column (low half = even K index; tools/probe_h16.cu test 4): 64 columns of hi halves, 64 of lo halves.
// Shared memory then only feeds the B operand, which halves the MMAs' shared-memory reads.
constexpr uint32_t kColAhi = 384, kColAlo = 448;
enum : uint8_t { SEG_FWD0 = 0, SEG_FWD1 = 1, SEG_BWD0 = 2 };

// high word of a SWIZZLE_128B shared-memory descriptor: SBO = 1024 B, version 1, layout type 2
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
// no-swizzle (INTERLEAVE) descriptors of the small [128 samples][16] tile (core matrices of 8 rows x 16 B,
// element (m, n) at (m >> 3) * 256 + (n >> 3) * 128 + (m & 7) * 16 + (n & 7) * 2):
//   K-major A operand (M = sample, K = 16):  LBO = 128 (next 8 K elements), SBO = 256 (next 8 rows)
//   MN-major B operand (N = 8, K = samples): LBO = 256 (next 8 samples),    SBO = 128 (next 8 N)
constexpr uint32_t kDescHiSmallK = (256u >> 4) | (1u << 14);
constexpr uint32_t kDescHiSmallMN = (128u >> 4) | (1u << 14);

__host__ __device__ constexpr uint32_t idesc_f16(int n, int a_mn = 0, int b_mn = 0) {
  // FP32 accumulate, FP16 x FP16, M = 128, N = n
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((128u >> 4) << 24);
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes = 16u) {
  return ((smem_addr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                           uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  const uint64_t ad = ((uint64_t)a_hi << 32) | a_lo, bd = ((uint64_t)b_hi << 32) | b_lo;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(ad), "l"(bd), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
  const uint64_t bd = ((uint64_t)b_hi << 32) | b_lo;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bd), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint4& v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst),
               "r"(ssrc), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// two floats -> packed FP16 pair (element `a` at the lower address), saturating to the finite range
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// 8 floats -> 16 bytes of hi halves and 16 bytes of lo halves (v = hi + lo up to 2^-22 |v|)
__device__ __forceinline__ void split8(const float v[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = pack_h2(v[2 * j], v[2 * j + 1]);
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h[j]));
    l[j] = pack_h2(v[2 * j] - f.x, v[2 * j + 1] - f.y);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// ---------------------------------------------------------------- shared state
template <int NN>
struct CtlT {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t dbar;         // forward / input-gradient accumulator ready
  uint64_t wbar;         // weight-gradient accumulators ready (and their operand tiles free)
  uint64_t abar;         // activation tile of the stash has landed in shared memory
  uint64_t kready[4];    // 32-column group g of the next layer's A operand written (16 warp arrivals)
  uint64_t dbar2;        // the same pair for the second tile of the two-tile forward evaluation
  uint64_t kready2[4];
  uint32_t tmem_base;
  uint32_t prog_len;
  const unsigned char* img[NN][NVFI_VEL_LAYERS];  // forward images (W) of the nets
  const unsigned char* imgT[NVFI_VEL_LAYERS];     // images of W^T of net 0 (input-gradient GEMMs), or NULL
  uint8_t prog[208];     // weight segments in the order the kernel consumes them, cyclically
  alignas(16) float bias[NN][NVFI_VEL_LAYERS][NVFI_TM];
};
using Ctl = CtlT<2>;    // velocity + acceleration net (forward kernels)
using Ctl1 = CtlT<1>;   // velocity net only (backward kernel: shared memory is tight)

// Producer / consumer state of the weight ring (registers of the issuer warp, updated uniformly).
struct Issuer {
  uint32_t tb, ring_u32, n_stages;
  uint32_t p_stage, p_round, p_seg, p_li, p_kb;
  uint32_t c_stage, c_round, in_flight;
  template <class C>
  __device__ void init(const C& c, uint32_t ring_addr, uint32_t stages) {
    tb = tc::uniform(c.tmem_base);
    ring_u32 = tc::uniform(ring_addr);
    n_stages = stages;
    p_stage = p_round = p_seg = p_li = p_kb = 0;
    c_stage = c_round = in_flight = 0;
  }
};

// One-time setup by the whole CTA (any thread count >= 64): barriers, TMEM, biases.
template <int NN>
__device__ inline void setup(CtlT<NN>& c, const NvfiLinear* net0, const NvfiLinear* net1, uint32_t tmem_cols) {
  const int tid = threadIdx.x;
  if (NN < 2) net1 = nullptr;
  if (tid == 0) {
    for (int l = 0; l < NVFI_VEL_LAYERS; ++l) {
      c.img[0][l] = reinterpret_cast<const unsigned char*>(net0[l].himg);
      if (NN > 1) c.img[NN - 1][l] = net1 ? reinterpret_cast<const unsigned char*>(net1[l].himg) : nullptr;
      c.imgT[l] = reinterpret_cast<const unsigned char*>(net0[l].himgT);
    }
    c.prog[0] = SEG_FWD0;       // default program: forward evaluations, nets round-robin
    c.prog[1] = SEG_FWD1;
    c.prog_len = net1 ? 2 : 1;
    for (int s = 0; s < kMaxStages; ++s) {
      tc::mbar_init(&c.full[s], 1);
      tc::mbar_init(&c.empty[s], 1);
    }
    tc::mbar_init(&c.dbar, 1);
    tc::mbar_init(&c.wbar, 1);
    tc::mbar_init(&c.abar, 1);
    for (int k = 0; k < 4; ++k) tc::mbar_init(&c.kready[k], kThreads / 32);
    tc::mbar_init(&c.dbar2, 1);
    for (int k = 0; k < 4; ++k) tc::mbar_init(&c.kready2[k], kThreads / 32);
    tc::fence_barrier_init();
  }
  if (tid < 32) tc::tmem_alloc(&c.tmem_base, tmem_cols);
  for (int i = tid; i < NN * NVFI_VEL_LAYERS * NVFI_TM; i += blockDim.x) {
    const int w = i / (NVFI_VEL_LAYERS * NVFI_TM), r = i - w * (NVFI_VEL_LAYERS * NVFI_TM);
    const int l = r / NVFI_TM, n = r - l * NVFI_TM;
    const NvfiLinear* net = w ? net1 : net0;
    c.bias[w][l][n] = (net && net[l].bias && n < net[l].n_pad) ? net[l].bias[n] : 0.f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
}

// Geometry of the K blocks of a segment: position li in the segment's order -> image, rows, K blocks.
template <class C>
__device__ __forceinline__ void seg_block(const C& c, uint32_t seg, uint32_t li, const unsigned char*& img,
                                          uint32_t& rows, uint32_t& nkb) {
  if (seg == SEG_BWD0) {   // head^T (K = its 6 outputs, padded: 1 block), layers 4, 3, 2, 1, then 0 (32 rows)
    const uint32_t layer = 5u - li;
    img = c.imgT[layer];
    rows = (layer == 0) ? 32u : 128u;
    nkb = (li == 0) ? 1u : 2u;
  } else {                 // layers 0..5; the head has 16 rows, layer 0 one K block
    img = c.img[seg < (uint32_t)(sizeof(c.img) / sizeof(c.img[0])) ? seg : 0][li];
    rows = (li == NVFI_VEL_LAYERS - 1) ? 16u : 128u;
    nkb = (li == 0) ? 1u : 2u;
  }
}

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {   // one non-blocking probe
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(tc::smem_u32(bar)), "r"(parity)
      : "memory");
  return tc::uniform(done) != 0;   // one answer for the whole warp
}

// Issuer warp, all lanes, uniformly: keep the ring full.  With block == false a stage whose previous
// MMAs have not completed yet is left for the next call (the issuer has MMAs to issue meanwhile).
template <class C>
__device__ __forceinline__ void ring_top_up(C& c, Issuer& is, bool block = true) {
  while (is.in_flight < is.n_stages) {
    const uint32_t seg = c.prog[is.p_seg];
    if (is.p_round > 0) {
      if (block) tc::mbar_wait(&c.empty[is.p_stage], (is.p_round - 1) & 1);
      else if (!mbar_test(&c.empty[is.p_stage], (is.p_round - 1) & 1)) break;
    }
    const unsigned char* img;
    uint32_t rows, nkb;
    seg_block(c, seg, is.p_li, img, rows, nkb);
    const uint32_t bytes = rows * 256u;
    if (tc::elect_one()) {
      tc::mbar_expect_tx(&c.full[is.p_stage], bytes);
      tc::bulk_g2s_u32(is.ring_u32 + is.p_stage * kStageBytes, img + (size_t)is.p_kb * bytes, bytes,
                       &c.full[is.p_stage]);
    }
    __syncwarp();
    if (++is.p_stage == is.n_stages) {
      is.p_stage = 0;
      ++is.p_round;
    }
    if (++is.p_kb == nkb) {
      is.p_kb = 0;
      if (++is.p_li == (uint32_t)NVFI_VEL_LAYERS) {
        is.p_li = 0;
        if (++is.p_seg == c.prog_len) is.p_seg = 0;
      }
    }
    ++is.in_flight;
  }
}
// wait for the next K block of the ring; returns its shared-window address
template <class C>
__device__ __forceinline__ uint32_t ring_acquire(C& c, Issuer& is) {
  ring_top_up(c, is, is.in_flight == 0);   // block only when nothing is on its way
  tc::mbar_wait(&c.full[is.c_stage], is.c_round & 1);
  tc::tc_fence_after();
  return is.ring_u32 + is.c_stage * kStageBytes;
}
// wait for the K block `ahead` positions after the next one (0 = the next) WITHOUT consuming it; returns
// its shared-window address.  The blocks are consumed in order with ring_advance.
template <class C>
__device__ __forceinline__ uint32_t ring_wait(C& c, Issuer& is, uint32_t ahead) {
  ring_top_up(c, is, is.in_flight <= ahead);   // block only when that block is not even on its way
  uint32_t st = is.c_stage + ahead, rd = is.c_round;
  if (st >= is.n_stages) {
    st -= is.n_stages;
    ++rd;
  }
  tc::mbar_wait(&c.full[st], rd & 1);
  tc::tc_fence_after();
  return is.ring_u32 + st * kStageBytes;
}
// after the elected thread has committed empty[c_stage]
__device__ __forceinline__ void ring_advance(Issuer& is) {
  if (++is.c_stage == is.n_stages) {
    is.c_stage = 0;
    ++is.c_round;
  }
  --is.in_flight;
}
template <class C>
__device__ inline void drain(C& c, Issuer& is_ref) {
  Issuer is = is_ref;   // registers: `is_ref` may live in shared memory, where 32 lanes updating it in lock
                        // step is a (same-value) write-write hazard for compute-sanitizer racecheck
  while (is.in_flight > 0) {
    tc::mbar_wait(&c.full[is.c_stage], is.c_round & 1);
    ring_advance(is);
  }
}
template <class C>
__device__ inline void teardown(C& c, Issuer& is, uint32_t tmem_cols) {
  const int warp = threadIdx.x >> 5;
  tc::tc_fence_before();
  __syncthreads();
  if (warp == kIssuerWarp) drain(c, is);
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc(c.tmem_base, tmem_cols);
  }
}

// Issuer warp: the MMAs of 32-column group g (2 K steps of 16) of a forward `layer` whose A operand is the
// tile at `tile_u32`, into accumulator D[layer & 1].  Even groups acquire a K block of the ring, odd
// groups (and the single group of layer 0) release it.  With wait_store the elected thread first waits
// until its outstanding bulk stores (1: all, 2: all but the newest) have read their shared-memory source,
// with commit_d the accumulator is published on dbar.
template <bool TS, class C>
__device__ __forceinline__ void issue_group(C& c, Issuer& is, uint32_t tile_u32, int layer, uint32_t g,
                                            uint32_t& stage_u32, bool commit_d, int wait_store) {
  const uint32_t n = (layer == NVFI_VEL_LAYERS - 1) ? 16u : 128u;
  const uint32_t idesc = idesc_f16((int)n);
  if ((g & 1u) == 0u) stage_u32 = ring_acquire(c, is);
  else ring_top_up(c, is, false);
  const bool release = (g & 1u) || layer == 0;
  const uint32_t kb = g >> 1;
  const uint32_t a_hi = desc_lo(tile_u32 + kb * kSlab), a_lo = desc_lo(tile_u32 + kLoOff + kb * kSlab);
  const uint32_t w_hi = desc_lo(stage_u32), w_lo = desc_lo(stage_u32 + n * 128u);
  const uint32_t d = is.tb + kColD + 128u * (uint32_t)(layer & 1);
  if (tc::elect_one()) {
#pragma unroll
    for (uint32_t ks = 0; ks < 2; ++ks) {
      const uint32_t o = ((g & 1u) * 2u + ks) * 2u;   // 32 bytes per K step, in 16-byte units
      if (TS) {
        const uint32_t ta = is.tb + (2u * g + ks) * 8u;   // 8 columns per K step
        mma_f16_ts(d, ta + kColAhi, w_hi + o, kDescHiSw128, idesc, (g | ks) ? 1u : 0u);
        mma_f16_ts(d, ta + kColAlo, w_hi + o, kDescHiSw128, idesc, 1u);
        mma_f16_ts(d, ta + kColAhi, w_lo + o, kDescHiSw128, idesc, 1u);
      } else {
        mma_f16_ss(d, a_hi + o, kDescHiSw128, w_hi + o, kDescHiSw128, idesc, (g | ks) ? 1u : 0u);
        mma_f16_ss(d, a_lo + o, kDescHiSw128, w_hi + o, kDescHiSw128, idesc, 1u);
        mma_f16_ss(d, a_hi + o, kDescHiSw128, w_lo + o, kDescHiSw128, idesc, 1u);
      }
    }
    if (release) tc::tc_commit(&c.empty[is.c_stage]);
    if (wait_store == 1) bulk_wait_read0();        // all bulk stores have read their source
    else if (wait_store == 2) bulk_wait_read1();   // all but the most recent one
    if (commit_d) tc::tc_commit(&c.dbar);
  }
  __syncwarp();
  if (release) ring_advance(is);
}

// activation and (optionally) its derivative with the approximate SFU ops (ex2 2 ulp, rcp 1 ulp)
template <int ACT>
__device__ __forceinline__ float act_h(float x, float& da) {
  if (ACT == ACT_SILU) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    const float a = x * r;
    da = fmaf(a, 1.f - r, r);   // s + x s (1 - s)
    return a;
  }
  da = x > 0.f ? 1.f : 0.f;
  return fmaxf(x, 0.f);
}

// Stash of one stashed evaluation (global memory, per CTA), for the backward pass:
//   stash_a: [enc: hi slab 0 (16 KB) | lo slab 0 (16 KB)] [A_0 .. A_3: 4 x 64 KB tile images]
//   stash_s: S_l[unit / 4][sample][unit % 4] = act'(h_l), l = 0..4 (FP32; a thread owns 16-byte quads of units,
//            a warp stores / loads 512 contiguous bytes per instruction)
constexpr size_t kStashABytes = 32768 + 4 * (size_t)kTileBytes;
constexpr size_t kStashSFloats = 5 * (size_t)NVFI_TM * NVFI_TM;

// Weight net of VelBasis on a tile (models/velocity_field.py:58-67, models/base_network.py:42-54):
// inputs (x,y,z,t)[m] in shared memory, outputs outS[0..5][m].  `tile_u32`: shared-window address of the
// 64 KB activation tile (1024-aligned); the epilogue of layer l writes A_l over A_{l-1} in place and on
// return the tile holds A_4 = act(h_4).  With a second tile (tile1_u32 != 0: the stashed evaluations of the
// backward kernel) the layers ping-pong instead — the encoding and A_1, A_3 in tile 0; A_0, A_2, A_4 in
// tile 1 — so that the bulk copy of A_l to the stash has a whole layer to read its tile before the tile
// is overwritten.  Whole CTA (2 block barriers).
//
// JVP = 1 (the PDE loss, k_pde_jac_h): forward-mode rows.  The 32 rows of a TMEM lane quadrant hold 6
// points x 5 rows — lane 5 p + j: j = 0 the value row, j = 1..4 the tangent d/d(x, y, z, t) — and 2 dead
// lanes.  The linear layers are unchanged (tangent rows simply carry no bias); the activation couples the
// rows of a point, a_j = silu'(h_0) h_j, through one warp shuffle per element.  The stash then holds
// S = silu'(h_0) for all 5 rows and S2 = silu''(h_0) h_j for the tangent rows (stash_s2), which is all the
// reverse pass needs: g_h0 = g_a0 S + sum_j g_aj S2_j, g_hj = g_aj S.
template <int ACT, bool TS, int JVP = 0, class C>
__device__ void vel_net_tile_h(C& c, Issuer& is_ref, int which, float* outS, const float* xs,
                               const float* ys, const float* zs, const float* ts, uint32_t tile_u32,
                               uint32_t& dphase, uint32_t& kphase, uint32_t tile1_u32 = 0u,
                               unsigned char* __restrict__ stash_a = nullptr,
                               float* __restrict__ stash_s = nullptr, float* __restrict__ stash_s2 = nullptr) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool stash = stash_a != nullptr;
  if (tile1_u32 == 0u) tile1_u32 = tile_u32;
#ifdef NVFI_TLH_AT
  __shared__ volatile long long tl_arr[16];
  __shared__ volatile long long tl_iss[10];
  __shared__ volatile long long tl_own[4];
#endif
  // layer L reads tile (L & 1), its epilogue writes tile ((L + 1) & 1)
  if (warp == kIssuerWarp) {      // ---- issuer warp: all 32 lanes run the issue code uniformly
    Issuer is = is_ref;
    NVFI_TLH(1010, 1);
    ring_top_up(c, is);           // weights stream in while the workers encode
    __syncthreads();              // (1) the encoding (group 0 of layer 0's A operand) is in the tile
    tc::tc_fence_after();
    NVFI_TLH(1011, 1);
    uint32_t stage = 0;
    if (stash && tc::elect_one()) {
      bulk_s2g(stash_a, tile_u32, kSlab);
      bulk_s2g(stash_a + kSlab, tile_u32 + kLoOff, kSlab);
      bulk_commit();
    }
    __syncwarp();
    issue_group<TS>(c, is, tile_u32, 0, 0, stage, true, (stash && !TS && tile1_u32 == tile_u32) ? 1 : 0);
#pragma unroll 1
    for (int l = 0; l < NVFI_VEL_LAYERS - 1; ++l) {
#pragma unroll 1
      for (uint32_t g = 0; g < 4; ++g) {
        ring_top_up(c, is, false);
        tc::mbar_wait(&c.kready[g], kphase & 1);
        tc::tc_fence_after();
        NVFI_TLH(1200 + 4 * l + (int)g, 1);
#ifdef NVFI_TLH_AT
        tl_iss[2 + 2 * g] = clock64();   // woken on kready[g]
#endif
        const bool st = stash && l < 4 && g == 3;
        if (st && tc::elect_one()) {   // A_l is complete: copy the tile image to the stash
          bulk_s2g(stash_a + 32768 + (size_t)l * kTileBytes, ((l + 1) & 1) ? tile1_u32 : tile_u32, kTileBytes);
          bulk_commit();
        }
        __syncwarp();
        // before layer l + 1's accumulator is published its epilogue's target tile must be free: in place
        // that is the tile just copied (wait for it), with two tiles it is the copy issued a layer ago
        issue_group<TS>(c, is, ((l + 1) & 1) ? tile1_u32 : tile_u32, l + 1, g, stage, g == 3,
                    (stash && g == 3) ? (tile1_u32 == tile_u32 ? 1 : 2) : 0);
        NVFI_TLH(1020 + 4 * l + (int)g, 1);
#ifdef NVFI_TLH_AT
        tl_iss[3 + 2 * g] = clock64();   // group g issued
        if (g == 3) {   // timeline builds: when does the tensor pipe publish layer l + 1's accumulator?
          tl_iss[0] = clock64();   // commit issued
          tc::mbar_wait(&c.dbar, (dphase + (uint32_t)l + 1u) & 1);
          tl_iss[1] = clock64();   // accumulator published, as seen by the issuer
          NVFI_TLH(1500 + l, 1);
        }
#endif
      }
      ++kphase;
    }
    dphase += NVFI_VEL_LAYERS;
    is_ref = is;
    __syncthreads();              // (2) outS is complete
    return;
  }
  const int q = warp & 3, h = warp >> 2;       // TMEM lane quadrant, 8-column slot in a 32-column group
  const int m = q * 32 + lane;                 // sample (= TMEM lane = tile row) of this thread
  const uint32_t tb = c.tmem_base;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const uint32_t row_off = (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
  const uint32_t row_u32 = tile_u32 + row_off;
  const uint32_t x7 = (uint32_t)(m & 7);
  // JVP: row type of this lane (0 value, 1..4 tangent, 5 dead) and the lane of its point's value row
  const int jt = JVP ? (lane < 30 ? lane % 5 : 5) : 0;
  const int jsrc = JVP ? (lane < 30 ? lane - jt : lane) : lane;

  // ---- PositionEncoder(3) of (x, y, z, t): 28 values + 4 zeros into columns [0, 32):
  // [q | sin q | cos q | sin 2q | cos 2q | sin 4q | cos 4q | 0], 8 columns per warp slot
  {
    const float p[4] = {xs[m], ys[m], zs[m], ts[m]};
    float v[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // slot h: [q, sin q], [cos q, sin 2q], [cos 2q, sin 4q], [cos 4q, 0]
      const float fa = (h <= 1) ? 1.f : ((h == 2) ? 2.f : 4.f), fb = (h == 0) ? 1.f : ((h == 1) ? 2.f : 4.f);
      float sa, ca, sb, cb;
      tc::sincos_bounded(p[i] * fa, sa, ca);
      tc::sincos_bounded(p[i] * fb, sb, cb);
      v[i] = (h == 0) ? p[i] : ca;
      v[4 + i] = (h == 3) ? 0.f : sb;
      if (JVP && jt != 0) {   // tangent row: d/dq_i of the same entries for i = jt - 1, zero elsewhere
        const bool mine = (i == jt - 1);
        v[i] = mine ? ((h == 0) ? 1.f : -fa * sa) : 0.f;
        v[4 + i] = (mine && h != 3) ? fb * cb : 0.f;
      }
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    if (TS) {
      tmem_st4(tb + lane_base + kColAhi + (uint32_t)(4 * h), hi);
      tmem_st4(tb + lane_base + kColAlo + (uint32_t)(4 * h), lo);
    }
    if (!TS || stash) {
      const uint32_t a = row_u32 + (((uint32_t)h ^ x7) << 4);
      st_shared_v4(a, hi);
      st_shared_v4(a + kLoOff, lo);
      fence_async_smem();
    }
    if (TS) tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  NVFI_TLH(10, 0);
  __syncthreads();   // (1)
  NVFI_TLH(11, 0);

#pragma unroll 1
  for (int l = 0; l < NVFI_VEL_LAYERS - 1; ++l) {
    // ---- accumulator of layer l
    tc::mbar_wait(&c.dbar, dphase & 1);
    ++dphase;
    tc::tc_fence_after();
#ifdef NVFI_TLH_AT
    const long long t_wake = clock64();
    if (l > 0 && tid == 0) {   // the latest of the 16 warps' arrivals that released layer l's last MMAs
      long long mx = 0;
      for (int w = 0; w < 16; ++w) mx = tl_arr[w] > mx ? tl_arr[w] : mx;
      NVFI_TLH_AT(200 + l, 0, mx);
      if (l == 2) {   // one layer in detail: arrivals of this thread's warp, wake-ups and issues of the issuer
        for (int g = 0; g < 4; ++g) {
          NVFI_TLH_AT(500 + g, 0, tl_own[g]);
          NVFI_TLH_AT(510 + g, 0, tl_iss[2 + 2 * g]);
          NVFI_TLH_AT(520 + g, 0, tl_iss[3 + 2 * g]);
        }
      }
      NVFI_TLH_AT(300 + l, 0, tl_iss[0]);
      NVFI_TLH_AT(400 + l, 0, tl_iss[1]);
    }
    NVFI_TLH_AT(20 + l, 0, t_wake);
#else
    NVFI_TLH(20 + l, 0);
#endif
    const uint32_t dcol = tb + lane_base + kColD + 128u * (uint32_t)(l & 1) + (uint32_t)(h * 8);
    uint32_t raw[4][8];
#pragma unroll
    for (int g = 0; g < 4; ++g) tc::tmem_ld8_nowait(dcol + 32u * g, raw[g]);
    tc::tmem_ld_wait();
    // ---- epilogue, one 32-column group of layer l + 1's A operand at a time.  (The loop is unrolled and ptxas
    // hoists the activations of all four groups in front of the first group's stores, so the 16 warps arrive
    // on kready[0..3] late in the epilogue; a loop that is NOT unrolled keeps the groups staged — arrivals
    // ~1 K cycles apart — but loses the instruction-level parallelism: measured time-neutral, as was handing
    // the operand over in two halves instead of four groups.)
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int col = g * 32 + h * 8;
      const float4 b0 = *reinterpret_cast<const float4*>(&c.bias[which][l][col]);
      const float4 b1 = *reinterpret_cast<const float4*>(&c.bias[which][l][col + 4]);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float av[8], sv[8];
      if (JVP) {
        float s2v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float own = __uint_as_float(raw[g][i]);
          const float hv = __shfl_sync(0xffffffffu, own, jsrc) + bb[i];   // pre-activation of the value row
          float e, r;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(hv * -1.4426950408889634f));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
          const float a = hv * r;
          const float d1 = fmaf(a, 1.f - r, r);                                   // silu'
          const float d2 = r * (1.f - r) * fmaf(hv, 1.f - 2.f * r, 2.f);          // silu''
          av[i] = (jt == 0) ? a : ((jt < 5) ? d1 * own : 0.f);
          sv[i] = (jt < 5) ? d1 : 0.f;
          s2v[i] = (jt >= 1 && jt < 5) ? d2 * own : 0.f;
        }
        if (stash) {
          float4* sp2 = reinterpret_cast<float4*>(stash_s2) + ((size_t)l * 32 + (col >> 2)) * NVFI_TM + m;
          __stcg(sp2, make_float4(s2v[0], s2v[1], s2v[2], s2v[3]));
          __stcg(sp2 + NVFI_TM, make_float4(s2v[4], s2v[5], s2v[6], s2v[7]));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = act_h<ACT>(__uint_as_float(raw[g][i]) + bb[i], sv[i]);
      }
      if (stash) {   // quads of units: a warp stores 512 contiguous bytes per quad
        float4* sp = reinterpret_cast<float4*>(stash_s) + ((size_t)l * 32 + (col >> 2)) * NVFI_TM + m;
        __stcg(sp, make_float4(sv[0], sv[1], sv[2], sv[3]));
        __stcg(sp + NVFI_TM, make_float4(sv[4], sv[5], sv[6], sv[7]));
      }
      uint4 hi, lo;
      split8(av, hi, lo);
      if (TS) {   // next layer's A operand: tensor memory
        tmem_st4(tb + lane_base + kColAhi + (uint32_t)(16 * g + 4 * h), hi);
        tmem_st4(tb + lane_base + kColAlo + (uint32_t)(16 * g + 4 * h), lo);
      }
      if (!TS || stash) {   // shared-memory tile: the A operand (SS form) / the image copied to the stash
        const uint32_t a = (((l + 1) & 1) ? tile1_u32 : tile_u32) + row_off + (uint32_t)(g >> 1) * kSlab +
                           ((((uint32_t)(4 * g + h) & 7u) ^ x7) << 4);
        st_shared_v4(a, hi);
        st_shared_v4(a + kLoOff, lo);
        fence_async_smem();
      }
      if (TS) tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&c.kready[g]);
#ifdef NVFI_TLH_AT
      if (g == 3 && lane == 0) tl_arr[warp] = clock64();
      if (tid == 0) tl_own[g] = clock64();
#endif
    }
    ++kphase;
    NVFI_TLH(30 + l, 0);
  }
  // ---- head: 6 basis weights
  tc::mbar_wait(&c.dbar, dphase & 1);
  ++dphase;
  tc::tc_fence_after();
  NVFI_TLH(40, 0);
  if (h == 0) {
    uint32_t raw[8];
    tc::tmem_ld8_nowait(tb + lane_base + kColD + 128u * (uint32_t)((NVFI_VEL_LAYERS - 1) & 1), raw);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 6; ++i)
      outS[i * NVFI_TM + m] = __uint_as_float(raw[i]) + ((JVP && jt != 0) ? 0.f : c.bias[which][NVFI_VEL_LAYERS - 1][i]);
  }
  tc::tc_fence_before();
  __syncthreads();   // (2)
}

// ---------------------------------------------------------------------------------------------------
// Two tiles in flight (the forward kernels).  With one tile the 16 worker warps wait while the last
// MMAs of a layer drain, and the tensor pipe waits while they run the epilogue.  Here the workers
// alternate between two tiles of 128 samples: while they run the epilogue of tile 0's layer l, the
// tensor cores execute tile 1's layer l MMAs (and vice versa), and both tiles consume the SAME weight
// block from the ring, which halves the weight traffic per sample.  TS form only:
//   TMEM columns of tile t:  D [256 t, +128)   A_hi [256 t + 128, +64)   A_lo [256 t + 192, +64)
// D is single-buffered: a warp arrives on kready[t][0] only after it has loaded ALL its accumulator
// columns, and the issuer starts layer l + 1 of that tile only then.
// xyzt[t][c][m]: inputs of tile t (c = x, y, z, t), outS[t]: its 6 basis weights [6][128].
template <int ACT, class C>
__device__ void vel_net_tile2_h(C& c, Issuer& is_ref, int which, float* const (&outS)[2],
                                const float* const (&xyzt)[2][4], uint32_t (&dph)[2], uint32_t (&kph)[2]) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == kIssuerWarp) {      // ---- issuer warp
    Issuer is = is_ref;
    ring_top_up(c, is, false);
    __syncthreads();              // (1) both encodings are in tensor memory
    tc::tc_fence_after();
    // one 32-column group (2 K steps) of `layer` for tile t against the K block at `stage`
    auto issue = [&](uint32_t t, int layer, uint32_t g, uint32_t stage, uint64_t* dbar) {
      const uint32_t n = (layer == NVFI_VEL_LAYERS - 1) ? 16u : 128u;
      const uint32_t idesc = idesc_f16((int)n);
      const uint32_t w_hi = desc_lo(stage), w_lo = desc_lo(stage + n * 128u);
      const uint32_t d = is.tb + 256u * t;
      if (tc::elect_one()) {
#pragma unroll
        for (uint32_t ks = 0; ks < 2; ++ks) {
          const uint32_t o = ((g & 1u) * 2u + ks) * 2u;
          const uint32_t ta = d + (2u * g + ks) * 8u;
          mma_f16_ts(d, ta + 128u, w_hi + o, kDescHiSw128, idesc, (g | ks) ? 1u : 0u);
          mma_f16_ts(d, ta + 192u, w_hi + o, kDescHiSw128, idesc, 1u);
          mma_f16_ts(d, ta + 128u, w_lo + o, kDescHiSw128, idesc, 1u);
        }
        if (dbar) tc::tc_commit(dbar);
      }
      __syncwarp();
    };
    auto release = [&]() {        // the MMAs issued so far were the last readers of the oldest block
      if (tc::elect_one()) tc::tc_commit(&c.empty[is.c_stage]);
      __syncwarp();
      ring_advance(is);
    };
    {
      const uint32_t st = ring_wait(c, is, 0);
      issue(0, 0, 0, st, &c.dbar);
      issue(1, 0, 0, st, &c.dbar2);
      release();
    }
#pragma unroll 1
    for (int l = 0; l < NVFI_VEL_LAYERS - 1; ++l) {
      uint32_t st[2];
      st[0] = ring_wait(c, is, 0);
#pragma unroll 1
      for (uint32_t t = 0; t < 2; ++t) {
#pragma unroll 1
        for (uint32_t g = 0; g < 4; ++g) {
          if (t == 0 && g == 2) st[1] = ring_wait(c, is, 1);
          else ring_top_up(c, is, false);
          tc::mbar_wait(t ? &c.kready2[g] : &c.kready[g], kph[t] & 1);
          tc::tc_fence_after();
          issue(t, l + 1, g, (g < 2) ? st[0] : st[1], g == 3 ? (t ? &c.dbar2 : &c.dbar) : nullptr);
          if (t == 1 && (g & 1u)) release();   // g = 1: K block 0, g = 3: K block 1
        }
        ++kph[t];
      }
    }
    dph[0] += NVFI_VEL_LAYERS;
    dph[1] += NVFI_VEL_LAYERS;
    is_ref = is;
    __syncthreads();              // (2) outputs complete
    return;
  }
  const int q = warp & 3, h = warp >> 2;
  const int m = q * 32 + lane;
  const uint32_t tb = c.tmem_base;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  // ---- PositionEncoder(3) of both tiles (see vel_net_tile_h)
#pragma unroll 1
  for (int t = 0; t < 2; ++t) {
    const float p[4] = {xyzt[t][0][m], xyzt[t][1][m], xyzt[t][2][m], xyzt[t][3][m]};
    float v[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float fa = (h <= 1) ? 1.f : ((h == 2) ? 2.f : 4.f), fb = (h == 0) ? 1.f : ((h == 1) ? 2.f : 4.f);
      float sa, ca, sb, cb;
      tc::sincos_bounded(p[i] * fa, sa, ca);
      tc::sincos_bounded(p[i] * fb, sb, cb);
      v[i] = (h == 0) ? p[i] : ca;
      v[4 + i] = (h == 3) ? 0.f : sb;
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    tmem_st4(tb + lane_base + 256u * t + 128u + (uint32_t)(4 * h), hi);
    tmem_st4(tb + lane_base + 256u * t + 192u + (uint32_t)(4 * h), lo);
  }
  tc::tmem_st_wait();
  tc::tc_fence_before();
  __syncthreads();   // (1)

#pragma unroll 1
  for (int l = 0; l < NVFI_VEL_LAYERS - 1; ++l) {
#pragma unroll 1
    for (int t = 0; t < 2; ++t) {
      tc::mbar_wait(t ? &c.dbar2 : &c.dbar, dph[t] & 1);
      ++dph[t];
      tc::tc_fence_after();
      const uint32_t tcol = tb + lane_base + 256u * (uint32_t)t;
      uint32_t raw[4][8];
#pragma unroll
      for (int g = 0; g < 4; ++g) tc::tmem_ld8_nowait(tcol + (uint32_t)(h * 8) + 32u * g, raw[g]);
      tc::tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = g * 32 + h * 8;
        const float4 b0 = *reinterpret_cast<const float4*>(&c.bias[which][l][col]);
        const float4 b1 = *reinterpret_cast<const float4*>(&c.bias[which][l][col + 4]);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float av[8], sv;
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = act_h<ACT>(__uint_as_float(raw[g][i]) + bb[i], sv);
        uint4 hi, lo;
        split8(av, hi, lo);
        tmem_st4(tcol + 128u + (uint32_t)(16 * g + 4 * h), hi);
        tmem_st4(tcol + 192u + (uint32_t)(16 * g + 4 * h), lo);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(t ? &c.kready2[g] : &c.kready[g]);
      }
      ++kph[t];
    }
  }
  // ---- heads
#pragma unroll 1
  for (int t = 0; t < 2; ++t) {
    tc::mbar_wait(t ? &c.dbar2 : &c.dbar, dph[t] & 1);
    ++dph[t];
    tc::tc_fence_after();
    if (h == 0) {
      uint32_t raw[8];
      tc::tmem_ld8_nowait(tb + lane_base + 256u * (uint32_t)t, raw);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 6; ++i)
        outS[t][i * NVFI_TM + m] = __uint_as_float(raw[i]) + c.bias[which][NVFI_VEL_LAYERS - 1][i];
    }
  }
  tc::tc_fence_before();
  __syncthreads();   // (2)
}

}  // namespace th
}  // namespace nvfi
