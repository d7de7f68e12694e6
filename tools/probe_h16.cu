// Hardware probe for the FP16-split tensor-core path (round 2).  Stand-alone binary:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I nvfi_b200/csrc -I include tools/probe_h16.cu -o gpurun_out/probe_h16
// It answers, on the B200, the questions the kernels in mlp_h.cuh / backward_h.cu rest on:
//   1. kind::f16 MMA, both operands from shared memory, K-major SWIZZLE_128B tiles [row][64 fp16]
//   2. the SAME bytes read MN-major (a_major = b_major = 1): D[i][j] = sum_m A[m][i] B[m][j]
//   3. N = 8 MMAs against a constant "ones" block (bias gradients) and small no-swizzle tiles
//   4. FP16 A operand packed two per column in tensor memory (TS form)
//   5. throughput of cp.reduce.async.bulk (.add.f32) into one shared 320 KB gradient buffer
// Every MMA test is generic: the host builds byte images of the operands for one hypothesis about
// the layout, the kernel copies them verbatim to shared memory / tensor memory and issues the MMAs
// with host-supplied descriptor words; the host compares with a reference GEMM.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "mlp_tc.cuh"

using namespace nvfi;

struct MmaJob {
  uint32_t a_bytes, b_bytes;       // image sizes (A image may be 0 when A comes from TMEM)
  uint32_t a_tmem_cols;            // > 0: A from tensor memory, image = [128 lanes][cols] u32
  uint32_t n_mma;
  uint32_t a_lo0, a_step;          // descriptor low word base offset (bytes from A image start) / step per MMA
  uint32_t b_lo0, b_step;
  uint32_t a_lbo, b_lbo;           // LBO fields (>> 4 units)
  uint32_t a_hi, b_hi;             // descriptor high words
  uint32_t idesc;
  uint32_t d_cols;                 // accumulator columns to read back
  uint32_t a_tmem_step;            // TMEM column step per MMA (TS form)
};

__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
    k_probe(const unsigned char* __restrict__ a_img, const unsigned char* __restrict__ b_img,
            float* __restrict__ Dout, MmaJob J) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* p = smem_raw;
  {
    const uint32_t a = tc::smem_u32(p);
    p += (1024u - (a & 1023u)) & 1023u;
  }
  unsigned char* sa = p;
  unsigned char* sb = p + ((J.a_bytes + 1023u) & ~1023u);
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  if (J.a_tmem_cols == 0)
    for (uint32_t i = tid * 16; i < J.a_bytes; i += 128 * 16)
      *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(a_img + i);
  for (uint32_t i = tid * 16; i < J.b_bytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(b_img + i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tb = tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  if (J.a_tmem_cols) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a_img) + (size_t)tid * J.a_tmem_cols;
    for (uint32_t c = 0; c < J.a_tmem_cols; c += 8) {
      uint32_t v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = src[c + i];
      tc::tmem_st8(tb + lane_base + 256u + c, v);
    }
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc::tc_fence_after();
    const uint32_t ab = tc::smem_u32(sa), bb = tc::smem_u32(sb);
    for (uint32_t i = 0; i < J.n_mma; ++i) {
      const uint32_t blo = (((bb + J.b_lo0 + i * J.b_step) >> 4) & 0x3FFFu) | (J.b_lbo << 16);
      const uint64_t bd = ((uint64_t)J.b_hi << 32) | blo;
      if (J.a_tmem_cols) {
        mma_f16_ts(tb, tb + 256u + i * J.a_tmem_step, bd, J.idesc, i ? 1u : 0u);
      } else {
        const uint32_t alo = (((ab + J.a_lo0 + i * J.a_step) >> 4) & 0x3FFFu) | (J.a_lbo << 16);
        const uint64_t ad = ((uint64_t)J.a_hi << 32) | alo;
        mma_f16_ss(tb, ad, bd, J.idesc, i ? 1u : 0u);
      }
    }
    tc::tc_commit(&bar);
  }
  __syncwarp();
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  for (uint32_t c = 0; c < J.d_cols; c += 8) {
    uint32_t v[8];
    tc::tmem_ld8_nowait(tb + lane_base + c, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) Dout[(size_t)tid * J.d_cols + c + i] = __uint_as_float(v[i]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tb, 512);
  }
}

// ---------------------------------------------------------------- host helpers
static uint16_t h16(float x) {
  __half h = __float2half_rn(x);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
static float f16(uint16_t u) {
  __half h;
  memcpy(&h, &u, 2);
  return __half2float(h);
}
// universal tile: [128 rows][ncols fp16], slabs of 64 columns (16 KB each), rows of 128 B, 8-row groups of
// 1 KB, 16-byte chunks XOR-swizzled with (row & 7)
static void tile_put(std::vector<unsigned char>& img, int row, int col, uint16_t v) {
  const size_t off = (size_t)(col >> 6) * 16384 + (size_t)(row >> 3) * 1024 + (size_t)(row & 7) * 128 +
                     (size_t)((((col & 63) >> 3) ^ (row & 7)) << 4) + (size_t)(col & 7) * 2;
  memcpy(&img[off], &v, 2);
}
static uint32_t idesc_f16(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
static uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout) { return (sbo_bytes >> 4) | (1u << 14) | (layout << 29); }

static unsigned char *d_a, *d_b;
static float* d_out;
static double run(const MmaJob& J, const std::vector<unsigned char>& a, const std::vector<unsigned char>& b,
                  const std::vector<double>& ref, int rows, int cols, const char* name) {
  cudaMemcpy(d_a, a.data(), a.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(d_b, b.data(), b.size(), cudaMemcpyHostToDevice);
  cudaMemset(d_out, 0xff, 128 * 256 * 4);
  const size_t smem = 1024 + ((J.a_bytes + 1023u) & ~1023u) + J.b_bytes + 1024;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_probe<<<1, 128, smem>>>(d_a, d_b, d_out, J);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-44s CUDA error %s\n", name, cudaGetErrorString(e));
    exit(1);
  }
  std::vector<float> out((size_t)128 * J.d_cols);
  cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
  double num = 0, den = 0;
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      const double d = (double)out[(size_t)r * J.d_cols + c] - ref[(size_t)r * cols + c];
      num += d * d;
      den += ref[(size_t)r * cols + c] * ref[(size_t)r * cols + c];
    }
  const double err = sqrt(num / (den + 1e-30));
  printf("%-44s rel err %.3e  %s\n", name, err, err < 1e-3 ? "OK" : "MISMATCH");
  return err;
}

// ---------------------------------------------------------------- MMA pacing
// cycles per MMA for back-to-back kind::f16 MMAs, M = 128: SS (both operands from shared memory, K-major
// SW128) and TS (A from tensor memory), N = 128 / 256.  mode: 0 SS, 1 TS, 2 SS MN-major both
__global__ void __launch_bounds__(128, 1) k_mma_pace(long long* out, int mode, int n, int iters) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tb = tmem_slot;
  if (tid == 0) {
    const uint32_t sa = tc::smem_u32(sm), sb = sa + 32768;
    const int mn = mode == 2;
    const uint32_t idesc = (1u << 4) | ((uint32_t)mn << 15) | ((uint32_t)mn << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t k = (uint32_t)(it & 3);
      const uint32_t alo = mn ? ((((sa + k * 2048u) >> 4) & 0x3FFFu) | (1024u << 16)) : ((((sa + k * 32u) >> 4) & 0x3FFFu) | (1u << 16));
      const uint32_t blo = mn ? ((((sb + k * 2048u) >> 4) & 0x3FFFu) | (1024u << 16)) : ((((sb + k * 32u) >> 4) & 0x3FFFu) | (1u << 16));
      const uint64_t bd = ((uint64_t)hi << 32) | blo;
      if (mode == 1) mma_f16_ts(tb, tb + 256u + k * 8u, bd, idesc, 1u);
      else mma_f16_ss(tb, ((uint64_t)hi << 32) | alo, bd, idesc, 1u);
    }
    tc::tc_commit(&bar);
    tc::mbar_wait(&bar, 0);
    out[0] = clock64() - t0;
  }
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tb, 512);
  }
}

// ---------------------------------------------------------------- TMA reduce throughput
__global__ void __launch_bounds__(128, 1) k_reduce_bw(float* __restrict__ dst, int iters, int priv, int layers) {
  extern __shared__ __align__(1024) unsigned char sm[];
  float* s = reinterpret_cast<float*>(sm);
  for (int i = threadIdx.x; i < 128 * 128; i += 128) s[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  float* base = dst + (priv ? (size_t)blockIdx.x * layers * 16384 : 0);
  for (int it = 0; it < iters; ++it) {
    float* g = base + (size_t)(it % layers) * 16384 + threadIdx.x * 128;
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(g),
                 "r"(tc::smem_u32(s + threadIdx.x * 128)), "r"(512u)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  cudaMalloc(&d_a, 1 << 17);
  cudaMalloc(&d_b, 1 << 17);
  cudaMalloc(&d_out, 128 * 256 * 4);
  srand(7);
  std::vector<float> A(128 * 128), B(128 * 128);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  std::vector<uint16_t> Ah(128 * 128), Bh(128 * 128);
  for (int i = 0; i < 128 * 128; ++i) {
    Ah[i] = h16(A[i]);
    Bh[i] = h16(B[i]);
  }
  auto a_at = [&](int r, int c) { return (double)f16(Ah[r * 128 + c]); };
  auto b_at = [&](int r, int c) { return (double)f16(Bh[r * 128 + c]); };
  std::vector<unsigned char> ta(32768), tb_(32768);
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 128; ++c) {
      tile_put(ta, r, c, Ah[r * 128 + c]);
      tile_put(tb_, r, c, Bh[r * 128 + c]);
    }
  // ---- 1. K-major SS: D[m][n] = sum_u A[m][u] B[n][u]
  {
    std::vector<double> ref(128 * 128, 0.0);
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 128; ++n) {
        double s = 0;
        for (int u = 0; u < 128; ++u) s += a_at(m, u) * b_at(n, u);
        ref[m * 128 + n] = s;
      }
    // 8 MMAs of K = 16: slab (i >> 2) * 16 KB + (i & 3) * 32 B.  The generic job has a linear step, so run
    // the two slabs as two jobs?  Use step table instead: do 4 + 4 with accumulate via two calls is not
    // possible -> encode as n_mma = 4 on slab 0 only and compare with the K = 64 reference, then slab 1.
    for (int slab = 0; slab < 2; ++slab) {
      std::vector<double> r2(128 * 128, 0.0);
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
          double s = 0;
          for (int u = slab * 64; u < slab * 64 + 64; ++u) s += a_at(m, u) * b_at(n, u);
          r2[m * 128 + n] = s;
        }
      MmaJob J{};
      J.a_bytes = 32768; J.b_bytes = 32768; J.n_mma = 4;
      J.a_lo0 = slab * 16384; J.a_step = 32; J.b_lo0 = slab * 16384; J.b_step = 32;
      J.a_lbo = 1; J.b_lbo = 1; J.a_hi = desc_hi(1024, 2); J.b_hi = desc_hi(1024, 2);
      J.idesc = idesc_f16(128, 128, 0, 0); J.d_cols = 128;
      char nm[64];
      snprintf(nm, sizeof nm, "1. SS K-major SW128 f16, slab %d", slab);
      run(J, ta, tb_, r2, 128, 128, nm);
    }
  }
  // ---- 2. MN-major SS from the same tiles: D[i][j] = sum_m A[m][i] B[m][j]; 8 MMAs of 16 samples
  {
    std::vector<double> ref(128 * 128, 0.0);
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < 128; ++j) {
        double s = 0;
        for (int m = 0; m < 128; ++m) s += a_at(m, i) * b_at(m, j);
        ref[i * 128 + j] = s;
      }
    struct V { uint32_t lbo, sbo; const char* nm; } vs[] = {
        {16384, 1024, "2. SS MN-major SW128 LBO=16K SBO=1K"},
        {1024, 16384, "2. SS MN-major SW128 LBO=1K SBO=16K (swapped)"}};
    for (auto& v : vs) {
      MmaJob J{};
      J.a_bytes = 32768; J.b_bytes = 32768; J.n_mma = 8;
      J.a_lo0 = 0; J.a_step = 2048; J.b_lo0 = 0; J.b_step = 2048;
      J.a_lbo = v.lbo >> 4; J.b_lbo = v.lbo >> 4; J.a_hi = desc_hi(v.sbo, 2); J.b_hi = desc_hi(v.sbo, 2);
      J.idesc = idesc_f16(128, 128, 1, 1); J.d_cols = 128;
      run(J, ta, tb_, ref, 128, 128, v.nm);
    }
    // mixed: A MN-major (units of A), B K-major is a different contraction; skip.
    // N = 32 variant (layer 0: D1'[n][k] = sum_m G[m][n] enc[m][k], B = enc tile, 32 columns of slab 0)
    {
      std::vector<double> r32(128 * 32, 0.0);
      for (int i = 0; i < 128; ++i)
        for (int j = 0; j < 32; ++j) r32[i * 32 + j] = ref[i * 128 + j];
      MmaJob J{};
      J.a_bytes = 32768; J.b_bytes = 32768; J.n_mma = 8;
      J.a_step = 2048; J.b_step = 2048;
      J.a_lbo = 16384 >> 4; J.b_lbo = 16384 >> 4; J.a_hi = desc_hi(1024, 2); J.b_hi = desc_hi(1024, 2);
      J.idesc = idesc_f16(128, 32, 1, 1); J.d_cols = 32;
      run(J, ta, tb_, r32, 128, 32, "2. SS MN-major N=32 (first 32 columns of B)");
    }
  }
  // ---- 3. ones block: D[i][j] = sum_m A[m][i] * 1, N = 8, B = 512 B of fp16 1.0 re-read by every MMA
  {
    std::vector<double> ref(128 * 8, 0.0);
    for (int i = 0; i < 128; ++i) {
      double s = 0;
      for (int m = 0; m < 128; ++m) s += a_at(m, i);
      for (int j = 0; j < 8; ++j) ref[i * 8 + j] = s;
    }
    std::vector<unsigned char> ones(1024);
    for (int i = 0; i < 512; ++i) {
      const uint16_t o = h16(1.f);
      memcpy(&ones[i * 2], &o, 2);
    }
    MmaJob J{};
    J.a_bytes = 32768; J.b_bytes = 1024; J.n_mma = 8;
    J.a_step = 2048; J.b_step = 0;
    J.a_lbo = 16384 >> 4; J.b_lbo = 128 >> 4; J.a_hi = desc_hi(1024, 2); J.b_hi = desc_hi(256, 0);
    J.idesc = idesc_f16(128, 8, 1, 0); J.d_cols = 8;
    run(J, ta, ones, ref, 128, 8, "3. column sums: A MN-major x ones (N=8)");
  }
  // ---- 3b. small no-swizzle tile g[128 samples][16]: core matrices of 8 rows x 16 B;
  //          element (m, n): (m >> 3) * 256 + (n >> 3) * 128 + (m & 7) * 16 + (n & 7) * 2
  {
    std::vector<unsigned char> g(4096);
    std::vector<uint16_t> gh(128 * 16);
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 16; ++n) {
        gh[m * 16 + n] = h16((float)rand() / RAND_MAX * 2.f - 1.f);
        const size_t off = (size_t)(m >> 3) * 256 + (size_t)(n >> 3) * 128 + (size_t)(m & 7) * 16 + (size_t)(n & 7) * 2;
        memcpy(&g[off], &gh[m * 16 + n], 2);
      }
    auto g_at = [&](int m, int n) { return (double)f16(gh[m * 16 + n]); };
    // (i) as the K-major A operand (M = sample, K = 16) against a K-major SW128 B whose first 16 columns
    //     are used (N = 128 rows of tile B, K = columns 0..15): D[m][u] = sum_{n<16} g[m][n] B[u][n]
    std::vector<double> ref(128 * 128, 0.0);
    for (int m = 0; m < 128; ++m)
      for (int u = 0; u < 128; ++u) {
        double s = 0;
        for (int n = 0; n < 16; ++n) s += g_at(m, n) * b_at(u, n);
        ref[m * 128 + u] = s;
      }
    struct V { uint32_t lbo, sbo; const char* nm; } vs[] = {
        {128, 256, "3b. A no-swizzle K-major LBO=128 SBO=256"},
        {256, 128, "3b. A no-swizzle K-major LBO=256 SBO=128"}};
    for (auto& v : vs) {
      MmaJob J{};
      J.a_bytes = 4096; J.b_bytes = 32768; J.n_mma = 1;
      J.a_lbo = v.lbo >> 4; J.b_lbo = 1; J.a_hi = desc_hi(v.sbo, 0); J.b_hi = desc_hi(1024, 2);
      J.idesc = idesc_f16(128, 128, 0, 0); J.d_cols = 128;
      run(J, g, tb_, ref, 128, 128, v.nm);
    }
    // (ii) as the MN-major B operand (N = 8 of the 16 columns, K = samples): D[i][n] = sum_m A[m][i] g[m][n]
    std::vector<double> r8(128 * 8, 0.0);
    for (int i = 0; i < 128; ++i)
      for (int n = 0; n < 8; ++n) {
        double s = 0;
        for (int m = 0; m < 128; ++m) s += a_at(m, i) * g_at(m, n);
        r8[i * 8 + n] = s;
      }
    struct W { uint32_t lbo, sbo; const char* nm; } ws[] = {
        {256, 128, "3b. B no-swizzle MN-major N=8 LBO=256 SBO=128"},
        {128, 256, "3b. B no-swizzle MN-major N=8 LBO=128 SBO=256"}};
    for (auto& v : ws) {
      MmaJob J{};
      J.a_bytes = 32768; J.b_bytes = 4096; J.n_mma = 8;
      J.a_step = 2048; J.b_step = 512;   // 16 samples = 2 row groups of 256 B
      J.a_lbo = 16384 >> 4; J.b_lbo = v.lbo >> 4; J.a_hi = desc_hi(1024, 2); J.b_hi = desc_hi(v.sbo, 0);
      J.idesc = idesc_f16(128, 8, 1, 1); J.d_cols = 8;
      run(J, ta, g, r8, 128, 8, v.nm);
    }
  }
  // ---- 4. TS: A packed two FP16 per TMEM column (low half = even k)
  {
    std::vector<double> ref(128 * 128, 0.0);
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 128; ++n) {
        double s = 0;
        for (int u = 0; u < 64; ++u) s += a_at(m, u) * b_at(n, u);
        ref[m * 128 + n] = s;
      }
    std::vector<unsigned char> at(128 * 32 * 4);
    for (int m = 0; m < 128; ++m)
      for (int c = 0; c < 32; ++c) {
        const uint32_t v = (uint32_t)Ah[m * 128 + 2 * c] | ((uint32_t)Ah[m * 128 + 2 * c + 1] << 16);
        memcpy(&at[((size_t)m * 32 + c) * 4], &v, 4);
      }
    MmaJob J{};
    J.a_bytes = 0; J.a_tmem_cols = 32; J.a_tmem_step = 8; J.b_bytes = 32768; J.n_mma = 4;
    J.b_step = 32; J.b_lbo = 1; J.b_hi = desc_hi(1024, 2);
    J.idesc = idesc_f16(128, 128, 0, 0); J.d_cols = 128;
    run(J, at, tb_, ref, 128, 128, "4. TS: A fp16x2 packed in TMEM (8 cols / K=16)");
  }
  // ---- 4b. MMA pacing
  {
    long long* d_t;
    cudaMalloc(&d_t, 8);
    cudaFuncSetAttribute(k_mma_pace, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    const char* names[3] = {"SS K-major", "TS (A in TMEM)", "SS MN-major"};
    for (int mode = 0; mode < 3; ++mode)
      for (int n = 128; n <= 256; n += 128) {
        const int iters = 4096;
        k_mma_pace<<<1, 128, 98304>>>(d_t, mode, n, iters);
        long long t = 0;
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&t, d_t, 8, cudaMemcpyDeviceToHost);
        printf("4b. %-16s M=128 N=%3d K=16: %.1f cycles per MMA (%s)\n", names[mode], n, (double)t / iters, cudaGetErrorString(e));
      }
  }
  // ---- 5. TMA reduce throughput
  {
    float* dst;
    const int layers = 5;
    cudaMalloc(&dst, (size_t)148 * layers * 65536);
    cudaMemset(dst, 0, (size_t)148 * layers * 65536);
    cudaFuncSetAttribute(k_reduce_bw, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int priv = 0; priv < 2; ++priv) {
      const int iters = 400;
      k_reduce_bw<<<148, 128, 65536>>>(dst, 20, priv, layers);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0);
      k_reduce_bw<<<148, 128, 65536>>>(dst, iters, priv, layers);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double bytes = 148.0 * iters * 65536.0;
      printf("5. cp.reduce.async.bulk add.f32 %s: %.3f ms, %.1f GB/s (%s)\n", priv ? "private per-CTA buffers" : "one shared 320 KB buffer",
             ms, bytes / ms * 1e-6, cudaGetErrorString(e));
    }
  }
  return 0;
}
