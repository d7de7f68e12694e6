// Development probe (not part of the reference surface): one 128x128x128 TF32 MMA with the A
// operand in tensor memory (A^T[k][m]) and the B operand G[m][n] read MN-major from the
// sample-major swizzled shared-memory tile used by the tensor-core backward, with the
// descriptor fields supplied by the caller.  tests/test_gpu_debug_mma.py pins the field
// semantics the backward relies on against a host GEMM.
#include "nvfi_b200_debug.h"
#include "mlp_tc.cuh"

namespace nvfi {

__global__ void __launch_bounds__(128, 1)
    k_debug_mma_mn(const float* __restrict__ At /*[k][m]*/, const float* __restrict__ G /*[m][n]*/,
                   float* __restrict__ Dout /*[k][n]*/, uint32_t lbo_field, uint32_t sbo_field,
                   uint32_t kstep_bytes, uint32_t layout_type, uint32_t b_mn_major) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* p = smem_raw;
  {
    const uint32_t a = tc::smem_u32(p);
    p += (1024u - (a & 1023u)) & 1023u;
  }
  unsigned char* g = p;   // 64 KB
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);
  // G -> shared memory, sample-major swizzled tile (same addressing as backward_tc.cu g_off)
  for (int i = tid; i < 128 * 128; i += 128) {
    const int m = i >> 7, n = i & 127;
    const uint32_t off = (uint32_t)(((n >> 5) << 14) + ((m >> 3) << 10) + ((m & 7) << 7) +
                                    ((((n & 31) >> 2) ^ (m & 7)) << 4) + ((n & 3) << 2));
    *reinterpret_cast<float*>(g + off) = __uint_as_float(tc::to_tf32(G[i]));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tb = tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  // A^T -> TMEM columns [128, 256): lane k = tid, column m
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = tc::to_tf32(At[tid * 128 + c * 32 + i]);
    tc::tmem_st32(tb + lane_base + 128u + (uint32_t)(c * 32), v);
  }
  tc::tmem_st_wait();
  tc::tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc::tc_fence_after();
    const uint32_t idesc = tc::instr_desc_tf32(128) | (b_mn_major << 16);
    const uint32_t desc_hi = sbo_field | (1u << 14) | (layout_type << 29);
    const uint32_t gb = tc::smem_u32(g);
    for (uint32_t ks = 0; ks < 16; ++ks) {
      // K step ks: 32-wide K block (ks >> 2) is 16 KB further, step (ks & 3) kstep_bytes further
      // (kstep_bytes >= 1024 means: linear stepping, used by the MN-major variants)
      const uint32_t adv = (kstep_bytes >= 1024u) ? ks * kstep_bytes : (ks >> 2) * 16384u + (ks & 3u) * kstep_bytes;
      const uint32_t lo = (((gb + adv) >> 4) & 0x3FFFu) | (lbo_field << 16);
      const uint64_t bd = ((uint64_t)desc_hi << 32) | lo;
      tc::mma_tf32_ts(tb, tb + 128u + ks * 8u, bd, idesc, ks ? 1u : 0u);
    }
    tc::tc_commit(&bar);
  }
  __syncwarp();
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  for (int c = 0; c < 4; ++c) {
    float v[32];
    tc::tmem_ld32(tb + lane_base + (uint32_t)(c * 32), v);
#pragma unroll
    for (int i = 0; i < 32; ++i) Dout[tid * 128 + c * 32 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tb, 256);
  }
}

}  // namespace nvfi

extern "C" int nvfi_debug_mma_mn(const float* At, const float* G, float* Dout, uint32_t lbo_field,
                                 uint32_t sbo_field, uint32_t kstep_bytes, uint32_t layout_type,
                                 uint32_t b_mn_major, void* stream) {
  if (!At || !G || !Dout) return NVFI_EINVAL;
  const size_t smem = 65536 + 1024;
  NVFI_CUDA_OK(cudaFuncSetAttribute(nvfi::k_debug_mma_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nvfi::k_debug_mma_mn<<<1, 128, smem, (cudaStream_t)stream>>>(At, G, Dout, lbo_field, sbo_field, kstep_bytes,
                                                               layout_type, b_mn_major);
  return (int)cudaGetLastError();
}
