/* nvfi_b200_debug.h -- development probes, built into a SEPARATE library (libnvfi_b200_debug.so) so that
 * nothing of it ships in the product library.  Not part of the reference surface. */
#ifndef NVFI_B200_DEBUG_H
#define NVFI_B200_DEBUG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* One 128x128x128 TF32 tcgen05 MMA, D[k][n] = sum_m At[k][m] G[m][n], A from tensor memory and
 * B = G read from the sample-major swizzled shared-memory tile of the round-1 tensor-core backward
 * (csrc/backward_tc.cu) with caller-supplied descriptor fields (tests/test_gpu_debug_mma.py pins
 * their meaning). */
int nvfi_debug_mma_mn(const float* At, const float* G, float* Dout, uint32_t lbo_field,
                      uint32_t sbo_field, uint32_t kstep_bytes, uint32_t layout_type,
                      uint32_t b_mn_major, void* stream);
#ifdef __cplusplus
}
#endif
#endif
