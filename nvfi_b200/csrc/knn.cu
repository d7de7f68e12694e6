// Brute-force k nearest neighbours of a point cloud in itself or in a second cloud, for the smoothness
// loss of the segmentation trainer (utils/seg_loss.py:78-90 calls pytorch3d.ops.knn_points(pc, pc, K=k),
// which the image does not have; published behaviour restated: squared L2 distances, the K smallest per
// query in ascending order, with their indices).  One thread owns one query and keeps its K best in
// registers (insertion into a sorted list); the reference points stream through shared memory in tiles of
// 1024, so every global read is coalesced and shared by the 256 queries of a block.  SURVEY.md section 8
// row f4; not on the render path.
#include "nvfi_common.cuh"

namespace nvfi {

constexpr int KNN_TILE = 1024;
constexpr int KNN_MAXK = 16;

template <int K>
__global__ void __launch_bounds__(256)
    k_knn(const float* __restrict__ q, const float* __restrict__ p, int nq, int np, float* __restrict__ dist,
          long long* __restrict__ idx) {
  __shared__ float sx[KNN_TILE], sy[KNN_TILE], sz[KNN_TILE];
  const int b = blockIdx.y;
  q += (size_t)b * nq * 3;
  p += (size_t)b * np * 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < nq;
  const float qx = live ? q[i * 3] : 0.f, qy = live ? q[i * 3 + 1] : 0.f, qz = live ? q[i * 3 + 2] : 0.f;
  float bd[K];
  int bi[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    bd[k] = INFINITY;
    bi[k] = -1;
  }
  for (int t0 = 0; t0 < np; t0 += KNN_TILE) {
    const int nt = min(KNN_TILE, np - t0);
    __syncthreads();
    for (int j = threadIdx.x; j < nt; j += blockDim.x) {
      sx[j] = p[(size_t)(t0 + j) * 3];
      sy[j] = p[(size_t)(t0 + j) * 3 + 1];
      sz[j] = p[(size_t)(t0 + j) * 3 + 2];
    }
    __syncthreads();
    if (!live) continue;
    for (int j = 0; j < nt; ++j) {
      // (a - b)^2 summed in x, y, z order with separate roundings, like the library it replaces
      const float dx = __fsub_rn(qx, sx[j]), dy = __fsub_rn(qy, sy[j]), dz = __fsub_rn(qz, sz[j]);
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (d < bd[K - 1]) {   // strict: among equal distances the lower index stays first
        bd[K - 1] = d;
        bi[K - 1] = t0 + j;
#pragma unroll
        for (int k = K - 1; k > 0; --k) {
          if (bd[k] < bd[k - 1]) {
            const float td = bd[k];
            bd[k] = bd[k - 1];
            bd[k - 1] = td;
            const int ti = bi[k];
            bi[k] = bi[k - 1];
            bi[k - 1] = ti;
          }
        }
      }
    }
  }
  if (live) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      dist[((size_t)b * nq + i) * K + k] = bd[k];
      idx[((size_t)b * nq + i) * K + k] = bi[k];
    }
  }
}

}  // namespace nvfi

using namespace nvfi;

extern "C" int nvfi_knn_points(const float* query, const float* points, int32_t batch, int32_t n_query,
                               int32_t n_points, int32_t k, float* dist, int64_t* idx, void* stream) {
  if (!query || !points || !dist || !idx || batch <= 0 || n_query < 0 || n_points < 0) return NVFI_EINVAL;
  if (k <= 0 || k > KNN_MAXK) return NVFI_EUNSUPPORTED;
  if (n_query == 0) return NVFI_OK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((n_query + 255) / 256), (unsigned)batch);
  long long* ip = reinterpret_cast<long long*>(idx);
#define NVFI_KNN_CASE(KK)                                                                          \
  if (k <= KK) {                                                                                   \
    if (k != KK) return NVFI_EUNSUPPORTED;                                                         \
    NVFI_LAUNCH(k_knn<KK>, grid, 256, 0, st, query, points, n_query, n_points, dist, ip);          \
    return (int)cudaGetLastError();                                                                \
  }
  NVFI_KNN_CASE(1)
  NVFI_KNN_CASE(2)
  NVFI_KNN_CASE(4)
  NVFI_KNN_CASE(8)
  NVFI_KNN_CASE(16)
#undef NVFI_KNN_CASE
  return NVFI_EUNSUPPORTED;
}
