// Kernel 1 of the render path: ray sampling + (eval) alpha-mask skip + RK2 backward
// advection of every valid sample through the velocity MLP.
//
// Replaces, for one Renderer.forward call (all chunks at once):
//   TensorBase.sample_ray            models/tensorf_base.py:290-314
//   normalize_coord                  models/tensorf_base.py:241-242
//   AlphaGridMask.sample_alpha skip  models/tensorf_keyframe.py:656-661
//   integrate_pos / VelocityAABB(.Sur) / VelBasis.get_vel
//                                    models/tensorf_keyframe.py:575-611,
//                                    models/velocity_field.py:21-98
//
// Design: persistent CTAs pull batches of 2048 raw (ray, sample) slots from an atomic
// counter, evaluate the sampler per thread, and compact the valid slots into a
// shared-memory queue.  Whenever 128 entries are queued they are advected as one tile
// by the FP32 tile-GEMM MLP (nvfi_common.cuh); leftovers carry over, so no tile is
// padded except the very last one of a CTA.
#include "nvfi_common.cuh"

namespace nvfi {

struct SampleQueue {
  int q_idx[NVFI_QCAP];
  float q_x[3][NVFI_QCAP];
  int warp_cnt[2][NVFI_THREADS / 32];
  int batch;
};

struct SampleAdvectSmem {
  float actT[NVFI_TM * NVFI_TM];
  float wS[2 * NVFI_KC * 128];
  AdvectTile tile;
  SampleQueue q;
};

__device__ __forceinline__ bool eval_sample(const NvfiField& F, const NvfiRenderArgs& A,
                                            const NvfiRenderBuffers& B, long long idx, int S,
                                            float xn[3]) {
  const long long ray = idx / S;
  const int s = (int)(idx - ray * S);
  const float o[3] = {__ldg(A.rays_o + ray * 3), __ldg(A.rays_o + ray * 3 + 1),
                      __ldg(A.rays_o + ray * 3 + 2)};
  const float d[3] = {__ldg(A.rays_d + ray * 3), __ldg(A.rays_d + ray * 3 + 1),
                      __ldg(A.rays_d + ray * 3 + 2)};
  const bool inside = B.chunk_inside[ray / A.ray_chunk] != 0;
  const float tmin = ray_tmin(F, o, d, inside);
  const bool train = A.jitter != nullptr;
  const float u = train ? __ldg(A.jitter + ray) : 0.f;
  const float z = sample_z(tmin, F.step_size, s, u, train);
  bool valid = sample_point(F, o, d, z, xn);
  if (valid && !A.training && F.alpha_volume != nullptr) valid = alpha_mask_keep(F, xn);
  return valid;
}

// No advection (keyframe time, or use_vel == 0): x_adv = normalised sample position.
__global__ void __launch_bounds__(256) k_sample_only(const NvfiField F, const NvfiRenderArgs A,
                                                     const NvfiRenderBuffers B, int S,
                                                     long long total) {
  long long n_valid = 0;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    float xn[3];
    const bool valid = eval_sample(F, A, B, idx, S, xn);
    B.valid[idx] = valid ? 1 : 0;
    if (valid) {
      B.x_adv[idx * 3 + 0] = xn[0];
      B.x_adv[idx * 3 + 1] = xn[1];
      B.x_adv[idx * 3 + 2] = xn[2];
      ++n_valid;
    }
  }
  if (B.stats) {
    float c = warp_sum((float)n_valid);  // exact for counts < 2^24 per warp
    if ((threadIdx.x & 31) == 0 && c > 0.f)
      atomicAdd(reinterpret_cast<unsigned long long*>(B.stats), (unsigned long long)c);
  }
}

__global__ void __launch_bounds__(NVFI_THREADS, 2)
    k_sample_advect(const NvfiField F, const NvfiRenderArgs A, const NvfiRenderBuffers B, int S,
                    long long total, int n_batches) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SampleAdvectSmem& sm = *reinterpret_cast<SampleAdvectSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  int sub = NVFI_SUBS;
  long long batch_base = 0;
  bool exhausted = false;
  unsigned n_valid = 0;
  int qc = 0;   // queue fill, identical in every thread
  int par = 0;  // parity of the warp_cnt buffer
  const float off0 = __fsub_rn(A.t, A.base_time);

  for (;;) {
    // ---- produce: fill the queue up to one tile
    while (qc < NVFI_TM && !exhausted) {
      if (sub == NVFI_SUBS) {
        if (tid == 0) sm.q.batch = atomicAdd(&B.counters[0], 1);
        __syncthreads();
        const int b = sm.q.batch;
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
      bool push = false;
      float xn[3] = {0.f, 0.f, 0.f};
      if (idx < total) {
        push = eval_sample(F, A, B, idx, S, xn);
        B.valid[idx] = push ? 1 : 0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, push);
      if (lane == 0) sm.q.warp_cnt[par][warp] = __popc(bal);
      const int tot = __syncthreads_count(push);
      if (push) {
        int pos = qc + __popc(bal & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) pos += sm.q.warp_cnt[par][w];
        sm.q.q_idx[pos] = (int)idx;
        sm.q.q_x[0][pos] = xn[0];
        sm.q.q_x[1][pos] = xn[1];
        sm.q.q_x[2][pos] = xn[2];
        ++n_valid;
      }
      qc += tot;
      par ^= 1;
    }
    if (qc == 0) break;
    __syncthreads();  // queue writes of the last sub-batch are visible
    const int n = min(NVFI_TM, qc);
    const int start = qc - n;
    qc = start;
    // ---- consume one tile
    if (tid < NVFI_TM) {
      const bool live = tid < n;
      sm.tile.x[0][tid] = live ? sm.q.q_x[0][start + tid] : 0.f;
      sm.tile.x[1][tid] = live ? sm.q.q_x[1][start + tid] : 0.f;
      sm.tile.x[2][tid] = live ? sm.q.q_x[2][start + tid] : 0.f;
      sm.tile.tcur[tid] = A.t;
      sm.tile.off[tid] = live ? off0 : 0.f;
    }
    __syncthreads();
    advect_tile(F, sm.tile, sm.actT, sm.wS);
    if (tid < n) {
      const long long gi = sm.q.q_idx[start + tid];
      B.x_adv[gi * 3 + 0] = sm.tile.x[0][tid];
      B.x_adv[gi * 3 + 1] = sm.tile.x[1][tid];
      B.x_adv[gi * 3 + 2] = sm.tile.x[2][tid];
    }
    __syncthreads();
  }
  if (B.stats) {
    float c = warp_sum((float)n_valid);
    if (lane == 0 && c > 0.f) {
      atomicAdd(reinterpret_cast<unsigned long long*>(B.stats), (unsigned long long)c);
      atomicAdd(reinterpret_cast<unsigned long long*>(B.stats) + 1, (unsigned long long)c);
    }
  }
}

// Chunk-global predicate of sample_ray (models/tensorf_base.py:294): one flag per chunk
// of `ray_chunk` rays: any component of any origin inside [aabb_min, aabb_max].
__global__ void k_chunk_inside(const NvfiField F, const float* __restrict__ rays_o, long long n,
                               int ray_chunk, unsigned char* __restrict__ flags) {
  const long long c = blockIdx.x;
  const long long r0 = c * ray_chunk;
  const long long r1 = min(n, r0 + (long long)ray_chunk);
  int any = 0;
  for (long long i = r0 * 3 + threadIdx.x; i < r1 * 3; i += blockDim.x) {
    const int a = (int)(i % 3);
    const float v = rays_o[i];
    any |= (F.aabb_min[a] <= v) & (v <= F.aabb_max[a]);
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) flags[c] = any ? 1 : 0;
}

// ---------------------------------------------------------------------------------------
// Stand-alone field queries
// ---------------------------------------------------------------------------------------
struct PointAdvectSmem {
  float actT[NVFI_TM * NVFI_TM];
  float wS[2 * NVFI_KC * 128];
  AdvectTile tile;
  int next;
};

// integrate_pos with per-point t / base (models/tensorf_keyframe.py:575-611).
__global__ void __launch_bounds__(NVFI_THREADS, 2)
    k_integrate_pos(const NvfiField F, const float* __restrict__ x, const float* __restrict__ t,
                    const float* __restrict__ base, long long n, float* __restrict__ out,
                    int* counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PointAdvectSmem& sm = *reinterpret_cast<PointAdvectSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const long long n_tiles = (n + NVFI_TM - 1) / NVFI_TM;
  for (;;) {
    if (tid == 0) sm.next = atomicAdd(counter, 1);
    __syncthreads();
    const long long tile = sm.next;
    __syncthreads();
    if (tile >= n_tiles) break;
    const long long i = tile * NVFI_TM + tid;
    if (tid < NVFI_TM) {
      const bool live = i < n;
      sm.tile.x[0][tid] = live ? x[i * 3 + 0] : 0.f;
      sm.tile.x[1][tid] = live ? x[i * 3 + 1] : 0.f;
      sm.tile.x[2][tid] = live ? x[i * 3 + 2] : 0.f;
      const float tt = live ? t[i] : 0.f;
      sm.tile.tcur[tid] = tt;
      sm.tile.off[tid] = live ? __fsub_rn(tt, base[i]) : 0.f;
    }
    __syncthreads();
    advect_tile(F, sm.tile, sm.actT, sm.wS);
    if (tid < NVFI_TM && i < n) {
      out[i * 3 + 0] = sm.tile.x[0][tid];
      out[i * 3 + 1] = sm.tile.x[1][tid];
      out[i * 3 + 2] = sm.tile.x[2][tid];
    }
    __syncthreads();
  }
}

// VelBasis.forward (full != 0 -> (n,6) = [v, a]) or the gated velocity (n,3).
__global__ void __launch_bounds__(NVFI_THREADS, 2)
    k_velocity(const NvfiField F, const float* __restrict__ xyzt, long long n, int full,
               float* __restrict__ out, int* counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PointAdvectSmem& sm = *reinterpret_cast<PointAdvectSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const long long n_tiles = (n + NVFI_TM - 1) / NVFI_TM;
  for (;;) {
    if (tid == 0) sm.next = atomicAdd(counter, 1);
    __syncthreads();
    const long long tile = sm.next;
    __syncthreads();
    if (tile >= n_tiles) break;
    const long long i = tile * NVFI_TM + tid;
    if (tid < NVFI_TM) {
      const bool live = i < n;
      sm.tile.x[0][tid] = live ? xyzt[i * 4 + 0] : 0.f;
      sm.tile.x[1][tid] = live ? xyzt[i * 4 + 1] : 0.f;
      sm.tile.x[2][tid] = live ? xyzt[i * 4 + 2] : 0.f;
      sm.tile.tcur[tid] = live ? xyzt[i * 4 + 3] : 0.f;
    }
    __syncthreads();
    vel_net_tile<ACT_SILU>(F.vel_net, sm.actT, sm.wS, &sm.tile.wout[0][0], sm.tile.x[0],
                           sm.tile.x[1], sm.tile.x[2], sm.tile.tcur);
    float v[3] = {0.f, 0.f, 0.f};
    const float px = sm.tile.x[0][tid & 127], py = sm.tile.x[1][tid & 127],
                pz = sm.tile.x[2][tid & 127];
    if (tid < NVFI_TM) {
      const float w[6] = {sm.tile.wout[0][tid], sm.tile.wout[1][tid], sm.tile.wout[2][tid],
                          sm.tile.wout[3][tid], sm.tile.wout[4][tid], sm.tile.wout[5][tid]};
      if (full || !gate_outside(F, px, py, pz)) basis_velocity(w, px, py, pz, v);
    }
    __syncthreads();
    if (full) {
      vel_net_tile<ACT_RELU>(F.acc_net, sm.actT, sm.wS, &sm.tile.wout[0][0], sm.tile.x[0],
                             sm.tile.x[1], sm.tile.x[2], sm.tile.tcur);
      if (tid < NVFI_TM && i < n) {
        const float aw[6] = {sm.tile.wout[0][tid], sm.tile.wout[1][tid], sm.tile.wout[2][tid],
                             sm.tile.wout[3][tid], sm.tile.wout[4][tid], sm.tile.wout[5][tid]};
        float a[3];
        basis_acceleration(aw, px, py, pz, a);
        out[i * 6 + 0] = v[0];
        out[i * 6 + 1] = v[1];
        out[i * 6 + 2] = v[2];
        out[i * 6 + 3] = a[0];
        out[i * 6 + 4] = a[1];
        out[i * 6 + 5] = a[2];
      }
    } else if (tid < NVFI_TM && i < n) {
      out[i * 3 + 0] = v[0];
      out[i * 3 + 1] = v[1];
      out[i * 3 + 2] = v[2];
    }
    __syncthreads();
  }
}

}  // namespace nvfi

using namespace nvfi;

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

extern "C" int nvfi_launch_sample_advect(const NvfiField* F, const NvfiRenderArgs* A,
                                         const NvfiRenderBuffers* B, cudaStream_t st) {
  const int S = F->n_samples;
  const long long total = (long long)A->n_rays * S;
  if (total <= 0) return NVFI_OK;
  if (total >= (1ll << 31)) return NVFI_EUNSUPPORTED;  // queue indices are int32
  const int n_chunks = (int)((A->n_rays + A->ray_chunk - 1) / A->ray_chunk);
  NVFI_LAUNCH(k_chunk_inside, n_chunks, 256, 0, st, *F, A->rays_o, A->n_rays, A->ray_chunk, B->chunk_inside);
  NVFI_CUDA_OK(cudaGetLastError());
  if (!A->advect) {
    const long long blocks = (total + 255) / 256;
    const int grid = (int)(blocks < (long long)num_sms() * 16 ? blocks : (long long)num_sms() * 16);
    NVFI_LAUNCH(k_sample_only, grid, 256, 0, st, *F, *A, *B, S, total);
    return (int)cudaGetLastError();
  }
  const size_t smem = sizeof(SampleAdvectSmem);
  static bool attr_set = false;
  if (!attr_set) {
    NVFI_CUDA_OK(cudaFuncSetAttribute(k_sample_advect, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    attr_set = true;
  }
  const int n_batches = (int)((total + NVFI_SUBS * NVFI_THREADS - 1) / (NVFI_SUBS * NVFI_THREADS));
  const int grid = min(n_batches, num_sms() * 2);
  NVFI_LAUNCH(k_sample_advect, grid, NVFI_THREADS, smem, st, *F, *A, *B, S, total, n_batches);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_integrate_pos(const NvfiField* F, const float* x, const float* t,
                                  const float* base, int64_t n, float* out, int32_t* counters,
                                  void* stream) {
  if (!F || !x || !t || !base || !out || !counters || n < 0) return NVFI_EINVAL;
  if (n == 0) return NVFI_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = sizeof(PointAdvectSmem);
  static bool attr_set = false;
  if (!attr_set) {
    NVFI_CUDA_OK(cudaFuncSetAttribute(k_integrate_pos, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    attr_set = true;
  }
  NVFI_CUDA_OK(cudaMemsetAsync(counters, 0, sizeof(int32_t), st));
  const long long n_tiles = (n + NVFI_TM - 1) / NVFI_TM;
  const int grid = (int)(n_tiles < (long long)num_sms() * 2 ? n_tiles : (long long)num_sms() * 2);
  NVFI_LAUNCH(k_integrate_pos, grid, NVFI_THREADS, smem, st, *F, x, t, base, n, out, counters);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_velocity(const NvfiField* F, const float* xyzt, int64_t n, int32_t full,
                             float* out, int32_t* counters, void* stream) {
  if (!F || !xyzt || !out || !counters || n < 0) return NVFI_EINVAL;
  if (n == 0) return NVFI_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = sizeof(PointAdvectSmem);
  static bool attr_set = false;
  if (!attr_set) {
    NVFI_CUDA_OK(cudaFuncSetAttribute(k_velocity, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    attr_set = true;
  }
  NVFI_CUDA_OK(cudaMemsetAsync(counters, 0, sizeof(int32_t), st));
  const long long n_tiles = (n + NVFI_TM - 1) / NVFI_TM;
  const int grid = (int)(n_tiles < (long long)num_sms() * 2 ? n_tiles : (long long)num_sms() * 2);
  NVFI_LAUNCH(k_velocity, grid, NVFI_THREADS, smem, st, *F, xyzt, n, full, out, counters);
  return (int)cudaGetLastError();
}
