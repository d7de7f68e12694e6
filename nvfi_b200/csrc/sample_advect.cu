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
#include "mlp_h.cuh"
#include "mlp_tc.cuh"
#include "nvfi_common.cuh"

namespace nvfi {

// ---------------------------------------------------------------------------------------
// MLP back ends of the tile kernels.  Both evaluate the weight net of VelBasis on a tile of
// 128 samples whose (x, y, z, t) sit in shared memory and leave the 6 basis weights in
// outS[0..5][m]:
//   SimtMlp  FP32 FMA tile GEMM (nvfi_common.cuh) — the verification path
//   TcMlp    tcgen05 tensor cores, activations in TMEM (mlp_tc.cuh) — the product path
// ---------------------------------------------------------------------------------------
struct SimtMlp {
  static constexpr int kThreads = NVFI_THREADS;        // threads that take part in the tile work
  static constexpr int kLaunchThreads = NVFI_THREADS;  // threads per CTA
  static constexpr size_t kBytes = (size_t)(NVFI_TM * NVFI_TM + 2 * NVFI_KC * 128) * sizeof(float);
  float* actT;
  float* wS;
  const NvfiLinear* nets[2];
  __device__ void init(unsigned char* p, const NvfiLinear* n0, const NvfiLinear* n1, int) {
    actT = reinterpret_cast<float*>(p);
    wS = actT + NVFI_TM * NVFI_TM;
    nets[0] = n0;
    nets[1] = n1;
  }
  __device__ void finish() {}
  template <int ACT>
  __device__ void eval(int which, float* outS, const float* xs, const float* ys, const float* zs,
                       const float* ts) {
    vel_net_tile<ACT>(nets[which], actT, wS, outS, xs, ys, zs, ts);
  }
};

struct TcMlp {
  static constexpr int kThreads = tc::kThreads;
  static constexpr int kLaunchThreads = tc::kLaunchThreads;   // + the issuer warp
  static constexpr size_t kBytes = sizeof(tc::Ring) + 1024 + sizeof(tc::Ctl);
  tc::Ring* ring;
  tc::Ctl* ctl;
  tc::Issuer is;
  uint32_t dphase, kphase;
  int mode3;
  __device__ void init(unsigned char* p, const NvfiLinear* n0, const NvfiLinear* n1, int mode) {
    // the swizzled operand slabs need 1024-byte alignment in the shared window
    const uint32_t a = tc::smem_u32(p);
    p += (1024u - (a & 1023u)) & 1023u;
    ring = reinterpret_cast<tc::Ring*>(p);
    ctl = reinterpret_cast<tc::Ctl*>(p + sizeof(tc::Ring));
    dphase = 0;
    kphase = 0;
    mode3 = (mode == NVFI_MLP_TF32X3) ? 1 : 0;
    tc::setup(*ctl, n0, n1);
    is.init(*ctl, tc::smem_u32(ring->stage[0]), tc::kStages);
  }
  __device__ void finish() { tc::teardown(*ctl, is); }
  template <int ACT>
  __device__ void eval(int which, float* outS, const float* xs, const float* ys, const float* zs,
                       const float* ts) {
    tc::vel_net_tile_tc<ACT>(*ctl, is, which, outS, xs, ys, zs, ts, dphase, kphase, mode3);
  }
};

// FP16-split tensor cores (mlp_h.cuh) — the product path.  TS form: activations in tensor memory, shared
// memory holds only the weight ring.
struct HMlp {
  static constexpr int kThreads = th::kThreads;
  static constexpr int kLaunchThreads = th::kLaunchThreads;
  static constexpr int kStages = 4;
  static constexpr uint32_t kTmemCols = 512;   // D ping-pong (256) + A operand hi | lo (128)
  static constexpr size_t kBytes = 1024 + (size_t)kStages * th::kStageBytes + sizeof(th::Ctl);
  th::Ctl* ctl;
  th::Issuer is;
  uint32_t dphase, kphase;
  __device__ void init(unsigned char* p, const NvfiLinear* n0, const NvfiLinear* n1, int) {
    const uint32_t a = tc::smem_u32(p);
    p += (1024u - (a & 1023u)) & 1023u;
    ctl = reinterpret_cast<th::Ctl*>(p + (size_t)kStages * th::kStageBytes);
    dphase = 0;
    kphase = 0;
    th::setup(*ctl, n0, n1, kTmemCols);
    is.init(*ctl, tc::smem_u32(p), kStages);
  }
  __device__ void finish() { th::teardown(*ctl, is, kTmemCols); }
  template <int ACT>
  __device__ void eval(int which, float* outS, const float* xs, const float* ys, const float* zs,
                       const float* ts) {
    th::vel_net_tile_h<ACT, true>(*ctl, is, which, outS, xs, ys, zs, ts, 0u, dphase, kphase);
  }
};

template <int NT>
struct SampleQueue {      // capacity: a carried partial tile + one sub-batch of NT raw samples
  int q_idx[NVFI_TM + NT];
  float q_x[3][NVFI_TM + NT];
  int warp_cnt[2][NT / 32];
  int batch;
};

template <int NT>
struct SampleAdvectTail {   // follows the back end's region in dynamic shared memory
  AdvectTile tile;
  SampleQueue<NT> q;
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

template <class Mlp>
__device__ __forceinline__ void sample_advect_body(const NvfiField& F, const NvfiRenderArgs& A,
                                                   const NvfiRenderBuffers& B, int S,
                                                   long long total, int n_batches, int mode,
                                                   int subs = NVFI_SUBS) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Mlp mlp;
  mlp.init(smem_raw, F.vel_net, nullptr, mode);
  constexpr int NT = Mlp::kThreads;
  SampleAdvectTail<NT>& sm = *reinterpret_cast<SampleAdvectTail<NT>*>(smem_raw + Mlp::kBytes);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  int sub = subs;
  long long batch_base = 0;
  bool exhausted = false;
  unsigned n_valid = 0;
  int qc = 0;   // queue fill, identical in every thread
  int par = 0;  // parity of the warp_cnt buffer
  const float off0 = __fsub_rn(A.t, A.base_time);

  for (;;) {
    // ---- produce: fill the queue up to one tile
    while (qc < NVFI_TM && !exhausted) {
      if (sub == subs) {
        if (tid == 0) sm.q.batch = atomicAdd(&B.counters[0], 1);
        __syncthreads();
        const int b = sm.q.batch;
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
      float xn[3] = {0.f, 0.f, 0.f};
      if (tid < NT && idx < total) {     // (a back end may run extra warps that only join barriers)
        push = eval_sample(F, A, B, idx, S, xn);
        B.valid[idx] = push ? 1 : 0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, push);
      if (lane == 0 && warp < NT / 32) sm.q.warp_cnt[par][warp] = __popc(bal);
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
    advect_tile_with(F, sm.tile,
                     [&](const float* xs, const float* ys, const float* zs, const float* ts) {
                       mlp.template eval<ACT_SILU>(0, &sm.tile.wout[0][0], xs, ys, zs, ts);
                     });
    if (tid < n) {
      const long long gi = sm.q.q_idx[start + tid];
      B.x_adv[gi * 3 + 0] = sm.tile.x[0][tid];
      B.x_adv[gi * 3 + 1] = sm.tile.x[1][tid];
      B.x_adv[gi * 3 + 2] = sm.tile.x[2][tid];
      if (B.x_mid != nullptr) {   // midpoint of the last RK2 step, for the backward pass
        B.x_mid[gi * 3 + 0] = sm.tile.xm[0][tid];
        B.x_mid[gi * 3 + 1] = sm.tile.xm[1][tid];
        B.x_mid[gi * 3 + 2] = sm.tile.xm[2][tid];
      }
    }
    __syncthreads();
  }
  mlp.finish();
  if (B.stats) {
    float c = warp_sum((float)n_valid);
    if (lane == 0 && c > 0.f) {
      atomicAdd(reinterpret_cast<unsigned long long*>(B.stats), (unsigned long long)c);
      atomicAdd(reinterpret_cast<unsigned long long*>(B.stats) + 1, (unsigned long long)c);
    }
  }
}

__global__ void __launch_bounds__(NVFI_THREADS, 2)
    k_sample_advect(const __grid_constant__ NvfiField F, const NvfiRenderArgs A,
                    const NvfiRenderBuffers B, int S, long long total, int n_batches) {
  sample_advect_body<SimtMlp>(F, A, B, S, total, n_batches, NVFI_MLP_FP32_SIMT);
}

__global__ void __launch_bounds__(tc::kLaunchThreads, 1)
    k_sample_advect_tc(const __grid_constant__ NvfiField F, const NvfiRenderArgs A,
                       const NvfiRenderBuffers B, int S, long long total, int n_batches, int mode, int subs) {
  sample_advect_body<TcMlp>(F, A, B, S, total, n_batches, mode, subs);
}

// ---------------------------------------------------------------------------------------
// Product path of the render forward: TWO tiles of 128 samples in flight per CTA (mlp_h.cuh,
// vel_net_tile2_h): the queue hands out 256 compacted samples at a time; threads 0..255 own one
// sample each for the per-sample glue (tile = tid >> 7).
// ---------------------------------------------------------------------------------------
struct SampleAdvect2Tail {
  AdvectTile tile[2];
  int q_idx[2 * NVFI_TM + th::kThreads];
  float q_x[3][2 * NVFI_TM + th::kThreads];
  int warp_cnt[2][th::kThreads / 32];
  int batch;
};

// integrate_pos on two tiles (advect_tile_with for 256 rows; models/tensorf_keyframe.py:575-611)
template <class Net2>
__device__ inline void advect_tile2_with(const NvfiField& F, AdvectTile (&T)[2], Net2 net2) {
  const int tid = threadIdx.x;
  const int tt = (tid >> 7) & 1, m = tid & 127;
  AdvectTile& Tm = T[tt];
  for (;;) {
    int active = (tid < 2 * NVFI_TM) ? (fabsf(Tm.off[m]) > 0.f) : 0;
    if (!__syncthreads_or(active)) break;
    net2(false);
    if (tid < 2 * NVFI_TM) {
      const float off = Tm.off[m];
      const float a = fabsf(off);
      float dt = fminf(a, F.dt_max);
      dt = (off > 0.f) ? dt : ((off < 0.f) ? -dt : 0.f);
      const float x = Tm.x[0][m], y = Tm.x[1][m], z = Tm.x[2][m];
      float v[3] = {0.f, 0.f, 0.f};
      if (!gate_outside(F, x, y, z)) {
        const float w[6] = {Tm.wout[0][m], Tm.wout[1][m], Tm.wout[2][m], Tm.wout[3][m], Tm.wout[4][m], Tm.wout[5][m]};
        basis_velocity(w, x, y, z, v);
      }
      const float hdt = 0.5f * dt;
      Tm.xm[0][m] = __fsub_rn(x, __fmul_rn(hdt, v[0]));
      Tm.xm[1][m] = __fsub_rn(y, __fmul_rn(hdt, v[1]));
      Tm.xm[2][m] = __fsub_rn(z, __fmul_rn(hdt, v[2]));
      Tm.tmid[m] = __fsub_rn(Tm.tcur[m], hdt);
      Tm.dt[m] = dt;
    }
    __syncthreads();
    net2(true);
    if (tid < 2 * NVFI_TM) {
      const float off = Tm.off[m];
      if (fabsf(off) > 0.f) {
        const float dt = Tm.dt[m];
        const float xm = Tm.xm[0][m], ym = Tm.xm[1][m], zm = Tm.xm[2][m];
        float v[3] = {0.f, 0.f, 0.f};
        if (!gate_outside(F, xm, ym, zm)) {
          const float w[6] = {Tm.wout[0][m], Tm.wout[1][m], Tm.wout[2][m], Tm.wout[3][m], Tm.wout[4][m], Tm.wout[5][m]};
          basis_velocity(w, xm, ym, zm, v);
        }
        const float x = Tm.x[0][m], y = Tm.x[1][m], z = Tm.x[2][m];
        float nx = __fsub_rn(x, __fmul_rn(dt, v[0]));
        float ny = __fsub_rn(y, __fmul_rn(dt, v[1]));
        float nz = __fsub_rn(z, __fmul_rn(dt, v[2]));
        if (F.vel_gate == NVFI_GATE_SUR && gate_outside(F, nx, ny, nz)) {  // :603-605
          nx = x;
          ny = y;
          nz = z;
        }
        Tm.x[0][m] = nx;
        Tm.x[1][m] = ny;
        Tm.x[2][m] = nz;
        Tm.off[m] = __fsub_rn(off, dt);
        Tm.tcur[m] = __fsub_rn(Tm.tcur[m], dt);
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(th::kLaunchThreads, 1)
    k_sample_advect_h(const __grid_constant__ NvfiField F, const NvfiRenderArgs A,
                      const NvfiRenderBuffers B, int S, long long total, int n_batches, int mode, int subs,
                      int s0, int sw) {
  // Depth wave [s0, s0 + sw) of every ray (the whole ray when sw == S): raw index w -> ray w / sw, sample
  // s0 + w % sw.  From the second wave on a ray that has terminated (ray_term != S) is not advected; its
  // `valid` flags are still written (they are geometric).
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HMlp mlp;
  mlp.init(smem_raw, F.vel_net, nullptr, mode);
  constexpr int NT = HMlp::kThreads;
  constexpr int TILE2 = 2 * NVFI_TM;
  SampleAdvect2Tail& sm = *reinterpret_cast<SampleAdvect2Tail*>(smem_raw + HMlp::kBytes);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t dph[2] = {0u, 0u}, kph[2] = {0u, 0u};
  float* const outS[2] = {&sm.tile[0].wout[0][0], &sm.tile[1].wout[0][0]};

  int sub = subs;
  long long batch_base = 0;
  bool exhausted = false;
  unsigned n_valid = 0, n_geo = 0, n_gate = 0;
  int qc = 0, par = 0;
  const float off0 = __fsub_rn(A.t, A.base_time);
  const bool check_term = (s0 > 0) && (B.ray_term != nullptr);

  for (;;) {
    // ---- produce: fill the queue up to two tiles
    while (qc < TILE2 && !exhausted) {
      if (sub == subs) {
        if (tid == 0) sm.batch = atomicAdd(&B.counters[0], 1);
        __syncthreads();
        const int b = sm.batch;
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
      float xn[3] = {0.f, 0.f, 0.f};
      long long gi = 0;
      if (tid < NT && idx < total) {
        const long long ray = idx / sw;
        gi = ray * S + s0 + (idx - ray * sw);
        push = eval_sample(F, A, B, gi, S, xn);
        B.valid[gi] = push ? 1 : 0;
        n_geo += push ? 1u : 0u;
        if (push && check_term && B.ray_term[ray] != S) push = false;
        // Outside the velocity gate the field is zero: v0 = 0 puts the midpoint on x0, which is outside the gate
        // again, so v1 = 0 and the sample does not move — in every RK2 step, exactly (x - dt * 0 = x).  No
        // network evaluation is needed to know that.
        if (push && gate_outside(F, xn[0], xn[1], xn[2])) {
          B.x_adv[gi * 3 + 0] = xn[0];
          B.x_adv[gi * 3 + 1] = xn[1];
          B.x_adv[gi * 3 + 2] = xn[2];
          if (B.x_mid != nullptr) {
            B.x_mid[gi * 3 + 0] = xn[0];
            B.x_mid[gi * 3 + 1] = xn[1];
            B.x_mid[gi * 3 + 2] = xn[2];
          }
          ++n_gate;    // counted with the advected samples (advected by a zero velocity), not with the MLP's
          push = false;
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, push);
      if (lane == 0 && warp < NT / 32) sm.warp_cnt[par][warp] = __popc(bal);
      const int tot = __syncthreads_count(push);
      if (push) {
        int pos = qc + __popc(bal & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) pos += sm.warp_cnt[par][w];
        sm.q_idx[pos] = (int)gi;
        sm.q_x[0][pos] = xn[0];
        sm.q_x[1][pos] = xn[1];
        sm.q_x[2][pos] = xn[2];
        ++n_valid;
      }
      qc += tot;
      par ^= 1;
    }
    if (qc == 0) break;
    __syncthreads();
    const int n = min(TILE2, qc);
    const int start = qc - n;
    qc = start;
    // ---- consume two tiles (the second may be partly or wholly padding at the very end)
    if (tid < TILE2) {
      AdvectTile& Tm = sm.tile[tid >> 7];
      const int m = tid & 127;
      const bool live = tid < n;
      Tm.x[0][m] = live ? sm.q_x[0][start + tid] : 0.f;
      Tm.x[1][m] = live ? sm.q_x[1][start + tid] : 0.f;
      Tm.x[2][m] = live ? sm.q_x[2][start + tid] : 0.f;
      Tm.tcur[m] = A.t;
      Tm.off[m] = live ? off0 : 0.f;
    }
    __syncthreads();
    advect_tile2_with(F, sm.tile, [&](bool mid) {
      AdvectTile& T0 = sm.tile[0];
      AdvectTile& T1 = sm.tile[1];
      const float* const in[2][4] = {
          {mid ? T0.xm[0] : T0.x[0], mid ? T0.xm[1] : T0.x[1], mid ? T0.xm[2] : T0.x[2], mid ? T0.tmid : T0.tcur},
          {mid ? T1.xm[0] : T1.x[0], mid ? T1.xm[1] : T1.x[1], mid ? T1.xm[2] : T1.x[2], mid ? T1.tmid : T1.tcur}};
      th::vel_net_tile2_h<ACT_SILU>(*mlp.ctl, mlp.is, 0, outS, in, dph, kph);
    });
    if (tid < n) {
      const AdvectTile& Tm = sm.tile[tid >> 7];
      const int m = tid & 127;
      const long long gi = sm.q_idx[start + tid];
      B.x_adv[gi * 3 + 0] = Tm.x[0][m];
      B.x_adv[gi * 3 + 1] = Tm.x[1][m];
      B.x_adv[gi * 3 + 2] = Tm.x[2][m];
      if (B.x_mid != nullptr) {   // midpoint of the last RK2 step, for the backward pass
        B.x_mid[gi * 3 + 0] = Tm.xm[0][m];
        B.x_mid[gi * 3 + 1] = Tm.xm[1][m];
        B.x_mid[gi * 3 + 2] = Tm.xm[2][m];
      }
    }
    __syncthreads();
  }
  mlp.finish();
  if (B.stats) {   // [0] in-box samples, [1] samples advected (fewer with early ray termination), [3] of those: by the MLP
    const float c = warp_sum((float)n_valid), cg = warp_sum((float)n_geo), cz = warp_sum((float)n_gate);
    if (lane == 0 && cg > 0.f) atomicAdd(reinterpret_cast<unsigned long long*>(B.stats), (unsigned long long)cg);
    if (lane == 0 && c + cz > 0.f)
      atomicAdd(reinterpret_cast<unsigned long long*>(B.stats) + 1, (unsigned long long)(c + cz));
    if (lane == 0 && c > 0.f) atomicAdd(reinterpret_cast<unsigned long long*>(B.stats) + 3, (unsigned long long)c);
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
struct PointAdvectTail {   // follows the back end's region in dynamic shared memory
  AdvectTile tile;
  int next;
};

// integrate_pos with per-point t / base (models/tensorf_keyframe.py:575-611).
template <class Mlp>
__device__ __forceinline__ void integrate_pos_body(const NvfiField& F, const float* __restrict__ x,
                                                   const float* __restrict__ t,
                                                   const float* __restrict__ base, long long n,
                                                   float* __restrict__ out, int* counter, int mode) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Mlp mlp;
  mlp.init(smem_raw, F.vel_net, nullptr, mode);
  PointAdvectTail& sm = *reinterpret_cast<PointAdvectTail*>(smem_raw + Mlp::kBytes);
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
    advect_tile_with(F, sm.tile,
                     [&](const float* xs, const float* ys, const float* zs, const float* ts) {
                       mlp.template eval<ACT_SILU>(0, &sm.tile.wout[0][0], xs, ys, zs, ts);
                     });
    if (tid < NVFI_TM && i < n) {
      out[i * 3 + 0] = sm.tile.x[0][tid];
      out[i * 3 + 1] = sm.tile.x[1][tid];
      out[i * 3 + 2] = sm.tile.x[2][tid];
    }
    __syncthreads();
  }
  mlp.finish();
}

__global__ void __launch_bounds__(NVFI_THREADS, 2)
    k_integrate_pos(const __grid_constant__ NvfiField F, const float* __restrict__ x,
                    const float* __restrict__ t, const float* __restrict__ base, long long n,
                    float* __restrict__ out, int* counter) {
  integrate_pos_body<SimtMlp>(F, x, t, base, n, out, counter, NVFI_MLP_FP32_SIMT);
}
__global__ void __launch_bounds__(tc::kLaunchThreads, 1)
    k_integrate_pos_tc(const __grid_constant__ NvfiField F, const float* __restrict__ x,
                       const float* __restrict__ t, const float* __restrict__ base, long long n,
                       float* __restrict__ out, int* counter, int mode) {
  integrate_pos_body<TcMlp>(F, x, t, base, n, out, counter, mode);
}

__global__ void __launch_bounds__(th::kLaunchThreads, 1)
    k_integrate_pos_h(const __grid_constant__ NvfiField F, const float* __restrict__ x,
                      const float* __restrict__ t, const float* __restrict__ base, long long n,
                      float* __restrict__ out, int* counter, int mode) {
  integrate_pos_body<HMlp>(F, x, t, base, n, out, counter, mode);
}

// VelBasis.forward (full != 0 -> (n,6) = [v, a]) or the gated velocity (n,3).
template <class Mlp>
__device__ __forceinline__ void velocity_body(const NvfiField& F, const float* __restrict__ xyzt,
                                              long long n, int full, float* __restrict__ out,
                                              int* counter, int mode) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Mlp mlp;
  mlp.init(smem_raw, F.vel_net, full ? F.acc_net : nullptr, mode);
  PointAdvectTail& sm = *reinterpret_cast<PointAdvectTail*>(smem_raw + Mlp::kBytes);
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
    mlp.template eval<ACT_SILU>(0, &sm.tile.wout[0][0], sm.tile.x[0], sm.tile.x[1], sm.tile.x[2],
                                sm.tile.tcur);
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
      mlp.template eval<ACT_RELU>(1, &sm.tile.wout[0][0], sm.tile.x[0], sm.tile.x[1], sm.tile.x[2],
                                  sm.tile.tcur);
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
  mlp.finish();
}

__global__ void __launch_bounds__(NVFI_THREADS, 2)
    k_velocity(const __grid_constant__ NvfiField F, const float* __restrict__ xyzt, long long n,
               int full, float* __restrict__ out, int* counter) {
  velocity_body<SimtMlp>(F, xyzt, n, full, out, counter, NVFI_MLP_FP32_SIMT);
}
__global__ void __launch_bounds__(tc::kLaunchThreads, 1)
    k_velocity_tc(const __grid_constant__ NvfiField F, const float* __restrict__ xyzt, long long n,
                  int full, float* __restrict__ out, int* counter, int mode) {
  velocity_body<TcMlp>(F, xyzt, n, full, out, counter, mode);
}

__global__ void __launch_bounds__(th::kLaunchThreads, 1)
    k_velocity_h(const __grid_constant__ NvfiField F, const float* __restrict__ xyzt, long long n,
                 int full, float* __restrict__ out, int* counter, int mode) {
  velocity_body<HMlp>(F, xyzt, n, full, out, counter, mode);
}

}  // namespace nvfi

using namespace nvfi;

static int num_sms() { return device_sms(); }

// The tensor-core path needs the weight images of every layer.
static bool has_umma(const NvfiLinear* net) {
  for (int l = 0; l < NVFI_VEL_LAYERS; ++l)
    if (!net[l].umma || net[l].umma_rows != (l == NVFI_VEL_LAYERS - 1 ? 16 : 128)) return false;
  return true;
}

static bool has_himg(const NvfiLinear* net) {
  for (int l = 0; l < NVFI_VEL_LAYERS; ++l)
    if (!net[l].himg) return false;
  return true;
}

// One depth wave [s0, s0 + sw) of the product path (nvfi_render_forward's early-termination loop).
extern "C" int nvfi_launch_sample_advect_wave(const NvfiField* F, const NvfiRenderArgs* A,
                                              const NvfiRenderBuffers* B, int s0, int sw, cudaStream_t st) {
  const int S = F->n_samples;
  const long long total = (long long)A->n_rays * sw;
  if (total <= 0) return NVFI_OK;
  if ((long long)A->n_rays * S >= (1ll << 31)) return NVFI_EUNSUPPORTED;
  if (!has_himg(F->vel_net)) return NVFI_EINVAL;
  NVFI_CUDA_OK(cudaMemsetAsync(B->counters, 0, sizeof(int32_t), st));   // the batch counter of this wave
  const int subs = grab_subs(total, HMlp::kThreads, num_sms());
  const int per_batch = subs * HMlp::kThreads;
  const int n_batches = (int)((total + per_batch - 1) / per_batch);
  const size_t smem = HMlp::kBytes + sizeof(SampleAdvect2Tail);
  int rc = ensure_smem<k_sample_advect_h>(smem);
  if (rc != NVFI_OK) return rc;
  const int grid = min(n_batches, num_sms());
  NVFI_LAUNCH(k_sample_advect_h, grid, HMlp::kLaunchThreads, smem, st, *F, *A, *B, S, total, n_batches,
              NVFI_MLP_F16X3, subs, s0, sw);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_launch_chunk_inside(const NvfiField* F, const NvfiRenderArgs* A, const NvfiRenderBuffers* B,
                                        cudaStream_t st) {
  const int n_chunks = (int)((A->n_rays + A->ray_chunk - 1) / A->ray_chunk);
  NVFI_LAUNCH(k_chunk_inside, n_chunks, 256, 0, st, *F, A->rays_o, A->n_rays, A->ray_chunk, B->chunk_inside);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_launch_sample_advect(const NvfiField* F, const NvfiRenderArgs* A,
                                         const NvfiRenderBuffers* B, cudaStream_t st) {
  const int S = F->n_samples;
  const long long total = (long long)A->n_rays * S;
  if (total <= 0) return NVFI_OK;
  if (total >= (1ll << 31)) return NVFI_EUNSUPPORTED;  // queue indices are int32
  NVFI_CUDA_OK((cudaError_t)nvfi_launch_chunk_inside(F, A, B, st));
  if (!A->advect) {
    const long long blocks = (total + 255) / 256;
    const int grid = (int)(blocks < (long long)num_sms() * 16 ? blocks : (long long)num_sms() * 16);
    NVFI_LAUNCH(k_sample_only, grid, 256, 0, st, *F, *A, *B, S, total);
    return (int)cudaGetLastError();
  }
  const int mode = mlp_mode_of(F);
  if (mode == NVFI_MLP_F16X3) {
    if (!has_himg(F->vel_net)) return NVFI_EINVAL;
    const int subs = grab_subs(total, HMlp::kThreads, num_sms());
    const int per_batch = subs * HMlp::kThreads;
    const int n_batches = (int)((total + per_batch - 1) / per_batch);
    const size_t smem = HMlp::kBytes + sizeof(SampleAdvect2Tail);
    int rc = ensure_smem<k_sample_advect_h>(smem);
    if (rc != NVFI_OK) return rc;
    const int grid = min(n_batches, num_sms());
    NVFI_LAUNCH(k_sample_advect_h, grid, HMlp::kLaunchThreads, smem, st, *F, *A, *B, S, total, n_batches, mode, subs,
                0, S);
    return (int)cudaGetLastError();
  }
  if (mode != NVFI_MLP_FP32_SIMT) {
    if (!has_umma(F->vel_net)) return NVFI_EINVAL;
    // raw samples per atomically grabbed batch: 8 x 512 for a frame, fewer for a training batch
    // (2 048 rays x 192 samples are 96 such batches: a third of the SMs would stay idle)
    const int subs = grab_subs(total, TcMlp::kThreads, num_sms());
    const int per_batch = subs * TcMlp::kThreads;
    const int n_batches = (int)((total + per_batch - 1) / per_batch);
    const size_t smem = TcMlp::kBytes + sizeof(SampleAdvectTail<TcMlp::kThreads>);
    int rc = ensure_smem<k_sample_advect_tc>(smem);
    if (rc != NVFI_OK) return rc;
    const int grid = min(n_batches, num_sms());
    NVFI_LAUNCH(k_sample_advect_tc, grid, TcMlp::kLaunchThreads, smem, st, *F, *A, *B, S, total, n_batches, mode, subs);
    return (int)cudaGetLastError();
  }
  const int n_batches = (int)((total + NVFI_SUBS * NVFI_THREADS - 1) / (NVFI_SUBS * NVFI_THREADS));
  const size_t smem = SimtMlp::kBytes + sizeof(SampleAdvectTail<NVFI_THREADS>);
  int rc = ensure_smem<k_sample_advect>(smem);
  if (rc != NVFI_OK) return rc;
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
  NVFI_CUDA_OK(cudaMemsetAsync(counters, 0, sizeof(int32_t), st));
  const long long n_tiles = (n + NVFI_TM - 1) / NVFI_TM;
  const int mode = mlp_mode_of(F);
  if (mode == NVFI_MLP_F16X3) {
    if (!has_himg(F->vel_net)) return NVFI_EINVAL;
    const size_t smem = HMlp::kBytes + sizeof(PointAdvectTail);
    int rc = ensure_smem<k_integrate_pos_h>(smem);
    if (rc != NVFI_OK) return rc;
    const int grid = (int)(n_tiles < (long long)num_sms() ? n_tiles : (long long)num_sms());
    NVFI_LAUNCH(k_integrate_pos_h, grid, HMlp::kLaunchThreads, smem, st, *F, x, t, base, n, out, counters, mode);
    return (int)cudaGetLastError();
  }
  if (mode != NVFI_MLP_FP32_SIMT) {
    if (!has_umma(F->vel_net)) return NVFI_EINVAL;
    const size_t smem = TcMlp::kBytes + sizeof(PointAdvectTail);
    int rc = ensure_smem<k_integrate_pos_tc>(smem);
    if (rc != NVFI_OK) return rc;
    const int grid = (int)(n_tiles < (long long)num_sms() ? n_tiles : (long long)num_sms());
    NVFI_LAUNCH(k_integrate_pos_tc, grid, TcMlp::kLaunchThreads, smem, st, *F, x, t, base, n, out, counters, mode);
    return (int)cudaGetLastError();
  }
  const size_t smem = SimtMlp::kBytes + sizeof(PointAdvectTail);
  int rc = ensure_smem<k_integrate_pos>(smem);
  if (rc != NVFI_OK) return rc;
  const int grid = (int)(n_tiles < (long long)num_sms() * 2 ? n_tiles : (long long)num_sms() * 2);
  NVFI_LAUNCH(k_integrate_pos, grid, NVFI_THREADS, smem, st, *F, x, t, base, n, out, counters);
  return (int)cudaGetLastError();
}

extern "C" int nvfi_velocity(const NvfiField* F, const float* xyzt, int64_t n, int32_t full,
                             float* out, int32_t* counters, void* stream) {
  if (!F || !xyzt || !out || !counters || n < 0) return NVFI_EINVAL;
  if (n == 0) return NVFI_OK;
  cudaStream_t st = (cudaStream_t)stream;
  NVFI_CUDA_OK(cudaMemsetAsync(counters, 0, sizeof(int32_t), st));
  const long long n_tiles = (n + NVFI_TM - 1) / NVFI_TM;
  const int mode = mlp_mode_of(F);
  if (mode == NVFI_MLP_F16X3) {
    if (!has_himg(F->vel_net) || (full && !has_himg(F->acc_net))) return NVFI_EINVAL;
    const size_t smem = HMlp::kBytes + sizeof(PointAdvectTail);
    int rc = ensure_smem<k_velocity_h>(smem);
    if (rc != NVFI_OK) return rc;
    const int grid = (int)(n_tiles < (long long)num_sms() ? n_tiles : (long long)num_sms());
    NVFI_LAUNCH(k_velocity_h, grid, HMlp::kLaunchThreads, smem, st, *F, xyzt, n, full, out, counters, mode);
    return (int)cudaGetLastError();
  }
  if (mode != NVFI_MLP_FP32_SIMT) {
    if (!has_umma(F->vel_net) || (full && !has_umma(F->acc_net))) return NVFI_EINVAL;
    const size_t smem = TcMlp::kBytes + sizeof(PointAdvectTail);
    int rc = ensure_smem<k_velocity_tc>(smem);
    if (rc != NVFI_OK) return rc;
    const int grid = (int)(n_tiles < (long long)num_sms() ? n_tiles : (long long)num_sms());
    NVFI_LAUNCH(k_velocity_tc, grid, TcMlp::kLaunchThreads, smem, st, *F, xyzt, n, full, out, counters, mode);
    return (int)cudaGetLastError();
  }
  const size_t smem = SimtMlp::kBytes + sizeof(PointAdvectTail);
  int rc = ensure_smem<k_velocity>(smem);
  if (rc != NVFI_OK) return rc;
  const int grid = (int)(n_tiles < (long long)num_sms() * 2 ? n_tiles : (long long)num_sms() * 2);
  NVFI_LAUNCH(k_velocity, grid, NVFI_THREADS, smem, st, *F, xyzt, n, full, out, counters);
  return (int)cudaGetLastError();
}
