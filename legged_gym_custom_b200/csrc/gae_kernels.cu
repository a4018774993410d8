// Learner-side scalar kernels (compiled with -fmad=false): GAE + advantage normalisation and
// the per-step reward bootstrap.
//
//   K4 gae_scan_kernel       one thread per env, reverse-time scan over T (rollout_storage.py:110-121);
//                            [T,N] layout -> every load/store of a warp is one contiguous 128 B line.
//                            All T rewards/values/dones of an env are independent loads issued up front.
//   K5 adv_normalize_kernel  (A - mean) / (std_unbiased + 1e-8) over T*N (rollout_storage.py:123-124);
//                            mean/std from fp64 partial sums reduced in a fixed order with warp shuffles.
#include "common.cuh"

constexpr int kGaeThreads = 128;
constexpr int kMaxT = 64;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int T_MAX>
__global__ void __launch_bounds__(kGaeThreads)
gae_scan_kernel(const float* __restrict__ rewards, const uint8_t* __restrict__ dones, const float* __restrict__ values,
                const float* __restrict__ last_values, float* __restrict__ returns, float* __restrict__ advantages,
                int T, int N, float gamma, float lam, double* __restrict__ partials) {
  const int e = blockIdx.x * kGaeThreads + threadIdx.x;
  double s = 0.0, s2 = 0.0;
  if (e < N) {
    float r[T_MAX], v[T_MAX], nnt[T_MAX];
#pragma unroll
    for (int t = 0; t < T_MAX; ++t) {
      if (t < T) {
        r[t] = rewards[(int64_t)t * N + e];
        v[t] = values[(int64_t)t * N + e];
        nnt[t] = 1.0f - (float)dones[(int64_t)t * N + e];
      }
    }
    float adv = 0.0f, next_v = last_values[e];
#pragma unroll
    for (int t = T_MAX - 1; t >= 0; --t) {
      if (t < T) {
        const float delta = (r[t] + (nnt[t] * gamma) * next_v) - v[t];
        adv = delta + ((nnt[t] * gamma) * lam) * adv;
        const float ret = adv + v[t];
        const float a = ret - v[t];                      // self.advantages = self.returns - self.values
        returns[(int64_t)t * N + e] = ret;
        advantages[(int64_t)t * N + e] = a;
        s += (double)a;
        s2 += (double)a * (double)a;
        next_v = v[t];
      }
    }
  }
  __shared__ double sm[2][kGaeThreads / 32];
  s = warp_sum(s);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) {
    sm[0][threadIdx.x >> 5] = s;
    sm[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < kGaeThreads / 32; ++w) {
      a += sm[0][w];
      b += sm[1][w];
    }
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = b;
  }
}

__global__ void __launch_bounds__(256)
adv_normalize_kernel(float* __restrict__ advantages, int64_t n, const double* __restrict__ partials, int num_partials) {
  __shared__ float s_mean, s_inv;
  if (threadIdx.x < 32) {
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < num_partials; i += 32) {
      a += partials[2 * i];
      b += partials[2 * i + 1];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (threadIdx.x == 0) {
      const double mean = a / (double)n;
      double var = (b - a * a / (double)n) / (double)(n - 1);
      var = var < 0.0 ? 0.0 : var;
      s_mean = (float)mean;
      s_inv = (float)sqrt(var) + 1e-8f;
    }
  }
  __syncthreads();
  const float mean = s_mean, denom = s_inv;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    advantages[i] = (advantages[i] - mean) / denom;
}

__global__ void __launch_bounds__(256)
store_step_scalars_kernel(const float* __restrict__ rew, const uint8_t* __restrict__ reset, const uint8_t* __restrict__ time_outs,
                          const float* __restrict__ values, float gamma, float* __restrict__ rewards_t,
                          uint8_t* __restrict__ dones_t, int N) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= N) return;
  float r = rew[e];
  if (time_outs) r = r + gamma * (values[e] * (float)time_outs[e]);   // ppo.py:165-166
  rewards_t[e] = r;
  dones_t[e] = reset[e];
}

extern "C" {

int64_t b200_gae_scratch_bytes(int T, int N) {
  (void)T;
  return (int64_t)((N + kGaeThreads - 1) / kGaeThreads) * 2 * (int64_t)sizeof(double);
}

int b200_compute_returns(const float* rewards, const uint8_t* dones, const float* values, const float* last_values,
                         float* returns, float* advantages, int T, int N, float gamma, float lam, void* scratch, void* stream) {
  B200_CHECK_ARG(rewards && dones && values && last_values && returns && advantages && scratch, "b200_compute_returns: null argument");
  B200_CHECK_ARG(T > 0 && T <= kMaxT && N > 0, "b200_compute_returns: need 0 < T <= %d and N > 0 (T=%d N=%d)", kMaxT, T, N);
  B200_CHECK_ARG((int64_t)T * N > 1, "b200_compute_returns: std needs more than one sample");
  const int blocks = (N + kGaeThreads - 1) / kGaeThreads;
  cudaStream_t st = (cudaStream_t)stream;
  if (T <= 24)
    gae_scan_kernel<24><<<blocks, kGaeThreads, 0, st>>>(rewards, dones, values, last_values, returns, advantages, T, N, gamma, lam, (double*)scratch);
  else
    gae_scan_kernel<kMaxT><<<blocks, kGaeThreads, 0, st>>>(rewards, dones, values, last_values, returns, advantages, T, N, gamma, lam, (double*)scratch);
  B200_CHECK_LAUNCH("gae_scan_kernel");
  const int64_t n = (int64_t)T * N;
  const int nblocks = (int)((n + 256 * 4 - 1) / (256 * 4));
  adv_normalize_kernel<<<nblocks < 1 ? 1 : nblocks, 256, 0, st>>>(advantages, n, (const double*)scratch, blocks);
  B200_CHECK_LAUNCH("adv_normalize_kernel");
  return 0;
}

int b200_store_step_scalars(const float* rew, const uint8_t* reset, const uint8_t* time_outs, const float* values, float gamma,
                            float* rewards_t, uint8_t* dones_t, int N, void* stream) {
  B200_CHECK_ARG(rew && reset && values && rewards_t && dones_t && N > 0, "b200_store_step_scalars: bad argument");
  store_step_scalars_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rew, reset, time_outs, values, gamma, rewards_t, dones_t, N);
  B200_CHECK_LAUNCH("store_step_scalars_kernel");
  return 0;
}

}  // extern "C"
