// Learner-side non-GEMM kernels (SURVEY.md §2.1 K7 / K8 and the storage traffic of a14 / a18).
//
//   copy_segments_kernel   strided 2-D copies, several (src,dst) pairs per launch   rollout_storage.py:87-105
//   gather_rows_kernel     dst[i,:] = src[idx[i],:]  (the ONE permutation of an iteration, a18)
//   sample_actions_kernel  a = mu + std * z (keyed Philox + Box-Muller), log-prob    actor_critic.py:190-226
//   ppo_loss_kernel        surrogate / clipped value / entropy / ROA regulariser, fwd + bwd   ppo.py:199-270
//   mse_rows_loss_kernel   estimator loss  mean ||pred - target||_2^2, fwd + bwd               ppo.py:224-226
//   l2_rows_loss_kernel    DAgger loss     mean ||target - pred||_2,   fwd + bwd               ppo.py:330-333
//   sumsq / clip_adam      global grad-norm clip fused with the Adam step on flat buffers      ppo.py:228-231, :273-276
//   kl_sum / adaptive_lr   schedule='adaptive': KL(old || new) of the minibatch, learning-rate rule on the device   ppo.py:233-246
//
// All are streaming, HBM/latency-bound kernels; reductions use warp shuffles + one atomic per CTA.
#include "common.cuh"
#include "philox.cuh"

namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// CTA-wide sum of up to 4 values; result valid in thread 0.
template <int NV>
__device__ __forceinline__ void cta_sum(float (&v)[NV], float* smem /* [NV][32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = warp_sum_f(v[i]);
    if (lane == 0) smem[i * 32 + warp] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float t = lane < nw ? smem[i * 32 + lane] : 0.0f;
      v[i] = warp_sum_f(t);
    }
  }
}

}  // namespace

struct CopyArgs {
  B200CopySeg seg[8];
  int nseg, rows;
};

__global__ void __launch_bounds__(256) copy_segments_kernel(const __grid_constant__ CopyArgs a) {
  const B200CopySeg s = a.seg[blockIdx.y];
  const bool v4 = (s.width % 4 == 0) && (s.src_ld % 4 == 0) && (s.dst_ld % 4 == 0) && (((uintptr_t)s.src | (uintptr_t)s.dst) % 16 == 0);
  if (v4) {
    const int w4 = s.width / 4;
    const int64_t total = (int64_t)a.rows * w4;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
      const int64_t r = i / w4;
      const int c = (int)(i % w4);
      reinterpret_cast<float4*>(s.dst + r * s.dst_ld)[c] = reinterpret_cast<const float4*>(s.src + r * s.src_ld)[c];
    }
  } else {
    const int64_t total = (int64_t)a.rows * s.width;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
      const int64_t r = i / s.width;
      const int c = (int)(i % s.width);
      s.dst[r * s.dst_ld + c] = s.src[r * s.src_ld + c];
    }
  }
}

// one warp per destination row
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, int src_ld, const int64_t* __restrict__ idx, float* __restrict__ dst, int dst_ld,
                   int width, int64_t rows) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* s = src + idx[row] * src_ld;
  float* d = dst + row * dst_ld;
  if ((width % 4 == 0) && (src_ld % 4 == 0) && (dst_ld % 4 == 0) && (((uintptr_t)src | (uintptr_t)dst) % 16 == 0)) {
    for (int c = lane; c < width / 4; c += 32) reinterpret_cast<float4*>(d)[c] = __ldg(reinterpret_cast<const float4*>(s) + c);
  } else {
    for (int c = lane; c < width; c += 32) d[c] = s[c];
  }
}

__global__ void __launch_bounds__(256)
gather_bytes_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ idx, uint8_t* __restrict__ dst, int64_t rows) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < rows) dst[i] = src[idx[i]];
}

// actions ~ Normal(mu, std): z from two keyed uniforms (Box-Muller); log_prob summed over actions.
// 16 lanes per env (one per action, A <= 16), 16 envs per CTA: coalesced rows, the per-env log-prob is summed by the group's
// first lane in action order (the order of the serial loop this replaces, so the bits do not change).
// `step_dev` != nullptr: [0] = the noise step counter in device memory, [1] = a ticket the kernel keeps at zero: the LAST CTA
// to finish (every CTA has read the counter by then) advances the counter -- no separate counter launch.
constexpr int kSampleEnvsPerCta = 16;
__global__ void __launch_bounds__(256)
sample_actions_kernel(const float* __restrict__ mu, int ldmu, const float* __restrict__ std, uint64_t seed, uint32_t step,
                      int64_t* __restrict__ step_dev, float* __restrict__ actions, float* __restrict__ logp,
                      float* __restrict__ mu_out, float* __restrict__ sigma_out, int N, int A) {
  const int a = threadIdx.x & 15;
  const int e = blockIdx.x * kSampleEnvsPerCta + (threadIdx.x >> 4);
  if (step_dev) step = (uint32_t)*reinterpret_cast<const volatile int64_t*>(step_dev);
  float term = 0.0f;
  if (e < N && a < A) {
    const float u1 = ((float)(keyed_u32(seed, SITE_ACTION_NOISE, step, e, 2 * a) >> 8) + 0.5f) * 5.9604644775390625e-08f;
    const float u2 = u32_to_uniform(keyed_u32(seed, SITE_ACTION_NOISE, step, e, 2 * a + 1));
    const float z = sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
    const float m = mu[(int64_t)e * ldmu + a], s = std[a];
    const float act = m + s * z;
    actions[(int64_t)e * A + a] = act;
    if (mu_out) mu_out[(int64_t)e * A + a] = m;
    if (sigma_out) sigma_out[(int64_t)e * A + a] = s;
    const float d = act - m;
    term = -(d * d) / (2.0f * s * s) - logf(s) - 0.9189385332046727f;
  }
  const int base = threadIdx.x & 16;           // first lane of this env's group within the warp
  float lp = 0.0f;
  for (int i = 0; i < A; ++i) lp += __shfl_sync(0xffffffffu, term, base + i);
  if (e < N && a == 0) logp[e] = lp;
  if (step_dev) {
    __syncthreads();                           // every thread of the CTA has read the counter
    if (threadIdx.x == 0) {
      unsigned long long* ticket = reinterpret_cast<unsigned long long*>(step_dev + 1);
      __threadfence();
      if (atomicAdd(ticket, 1ull) == (unsigned long long)gridDim.x - 1ull) {
        *ticket = 0ull;
        *step_dev += 1;
      }
    }
  }
}

typedef B200PpoLossArgs PpoLossArgs;

constexpr int kLossThreads = 128;
constexpr int kMaxA = 16;

__global__ void __launch_bounds__(kLossThreads) ppo_loss_kernel(const __grid_constant__ PpoLossArgs p) {
  __shared__ float red[4 * 32];
  __shared__ float dstd_s[kMaxA];
  __shared__ float c_s[kMaxA], c_inv2s2[kMaxA], c_logs[kMaxA], c_invs2[kMaxA], c_invs3[kMaxA];
  if (threadIdx.x < kMaxA) {
    dstd_s[threadIdx.x] = 0.0f;
    if (threadIdx.x < p.A) {
      const float s = p.std[threadIdx.x];
      c_s[threadIdx.x] = s;
      c_inv2s2[threadIdx.x] = 1.0f / (2.0f * s * s);
      c_logs[threadIdx.x] = logf(s);
      c_invs2[threadIdx.x] = 1.0f / (s * s);
      c_invs3[threadIdx.x] = 1.0f / (s * s * s);
    }
  }
  __syncthreads();
  const int i = blockIdx.x * kLossThreads + threadIdx.x;
  const float invM = 1.0f / (float)p.M;
  float part[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  float dsig[kMaxA];
#pragma unroll
  for (int a = 0; a < kMaxA; ++a) dsig[a] = 0.0f;
  if (i < p.M) {
    // log-prob and entropy of Normal(mu, std) (torch.distributions.Normal)
    float d[kMaxA];
    float lp = 0.0f, ent = 0.0f;
#pragma unroll
    for (int a = 0; a < kMaxA; ++a) {
      if (a < p.A) {
        d[a] = p.actions[(int64_t)i * p.A + a] - p.mu[(int64_t)i * p.ldmu + a];
        lp += -(d[a] * d[a]) * c_inv2s2[a] - c_logs[a] - 0.9189385332046727f;
        ent += 1.4189385332046727f + c_logs[a];
      }
    }
    const float adv = p.adv[i];
    const float ratio = expf(lp - p.old_logp[i]);
    const float s1 = -adv * ratio;
    const float rc = fminf(fmaxf(ratio, 1.0f - p.clip), 1.0f + p.clip);
    const float s2 = -adv * rc;
    part[0] = fmaxf(s1, s2);
    const bool inside = (ratio >= 1.0f - p.clip) && (ratio <= 1.0f + p.clip);
    // d max(s1,s2)/d logp: both branches carry -adv*ratio inside the clip range; outside only s1 does
    const float dlp = (inside || s1 > s2) ? (-adv * ratio) * invM : 0.0f;
    const float dsig_entropy = -p.entropy_coef * invM;
#pragma unroll
    for (int a = 0; a < kMaxA; ++a) {
      if (a < p.A) {
        p.dmu[(int64_t)i * p.lddmu + a] = dlp * d[a] * c_invs2[a];
        dsig[a] = dlp * (d[a] * d[a] - c_s[a] * c_s[a]) * c_invs3[a] + dsig_entropy / c_s[a];
      }
    }
    // value loss (ppo.py:256-264)
    const float v = p.value[(int64_t)i * p.ldv], R = p.returns[i];
    float dv;
    if (p.use_clipped_value_loss) {
      const float tv = p.target_values[i];
      const float dvt = v - tv;
      const float vc = tv + fminf(fmaxf(dvt, -p.clip), p.clip);
      const float l1 = (v - R) * (v - R), l2 = (vc - R) * (vc - R);
      part[1] = fmaxf(l1, l2);
      const bool in_v = (dvt >= -p.clip) && (dvt <= p.clip);
      const float d1 = 2.0f * (v - R), d2 = in_v ? 2.0f * (vc - R) : 0.0f;
      dv = l1 > l2 ? d1 : (l1 < l2 ? d2 : 0.5f * (d1 + d2));
    } else {
      part[1] = (R - v) * (R - v);
      dv = 2.0f * (v - R);
    }
    p.dvalue[(int64_t)i * p.lddv] = p.value_coef * dv * invM;
    // ROA regulariser: mean ||latent_p - latent_a||_2 (ppo.py:216)
    float n2 = 0.0f;
    const float reg_coef = p.reg_coef_dev ? p.reg_coef_dev[0] : p.reg_coef;
    const float* latp = p.latent_p + (int64_t)i * p.ldlp;
    const float* lata = p.latent_a + (int64_t)i * p.ldla;
    float* dl = p.dlatent_p + (int64_t)i * p.lddlp;
    constexpr int kL = 20;                                 // the go2 latent width: 5 x 16-byte vectors per row, loaded once
    if (p.L == kL && (p.ldlp % 4 == 0) && (p.ldla % 4 == 0) && (p.lddlp % 4 == 0) &&
        ((((uintptr_t)p.latent_p) | ((uintptr_t)p.latent_a) | ((uintptr_t)p.dlatent_p)) & 15) == 0) {
      float4 e4[kL / 4];
#pragma unroll
      for (int l = 0; l < kL / 4; ++l) {
        const float4 a4 = reinterpret_cast<const float4*>(latp)[l], b4 = reinterpret_cast<const float4*>(lata)[l];
        e4[l] = make_float4(a4.x - b4.x, a4.y - b4.y, a4.z - b4.z, a4.w - b4.w);
      }
#pragma unroll
      for (int l = 0; l < kL / 4; ++l) {                   // same summation order as the scalar loop
        n2 += e4[l].x * e4[l].x;
        n2 += e4[l].y * e4[l].y;
        n2 += e4[l].z * e4[l].z;
        n2 += e4[l].w * e4[l].w;
      }
      const float nrm = sqrtf(n2);
      part[2] = nrm;
      const float g = nrm > 0.0f ? reg_coef * invM / nrm : 0.0f;
#pragma unroll
      for (int l = 0; l < kL / 4; ++l) reinterpret_cast<float4*>(dl)[l] = make_float4(g * e4[l].x, g * e4[l].y, g * e4[l].z, g * e4[l].w);
    } else {
      for (int l = 0; l < p.L; ++l) {
        const float e = latp[l] - lata[l];
        n2 += e * e;
      }
      const float nrm = sqrtf(n2);
      part[2] = nrm;
      const float g = nrm > 0.0f ? reg_coef * invM / nrm : 0.0f;
      for (int l = 0; l < p.L; ++l) dl[l] = g * (latp[l] - lata[l]);
    }
    part[3] = ent;
  }
  // d(loss)/d(std): warp shuffle reduction, one shared-memory atomic per warp and action
#pragma unroll
  for (int a = 0; a < kMaxA; ++a) {
    if (a < p.A) {
      const float t = warp_sum_f(dsig[a]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&dstd_s[a], t);
    }
  }
  cta_sum<4>(part, red);
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 0; k < 4; ++k) atomicAdd(p.sums + k, part[k]);
  if (threadIdx.x < p.A) atomicAdd(p.dstd + threadIdx.x, dstd_s[threadIdx.x]);
}

// loss = mean_i ||pred_i - target_i||_2^2 ; dpred = 2 (pred - target) / M      (estimator, ppo.py:224-226)
__global__ void __launch_bounds__(256)
mse_rows_loss_kernel(const float* __restrict__ pred, int ldp, const float* __restrict__ target, int ldt, float* __restrict__ dpred,
                     int lddp, float* __restrict__ sum, int M, int D) {
  __shared__ float red[32];
  const int i = blockIdx.x * 256 + threadIdx.x;
  float part[1] = {0.0f};
  if (i < M) {
    for (int d = 0; d < D; ++d) {
      const float e = pred[(int64_t)i * ldp + d] - target[(int64_t)i * ldt + d];
      part[0] += e * e;
      dpred[(int64_t)i * lddp + d] = 2.0f * e / (float)M;
    }
  }
  cta_sum<1>(part, red);
  if (threadIdx.x == 0) atomicAdd(sum, part[0]);
}

// loss = mean_i ||target_i - pred_i||_2 ; dpred = -(target - pred) / (||.|| M)   (DAgger, ppo.py:330-333)
__global__ void __launch_bounds__(256)
l2_rows_loss_kernel(const float* __restrict__ pred, int ldp, const float* __restrict__ target, int ldt, float* __restrict__ dpred,
                    int lddp, float* __restrict__ sum, int M, int D) {
  __shared__ float red[32];
  const int i = blockIdx.x * 256 + threadIdx.x;
  float part[1] = {0.0f};
  if (i < M) {
    float n2 = 0.0f;
    for (int d = 0; d < D; ++d) {
      const float e = target[(int64_t)i * ldt + d] - pred[(int64_t)i * ldp + d];
      n2 += e * e;
    }
    const float nrm = sqrtf(n2);
    part[0] = nrm;
    const float g = nrm > 0.0f ? 1.0f / (nrm * (float)M) : 0.0f;
    for (int d = 0; d < D; ++d) dpred[(int64_t)i * lddp + d] = -g * (target[(int64_t)i * ldt + d] - pred[(int64_t)i * ldp + d]);
  }
  cta_sum<1>(part, red);
  if (threadIdx.x == 0) atomicAdd(sum, part[0]);
}

// dY *= elu'(Y) in place, for heads whose last layer has an activation (AdaptationEncoder.fc_final)
__global__ void __launch_bounds__(256)
elu_backward_kernel(float* __restrict__ dY, int lddy, const float* __restrict__ Y, int ldy, int M, int N) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  const int64_t r = i / N;
  const int c = (int)(i % N);
  const float y = Y[r * ldy + c];
  dY[r * lddy + c] *= (y > 0.0f ? 1.0f : y + 1.0f);
}

// Optimiser state block (device, doubles) so that a captured CUDA graph can be replayed step after step:
//   [0] sum of squared gradients (scratch)  [1] step  [2] beta1^step  [3] beta2^step  [4] lr
// schedule == 'adaptive' (ppo.py:233-246).  Per sample, in the reference's fp32 op order:
//   kl = sum_a [ log(sigma / old_sigma + 1e-5) + (old_sigma^2 + (old_mu - mu)^2) / (2 sigma^2) - 0.5 ]
// `acc[0]` accumulates the sum over the samples of this rank (fp64 atomics: the order of arrival moves it by ~1e-16 relative).
__global__ void __launch_bounds__(256)
kl_sum_kernel(const float* __restrict__ mu, int ldmu, const float* __restrict__ std, const float* __restrict__ old_mu, int ldom,
              const float* __restrict__ old_sigma, int ldos, int M, int A, double* __restrict__ acc) {
  __shared__ float red[32];
  const int i = blockIdx.x * 256 + threadIdx.x;
  float part[1] = {0.0f};
  if (i < M) {
    float kl = 0.0f;
    for (int a = 0; a < A; ++a) {
      const float sg = std[a], os = old_sigma[(int64_t)i * ldos + a];
      const float d = old_mu[(int64_t)i * ldom + a] - mu[(int64_t)i * ldmu + a];
      kl += (logf(sg / os + 1.e-5f) + (os * os + d * d) / (2.0f * (sg * sg))) - 0.5f;
    }
    part[0] = kl;
  }
  cta_sum<1>(part, red);
  if (threadIdx.x == 0) atomicAdd(acc, (double)part[0]);
}

// acc[0] = KL sum over `count` samples (all ranks) -> kl_mean (fp32, like the reference's tensor) -> the rule on the Python
// float `learning_rate` (fp64 here) -> adam_state[4]; acc[1] keeps kl_mean for the host, acc[0] is cleared for the next minibatch.
__global__ void adaptive_lr_kernel(double* __restrict__ acc, double count, float hi, float lo, double* __restrict__ adam_state) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float kl_mean = (float)(acc[0] / count);
  double lr = adam_state[4];
  if (kl_mean > hi) lr = fmax(1e-5, lr / 1.5);
  else if (kl_mean < lo && kl_mean > 0.0f) lr = fmin(1e-2, lr * 1.5);
  adam_state[4] = lr;
  acc[1] = (double)kl_mean;
  acc[0] = 0.0;
}

__global__ void adam_advance_kernel(double* __restrict__ state, double beta1, double beta2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    state[0] = 0.0;
    state[1] += 1.0;
    state[2] *= beta1;
    state[3] *= beta2;
    state[5] = state[7];      // stale squared norm covered by this step's clip (see clip_adam_kernel)
  }
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ state) {
  __shared__ float red[32];
  float part[1] = {0.0f};
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) part[0] += g[i] * g[i];
  cta_sum<1>(part, red);
  if (threadIdx.x == 0) atomicAdd(state, (double)part[0]);
}

// torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (default, non-amsgrad, weight_decay 0) on flat buffers.
// `grad_scale` pre-multiplies the gradient (1/world_size after an NCCL sum all-reduce); the gradient
// buffer is zeroed for the next accumulation.
__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                 double* state, float grad_scale, float max_norm, float beta1, float beta2, float eps) {
  // state[5]: squared norm of gradients OUTSIDE this buffer that the same clip_grad_norm_ call covers and rescales --
  // ppo.py:274 clips actor_critic.parameters(), which includes the adaptation encoder's stale .grad left (post-clip) by the
  // last update_dagger; optimizer.zero_grad() never clears it.  0 unless the caller seeded state[7].
  const double extra = state[5];
  const float own_norm = (float)sqrt(state[0]) * grad_scale;
  const float total_norm = extra > 0.0 ? (float)sqrt((double)own_norm * (double)own_norm + extra) : own_norm;
  float coef = max_norm / (total_norm + 1e-6f);
  coef = coef > 1.0f ? 1.0f : coef;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double* st_out = state;
    st_out[6] = (double)own_norm * (double)own_norm * (double)coef * (double)coef;   // this buffer's post-clip squared norm
    st_out[7] = extra * (double)coef * (double)coef;                                  // the stale gradients shrink with the clip
  }
  coef *= grad_scale;
  const float bc1 = (float)(1.0 - state[2]), bc2_sqrt = (float)sqrt(1.0 - state[3]);
  const float step_size = (float)state[4] / bc1;
  auto step = [&](float& pi, float& gi_, float& mi_, float& vi_) {
    const float gi = gi_ * coef;
    const float mi = mi_ + (gi - mi_) * (1.0f - beta1);          // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = vi_ * beta2 + (1.0f - beta2) * gi * gi;
    mi_ = mi;
    vi_ = vi;
    pi = pi - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    gi_ = 0.0f;
  };
  if ((n & 3) == 0 && ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0) {      // flat buffers: always
    float4 *p4 = reinterpret_cast<float4*>(p), *g4 = reinterpret_cast<float4*>(g), *m4 = reinterpret_cast<float4*>(m),
           *v4 = reinterpret_cast<float4*>(v);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n / 4; i += (int64_t)gridDim.x * 256) {
      float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
      step(pp.x, gg.x, mm.x, vv.x);
      step(pp.y, gg.y, mm.y, vv.y);
      step(pp.z, gg.z, mm.z, vv.z);
      step(pp.w, gg.w, mm.w, vv.w);
      p4[i] = pp; g4[i] = gg; m4[i] = mm; v4[i] = vv;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) step(p[i], g[i], m[i], v[i]);
  }
}

// ---- fused AdaptationEncoder forward (support_networks.py:128-175): Linear(52,30)+ELU on each of the 10 history
// steps -> Conv1d(30,20,k4,s2)+ELU -> Conv1d(20,10,k2)+ELU -> Flatten -> Linear(30,20)+ELU, all in fp32 FMAs.
// 27 k MAC per sample and 5 k parameters: too small for tensor-core tiles (it was 18 GEMM launches), so one CTA
// keeps all weights in shared memory and walks SAMPLES rows through the four stages.  Weight layouts are the
// kernel layouts of networks.py: W1 [30][52], W2 [20][4*32] (k*32+ci), W3 [10][2*20] (k*20+ci), W4 [20][3*12] (t*12+c).
// With save != 0 the intermediate activations are also written in the [M,320] / [M,80] / [M,36] layouts the
// GEMM-decomposed backward (DAgger) consumes.
struct AdaptArgs {
  const float* X; int ldx;
  const float *W1, *b1, *W2, *b2, *W3, *b3, *W4, *b4;
  float* out; int ldo;
  float *proj, *c1, *c2;      // optional [M,320], [M,80], [M,36]
  int M;
};
constexpr int AD_S = 32;       // samples per CTA pass
constexpr int AD_THREADS = AD_S * 10;   // stage 1 has one thread per (sample, history step): exactly one pass
constexpr int AD_W1 = 52, AD_W2 = 120, AD_W3 = 40, AD_W4 = 32;     // weight row pitches (floats, 16-byte rows)
constexpr int AD_SMEM_FLOATS = 30 * AD_W1 + 20 * AD_W2 + 10 * AD_W3 + 20 * AD_W4 + 32 + 20 + 12 + 20 + AD_S * 10 * 31 + AD_S * 4 * 21 +
                               AD_S * 3 * 11;

__device__ __forceinline__ float elu1(float x) { return x > 0.0f ? x : expf(x) - 1.0f; }

// Every stage is register-tiled the same way: a thread owns one output POSITION (sample, time step) with ALL its output
// channels, keeps the position's inputs in registers and streams the weight rows as 16-byte shared-memory BROADCASTS (all
// lanes read the same address), so the inner loops are 4 FMAs per shared-memory instruction instead of 1.
__global__ void __launch_bounds__(AD_THREADS, 2) adapt_forward_kernel(const __grid_constant__ AdaptArgs a) {
  extern __shared__ __align__(16) float ad_smem[];
  float* W1 = ad_smem;                    // [30][52]
  float* W2 = W1 + 30 * AD_W1;            // [20][4 taps x 30]
  float* W3 = W2 + 20 * AD_W2;            // [10][2 taps x 20]
  float* W4 = W3 + 10 * AD_W3;            // [20][32] (3 x 10 used, zero padded)
  float* B1 = W4 + 20 * AD_W4;            // 32
  float* B2 = B1 + 32;                    // 20
  float* B3 = B2 + 20;                    // 12
  float* B4 = B3 + 12;                    // 20
  float* proj = B4 + 20;                  // [AD_S][10][31]
  float* c1 = proj + AD_S * 10 * 31;      // [AD_S][4][21]
  float* c2 = c1 + AD_S * 4 * 21;         // [AD_S][3][11]
  const int tid = threadIdx.x;
  for (int i = tid; i < 30 * 52; i += AD_THREADS) W1[i] = a.W1[i];
  for (int i = tid; i < 20 * 120; i += AD_THREADS) {           // drop the two zero-pad channels of each tap
    const int co = i / 120, r = i % 120, k = r / 30, ci = r % 30;
    W2[i] = a.W2[co * 128 + k * 32 + ci];
  }
  for (int i = tid; i < 10 * 40; i += AD_THREADS) W3[i] = a.W3[i];
  for (int i = tid; i < 20 * 32; i += AD_THREADS) {
    const int o = i / 32, r = i % 32, t = r / 10, c = r % 10;
    W4[i] = r < 30 ? a.W4[o * 36 + t * 12 + c] : 0.0f;
  }
  if (tid < 30) B1[tid] = a.b1[tid];
  if (tid < 20) { B2[tid] = a.b2[tid]; B4[tid] = a.b4[tid]; }
  if (tid < 10) B3[tid] = a.b3[tid];
  __syncthreads();
  for (int row0 = blockIdx.x * AD_S; row0 < a.M; row0 += gridDim.x * AD_S) {
    const int ns = min(AD_S, a.M - row0);
    // stage 1: Linear(52 -> 30) + ELU on one (sample, step): the 52 inputs sit in registers
    for (int idx = tid; idx < ns * 10; idx += AD_THREADS) {
      const int s = idx / 10, t = idx % 10;
      const float4* xp = reinterpret_cast<const float4*>(a.X + (int64_t)(row0 + s) * a.ldx + 52 * t);
      float4 x[13];
#pragma unroll
      for (int i = 0; i < 13; ++i) x[i] = __ldg(xp + i);
      float* pr = proj + (s * 10 + t) * 31;
#pragma unroll 3
      for (int c = 0; c < 30; ++c) {
        const float4* w = reinterpret_cast<const float4*>(W1 + c * AD_W1);
        float acc0 = B1[c], acc1 = 0.0f;
#pragma unroll
        for (int i = 0; i < 13; ++i) {
          const float4 w4 = w[i];
          acc0 = fmaf(x[i].x, w4.x, acc0);
          acc1 = fmaf(x[i].y, w4.y, acc1);
          acc0 = fmaf(x[i].z, w4.z, acc0);
          acc1 = fmaf(x[i].w, w4.w, acc1);
        }
        const float y = elu1(acc0 + acc1);
        pr[c] = y;
        if (a.proj) a.proj[(int64_t)(row0 + s) * 320 + 32 * t + c] = y;
      }
    }
    __syncthreads();
    // stage 2: Conv1d(30 -> 20, k = 4, stride 2) + ELU on one (sample, t'): 20 accumulators, one tap's 30 inputs at a time
    for (int idx = tid; idx < ns * 4; idx += AD_THREADS) {
      const int s = idx >> 2, tp = idx & 3;
      float acc[20];
#pragma unroll
      for (int co = 0; co < 20; ++co) acc[co] = B2[co];
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        const float* p = proj + (s * 10 + 2 * tp + k) * 31;
        float in[32];
#pragma unroll
        for (int ci = 0; ci < 30; ++ci) in[ci] = p[ci];
        in[30] = in[31] = 0.0f;
#pragma unroll
        for (int co = 0; co < 20; ++co) {
          const float* wr = W2 + co * AD_W2 + k * 30;     // 30 floats per tap: 8-byte aligned rows -> float2 broadcasts
          const float2* w2 = reinterpret_cast<const float2*>(wr);
          float accl = acc[co];
#pragma unroll
          for (int i = 0; i < 15; ++i) {
            const float2 ww = w2[i];
            accl = fmaf(in[2 * i], ww.x, accl);
            accl = fmaf(in[2 * i + 1], ww.y, accl);
          }
          acc[co] = accl;
        }
      }
      float* o = c1 + (s * 4 + tp) * 21;
#pragma unroll
      for (int co = 0; co < 20; ++co) {
        const float y = elu1(acc[co]);
        o[co] = y;
        if (a.c1) a.c1[(int64_t)(row0 + s) * 80 + 20 * tp + co] = y;
      }
    }
    __syncthreads();
    // stage 3: Conv1d(20 -> 10, k = 2) + ELU on one (sample, t'')
    for (int idx = tid; idx < ns * 3; idx += AD_THREADS) {
      const int s = idx / 3, tp = idx % 3;
      float in[40];
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int ci = 0; ci < 20; ++ci) in[k * 20 + ci] = c1[(s * 4 + tp + k) * 21 + ci];
      float* o = c2 + (s * 3 + tp) * 11;
#pragma unroll 2
      for (int co = 0; co < 10; ++co) {
        const float4* w = reinterpret_cast<const float4*>(W3 + co * AD_W3);
        float acc = B3[co];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          const float4 w4 = w[i];
          acc = fmaf(in[4 * i], w4.x, acc);
          acc = fmaf(in[4 * i + 1], w4.y, acc);
          acc = fmaf(in[4 * i + 2], w4.z, acc);
          acc = fmaf(in[4 * i + 3], w4.w, acc);
        }
        const float y = elu1(acc);
        o[co] = y;
        if (a.c2) a.c2[(int64_t)(row0 + s) * 36 + 12 * tp + co] = y;
      }
    }
    __syncthreads();
    // stage 4: Flatten + Linear(30 -> 20) + ELU: (sample, o)
    for (int idx = tid; idx < ns * 20; idx += AD_THREADS) {
      const int s = idx / 20, o = idx % 20;
      float acc = B4[o];
      const float* w = W4 + o * AD_W4;
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int c = 0; c < 10; ++c) acc = fmaf(c2[(s * 3 + t) * 11 + c], w[t * 10 + c], acc);
      a.out[(int64_t)(row0 + s) * a.ldo + o] = elu1(acc);
    }
    __syncthreads();
  }
}

// db[n] += sum_m dY[m,n]: lane = column, warps stride over a slab of rows, one atomic per column per CTA.
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ dY, int lddy, float* __restrict__ db, int M, int N, int rows_per_cta) {
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float s = 0.0f;
  if (col < N)
    for (int r = r0 + warp; r < r1; r += 8) s += dY[(int64_t)r * lddy + col];
  part[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && col < N) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][lane];
    atomicAdd(db + col, t);
  }
}

__global__ void fill_kernel(float* __restrict__ p, float value, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = value;
}

extern "C" {

int b200_copy_segments(const B200CopySeg* segs, int nseg, int rows, void* stream) {
  B200_CHECK_ARG(segs && nseg > 0 && nseg <= 8 && rows > 0, "b200_copy_segments: need 1..8 segments and rows > 0");
  CopyArgs a{};
  a.nseg = nseg;
  a.rows = rows;
  int64_t maxw = 0;
  for (int i = 0; i < nseg; ++i) {
    B200_CHECK_ARG(segs[i].src && segs[i].dst && segs[i].width > 0 && segs[i].src_ld >= segs[i].width && segs[i].dst_ld >= segs[i].width,
                   "b200_copy_segments: bad segment %d", i);
    a.seg[i] = segs[i];
    maxw = segs[i].width > maxw ? segs[i].width : maxw;
  }
  int64_t work = ((int64_t)rows * maxw / 4 + 255) / 256;
  int bx = (int)(work < 1 ? 1 : (work > 148 * 8 ? 148 * 8 : work));
  copy_segments_kernel<<<dim3(bx, nseg), 256, 0, (cudaStream_t)stream>>>(a);
  B200_CHECK_LAUNCH("copy_segments_kernel");
  return 0;
}

int b200_gather_rows(const float* src, int src_ld, const int64_t* idx, float* dst, int dst_ld, int width, int64_t rows, void* stream) {
  B200_CHECK_ARG(src && idx && dst && width > 0 && rows > 0 && src_ld >= width && dst_ld >= width, "b200_gather_rows: bad argument");
  gather_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(src, src_ld, idx, dst, dst_ld, width, rows);
  B200_CHECK_LAUNCH("gather_rows_kernel");
  return 0;
}

int b200_gather_bytes(const uint8_t* src, const int64_t* idx, uint8_t* dst, int64_t rows, void* stream) {
  B200_CHECK_ARG(src && idx && dst && rows > 0, "b200_gather_bytes: bad argument");
  gather_bytes_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, idx, dst, rows);
  B200_CHECK_LAUNCH("gather_bytes_kernel");
  return 0;
}

int b200_sample_actions(const float* mu, int ldmu, const float* std, uint64_t seed, int64_t step, float* actions, float* logp,
                        float* mu_out, float* sigma_out, int N, int A, void* stream) {
  B200_CHECK_ARG(mu && std && actions && logp && N > 0 && A > 0 && A <= 16 && ldmu >= A, "b200_sample_actions: bad argument");
  sample_actions_kernel<<<(N + kSampleEnvsPerCta - 1) / kSampleEnvsPerCta, 256, 0, (cudaStream_t)stream>>>(
      mu, ldmu, std, seed, (uint32_t)step, nullptr, actions, logp, mu_out, sigma_out, N, A);
  B200_CHECK_LAUNCH("sample_actions_kernel");
  return 0;
}

// same, with the noise step counter in device memory: step_counter_dev[0] is read, then incremented by the kernel's last CTA;
// step_counter_dev[1] is the kernel's ticket word (zero-initialised by the caller, left at zero by every launch)
int b200_sample_actions_dev(const float* mu, int ldmu, const float* std, uint64_t seed, int64_t* step_counter_dev, float* actions,
                            float* logp, float* mu_out, float* sigma_out, int N, int A, void* stream) {
  B200_CHECK_ARG(mu && std && actions && logp && step_counter_dev && N > 0 && A > 0 && A <= 16 && ldmu >= A, "b200_sample_actions_dev: bad argument");
  sample_actions_kernel<<<(N + kSampleEnvsPerCta - 1) / kSampleEnvsPerCta, 256, 0, (cudaStream_t)stream>>>(
      mu, ldmu, std, seed, 0, step_counter_dev, actions, logp, mu_out, sigma_out, N, A);
  B200_CHECK_LAUNCH("sample_actions_kernel");
  return 0;
}

int b200_ppo_loss(const PpoLossArgs* a, void* stream) {
  B200_CHECK_ARG(a && a->M > 0 && a->A > 0 && a->A <= 16 && a->L > 0, "b200_ppo_loss: bad sizes");
  B200_CHECK_ARG(a->mu && a->std && a->actions && a->old_logp && a->adv && a->returns && a->target_values && a->value && a->latent_p &&
                     a->latent_a && a->dmu && a->dvalue && a->dlatent_p && a->dstd && a->sums,
                 "b200_ppo_loss: null pointer");
  ppo_loss_kernel<<<(a->M + kLossThreads - 1) / kLossThreads, kLossThreads, 0, (cudaStream_t)stream>>>(*a);
  B200_CHECK_LAUNCH("ppo_loss_kernel");
  return 0;
}

int b200_mse_rows_loss(const float* pred, int ldp, const float* target, int ldt, float* dpred, int lddp, float* sum, int M, int D,
                       void* stream) {
  B200_CHECK_ARG(pred && target && dpred && sum && M > 0 && D > 0, "b200_mse_rows_loss: bad argument");
  mse_rows_loss_kernel<<<(M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pred, ldp, target, ldt, dpred, lddp, sum, M, D);
  B200_CHECK_LAUNCH("mse_rows_loss_kernel");
  return 0;
}

int b200_l2_rows_loss(const float* pred, int ldp, const float* target, int ldt, float* dpred, int lddp, float* sum, int M, int D,
                      void* stream) {
  B200_CHECK_ARG(pred && target && dpred && sum && M > 0 && D > 0, "b200_l2_rows_loss: bad argument");
  l2_rows_loss_kernel<<<(M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pred, ldp, target, ldt, dpred, lddp, sum, M, D);
  B200_CHECK_LAUNCH("l2_rows_loss_kernel");
  return 0;
}

int b200_elu_backward(float* dY, int lddy, const float* Y, int ldy, int M, int N, void* stream) {
  B200_CHECK_ARG(dY && Y && M > 0 && N > 0, "b200_elu_backward: bad argument");
  const int64_t n = (int64_t)M * N;
  elu_backward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dY, lddy, Y, ldy, M, N);
  B200_CHECK_LAUNCH("elu_backward_kernel");
  return 0;
}

int b200_clip_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double* state, float grad_scale,
                   float max_norm, float beta1, float beta2, float eps, void* stream) {
  B200_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && state && n > 0, "b200_clip_adam: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  adam_advance_kernel<<<1, 32, 0, st>>>(state, (double)beta1, (double)beta2);
  int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
  blocks = blocks < 1 ? 1 : (blocks > 148 * 4 ? 148 * 4 : blocks);
  sumsq_kernel<<<blocks, 256, 0, st>>>(grads, n, state);
  B200_CHECK_LAUNCH("sumsq_kernel");
  clip_adam_kernel<<<blocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, state, grad_scale, max_norm, beta1, beta2, eps);
  B200_CHECK_LAUNCH("clip_adam_kernel");
  return 0;
}

int b200_kl_sum(const float* mu, int ldmu, const float* std, const float* old_mu, int ldom, const float* old_sigma, int ldos, int M, int A,
                double* acc, void* stream) {
  B200_CHECK_ARG(mu && std && old_mu && old_sigma && acc && M > 0 && A > 0, "b200_kl_sum: bad argument");
  kl_sum_kernel<<<(M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mu, ldmu, std, old_mu, ldom, old_sigma, ldos, M, A, acc);
  B200_CHECK_LAUNCH("kl_sum_kernel");
  return 0;
}

int b200_adaptive_lr(double* acc, int64_t count, double desired_kl, double* adam_state, void* stream) {
  B200_CHECK_ARG(acc && adam_state && count > 0 && desired_kl > 0.0, "b200_adaptive_lr: bad argument");
  // the reference compares an fp32 tensor with Python floats: the thresholds take the tensor's dtype
  adaptive_lr_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc, (double)count, (float)(desired_kl * 2.0), (float)(desired_kl / 2.0), adam_state);
  B200_CHECK_LAUNCH("adaptive_lr_kernel");
  return 0;
}

int b200_adaptation_forward(const float* X, int ldx, const float* W1, const float* b1, const float* W2, const float* b2,
                            const float* W3, const float* b3, const float* W4, const float* b4, float* out, int ldo, float* proj,
                            float* c1, float* c2, int M, void* stream) {
  B200_CHECK_ARG(X && W1 && b1 && W2 && b2 && W3 && b3 && W4 && b4 && out && M > 0, "b200_adaptation_forward: null argument");
  B200_CHECK_ARG(ldx % 4 == 0 && ldx >= 520 && (((uintptr_t)X) & 15) == 0, "b200_adaptation_forward: X rows must be 16-byte aligned, >= 520 wide");
  AdaptArgs a{X, ldx, W1, b1, W2, b2, W3, b3, W4, b4, out, ldo, proj, c1, c2, M};
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(adapt_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(AD_SMEM_FLOATS * sizeof(float)));
    if (e != cudaSuccess) {
      b200_set_error("adapt_forward_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_done = true;
  }
  int blocks = (M + AD_S - 1) / AD_S;
  blocks = blocks > 148 * 2 ? 148 * 2 : blocks;
  adapt_forward_kernel<<<blocks, AD_THREADS, AD_SMEM_FLOATS * sizeof(float), (cudaStream_t)stream>>>(a);
  B200_CHECK_LAUNCH("adapt_forward_kernel");
  return 0;
}

int b200_colsum(const float* dY, int lddy, float* db, int M, int N, void* stream) {
  B200_CHECK_ARG(dY && db && M > 0 && N > 0 && lddy >= N, "b200_colsum: bad argument");
  const int col_blocks = (N + 31) / 32;
  int row_blocks = (148 * 4 + col_blocks - 1) / col_blocks;
  const int max_rb = (M + 63) / 64;
  row_blocks = row_blocks < 1 ? 1 : (row_blocks > max_rb ? max_rb : row_blocks);
  const int rows_per_cta = (M + row_blocks - 1) / row_blocks;
  colsum_kernel<<<dim3(col_blocks, (M + rows_per_cta - 1) / rows_per_cta), 256, 0, (cudaStream_t)stream>>>(dY, lddy, db, M, N, rows_per_cta);
  B200_CHECK_LAUNCH("colsum_kernel");
  return 0;
}

int b200_fill(float* p, float value, int64_t n, void* stream) {
  B200_CHECK_ARG(p && n > 0, "b200_fill: bad argument");
  int blocks = (int)((n + 255) / 256);
  blocks = blocks > 148 * 8 ? 148 * 8 : blocks;
  fill_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, value, n);
  B200_CHECK_LAUNCH("fill_kernel");
  return 0;
}

}  // extern "C"
