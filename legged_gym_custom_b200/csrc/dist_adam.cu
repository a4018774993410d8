// K8 -- the ONE collective of the data-parallel path, fused with what follows it (ppo.py:228-231,
// :273-276, :336-339): gradient exchange + global-norm clip + Adam, one kernel per optimiser step, over NVLink / NVSwitch
// peer memory.  Replaces  all_reduce(grads)  ->  sumsq  ->  clip + Adam on the full buffer  on every rank.
//
// Every rank launches the same kernel (inside its CUDA graph).  Gradients, parameters and a small sync area live in
// SYMMETRIC memory (same allocation on every rank, mapped into every rank's address space; with NVSwitch multicast also
// behind one multicast address).  Rank r owns shard r = [r n / W, (r + 1) n / W) of the flat buffers:
//
//   0. start barrier   every rank's gradients are complete            flags (st.release.sys on the peers, ld.acquire.sys)
//   1. reduce-scatter  g[shard] = sum over ranks of grads[shard]      multimem.ld_reduce (the SWITCH adds) or W peer loads
//                      + this shard's sum of squares
//   2. norm exchange   W doubles, every rank gets every partial       plain peer stores + flags; total in fixed rank order
//   3. clip + Adam     on the shard only (1 / W of the work); moments stay sharded
//   4. all-gather      the updated PARAMETERS go to every rank, the   multimem.st (one store, the switch replicates) or
//                      shard of every rank's gradient buffer is zeroed W peer stores
//   5. end barrier     every rank's copy of the parameters is complete
//
// Only the owner ever reads a shard's gradients and it broadcasts ONE result, so the replicas stay bit-identical by
// construction.  The optimiser state (step, beta powers, lr, stale-norm slots; b200_clip_adam) is advanced by the kernel
// itself and the barrier epoch lives in device memory: graph replays need no host involvement.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxWorld = 16;

struct DistAdamArgs {
  float* grads;                 // local addresses of the symmetric buffers
  float* params;
  float* grads_peer[kMaxWorld];  // the same buffers of every rank (own rank included), peer-mapped
  float* params_peer[kMaxWorld];
  unsigned long long* sync_peer[kMaxWorld];   // sync area of every rank: see kSync* below
  float* grads_mc;              // multicast addresses (nullptr: no NVSwitch multicast -> peer loops)
  float* params_mc;
  float* exp_avg;               // local; only this rank's shard is maintained
  float* exp_avg_sq;
  float* gsum;                  // local scratch, one shard of reduced gradients
  double* state;                // b200_clip_adam's 8 doubles
  unsigned int* local;          // local scratch: [0] barrier epoch, [1] / [2] grid counters, [4..5] = double sum of squares
  long long n;                  // floats in the flat buffers (multiple of 4)
  int world, rank;
  float max_norm, beta1, beta2, eps;
};

// sync area (unsigned long long units): [0, W) start flags | [W, 2W) partial flags | [2W, 3W) end flags | [3W, 4W) partials
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_v4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 multimem_ld_reduce_add_v4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_v4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// every block waits until all W flags of `mine` carry this launch's epoch
__device__ __forceinline__ void wait_flags(const unsigned long long* mine, int world, unsigned long long epoch) {
  if (threadIdx.x < world) {
    while (ld_acquire_sys(mine + threadIdx.x) < epoch) {
    }
  }
  __syncthreads();
}

template <bool MC>
__global__ void __launch_bounds__(kThreads) dist_adam_kernel(const __grid_constant__ DistAdamArgs a) {
  __shared__ float red[32];
  __shared__ int is_last;
  const int W = a.world, r = a.rank, t = threadIdx.x;
  unsigned long long* mine = a.sync_peer[r];
  const unsigned long long epoch = (unsigned long long)a.local[0] + 1ull;      // advanced by the last block at the very end
  // optimiser state of THIS step (adam_advance_kernel's arithmetic), read before anybody writes it back
  const double step_old = a.state[1], b1p = a.state[2] * (double)a.beta1, b2p = a.state[3] * (double)a.beta2, lr = a.state[4];
  const double extra = a.state[7];
  const long long n4 = a.n / 4;
  const long long per = (n4 + W - 1) / W;
  const long long lo = per * r, hi = (lo + per < n4) ? lo + per : n4;          // this rank's shard, in float4 units

  // ---- 0. start barrier: my gradients are complete (kernel boundary) -> tell everyone; wait for everyone
  //         (everything earlier kernels of this stream wrote is already in this GPU's L2, the point of coherence for the
  //         peers' NVLink accesses: the flag needs no fence of its own)
  if (blockIdx.x == 0 && t < W) st_release_sys(a.sync_peer[t] + r, epoch);
  wait_flags(mine, W, epoch);

  // ---- 1. reduce-scatter + sum of squares of the shard (4 independent requests per thread in flight: the round trip through
  //         the switch is ~2 us)
  float part[1] = {0.0f};
  auto fetch = [&](long long i) {
    float4 g;
    if (MC) {
      g = multimem_ld_reduce_add_v4(a.grads_mc + 4 * i);
    } else {
      g = ld_relaxed_sys_v4(a.grads_peer[0] + 4 * i);
      for (int p = 1; p < W; ++p) {                     // fixed order: the one result is broadcast, replicas cannot diverge
        const float4 q = ld_relaxed_sys_v4(a.grads_peer[p] + 4 * i);
        g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
      }
    }
    return g;
  };
  auto keep = [&](long long i, const float4& g) {
    reinterpret_cast<float4*>(a.gsum)[i - lo] = g;
    part[0] += (g.x * g.x + g.y * g.y) + (g.z * g.z + g.w * g.w);
  };
  {
    const long long stride = (long long)gridDim.x * kThreads;
    long long i = lo + (long long)blockIdx.x * kThreads + t;
    for (; i + 3 * stride < hi; i += 4 * stride) {
      const float4 g0 = fetch(i), g1 = fetch(i + stride), g2 = fetch(i + 2 * stride), g3 = fetch(i + 3 * stride);
      keep(i, g0); keep(i + stride, g1); keep(i + 2 * stride, g2); keep(i + 3 * stride, g3);
    }
    for (; i < hi; i += stride) keep(i, fetch(i));
  }
  {   // CTA sum (fixed tree) -> one double atomic per block
    const int lane = t & 31, warp = t >> 5;
    float v = part[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
      v = lane < kThreads / 32 ? red[lane] : 0.0f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    double* acc = reinterpret_cast<double*>(a.local + 4);
    if (t == 0) {
      atomicAdd(acc, (double)v);
      __threadfence();
      is_last = atomicAdd(a.local + 1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    // ---- 2. norm exchange: the last block of this rank publishes the shard's partial to every rank
    if (is_last && t < W) {
      __threadfence();
      const double mine_sq = *reinterpret_cast<volatile double*>(acc);
      unsigned long long bits = (unsigned long long)__double_as_longlong(mine_sq);
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a.sync_peer[t] + 3 * W + r), "l"(bits) : "memory");
      st_release_sys(a.sync_peer[t] + W + r, epoch);      // release: the partial (same thread, same peer) is visible before the flag
    }
    if (is_last) {
      __syncthreads();
      if (t == 0) {
        *acc = 0.0;
        a.local[1] = 0u;
      }
    }
  }
  wait_flags(mine + W, W, epoch);
  // the flags were acquired (and the block synchronised) in wait_flags: the partials behind them are visible, so the W loads
  // go out together as relaxed loads instead of W dependent acquire round trips; summed in rank order: identical everywhere
  double partial[kMaxWorld];
#pragma unroll
  for (int p = 0; p < kMaxWorld; ++p) {
    unsigned long long bits = 0ull;
    if (p < W) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(bits) : "l"(mine + 3 * W + p) : "memory");
    partial[p] = __longlong_as_double((long long)bits);
  }
  double total_sq = 0.0;
#pragma unroll
  for (int p = 0; p < kMaxWorld; ++p)
    if (p < W) total_sq += partial[p];

  // ---- 3. clip coefficient (gradient = sum / W, b200_clip_adam's grad_scale) and Adam on the shard
  const float grad_scale = 1.0f / (float)W;
  const float own_norm = (float)sqrt(total_sq) * grad_scale;
  const float total_norm = extra > 0.0 ? (float)sqrt((double)own_norm * (double)own_norm + extra) : own_norm;
  float coef = a.max_norm / (total_norm + 1e-6f);
  coef = coef > 1.0f ? 1.0f : coef;
  const float clip = coef;
  coef *= grad_scale;
  const float bc1 = (float)(1.0 - b1p), bc2_sqrt = (float)sqrt(1.0 - b2p);
  const float step_size = (float)lr / bc1;
  const float beta1 = a.beta1, beta2 = a.beta2, eps = a.eps;
  auto step = [&](float& pi, float gi_, float& mi_, float& vi_) {
    const float gi = gi_ * coef;
    const float mi = mi_ + (gi - mi_) * (1.0f - beta1);
    const float vi = vi_ * beta2 + (1.0f - beta2) * gi * gi;
    mi_ = mi;
    vi_ = vi;
    pi = pi - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  };
  float4* p4 = reinterpret_cast<float4*>(a.params);
  float4* m4 = reinterpret_cast<float4*>(a.exp_avg);
  float4* v4 = reinterpret_cast<float4*>(a.exp_avg_sq);
  for (long long i = lo + (long long)blockIdx.x * kThreads + t; i < hi; i += (long long)gridDim.x * kThreads) {
    const float4 g = reinterpret_cast<const float4*>(a.gsum)[i - lo];
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    step(pp.x, g.x, mm.x, vv.x);
    step(pp.y, g.y, mm.y, vv.y);
    step(pp.z, g.z, mm.z, vv.z);
    step(pp.w, g.w, mm.w, vv.w);
    m4[i] = mm;
    v4[i] = vv;
    // ---- 4. all-gather of the parameters; the shard of every rank's gradient buffer is zeroed for the next accumulation
    //         (only its owner ever reads it, and every read of it is behind the grid-wide sync of phase 2)
    const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (MC) {
      multimem_st_v4(a.params_mc + 4 * i, pp);
      multimem_st_v4(a.grads_mc + 4 * i, z);
    } else {
      for (int p = 0; p < W; ++p) {
        st_relaxed_sys_v4(a.params_peer[p] + 4 * i, pp);
        st_relaxed_sys_v4(a.grads_peer[p] + 4 * i, z);
      }
    }
  }

  // ---- 5. end barrier: my stores are out (all blocks) -> tell everyone; the LAST block stays until everyone has told us
  __syncthreads();
  if (t == 0) {
    __threadfence_system();
    is_last = atomicAdd(a.local + 2, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  if (t < W) st_release_sys(a.sync_peer[t] + 2 * W + r, epoch);
  wait_flags(mine + 2 * W, W, epoch);
  if (t == 0) {
    a.local[2] = 0u;
    a.local[0] = (unsigned int)epoch;
    a.state[0] = total_sq;
    a.state[1] = step_old + 1.0;
    a.state[2] = b1p;
    a.state[3] = b2p;
    a.state[5] = extra;
    a.state[6] = (double)own_norm * (double)own_norm * (double)clip * (double)clip;
    a.state[7] = extra * (double)clip * (double)clip;
  }
}

}  // namespace

extern "C" {

// see include/b200gym.h
int b200_dist_adam(const B200DistAdam* d, void* stream) {
  B200_CHECK_ARG(d && d->grads && d->params && d->exp_avg && d->exp_avg_sq && d->gsum && d->state && d->local, "b200_dist_adam: null argument");
  B200_CHECK_ARG(d->world >= 1 && d->world <= kMaxWorld && d->rank >= 0 && d->rank < d->world, "b200_dist_adam: world %d rank %d", d->world, d->rank);
  B200_CHECK_ARG(d->n > 0 && d->n % 4 == 0, "b200_dist_adam: n must be a positive multiple of 4");
  B200_CHECK_ARG(d->grads_peer && d->params_peer && d->sync_peer, "b200_dist_adam: peer pointer tables are null");
  DistAdamArgs a{};
  a.grads = d->grads; a.params = d->params; a.grads_mc = d->grads_mc; a.params_mc = d->params_mc;
  a.exp_avg = d->exp_avg; a.exp_avg_sq = d->exp_avg_sq; a.gsum = d->gsum; a.state = d->state; a.local = d->local;
  a.n = d->n; a.world = d->world; a.rank = d->rank;
  a.max_norm = d->max_norm; a.beta1 = d->beta1; a.beta2 = d->beta2; a.eps = d->eps;
  for (int p = 0; p < d->world; ++p) {
    B200_CHECK_ARG(d->grads_peer[p] && d->params_peer[p] && d->sync_peer[p], "b200_dist_adam: peer %d is not mapped", p);
    a.grads_peer[p] = d->grads_peer[p];
    a.params_peer[p] = d->params_peer[p];
    a.sync_peer[p] = d->sync_peer[p];
  }
  // The ranks (and the blocks of one rank) wait for each other inside the kernel, so every block must become resident
  // without waiting for another block of the SAME kernel: at most 128 blocks of 256 threads and no shared memory to speak
  // of -- they fit beside any persistent one-CTA-per-SM GEMM that may be running on the side streams.
  const long long shard4 = (d->n / 4 + d->world - 1) / d->world;
  int blocks = (int)((shard4 + kThreads - 1) / kThreads);
  blocks = blocks < 1 ? 1 : (blocks > 128 ? 128 : blocks);
  const bool mc = d->grads_mc != nullptr && d->params_mc != nullptr;
  if (mc) dist_adam_kernel<true><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(a);
  else dist_adam_kernel<false><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(a);
  B200_CHECK_LAUNCH("dist_adam_kernel");
  return 0;
}

}  // extern "C"
