// Fused Linear-layer GEMMs for the actor / critic / estimator / encoders.
//
//   forward  (NT): Y[M,N]  = act(X[M,K] . W[N,K]^T + b)                 actor_critic.py:84-108, support_networks.py
//   dgrad    (NN): dX[M,K] (+)= (dY[M,N] . W[N,K]) * elu'(Yprev[M,K])
//   wgrad    (TN): dW[N,K] += dY[M,N]^T . X[M,K],  db[N] += colsum(dY)   (split over M, fp32 atomics)
//
// Tensor-core path: warp-level mma.sync.m16n8k8 TF32 with fp32 accumulation -- the precision
// the reference trains with on GPU (torch.set_float32_matmul_precision('high'), train.py:39).
// `precise` selects 3xTF32 error compensation (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi), which is
// fp32-accurate and is what the parity tests against the fp32 CPU oracle use.
// Operands stream through a 3-stage cp.async (LDGSTS) shared-memory pipeline; all three
// variants read 16-byte chunks along the contiguous dimension, so rows need 16 B alignment
// (ld % 4 == 0) -- the host pads 627/29/30-wide tensors to the next multiple of 4.
//
// This is the portable tensor-core baseline of the repo; csrc/mlp_tcgen05.cu carries the
// tcgen05/TMEM path for the wide layers.
#include "common.cuh"

namespace {

constexpr int BK = 32;
constexpr int STAGES = 3;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ unsigned f2tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

enum Mode { NT = 0, NN = 1, TN = 2 };

struct GemmArgs {
  const float* A;   // NT/NN: [M, red] rows (X or dY); TN: dY [Mred, Nout]
  const float* B;   // NT: W [N, K]; NN: W [Nred, Kout]; TN: X [Mred, Kout]
  float* C;         // NT: Y [M,N]; NN: dX [M,K]; TN: dW [N,K]
  const float* bias;    // NT: bias[N] or null
  const float* aux;     // NN: Yprev [M,K] (elu' applied) or null
  float* dbias;         // TN: db[N] or null
  int lda, ldb, ldc, ldaux;
  int M, N, K;          // output rows, output cols, reduction length
  int act;              // NT: 0 none, 1 ELU
  int accumulate;       // NN: 0 overwrite dX, 1 add to dX, n > 1 add to its first n columns
  int red_per_split;    // TN: reduction rows per blockIdx.z
};

// Tile of the output: BM x BN, reduction step BK.  Warp grid WM x WN, each warp (BM/WM) x (BN/WN).
// smem operand tiles:
//   "row" operand with the reduction dim contiguous : T[rows][BK + 4]        (NT: A and B; NN: A)
//   operand with the OUTPUT dim contiguous          : T[BK][cols + 8]        (NN: B; TN: A and B)
template <int MODE, int BM, int BN, int WM, int WN, bool PRECISE>
__global__ void __launch_bounds__(WM* WN * 32) gemm_kernel(const GemmArgs g) {
  constexpr int THREADS = WM * WN * 32;
  constexpr bool A_RED_CONTIG = (MODE != TN);
  constexpr bool B_RED_CONTIG = (MODE == NT);
  constexpr int A_ROWS = A_RED_CONTIG ? BM : BK, A_COLS = A_RED_CONTIG ? BK + 4 : BM + 8;
  constexpr int B_ROWS = B_RED_CONTIG ? BN : BK, B_COLS = B_RED_CONTIG ? BK + 4 : BN + 8;
  constexpr int TM = BM / WM, TN_ = BN / WN;   // warp tile
  constexpr int MI = TM / 16, NI = TN_ / 8;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = smem + STAGES * A_ROWS * A_COLS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WN, wn = warp % WN;
  const int gq = lane >> 2, tq = lane & 3;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int red_lo = 0, red_hi = g.K;
  if (MODE == TN) {
    red_lo = blockIdx.z * g.red_per_split;
    red_hi = min(g.K, red_lo + g.red_per_split);
  }
  const int ktiles = (red_hi - red_lo + BK - 1) / BK;

  auto load_stage = [&](int stage, int kt) {
    const int r0 = red_lo + kt * BK;
    float* as = As + stage * A_ROWS * A_COLS;
    float* bs = Bs + stage * B_ROWS * B_COLS;
    if (A_RED_CONTIG) {   // rows = output rows m, cols = reduction
      constexpr int CH = BK / 4;
      for (int i = tid; i < BM * CH; i += THREADS) {
        const int r = i / CH, c = (i % CH) * 4;
        const int gm = m0 + r, gk = r0 + c;
        int bytes = 0;
        if (gm < g.M && gk < red_hi) bytes = min(16, (red_hi - gk) * 4);
        cp_async16(as + r * A_COLS + c, bytes ? g.A + (int64_t)gm * g.lda + gk : g.A, bytes);
      }
    } else {              // TN: rows = reduction (m), cols = output rows (n of dW)
      constexpr int CH = BM / 4;
      for (int i = tid; i < BK * CH; i += THREADS) {
        const int r = i / CH, c = (i % CH) * 4;
        const int gr = r0 + r, gc = m0 + c;
        int bytes = 0;
        if (gr < red_hi && gc < g.M) bytes = min(16, (g.M - gc) * 4);
        cp_async16(as + r * A_COLS + c, bytes ? g.A + (int64_t)gr * g.lda + gc : g.A, bytes);
      }
    }
    if (B_RED_CONTIG) {   // NT: rows = output cols n, cols = reduction k
      constexpr int CH = BK / 4;
      for (int i = tid; i < BN * CH; i += THREADS) {
        const int r = i / CH, c = (i % CH) * 4;
        const int gn = n0 + r, gk = r0 + c;
        int bytes = 0;
        if (gn < g.N && gk < red_hi) bytes = min(16, (red_hi - gk) * 4);
        cp_async16(bs + r * B_COLS + c, bytes ? g.B + (int64_t)gn * g.ldb + gk : g.B, bytes);
      }
    } else {              // NN / TN: rows = reduction, cols = output cols
      constexpr int CH = BN / 4;
      for (int i = tid; i < BK * CH; i += THREADS) {
        const int r = i / CH, c = (i % CH) * 4;
        const int gr = r0 + r, gc = n0 + c;
        int bytes = 0;
        if (gr < red_hi && gc < g.N) bytes = min(16, (g.N - gc) * 4);
        cp_async16(bs + r * B_COLS + c, bytes ? g.B + (int64_t)gr * g.ldb + gc : g.B, bytes);
      }
    }
  };

  float acc[MI][NI][4];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.0f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < ktiles) load_stage(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + STAGES - 1;
      if (nk < ktiles) load_stage(nk % STAGES, nk);
      cp_async_commit();
    }
    const float* as = As + (kt % STAGES) * A_ROWS * A_COLS;
    const float* bs = Bs + (kt % STAGES) * B_ROWS * B_COLS;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 8) {
      float af[MI][4], bf[NI][2];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int r = wm * TM + i * 16 + gq;
        if (A_RED_CONTIG) {
          af[i][0] = as[r * A_COLS + kk + tq];
          af[i][1] = as[(r + 8) * A_COLS + kk + tq];
          af[i][2] = as[r * A_COLS + kk + tq + 4];
          af[i][3] = as[(r + 8) * A_COLS + kk + tq + 4];
        } else {
          af[i][0] = as[(kk + tq) * A_COLS + r];
          af[i][1] = as[(kk + tq) * A_COLS + r + 8];
          af[i][2] = as[(kk + tq + 4) * A_COLS + r];
          af[i][3] = as[(kk + tq + 4) * A_COLS + r + 8];
        }
      }
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int c = wn * TN_ + j * 8 + gq;
        if (B_RED_CONTIG) {
          bf[j][0] = bs[c * B_COLS + kk + tq];
          bf[j][1] = bs[c * B_COLS + kk + tq + 4];
        } else {
          bf[j][0] = bs[(kk + tq) * B_COLS + c];
          bf[j][1] = bs[(kk + tq + 4) * B_COLS + c];
        }
      }
      unsigned ah[MI][4], bh[NI][2];
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) ah[i][k] = f2tf32(af[i][k]);   // round-to-nearest, as cuBLAS TF32 does
#pragma unroll
      for (int j = 0; j < NI; ++j)
#pragma unroll
        for (int k = 0; k < 2; ++k) bh[j][k] = f2tf32(bf[j][k]);
      if (PRECISE) {
        unsigned al[MI][4], bl[NI][2];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int k = 0; k < 4; ++k) al[i][k] = f2tf32(af[i][k] - __uint_as_float(ah[i][k]));
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
          for (int k = 0; k < 2; ++k) bl[j][k] = f2tf32(bf[j][k] - __uint_as_float(bh[j][k]));
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            mma_tf32(acc[i][j], al[i], bh[j]);
            mma_tf32(acc[i][j], ah[i], bl[j]);
          }
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) mma_tf32(acc[i][j], ah[i], bh[j]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue --------------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < MI; ++i) {
#pragma unroll
    for (int j = 0; j < NI; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = m0 + wm * TM + i * 16 + gq + h * 8;
        const int col = n0 + wn * TN_ + j * 8 + tq * 2;
        if (row >= g.M) continue;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = col + e;
          if (c >= g.N) continue;
          float v = acc[i][j][h * 2 + e];
          if (MODE == NT) {
            if (g.bias) v += g.bias[c];
            if (g.act == 1) v = v > 0.0f ? v : expf(v) - 1.0f;   // nn.ELU(alpha=1)
            g.C[(int64_t)row * g.ldc + c] = v;
          } else if (MODE == NN) {
            if (g.aux) {
              const float y = g.aux[(int64_t)row * g.ldaux + c];
              v *= (y > 0.0f ? 1.0f : y + 1.0f);                 // elu'(z) = 1 or exp(z) = y + 1
            }
            float* p = g.C + (int64_t)row * g.ldc + c;
            *p = (g.accumulate == 1 || c < g.accumulate) ? *p + v : v;    // 1 = every column, n > 1 = the first n columns
          } else {
            atomicAdd(g.C + (int64_t)row * g.ldc + c, v);
          }
        }
      }
    }
  }
  if (MODE == TN && g.dbias && blockIdx.x == 0) {
    // bias gradient: column sums of this CTA's dY rows [red_lo, red_hi) x [m0, m0+BM)
    for (int c = tid; c < BM; c += THREADS) {
      const int gc = m0 + c;
      if (gc >= g.M) continue;
      float s = 0.0f;
      for (int r = red_lo; r < red_hi; ++r) s += g.A[(int64_t)r * g.lda + gc];
      atomicAdd(g.dbias + gc, s);
    }
  }
}

template <int MODE, int BM, int BN, int WM, int WN>
int launch(const GemmArgs& g, int precise, int splits, cudaStream_t st, const char* name) {
  constexpr bool A_RED = (MODE != TN), B_RED = (MODE == NT);
  constexpr int a_elems = (A_RED ? BM * (BK + 4) : BK * (BM + 8));
  constexpr int b_elems = (B_RED ? BN * (BK + 4) : BK * (BN + 8));
  const int smem = STAGES * (a_elems + b_elems) * (int)sizeof(float);
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, splits);
  auto k0 = gemm_kernel<MODE, BM, BN, WM, WN, false>;
  auto k1 = gemm_kernel<MODE, BM, BN, WM, WN, true>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_done = true;
  }
  if (precise)
    k1<<<grid, WM * WN * 32, smem, st>>>(g);
  else
    k0<<<grid, WM * WN * 32, smem, st>>>(g);
  B200_CHECK_LAUNCH(name);
  return 0;
}

int check_common(const char* who, const void* a, const void* b, const void* c, int lda, int ldb, int M, int N, int K) {
  B200_CHECK_ARG(a && b && c, "%s: null pointer", who);
  B200_CHECK_ARG(M > 0 && N > 0 && K > 0, "%s: M, N, K must be > 0 (got %d %d %d)", who, M, N, K);
  B200_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0, "%s: leading dimensions must be multiples of 4 floats (lda=%d ldb=%d)", who, lda, ldb);
  B200_CHECK_ARG(((uintptr_t)a % 16) == 0 && ((uintptr_t)b % 16) == 0, "%s: operands must be 16-byte aligned", who);
  return 0;
}

}  // namespace

// dX[m,k] (+)= (sum_{n < N <= 4} dY[m,n] W[n,k]) * elu'(Yprev[m,k]): the dgrad of a 1- or 3-wide head is an outer product,
// not a GEMM.  One thread per 4 consecutive k of one row (16-byte accesses; the threads of a row read the same dY values,
// which the hardware broadcasts); `vec` = 0 falls back to one element per thread.
__global__ void __launch_bounds__(256)
small_n_dgrad_kernel(const float* __restrict__ dY, int lddy, const float* __restrict__ W, int ldw, const float* __restrict__ Yprev, int ldyp,
                     float* __restrict__ dX, int lddx, int M, int N, int K, int accumulate, int vec) {
  const int kw = vec ? 4 : 1, kq = K / kw;                 // work items per row
  const unsigned total = (unsigned)M * (unsigned)kq;
  for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
    const int m = (int)(i / (unsigned)kq), k = (int)(i - (unsigned)m * (unsigned)kq) * kw;
    float dy[4];
    for (int n = 0; n < N; ++n) dy[n] = dY[(int64_t)m * lddy + n];
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int n = 0; n < N; ++n) {
      if (vec) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * ldw + k));
        acc[0] = fmaf(dy[n], w4.x, acc[0]);
        acc[1] = fmaf(dy[n], w4.y, acc[1]);
        acc[2] = fmaf(dy[n], w4.z, acc[2]);
        acc[3] = fmaf(dy[n], w4.w, acc[3]);
      } else {
        acc[0] = fmaf(dy[n], __ldg(W + (int64_t)n * ldw + k), acc[0]);
      }
    }
    float* o = dX + (int64_t)m * lddx + k;
    if (vec) {
      if (Yprev) {
        const float4 y = *reinterpret_cast<const float4*>(Yprev + (int64_t)m * ldyp + k);
        acc[0] *= (y.x > 0.0f ? 1.0f : y.x + 1.0f);
        acc[1] *= (y.y > 0.0f ? 1.0f : y.y + 1.0f);
        acc[2] *= (y.z > 0.0f ? 1.0f : y.z + 1.0f);
        acc[3] *= (y.w > 0.0f ? 1.0f : y.w + 1.0f);
      }
      float4 r = make_float4(acc[0], acc[1], acc[2], acc[3]);
      if (accumulate) {
        const float4 old = *reinterpret_cast<const float4*>(o);
        if (accumulate == 1 || k + 0 < accumulate) r.x += old.x;
        if (accumulate == 1 || k + 1 < accumulate) r.y += old.y;
        if (accumulate == 1 || k + 2 < accumulate) r.z += old.z;
        if (accumulate == 1 || k + 3 < accumulate) r.w += old.w;
      }
      *reinterpret_cast<float4*>(o) = r;
    } else {
      if (Yprev) {
        const float y = Yprev[(int64_t)m * ldyp + k];
        acc[0] *= (y > 0.0f ? 1.0f : y + 1.0f);
      }
      *o = (accumulate == 1 || k < accumulate) ? *o + acc[0] : acc[0];
    }
  }
}

extern "C" {

// Y[m,n] = act(sum_k X[m,k] W[n,k] + b[n]) for a head with N <= 4 outputs (value, estimated velocity): N dot products per row,
// not a GEMM.  One warp per row, lanes stride over k (16-byte loads when `vec`), fixed shuffle tree -- exact fp32 in every mode.
__global__ void __launch_bounds__(256)
small_n_forward_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw, const float* __restrict__ bias,
                       float* __restrict__ Y, int ldy, int M, int N, int K, int act, int vec) {
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * 8;
  for (int m = blockIdx.x * 8 + (threadIdx.x >> 5); m < M; m += warps) {
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const float* x = X + (int64_t)m * ldx;
    if (vec) {
      for (int k = lane * 4; k < K; k += 128) {
        const float4 xv = *reinterpret_cast<const float4*>(x + k);
#pragma unroll
        for (int n = 0; n < 4; ++n)
          if (n < N) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * ldw + k));
            acc[n] = fmaf(xv.w, w4.w, fmaf(xv.z, w4.z, fmaf(xv.y, w4.y, fmaf(xv.x, w4.x, acc[n]))));
          }
      }
    } else {
      for (int k = lane; k < K; k += 32) {
        const float xv = x[k];
#pragma unroll
        for (int n = 0; n < 4; ++n)
          if (n < N) acc[n] = fmaf(xv, __ldg(W + (int64_t)n * ldw + k), acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
    if (lane < N) {
      float v = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
      if (bias) v += bias[lane];
      if (act == 1) v = v > 0.0f ? v : expf(v) - 1.0f;   // nn.ELU(alpha=1), as gemm_kernel
      Y[(int64_t)m * ldy + lane] = v;
    }
  }
}

int b200_linear_forward(const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y, int ldy, int M, int N,
                        int K, int act, int precise, void* stream) {
  if (int rc = check_common("b200_linear_forward", X, W, Y, ldx, ldw, M, N, K)) return rc;
  B200_CHECK_ARG(ldx >= K && ldw >= K && ldy >= N && (act == 0 || act == 1), "b200_linear_forward: bad ld/act");
  if (N <= 4) {      // value / estimator heads: streaming dot products (exact fp32 in every mode), ~3 us at 4096 rows
    const int vec = (K % 4 == 0) && (ldx % 4 == 0) && (ldw % 4 == 0) && ((((uintptr_t)X | (uintptr_t)W) & 15) == 0);
    int blocks = (M + 7) / 8;
    blocks = blocks > 148 * 8 ? 148 * 8 : blocks;
    small_n_forward_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(X, ldx, W, ldw, bias, Y, ldy, M, N, K, act, vec);
    B200_CHECK_LAUNCH("small_n_forward_kernel");
    return 0;
  }
  GemmArgs g{};
  g.A = X; g.B = W; g.C = Y; g.bias = bias;
  g.lda = ldx; g.ldb = ldw; g.ldc = ldy;
  g.M = M; g.N = N; g.K = K; g.act = act;
  cudaStream_t st = (cudaStream_t)stream;
  if (N > 64) return launch<NT, 128, 128, 2, 4>(g, precise, 1, st, "linear_forward<128x128>");
  if (N > 32) return launch<NT, 128, 64, 4, 2>(g, precise, 1, st, "linear_forward<128x64>");
  return launch<NT, 128, 32, 8, 1>(g, precise, 1, st, "linear_forward<128x32>");
}

int b200_linear_dgrad(const float* dY, int lddy, const float* W, int ldw, const float* Yprev, int ldyp, float* dX, int lddx, int M,
                      int N, int K, int accumulate, int precise, void* stream) {
  // dX[M,K] = dY[M,N] . W[N,K]: output cols = K, reduction = N
  if (int rc = check_common("b200_linear_dgrad", dY, W, dX, lddy, ldw, M, N, K)) return rc;
  B200_CHECK_ARG(lddy >= N && ldw >= K && lddx >= K, "b200_linear_dgrad: bad leading dimension");
  if (N <= 4) {      // value / estimator heads: a rank-N outer product, pure streaming (exact fp32 in every mode)
    const int vec = (K % 4 == 0) && (ldw % 4 == 0) && (lddx % 4 == 0) && (!Yprev || ldyp % 4 == 0) &&
                    (((uintptr_t)W | (uintptr_t)dX | (uintptr_t)(Yprev ? Yprev : W)) & 15) == 0;
    const int64_t total = (int64_t)M * (vec ? K / 4 : K);
    B200_CHECK_ARG(total < (1ll << 32), "b200_linear_dgrad: problem too large for the streaming head kernel");
    int blocks = (int)((total + 255) / 256);
    blocks = blocks > 148 * 16 ? 148 * 16 : blocks;
    small_n_dgrad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dY, lddy, W, ldw, Yprev, ldyp, dX, lddx, M, N, K, accumulate, vec);
    B200_CHECK_LAUNCH("small_n_dgrad_kernel");
    return 0;
  }
  GemmArgs g{};
  g.A = dY; g.B = W; g.C = dX; g.aux = Yprev;
  g.lda = lddy; g.ldb = ldw; g.ldc = lddx; g.ldaux = ldyp;
  g.M = M; g.N = K; g.K = N; g.accumulate = accumulate;
  cudaStream_t st = (cudaStream_t)stream;
  if (K > 64) return launch<NN, 128, 128, 2, 4>(g, precise, 1, st, "linear_dgrad<128x128>");
  if (K > 32) return launch<NN, 128, 64, 4, 2>(g, precise, 1, st, "linear_dgrad<128x64>");
  return launch<NN, 128, 32, 8, 1>(g, precise, 1, st, "linear_dgrad<128x32>");
}

int b200_linear_wgrad(const float* dY, int lddy, const float* X, int ldx, float* dW, int ldw, float* db, int M, int N, int K,
                      int precise, void* stream) {
  // dW[N,K] += dY[M,N]^T . X[M,K]: output rows = N, output cols = K, reduction = M (split over blockIdx.z)
  if (int rc = check_common("b200_linear_wgrad", dY, X, dW, lddy, ldx, M, N, K)) return rc;
  B200_CHECK_ARG(lddy >= N && ldx >= K && ldw >= K, "b200_linear_wgrad: bad leading dimension");
  GemmArgs g{};
  g.A = dY; g.B = X; g.C = dW; g.dbias = db;
  g.lda = lddy; g.ldb = ldx; g.ldc = ldw;
  g.M = N; g.N = K; g.K = M;
  cudaStream_t st = (cudaStream_t)stream;
  // enough CTAs to fill 148 SMs: tiles(N,K) x splits(M)
  auto plan = [&](int bm, int bn) {
    const int tiles = ((N + bm - 1) / bm) * ((K + bn - 1) / bn);
    int splits = (296 + tiles - 1) / tiles;
    const int max_splits = (M + 4 * BK - 1) / (4 * BK);
    splits = splits < 1 ? 1 : (splits > max_splits ? max_splits : splits);
    int per = (M + splits - 1) / splits;
    per = ((per + BK - 1) / BK) * BK;
    g.red_per_split = per;
    return (M + per - 1) / per;
  };
  if (N > 32 && K > 64) {
    const int s = plan(64, 128);
    return launch<TN, 64, 128, 2, 4>(g, precise, s, st, "linear_wgrad<64x128>");
  }
  if (N > 32) {
    const int s = plan(64, 32);
    return launch<TN, 64, 32, 4, 1>(g, precise, s, st, "linear_wgrad<64x32>");
  }
  if (K > 64) {
    const int s = plan(32, 128);
    return launch<TN, 32, 128, 2, 4>(g, precise, s, st, "linear_wgrad<32x128>");
  }
  const int s = plan(32, 32);
  return launch<TN, 32, 32, 2, 1>(g, precise, s, st, "linear_wgrad<32x32>");
}

}  // extern "C"
