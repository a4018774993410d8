// tcgen05 / TMEM / TMA path for the wide Linear layers -- sm_100a only.
//
//   forward (NT): Y[M,N]  = act(X[M,K] . W[N,K]^T + b)      A, B K-major      (both row-major, K contiguous)
//   dgrad   (NN): dX[M,K] = (dY[M,N] . W[N,K]) * elu'(Y)     A K-major, B MN-major (W rows = reduction)
//   wgrad   (TN): dW[N,K] += dY[M,N]^T . X[M,K]              A, B MN-major     (reduction = rows of dY and X)
//
// kind::tf32 with fp32 accumulation in TMEM: fp32 tensors are consumed in place (no conversion
// pass), which is the reference's own GPU matmul precision (TF32, train.py:39).
//
// One persistent CTA per SM, 320 threads:
//   warp 0  (one elected lane) TMA producer: cp.async.bulk.tensor 128-byte-swizzled boxes into a 4-stage ring
//   warp 1  (one elected lane) MMA issuer:   tcgen05.mma cta_group::1, M=128, N=BN, K=8 per instruction,
//                                             tcgen05.commit frees the smem stage / publishes the accumulator
//   warps 2-9                  epilogue:      tcgen05.ld 32x32b -> registers -> bias / ELU / ELU' -> global
// Two accumulators (2 x BN TMEM columns) let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Shared-memory operand layouts are exactly what TMA SWIZZLE_128B writes (1024-byte aligned stages):
//   K-major  tile [rows][32 fp32]            -> descriptor SBO = 1024 B, start address += 32 B per K=8 step
//   MN-major tile [32 k][32 fp32] per chunk  -> descriptor LBO = chunk stride, SBO = 1024 B, += 1024 B per K=8 step
// (cute/atom/mma_traits_sm100.hpp canonical layouts; instruction descriptor bits as cute::UMMA::InstrDescriptor).
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 32;           // 32 fp32 = 128 B = one swizzle row
constexpr int STAGES = 4;
constexpr int NUM_THREADS = 320;   // TMA warp + MMA warp + 8 epilogue warps

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), version 1.
// layout_type: 2 = SWIZZLE_128B (K-major operands), 1 = SWIZZLE_128B_BASE32B (the only layout for MN-major tf32 operands)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

enum Mode { NT = 0, NN = 1, TN = 2 };

struct TcArgs {
  float* C;
  const float* bias;   // NT
  const float* aux;    // NN: Yprev for elu'
  float* dbias;        // TN (unused here: bias gradients are reduced by the baseline kernel's epilogue path)
  int ldc, ldaux;
  int M, N, K;         // output rows, output cols, reduction
  int act, accumulate;
  int k_per_split;     // TN: reduction elements per blockIdx.y
};

// A-operand smem per stage: K-major  [BM rows][32]         = 16 KB
//                           MN-major 4 chunks x [32 k][32]  = 16 KB   (chunk = 32 fp32 of the M dimension)
// B-operand smem per stage: K-major  [BN rows][32]; MN-major (BN/32) chunks x [32 k][32]      = BN * 128 B
template <int MODE, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const TcArgs g) {
  constexpr bool A_MN = (MODE == TN), B_MN = (MODE != NT);
  constexpr uint32_t A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  constexpr uint32_t IDESC = make_idesc(BM, BN, A_MN, B_MN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (g.M + BM - 1) / BM, n_tiles = (g.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  int k_lo = 0, k_hi = g.K;
  if (MODE == TN) {
    k_lo = blockIdx.y * g.k_per_split;
    k_hi = min(g.K, k_lo + g.k_per_split);
  }
  const int num_kb = (k_hi - k_lo + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full + a, 1);
      mbar_init(tmem_empty + a, 8);      // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(empty_bar + s, ph ^ 1);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(full_bar + s, STAGE_BYTES);
          const int k0 = k_lo + kb * BK;
          if (!A_MN) {
            tma_load_2d(sa, &tmap_a, full_bar + s, k0, m0);                 // box [BM rows][32 k]
          } else {
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) tma_load_2d(sa + c * (BK * 128), &tmap_a, full_bar + s, m0 + c * 32, k0);   // box [32 k][32 m]
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmap_b, full_bar + s, k0, n0);                 // box [BN rows][32 k]
          } else {
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) tma_load_2d(sb + c * (BK * 128), &tmap_b, full_bar + s, n0 + c * 32, k0);   // box [32 k][32 n]
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t it = 0, local_tile = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_tile) {
        const uint32_t acc = local_tile & 1, acc_ph = (local_tile >> 1) & 1;
        mbar_wait(tmem_empty + acc, acc_ph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(full_bar + s, ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            // K-major: 8-row x 128 B atoms (SBO 1024), a K=8 step is 32 B inside the row.
            // MN-major (tf32): 4-k-row x 128 B atoms swizzled in 32 B chunks (SBO 512 between atoms along K,
            // LBO between 32-element chunks along MN), a K=8 step is 8 rows = 1024 B.
            const uint64_t da = A_MN ? make_desc(sa + k * 1024, BK * 128, 512, 1) : make_desc(sa + k * 32, 16, 1024, 2);
            const uint64_t db = B_MN ? make_desc(sb + k * 1024, BK * 128, 512, 1) : make_desc(sb + k * 32, 16, 1024, 2);
            umma_tf32(tmem_d, da, db, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar + s);               // smem stage reusable once these MMAs retire
        }
        umma_commit(tmem_full + acc);               // accumulator complete
      }
    }
  } else {
    // ===== epilogue warps 2..9: TMEM lane quarter = warp % 4; the two warps of a quarter split the columns =====
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    constexpr int CHUNKS = BN / 32, CH_PER_HALF = (CHUNKS + 1) / 2;
    uint32_t local_tile = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_tile) {
      const uint32_t acc = local_tile & 1, acc_ph = (local_tile >> 1) & 1;
      const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
      mbar_wait(tmem_full + acc, acc_ph);
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
#pragma unroll 1
      for (int ci = half * CH_PER_HALF; ci < min(CHUNKS, (half + 1) * CH_PER_HALF); ++ci) {
        const int c = ci * 32;
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c, v);
        if (row >= g.M) continue;
        const int col0 = n0 + c;
        if (col0 >= g.N) continue;
        float* crow = g.C + (int64_t)row * g.ldc + col0;
        const bool full = (col0 + 32 <= g.N);
        if (MODE == NT) {
          if (g.bias) {
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + col0 + j));
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < g.N) v[j] += __ldg(g.bias + col0 + j);
            }
          }
          if (g.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {                    // ELU, branch-free: exp only ever sees x <= 0
              const float e = __expf(fminf(v[j], 0.0f)) - 1.0f;
              v[j] = v[j] > 0.0f ? v[j] : e;
            }
          }
        } else if (MODE == NN) {
          if (g.aux) {
            const float* yrow = g.aux + (int64_t)row * g.ldaux + col0;
            if (full && (((uintptr_t)yrow) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 y4 = *reinterpret_cast<const float4*>(yrow + j);
                v[j] *= (y4.x > 0.0f ? 1.0f : y4.x + 1.0f);
                v[j + 1] *= (y4.y > 0.0f ? 1.0f : y4.y + 1.0f);
                v[j + 2] *= (y4.z > 0.0f ? 1.0f : y4.z + 1.0f);
                v[j + 3] *= (y4.w > 0.0f ? 1.0f : y4.w + 1.0f);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < g.N) {
                  const float y = yrow[j];
                  v[j] *= (y > 0.0f ? 1.0f : y + 1.0f);
                }
            }
          }
          if (g.accumulate) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < g.N) v[j] += crow[j];
          }
        }
        const bool vec = full && (((uintptr_t)crow) & 15) == 0;
        if (MODE == TN) {
          if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) atomicAdd(reinterpret_cast<float4*>(crow + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < g.N) atomicAdd(crow + j, v[j]);
          }
        } else if (vec) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(crow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < g.N) crow[j] = v[j];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 row-major tensor [rows][cols] with leading dimension ld; box = [box_rows][32 cols].
// mn_major = 0: SWIZZLE_128B (16 B chunks);  1: SWIZZLE_128B_ATOM_32B (32 B chunks, for MN-major tf32 operands)
int make_tmap(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int mn_major) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    b200_set_error("cuTensorMapEncodeTiled is unavailable");
    return -2;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200_set_error("cuTensorMapEncodeTiled failed with %d (base %p rows %lld cols %lld ld %lld)", (int)r, (const void*)base, (long long)rows,
                   (long long)cols, (long long)ld);
    return -3;
  }
  return 0;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <int MODE, int BN>
int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& g, int splits, cudaStream_t st, const char* name) {
  constexpr int smem = STAGES * (BM * BK * 4 + BN * BK * 4) + 256 + 1024;
  auto kern = tc_gemm_kernel<MODE, BN>;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      b200_set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
      return (int)e;
    }
    done = true;
  }
  const int tiles = ((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN);
  int ctas = num_sms() / (splits > 1 ? splits : 1);
  ctas = ctas < 1 ? 1 : ctas;
  dim3 grid(tiles < ctas ? tiles : ctas, splits);
  kern<<<grid, NUM_THREADS, smem, st>>>(ta, tb, g);
  B200_CHECK_LAUNCH(name);
  return 0;
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// N-tile width: the widest of {256,128,64,32} whose padding of `cols` (TMA zero-fills, stores are guarded) wastes
// <= 13 % of the MMA work, narrowed while the problem would otherwise leave most SMs without a tile.
int pick_bn(int rows, int cols, int splits_hint = 1) {
  const int m_tiles = (rows + BM - 1) / BM;
  int bn = 32;
  for (int cand : {256, 128, 64, 32}) {
    const int padded = (cols + cand - 1) / cand * cand;
    if (cand > 32 && padded * 100 > cols * 113) continue;
    bn = cand;
    if (m_tiles * (padded / cand) * splits_hint >= (num_sms() * 3) / 4) break;
  }
  return bn;
}

}  // namespace

extern "C" {

// 1 if the tcgen05 path can run this forward problem (else callers use b200_linear_forward)
int b200_tc_linear_supported(int M, int N, int K) { return (N >= 8 && K >= 8 && M >= 1) ? 1 : 0; }

int b200_tc_linear_forward(const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y, int ldy, int M, int N, int K,
                           int act, void* stream) {
  B200_CHECK_ARG(X && W && Y && M > 0 && N > 0 && K > 0, "b200_tc_linear_forward: bad argument");
  B200_CHECK_ARG(b200_tc_linear_supported(M, N, K), "b200_tc_linear_forward: needs N >= 8 and K >= 8 (N=%d K=%d)", N, K);
  B200_CHECK_ARG(ldx % 4 == 0 && ldw % 4 == 0 && aligned16(X) && aligned16(W), "b200_tc_linear_forward: operands need 16-byte rows");
  const int bn = pick_bn(M, N);
  CUtensorMap ta, tb;
  if (int rc = make_tmap(&ta, X, M, K, ldx, BM, 0)) return rc;
  if (int rc = make_tmap(&tb, W, N, K, ldw, bn, 0)) return rc;
  TcArgs g{};
  g.C = Y; g.bias = bias; g.ldc = ldy; g.M = M; g.N = N; g.K = K; g.act = act;
  cudaStream_t st = (cudaStream_t)stream;
  switch (bn) {
    case 256: return launch_tc<NT, 256>(ta, tb, g, 1, st, "tc_forward<256>");
    case 128: return launch_tc<NT, 128>(ta, tb, g, 1, st, "tc_forward<128>");
    case 64: return launch_tc<NT, 64>(ta, tb, g, 1, st, "tc_forward<64>");
    default: return launch_tc<NT, 32>(ta, tb, g, 1, st, "tc_forward<32>");
  }
}

// dX[M,K] (+)= (dY[M,N] . W[N,K]) * elu'(Yprev)
int b200_tc_linear_dgrad(const float* dY, int lddy, const float* W, int ldw, const float* Yprev, int ldyp, float* dX, int lddx, int M,
                         int N, int K, int accumulate, void* stream) {
  B200_CHECK_ARG(dY && W && dX && M > 0 && N > 0 && K > 0, "b200_tc_linear_dgrad: bad argument");
  B200_CHECK_ARG(lddy % 4 == 0 && ldw % 4 == 0 && aligned16(dY) && aligned16(W), "b200_tc_linear_dgrad: operands need 16-byte rows");
  const int bn = pick_bn(M, K);
  CUtensorMap ta, tb;
  if (int rc = make_tmap(&ta, dY, M, N, lddy, BM, 0)) return rc;          // A K-major: [M rows][N reduction]
  if (int rc = make_tmap(&tb, W, N, K, ldw, BK, 1)) return rc;            // B MN-major: box [32 n][32 k]
  TcArgs g{};
  g.C = dX; g.aux = Yprev; g.ldc = lddx; g.ldaux = ldyp; g.M = M; g.N = K; g.K = N; g.accumulate = accumulate;
  cudaStream_t st = (cudaStream_t)stream;
  switch (bn) {
    case 256: return launch_tc<NN, 256>(ta, tb, g, 1, st, "tc_dgrad<256>");
    case 128: return launch_tc<NN, 128>(ta, tb, g, 1, st, "tc_dgrad<128>");
    case 64: return launch_tc<NN, 64>(ta, tb, g, 1, st, "tc_dgrad<64>");
    default: return launch_tc<NN, 32>(ta, tb, g, 1, st, "tc_dgrad<32>");
  }
}

// dW[N,K] += dY[M,N]^T . X[M,K]   (bias gradient is NOT produced here)
int b200_tc_linear_wgrad(const float* dY, int lddy, const float* X, int ldx, float* dW, int ldw, int M, int N, int K, void* stream) {
  B200_CHECK_ARG(dY && X && dW && M > 0 && N > 0 && K > 0, "b200_tc_linear_wgrad: bad argument");
  B200_CHECK_ARG(M >= 32, "b200_tc_linear_wgrad: needs M >= 32");
  B200_CHECK_ARG(lddy % 4 == 0 && ldx % 4 == 0 && aligned16(dY) && aligned16(X), "b200_tc_linear_wgrad: operands need 16-byte rows");
  // output rows = N (UMMA M, tiles of 128), output cols = K (UMMA N), reduction = M
  const int bn = pick_bn(N, K, 8);      // split-K supplies the parallelism: prefer wide tiles (A is re-read per N tile)
  CUtensorMap ta, tb;
  if (int rc = make_tmap(&ta, dY, M, N, lddy, BK, 1)) return rc;          // A MN-major: box [32 m][32 n]
  if (int rc = make_tmap(&tb, X, M, K, ldx, BK, 1)) return rc;            // B MN-major: box [32 m][32 k]
  TcArgs g{};
  g.C = dW; g.ldc = ldw; g.M = N; g.N = K; g.K = M;
  const int tiles = ((N + BM - 1) / BM) * ((K + bn - 1) / bn);
  int splits = num_sms() / (tiles < 1 ? 1 : tiles);
  const int max_splits = (M + 8 * BK - 1) / (8 * BK);
  splits = splits < 1 ? 1 : (splits > max_splits ? max_splits : splits);
  int per = (M + splits - 1) / splits;
  per = (per + BK - 1) / BK * BK;
  g.k_per_split = per;
  splits = (M + per - 1) / per;
  cudaStream_t st = (cudaStream_t)stream;
  switch (bn) {
    case 256: return launch_tc<TN, 256>(ta, tb, g, splits, st, "tc_wgrad<256>");
    case 128: return launch_tc<TN, 128>(ta, tb, g, splits, st, "tc_wgrad<128>");
    case 64: return launch_tc<TN, 64>(ta, tb, g, splits, st, "tc_wgrad<64>");
    default: return launch_tc<TN, 32>(ta, tb, g, splits, st, "tc_wgrad<32>");
  }
}

}  // extern "C"
