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

#include <unordered_map>

#include "common.cuh"

namespace {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 32;           // 32 fp32 = 128 B = one swizzle row
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

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// shared -> global tile store (bulk async-group of the issuing thread); out-of-bounds parts of the box are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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

// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster on one TPC run ONE 256-row UMMA; each stages its own 128
// rows of A and HALF of B, so a 256 x BN tile costs the L2 -> SM path 2/3 of what two 128 x BN tiles do.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;    // shared::cluster address -> the same offset in the pair's rank-0 CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// executed by both CTAs: data lands in the issuing CTA's smem, the bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

enum Mode { NT = 0, NN = 1, TN = 2 };

struct TcArgs {
  float* C;
  const float* bias;   // NT
  const float* aux;    // NN: Yprev for elu'
  float* dbias;        // NN: dbias[N] += column sums of the OUTPUT (= bias gradient of the layer below), or nullptr
  int ldc, ldaux;
  int M, N, K;         // output rows, output cols, reduction
  int act, accumulate;
  int k_per_split;     // TN: reduction elements per blockIdx.y
  unsigned long long* trace;   // optional launch timeline (b200_tc_set_trace): this launch's {min CTA start, max CTA end} in %globaltimer ns
};

// A-operand smem per stage: K-major  [BM rows][32]         = 16 KB
//                           MN-major 4 chunks x [32 k][32]  = 16 KB   (chunk = 32 fp32 of the M dimension)
// B-operand smem per stage: K-major  [BN rows][32]; MN-major (BN/32) chunks x [32 k][32]      = BN * 128 B
// EPI (dgrad only): the epilogue moves its tiles by TMA -- the stored activation (for elu') arrives in shared memory as
// 128-byte-swizzled 32 x 32 boxes on per-warp mbarriers, the result leaves as the same kind of box (cp.async.bulk.tensor
// store): no per-lane global loads / stores, no staging round trip.  Costs 12 KB of shared memory per epilogue warp.
constexpr uint32_t kEpiBytesPerWarp = 3 * 4096;      // aux double buffer + one output tile
template <int BN, bool PAIR, int CPS = 1, bool EPI = false>
struct TileCfg {
  static constexpr int BNH = PAIR ? BN / 2 : BN;                  // rows of B this CTA stages
  static constexpr int BMT = PAIR ? 2 * BM : BM;                  // rows of the output tile (pair: 256)
  static constexpr uint32_t A_BYTES = BM * BK * 4, B_BYTES = BNH * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  // shared-memory budget of the operand ring: what is left of 227 KB after the epilogue's staging tiles (36 KB), the
  // dgrad's column-sum buffer (<= 8 KB) and the barriers
  // (CPS = 2: two CTAs share the SM -- one's epilogue runs under the other's main loop -- with half the ring each)
  static constexpr uint32_t RING_BYTES = CPS == 2 ? 70u * 1024u : (EPI ? 116u * 1024u : 176u * 1024u);
  static constexpr int STAGES_ = (int)(RING_BYTES / STAGE_BYTES) < 8 ? (int)(RING_BYTES / STAGE_BYTES) : 8;
  static_assert(STAGES_ >= 2, "operand ring needs at least two stages");
};

template <int MODE, int BN, bool PAIR, int CPS = 1, bool EPI = false>
__global__ void __launch_bounds__(NUM_THREADS, CPS)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const TcArgs g,
               const __grid_constant__ CUtensorMap tmap_aux, const __grid_constant__ CUtensorMap tmap_out) {
  using Cfg = TileCfg<BN, PAIR, CPS, EPI>;
  static_assert(!EPI || (MODE == NN && CPS == 1), "the TMA epilogue exists for the dgrad, one CTA (or CTA pair) per SM");
  constexpr bool A_MN = (MODE == TN), B_MN = (MODE != NT);
  constexpr int BNH = Cfg::BNH, BMT = Cfg::BMT, STAGES = Cfg::STAGES_;
  constexpr uint32_t A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  constexpr uint32_t IDESC = make_idesc(BMT, BN, A_MN, B_MN);
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;            // 0 = leader (issues the MMAs), 1 = peer
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, num_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* aux_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + 256);   // EPI: [8 epilogue warps][2]
  float* red = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 512);   // NN + dbias: [2 tiles][4 quarters][BN]
  constexpr int STG_PITCH = 36;                                                // floats per staged row (32 + 4 pad)
  float* stage = red + (MODE == NN ? 2 * 4 * BN : 0);                          // [8 epilogue warps][32 rows][STG_PITCH]
  // EPI: [8 epilogue warps][aux 0 | aux 1 | out], 1024-byte aligned (128-byte swizzle atoms) -- takes the staging tiles' place
  uint8_t* epi = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(stage) + 1023) & ~(uintptr_t)1023);

  // Programmatic dependent launch: let the NEXT kernel of the stream start launching now (its CTAs land on each SM as ours
  // leave and run their prologue there); our own reads / writes of global memory wait below until the PREVIOUS kernel of
  // the stream has completed and flushed (griddepcontrol.wait).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (g.trace && threadIdx.x == 0) {
    unsigned long long t_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    atomicMin(g.trace, t_);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (g.M + BMT - 1) / BMT, n_tiles = (g.N + BN - 1) / BN;
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
      mbar_init(tmem_empty + a, PAIR ? 16 : 8);      // one arrive per epilogue warp (pair: of both CTAs, on the leader)
    }
    if (EPI) {
      prefetch_tmap(&tmap_aux);
      prefetch_tmap(&tmap_out);
      for (int a = 0; a < 16; ++a) mbar_init(aux_bar + a, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_ptr, TMEM_COLS);
    else tmem_alloc(tmem_ptr, TMEM_COLS);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();          // the peer's barriers must be initialised before anything is signalled on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  asm volatile("griddepcontrol.wait;" ::: "memory");      // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        const int m0 = (tile / n_tiles) * BMT + (int)rank * BM;             // pair: this CTA's 128 rows of the 256-row tile
        const int n0 = (tile % n_tiles) * BN + (int)rank * BNH;             // pair: this CTA's half of the B rows
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(empty_bar + s, ph ^ 1);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (!PAIR || rank == 0) mbar_expect_tx(full_bar + s, (PAIR ? 2u : 1u) * STAGE_BYTES);   // pair: both CTAs' bytes
          const int k0 = k_lo + kb * BK;
          auto load = [&](void* dst, const CUtensorMap* map, int c0, int c1) {
            if (PAIR) tma_load_2d_pair(dst, map, full_bar + s, c0, c1);
            else tma_load_2d(dst, map, full_bar + s, c0, c1);
          };
          if (!A_MN) {
            load(sa, &tmap_a, k0, m0);                                      // box [BM rows][32 k]
          } else {
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) load(sa + c * (BK * 128), &tmap_a, m0 + c * 32, k0);   // box [32 k][32 m]
          }
          if (!B_MN) {
            load(sb, &tmap_b, k0, n0);                                      // box [BNH rows][32 k]
          } else {
#pragma unroll
            for (int c = 0; c < BNH / 32; ++c) load(sb + c * (BK * 128), &tmap_b, n0 + c * 32, k0);   // box [32 k][32 n]
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0 && rank == 0) {
      uint32_t it = 0, local_tile = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers, ++local_tile) {
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
            if (PAIR) umma_tf32_pair(tmem_d, da, db, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
            else umma_tf32(tmem_d, da, db, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          }
          if (PAIR) umma_commit_pair(empty_bar + s);   // smem stage reusable (in both CTAs) once these MMAs retire
          else umma_commit(empty_bar + s);
        }
        if (PAIR) umma_commit_pair(tmem_full + acc);   // accumulator complete (both CTAs' epilogues)
        else umma_commit(tmem_full + acc);
      }
    }
  } else {
    // ===== epilogue warps 2..9: TMEM lane quarter = warp % 4; the two warps of a quarter split the columns =====
    // tcgen05.ld hands every lane one ROW of a 32 x 32 chunk; global memory wants consecutive lanes on consecutive
    // addresses.  Each warp therefore owns a 32 x 32 staging tile in shared memory (pitch 36 floats: conflict-free for
    // both the row-per-lane and the 8-lanes-per-row access): rows go in, 128-byte row segments come out, four rows per
    // instruction.  The dgrad's elu' input travels the other way (coalesced load -> staging -> row per lane).
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    constexpr int CHUNKS = BN / 32, CH_PER_HALF = (CHUNKS + 1) / 2;
    if constexpr (EPI) {
      // ---- TMA epilogue (dgrad): per 32 x 32 chunk -- aux box (prefetched one chunk ahead) -> row per lane from the swizzled
      // tile, x accumulator row from TMEM, x elu', -> swizzled output tile -> one TMA store; column sums from the output tile.
      // Box element (r, c) of a SWIZZLE_128B tile lives at r * 128 + ((c / 4) ^ (r & 7)) * 16 + (c % 4) * 4.
      uint8_t* ebuf = epi + (size_t)(warp - 2) * kEpiBytesPerWarp;
      uint8_t* obuf = ebuf + 2 * 4096;
      uint64_t* abar = aux_bar + (warp - 2) * 2;
      const bool has_aux = g.aux != nullptr;
      uint32_t q = 0;                                          // chunks this warp has consumed: buffer q & 1, parity (q >> 1) & 1
      uint32_t local_tile = 0;
      const uint32_t sw = (uint32_t)(lane & 7);
      for (int tile = worker; tile < num_tiles; tile += num_workers, ++local_tile) {
        const uint32_t acc = local_tile & 1, acc_ph = (local_tile >> 1) & 1;
        const int m0 = (tile / n_tiles) * BMT + (int)rank * BM, n0 = (tile % n_tiles) * BN;      // pair: this CTA's 128 rows
        const int row0 = m0 + quarter * 32;
        const int ci_lo = half * CH_PER_HALF, ci_hi = min(CHUNKS, (half + 1) * CH_PER_HALF);
        auto issue_aux = [&](int ci, uint32_t qq) {
          if (has_aux && lane == 0) {
            mbar_expect_tx(abar + (qq & 1), 4096);
            tma_load_2d(ebuf + (qq & 1) * 4096, &tmap_aux, abar + (qq & 1), n0 + ci * 32, row0);
          }
        };
        if (ci_lo < ci_hi) issue_aux(ci_lo, q);                 // before the accumulator is even complete
        mbar_wait(tmem_full + acc, acc_ph);
        tc_fence_after();
        const bool want_colsum = g.dbias != nullptr;
        float* red_tile = red + (local_tile & 1) * 4 * BN + quarter * BN;
#pragma unroll 1
        for (int ci = ci_lo; ci < ci_hi; ++ci, ++q) {
          const int c = ci * 32;
          if (ci + 1 < ci_hi) issue_aux(ci + 1, q + 1);         // the other buffer: its last reader was chunk q - 1
          float v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c, v);
          if (has_aux) {
            mbar_wait(abar + (q & 1), (q >> 1) & 1);
            const uint8_t* arow = ebuf + (q & 1) * 4096 + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 y4 = *reinterpret_cast<const float4*>(arow + (((uint32_t)j ^ sw) << 4));
              v[4 * j] *= (y4.x > 0.0f ? 1.0f : y4.x + 1.0f);      // elu'(y) from the stored y (zero-filled rows / columns: x 1)
              v[4 * j + 1] *= (y4.y > 0.0f ? 1.0f : y4.y + 1.0f);
              v[4 * j + 2] *= (y4.z > 0.0f ? 1.0f : y4.z + 1.0f);
              v[4 * j + 3] *= (y4.w > 0.0f ? 1.0f : y4.w + 1.0f);
            }
          }
          if (lane == 0) bulk_wait_read0();                     // the previous chunk's store has read the output tile
          __syncwarp();
          uint8_t* orow = obuf + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(orow + (((uint32_t)j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && n0 + c < g.N && row0 < g.M) {
            tma_store_2d(&tmap_out, obuf, n0 + c, row0);
            bulk_commit();
          }
          if (want_colsum) {                                    // column `lane` of the tile (rows past M carry zeros)
            float cs = 0.0f;
            const uint32_t cchunk = (uint32_t)lane >> 2, cw = ((uint32_t)lane & 3) << 2;
#pragma unroll
            for (int r = 0; r < 32; ++r) cs += *reinterpret_cast<const float*>(obuf + r * 128 + ((cchunk ^ (uint32_t)(r & 7)) << 4) + cw);
            red_tile[c + lane] = cs;
          }
        }
        if (want_colsum) {
          asm volatile("bar.sync 2, 256;" ::: "memory");
          const int tcol = (int)threadIdx.x - 64;
          if (tcol < BN && n0 + tcol < g.N) {
            const float* r = red + (local_tile & 1) * 4 * BN + tcol;
            atomicAdd(g.dbias + n0 + tcol, (r[0] + r[BN]) + (r[2 * BN] + r[3 * BN]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_leader(tmem_empty + acc);
          else mbar_arrive(tmem_empty + acc);
        }
      }
      if (lane == 0) bulk_wait_all();                           // every store of this warp is complete before the CTA retires
    } else {
    float* stg = stage + (warp - 2) * (32 * STG_PITCH);
    const int sr = lane >> 3, sc = (lane & 7) * 4;            // staging coordinates of this lane in the "8 lanes per row" view
    uint32_t local_tile = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers, ++local_tile) {
      const uint32_t acc = local_tile & 1, acc_ph = (local_tile >> 1) & 1;
      const int m0 = (tile / n_tiles) * BMT + (int)rank * BM, n0 = (tile % n_tiles) * BN;
      const int row0 = m0 + quarter * 32;                     // first row of this warp's chunk rows
      const int row = row0 + lane;
      const int ci_lo = half * CH_PER_HALF, ci_hi = min(CHUNKS, (half + 1) * CH_PER_HALF);
      // dgrad: the stored activation of the layer below (for elu') does not depend on the MMAs -- fetch chunk i + 1 while
      // chunk i is in flight, and the first chunk before the accumulator is even complete
      float4 ax_next[8];
      auto load_aux = [&](int ci) {
        if (MODE != NN || g.aux == nullptr) return;
        const int col = n0 + ci * 32 + sc;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = row0 + it * 4 + sr;
          const float* y = g.aux + (int64_t)r * g.ldaux + col;
          float4 v4 = make_float4(1.0f, 1.0f, 1.0f, 1.0f);    // y > 0 -> factor 1
          if (r < g.M) {
            if (col + 3 < g.N && (((uintptr_t)y) & 15) == 0) {
              v4 = *reinterpret_cast<const float4*>(y);
            } else {
              if (col < g.N) v4.x = y[0];
              if (col + 1 < g.N) v4.y = y[1];
              if (col + 2 < g.N) v4.z = y[2];
              if (col + 3 < g.N) v4.w = y[3];
            }
          }
          ax_next[it] = v4;
        }
      };
      if (CPS == 1 && ci_lo < ci_hi) load_aux(ci_lo);      // (CPS = 2: the partner CTA hides the latency; no prefetch registers)
      mbar_wait(tmem_full + acc, acc_ph);
      tc_fence_after();
      const bool want_colsum = (MODE == NN) && g.dbias != nullptr;
      float* red_tile = red + (local_tile & 1) * 4 * BN + quarter * BN;
#pragma unroll 1
      for (int ci = ci_lo; ci < ci_hi; ++ci) {
        const int c = ci * 32;
        const int col0 = n0 + c;
        float ay[32];
        if (MODE == NN && g.aux != nullptr) {
          if (CPS == 2) load_aux(ci);
#pragma unroll
          for (int it = 0; it < 8; ++it) *reinterpret_cast<float4*>(stg + (it * 4 + sr) * STG_PITCH + sc) = ax_next[it];
          if (CPS == 1 && ci + 1 < ci_hi) load_aux(ci + 1);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 y4 = *reinterpret_cast<const float4*>(stg + lane * STG_PITCH + j);
            ay[j] = y4.x; ay[j + 1] = y4.y; ay[j + 2] = y4.z; ay[j + 3] = y4.w;
          }
          __syncwarp();
        }
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c, v);
        if (col0 >= g.N && !want_colsum) continue;            // warp-uniform
        if (MODE == NT) {
          if (g.bias) {
            if (col0 + 32 <= g.N) {
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
          if (g.aux != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= (ay[j] > 0.0f ? 1.0f : ay[j] + 1.0f);     // elu'(y) from the stored y
          }
          if (g.accumulate && row < g.M) {                    // rare (the actor's latent columns): row-per-lane access
            const float* crow = g.C + (int64_t)row * g.ldc + col0;
            const int acc_cols = g.accumulate == 1 ? g.N : min(g.accumulate, g.N);   // 1 = every column, n > 1 = the first n
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < acc_cols) v[j] += crow[j];
          }
        }
        // rows -> staging -> coalesced row segments
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(stg + lane * STG_PITCH + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        {
          const int col = col0 + sc;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = row0 + it * 4 + sr;
            if (r >= g.M || col >= g.N) continue;
            const float4 x = *reinterpret_cast<const float4*>(stg + (it * 4 + sr) * STG_PITCH + sc);
            float* dst = g.C + (int64_t)r * g.ldc + col;
            if (MODE == TN) {
              if (col + 3 < g.N && (((uintptr_t)dst) & 15) == 0) {
                atomicAdd(reinterpret_cast<float4*>(dst), x);
              } else {
                atomicAdd(dst, x.x);
                if (col + 1 < g.N) atomicAdd(dst + 1, x.y);
                if (col + 2 < g.N) atomicAdd(dst + 2, x.z);
                if (col + 3 < g.N) atomicAdd(dst + 3, x.w);
              }
            } else if (col + 3 < g.N && (((uintptr_t)dst) & 15) == 0) {
              *reinterpret_cast<float4*>(dst) = x;
            } else {
              dst[0] = x.x;
              if (col + 1 < g.N) dst[1] = x.y;
              if (col + 2 < g.N) dst[2] = x.z;
              if (col + 3 < g.N) dst[3] = x.w;
            }
          }
        }
        if (MODE == NN && want_colsum) {
          // column `lane` of the staged 32 x 32 chunk (rows past M carry zeros: TMA zero-fills A out of bounds)
          float cs = 0.0f;
#pragma unroll
          for (int r = 0; r < 32; ++r) cs += stg[r * STG_PITCH + lane];
          red_tile[c + lane] = cs;
        }
        __syncwarp();                                         // staging is rewritten by the next chunk
      }
      if (MODE == NN && want_colsum) {
        // the 4 row quarters meet in shared memory: one atomic per column per tile (double-buffered by tile parity, so one
        // barrier per tile orders both the reads after the writes and the next-but-one tile's writes after these reads)
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const int tcol = (int)threadIdx.x - 64;
        if (tcol < BN && n0 + tcol < g.N) {
          const float* r = red + (local_tile & 1) * 4 * BN + tcol;
          atomicAdd(g.dbias + n0 + tcol, (r[0] + r[BN]) + (r[2 * BN] + r[3 * BN]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(tmem_empty + acc);
        else mbar_arrive(tmem_empty + acc);
      }
    }
    }      // !EPI
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();          // nobody leaves while its partner can still signal its barriers / read its smem
  else __syncthreads();
  if (warp == 1) {
    if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (g.trace && threadIdx.x == 0) {
    unsigned long long t_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    atomicMax(g.trace + 1, t_);
  }
}


// ---- tc_chain_kernel: a whole Linear / ELU CHAIN (an MLP's forward) in ONE persistent launch -------------------------------
// The rollout's policy inference runs 17 GEMMs per env step at M = 4096, each 10-18 us of which ~3 us is work: launch latency,
// TMEM allocation, barrier set-up and the pipeline fill are paid per layer.  This kernel keeps the CTAs (TMEM, barriers, the
// warp roles of tc_gemm_kernel's forward path) alive across the layers of a chain: every layer is the usual persistent tile
// loop over ITS tensor maps, and between two layers the CTAs meet at a grid-wide barrier in global memory -- the epilogue
// threads of a CTA finish their stores, one of them fences and bumps a monotonic counter, and the TMA producer of every CTA
// spins until all CTAs have bumped it before it loads the next layer's activations (which are this layer's outputs, read back
// from L2).  Every CTA of the (persistent, at most two per SM) grid is resident, so the spin cannot deadlock.  The counter's
// base for the next launch is published by the last CTA to leave: CUDA-graph replays need no host reset.
constexpr int kMaxChain = 8;
struct ChainLayer {
  float* C;
  const float* bias;
  int ldc, N, K, act;
};
struct ChainArgs {
  ChainLayer layer[kMaxChain];
  int num_layers, M;
  unsigned int* sync;      // device [4], zero-initialised once: [0] arrivals (monotonic), [1] base of this launch, [2] exit counter
};
struct ChainMaps {
  CUtensorMap a[kMaxChain], b[kMaxChain];
};

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int BN, int CPS>
__global__ void __launch_bounds__(NUM_THREADS, CPS)
tc_chain_kernel(const __grid_constant__ ChainMaps maps, const ChainArgs g) {
  using Cfg = TileCfg<BN, false, CPS>;
  constexpr int STAGES = Cfg::STAGES_;
  constexpr uint32_t A_BYTES = Cfg::A_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  constexpr uint32_t IDESC = make_idesc(BM, BN, 0, 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  constexpr int STG_PITCH = 36;
  float* stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 512);      // [8 epilogue warps][32 rows][STG_PITCH]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int worker = (int)blockIdx.x, num_workers = (int)gridDim.x;
  const int m_tiles = (g.M + BM - 1) / BM;
  const unsigned int base = g.sync[1];                      // arrivals before this launch (written by the previous launch's last CTA)

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full + a, 1);
      mbar_init(tmem_empty + a, 8);
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
      for (int l = 0; l < g.num_layers; ++l) {
        const int n_tiles = (g.layer[l].N + BN - 1) / BN, num_tiles = m_tiles * n_tiles;
        const int num_kb = (g.layer[l].K + BK - 1) / BK;
        prefetch_tmap(&maps.a[l]);
        prefetch_tmap(&maps.b[l]);
        if (l > 0) {      // this layer's activations are the previous layer's outputs: every CTA must have stored them
          const unsigned int want = (unsigned int)l * (unsigned int)num_workers;
          while (ld_acquire_gpu_u32(g.sync) - base < want) {
          }
          asm volatile("fence.proxy.async;" ::: "memory");      // generic-proxy stores (observed above) before our TMA loads
        }
        for (int tile = worker; tile < num_tiles; tile += num_workers) {
          const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(empty_bar + s, ph ^ 1);
            uint8_t* sa = smem + s * STAGE_BYTES;
            mbar_expect_tx(full_bar + s, STAGE_BYTES);
            tma_load_2d(sa, &maps.a[l], full_bar + s, kb * BK, m0);
            tma_load_2d(sa + A_BYTES, &maps.b[l], full_bar + s, kb * BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t it = 0, local_tile = 0;
      for (int l = 0; l < g.num_layers; ++l) {
        const int n_tiles = (g.layer[l].N + BN - 1) / BN, num_tiles = m_tiles * n_tiles;
        const int num_kb = (g.layer[l].K + BK - 1) / BK;
        for (int tile = worker; tile < num_tiles; tile += num_workers, ++local_tile) {
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
            for (int k = 0; k < BK / 8; ++k)
              umma_tf32(tmem_d, make_desc(sa + k * 32, 16, 1024, 2), make_desc(sb + k * 32, 16, 1024, 2), IDESC, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(empty_bar + s);
          }
          umma_commit(tmem_full + acc);
        }
      }
    }
  } else {
    // ===== epilogue warps 2..9 (tc_gemm_kernel's forward epilogue: bias, ELU, staged coalesced stores) =====
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    constexpr int CHUNKS = BN / 32, CH_PER_HALF = (CHUNKS + 1) / 2;
    float* stg = stage + (warp - 2) * (32 * STG_PITCH);
    const int sr = lane >> 3, sc = (lane & 7) * 4;
    uint32_t local_tile = 0;
    for (int l = 0; l < g.num_layers; ++l) {
      const ChainLayer L = g.layer[l];
      const int n_tiles = (L.N + BN - 1) / BN, num_tiles = m_tiles * n_tiles;
      for (int tile = worker; tile < num_tiles; tile += num_workers, ++local_tile) {
        const uint32_t acc = local_tile & 1, acc_ph = (local_tile >> 1) & 1;
        const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
        const int row0 = m0 + quarter * 32;
        const int ci_lo = half * CH_PER_HALF, ci_hi = min(CHUNKS, (half + 1) * CH_PER_HALF);
        mbar_wait(tmem_full + acc, acc_ph);
        tc_fence_after();
#pragma unroll 1
        for (int ci = ci_lo; ci < ci_hi; ++ci) {
          const int c = ci * 32, col0 = n0 + c;
          float v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c, v);
          if (col0 >= L.N) continue;                            // warp-uniform
          if (L.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < L.N) v[j] += __ldg(L.bias + col0 + j);
          }
          if (L.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float e = __expf(fminf(v[j], 0.0f)) - 1.0f;
              v[j] = v[j] > 0.0f ? v[j] : e;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(stg + lane * STG_PITCH + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          __syncwarp();
          const int col = col0 + sc;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = row0 + it * 4 + sr;
            if (r >= g.M || col >= L.N) continue;
            const float4 x = *reinterpret_cast<const float4*>(stg + (it * 4 + sr) * STG_PITCH + sc);
            float* dst = L.C + (int64_t)r * L.ldc + col;
            if (col + 3 < L.N && (((uintptr_t)dst) & 15) == 0) {
              *reinterpret_cast<float4*>(dst) = x;
            } else {
              dst[0] = x.x;
              if (col + 1 < L.N) dst[1] = x.y;
              if (col + 2 < L.N) dst[2] = x.z;
              if (col + 3 < L.N) dst[3] = x.w;
            }
          }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty + acc);
      }
      // ---- grid barrier, arrive side: all 8 epilogue warps of this CTA have stored this layer's tiles.  The counter is ONE
      //      monotonic count of arrivals, so nobody may arrive for layer l before the barrier of layer l - 1 is complete -- a
      //      CTA without tiles in a layer would otherwise run ahead and its early arrivals would release the others' waits
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (warp == 2 && lane == 0) {
        const unsigned int prev = (unsigned int)l * (unsigned int)num_workers;
        while (ld_acquire_gpu_u32(g.sync) - base < prev) {
        }
        __threadfence();
        atomicAdd(g.sync, 1u);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
  if (threadIdx.x == 0) {      // the last CTA to leave publishes the base of the next launch
    __threadfence();
    if (atomicAdd(g.sync + 2, 1u) == (unsigned int)num_workers - 1u) {
      g.sync[1] = base + (unsigned int)g.num_layers * (unsigned int)num_workers;
      g.sync[2] = 0u;
    }
  }
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

// launch timeline (b200_tc_set_trace): launch i of the tcgen05 GEMMs since the call min/max-es its CTAs' %globaltimer into slot i
// {start, end} of the caller's DEVICE buffer and describes itself in slot i {M, N, K, mode * 1000 + BN + (pair ? 500 : 0)} of the
// caller's HOST array (a plain store at call time: capture-safe); under CUDA-graph replay a node keeps the slot it was captured with
unsigned long long* g_trace = nullptr;
long long* g_trace_meta = nullptr;
int g_trace_cap = 0, g_trace_next = 0;

// b200_tc_set_sm_cap: launches issued while a cap is set size their persistent grid (and the wgrad's split-K) for `cap` SMs
// instead of all of them.  A persistent one-wave kernel holds every SM until it ends, so a LOW-priority big GEMM starves the
// small kernels of the critical path no matter what the stream priorities say (priorities only order PENDING CTAs): the
// learner caps the side-stream chains and leaves the rest of the machine to the critical path.
// The cap is a property of the STREAM a launch goes to (b200_tc_set_stream_sm_cap, set once when the learner creates its
// low-priority side streams): no process-wide switch to toggle around blocks of launches, nothing to restore after an
// exception.  b200_tc_set_sm_cap (process-wide) remains for A/B tools.
int g_sm_cap = 0;
std::unordered_map<cudaStream_t, int> g_stream_caps;
int avail_sms(cudaStream_t st) {
  const int n = num_sms();
  int cap = g_sm_cap;
  auto it = g_stream_caps.find(st);
  if (it != g_stream_caps.end()) cap = it->second;
  return (cap > 0 && cap < n) ? cap : n;
}

// b200_tc_set_ctas_per_sm: 0 (default) = by shape: outputs up to 256 columns wide run two CTAs per SM on <= 128-wide tiles
// (measured on B200, M = 24576: forward 256 x 512 26.1 -> 23.3 us, 128 x 256 14.8 -> 12.8 us), wider ones one CTA per SM on
// 256-wide tiles / CTA pairs (512 x 627: 43.5 us against 62.9 us with two CTAs); 1 / 2 force either (A/B runs)
int g_cps = 0;
int cps_for(int cols) { return g_cps == 0 ? (cols <= 256 ? 2 : 1) : g_cps; }
int g_wgrad_pairs = 1;    // b200_tc_set_wgrad_pairs: weight gradients with >= 256 output rows on CTA pairs (256-row UMMA, X staged once per pair)
int g_tma_epi = 1;        // b200_tc_set_tma_epilogue: dgrad tiles >= 64 wide move their epilogue tiles by TMA (0: legacy staging path)
int g_pdl = 0;            // b200_tc_set_pdl: programmatic dependent launch of the tcgen05 GEMMs (measured: no gain)

template <int MODE, int BN, bool PAIR = false, int CPS = 1, bool EPI = false>
int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& g, int splits, cudaStream_t st, const char* name,
              const CUtensorMap* taux = nullptr, const CUtensorMap* tout = nullptr) {
  using Cfg = TileCfg<BN, PAIR, CPS, EPI>;
  static_assert(CPS == 1 || (!PAIR && 2 * BN <= 256), "two CTAs per SM: 2 x 256 TMEM columns, no CTA pairs");
  constexpr int smem = Cfg::STAGES_ * (int)Cfg::STAGE_BYTES + 512 + (MODE == NN ? 2 * 4 * BN * 4 : 0) +
                       (EPI ? 8 * (int)kEpiBytesPerWarp + 1024 : 8 * 32 * 36 * 4) + 1024;
  static_assert(smem <= 227 * 1024, "tile configuration does not fit shared memory");
  auto kern = tc_gemm_kernel<MODE, BN, PAIR, CPS, EPI>;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess && CPS == 2) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) {
      b200_set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
      return (int)e;
    }
    done = true;
  }
  TcArgs gt = g;
  if (g_trace && g_trace_next < g_trace_cap) {
    const int i = g_trace_next++;
    if (g_trace_meta) {
      long long* m = g_trace_meta + 4 * (size_t)i;
      m[0] = g.M; m[1] = g.N; m[2] = g.K; m[3] = MODE * 1000 + BN + (PAIR ? 500 : 0);
    }
    gt.trace = g_trace + 2 * (size_t)i;
  }
  const int tiles = ((g.M + Cfg::BMT - 1) / Cfg::BMT) * ((g.N + BN - 1) / BN);
  int workers = (PAIR ? avail_sms(st) / 2 : avail_sms(st) * CPS) / (splits > 1 ? splits : 1);
  workers = workers < 1 ? 1 : workers;
  workers = tiles < workers ? tiles : workers;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(PAIR ? 2 * workers : workers, splits);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_pdl) {                                               // overlap our prologue with the previous kernel's tail
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;      // the pair: two CTAs of one TPC
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, gt, taux ? *taux : ta, tout ? *tout : ta);
  if (e != cudaSuccess) {
    b200_set_error("%s: cudaLaunchKernelEx: %s", name, cudaGetErrorString(e));
    return (int)e;
  }
  B200_CHECK_LAUNCH(name);
  return 0;
}

// CTA pairs (256-row tiles) pay off once there are enough pair tiles to occupy every pair; the N-tile width is the one
// that wastes the fewest tile slots over the rounds of the persistent loop (wide tiles preferred at equal waste).
int g_pair_mode = 1;      // b200_tc_set_pair_mode: 0 = never use the CTA-pair kernels (A/B measurements, tests)
int pick_pair_bn(int rows, int cols) {
  if (!g_pair_mode) return 0;
  const int pairs = num_sms() / 2, m_tiles = (rows + 2 * BM - 1) / (2 * BM);
  int best = 0;
  double best_score = 0.0;
  for (int cand : {256, 128}) {
    const int padded = (cols + cand - 1) / cand * cand;
    if (padded * 100 > cols * 113) continue;
    const int tiles = m_tiles * (padded / cand);
    if (tiles < pairs || m_tiles < pairs / 2) continue;
    const int rounds = (tiles + pairs - 1) / pairs;
    const double score = (double)tiles / ((double)rounds * pairs) * (cand == 256 ? 1.0 : (cand == 128 ? 0.97 : 0.90));
    if (score > best_score) {
      best_score = score;
      best = cand;
    }
  }
  return best;
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// N-tile width for the persistent single-CTA kernels: among {256,128,64,32} with <= 13 % padding of `cols` (TMA zero-fills,
// stores are guarded) the one that fills the SMs best over the rounds of the persistent loop, wide tiles preferred at
// equal fill (operand re-reads).  `split_k` (wgrad): split-K supplies the parallelism, take the widest tile.
int pick_bn(int rows, int cols, int split_k = 0) {
  const int m_tiles = (rows + BM - 1) / BM, sms = num_sms();
  int best = 32;
  double best_score = -1.0;
  for (int cand : {256, 128, 64, 32}) {
    const int padded = (cols + cand - 1) / cand * cand;
    if (cand > 32 && padded * 100 > cols * 113) continue;
    if (split_k) return cand;
    const int tiles = m_tiles * (padded / cand);
    const int rounds = (tiles + sms - 1) / sms;
    const double pref = cand == 256 ? 1.0 : (cand == 128 ? 0.95 : (cand == 64 ? 0.85 : 0.70));
    const double score = (double)tiles / ((double)rounds * sms) * pref;
    if (score > best_score) {
      best_score = score;
      best = cand;
    }
  }
  return best;
}


template <int BN, int CPS>
int launch_chain(const ChainMaps& maps, const ChainArgs& g, int workers, cudaStream_t st, const char* name) {
  using Cfg = TileCfg<BN, false, CPS>;
  constexpr int smem = Cfg::STAGES_ * (int)Cfg::STAGE_BYTES + 512 + 8 * 32 * 36 * 4 + 1024;
  static_assert(smem <= 227 * 1024, "chain tile configuration does not fit shared memory");
  auto kern = tc_chain_kernel<BN, CPS>;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess && CPS == 2) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) {
      b200_set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
      return (int)e;
    }
    done = true;
  }
  kern<<<workers, NUM_THREADS, smem, st>>>(maps, g);
  B200_CHECK_LAUNCH(name);
  return 0;
}

}  // namespace

extern "C" {

int b200_tc_set_trace(unsigned long long* dev_buf, long long* host_meta, int capacity) {
  g_trace = dev_buf;
  g_trace_meta = host_meta;
  g_trace_cap = dev_buf ? capacity : 0;
  g_trace_next = 0;
  return 0;
}

int b200_tc_set_pdl(int on) {
  g_pdl = on ? 1 : 0;
  return 0;
}

int b200_tc_set_sm_cap(int sms) {
  g_sm_cap = sms > 0 ? sms : 0;
  return 0;
}

int b200_tc_set_stream_sm_cap(void* stream, int sms) {
  cudaStream_t st = (cudaStream_t)stream;
  if (sms > 0) g_stream_caps[st] = sms;
  else g_stream_caps.erase(st);
  return 0;
}

int b200_tc_set_ctas_per_sm(int n) {
  g_cps = (n == 1 || n == 2) ? n : 0;
  return 0;
}

int b200_tc_set_wgrad_pairs(int on) {
  g_wgrad_pairs = on ? 1 : 0;
  return 0;
}

int b200_tc_set_tma_epilogue(int on) {
  g_tma_epi = on ? 1 : 0;
  return 0;
}

int b200_tc_set_pair_mode(int on) {
  g_pair_mode = on < 0 ? 0 : (on > 2 ? 2 : on);      // 0 off, 1 forward only (default), 2 forward + dgrad
  return 0;
}

// 1 if the tcgen05 path can run this forward problem (else callers use b200_linear_forward)
int b200_tc_linear_supported(int M, int N, int K) { return (N >= 8 && K >= 8 && M >= 1) ? 1 : 0; }

int b200_tc_linear_forward(const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y, int ldy, int M, int N, int K,
                           int act, void* stream) {
  B200_CHECK_ARG(X && W && Y && M > 0 && N > 0 && K > 0, "b200_tc_linear_forward: bad argument");
  B200_CHECK_ARG(b200_tc_linear_supported(M, N, K), "b200_tc_linear_forward: needs N >= 8 and K >= 8 (N=%d K=%d)", N, K);
  B200_CHECK_ARG(ldx % 4 == 0 && ldw % 4 == 0 && aligned16(X) && aligned16(W), "b200_tc_linear_forward: operands need 16-byte rows");
  const int cps = cps_for(N);
  const int pbn = cps == 2 ? 0 : pick_pair_bn(M, N);
  int bn = pbn ? pbn : pick_bn(M, N);
  if (cps == 2 && bn > 128) bn = 128;
  CUtensorMap ta, tb;
  if (int rc = make_tmap(&ta, X, M, K, ldx, BM, 0)) return rc;
  if (int rc = make_tmap(&tb, W, N, K, ldw, pbn ? bn / 2 : bn, 0)) return rc;   // pair: each CTA stages half of the B rows
  TcArgs g{};
  g.C = Y; g.bias = bias; g.ldc = ldy; g.M = M; g.N = N; g.K = K; g.act = act;
  cudaStream_t st = (cudaStream_t)stream;
  switch (pbn) {
    case 256: return launch_tc<NT, 256, true>(ta, tb, g, 1, st, "tc_forward_pair<256>");
    case 128: return launch_tc<NT, 128, true>(ta, tb, g, 1, st, "tc_forward_pair<128>");
    case 64: return launch_tc<NT, 64, true>(ta, tb, g, 1, st, "tc_forward_pair<64>");
    default: break;
  }
  if (cps == 2) switch (bn) {
      case 128: return launch_tc<NT, 128, false, 2>(ta, tb, g, 1, st, "tc_forward<128,2/SM>");
      case 64: return launch_tc<NT, 64, false, 2>(ta, tb, g, 1, st, "tc_forward<64,2/SM>");
      default: return launch_tc<NT, 32, false, 2>(ta, tb, g, 1, st, "tc_forward<32,2/SM>");
    }
  switch (bn) {
    case 256: return launch_tc<NT, 256>(ta, tb, g, 1, st, "tc_forward<256>");
    case 128: return launch_tc<NT, 128>(ta, tb, g, 1, st, "tc_forward<128>");
    case 64: return launch_tc<NT, 64>(ta, tb, g, 1, st, "tc_forward<64>");
    default: return launch_tc<NT, 32>(ta, tb, g, 1, st, "tc_forward<32>");
  }
}


// A chain of Linear (+ ELU) layers, Y_l = act_l(Y_{l-1} . W_l^T + b_l), in ONE launch (tc_chain_kernel); see include/b200gym.h
int b200_tc_mlp_forward(const B200MlpLayer* layers, int num_layers, const float* X, int ldx, int M, unsigned int* sync, int max_ctas,
                        void* stream) {
  B200_CHECK_ARG(layers && X && sync && M > 0 && num_layers >= 1 && num_layers <= kMaxChain, "b200_tc_mlp_forward: bad argument");
  ChainMaps maps;
  ChainArgs g{};
  g.num_layers = num_layers;
  g.M = M;
  g.sync = sync;
  const float* in = X;
  int ld_in = ldx, max_n = 0;
  for (int l = 0; l < num_layers; ++l) {
    const B200MlpLayer& L = layers[l];
    B200_CHECK_ARG(L.W && L.Y && L.N > 0 && L.K >= 8, "b200_tc_mlp_forward: layer %d: needs W, Y, N > 0, K >= 8", l);
    B200_CHECK_ARG(ld_in % 4 == 0 && L.ldw % 4 == 0 && aligned16(in) && aligned16(L.W), "b200_tc_mlp_forward: layer %d: operands need 16-byte rows", l);
    if (int rc = make_tmap(&maps.a[l], in, M, L.K, ld_in, BM, 0)) return rc;
    g.layer[l].C = L.Y; g.layer[l].bias = L.bias; g.layer[l].ldc = L.ldy; g.layer[l].N = L.N; g.layer[l].K = L.K; g.layer[l].act = L.act;
    in = L.Y;
    ld_in = L.ldy;
    max_n = L.N > max_n ? L.N : max_n;
  }
  // one tile width for the whole chain: 64 columns, two CTAs per SM (the rollout's problems are latency-bound; narrow tiles
  // give every layer of a 4096-row batch >= 32 tiles); 32 for chains no wider than 32
  const int bn = max_n <= 32 ? 32 : 64;
  for (int l = 0; l < num_layers; ++l)
    if (int rc = make_tmap(&maps.b[l], layers[l].W, layers[l].N, layers[l].K, layers[l].ldw, bn, 0)) return rc;
  const int m_tiles = (M + BM - 1) / BM;
  int max_tiles = 1;
  for (int l = 0; l < num_layers; ++l) {
    const int t = m_tiles * ((layers[l].N + bn - 1) / bn);
    max_tiles = t > max_tiles ? t : max_tiles;
  }
  // Every CTA of the grid must be RESIDENT (they wait for each other): an SM holds two of these CTAs, so all chain launches
  // that can run concurrently must together stay within 2 x SMs -- the caller splits that budget with `max_ctas`.
  cudaStream_t st = (cudaStream_t)stream;
  int workers = num_sms() * 2;
  if (max_ctas > 0 && max_ctas < workers) workers = max_ctas;
  workers = max_tiles < workers ? max_tiles : workers;
  if (bn == 32) return launch_chain<32, 2>(maps, g, workers, st, "tc_chain<32>");
  return launch_chain<64, 2>(maps, g, workers, st, "tc_chain<64>");
}

// dX[M,K] (+)= (dY[M,N] . W[N,K]) * elu'(Yprev);  dbias_prev[K] += column sums of dX (the bias gradient of the layer below)
int b200_tc_linear_dgrad_bias(const float* dY, int lddy, const float* W, int ldw, const float* Yprev, int ldyp, float* dX, int lddx, int M,
                              int N, int K, int accumulate, float* dbias_prev, void* stream);
int b200_tc_linear_dgrad(const float* dY, int lddy, const float* W, int ldw, const float* Yprev, int ldyp, float* dX, int lddx, int M,
                         int N, int K, int accumulate, void* stream) {
  return b200_tc_linear_dgrad_bias(dY, lddy, W, ldw, Yprev, ldyp, dX, lddx, M, N, K, accumulate, nullptr, stream);
}
int b200_tc_linear_dgrad_bias(const float* dY, int lddy, const float* W, int ldw, const float* Yprev, int ldyp, float* dX, int lddx, int M,
                              int N, int K, int accumulate, float* dbias_prev, void* stream) {
  B200_CHECK_ARG(dY && W && dX && M > 0 && N > 0 && K > 0, "b200_tc_linear_dgrad: bad argument");
  B200_CHECK_ARG(!(dbias_prev && accumulate), "b200_tc_linear_dgrad_bias: the fused bias gradient needs accumulate = 0");
  B200_CHECK_ARG(lddy % 4 == 0 && ldw % 4 == 0 && aligned16(dY) && aligned16(W), "b200_tc_linear_dgrad: operands need 16-byte rows");
  const int cps = cps_for(K);
  const int pbn = (g_pair_mode == 2 && cps != 2) ? pick_pair_bn(M, K) : 0;   // measured: pairs do not pay for the dgrads (epilogue-bound); 2 = force
  int bn = pbn ? pbn : pick_bn(M, K);
  if (cps == 2 && bn > 128) bn = 128;
  CUtensorMap ta, tb;
  if (int rc = make_tmap(&ta, dY, M, N, lddy, BM, 0)) return rc;          // A K-major: [M rows][N reduction]
  if (int rc = make_tmap(&tb, W, N, K, ldw, BK, 1)) return rc;            // B MN-major: box [32 n][32 k]
  TcArgs g{};
  g.C = dX; g.aux = Yprev; g.ldc = lddx; g.ldaux = ldyp; g.M = M; g.N = K; g.K = N; g.accumulate = accumulate; g.dbias = dbias_prev;
  cudaStream_t st = (cudaStream_t)stream;
  // TMA epilogue: whole 32 x 32 boxes of the stored activation in, of dX out (needs 16-byte rows on both, no accumulate).
  // Measured on B200 at M = 24576 (with the fused bias gradient): 256 -> 512: 40.8 -> 33.2 us, 128 -> 256: 16.9 -> 14.9 us; not
  // for long reductions (N = 512: the epilogue buffers cost ring stages, 59.7 -> 64.7 us) and not for small M (two more
  // tensor maps to encode per launch).
  if (g_tma_epi && !accumulate && K >= 64 && N <= 256 && M >= 2048 && lddx % 4 == 0 && aligned16(dX) &&
      (!Yprev || (ldyp % 4 == 0 && aligned16(Yprev)))) {
    CUtensorMap taux, tout;
    if (int rc = make_tmap(&tout, dX, M, K, lddx, 32, 0)) return rc;
    if (Yprev) {
      if (int rc = make_tmap(&taux, Yprev, M, K, ldyp, 32, 0)) return rc;
    } else {
      taux = tout;
    }
    // CTA pairs (each CTA stages half of W) measured no better with this epilogue either (256 -> 512: 33.3 us both ways,
    // 128 -> 256: 14.8 -> 17.0 us): only when forced (pair mode 2, A/B runs)
    const int epbn = g_pair_mode == 2 ? pick_pair_bn(M, K) : 0;
    if (epbn == 256) return launch_tc<NN, 256, true, 1, true>(ta, tb, g, 1, st, "tc_dgrad_tma_pair<256>", &taux, &tout);
    if (epbn == 128) return launch_tc<NN, 128, true, 1, true>(ta, tb, g, 1, st, "tc_dgrad_tma_pair<128>", &taux, &tout);
    const int ebn = pick_bn(M, K) >= 128 ? pick_bn(M, K) : 64;
    switch (ebn) {
      case 256: return launch_tc<NN, 256, false, 1, true>(ta, tb, g, 1, st, "tc_dgrad_tma<256>", &taux, &tout);
      case 128: return launch_tc<NN, 128, false, 1, true>(ta, tb, g, 1, st, "tc_dgrad_tma<128>", &taux, &tout);
      default: return launch_tc<NN, 64, false, 1, true>(ta, tb, g, 1, st, "tc_dgrad_tma<64>", &taux, &tout);
    }
  }
  switch (pbn) {
    case 256: return launch_tc<NN, 256, true>(ta, tb, g, 1, st, "tc_dgrad_pair<256>");
    case 128: return launch_tc<NN, 128, true>(ta, tb, g, 1, st, "tc_dgrad_pair<128>");
    case 64: return launch_tc<NN, 64, true>(ta, tb, g, 1, st, "tc_dgrad_pair<64>");
    default: break;
  }
  if (cps == 2) switch (bn) {
      case 128: return launch_tc<NN, 128, false, 2>(ta, tb, g, 1, st, "tc_dgrad<128,2/SM>");
      case 64: return launch_tc<NN, 64, false, 2>(ta, tb, g, 1, st, "tc_dgrad<64,2/SM>");
      default: return launch_tc<NN, 32, false, 2>(ta, tb, g, 1, st, "tc_dgrad<32,2/SM>");
    }
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
  const bool pair = g_wgrad_pairs && N >= 2 * BM && bn >= 64;
  const int tiles = ((N + (pair ? 2 : 1) * BM - 1) / ((pair ? 2 : 1) * BM)) * ((K + bn - 1) / bn);
  int splits = (avail_sms((cudaStream_t)stream) / (pair ? 2 : 1)) / (tiles < 1 ? 1 : tiles);
  const int max_splits = (M + 8 * BK - 1) / (8 * BK);
  splits = splits < 1 ? 1 : (splits > max_splits ? max_splits : splits);
  int per = (M + splits - 1) / splits;
  per = (per + BK - 1) / BK * BK;
  g.k_per_split = per;
  splits = (M + per - 1) / per;
  cudaStream_t st = (cudaStream_t)stream;
  if (pair) switch (bn) {
      case 256: return launch_tc<TN, 256, true>(ta, tb, g, splits, st, "tc_wgrad_pair<256>");
      case 128: return launch_tc<TN, 128, true>(ta, tb, g, splits, st, "tc_wgrad_pair<128>");
      default: return launch_tc<TN, 64, true>(ta, tb, g, splits, st, "tc_wgrad_pair<64>");
    }
  switch (bn) {
    case 256: return launch_tc<TN, 256>(ta, tb, g, splits, st, "tc_wgrad<256>");
    case 128: return launch_tc<TN, 128>(ta, tb, g, splits, st, "tc_wgrad<128>");
    case 64: return launch_tc<TN, 64>(ta, tb, g, splits, st, "tc_wgrad<64>");
    default: return launch_tc<TN, 32>(ta, tb, g, splits, st, "tc_wgrad<32>");
  }
}

}  // extern "C"
