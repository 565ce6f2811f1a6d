// tcgen05 / TMA / TMEM GEMM core for sm_100a (hand-written PTX, no CUTLASS on the product path).
//
//   D[m, n] = sum_k A[m, k] * B[n, k]       A (M x K) and B (N x K) bf16, both K-major ("TN")
//
// One persistent CTA per SM, 64 + 32*EW threads, warp-specialised:
//   warp 0  lane 0 : TMA producer   cp.async.bulk.tensor.2d -> 128B-swizzled smem ring (STAGES deep)
//   warp 1  lane 0 : MMA issuer     tcgen05.mma.cta_group::1.kind::f16, fp32 accumulators in TMEM
//   warps 2..1+EW  : epilogue       tcgen05.ld (32x32b): one thread = one accumulator ROW, so row
//                                   reductions (max / sum-exp / dot) need no shuffles.  EW = 4, or 8 when the
//                                   epilogue (exp2 / pack / stores per element) is heavier than the tile's MMAs:
//                                   the two warps of a TMEM lane quarter then split the tile's columns
// Three pipelines: smem full/empty (TMA<->MMA), TMEM full/empty (MMA<->epilogue; ACC_STAGES
// accumulators of BLOCK_N fp32 columns each so the epilogue of tile i overlaps the MMAs of tile
// i+1), and the static persistent tile loop (m fastest so that concurrently running CTAs share
// the same B tile in L2).  Split-K work units are supported for short-M / long-K contractions.
//
// Layout contract: tile rows are 64 bf16 = 128 bytes, TMA writes them with SWIZZLE_128B and the
// UMMA smem descriptors read them as the canonical K-major SW128 layout (8-row x 128 B atoms,
// SBO = 1024 B).  K-steps inside the 128-byte atom advance the descriptor start address by 32 B.
// Out-of-bounds rows / k are zero-filled by TMA, so M, N, K need not be tile multiples.
//
// Either operand may instead be MN-major (A stored (K x M), B stored (K x N), the MN index
// contiguous): TMA then fetches 64(mn) x 64(k) boxes (128-byte rows again, SWIZZLE_128B) laid out
// box after box, and the UMMA descriptor describes the canonical MN-major SW128 layout
// (LBO = 8 KB between 64-wide MN blocks, SBO = 1 KB between 8-deep K groups; a K-step of 16
// advances the start address by 2 KB).  This is what lets dX = dcos Wn and dW = dcos^T X read
// dcos, Wn and X in place, without transposed copies.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace msml {
namespace tc {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;          // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kEpiWarp0 = 2;
constexpr int kTmemCols = 512;
constexpr uint32_t kSpinLimit = 1u << 26;  // bounded waits: trap instead of hanging the GPU

// ------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` when every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- warp-converged issue: every lane of the issuing warp executes these, ONE elected lane performs the operation.
// The round-1 kernels branched on `lane == 0` around the whole role; ptxas then wraps every uniform-datapath instruction
// (UTCHMMA, UTCBAR, UTMALDG and the descriptor arithmetic feeding them) in an ELECT / BRA.U.ANY loop, ~25 SASS instructions
// per MMA: the issuing thread needed ~750 cycles per k-block for 512 cycles of tensor work and WAS the limiter (ncu: the MMA
// thread never waits on a barrier, tensor pipe 68 % busy; profiles/r02_ncu_head_summary.md).
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair_elect(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// issue only (no wait): pair with tmem_ld_wait(v) which also pins the data dependency
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
        "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]),
        "=f"(r[16]), "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]),
        "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
      : "r"(taddr)
      : "memory");
}
// wait for all outstanding tcgen05.ld of this thread; the registers pass through as in/out operands
// so no consumer can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait(float* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
                 "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]),
                 "+f"(r[16]), "+f"(r[17]), "+f"(r[18]), "+f"(r[19]), "+f"(r[20]), "+f"(r[21]), "+f"(r[22]), "+f"(r[23]),
                 "+f"(r[24]), "+f"(r[25]), "+f"(r[26]), "+f"(r[27]), "+f"(r[28]), "+f"(r[29]), "+f"(r[30]), "+f"(r[31])
               :
               : "memory");
}

// ---- epilogue-side async copies ---------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
  const uint32_t bytes = pred ? 16u : 0u;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// byte offset of 16-byte chunk `c` of row `r` in a 128-byte-row SWIZZLE_128B tile
__device__ __forceinline__ uint32_t swz128(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// K-major, SWIZZLE_128B canonical layout: start>>4 | LBO(16B, ignored)=1 | SBO=1024B | version=1 | layout=2
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // [0,14)  start address
  d |= (uint64_t)1 << 16;                             // [16,30) leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                   // [32,46) stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                             // [46,48) descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                             // [61,64) SWIZZLE_128B
  return d;
}
// MN-major, SWIZZLE_128B canonical layout: 64(mn) x 64(k) boxes of 8 KB; LBO = 8 KB (next MN block),
// SBO = 1 KB (next group of 8 k)
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, M x N, per-operand major (0 = K, 1 = MN)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn = false, bool b_mn = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------- problem
struct GemmShape {
  int M, N, K;           // logical extents (TMA zero-fills beyond them)
  int m_blocks, n_blocks;
  int k_splits;          // split-K work units
  int k_blocks_per_split;
  int n_fastest;         // unit order: 0 = m fastest (CTAs running together share the B tile in L2), 1 = n fastest (share A)
  int a_blk_pitch;       // A_BLOCKED only: rows per 64-column block of the tile-blocked A matrix (see below)
};

// Tile-blocked A operand (the head's dcos matrix, written by one GEMM epilogue and read by two other GEMMs, once
// along each of its axes): a logical (R x C) bf16 matrix stored as [C/64][R_pad][64] — for every block of 64 columns,
// all rows back to back, 128 bytes each.  Seen by TMA as a 2-D array (inner 64, outer C/64 * R_pad), so that
//   * the K-major reader  (M = rows,    K = columns): 128 rows x 64 k   = ONE contiguous 16 KB box at (0, kb * R_pad + m0)
//   * the MN-major reader (M = columns, K = rows):    64 mn x 64 k      = ONE contiguous  8 KB box at (0, cb * R_pad + k0)
//   * the writer (32 rows x 64 columns per warp)                        = ONE contiguous  4 KB box
// and every DRAM access of all three kernels is a whole multi-KB burst, where the row-major layout gave the MN-major
// reader (and the writer) 128-byte pieces a full row pitch (250 KB at 125,000 classes) apart.  The shared-memory images
// are exactly the canonical SW128 K-major / MN-major layouts, so the UMMA descriptors do not change.

template <int BLOCK_N, int STAGES, int EPI_BYTES>
struct SmemLayout {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiOffset = STAGES * kStageBytes;          // 1024-aligned epilogue scratch
  static constexpr int kBarOffset = kEpiOffset + EPI_BYTES;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // barriers + tmem ptr, + slack for 1024-B alignment
  static_assert(EPI_BYTES % 1024 == 0, "epilogue scratch must keep 1024-byte alignment");
  static_assert(kTotal <= 232448, "exceeds 227 KB of shared memory");
};

// Epilogue concept:
//   struct Epi {
//     static constexpr int kSmemBytes;     // per-CTA scratch (multiple of 1024), split by the callee per warp
//     __device__ void operator()(uint32_t tmem_acc, int m_blk, int n_blk, int k_split,
//                                int quarter, int lane, uint8_t* scratch, int half, int n_halves) const;
//     __device__ void prefetch(int m_blk, int n_blk, int quarter, int lane, uint8_t* scratch, int half, int n_halves) const;  // before the tile's
//                                                             // accumulator is complete (may be a no-op)
//     __device__ void finish(int quarter, int lane) const;   // once per warp after the last tile
//   };
// tmem_acc already carries the accumulator-stage column offset; the callee adds
// ((quarter*32) << 16) + column.  Called by all epilogue threads (warp-convergent); with 8 epilogue warps
// (n_halves == 2) warp `half` of a quarter owns columns [half, half+1) * BLOCK_N / 2 of the tile.

template <int BLOCK_N, int ACC_STAGES, int STAGES, bool A_MN, bool B_MN, class Epi, int EW = 4, bool A_BLOCKED = false>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const GemmShape shape, const __grid_constant__ Epi epi) {
  static_assert(BLOCK_N == 128 || BLOCK_N == 256 || BLOCK_N == 512, "BLOCK_N in {128,256,512}");
  static_assert(BLOCK_N * ACC_STAGES <= kTmemCols, "accumulators exceed TMEM");
  constexpr int UMMA_N = BLOCK_N >= 256 ? 256 : BLOCK_N;
  constexpr int N_SUB = BLOCK_N / UMMA_N;
  using L = SmemLayout<BLOCK_N, STAGES, Epi::kSmemBytes>;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_units = shape.m_blocks * shape.n_blocks * shape.k_splits;
  const int total_k_blocks = (shape.K + kBlockK - 1) / kBlockK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], EW); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        const int mn = u % (shape.m_blocks * shape.n_blocks);
        const int m_blk = shape.n_fastest ? mn / shape.n_blocks : mn % shape.m_blocks;
        const int n_blk = shape.n_fastest ? mn % shape.n_blocks : mn / shape.m_blocks;
        const int ks = u / (shape.m_blocks * shape.n_blocks);
        const int kb0 = ks * shape.k_blocks_per_split;
        int kb1 = kb0 + shape.k_blocks_per_split;
        if (kb1 > total_k_blocks) kb1 = total_k_blocks;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          if (A_MN) {   // 64(m) x 64(k) boxes, one per 64-wide M block
#pragma unroll
            for (int j = 0; j < kBlockM / 64; ++j) {
              if (A_BLOCKED) tma_load_2d(sa + j * 8192, &map_a, &full_bar[stage], 0, (m_blk * (kBlockM / 64) + j) * shape.a_blk_pitch + kb * kBlockK);
              else tma_load_2d(sa + j * 8192, &map_a, &full_bar[stage], m_blk * kBlockM + j * 64, kb * kBlockK);
            }
          } else {
            if (A_BLOCKED) tma_load_2d(sa, &map_a, &full_bar[stage], 0, kb * shape.a_blk_pitch + m_blk * kBlockM);
            else tma_load_2d(sa, &map_a, &full_bar[stage], kb * kBlockK, m_blk * kBlockM);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(sb + j * 8192, &map_b, &full_bar[stage], n_blk * BLOCK_N + j * 64, kb * kBlockK);
          } else {
#pragma unroll
            for (int j = 0; j < N_SUB; ++j)
              tma_load_2d(sb + j * UMMA_N * kBlockK * 2, &map_b, &full_bar[stage], kb * kBlockK, n_blk * BLOCK_N + j * UMMA_N);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the whole warp walks the loop converged, one elected lane issues ==============
    constexpr uint32_t idesc = make_idesc(kBlockM, UMMA_N, A_MN, B_MN);
    // descriptor templates for stage 0, k-step 0: the start-address field holds (address >> 4) in its low 14 bits, so the
    // operand of another stage / k-step / N half is the template plus a byte offset >> 4 (shared memory is < 256 KB: no carry)
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t da0 = A_MN ? make_smem_desc_mn(smem_base) : make_smem_desc(smem_base);
    const uint64_t db0 = B_MN ? make_smem_desc_mn(smem_base + L::kABytes) : make_smem_desc(smem_base + L::kABytes);
    constexpr uint32_t kAStep = (A_MN ? 2048 : kUmmaK * 2) >> 4;          // K-major: +32 B per K-step; MN-major: +16 k-rows = 2 KB
    constexpr uint32_t kBStep = (B_MN ? 2048 : kUmmaK * 2) >> 4;
    constexpr uint32_t kBSub = (B_MN ? (UMMA_N / 64) * 8192 : UMMA_N * kBlockK * 2) >> 4;
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      const int ks = u / (shape.m_blocks * shape.n_blocks);
      const int kb0 = ks * shape.k_blocks_per_split;
      int kb1 = kb0 + shape.k_blocks_per_split;
      if (kb1 > total_k_blocks) kb1 = total_k_blocks;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t so = (uint64_t)((uint32_t)(stage * L::kStageBytes) >> 4);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
#pragma unroll
          for (int j = 0; j < N_SUB; ++j)
            umma_bf16_elect(d_tmem + j * UMMA_N, da0 + so + k * kAStep, db0 + so + j * kBSub + k * kBStep, idesc,
                            (kb > kb0 || k > 0) ? 1u : 0u);
        }
        umma_commit_elect(&empty_bar[stage]);                 // smem slot free once these MMAs retire
        if (kb == kb1 - 1) umma_commit_elect(&tmem_full[acc]);  // accumulator ready for the epilogue
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = (warp - kEpiWarp0) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      const int mn = u % (shape.m_blocks * shape.n_blocks);
      const int m_blk = shape.n_fastest ? mn / shape.n_blocks : mn % shape.m_blocks;
      const int n_blk = shape.n_fastest ? mn % shape.n_blocks : mn / shape.m_blocks;
      const int ks = u / (shape.m_blocks * shape.n_blocks);
      epi.prefetch(m_blk, n_blk, quarter, lane, smem + L::kEpiOffset, half, EW / 4);   // operand fetches that need not wait for the MMAs
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epi(tmem_base + acc * BLOCK_N, m_blk, n_blk, ks, quarter, lane, smem + L::kEpiOffset, half, EW / 4);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    epi.finish(quarter, lane);
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------- CTA pairs (cta_group::2)
// Two CTAs of one cluster (the two SMs of a TPC) own ONE 256 x BLOCK_N tile: rank r holds rows [128 r, 128 r + 128) of A
// and rows [BLOCK_N/2 r, BLOCK_N/2 (r + 1)) of B in ITS shared memory, `tcgen05.mma.cta_group::2` (M = 256, issued by
// rank 0 only) reads both halves of B from the two shared memories, and each CTA's TMEM receives its own 128 rows x BLOCK_N
// columns.  Per k-block a CTA therefore pulls (128 + BLOCK_N/2) x 128 bytes through its L2 -> SM port instead of
// (128 + BLOCK_N) x 128 (-33 % at BLOCK_N = 256) and its tensor core reads half the B bytes from shared memory per flop:
// the two limits measured on the single-CTA kernels (0.70-0.73 of peak).  The epilogues are unchanged: a thread still owns
// one accumulator row of its CTA's 128.
//   full_bar[s]   lives in rank 0: its producer arms it with the bytes of BOTH CTAs, both producers' TMA loads complete on it
//                 (cp.async.bulk.tensor ... .cta_group::2 with the barrier address mapped into rank 0)
//   empty_bar[s]  one per CTA; `tcgen05.commit.cta_group::2 ... multicast::cluster` (mask 0b11) arrives on both
//   tmem_full[a]  one per CTA, same multicast commit;   tmem_empty[a] in rank 0, counts the epilogue warps of both CTAs
//                 (rank 1 arrives remotely)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_count_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at the same shared-memory offset in BOTH CTAs when every MMA issued so far has completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

template <int BLOCK_N, int STAGES, int EPI_BYTES>
struct SmemLayoutPair {
  static constexpr int kABytes = kBlockM * kBlockK * 2;            // this CTA's 128 rows of A
  static constexpr int kBBytes = (BLOCK_N / 2) * kBlockK * 2;      // this CTA's half of B
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiOffset = STAGES * kStageBytes;
  static constexpr int kBarOffset = kEpiOffset + EPI_BYTES;
  static constexpr int kTotal = kBarOffset + 256 + 1024;
  static_assert(EPI_BYTES % 1024 == 0, "epilogue scratch must keep 1024-byte alignment");
  static_assert(kTotal <= 232448, "exceeds 227 KB of shared memory");
};

// shape.m_blocks counts 128-row blocks as in the single-CTA kernel; a pair owns blocks (2 p, 2 p + 1).
template <int BLOCK_N, int ACC_STAGES, int STAGES, bool A_MN, bool B_MN, class Epi, int EW = 4, bool A_BLOCKED = false>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const GemmShape shape, const __grid_constant__ Epi epi) {
  static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N in {128,256} (cta_group::2: N <= 256)");
  static_assert(BLOCK_N * ACC_STAGES <= kTmemCols, "accumulators exceed TMEM");
  constexpr int HALF_N = BLOCK_N / 2;
  using L = SmemLayoutPair<BLOCK_N, STAGES, Epi::kSmemBytes>;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int m_pairs = (shape.m_blocks + 1) / 2;
  const int num_units = m_pairs * shape.n_blocks * shape.k_splits;
  const int total_k_blocks = (shape.K + kBlockK - 1) / kBlockK;
  const int first_unit = (int)cluster_id_x(), unit_step = (int)cluster_count_x();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * EW); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer (both CTAs) =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int u = first_unit; u < num_units; u += unit_step) {
        const int mn = u % (m_pairs * shape.n_blocks);
        const int pair = shape.n_fastest ? mn / shape.n_blocks : mn % m_pairs;
        const int n_blk = shape.n_fastest ? mn % shape.n_blocks : mn / m_pairs;
        const int m_blk = pair * 2 + (int)rank;
        const int ks = u / (m_pairs * shape.n_blocks);
        const int kb0 = ks * shape.k_blocks_per_split;
        int kb1 = kb0 + shape.k_blocks_per_split;
        if (kb1 > total_k_blocks) kb1 = total_k_blocks;
        const int n0 = n_blk * BLOCK_N + (int)rank * HALF_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
          const uint32_t bar = map_to_cta(smem_u32(&full_bar[stage]), 0);
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < kBlockM / 64; ++j) {
              if (A_BLOCKED) tma_load_2d_pair(sa + j * 8192, &map_a, bar, 0, (m_blk * (kBlockM / 64) + j) * shape.a_blk_pitch + kb * kBlockK);
              else tma_load_2d_pair(sa + j * 8192, &map_a, bar, m_blk * kBlockM + j * 64, kb * kBlockK);
            }
          } else {
            if (A_BLOCKED) tma_load_2d_pair(sa, &map_a, bar, 0, kb * shape.a_blk_pitch + m_blk * kBlockM);
            else tma_load_2d_pair(sa, &map_a, bar, kb * kBlockK, m_blk * kBlockM);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < HALF_N / 64; ++j) tma_load_2d_pair(sb + j * 8192, &map_b, bar, n0 + j * 64, kb * kBlockK);
          } else {
            tma_load_2d_pair(sb, &map_b, bar, kb * kBlockK, n0);        // box = HALF_N rows x 64 k
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================== MMA issuer (rank 0 only; whole warp converged, one elected lane issues) =================
      constexpr uint32_t idesc = make_idesc(2 * kBlockM, BLOCK_N, A_MN, B_MN);
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t da0 = A_MN ? make_smem_desc_mn(smem_base) : make_smem_desc(smem_base);
      const uint64_t db0 = B_MN ? make_smem_desc_mn(smem_base + L::kABytes) : make_smem_desc(smem_base + L::kABytes);
      constexpr uint32_t kAStep = (A_MN ? 2048 : kUmmaK * 2) >> 4;
      constexpr uint32_t kBStep = (B_MN ? 2048 : kUmmaK * 2) >> 4;
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int u = first_unit; u < num_units; u += unit_step) {
        const int ks = u / (m_pairs * shape.n_blocks);
        const int kb0 = ks * shape.k_blocks_per_split;
        int kb1 = kb0 + shape.k_blocks_per_split;
        if (kb1 > total_k_blocks) kb1 = total_k_blocks;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t so = (uint64_t)((uint32_t)(stage * L::kStageBytes) >> 4);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_bf16_pair_elect(d_tmem, da0 + so + k * kAStep, db0 + so + k * kBStep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit_pair_elect(&empty_bar[stage]);
          if (kb == kb1 - 1) umma_commit_pair_elect(&tmem_full[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int quarter = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = first_unit; u < num_units; u += unit_step) {
      const int mn = u % (m_pairs * shape.n_blocks);
      const int pair = shape.n_fastest ? mn / shape.n_blocks : mn % m_pairs;
      const int n_blk = shape.n_fastest ? mn % shape.n_blocks : mn / m_pairs;
      const int m_blk = pair * 2 + (int)rank;
      const int ks = u / (m_pairs * shape.n_blocks);
      const bool live = m_blk < shape.m_blocks;          // odd number of 128-row blocks: the last pair's second half is empty
      if (live) epi.prefetch(m_blk, n_blk, quarter, lane, smem + L::kEpiOffset, half, EW / 4);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (live) epi(tmem_base + acc * BLOCK_N, m_blk, n_blk, ks, quarter, lane, smem + L::kEpiOffset, half, EW / 4);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty[acc]), 0));
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    epi.finish(quarter, lane);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // nobody leaves (or frees TMEM) while the peer may still signal / read this CTA
  tc_fence_after();
  if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------- host side
// cuTensorMapEncodeTiled is fetched through the runtime (no link-time libcuda dependency).
// 2-D tensor map over a row-major (outer x inner) array: inner contiguous, row pitch ld_elems, 128-byte
// swizzled boxes of box_inner x box_outer elements (box_inner * elem_bytes must be 128).
int encode_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, int64_t inner, int64_t outer, int64_t ld_elems,
                   int box_inner, int box_outer);
// K-major bf16 operand (rows x k): boxes of 64(k) x box_rows
inline int encode_tmap_bf16_kmajor(CUtensorMap* out, const void* base, int64_t rows, int64_t k, int64_t ld_elems, int box_rows) {
  return encode_tmap_2d(out, base, 2, k, rows, ld_elems, kBlockK, box_rows);
}
// MN-major bf16 operand stored (k x mn): boxes of 64(mn) x 64(k)
inline int encode_tmap_bf16_mnmajor(CUtensorMap* out, const void* base, int64_t mn, int64_t k, int64_t ld_elems) {
  return encode_tmap_2d(out, base, 2, mn, k, ld_elems, 64, kBlockK);
}

template <int BLOCK_N, int ACC_STAGES, int STAGES, bool A_MN, bool B_MN, int EW = 4, bool A_BLOCKED = false, class Epi>
int launch_gemm(const char* name, const CUtensorMap& ma, const CUtensorMap& mb, const GemmShape& shape, const Epi& epi,
                cudaStream_t st, double min_bytes = 0.0) {
  using L = SmemLayout<BLOCK_N, STAGES, Epi::kSmemBytes>;
  auto kern = gemm_kernel<BLOCK_N, ACC_STAGES, STAGES, A_MN, B_MN, Epi, EW, A_BLOCKED>;
  static thread_local bool configured = false;   // per template instantiation
  if (!configured) {
    MSML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  const int units = shape.m_blocks * shape.n_blocks * shape.k_splits;
  int grid = num_sms();
  if (units < grid) grid = units;
  if (grid < 1) grid = 1;
  MSML_PROF2(name, 2.0 * shape.M * shape.N * shape.K, min_bytes, st);
  kern<<<grid, 64 + 32 * EW, L::kTotal, st>>>(ma, mb, shape, epi);
  MSML_LAUNCH_CHECK();
  return 0;
}

template <int BLOCK_N, int ACC_STAGES, int STAGES, bool A_MN, bool B_MN, int EW = 4, bool A_BLOCKED = false, class Epi>
int launch_gemm_pair(const char* name, const CUtensorMap& ma, const CUtensorMap& mb, const GemmShape& shape, const Epi& epi,
                     cudaStream_t st, double min_bytes = 0.0) {
  using L = SmemLayoutPair<BLOCK_N, STAGES, Epi::kSmemBytes>;
  auto kern = gemm_pair_kernel<BLOCK_N, ACC_STAGES, STAGES, A_MN, B_MN, Epi, EW, A_BLOCKED>;
  static thread_local bool configured = false;   // per template instantiation
  if (!configured) {
    MSML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  const int units = ((shape.m_blocks + 1) / 2) * shape.n_blocks * shape.k_splits;
  int pairs = num_sms() / 2;
  if (units < pairs) pairs = units;
  if (pairs < 1) pairs = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(64 + 32 * EW);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MSML_PROF2(name, 2.0 * shape.M * shape.N * shape.K, min_bytes, st);
  MSML_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, shape, epi));
  MSML_LAUNCH_CHECK();
  return 0;
}

inline GemmShape make_shape(int64_t M, int64_t N, int64_t K, int block_n, int k_splits = 1, bool n_fastest = false) {
  GemmShape s;
  s.M = (int)M; s.N = (int)N; s.K = (int)K;
  s.a_blk_pitch = 0;
  s.n_fastest = n_fastest ? 1 : 0;
  s.m_blocks = (int)((M + kBlockM - 1) / kBlockM);
  s.n_blocks = (int)((N + block_n - 1) / block_n);
  const int total_kb = (int)((K + kBlockK - 1) / kBlockK);
  if (k_splits < 1) k_splits = 1;
  if (k_splits > total_kb) k_splits = total_kb;
  s.k_blocks_per_split = (total_kb + k_splits - 1) / k_splits;
  s.k_splits = (total_kb + s.k_blocks_per_split - 1) / s.k_blocks_per_split;  // no empty splits
  return s;
}

}  // namespace tc
}  // namespace msml
