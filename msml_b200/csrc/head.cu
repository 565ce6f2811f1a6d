// K-E..K-H  PartialFC head on tcgen05 tensor cores.      ref headers/partial_fc.py:96-99,115,118-177
//                                                        ref headers/margin_losses.py:275-303,390-418
// Per rank (class shard):  X (B_tot, D) bf16 gathered features, Wn (n_s, D) bf16 unit class centres.
//
//   head_fwd   S = X Wn^T (tcgen05, fp32 in TMEM); epilogue: margin on the target column, * s,
//              online row max / sum-exp per (row, class tile)  -> partials -> merged row stats.
//              The (B_tot x n_s) logits matrix is NEVER written to memory.
//   merge      combines the row stats of all ranks (after one all-gather), computes the loss.
//   head_bwd   1. recompute S tile by tile; epilogue: p = softmax, grad = (p - smoothed one-hot)/B_tot,
//                 chain through margin and scale -> dcos in bf16 (row-major, the only materialised matrix),
//                 plus rdot[n] = <Wn[n], dWn[n]> by a warp transpose-reduce
//              2. dX  = dcos Wn        Wn read in place (MN-major), split-K over the class dimension sized to
//                 exactly one wave, fp32 red.add
//              3. dW  = normalize_bwd(dcos^T X): dcos and X read in place (MN-major), 128 x 256 tiles ordered so
//                 that the two halves of a class block run side by side (dcos tile fetched from HBM once);
//                 epilogue applies (dWn - Wn * rdot) / ||W|| and streams fp32 rows out through TMA.
// Everything runs on the caller's stream; no allocation, no host sync.
#include <cmath>
#include <cstdlib>

#include "tc_gemm.cuh"
#include "head_small_kernels.cuh"

namespace msml {
namespace tc {

// ------------------------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, int64_t inner, int64_t outer, int64_t ld_elems,
                   int box_inner, int box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  MSML_REQUIRE(fn != nullptr, MSML_EUNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  MSML_REQUIRE(elem_bytes == 2 || elem_bytes == 4, MSML_EINVAL, "tensor map element size %d", elem_bytes);
  MSML_REQUIRE(aligned16(base), MSML_EALIGN, "TMA operand base must be 16-byte aligned");
  MSML_REQUIRE((ld_elems * elem_bytes) % 16 == 0 && ld_elems >= inner, MSML_EALIGN,
               "TMA operand pitch %lld elements must be 16-byte aligned and >= inner extent %lld", (long long)ld_elems,
               (long long)inner);
  MSML_REQUIRE(inner > 0 && outer > 0 && box_inner * elem_bytes == 128 && box_outer > 0 && box_outer <= 256, MSML_EINVAL,
               "bad TMA operand shape");
  const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld_elems * elem_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MSML_REQUIRE(r == CUDA_SUCCESS, MSML_EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// margin math (Margin, margin_target, margin_target_grad) and the small kernels: head_small_kernels.cuh
__device__ __forceinline__ float fast_exp2(float x) {   // MUFU.EX2; -inf -> 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------- epilogues
// Shared structure: a thread owns accumulator row (quarter*32 + lane); the 32-column TMEM chunks are
// double-buffered in registers (the tcgen05.ld of chunk c+1 is in flight while chunk c is processed).
// Tile outputs leave through a per-warp swizzled smem staging buffer and a TMA store (fully coalesced
// 128-byte rows, out-of-range rows/columns clipped by the tensor map) — never through per-thread
// scattered global stores.

// pick / replace element `idx` (0..31, warp-divergent allowed) without dynamic register indexing
__device__ __forceinline__ float pick32(const float* v, int idx) {
  float r = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) r = (j == idx) ? v[j] : r;
  return r;
}
__device__ __forceinline__ void put32(float* v, int idx, float x) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = (j == idx) ? x : v[j];
}

struct NoScratch {
  static constexpr int kSmemBytes = 0;
  __device__ void prefetch(int, int, int, int, uint8_t*, int, int) const {}
  __device__ void finish(int, int) const {}
};

// plain store (tests / in-model heads): C[row, col] = acc
struct EpiStore : NoScratch {
  float* c;
  int64_t ldc;
  int M, N, block_n;
  __device__ void operator()(uint32_t tmem_acc, int m_blk, int n_blk, int, int quarter, int lane, uint8_t*, int half, int nh) const {
    const int row = m_blk * kBlockM + quarter * 32 + lane;
    const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
    for (int c0 = half * (block_n / nh); c0 < (half + 1) * (block_n / nh); c0 += 32) {
      float v[32];
      tmem_ld32(taddr + c0, v);
      const int col0 = n_blk * block_n + c0;
      if (row < M) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < N) c[(int64_t)row * ldc + col0 + j] = v[j];
      }
    }
  }
};

// forward: margin + scale + online (max, sum exp2) in the log2 domain
struct EpiFwdStats : NoScratch {
  const int64_t* tl;
  Margin mg;
  int B_tot, n_s, block_n;
  float* part_max;   // [n_blocks][B_tot]  (log2 domain: logit * log2e)
  float* part_sum;   // [n_blocks][B_tot]
  float* tgt;        // [B_tot] target logit (natural units), written by the tile that owns the column
  __device__ void operator()(uint32_t tmem_acc, int m_blk, int n_blk, int, int quarter, int lane, uint8_t*, int half, int nh) const {
    const int row = m_blk * kBlockM + quarter * 32 + lane;
    const bool row_ok = row < B_tot;
    const int64_t label = row_ok ? tl[row] : -1;
    const float s2 = mg.s * kLog2e;
    const int cph = block_n / 32 / nh;                     // 32-column chunks per epilogue warp
    const int tile_col0 = n_blk * block_n + half * cph * 32;
    const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16) + half * cph * 32;
    int n_chunks = (n_s - tile_col0 + 31) / 32;            // warp-uniform; <= 0 when this half lies beyond n_s
    if (n_chunks > cph) n_chunks = cph;
    float run_max = -INFINITY, run_sum = 0.f;
    if (n_chunks > 0) {
      float buf[2][32];
      tmem_ld32_issue(taddr, buf[0]);
      tmem_ld_wait(buf[0]);
#pragma unroll 1
      for (int c = 0; c < n_chunks; c += 2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (c + h >= n_chunks) break;
          float* cur = buf[h];
          if (c + h + 1 < n_chunks) tmem_ld32_issue(taddr + (c + h + 1) * 32, buf[h ^ 1]);
          float* v = cur;
          const int col0 = tile_col0 + (c + h) * 32;
          const int64_t rel = label - col0;
          if (rel >= 0 && rel < 32) {                       // at most one row-chunk pair per row
            const float t = margin_target(mg, pick32(v, (int)rel));
            put32(v, (int)rel, t);
            tgt[row] = t * mg.s;
          }
          const int valid = n_s - col0;                      // >= 1
          if (valid < 32) {                                  // ragged last chunk (warp-uniform, once per row block)
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (j < valid) ? v[j] : -INFINITY;
          }
          // max(s2 * v) = s2 * max(v) (s > 0): FMNMX on the raw cosines, four independent chains
          float mx[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
          for (int j = 4; j < 32; ++j) mx[j & 3] = fmaxf(mx[j & 3], v[j]);
          const float cmax = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * s2;
          const float new_max = fmaxf(run_max, cmax);        // finite: column col0 < n_s exists
          // exp2(s2 * v - new_max): one FFMA + EX2 + FADD per element, four independent sums (-inf padding gives 0)
          float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; ++j) a4[j & 3] += fast_exp2(fmaf(v[j], s2, -new_max));
          run_sum = run_sum * fast_exp2(run_max - new_max) + ((a4[0] + a4[1]) + (a4[2] + a4[3]));
          run_max = new_max;
          if (c + h + 1 < n_chunks) tmem_ld_wait(buf[h ^ 1]);
        }
      }
    }
    if (row_ok) {       // one partial per (class tile, half); an empty half contributes (-inf, 0)
      part_max[(int64_t)(n_blk * nh + half) * B_tot + row] = run_max;
      part_sum[(int64_t)(n_blk * nh + half) * B_tot + row] = run_sum;
    }
  }
};

// per-warp staging: 2 buffers of 4 KB (32 rows x 128 B, SWIZZLE_128B) feeding TMA stores
struct StageOut {
  static constexpr int kBytesPerWarp = 2 * 4096;
  // store this thread's 128-byte row `lane` (8 x uint4) into buffer `b`, then TMA-store the 32-row box
  __device__ static __forceinline__ void put_and_store(uint8_t* warp_scratch, int b, int lane, const uint4* row8,
                                                       const CUtensorMap* map, int c0, int c1) {
    uint8_t* buf = warp_scratch + b * 4096;
    if (lane == 0) bulk_wait_read<1>();    // the store that last used this buffer (2 stores ago) has read it
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(buf + swz128(lane, q)) = row8[q];
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) { tma_store_2d(map, buf, c0, c1); bulk_commit(); }
  }
  __device__ static __forceinline__ void drain(int lane) {
    if (lane == 0) bulk_wait<0>();
    __syncwarp();
  }
};

// backward pass 1: recompute logits, emit dcos (bf16, row-major) and rdot[n] = sum_m dcos[m,n] cos[m,n] (= <Wn[n], dWn[n]>
// for the normalise backward).  The inner loop is the bare softmax gradient (FFMA, EX2, FFMA, FMUL per element): the target
// column is patched afterwards (rare, one element per row) and ragged tiles take a separate path, so that the epilogue of
// a tile issues fewer instructions than its MMAs take cycles.
// WITH_RDOT = false: the normalise-backward projection is applied by the fused optimizer (msml_pfc_sgd_update_raw), so the
// <Wn, dWn> column reduction — a third of this epilogue's instructions — is not needed
template <bool WITH_RDOT>
struct EpiBwdDcosT {
  static constexpr int kSmemBytes = 8 * StageOut::kBytesPerWarp;   // 64 KB (up to 8 epilogue warps)
  CUtensorMap map_dcos;        // bf16 (B_tot x n_s), boxes 64 cols x 32 rows
  const int64_t* tl;
  Margin mg;
  int B_tot, n_s, block_n;
  const float* gmax;   // [B_tot] global row max   (natural units)
  const float* gsum;   // [B_tot] global row sum of exp(logit - max)
  float* rdot;         // [n_s] zero-initialised; <Wn[n], dWn[n]> for the normalise backward
  float smooth_on, smooth_off, inv_btot;
  int blk_pitch;       // > 0: dcos is tile-blocked [n_s/64][blk_pitch rows][64] (tc_gemm.cuh); 0: row-major (B_tot x ld)
  __device__ void prefetch(int, int, int, int, uint8_t*, int, int) const {}
  __device__ void finish(int, int lane) const { StageOut::drain(lane); }
  __device__ void operator()(uint32_t tmem_acc, int m_blk, int n_blk, int, int quarter, int lane, uint8_t* scratch, int half, int nh) const {
    const int row0 = m_blk * kBlockM + quarter * 32;
    const int row = row0 + lane;
    const bool row_ok = row < B_tot;
    const int64_t label = row_ok ? tl[row] : -1;
    const int cph = block_n / 32 / nh;                     // 32-column chunks per epilogue warp (even)
    const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16) + half * cph * 32;
    uint8_t* wscr = scratch + (half * 4 + quarter) * StageOut::kBytesPerWarp;
    const float s2 = mg.s * kLog2e;
    // p = exp2(logit*log2e - off),  off = max*log2e + log2(sum);  dead rows: p = 0 and gs = 0
    const float neg_off = row_ok ? -fmaf(gmax[row], kLog2e, log2f(gsum[row])) : -INFINITY;
    const float gs = row_ok ? mg.s * inv_btot : 0.f;
    const float tg = (label >= 0 ? smooth_off : 0.f) * gs;   // rows without a local target: no one-hot row (ref :166)
    const int tile_col0 = n_blk * block_n + half * cph * 32;
    int n_chunks = (n_s - tile_col0 + 31) / 32;
    if (n_chunks > cph) n_chunks = cph;
    if (n_chunks <= 0) return;
    float buf[2][32];
    uint4 packed[8];
    tmem_ld32_issue(taddr, buf[0]);
    tmem_ld_wait(buf[0]);
#pragma unroll 1
    for (int c = 0; c < n_chunks; c += 2) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (c + h < n_chunks) {
          float* v = buf[h];
          if (c + h + 1 < n_chunks) tmem_ld32_issue(taddr + (c + h + 1) * 32, buf[h ^ 1]);
          const int col0 = tile_col0 + (c + h) * 32;
          const int64_t rel = label - col0;
          const bool has_t = rel >= 0 && rel < 32;
          float cos_t = 0.f;
          if (has_t) cos_t = pick32(v, (int)rel);
          float pr[WITH_RDOT ? 32 : 1];                      // dcos * raw cosine
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float g = fmaf(fast_exp2(fmaf(v[j], s2, neg_off)), gs, -tg);
            if constexpr (WITH_RDOT) pr[j] = g * v[j];
            v[j] = g;
          }
          if (has_t) {                                       // the one target element of this row (warp-divergent, rare)
            const float p = fast_exp2(fmaf(margin_target(mg, cos_t), s2, neg_off));
            const float g = (p - smooth_on) * gs * margin_target_grad(mg, cos_t);
            put32(v, (int)rel, g);
            if constexpr (WITH_RDOT) put32(pr, (int)rel, g * cos_t);
          }
          if (col0 + 32 > n_s) {                             // ragged last tile (warp-uniform)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] = (col0 + j < n_s) ? v[j] : 0.f;
              if constexpr (WITH_RDOT) pr[j] = (col0 + j < n_s) ? pr[j] : 0.f;
            }
          }
          if constexpr (WITH_RDOT) {
          // column sums over the 32 rows of this warp, then one red per lane.  The products are packed to bf16x2 first
          // (rdot = <Wn, dWn> is the projection term of the normalise backward; bf16 products leave it exact to ~2e-3, a
          // fifth of the tolerance of the bf16 dcos it corrects), so the recursive halving moves 16 packed values: 16 shuffles
          // + 16 HADD2.BF16 + 30 selects instead of 31 + 31 + 62 on fp32 — the epilogue, not the MMAs, bounds this kernel.
          uint32_t q2[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const __nv_bfloat162 t = __floats2bfloat162_rn(pr[2 * i], pr[2 * i + 1]);
            q2[i] = *reinterpret_cast<const uint32_t*>(&t);
          }
#pragma unroll
          for (int n = 8; n >= 1; n >>= 1) {
            const int sft = 2 * n;
            const bool up = (lane & sft) != 0;
#pragma unroll
            for (int i = 0; i < n; ++i) {
              const uint32_t send = up ? q2[i] : q2[i + n];
              const uint32_t keep = up ? q2[i + n] : q2[i];
              const uint32_t got = __shfl_xor_sync(0xffffffffu, send, sft);
              const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&keep), *reinterpret_cast<const __nv_bfloat162*>(&got));
              q2[i] = *reinterpret_cast<const uint32_t*>(&r);
            }
          }
          {   // lanes 2p and 2p+1 hold the same column pair (2p, 2p+1), summed over complementary halves of the warp
            const uint32_t got = __shfl_xor_sync(0xffffffffu, q2[0], 1);
            const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&q2[0]), *reinterpret_cast<const __nv_bfloat162*>(&got));
            const float colsum = (lane & 1) ? __high2float(r) : __low2float(r);
            if (col0 + lane < n_s) atomicAdd(rdot + col0 + lane, colsum);
          }
          }   // WITH_RDOT
          // bf16 pack: 32 values = 64 bytes = 4 x uint4 -> half `h` of the 128-byte staging row
#pragma unroll
          for (int q = 0; q < 4; ++q) packed[h * 4 + q] = Vec<__nv_bfloat16>::pack(v + q * 8);
          if (c + h + 1 < n_chunks) tmem_ld_wait(buf[h ^ 1]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) packed[h * 4 + q] = make_uint4(0, 0, 0, 0);
        }
      }
      if (blk_pitch) StageOut::put_and_store(wscr, (c >> 1) & 1, lane, packed, &map_dcos, 0, ((tile_col0 + c * 32) >> 6) * blk_pitch + row0);
      else StageOut::put_and_store(wscr, (c >> 1) & 1, lane, packed, &map_dcos, tile_col0 + c * 32, row0);
    }
  }
};

// backward pass 2: dX += partial (split-K over classes)
struct EpiDxAccum : NoScratch {
  float* dx;   // (B_tot, D) fp32, zero-initialised
  int B_tot, D, block_n;
  __device__ void operator()(uint32_t tmem_acc, int m_blk, int n_blk, int, int quarter, int lane, uint8_t*, int half, int nh) const {
    const int row = m_blk * kBlockM + quarter * 32 + lane;
    const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
    for (int c0 = half * (block_n / nh); c0 < (half + 1) * (block_n / nh); c0 += 32) {
      float v[32];
      tmem_ld32(taddr + c0, v);
      const int col0 = n_blk * block_n + c0;
      if (col0 >= D) break;
      if (row < B_tot) {
        float* p = dx + (int64_t)row * D + col0;
        if (col0 + 32 <= D) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + q * 4), "f"(v[q * 4]), "f"(v[q * 4 + 1]),
                         "f"(v[q * 4 + 2]), "f"(v[q * 4 + 3]) : "memory");
        } else {
          for (int j = 0; j < 32 && col0 + j < D; ++j) atomicAdd(p + j, v[j]);
        }
      }
    }
  }
};

// backward pass 3: dW = (dWn - Wn * rdot) * inv_norm      (ref :115 normalize backward), one pass.
// Wn tile rows arrive through cp.async (coalesced 128-byte rows -> swizzled smem, 2 boxes in flight; the first two are
// requested before the tile's MMAs have finished), dW leaves through TMA stores.  rdot = <Wn, dWn> comes from pass 1.
template <int NW>               // epilogue warps: 4 (each owns all 256 columns of its rows) or 8 (two per TMEM lane quarter, 128 columns each)
struct EpiDwNormBwdT {
  static constexpr int kWnBytesPerWarp = 2 * 4096;                  // 2 boxes of 32 rows x 64 bf16
  static constexpr int kSmemBytes = NW * (StageOut::kBytesPerWarp + kWnBytesPerWarp);   // 64 / 128 KB
  CUtensorMap map_dw;          // fp32 (n_s x D), boxes 32 cols x 32 rows
  const __nv_bfloat16* wn;     // (n_s, D)
  const float* inv_norm;       // (n_s)
  const float* rdot;           // (n_s)
  int n_s, D, block_n;         // block_n: D columns per tile
  __device__ void finish(int, int lane) const { StageOut::drain(lane); }
  // warp-cooperative async load of Wn[row0 .. row0+32, col0 .. col0+64) into a swizzled 4 KB box
  __device__ __forceinline__ void load_wn_box(uint8_t* box, int row0, int col0, int lane) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + (lane >> 3), ch = lane & 7;
      const bool ok = row0 + r < n_s;
      cp_async16(box + swz128(r, ch), wn + (int64_t)(ok ? row0 + r : 0) * D + col0 + ch * 8, ok);
    }
    cp_async_commit();
  }
  __device__ void prefetch(int m_blk, int n_blk, int quarter, int lane, uint8_t* scratch, int half, int nh) const {
    uint8_t* wwn = scratch + NW * StageOut::kBytesPerWarp + (half * 4 + quarter) * kWnBytesPerWarp;
    const int row0 = m_blk * kBlockM + quarter * 32, dcol0 = n_blk * block_n + half * (block_n / nh);
    if (dcol0 < D) load_wn_box(wwn, row0, dcol0, lane);
    if (D - dcol0 > 64) load_wn_box(wwn + 4096, row0, dcol0 + 64, lane);
  }
  __device__ void operator()(uint32_t tmem_acc, int m_blk, int n_blk, int, int quarter, int lane, uint8_t* scratch, int half, int nh) const {
    const int row0 = m_blk * kBlockM + quarter * 32;
    const int row = row0 + lane;                     // class index
    const int cols = block_n / nh;                   // columns this warp owns
    const int dcol0 = n_blk * block_n + half * cols; // ... [dcol0, dcol0 + cols) of D
    const bool ok = row < n_s;
    const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16) + half * cols;
    uint8_t* wout = scratch + (half * 4 + quarter) * StageOut::kBytesPerWarp;
    uint8_t* wwn = scratch + NW * StageOut::kBytesPerWarp + (half * 4 + quarter) * kWnBytesPerWarp;
    const float inv = ok ? inv_norm[row] : 0.f;
    const float cdot = ok ? rdot[row] * inv : 0.f;   // (acc - w*rdot)*inv = acc*inv - w*cdot
    int n_boxes = (D - dcol0) / 64;
    if (n_boxes > cols / 64) n_boxes = cols / 64;
    if (n_boxes <= 0) return;
    float buf[2][32];
    tmem_ld32_issue(taddr, buf[0]);                  // boxes 0 and 1 of Wn are already in flight (prefetch())
    tmem_ld_wait(buf[0]);
#pragma unroll 1
    for (int bx = 0; bx < n_boxes; ++bx) {
      if (bx + 1 < n_boxes) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncwarp();
      const uint8_t* wbox = wwn + (bx & 1) * 4096;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = bx * 2 + h;
        float* cur = buf[h];
        if (c + 1 < 2 * n_boxes) tmem_ld32_issue(taddr + (c + 1) * 32, buf[h ^ 1]);
        const float* v = cur;
        float w[32];
#pragma unroll
        for (int q = 0; q < 4; ++q) Vec<__nv_bfloat16>::unpack(*reinterpret_cast<const uint4*>(wbox + swz128(lane, h * 4 + q)), w + q * 8);
        uint4 out[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 r;
          r.x = fmaf(v[q * 4 + 0], inv, -w[q * 4 + 0] * cdot);
          r.y = fmaf(v[q * 4 + 1], inv, -w[q * 4 + 1] * cdot);
          r.z = fmaf(v[q * 4 + 2], inv, -w[q * 4 + 2] * cdot);
          r.w = fmaf(v[q * 4 + 3], inv, -w[q * 4 + 3] * cdot);
          out[q] = make_uint4(__float_as_uint(r.x), __float_as_uint(r.y), __float_as_uint(r.z), __float_as_uint(r.w));
        }
        StageOut::put_and_store(wout, c & 1, lane, out, &map_dw, dcol0 + c * 32, row0);
        if (c + 1 < 2 * n_boxes) tmem_ld_wait(buf[h ^ 1]);
      }
      __syncwarp();   // every lane is done with this Wn box before it is refilled
      if (bx + 2 < n_boxes) load_wn_box(wwn + (bx & 1) * 4096, row0, dcol0 + (bx + 2) * 64, lane);
    }
  }
};

// dW in raw mode: dWn = dcos^T X as it leaves the tensor core (fp32), no projection, no 1/||w||: a plain TMEM -> TMA-store copy
struct EpiDwRaw {
  static constexpr int kSmemBytes = 8 * StageOut::kBytesPerWarp;    // 64 KB (8 epilogue warps)
  CUtensorMap map_dw;          // fp32 (n_s x D), boxes 32 cols x 32 rows
  int n_s, D, block_n;
  __device__ void prefetch(int, int, int, int, uint8_t*, int, int) const {}
  __device__ void finish(int, int lane) const { StageOut::drain(lane); }
  __device__ void operator()(uint32_t tmem_acc, int m_blk, int n_blk, int, int quarter, int lane, uint8_t* scratch, int half, int nh) const {
    const int row0 = m_blk * kBlockM + quarter * 32;
    const int cols = block_n / nh;
    const int dcol0 = n_blk * block_n + half * cols;
    const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16) + half * cols;
    uint8_t* wout = scratch + (half * 4 + quarter) * StageOut::kBytesPerWarp;
    int n_chunks = (D - dcol0) / 32;
    if (n_chunks > cols / 32) n_chunks = cols / 32;
    if (n_chunks <= 0) return;
    float buf[2][32];
    tmem_ld32_issue(taddr, buf[0]);
    tmem_ld_wait(buf[0]);
#pragma unroll 1
    for (int c = 0; c < n_chunks; c += 2) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (c + h >= n_chunks) break;
        const float* v = buf[h];
        if (c + h + 1 < n_chunks) tmem_ld32_issue(taddr + (c + h + 1) * 32, buf[h ^ 1]);
        uint4 out[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
          out[q] = make_uint4(__float_as_uint(v[q * 4]), __float_as_uint(v[q * 4 + 1]), __float_as_uint(v[q * 4 + 2]), __float_as_uint(v[q * 4 + 3]));
        StageOut::put_and_store(wout, (c + h) & 1, lane, out, &map_dw, dcol0 + (c + h) * 32, row0);
        if (c + h + 1 < n_chunks) tmem_ld_wait(buf[h ^ 1]);
      }
    }
  }
};

static int to_margin(const msml_margin_params* p, Margin* out) {
  MSML_REQUIRE(p != nullptr, MSML_EINVAL, "margin params missing");
  MSML_REQUIRE(p->kind == MSML_MARGIN_ARC || p->kind == MSML_MARGIN_COS, MSML_EINVAL, "margin kind error (%d)", p->kind);
  MSML_REQUIRE(p->s > 0.f, MSML_EINVAL, "margin scale s must be positive (got %g)", (double)p->s);
  out->kind = p->kind; out->s = p->s; out->m = p->m; out->a = p->a; out->k = p->k;
  return 0;
}

// workspace carving
struct HeadWs {
  float* part_max; float* part_sum; float* tgt; float* rdot;
  __nv_bfloat16* dcos;
  int64_t ld_dc;          // row-major layout: row pitch in elements
  int64_t blk_pitch;      // tile-blocked layout: rows per 64-class block (B_tot rounded up to 128)
  int64_t n_cb;           // tile-blocked layout: number of 64-class blocks
  int n_blocks;
  size_t bytes;
};

// MSML_HEAD_DCOS_ROWMAJOR=1 keeps round 1's row-major dcos (A/B measurements); default: tile-blocked
static bool dcos_blocked() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MSML_HEAD_DCOS_ROWMAJOR"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}
// Which of the tensor-bound GEMMs run on CTA pairs (cta_group::2).  Measured at config-4 per-rank shapes (B_tot = 1024,
// 125,000 classes; profiles/r02_head_ab.md): dX 104.7 -> 98.4 us (long K loop, light epilogue), but fwd 109 -> 120 us and
// dcos 140 -> 155 us (8 k-blocks per tile: the pair's cross-SM accumulator hand-off is paid per tile), so the default is
// pairs for dX only.  MSML_HEAD_PAIR=all | dx | 0 overrides (A/B measurements, tests).
static int head_pair_mode() {       // 0 none, 1 dX only, 2 all
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MSML_HEAD_PAIR");
    v = !e ? 1 : (e[0] == '0' ? 0 : (e[0] == 'a' || e[0] == '1' ? 2 : 1));
  }
  return v;
}
constexpr int kFwdBlockN = 128;   // smallest class tile; each tile yields two softmax partials (one per epilogue half)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static HeadWs carve(void* ws, int64_t B_tot, int64_t n_s) {
  HeadWs h;
  h.n_blocks = 2 * (int)((n_s + kFwdBlockN - 1) / kFwdBlockN);
  h.ld_dc = (n_s + 7) / 8 * 8;
  h.blk_pitch = (B_tot + kBlockM - 1) / kBlockM * kBlockM;
  h.n_cb = (n_s + 63) / 64;
  size_t off = 0;
  char* base = static_cast<char*>(ws);
  h.part_max = reinterpret_cast<float*>(base + off); off = align_up(off + sizeof(float) * h.n_blocks * B_tot, 256);
  h.part_sum = reinterpret_cast<float*>(base + off); off = align_up(off + sizeof(float) * h.n_blocks * B_tot, 256);
  h.tgt = reinterpret_cast<float*>(base + off); off = align_up(off + sizeof(float) * B_tot, 256);
  h.rdot = reinterpret_cast<float*>(base + off); off = align_up(off + sizeof(float) * n_s, 256);
  {   // either layout fits (the blocked one pads rows to 128 and classes to 64)
    const size_t rowmajor = 2 * (size_t)B_tot * h.ld_dc, blocked = 2 * (size_t)h.n_cb * h.blk_pitch * 64;
    h.dcos = reinterpret_cast<__nv_bfloat16*>(base + off); off = align_up(off + (rowmajor > blocked ? rowmajor : blocked), 256);
  }
  h.bytes = off;
  return h;
}

}  // namespace tc
}  // namespace msml

using namespace msml;
using namespace msml::tc;

extern "C" int msml_gemm_bf16_tn(const void* a, int64_t lda, const void* b, int64_t ldb, float* c, int64_t ldc,
                                 int64_t M, int64_t N, int64_t K, void* stream) {
  MSML_REQUIRE(a && b && c && M > 0 && N > 0 && K > 0, MSML_EINVAL, "bad GEMM arguments");
  CUtensorMap ma, mb;
  if (int e = encode_tmap_bf16_kmajor(&ma, a, M, K, lda, kBlockM)) return e;
  if (int e = encode_tmap_bf16_kmajor(&mb, b, N, K, ldb, 256)) return e;
  EpiStore epi;
  epi.c = c; epi.ldc = ldc; epi.M = (int)M; epi.N = (int)N; epi.block_n = 256;
  return launch_gemm<256, 2, 4, false, false>("gemm_bf16_tn", ma, mb, make_shape(M, N, K, 256), epi, (cudaStream_t)stream);
}

// Generic layouts for the tests of the MN-major operand paths:  C (M x N) = op(A) * op(B)^T where
// a_mn != 0 means A is stored (K x M) row-major, b_mn != 0 means B is stored (K x N) row-major.
extern "C" int msml_gemm_bf16(const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb, int b_mn, float* c,
                              int64_t ldc, int64_t M, int64_t N, int64_t K, int block_n, void* stream) {
  MSML_REQUIRE(a && b && c && M > 0 && N > 0 && K > 0, MSML_EINVAL, "bad GEMM arguments");
  MSML_REQUIRE(block_n == 256 || block_n == 512, MSML_EINVAL, "block_n must be 256 or 512");
  CUtensorMap ma, mb;
  if (int e = a_mn ? encode_tmap_bf16_mnmajor(&ma, a, M, K, lda) : encode_tmap_bf16_kmajor(&ma, a, M, K, lda, kBlockM)) return e;
  if (int e = b_mn ? encode_tmap_bf16_mnmajor(&mb, b, N, K, ldb) : encode_tmap_bf16_kmajor(&mb, b, N, K, ldb, 256)) return e;
  EpiStore epi;
  epi.c = c; epi.ldc = ldc; epi.M = (int)M; epi.N = (int)N; epi.block_n = block_n;
  cudaStream_t st = (cudaStream_t)stream;
  const GemmShape sh = make_shape(M, N, K, block_n);
  const int sel = (block_n == 512 ? 4 : 0) + (a_mn ? 2 : 0) + (b_mn ? 1 : 0);
  switch (sel) {
    case 0: return launch_gemm<256, 2, 4, false, false>("gemm_bf16", ma, mb, sh, epi, st);
    case 1: return launch_gemm<256, 2, 4, false, true>("gemm_bf16", ma, mb, sh, epi, st);
    case 2: return launch_gemm<256, 2, 4, true, false>("gemm_bf16", ma, mb, sh, epi, st);
    case 3: return launch_gemm<256, 2, 4, true, true>("gemm_bf16", ma, mb, sh, epi, st);
    case 4: return launch_gemm<512, 1, 2, false, false>("gemm_bf16", ma, mb, sh, epi, st);
    case 5: return launch_gemm<512, 1, 2, false, true>("gemm_bf16", ma, mb, sh, epi, st);
    case 6: return launch_gemm<512, 1, 2, true, false>("gemm_bf16", ma, mb, sh, epi, st);
    default: return launch_gemm<512, 1, 2, true, true>("gemm_bf16", ma, mb, sh, epi, st);
  }
}

// The CTA-pair kernel with the plain fp32-store epilogue (bring-up / tests): same contract as msml_gemm_bf16, block_n in
// {128, 256}; both operand majors.
extern "C" int msml_gemm_bf16_pair(const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb, int b_mn, float* c,
                                   int64_t ldc, int64_t M, int64_t N, int64_t K, int block_n, void* stream) {
  MSML_REQUIRE(a && b && c && M > 0 && N > 0 && K > 0, MSML_EINVAL, "bad GEMM arguments");
  MSML_REQUIRE(block_n == 128 || block_n == 256, MSML_EINVAL, "block_n must be 128 or 256");
  CUtensorMap ma, mb;
  if (int e = a_mn ? encode_tmap_bf16_mnmajor(&ma, a, M, K, lda) : encode_tmap_bf16_kmajor(&ma, a, M, K, lda, kBlockM)) return e;
  if (int e = b_mn ? encode_tmap_bf16_mnmajor(&mb, b, N, K, ldb) : encode_tmap_bf16_kmajor(&mb, b, N, K, ldb, block_n / 2)) return e;
  EpiStore epi;
  epi.c = c; epi.ldc = ldc; epi.M = (int)M; epi.N = (int)N; epi.block_n = block_n;
  cudaStream_t st = (cudaStream_t)stream;
  const GemmShape sh = make_shape(M, N, K, block_n);
  const int sel = (block_n == 128 ? 4 : 0) + (a_mn ? 2 : 0) + (b_mn ? 1 : 0);
  switch (sel) {
    case 0: return launch_gemm_pair<256, 2, 6, false, false>("gemm_bf16_pair", ma, mb, sh, epi, st);
    case 1: return launch_gemm_pair<256, 2, 6, false, true>("gemm_bf16_pair", ma, mb, sh, epi, st);
    case 2: return launch_gemm_pair<256, 2, 6, true, false>("gemm_bf16_pair", ma, mb, sh, epi, st);
    case 3: return launch_gemm_pair<256, 2, 6, true, true>("gemm_bf16_pair", ma, mb, sh, epi, st);
    case 4: return launch_gemm_pair<128, 4, 8, false, false>("gemm_bf16_pair", ma, mb, sh, epi, st);
    case 5: return launch_gemm_pair<128, 4, 8, false, true>("gemm_bf16_pair", ma, mb, sh, epi, st);
    case 6: return launch_gemm_pair<128, 4, 8, true, false>("gemm_bf16_pair", ma, mb, sh, epi, st);
    default: return launch_gemm_pair<128, 4, 8, true, true>("gemm_bf16_pair", ma, mb, sh, epi, st);
  }
}

extern "C" int msml_wnorm_cast(const float* w, void* wn, void* wn_t, int64_t ld_t, float* inv_norm, int64_t n, int64_t D,
                               void* stream) {
  MSML_REQUIRE(w && wn && n > 0 && D > 0 && D % 8 == 0, MSML_EINVAL, "bad wnorm arguments (D %% 8 must be 0)");
  MSML_REQUIRE(aligned16(w) && aligned16(wn), MSML_EALIGN, "wnorm buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  {
    MSML_PROF("wnorm_cast", (double)n * D * 6, st);
    wnorm_cast_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(wn), inv_norm, n, (int)D, true);
  }
  MSML_LAUNCH_CHECK();
  if (wn_t) {
    MSML_REQUIRE(ld_t >= n, MSML_EINVAL, "ld_t %lld < n %lld", (long long)ld_t, (long long)n);
    dim3 grid((unsigned)((n + 63) / 64), (unsigned)((D + 63) / 64));
    MSML_PROF("transpose_bf16", (double)n * D * 4, st);
    transpose_bf16_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(wn), static_cast<__nv_bfloat16*>(wn_t), n, (int)D, ld_t);
    MSML_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int msml_cast_bf16(const float* x, void* x_bf16, void* x_t, int64_t ld_t, int64_t rows, int64_t D, void* stream) {
  MSML_REQUIRE(x && x_bf16 && rows > 0 && D > 0 && D % 8 == 0, MSML_EINVAL, "bad cast arguments (D %% 8 must be 0)");
  MSML_REQUIRE(aligned16(x) && aligned16(x_bf16), MSML_EALIGN, "cast buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  wnorm_cast_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, static_cast<__nv_bfloat16*>(x_bf16), nullptr, rows, (int)D, false);
  MSML_LAUNCH_CHECK();
  if (x_t) {
    MSML_REQUIRE(ld_t >= rows, MSML_EINVAL, "ld_t %lld < rows %lld", (long long)ld_t, (long long)rows);
    dim3 grid((unsigned)((rows + 63) / 64), (unsigned)((D + 63) / 64));
    transpose_bf16_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x_bf16), static_cast<__nv_bfloat16*>(x_t), rows, (int)D, ld_t);
    MSML_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int msml_transpose_bf16(const void* src, void* dst, int64_t rows, int64_t cols, int64_t ld_t, void* stream) {
  MSML_REQUIRE(src && dst && rows > 0 && cols > 0 && ld_t >= rows, MSML_EINVAL, "bad transpose arguments");
  dim3 grid((unsigned)((rows + 63) / 64), (unsigned)((cols + 63) / 64));
  transpose_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const __nv_bfloat16*>(src),
                                                                static_cast<__nv_bfloat16*>(dst), rows, (int)cols, ld_t);
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t msml_head_workspace(int64_t B_tot, int64_t n_s, int64_t D) {
  (void)D;
  if (B_tot <= 0 || n_s <= 0) return 0;
  return carve(nullptr, B_tot, n_s).bytes;
}

// class-tile width for the logits GEMMs: the one that fills the last wave of the persistent grid best
static int pick_block_n(int64_t B_tot, int64_t n_s, bool pair = false) {
  int64_t mb = (B_tot + kBlockM - 1) / kBlockM, sms = num_sms();
  if (pair) { mb = (mb + 1) / 2; sms /= 2; }              // work units are 256-row tiles, workers are CTA pairs
  auto eff = [&](int bn) {
    const int64_t tiles = mb * ((n_s + bn - 1) / bn);
    return (double)tiles / (double)(((tiles + sms - 1) / sms) * sms);
  };
  return eff(256) >= eff(128) - 0.02 ? 256 : 128;
}

static int head_check(int64_t B_tot, int64_t n_s, int64_t D, void* ws, size_t ws_bytes) {
  MSML_REQUIRE(B_tot > 0 && n_s > 1 && D > 0, MSML_EINVAL, "bad head shape B_tot=%lld n_s=%lld D=%lld", (long long)B_tot,
               (long long)n_s, (long long)D);
  MSML_REQUIRE(D % 64 == 0 && D <= 512, MSML_EUNSUPPORTED, "embedding size D=%lld must be a multiple of 64 and <= 512", (long long)D);
  MSML_REQUIRE(ws && aligned16(ws), MSML_EALIGN, "workspace missing or misaligned");
  MSML_REQUIRE(ws_bytes >= msml_head_workspace(B_tot, n_s, D), MSML_EWORKSPACE, "head workspace too small: %zu < %zu", ws_bytes,
               msml_head_workspace(B_tot, n_s, D));
  return 0;
}

extern "C" int msml_head_fwd(const void* x, const void* wn, const int64_t* tl, int64_t B_tot, int64_t n_s, int64_t D,
                             const msml_margin_params* margin, float* stats, void* ws, size_t ws_bytes, void* stream) {
  if (int e = head_check(B_tot, n_s, D, ws, ws_bytes)) return e;
  MSML_REQUIRE(x && wn && tl && stats, MSML_EINVAL, "null pointer");
  Margin mg;
  if (int e = to_margin(margin, &mg)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  HeadWs h = carve(ws, B_tot, n_s);
  const bool pair = head_pair_mode() == 2 && B_tot > kBlockM;      // a pair needs two 128-row blocks to be worth it
  const int bn = pick_block_n(B_tot, n_s, pair);
  CUtensorMap ma, mb;
  if (int e = encode_tmap_bf16_kmajor(&ma, x, B_tot, D, D, kBlockM)) return e;
  if (int e = encode_tmap_bf16_kmajor(&mb, wn, n_s, D, D, pair ? bn / 2 : bn)) return e;     // box rows: the B rows ONE CTA loads
  EpiFwdStats epi;
  epi.tl = tl; epi.mg = mg; epi.B_tot = (int)B_tot; epi.n_s = (int)n_s; epi.block_n = bn;
  epi.part_max = h.part_max; epi.part_sum = h.part_sum; epi.tgt = h.tgt;
  const double min_bytes = 2.0 * (double)n_s * D + 2.0 * (double)B_tot * D;      // stream Wn + X once (bf16)
  if (pair) {
    if (bn == 256) {
      if (int e = launch_gemm_pair<256, 2, 6, false, false, 8>("head_fwd_gemm", ma, mb, make_shape(B_tot, n_s, D, 256), epi, st, min_bytes)) return e;
    } else {
      if (int e = launch_gemm_pair<128, 4, 8, false, false, 8>("head_fwd_gemm", ma, mb, make_shape(B_tot, n_s, D, 128), epi, st, min_bytes)) return e;
    }
  } else if (bn == 256) {
    if (int e = launch_gemm<256, 2, 4, false, false, 8>("head_fwd_gemm", ma, mb, make_shape(B_tot, n_s, D, 256), epi, st, min_bytes)) return e;
  } else {
    if (int e = launch_gemm<128, 4, 6, false, false, 8>("head_fwd_gemm", ma, mb, make_shape(B_tot, n_s, D, 128), epi, st, min_bytes)) return e;
  }
  const int n_blocks = (int)((n_s + bn - 1) / bn) * 2;      // two epilogue halves per class tile
  head_local_stats_kernel<<<(unsigned)((B_tot + 7) / 8), 256, 0, st>>>(h.part_max, h.part_sum, h.tgt, tl, n_blocks, (int)B_tot, stats);
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" int msml_head_merge_stats(const float* gathered, int64_t W, int64_t B_tot, float* gstats, float* loss, void* stream) {
  MSML_REQUIRE(gathered && gstats && loss && W > 0 && B_tot > 0, MSML_EINVAL, "bad merge arguments");
  head_merge_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(gathered, (int)W, (int)B_tot, gstats, loss);
  MSML_LAUNCH_CHECK();
  return 0;
}

static int head_bwd_impl(const void* x, const void* wn, const float* inv_norm, const int64_t* tl, int64_t B_tot,
                         int64_t n_s, int64_t D, const msml_margin_params* margin, const float* gstats, float* dx_full,
                         float* dw, void* ws, size_t ws_bytes, void* stream, bool raw) {
  if (int e = head_check(B_tot, n_s, D, ws, ws_bytes)) return e;
  MSML_REQUIRE(x && wn && (raw || inv_norm) && tl && gstats && dx_full && dw, MSML_EINVAL, "null pointer");
  MSML_REQUIRE(aligned16(dx_full) && aligned16(dw), MSML_EALIGN, "gradient buffers must be 16-byte aligned");
  Margin mg;
  if (int e = to_margin(margin, &mg)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  HeadWs h = carve(ws, B_tot, n_s);

  const bool blocked = dcos_blocked();
  const bool pair = head_pair_mode() == 2 && B_tot > kBlockM;
  const bool pair_dx = head_pair_mode() >= 1 && B_tot > kBlockM;
  // 1. recompute logits -> dcos (bf16; tile-blocked, see tc_gemm.cuh) [+ rdot[n] = <Wn[n], dWn[n]> unless raw]
  {
    if (!raw) MSML_CUDA(cudaMemsetAsync(h.rdot, 0, sizeof(float) * (size_t)n_s, st));
    const int bn = pick_block_n(B_tot, n_s, pair);
    CUtensorMap ma, mb;
    if (int e = encode_tmap_bf16_kmajor(&ma, x, B_tot, D, D, kBlockM)) return e;
    if (int e = encode_tmap_bf16_kmajor(&mb, wn, n_s, D, D, pair ? bn / 2 : bn)) return e;   // box rows: the B rows ONE CTA loads
    const float eps = 0.1f;   // ref partial_fc.py:154
    const double min_bytes = 2.0 * (double)n_s * D + 2.0 * (double)B_tot * D + 2.0 * (double)B_tot * n_s;   // + dcos out
    auto run_dcos = [&](auto epi) -> int {
      if (blocked) {
        if (int e = encode_tmap_2d(&epi.map_dcos, h.dcos, 2, 64, h.n_cb * h.blk_pitch, 64, 64, 32)) return e;
      } else {
        if (int e = encode_tmap_2d(&epi.map_dcos, h.dcos, 2, n_s, B_tot, h.ld_dc, 64, 32)) return e;
      }
      epi.blk_pitch = blocked ? (int)h.blk_pitch : 0;
      epi.tl = tl; epi.mg = mg; epi.B_tot = (int)B_tot; epi.n_s = (int)n_s; epi.block_n = bn;
      epi.gmax = gstats; epi.gsum = gstats + B_tot; epi.rdot = raw ? nullptr : h.rdot;
      epi.smooth_on = 1.0f - eps; epi.smooth_off = eps / (float)(n_s - 1); epi.inv_btot = 1.0f / (float)B_tot;
      if (pair) {
        if (bn == 256) return launch_gemm_pair<256, 2, 4, false, false, 8>("head_bwd_dcos_gemm", ma, mb, make_shape(B_tot, n_s, D, 256), epi, st, min_bytes);
        return launch_gemm_pair<128, 4, 6, false, false, 8>("head_bwd_dcos_gemm", ma, mb, make_shape(B_tot, n_s, D, 128), epi, st, min_bytes);
      }
      if (bn == 256) return launch_gemm<256, 2, 3, false, false, 8>("head_bwd_dcos_gemm", ma, mb, make_shape(B_tot, n_s, D, 256), epi, st, min_bytes);
      return launch_gemm<128, 4, 4, false, false, 8>("head_bwd_dcos_gemm", ma, mb, make_shape(B_tot, n_s, D, 128), epi, st, min_bytes);
    };
    if (int e = raw ? run_dcos(EpiBwdDcosT<false>()) : run_dcos(EpiBwdDcosT<true>())) return e;
  }
  // 2. dX_full = dcos (B_tot x n_s) * Wn (n_s x D): A = dcos (K-major), B = Wn read in place (MN-major), split-K
  {
    MSML_CUDA(cudaMemsetAsync(dx_full, 0, sizeof(float) * (size_t)B_tot * D, st));
    CUtensorMap ma, mb;
    if (blocked) {
      if (int e = encode_tmap_2d(&ma, h.dcos, 2, 64, h.n_cb * h.blk_pitch, 64, 64, kBlockM)) return e;      // 128 rows x 64 classes: 16 KB contiguous
    } else {
      if (int e = encode_tmap_bf16_kmajor(&ma, h.dcos, B_tot, n_s, h.ld_dc, kBlockM)) return e;
    }
    if (int e = encode_tmap_bf16_mnmajor(&mb, wn, D, n_s, D)) return e;
    const int m_units = pair_dx ? (int)((B_tot + 2 * kBlockM - 1) / (2 * kBlockM)) : (int)((B_tot + kBlockM - 1) / kBlockM);
    const int tiles = m_units * (int)((D + 255) / 256);
    // as many split-K units as fit in ONE wave of the persistent grid (rounding up would leave a second, almost
    // empty wave: 16 tiles x 10 splits = 160 units on 148 SMs ran at 54 % occupancy)
    int splits = (pair_dx ? num_sms() / 2 : num_sms()) / tiles;
    if (splits < 1) splits = 1;
    EpiDxAccum epi;
    epi.dx = dx_full; epi.B_tot = (int)B_tot; epi.D = (int)D; epi.block_n = 256;
    const double min_bytes = 2.0 * (double)n_s * D + 2.0 * (double)B_tot * n_s + 4.0 * (double)B_tot * D;
    GemmShape sh = make_shape(B_tot, D, n_s, 256, splits);
    sh.a_blk_pitch = (int)h.blk_pitch;
    if (pair_dx) {
      if (blocked) { if (int e = launch_gemm_pair<256, 2, 6, false, true, 8, true>("head_bwd_dx_gemm", ma, mb, sh, epi, st, min_bytes)) return e; }
      else { if (int e = launch_gemm_pair<256, 2, 6, false, true, 8, false>("head_bwd_dx_gemm", ma, mb, sh, epi, st, min_bytes)) return e; }
    } else {
      if (blocked) { if (int e = launch_gemm<256, 2, 4, false, true, 8, true>("head_bwd_dx_gemm", ma, mb, sh, epi, st, min_bytes)) return e; }
      else { if (int e = launch_gemm<256, 2, 4, false, true, 8, false>("head_bwd_dx_gemm", ma, mb, sh, epi, st, min_bytes)) return e; }
    }
  }
  // 3. dW = normalize_bwd(dcos^T X): A = dcos read in place (MN-major over classes), B = X in place (MN-major), K = B_tot
  {
    CUtensorMap ma, mb;
    if (blocked) {
      if (int e = encode_tmap_2d(&ma, h.dcos, 2, 64, h.n_cb * h.blk_pitch, 64, 64, 64)) return e;           // 64 classes x 64 rows: 8 KB contiguous
    } else {
      if (int e = encode_tmap_bf16_mnmajor(&ma, h.dcos, n_s, B_tot, h.ld_dc)) return e;
    }
    if (int e = encode_tmap_bf16_mnmajor(&mb, x, D, B_tot, D)) return e;
    // Two shapes of this kernel.  Short K (B_tot <= 128: the 1-GPU training step, 2 k-blocks per tile): the epilogue is
    // everything, so eight epilogue warps (two per TMEM lane quarter) and a 2-stage operand ring (93,431 classes: 65.9 ->
    // 53.9 us).  Long K (B_tot = 1024: 16 k-blocks per tile): four warps and three stages — the 2-stage ring starves the
    // MMAs there (188.7 vs 145.6 us at config-4 shapes; gpurun_out/r02l_head_*.json).  MSML_HEAD_DW_EW=4|8 forces one.
    static int dw_ew = -1;
    if (dw_ew < 0) { const char* e = getenv("MSML_HEAD_DW_EW"); dw_ew = e ? atoi(e) : 0; }
    const double min_bytes_dw = 4.0 * (double)n_s * D + 2.0 * (double)n_s * D + 2.0 * (double)B_tot * n_s + 2.0 * (double)B_tot * D;
    const bool wide_epilogue = dw_ew == 8 || (dw_ew != 4 && B_tot <= 2 * kBlockK);
    // CTA pairs for dW (each CTA stages half of the X tile, 256 classes x 256 columns per pair, a 5-deep ring): the X slice
    // is re-read from L2 for every tile, and halving that stream is worth 144.6 -> 132.0 us (exact) / 142.5 -> 127.5 us (raw)
    // at config-4 shapes; with few tiles (config-3 W=8 shard: 184) the pair's hand-off costs 2 us instead.
    // MSML_HEAD_DW_PAIR=0|1 forces it.
    static int dw_pair_env = -1;
    if (dw_pair_env < 0) { const char* e = getenv("MSML_HEAD_DW_PAIR"); dw_pair_env = !e ? 2 : (e[0] == '1' ? 1 : 0); }
    const int64_t dw_tiles = ((n_s + kBlockM - 1) / kBlockM) * ((D + 255) / 256);
    const bool dw_pair = n_s > kBlockM && !wide_epilogue && (dw_pair_env == 1 || (dw_pair_env == 2 && dw_tiles >= 4 * (int64_t)num_sms()));
    if (raw) {     // dWn as it leaves the tensor core; the fused optimizer projects and scales (msml_pfc_sgd_update_raw)
      EpiDwRaw epir;
      if (int e = encode_tmap_2d(&epir.map_dw, dw, 4, D, n_s, D, 32, 32)) return e;
      epir.n_s = (int)n_s; epir.D = (int)D; epir.block_n = 256;
      GemmShape shr = make_shape(n_s, D, B_tot, 256, 1, /*n_fastest=*/true);
      shr.a_blk_pitch = (int)h.blk_pitch;
      const double mb_raw = 4.0 * (double)n_s * D + 2.0 * (double)B_tot * n_s + 2.0 * (double)B_tot * D;
      if (blocked && dw_pair) return launch_gemm_pair<256, 2, 5, true, true, 8, true>("head_bwd_dw_gemm", ma, mb, shr, epir, st, mb_raw);
      if (blocked) return launch_gemm<256, 2, 3, true, true, 8, true>("head_bwd_dw_gemm", ma, mb, shr, epir, st, mb_raw);
      return launch_gemm<256, 2, 3, true, true, 8, false>("head_bwd_dw_gemm", ma, mb, shr, epir, st, mb_raw);
    }
    if (blocked && wide_epilogue) {
      EpiDwNormBwdT<8> epi8;
      if (int e = encode_tmap_2d(&epi8.map_dw, dw, 4, D, n_s, D, 32, 32)) return e;
      epi8.wn = static_cast<const __nv_bfloat16*>(wn); epi8.inv_norm = inv_norm; epi8.rdot = h.rdot; epi8.n_s = (int)n_s; epi8.D = (int)D;
      epi8.block_n = 256;
      GemmShape sh8 = make_shape(n_s, D, B_tot, 256, 1, /*n_fastest=*/true);
      sh8.a_blk_pitch = (int)h.blk_pitch;
      if (int e = launch_gemm<256, 2, 2, true, true, 8, true>("head_bwd_dw_gemm", ma, mb, sh8, epi8, st, min_bytes_dw)) return e;
      return 0;
    }
    EpiDwNormBwdT<4> epi;
    if (int e = encode_tmap_2d(&epi.map_dw, dw, 4, D, n_s, D, 32, 32)) return e;
    epi.wn = static_cast<const __nv_bfloat16*>(wn); epi.inv_norm = inv_norm; epi.rdot = h.rdot; epi.n_s = (int)n_s; epi.D = (int)D;
    const double min_bytes = 4.0 * (double)n_s * D + 2.0 * (double)n_s * D + 2.0 * (double)B_tot * n_s + 2.0 * (double)B_tot * D;
    // (128-column tiles with a 5-deep ring, to keep more dcos loads in flight, measured slower: 161.6 vs 144.7 us at
    //  config-4 shapes, gpurun_out/r02k_head_125k_dw{128,256}.json)
    epi.block_n = 256;
    GemmShape sh = make_shape(n_s, D, B_tot, 256, 1, /*n_fastest=*/true);
    sh.a_blk_pitch = (int)h.blk_pitch;
    if (blocked && dw_pair) { if (int e = launch_gemm_pair<256, 2, 5, true, true, 4, true>("head_bwd_dw_gemm", ma, mb, sh, epi, st, min_bytes)) return e; }
    else if (blocked) { if (int e = launch_gemm<256, 2, 3, true, true, 4, true>("head_bwd_dw_gemm", ma, mb, sh, epi, st, min_bytes)) return e; }
    else { if (int e = launch_gemm<256, 2, 3, true, true, 4, false>("head_bwd_dw_gemm", ma, mb, sh, epi, st, min_bytes)) return e; }
  }
  return 0;
}

extern "C" int msml_head_bwd(const void* x, const void* wn, const float* inv_norm, const int64_t* tl, int64_t B_tot,
                             int64_t n_s, int64_t D, const msml_margin_params* margin, const float* gstats, float* dx_full,
                             float* dw, void* ws, size_t ws_bytes, void* stream) {
  return head_bwd_impl(x, wn, inv_norm, tl, B_tot, n_s, D, margin, gstats, dx_full, dw, ws, ws_bytes, stream, false);
}

extern "C" int msml_head_bwd_raw(const void* x, const void* wn, const int64_t* tl, int64_t B_tot, int64_t n_s, int64_t D,
                                 const msml_margin_params* margin, const float* gstats, float* dx_full, float* dwn,
                                 void* ws, size_t ws_bytes, void* stream) {
  return head_bwd_impl(x, wn, nullptr, tl, B_tot, n_s, D, margin, gstats, dx_full, dwn, ws, ws_bytes, stream, true);
}

extern "C" int msml_margin_fwd(float* cosm, const int64_t* label, int64_t B, int64_t C, int64_t ld,
                               const msml_margin_params* margin, void* stream) {
  MSML_REQUIRE(cosm && label && B > 0 && C > 0 && ld >= C, MSML_EINVAL, "bad margin arguments");
  Margin mg;
  if (int e = to_margin(margin, &mg)) return e;
  margin_fwd_kernel<<<(unsigned)((B * C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cosm, label, B, C, ld, mg);
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" int msml_margin_bwd(float* dl, const float* cosm, const int64_t* label, int64_t B, int64_t C, int64_t ld,
                               const msml_margin_params* margin, void* stream) {
  MSML_REQUIRE(dl && cosm && label && B > 0 && C > 0 && ld >= C, MSML_EINVAL, "bad margin arguments");
  Margin mg;
  if (int e = to_margin(margin, &mg)) return e;
  margin_bwd_kernel<<<(unsigned)((B * C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dl, cosm, label, B, C, ld, mg);
  MSML_LAUNCH_CHECK();
  return 0;
}
