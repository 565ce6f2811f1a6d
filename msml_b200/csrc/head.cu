// K-E..K-H  PartialFC head on tcgen05 tensor cores.      ref headers/partial_fc.py:96-99,115,118-177
//                                                        ref headers/margin_losses.py:275-303,390-418
// Per rank (class shard):  X (B_tot, D) bf16 gathered features, Wn (n_s, D) bf16 unit class centres.
//
//   head_fwd   S = X Wn^T (tcgen05, fp32 in TMEM); epilogue: margin on the target column, * s,
//              online row max / sum-exp per (row, class tile)  -> partials -> merged row stats.
//              The (B_tot x n_s) logits matrix is NEVER written to memory.
//   merge      combines the row stats of all ranks (after one all-gather), computes the loss.
//   head_bwd   1. recompute S tile by tile; epilogue: p = softmax, grad = (p - smoothed one-hot)/B_tot,
//                 chain through margin and scale -> dcos in bf16 (row-major + transposed)
//              2. dX  = dcos Wn        split-K over the class dimension, fp32 red.add
//              3. dW  = normalize_bwd(dcos^T X): 128 x 512 accumulator (all of TMEM), epilogue
//                 applies (dWn - Wn * <Wn, dWn>) / ||W|| and streams fp32 rows out.
// Everything runs on the caller's stream; no allocation, no host sync.
#include <cmath>

#include "tc_gemm.cuh"

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

int encode_tmap_bf16_kmajor(CUtensorMap* out, const void* base, int64_t rows, int64_t k, int64_t ld_elems, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  MSML_REQUIRE(fn != nullptr, MSML_EUNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  MSML_REQUIRE(aligned16(base), MSML_EALIGN, "TMA operand base must be 16-byte aligned");
  MSML_REQUIRE(ld_elems % 8 == 0 && ld_elems >= k, MSML_EALIGN, "TMA operand pitch %lld must be a multiple of 8 elements and >= k=%lld",
               (long long)ld_elems, (long long)k);
  MSML_REQUIRE(rows > 0 && k > 0 && box_rows > 0 && box_rows <= 256, MSML_EINVAL, "bad TMA operand shape");
  const cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld_elems * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MSML_REQUIRE(r == CUDA_SUCCESS, MSML_EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// ------------------------------------------------------------------------------- margin math
struct Margin {
  int kind;
  float s, m, a, k;
};
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float fast_exp2(float x) {   // MUFU.EX2; -inf -> 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// logit / s at the target column (ref margin_losses.py:411-417 arc, :298-299 cos)
__device__ __forceinline__ float margin_target(const Margin& mg, float c) {
  const float theta = acosf(c);   // no clamp, as the reference: |c| > 1 -> NaN
  const float m_eff = mg.m - mg.k * (theta - mg.a);
  return mg.kind == MSML_MARGIN_ARC ? cosf(theta + m_eff) : c - m_eff;
}
// d(logit)/d(cos) / s at the target column (the adaptive term carries gradient)
__device__ __forceinline__ float margin_target_grad(const Margin& mg, float c) {
  const float theta = acosf(c);
  const float sin_t = sinf(theta);
  if (mg.kind == MSML_MARGIN_ARC) return (1.0f - mg.k) * sinf((1.0f - mg.k) * theta + mg.m + mg.k * mg.a) / sin_t;
  return 1.0f - mg.k / sin_t;
}

// ------------------------------------------------------------------------------- epilogues
// plain store (tests / in-model heads): C[row, col] = acc
struct EpiStore {
  float* c;
  int64_t ldc;
  int M, N, block_n;
  __device__ void operator()(uint32_t tmem_acc, int m_blk, int n_blk, int, int quarter, int lane) const {
    const int row = m_blk * kBlockM + quarter * 32 + lane;
    const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
    for (int c0 = 0; c0 < block_n; c0 += 32) {
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
struct EpiFwdStats {
  const int64_t* tl;
  Margin mg;
  int B_tot, n_s, block_n;
  float* part_max;   // [n_blocks][B_tot]  (log2 domain: logit * log2e)
  float* part_sum;   // [n_blocks][B_tot]
  float* tgt;        // [B_tot] target logit (natural units), written by the tile that owns the column
  __device__ void operator()(uint32_t tmem_acc, int m_blk, int n_blk, int, int quarter, int lane) const {
    const int row = m_blk * kBlockM + quarter * 32 + lane;
    const bool row_ok = row < B_tot;
    const int64_t label = row_ok ? tl[row] : -1;
    const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
    const float s2 = mg.s * kLog2e;
    float run_max = -INFINITY, run_sum = 0.f;
    for (int c0 = 0; c0 < block_n; c0 += 32) {
      float v[32];
      tmem_ld32(taddr + c0, v);
      const int col0 = n_blk * block_n + c0;
      if (col0 >= n_s) break;   // warp-uniform
      const int64_t rel = label - col0;
      if (rel >= 0 && rel < 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j == (int)rel) { v[j] = margin_target(mg, v[j]); tgt[row] = v[j] * mg.s; }
      }
      float cmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = (col0 + j < n_s) ? v[j] * s2 : -INFINITY;
        cmax = fmaxf(cmax, v[j]);
      }
      const float new_max = fmaxf(run_max, cmax);   // finite: column col0 < n_s exists
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += fast_exp2(v[j] - new_max);
      run_sum = run_sum * fast_exp2(run_max - new_max) + acc;
      run_max = new_max;
    }
    if (row_ok) {
      part_max[(int64_t)n_blk * B_tot + row] = run_max;
      part_sum[(int64_t)n_blk * B_tot + row] = run_sum;
    }
  }
};

// backward pass 1: recompute logits, emit dcos (bf16) row-major and transposed
struct EpiBwdDcos {
  const int64_t* tl;
  Margin mg;
  int B_tot, n_s, block_n;
  const float* gmax;   // [B_tot] global row max   (natural units)
  const float* gsum;   // [B_tot] global row sum of exp(logit - max)
  __nv_bfloat16* dcos;     // (B_tot, ld_dc)
  int64_t ld_dc;
  __nv_bfloat16* dcos_t;   // (n_s, ld_t)
  int64_t ld_t;
  float smooth_on, smooth_off, inv_btot;
  __device__ void operator()(uint32_t tmem_acc, int m_blk, int n_blk, int, int quarter, int lane) const {
    const int row = m_blk * kBlockM + quarter * 32 + lane;
    const bool row_ok = row < B_tot;
    const int64_t label = row_ok ? tl[row] : -1;
    const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
    const float s2 = mg.s * kLog2e;
    // p = exp2(logit*log2e - off),  off = max*log2e + log2(sum)
    const float off = row_ok ? fmaf(gmax[row], kLog2e, log2f(gsum[row])) : 0.f;
    const float t_off = label >= 0 ? smooth_off : 0.f;   // rows without a local target: one_hot row absent
    const float gs = mg.s * inv_btot;
    for (int c0 = 0; c0 < block_n; c0 += 32) {
      float v[32];
      tmem_ld32(taddr + c0, v);
      const int col0 = n_blk * block_n + c0;
      if (col0 >= n_s) break;   // warp-uniform
      const int64_t rel = label - col0;
      float tgt_mult = 1.f;
      const bool has_t = rel >= 0 && rel < 32;
      if (has_t) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j == (int)rel) { tgt_mult = margin_target_grad(mg, v[j]); v[j] = margin_target(mg, v[j]); }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float p = fast_exp2(fmaf(v[j], s2, -off));
        const bool is_t = has_t && j == (int)rel;
        float g = (p - (is_t ? smooth_on : t_off)) * gs;
        if (is_t) g *= tgt_mult;
        v[j] = g;
      }
      if (row_ok) {
        __nv_bfloat16* drow = dcos + (int64_t)row * ld_dc + col0;
        if (col0 + 32 <= n_s) {
#pragma unroll
          for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(drow + q * 8) = Vec<__nv_bfloat16>::pack(v + q * 8);
        } else {
          for (int j = 0; j < 32 && col0 + j < n_s; ++j) drow[j] = __float2bfloat16_rn(v[j]);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j)   // lanes = consecutive rows: 64-byte coalesced segments
          if (col0 + j < n_s) dcos_t[(int64_t)(col0 + j) * ld_t + row] = __float2bfloat16_rn(v[j]);
      }
    }
  }
};

// backward pass 2: dX += partial (split-K over classes)
struct EpiDxAccum {
  float* dx;   // (B_tot, D) fp32, zero-initialised
  int B_tot, D, block_n;
  __device__ void operator()(uint32_t tmem_acc, int m_blk, int n_blk, int, int quarter, int lane) const {
    const int row = m_blk * kBlockM + quarter * 32 + lane;
    const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
    for (int c0 = 0; c0 < block_n; c0 += 32) {
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

// backward pass 3: dW = (dWn - Wn * <Wn, dWn>) * inv_norm      (ref :115 normalize backward)
struct EpiDwNormBwd {
  const __nv_bfloat16* wn;   // (n_s, D)
  const float* inv_norm;     // (n_s)
  float* dw;                 // (n_s, D) fp32
  int n_s, D;
  __device__ void operator()(uint32_t tmem_acc, int m_blk, int, int, int quarter, int lane) const {
    const int row = m_blk * kBlockM + quarter * 32 + lane;   // class index
    const bool ok = row < n_s;
    const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
    const __nv_bfloat16* wrow = wn + (int64_t)(ok ? row : 0) * D;
    float dot = 0.f;
    for (int c0 = 0; c0 < D; c0 += 32) {
      float v[32], w[32];
      tmem_ld32(taddr + c0, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) Vec<__nv_bfloat16>::unpack(*reinterpret_cast<const uint4*>(wrow + c0 + q * 8), w + q * 8);
#pragma unroll
      for (int j = 0; j < 32; ++j) dot = fmaf(v[j], w[j], dot);
    }
    const float inv = ok ? inv_norm[row] : 0.f;
    for (int c0 = 0; c0 < D; c0 += 32) {
      float v[32], w[32];
      tmem_ld32(taddr + c0, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) Vec<__nv_bfloat16>::unpack(*reinterpret_cast<const uint4*>(wrow + c0 + q * 8), w + q * 8);
      if (ok) {
        float* o = dw + (int64_t)row * D + c0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 r;
          r.x = (v[q * 4 + 0] - w[q * 4 + 0] * dot) * inv;
          r.y = (v[q * 4 + 1] - w[q * 4 + 1] * dot) * inv;
          r.z = (v[q * 4 + 2] - w[q * 4 + 2] * dot) * inv;
          r.w = (v[q * 4 + 3] - w[q * 4 + 3] * dot) * inv;
          __stcs(reinterpret_cast<float4*>(o + q * 4), r);
        }
      }
    }
  }
};

// ------------------------------------------------------------------------------- small kernels
// one warp per row: wn = bf16(w / max(||w||, eps)), inv_norm
__global__ void __launch_bounds__(256)
wnorm_cast_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wn, float* __restrict__ inv_norm,
                  int64_t n, int D, bool normalize) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const float4* src = reinterpret_cast<const float4*>(w + row * D);
  float ss = 0.f;
  for (int j = lane; j < D / 4; j += 32) {
    const float4 v = __ldg(src + j);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = normalize ? 1.0f / fmaxf(sqrtf(ss), 1e-12f) : 1.0f;
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
  for (int j = lane; j < D / 8; j += 32) {
    const float4 a = __ldg(src + 2 * j), b = __ldg(src + 2 * j + 1);
    const float f[8] = {a.x * inv, a.y * inv, a.z * inv, a.w * inv, b.x * inv, b.y * inv, b.z * inv, b.w * inv};
    reinterpret_cast<uint4*>(wn + row * D)[j] = Vec<__nv_bfloat16>::pack(f);
  }
}

// bf16 (rows x cols) -> (cols x ld_t) transpose through a padded 64x64 smem tile
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows, int cols, int64_t ld_t) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int64_t r0 = (int64_t)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int r = i >> 6, c = i & 63;
    tile[r][c] = (r0 + r < rows && c0 + c < cols) ? src[(r0 + r) * cols + c0 + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int c = i >> 6, r = i & 63;
    if (r0 + r < rows && c0 + c < cols) dst[(int64_t)(c0 + c) * ld_t + r0 + r] = tile[r][c];
  }
}

// merge per-tile partials of one rank into (max, sum, target logit) in natural units
__global__ void __launch_bounds__(128)
head_local_stats_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum, const float* __restrict__ tgt,
                        const int64_t* __restrict__ tl, int n_blocks, int B_tot, float* __restrict__ stats) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= B_tot) return;
  float m = -INFINITY;
  for (int b = 0; b < n_blocks; ++b) m = fmaxf(m, part_max[(int64_t)b * B_tot + row]);
  float s = 0.f;
  for (int b = 0; b < n_blocks; ++b) s += part_sum[(int64_t)b * B_tot + row] * exp2f(part_max[(int64_t)b * B_tot + row] - m);
  stats[row] = m * (1.0f / kLog2e);                 // row max of this shard's logits
  stats[B_tot + row] = s;                           // sum exp(logit - max)
  stats[2 * B_tot + row] = tl[row] >= 0 ? tgt[row] : -INFINITY;
}

// merge the stats of W ranks; loss = -mean(log(max(p_target, 1e-30)))     ref :136,141,162-163
__global__ void __launch_bounds__(256)
head_merge_kernel(const float* __restrict__ gathered, int W, int B_tot, float* __restrict__ gstats, float* __restrict__ loss) {
  __shared__ float red[256];
  float local = 0.f;
  for (int row = threadIdx.x; row < B_tot; row += 256) {
    float m = -INFINITY;
    for (int r = 0; r < W; ++r) m = fmaxf(m, gathered[((int64_t)r * 3 + 0) * B_tot + row]);
    float s = 0.f, pt = 0.f;
    for (int r = 0; r < W; ++r) {
      const float mr = gathered[((int64_t)r * 3 + 0) * B_tot + row];
      s += gathered[((int64_t)r * 3 + 1) * B_tot + row] * expf(mr - m);
    }
    for (int r = 0; r < W; ++r) {
      const float t = gathered[((int64_t)r * 3 + 2) * B_tot + row];
      if (t != -INFINITY) pt += expf(t - m) / s;      // sum over ranks == all_reduce(SUM) of :162
    }
    gstats[row] = m;
    gstats[B_tot + row] = s;
    local += logf(fmaxf(pt, 1e-30f));
  }
  red[threadIdx.x] = local;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = -red[0] / (float)B_tot;
}

// in-model margin heads on a materialised cosine matrix
__global__ void __launch_bounds__(256)
margin_fwd_kernel(float* __restrict__ cosm, const int64_t* __restrict__ label, int64_t B, int64_t C, int64_t ld, Margin mg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int64_t r = i / C, c = i - r * C;
  float v = cosm[r * ld + c];
  if (label[r] == c) v = margin_target(mg, v);
  cosm[r * ld + c] = v * mg.s;
}
__global__ void __launch_bounds__(256)
margin_bwd_kernel(float* __restrict__ dl, const float* __restrict__ cosm, const int64_t* __restrict__ label, int64_t B,
                  int64_t C, int64_t ld, Margin mg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int64_t r = i / C, c = i - r * C;
  float g = dl[r * ld + c] * mg.s;
  if (label[r] == c) g *= margin_target_grad(mg, cosm[r * ld + c]);
  dl[r * ld + c] = g;
}

static int to_margin(const msml_margin_params* p, Margin* out) {
  MSML_REQUIRE(p != nullptr, MSML_EINVAL, "margin params missing");
  MSML_REQUIRE(p->kind == MSML_MARGIN_ARC || p->kind == MSML_MARGIN_COS, MSML_EINVAL, "margin kind error (%d)", p->kind);
  out->kind = p->kind; out->s = p->s; out->m = p->m; out->a = p->a; out->k = p->k;
  return 0;
}

// workspace carving
struct HeadWs {
  float* part_max; float* part_sum; float* tgt;
  __nv_bfloat16* dcos; __nv_bfloat16* dcos_t;
  int64_t ld_dc, ld_t;
  int n_blocks;
  size_t bytes;
};
constexpr int kFwdBlockN = 256;

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static HeadWs carve(void* ws, int64_t B_tot, int64_t n_s) {
  HeadWs h;
  h.n_blocks = (int)((n_s + kFwdBlockN - 1) / kFwdBlockN);
  h.ld_dc = (n_s + 7) / 8 * 8;
  h.ld_t = (B_tot + 7) / 8 * 8;
  size_t off = 0;
  char* base = static_cast<char*>(ws);
  h.part_max = reinterpret_cast<float*>(base + off); off = align_up(off + sizeof(float) * h.n_blocks * B_tot, 256);
  h.part_sum = reinterpret_cast<float*>(base + off); off = align_up(off + sizeof(float) * h.n_blocks * B_tot, 256);
  h.tgt = reinterpret_cast<float*>(base + off); off = align_up(off + sizeof(float) * B_tot, 256);
  h.dcos = reinterpret_cast<__nv_bfloat16*>(base + off); off = align_up(off + 2 * (size_t)B_tot * h.ld_dc, 256);
  h.dcos_t = reinterpret_cast<__nv_bfloat16*>(base + off); off = align_up(off + 2 * (size_t)n_s * h.ld_t, 256);
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
  EpiStore epi{c, ldc, (int)M, (int)N, 256};
  return launch_gemm<256, 2, 4>("gemm_bf16_tn", ma, mb, make_shape(M, N, K, 256), epi, (cudaStream_t)stream);
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
  CUtensorMap ma, mb;
  if (int e = encode_tmap_bf16_kmajor(&ma, x, B_tot, D, D, kBlockM)) return e;
  if (int e = encode_tmap_bf16_kmajor(&mb, wn, n_s, D, D, 256)) return e;
  EpiFwdStats epi{tl, mg, (int)B_tot, (int)n_s, kFwdBlockN, h.part_max, h.part_sum, h.tgt};
  if (int e = launch_gemm<kFwdBlockN, 2, 4>("head_fwd_gemm", ma, mb, make_shape(B_tot, n_s, D, kFwdBlockN), epi, st)) return e;
  head_local_stats_kernel<<<(unsigned)((B_tot + 127) / 128), 128, 0, st>>>(h.part_max, h.part_sum, h.tgt, tl, h.n_blocks, (int)B_tot, stats);
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" int msml_head_merge_stats(const float* gathered, int64_t W, int64_t B_tot, float* gstats, float* loss, void* stream) {
  MSML_REQUIRE(gathered && gstats && loss && W > 0 && B_tot > 0, MSML_EINVAL, "bad merge arguments");
  head_merge_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(gathered, (int)W, (int)B_tot, gstats, loss);
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" int msml_head_bwd(const void* x, const void* x_t, int64_t ld_xt, const void* wn, const void* wn_t, int64_t ld_wt,
                             const float* inv_norm, const int64_t* tl, int64_t B_tot, int64_t n_s, int64_t D,
                             const msml_margin_params* margin, const float* gstats, float* dx_full, float* dw, void* ws,
                             size_t ws_bytes, void* stream) {
  if (int e = head_check(B_tot, n_s, D, ws, ws_bytes)) return e;
  MSML_REQUIRE(x && x_t && wn && wn_t && inv_norm && tl && gstats && dx_full && dw, MSML_EINVAL, "null pointer");
  MSML_REQUIRE(aligned16(dx_full) && aligned16(dw), MSML_EALIGN, "gradient buffers must be 16-byte aligned");
  Margin mg;
  if (int e = to_margin(margin, &mg)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  HeadWs h = carve(ws, B_tot, n_s);

  // 1. recompute logits -> dcos (bf16), row-major and transposed
  {
    CUtensorMap ma, mb;
    if (int e = encode_tmap_bf16_kmajor(&ma, x, B_tot, D, D, kBlockM)) return e;
    if (int e = encode_tmap_bf16_kmajor(&mb, wn, n_s, D, D, 256)) return e;
    const float eps = 0.1f;   // ref partial_fc.py:154
    EpiBwdDcos epi{tl, mg, (int)B_tot, (int)n_s, kFwdBlockN, gstats, gstats + B_tot, h.dcos, h.ld_dc, h.dcos_t, h.ld_t,
                   1.0f - eps, eps / (float)(n_s - 1), 1.0f / (float)B_tot};
    if (int e = launch_gemm<kFwdBlockN, 2, 4>("head_bwd_dcos_gemm", ma, mb, make_shape(B_tot, n_s, D, kFwdBlockN), epi, st)) return e;
  }
  // 2. dX_full = dcos (B_tot x n_s) * Wn (n_s x D): A = dcos, B = Wn^T (D x n_s), K = n_s, split-K
  {
    MSML_CUDA(cudaMemsetAsync(dx_full, 0, sizeof(float) * (size_t)B_tot * D, st));
    CUtensorMap ma, mb;
    if (int e = encode_tmap_bf16_kmajor(&ma, h.dcos, B_tot, n_s, h.ld_dc, kBlockM)) return e;
    if (int e = encode_tmap_bf16_kmajor(&mb, wn_t, D, n_s, ld_wt, 256)) return e;
    const int tiles = (int)((B_tot + kBlockM - 1) / kBlockM) * (int)((D + 255) / 256);
    int splits = (num_sms() + tiles - 1) / tiles;
    EpiDxAccum epi{dx_full, (int)B_tot, (int)D, 256};
    if (int e = launch_gemm<256, 2, 4>("head_bwd_dx_gemm", ma, mb, make_shape(B_tot, D, n_s, 256, splits), epi, st)) return e;
  }
  // 3. dW = normalize_bwd(dcos^T (n_s x B_tot) * X (B_tot x D)): A = dcos^T, B = X^T (D x B_tot), K = B_tot
  {
    CUtensorMap ma, mb;
    if (int e = encode_tmap_bf16_kmajor(&ma, h.dcos_t, n_s, B_tot, h.ld_t, kBlockM)) return e;
    if (int e = encode_tmap_bf16_kmajor(&mb, x_t, D, B_tot, ld_xt, 256)) return e;
    EpiDwNormBwd epi{static_cast<const __nv_bfloat16*>(wn), inv_norm, dw, (int)n_s, (int)D};
    if (int e = launch_gemm<512, 1, 2>("head_bwd_dw_gemm", ma, mb, make_shape(n_s, D, B_tot, 512), epi, st)) return e;
  }
  return 0;
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
