// Head (PartialFC / in-model margin heads): margin math and the small non-GEMM kernels — unit-norm bf16 cast, bf16 transpose,
// per-rank and cross-rank softmax-statistics merges (the loss), margin forward / backward on a materialised cosine matrix.
//   ref headers/partial_fc.py:115,135-144,159-163; headers/margin_losses.py:275-303,390-418
// Kept apart from the tcgen05 code (head.cu, tc_gemm.cuh) so that tests/emu can compile them for the host and run them
// under the CPU emulation; head.cu includes this header where the definitions used to stand.
#pragma once
#include <cmath>

#include "common.cuh"

namespace msml {
namespace tc {

// ------------------------------------------------------------------------------- margin math
struct Margin {
  int kind;
  float s, m, a, k;
};
constexpr float kLog2e = 1.4426950408889634f;

// The reference takes acos of an fp32 cosine of unit vectors (|c| <= 1 up to 1e-7, no clamp).  Here the
// operands are rounded to bf16, which can push a well-aligned target to |c| ~ 1 + 4e-3; clamping to the open
// interval restores the reference's domain (and keeps d/dc finite) instead of manufacturing NaNs.
__device__ __forceinline__ float clamp_cos(float c) { return fminf(fmaxf(c, -1.0f + 1e-6f), 1.0f - 1e-6f); }

// logit / s at the target column (ref margin_losses.py:411-417 arc, :298-299 cos)
__device__ __forceinline__ float margin_target(const Margin& mg, float c) {
  c = clamp_cos(c);
  const float theta = acosf(c);
  const float m_eff = mg.m - mg.k * (theta - mg.a);
  return mg.kind == MSML_MARGIN_ARC ? cosf(theta + m_eff) : c - m_eff;
}
// d(logit)/d(cos) / s at the target column (the adaptive term carries gradient)
__device__ __forceinline__ float margin_target_grad(const Margin& mg, float c) {
  c = clamp_cos(c);
  const float theta = acosf(c);
  const float sin_t = sinf(theta);
  if (mg.kind == MSML_MARGIN_ARC) return (1.0f - mg.k) * sinf((1.0f - mg.k) * theta + mg.m + mg.k * mg.a) / sin_t;
  return 1.0f - mg.k / sin_t;
}

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

// merge per-tile partials of one rank into (max, sum, target logit) in natural units; one warp per row
// (config 3 at W = 1 has 1460 partials per row: a thread per row took 0.26 ms)
__global__ void __launch_bounds__(256)
head_local_stats_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum, const float* __restrict__ tgt,
                        const int64_t* __restrict__ tl, int n_blocks, int B_tot, float* __restrict__ stats) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B_tot) return;
  float m = -INFINITY;
  for (int b = lane; b < n_blocks; b += 32) m = fmaxf(m, part_max[(int64_t)b * B_tot + row]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int b = lane; b < n_blocks; b += 32) s += part_sum[(int64_t)b * B_tot + row] * exp2f(part_max[(int64_t)b * B_tot + row] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    stats[row] = m * (1.0f / kLog2e);                 // row max of this shard's logits
    stats[B_tot + row] = s;                           // sum exp(logit - max)
    stats[2 * B_tot + row] = tl[row] >= 0 ? tgt[row] : -INFINITY;
  }
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

}  // namespace tc
}  // namespace msml
