// Multi-tensor bf16 -> fp32 gradient accumulate: the kernel (launcher: optim.cu; tests/emu runs this header under the CPU emulation).
//   dst_f32[i] += float(src_bf16[i]) for up to MSML_ACCUM_MAX_SEGMENTS tensors in ONE launch (ref train.py:283-300 under autocast:
//   per-weight `grad.float()` + AccumulateGrad).
#pragma once
#include "common.cuh"

namespace msml {

constexpr int kAccThreads = 256;
constexpr int kAccVecPerThread = 4;                          // 4 x 8 bf16 per thread per block
constexpr int kAccElemsPerBlock = kAccThreads * kAccVecPerThread * 8;

struct AccumSegs {
  float* dst[MSML_ACCUM_MAX_SEGMENTS];
  const __nv_bfloat16* src[MSML_ACCUM_MAX_SEGMENTS];
  int64_t n[MSML_ACCUM_MAX_SEGMENTS];
  int block_end[MSML_ACCUM_MAX_SEGMENTS];                    // exclusive prefix end of the blocks of each segment
  int nseg;
};

__global__ void __launch_bounds__(kAccThreads)
accum_bf16_multi_kernel(const __grid_constant__ AccumSegs s) {
  // binary search of this block's segment (the table lives in the constant bank: dynamic indexing is free)
  int lo = 0, hi = s.nseg - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((int)blockIdx.x < s.block_end[mid]) hi = mid; else lo = mid + 1;
  }
  const int seg = lo;
  const int first = seg ? s.block_end[seg - 1] : 0;
  float* dst = s.dst[seg];
  const __nv_bfloat16* src = s.src[seg];
  const int64_t n = s.n[seg];
  const int64_t base = (int64_t)(blockIdx.x - first) * kAccElemsPerBlock;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
  if (vec_ok) {
#pragma unroll
    for (int u = 0; u < kAccVecPerThread; ++u) {
      const int64_t i = base + ((int64_t)u * kAccThreads + threadIdx.x) * 8;
      if (i + 8 <= n) {
        float f[8];
        Vec<__nv_bfloat16>::unpack(ld_stream(src + i), f);
        float4 a = *reinterpret_cast<const float4*>(dst + i), b = *reinterpret_cast<const float4*>(dst + i + 4);
        a.x += f[0]; a.y += f[1]; a.z += f[2]; a.w += f[3];
        b.x += f[4]; b.y += f[5]; b.z += f[6]; b.w += f[7];
        *reinterpret_cast<float4*>(dst + i) = a;
        *reinterpret_cast<float4*>(dst + i + 4) = b;
      } else {
        for (int64_t j = i; j < n; ++j) dst[j] += __bfloat162float(src[j]);
      }
    }
  } else {
    for (int64_t j = base + threadIdx.x; j < n && j < base + kAccElemsPerBlock; j += kAccThreads) dst[j] += __bfloat162float(src[j]);
  }
}

}  // namespace msml
