// K-B  DAP head of the occlusion-segmentation branch (+ fused 2-way argmax mask): the kernels (launchers: dap.cu;
//      tests/emu runs this header under the CPU emulation: the bit-exact argmax mask is checked on the CPU test tier too).
//   ref backbones/osb/unet.py:158-161,223: PixelShuffle(k) -> AvgPool2d(k)  ==  mean over the k*k
//   consecutive input channels of each output channel (the (B,G,kH,kW) intermediate the reference
//   materialises never exists here);  train.py:357: final_seg[b].max(0)[1].
// One thread per pixel (per pixel pair for 16-bit NCHW): reads G*kk values, writes G (+1 int64).
// The mean and the argmax are taken on the fp32 accumulator, before rounding to the output dtype.
#pragma once
#include "common.cuh"

namespace msml {

template <typename T, bool CL, bool MASK>
__global__ void __launch_bounds__(256)
dap_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t* __restrict__ mask,
               int64_t B, int G, int kk, int64_t HW) {
  const int64_t total = B * HW;
  const float inv = 1.0f / (float)kk;
  (void)inv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / HW, p = i - b * HW;
    float best = 0.f;
    int64_t arg = 0;
    for (int g = 0; g < G; ++g) {
      float acc = 0.f;
      for (int j = 0; j < kk; ++j) {
        const int64_t c = (int64_t)g * kk + j;
        const int64_t idx = CL ? (i * G * kk + c) : ((b * G * kk + c) * HW + p);
        acc += to_f32(x[idx]);
      }
      acc = acc / (float)kk;   // ATen avg_pool2d: sum / divide_factor
      const int64_t oidx = CL ? (i * G + g) : ((b * G + g) * HW + p);
      y[oidx] = from_f32<T>(acc);
      if (MASK && (g == 0 || acc > best)) { best = acc; arg = g; }   // strict '>' : first index on ties
    }
    if (MASK) mask[i] = arg;
  }
}

template <typename T, bool CL>
__global__ void __launch_bounds__(256)
dap_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int64_t B, int G, int kk, int64_t HW) {
  const int64_t total = B * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / HW, p = i - b * HW;
    for (int g = 0; g < G; ++g) {
      const int64_t oidx = CL ? (i * G + g) : ((b * G + g) * HW + p);
      const T v = from_f32<T>(to_f32(dy[oidx]) / (float)kk);
      for (int j = 0; j < kk; ++j) {
        const int64_t c = (int64_t)g * kk + j;
        const int64_t idx = CL ? (i * G * kk + c) : ((b * G * kk + c) * HW + p);
        dx[idx] = v;
      }
    }
  }
}

}  // namespace msml
