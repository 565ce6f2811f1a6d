// Momentum SGD over ONE flat fp32 parameter buffer: the kernel (launcher: optim.cu msml_sgd_flat; tests/emu runs this header
// under the CPU emulation).
//   The backbone optimizer of the training step (ref train.py:186-191 `torch.optim.SGD(backbone.parameters(), lr, momentum=0.9,
//   weight_decay=5e-4)`, stepped at train.py:299 after clip_grad_norm_).  engine.TrainStep already keeps every gradient in one
//   flat fp32 buffer; with the parameters and momentum buffers laid out the same way (engine.FlatSGD) the update of all ~48 M
//   backbone parameters is one streaming pass
//       g  = grad / grad_scale + weight_decay * w         (grad_scale: clip coefficient and 1/world_size, a device scalar)
//       m  = momentum * m + g                              (torch.optim.SGD with dampening 0; m starts at 0 == "first step m = g")
//       w -= lr * (nesterov ? g + momentum * m : m)
//       shadow = bf16(w)                                   (next step's autocast weights: no separate cast pass)
//   instead of torch's multi-tensor fused SGD (~300 tensors, 2x the HBM time of this pass) plus a multi-tensor fp32 -> bf16 copy.
//   Algorithmic bytes per element: 12 read + 8 written (+ 2 for the shadow); HBM-bound.
#pragma once
#include "common.cuh"

namespace msml {

constexpr int kFlatSgdThreads = 256;
constexpr int kFlatSgdUnroll = 4;       // float4 vectors in flight per thread and stream (3 read streams)

struct FlatSgdParams {
  float momentum, weight_decay;
  int nesterov;
};

__device__ __forceinline__ float flat_sgd_one(float& w, float& m, float g, float lr, float inv_valid, float scale, const FlatSgdParams& p) {
  if (inv_valid != 0.f) g = g / scale;                      // same operation as torch's fused SGD (grad /= *grad_scale)
  g = fmaf(p.weight_decay, w, g);
  if (p.momentum != 0.f) {
    m = fmaf(p.momentum, m, g);
    g = p.nesterov ? fmaf(p.momentum, m, g) : m;
  }
  w = fmaf(-lr, g, w);
  return w;
}

// n4 = number of float4 vectors (the flat buffers are padded to multiples of 4 elements; pad lanes hold zeros and stay zero)
__global__ void __launch_bounds__(kFlatSgdThreads)
sgd_flat_kernel(float* __restrict__ w, float* __restrict__ m, const float* __restrict__ g, __nv_bfloat16* __restrict__ shadow,
                int64_t n4, const float* __restrict__ lr_p, const float* __restrict__ grad_scale_p, FlatSgdParams p) {
  const float lr = *lr_p;
  const float scale = grad_scale_p ? *grad_scale_p : 1.f;
  const float has_scale = grad_scale_p ? 1.f : 0.f;
  const int64_t stride = (int64_t)gridDim.x * kFlatSgdThreads;
  for (int64_t i0 = (int64_t)blockIdx.x * kFlatSgdThreads + threadIdx.x; i0 < n4; i0 += stride * kFlatSgdUnroll) {
    uint4 wv[kFlatSgdUnroll], mv[kFlatSgdUnroll], gv[kFlatSgdUnroll];
#pragma unroll
    for (int u = 0; u < kFlatSgdUnroll; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4) {
        wv[u] = *reinterpret_cast<const uint4*>(w + i * 4);          // w and m are rewritten below: ordinary loads
        gv[u] = ld_stream(g + i * 4);
        if (p.momentum != 0.f) mv[u] = *reinterpret_cast<const uint4*>(m + i * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < kFlatSgdUnroll; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= n4) break;
      float wf[4], mf[4] = {0.f, 0.f, 0.f, 0.f}, gf[4];
      Vec<float>::unpack(wv[u], wf);
      Vec<float>::unpack(gv[u], gf);
      if (p.momentum != 0.f) Vec<float>::unpack(mv[u], mf);
#pragma unroll
      for (int k = 0; k < 4; ++k) flat_sgd_one(wf[k], mf[k], gf[k], lr, has_scale, scale, p);
      st_stream(w + i * 4, Vec<float>::pack(wf));                           // not read again before the next update: evict first
      if (p.momentum != 0.f) st_stream(m + i * 4, Vec<float>::pack(mf));     // (the bf16 shadows below feed the next step)
      if (shadow) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(wf[0], wf[1]), hi = __floats2bfloat162_rn(wf[2], wf[3]);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&lo);
        o.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(shadow + i * 4) = o;
      }
    }
  }
}

}  // namespace msml
