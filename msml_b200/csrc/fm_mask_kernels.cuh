// K-A (extension)  low-resolution / single-channel mask fused into NHWC feature maps: the kernels (launchers: fm_mask.cu;
// tests/emu runs this header under the CPU emulation and its sanitizers).
// BASELINE.json north_star: "decoder mask logits resized to each feature scale, normalised into
// gates and multiplied into the feature maps, forward and backward".  The reference itself never
// resizes or broadcasts (its gate is C-channel at feature resolution, handled by
// fm_gate.cu); this family covers the general case with the same act/arith options
// (ref backbones/fm/fmoperator.py:71-81,113-117).
//
// A CTA owns kPix consecutive pixels x C channels.  Phase 1 stages the RESIZED mask tile in
// shared memory as gates (the activation is evaluated once per pixel, not once per element);
// phase 2 streams yf/out with 128-bit NHWC accesses.  Backward reduces dgate over channels with
// warp shuffles, accumulates the pixel sums in shared memory and scatters them through the
// resize fan-out into the fp32 mask gradient.
#pragma once
#include "common.cuh"

namespace msml {

constexpr int kPix = 32;
constexpr int kMaskThreads = 256;

struct MaskGeom {
  int64_t B, H, W, C, Hm, Wm, Cm;
};

__device__ __forceinline__ int64_t mask_index(const MaskGeom& g, int64_t pix) {
  const int64_t w = pix % g.W, h = (pix / g.W) % g.H, b = pix / (g.W * g.H);
  const int64_t mh = (h * g.Hm) / g.H, mw = (w * g.Wm) / g.W;   // nearest (floor), exact in integers
  return (b * g.Hm + mh) * g.Wm + mw;
}

template <typename T, int ACT, int ARITH>
__global__ void __launch_bounds__(kMaskThreads)
fm_mask_fwd_kernel(const T* __restrict__ yf, const T* __restrict__ m, T* __restrict__ out, MaskGeom g) {
  constexpr int VN = Vec<T>::N;
  __shared__ float gate[kPix];
  const int64_t npix = g.B * g.H * g.W;
  const int64_t pix0 = (int64_t)blockIdx.x * kPix;
  const int vec_per_pix = (int)(g.C / VN);
  if (g.Cm == 1) {
    if (threadIdx.x < kPix && pix0 + threadIdx.x < npix)
      gate[threadIdx.x] = gate_act<ACT>(to_f32(m[mask_index(g, pix0 + threadIdx.x)]));
    __syncthreads();
  }
  const int nvec = kPix * vec_per_pix;
  for (int v = threadIdx.x; v < nvec; v += kMaskThreads) {
    const int p = v / vec_per_pix, cv = v - p * vec_per_pix;
    const int64_t pix = pix0 + p;
    if (pix >= npix) break;
    const int64_t off = pix * g.C + (int64_t)cv * VN;
    float f[VN], z[VN], o[VN];
    Vec<T>::unpack(ld_stream(yf + off), f);
    if (g.Cm != 1) Vec<T>::unpack(*reinterpret_cast<const uint4*>(m + mask_index(g, pix) * g.C + (int64_t)cv * VN), z);
#pragma unroll
    for (int i = 0; i < VN; ++i) o[i] = gate_fuse<ARITH>(f[i], g.Cm == 1 ? gate[p] : gate_act<ACT>(z[i]));
    st_stream(out + off, Vec<T>::pack(o));
  }
}

template <typename T, int ACT, int ARITH>
__global__ void __launch_bounds__(kMaskThreads)
fm_mask_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ yf, const T* __restrict__ m, T* __restrict__ dyf,
                   float* __restrict__ dm, MaskGeom g) {
  constexpr int VN = Vec<T>::N;
  __shared__ float gate[kPix];
  __shared__ float dgate[kPix];
  const int64_t npix = g.B * g.H * g.W;
  const int64_t pix0 = (int64_t)blockIdx.x * kPix;
  const int vec_per_pix = (int)(g.C / VN);
  if (g.Cm == 1) {
    if (threadIdx.x < kPix) {
      dgate[threadIdx.x] = 0.f;
      if (pix0 + threadIdx.x < npix) gate[threadIdx.x] = gate_act<ACT>(to_f32(m[mask_index(g, pix0 + threadIdx.x)]));
    }
    __syncthreads();
  }
  // lanes of one pixel form a contiguous, power-of-two lane group (C/VN is a power of two here)
  const int group = vec_per_pix < 32 ? vec_per_pix : 32;
  const int nvec = kPix * vec_per_pix;
  const int nvec_pad = (nvec + kMaskThreads - 1) / kMaskThreads * kMaskThreads;   // keep warps convergent for shuffles
  for (int v = threadIdx.x; v < nvec_pad; v += kMaskThreads) {
    const int p = v / vec_per_pix, cv = v - p * vec_per_pix;
    const int64_t pix = pix0 + p;
    const bool ok = v < nvec && pix < npix;
    float dg_sum = 0.f;
    if (ok) {
      const int64_t off = pix * g.C + (int64_t)cv * VN;
      float d[VN], f[VN], z[VN], o[VN];
      Vec<T>::unpack(ld_stream(dout + off), d);
      Vec<T>::unpack(ld_stream(yf + off), f);
      const int64_t moff = g.Cm != 1 ? mask_index(g, pix) * g.C + (int64_t)cv * VN : 0;
      if (g.Cm != 1) Vec<T>::unpack(*reinterpret_cast<const uint4*>(m + moff), z);
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        const float gt = g.Cm == 1 ? gate[p] : gate_act<ACT>(z[i]);
        float dg;
        gate_fuse_grad<ARITH>(d[i], f[i], gt, o[i], dg);
        if (g.Cm == 1) dg_sum += dg;
        else atomicAdd(dm + moff + i, dg * gate_act_grad<ACT>(gt));   // per-channel mask: scatter through the resize
      }
      st_stream(dyf + off, Vec<T>::pack(o));
    }
    if (g.Cm == 1) {
      for (int o = group >> 1; o > 0; o >>= 1) dg_sum += __shfl_xor_sync(0xffffffffu, dg_sum, o);
      if (ok && (threadIdx.x & (group - 1)) == 0) atomicAdd(&dgate[p], dg_sum);
    }
  }
  if (g.Cm == 1) {
    __syncthreads();
    if (threadIdx.x < kPix && pix0 + threadIdx.x < npix)
      atomicAdd(dm + mask_index(g, pix0 + threadIdx.x), dgate[threadIdx.x] * gate_act_grad<ACT>(gate[threadIdx.x]));
  }
}

}  // namespace msml
