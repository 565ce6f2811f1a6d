// K-P  peer-guided branch of the Feature-Masking operator: the elementwise ends of it (launchers: fm_peer.cu; tests/emu runs
//      this header under the CPU emulation).
//   ref backbones/fm/fmoperator.py:293-302
//       m_bar = conv_m(gate)            (mask_trans 'conv': conv + BN;  'invert': 1 - gate, ref :160-166)
//       f_out = conv1(m_bar * yf);   f_occ = conv2(m_bar * yt);   l2 = MSELoss()(f_occ, f_out)
//   SURVEY.md 8(f)-3.  The reference runs two ATen multiplies (plus, for 'invert', the activation and the subtraction) and three
//   ATen kernels for the MSE (sub, pow, mean), each a full pass; here
//     fm_peer_mul_fwd / bwd : pf = m_bar*yf and pt = m_bar*yt in ONE pass (m_bar read once); MODE 1 takes the PRE-activation z
//                             and forms m_bar = 1 - act(z) in registers, so neither the gate nor its inverse is materialised;
//                             backward dm = dpf*yf + dpt*yt (through -act' for MODE 1), dyf = dpf*m_bar.  yt is the frozen
//                             teacher's feature map (ref peer/arcface.py:176-190 detaches it): it gets no gradient.
//     mse_partial / finish  : sum (a-b)^2 in fp32 from the storage dtype (what autocast's fp32 mse_loss computes), per-CTA
//                             partials then one CTA in fixed order: deterministic;  mse_bwd: da = 2(a-b)/n * g, db = -da.
//   All HBM-bound streaming kernels over a flat index (every operand shares one physical layout), 128-bit accesses.
#pragma once
#include "common.cuh"

namespace msml {

constexpr int kPeerThreads = 256;
constexpr int kPeerUnroll = 2;
constexpr int kMseMaxCtas = 148 * 8;

template <int MODE, int ACT>
__device__ __forceinline__ float peer_mbar(float src) {      // MODE 0: src is m_bar;  MODE 1: src is z, m_bar = 1 - act(z)
  return MODE == 0 ? src : 1.0f - gate_act<ACT>(src);
}

template <typename T, int MODE, int ACT, bool HAS_T>
__global__ void __launch_bounds__(kPeerThreads)
fm_peer_mul_fwd_kernel(const T* __restrict__ src, const T* __restrict__ yf, const T* __restrict__ yt, T* __restrict__ pf,
                       T* __restrict__ pt, int64_t n) {
  constexpr int VN = Vec<T>::N;
  const int64_t nvec = n / VN, stride = (int64_t)gridDim.x * kPeerThreads;
  for (int64_t base = (int64_t)blockIdx.x * kPeerThreads + threadIdx.x; base < nvec; base += stride * kPeerUnroll) {
    uint4 a[kPeerUnroll], b[kPeerUnroll], c[kPeerUnroll];
#pragma unroll
    for (int u = 0; u < kPeerUnroll; ++u) {
      const int64_t v = base + u * stride;
      if (v < nvec) {
        a[u] = ld_stream(src + v * VN);
        b[u] = ld_stream(yf + v * VN);
        if (HAS_T) c[u] = ld_stream(yt + v * VN);
      }
    }
#pragma unroll
    for (int u = 0; u < kPeerUnroll; ++u) {
      const int64_t v = base + u * stride;
      if (v >= nvec) break;
      float s[VN], f[VN], t[VN], of[VN], ot[VN];
      Vec<T>::unpack(a[u], s);
      Vec<T>::unpack(b[u], f);
      if (HAS_T) Vec<T>::unpack(c[u], t);
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        const float m = peer_mbar<MODE, ACT>(s[i]);
        of[i] = m * f[i];
        if (HAS_T) ot[i] = m * t[i];
      }
      *reinterpret_cast<uint4*>(pf + v * VN) = Vec<T>::pack(of);          // read next by a convolution: default policy
      if (HAS_T) *reinterpret_cast<uint4*>(pt + v * VN) = Vec<T>::pack(ot);
    }
  }
  if (blockIdx.x == 0) {                                                   // scalar tail (n % VN elements)
    for (int64_t i = nvec * VN + threadIdx.x; i < n; i += kPeerThreads) {
      const float m = peer_mbar<MODE, ACT>(to_f32(src[i]));
      pf[i] = from_f32<T>(m * to_f32(yf[i]));
      if (HAS_T) pt[i] = from_f32<T>(m * to_f32(yt[i]));
    }
  }
}

template <int MODE, int ACT>
__device__ __forceinline__ void peer_grad_one(float s, float f, float t, float dpf, float dpt, float& dsrc, float& dyf) {
  const float m = peer_mbar<MODE, ACT>(s);
  const float dm = fmaf(dpf, f, dpt * t);
  dyf = dpf * m;
  dsrc = MODE == 0 ? dm : -dm * gate_act_grad<ACT>(1.0f - m);              // d(1 - act(z))/dz = -act'(z), act' in terms of the gate
}

template <typename T, int MODE, int ACT, bool HAS_T>
__global__ void __launch_bounds__(kPeerThreads)
fm_peer_mul_bwd_kernel(const T* __restrict__ dpf, const T* __restrict__ dpt, const T* __restrict__ src, const T* __restrict__ yf,
                       const T* __restrict__ yt, T* __restrict__ dsrc, T* __restrict__ dyf, int64_t n) {
  constexpr int VN = Vec<T>::N;
  const int64_t nvec = n / VN, stride = (int64_t)gridDim.x * kPeerThreads;
  for (int64_t v = (int64_t)blockIdx.x * kPeerThreads + threadIdx.x; v < nvec; v += stride) {
    const uint4 a = ld_stream(dpf + v * VN), b = ld_stream(src + v * VN), c = ld_stream(yf + v * VN);
    uint4 d = a, e = a;
    if (HAS_T) { d = ld_stream(dpt + v * VN); e = ld_stream(yt + v * VN); }
    float gf[VN], gt[VN], s[VN], f[VN], t[VN], os[VN], oy[VN];
    Vec<T>::unpack(a, gf);
    Vec<T>::unpack(b, s);
    Vec<T>::unpack(c, f);
    if (HAS_T) { Vec<T>::unpack(d, gt); Vec<T>::unpack(e, t); }
#pragma unroll
    for (int i = 0; i < VN; ++i) peer_grad_one<MODE, ACT>(s[i], f[i], HAS_T ? t[i] : 0.f, gf[i], HAS_T ? gt[i] : 0.f, os[i], oy[i]);
    st_stream(dsrc + v * VN, Vec<T>::pack(os));
    st_stream(dyf + v * VN, Vec<T>::pack(oy));
  }
  if (blockIdx.x == 0) {
    for (int64_t i = nvec * VN + threadIdx.x; i < n; i += kPeerThreads) {
      float os, oy;
      peer_grad_one<MODE, ACT>(to_f32(src[i]), to_f32(yf[i]), HAS_T ? to_f32(yt[i]) : 0.f, to_f32(dpf[i]), HAS_T ? to_f32(dpt[i]) : 0.f, os, oy);
      dsrc[i] = from_f32<T>(os);
      dyf[i] = from_f32<T>(oy);
    }
  }
}

// ------------------------------------------------------------------------------------------- mean squared error
template <typename T>
__global__ void __launch_bounds__(kPeerThreads)
mse_partial_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t n, float* __restrict__ partial) {
  constexpr int VN = Vec<T>::N;
  __shared__ float red[kPeerThreads / 32];
  const int64_t nvec = n / VN, stride = (int64_t)gridDim.x * kPeerThreads;
  float acc = 0.f;
  for (int64_t v = (int64_t)blockIdx.x * kPeerThreads + threadIdx.x; v < nvec; v += stride) {
    float x[VN], y[VN];
    Vec<T>::unpack(ld_stream(a + v * VN), x);
    Vec<T>::unpack(ld_stream(b + v * VN), y);
#pragma unroll
    for (int i = 0; i < VN; ++i) { const float d = x[i] - y[i]; acc = fmaf(d, d, acc); }
  }
  if (blockIdx.x == 0) {
    for (int64_t i = nvec * VN + threadIdx.x; i < n; i += kPeerThreads) { const float d = to_f32(a[i]) - to_f32(b[i]); acc = fmaf(d, d, acc); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kPeerThreads / 32; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(kPeerThreads)
mse_finish_kernel(const float* __restrict__ partial, int count, int64_t n, float* __restrict__ out) {
  __shared__ float red[kPeerThreads / 32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < count; i += kPeerThreads) acc += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kPeerThreads / 32; ++w) s += red[w];
    *out = s / (float)n;
  }
}

// da = 2 (a - b) / n * (*gout),  db = -da   (db nullable)
template <typename T>
__global__ void __launch_bounds__(kPeerThreads)
mse_bwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const float* __restrict__ gout, T* __restrict__ da, T* __restrict__ db,
               int64_t n) {
  constexpr int VN = Vec<T>::N;
  const float k = 2.0f * *gout / (float)n;
  const int64_t nvec = n / VN, stride = (int64_t)gridDim.x * kPeerThreads;
  for (int64_t v = (int64_t)blockIdx.x * kPeerThreads + threadIdx.x; v < nvec; v += stride) {
    float x[VN], y[VN], o[VN];
    Vec<T>::unpack(ld_stream(a + v * VN), x);
    Vec<T>::unpack(ld_stream(b + v * VN), y);
#pragma unroll
    for (int i = 0; i < VN; ++i) o[i] = (x[i] - y[i]) * k;
    st_stream(da + v * VN, Vec<T>::pack(o));
    if (db) {
#pragma unroll
      for (int i = 0; i < VN; ++i) o[i] = -o[i];
      st_stream(db + v * VN, Vec<T>::pack(o));
    }
  }
  if (blockIdx.x == 0) {
    for (int64_t i = nvec * VN + threadIdx.x; i < n; i += kPeerThreads) {
      const float d = (to_f32(a[i]) - to_f32(b[i])) * k;
      da[i] = from_f32<T>(d);
      if (db) db[i] = from_f32<T>(-d);
    }
  }
}

}  // namespace msml
