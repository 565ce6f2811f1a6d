// K-C  input assembly of the Feature-Masking operator: channel concat + zero pad, NHWC, forward + backward:
//      the kernels and their geometry (launchers: fm_cat.cu; tests/emu runs this header under the CPU emulation).
//   ref backbones/fm/fmoperator.py:277-279  `x = torch.cat((yf, yo), dim=1)` feeding same_conv
//
//   fwd: cat[p, 0:C] = yf[p, :]   cat[p, C:C+Co] = yo[p, :]   cat[p, C+Co:Ct] = 0          (p = pixel row of N*H*W)
//   bwd: dyf[p, :] = dcat[p, 0:C] [+ dadd[p, :]]     [dyo[p, :] = dcat[p, C:C+Co]]
// Ct is C + Co rounded up to a whole number of 16-byte vectors, so that same_conv sees a channel count cuDNN's
// tensor-core kernels take without a padding pass of their own.  yf has two consumers in the operator (this concat and
// the fused tail, fm_gate.cu); `dadd` is the tail's gradient, so the backward kernel also does the sum autograd would
// do in one more (strided) pass over the sliced gradient.
// In ATen this is three strided copy kernels forward (one per part, 2-byte accesses into rows of Ct channels) and a
// strided add backward; here each direction is one pass with 128-bit accesses on the yf columns: HBM-bound,
// minimum traffic = every input read once, every output written once.
#pragma once
#include "common.cuh"

namespace msml {

constexpr int kCatThreads = 256;
constexpr int kCatUnroll = 4;
constexpr int kCatCtasPerSm = 4;       // 64 registers per thread: 4 + 4 vectors in flight in the backward loop

struct CatGeom {
  int64_t P;        // pixel rows
  int vf;           // 16-byte vectors of yf per row (C / VN)
  int vt;           // 16-byte vectors of cat per row (Ct / VN)
  int Co;           // channels of yo (any count; read with element accesses)
  int64_t nvec;     // vectors this launch walks: P * vt (fwd), P * vf (bwd)
};

// v -> (row, column vector) for a row of `per_row` vectors; 32-bit division when the index fits
__device__ __forceinline__ void split_index(int64_t v, int per_row, int64_t& row, int& j) {
  if (v <= 0xffffffffLL) {
    const uint32_t r = (uint32_t)v / (uint32_t)per_row;
    row = r;
    j = (int)((uint32_t)v - r * (uint32_t)per_row);
  } else {
    row = v / per_row;
    j = (int)(v - row * per_row);
  }
}

template <typename T>
__global__ void __launch_bounds__(kCatThreads, kCatCtasPerSm)
fm_cat_fwd_kernel(const T* __restrict__ yf, const T* __restrict__ yo, T* __restrict__ cat, CatGeom g) {
  constexpr int VN = Vec<T>::N;
  constexpr int U = kCatUnroll;
  const uint4* yf4 = reinterpret_cast<const uint4*>(yf);
  uint4* cat4 = reinterpret_cast<uint4*>(cat);
  const int64_t stride = (int64_t)gridDim.x * kCatThreads;
  for (int64_t base = (int64_t)blockIdx.x * kCatThreads + threadIdx.x; base < g.nvec; base += stride * U) {
    uint4 val[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = base + u * stride;
      if (v >= g.nvec) break;
      int64_t row;
      int j;
      split_index(v, g.vt, row, j);
      if (j < g.vf) {
        val[u] = ld_stream(yf4 + row * g.vf + j);
      } else {                                   // the yo columns and the zero padding: Co is not a whole vector
        const int e0 = (j - g.vf) * VN;
        const T* src = yo + row * g.Co + e0;
        float f[VN];
#pragma unroll
        for (int i = 0; i < VN; ++i) f[i] = (e0 + i < g.Co) ? to_f32(src[i]) : 0.f;     // widening and back is exact
        val[u] = Vec<T>::pack(f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = base + u * stride;
      if (v >= g.nvec) break;
      cat4[v] = val[u];                          // read next by same_conv: default cache policy
    }
  }
}

template <typename T, bool HAS_ADD>
__global__ void __launch_bounds__(kCatThreads, kCatCtasPerSm)
fm_cat_bwd_kernel(const T* __restrict__ dcat, const T* __restrict__ dadd, T* __restrict__ dyf, T* __restrict__ dyo, CatGeom g) {
  constexpr int VN = Vec<T>::N;
  constexpr int U = kCatUnroll;
  const uint4* dcat4 = reinterpret_cast<const uint4*>(dcat);
  const uint4* dadd4 = reinterpret_cast<const uint4*>(dadd);
  uint4* dyf4 = reinterpret_cast<uint4*>(dyf);
  const int64_t stride = (int64_t)gridDim.x * kCatThreads;
  for (int64_t base = (int64_t)blockIdx.x * kCatThreads + threadIdx.x; base < g.nvec; base += stride * U) {
    uint4 a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = base + u * stride;
      if (v >= g.nvec) break;
      int64_t row;
      int j;
      split_index(v, g.vf, row, j);
      a[u] = ld_stream(dcat4 + row * g.vt + j);
      if (HAS_ADD) b[u] = ld_stream(dadd4 + v);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = base + u * stride;
      if (v >= g.nvec) break;
      if (HAS_ADD) {
        float x[VN], y[VN];
        Vec<T>::unpack(a[u], x);
        Vec<T>::unpack(b[u], y);
#pragma unroll
        for (int i = 0; i < VN; ++i) x[i] += y[i];
        dyf4[v] = Vec<T>::pack(x);
      } else {
        dyf4[v] = a[u];
      }
    }
  }
  if (dyo != nullptr) {                          // only when the occlusion maps take part in the loss
    const int64_t n = g.P * g.Co;
    const int64_t c0 = (int64_t)g.vf * VN, ct = (int64_t)g.vt * VN;
    for (int64_t i = (int64_t)blockIdx.x * kCatThreads + threadIdx.x; i < n; i += stride) {
      const int64_t row = i / g.Co;
      const int e = (int)(i - row * g.Co);
      dyo[i] = dcat[row * ct + c0 + e];
    }
  }
}

inline int cat_geom(int64_t P, int64_t C, int64_t Co, int64_t Ct, int dtype, CatGeom* g) {
  MSML_REQUIRE(P > 0 && C >= 0 && Co >= 0 && Ct > 0, MSML_EINVAL, "bad concat shape P=%lld C=%lld Co=%lld Ct=%lld", (long long)P,
               (long long)C, (long long)Co, (long long)Ct);
  MSML_REQUIRE(dtype == MSML_F32 || dtype == MSML_BF16 || dtype == MSML_F16, MSML_EINVAL, "unknown dtype %d", dtype);
  const int vn = dtype == MSML_F32 ? 4 : 8;
  MSML_REQUIRE(C % vn == 0 && Ct % vn == 0, MSML_EUNSUPPORTED, "C=%lld and Ct=%lld must be multiples of %d", (long long)C,
               (long long)Ct, vn);
  MSML_REQUIRE(Ct >= C + Co, MSML_EINVAL, "Ct=%lld is smaller than C+Co=%lld", (long long)Ct, (long long)(C + Co));
  MSML_REQUIRE(Ct < (1 << 20), MSML_EUNSUPPORTED, "Ct=%lld too large", (long long)Ct);
  g->P = P; g->vf = (int)(C / vn); g->vt = (int)(Ct / vn); g->Co = (int)Co;
  return 0;
}

inline int cat_grid(int64_t nvec, int sms) {
  int64_t blocks = (nvec + (int64_t)kCatThreads * kCatUnroll - 1) / ((int64_t)kCatThreads * kCatUnroll);
  const int64_t cap = (int64_t)sms * kCatCtasPerSm;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace msml
