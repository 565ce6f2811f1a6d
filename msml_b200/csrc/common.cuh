// Shared helpers for the msml_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/msml_b200.h"

namespace msml {

// ---- error reporting (thread-local, returned by msml_last_error) --------------------------
char* err_buf();
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define MSML_REQUIRE(cond, code, ...)                         \
  do {                                                        \
    if (!(cond)) return ::msml::set_error((code), __VA_ARGS__); \
  } while (0)

#define MSML_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::msml::set_error((int)_e, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                       \
  } while (0)

#define MSML_LAUNCH_CHECK()                                                               \
  do {                                                                                    \
    ::msml::count_launch();                                                               \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess)                                                                \
      return ::msml::set_error((int)_e, "kernel launch failed: %s (%s:%d)",               \
                               cudaGetErrorString(_e), __FILE__, __LINE__);               \
  } while (0)

// ---- optional per-launch timing (CUDA events on the launching stream; msml_profile_*) ----------
// `work` is the ALGORITHMIC work of the launch: bytes for HBM-bound kernels, flops for GEMMs.
// `bytes` (optional) is the minimum HBM traffic of a contraction, so that short-M GEMMs can be
// reported against the roofline that actually binds them.
struct ProfScope {
  int slot;
  cudaStream_t st;
  ProfScope(const char* name, double work, cudaStream_t stream, double bytes = 0.0);
  ~ProfScope();
};
#define MSML_PROF(name, work, stream) ::msml::ProfScope _prof_scope((name), (double)(work), (stream))
#define MSML_PROF2(name, work, bytes, stream) ::msml::ProfScope _prof_scope((name), (double)(work), (stream), (double)(bytes))

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int num_sms();

// ---- 128-bit streaming accesses -------------------------------------------------------------
// Read-once / write-once streams: bypass L1 allocation on loads, evict-first on stores.
#ifndef MSML_CPU_EMU
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(void* p, const uint4& v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
// L2 residency hints for a slab that one kernel reads and a LATER kernel of the same op reads again (two-pass BatchNorm):
// the first read tags its lines evict_last so that unrelated streams (a convolution on the side stream) are evicted before
// them, the second read tags them evict_first so that they leave the cache once used.
__device__ __forceinline__ uint64_t l2_policy_keep() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_drop() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint4 ld_stream_hint(const void* p, uint64_t policy) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(policy));
  return r;
}
#else   // host build under tests/emu/cuda_emu.h (logic checks without a GPU): cache hints have no meaning there
inline uint4 ld_stream(const void* p) { return *static_cast<const uint4*>(p); }
inline void st_stream(void* p, const uint4& v) { *static_cast<uint4*>(p) = v; }
inline uint64_t l2_policy_keep() { return 0; }
inline uint64_t l2_policy_drop() { return 0; }
inline uint64_t l2_policy_normal() { return 0; }
inline uint4 ld_stream_hint(const void* p, uint64_t) { return *static_cast<const uint4*>(p); }
#endif

// ---- element packs: one 16-byte vector = VEC<T>::N elements ------------------------------------
template <typename T>
struct Vec;
template <>
struct Vec<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <>
struct Vec<__half> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// dtype dispatch for host launchers
#define MSML_DISPATCH_DTYPE(dtype, T, ...)                                         \
  switch (dtype) {                                                                 \
    case MSML_F32: { using T = float; __VA_ARGS__; } break;                        \
    case MSML_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break;               \
    case MSML_F16: { using T = __half; __VA_ARGS__; } break;                       \
    default: return ::msml::set_error(MSML_EINVAL, "unknown dtype %d", (int)dtype); \
  }

// ---- gate math (ref fmoperator.py:71-81,113-117) --------------------------------------------
template <int ACT>
__device__ __forceinline__ float gate_act(float z) {
  if (ACT == MSML_ACT_SIGMOID) return 1.0f / (1.0f + __expf(-z));
  // tanh(z) = 2*sigmoid(2z) - 1, saturates cleanly for |z| large
  return 2.0f / (1.0f + __expf(-2.0f * z)) - 1.0f;
}
template <int ACT>
__device__ __forceinline__ float gate_act_grad(float g) {
  return ACT == MSML_ACT_SIGMOID ? g * (1.0f - g) : 1.0f - g * g;
}
template <int ARITH>
__device__ __forceinline__ float gate_fuse(float yf, float g) {
  // arith(yf, g) + yf
  if (ARITH == MSML_ARITH_ADD) return yf + g + yf;
  if (ARITH == MSML_ARITH_SUB) return yf - g + yf;
  if (ARITH == MSML_ARITH_DIV) return __fdividef(yf, g) + yf;
  return fmaf(yf, g, yf);
}
// d = dout; returns dyf (direct) and dg
template <int ARITH>
__device__ __forceinline__ void gate_fuse_grad(float d, float yf, float g, float& dyf, float& dg) {
  if (ARITH == MSML_ARITH_ADD) { dyf = 2.0f * d; dg = d; }
  else if (ARITH == MSML_ARITH_SUB) { dyf = 2.0f * d; dg = -d; }
  else if (ARITH == MSML_ARITH_DIV) { float r = __fdividef(1.0f, g); dyf = fmaf(d, r, d); dg = -d * yf * r * r; }
  else { dyf = fmaf(d, g, d); dg = d * yf; }
}

}  // namespace msml
