// K-P  peer-guided branch of the Feature-Masking operator: C-ABI launchers.  Kernels and algorithm: fm_peer_kernels.cuh
#include "fm_peer_kernels.cuh"

namespace msml {

static int peer_grid(int64_t n, int vn, int unroll) {
  const int64_t per_cta = (int64_t)kPeerThreads * vn * unroll;
  int64_t blocks = (n + per_cta - 1) / per_cta;
  const int64_t cap = (int64_t)num_sms() * 8;            // one resident wave (256 threads, <= 64 registers), grid-stride beyond
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

static int peer_check(int64_t n, int dtype, int mode, int act) {
  MSML_REQUIRE(n >= 0, MSML_EINVAL, "bad element count %lld", (long long)n);
  MSML_REQUIRE(dtype == MSML_F32 || dtype == MSML_BF16 || dtype == MSML_F16, MSML_EINVAL, "unknown dtype %d", dtype);
  MSML_REQUIRE(mode == 0 || mode == 1, MSML_EINVAL, "peer mask mode must be 0 (m_bar given) or 1 (1 - act(z)), got %d", mode);
  MSML_REQUIRE(mode == 0 || act == MSML_ACT_TANH || act == MSML_ACT_SIGMOID, MSML_EINVAL, "activation type error (%d)", act);
  return 0;
}

}  // namespace msml

using namespace msml;

// (MODE, ACT) pairs: MODE 0 ignores the activation
#define MSML_DISPATCH_PEER(mode, act, has_t, ...)                                                              \
  if ((mode) == 0) { constexpr int MODE = 0, ACT = MSML_ACT_SIGMOID;                                           \
    if (has_t) { constexpr bool HAS_T = true; __VA_ARGS__; } else { constexpr bool HAS_T = false; __VA_ARGS__; } } \
  else if ((act) == MSML_ACT_SIGMOID) { constexpr int MODE = 1, ACT = MSML_ACT_SIGMOID;                        \
    if (has_t) { constexpr bool HAS_T = true; __VA_ARGS__; } else { constexpr bool HAS_T = false; __VA_ARGS__; } } \
  else { constexpr int MODE = 1, ACT = MSML_ACT_TANH;                                                          \
    if (has_t) { constexpr bool HAS_T = true; __VA_ARGS__; } else { constexpr bool HAS_T = false; __VA_ARGS__; } }

extern "C" int msml_fm_peer_mul_fwd(const void* src, const void* yf, const void* yt, void* pf, void* pt, int64_t n, int dtype,
                                    int mode, int act, void* stream) {
  if (int e = peer_check(n, dtype, mode, act)) return e;
  if (n == 0) return 0;
  MSML_REQUIRE(src && yf && pf && ((yt == nullptr) == (pt == nullptr)), MSML_EINVAL, "null pointer (yt and pt come together)");
  MSML_REQUIRE(aligned16(src) && aligned16(yf) && aligned16(yt) && aligned16(pf) && aligned16(pt), MSML_EALIGN,
               "pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const bool has_t = yt != nullptr;
  const double elem = dtype == MSML_F32 ? 4.0 : 2.0;
  MSML_PROF("fm_peer_mul_fwd", (double)n * elem * (has_t ? 5 : 3), st);
  const int grid = peer_grid(n, dtype == MSML_F32 ? 4 : 8, kPeerUnroll);
  MSML_DISPATCH_DTYPE(dtype, T, MSML_DISPATCH_PEER(mode, act, has_t,
      (fm_peer_mul_fwd_kernel<T, MODE, ACT, HAS_T><<<grid, kPeerThreads, 0, st>>>(
          static_cast<const T*>(src), static_cast<const T*>(yf), static_cast<const T*>(yt), static_cast<T*>(pf), static_cast<T*>(pt), n))));
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" int msml_fm_peer_mul_bwd(const void* dpf, const void* dpt, const void* src, const void* yf, const void* yt, void* dsrc,
                                    void* dyf, int64_t n, int dtype, int mode, int act, void* stream) {
  if (int e = peer_check(n, dtype, mode, act)) return e;
  if (n == 0) return 0;
  MSML_REQUIRE(dpf && src && yf && dsrc && dyf && ((yt == nullptr) == (dpt == nullptr)), MSML_EINVAL,
               "null pointer (yt and dpt come together)");
  MSML_REQUIRE(aligned16(dpf) && aligned16(dpt) && aligned16(src) && aligned16(yf) && aligned16(yt) && aligned16(dsrc) && aligned16(dyf),
               MSML_EALIGN, "pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const bool has_t = yt != nullptr;
  const double elem = dtype == MSML_F32 ? 4.0 : 2.0;
  MSML_PROF("fm_peer_mul_bwd", (double)n * elem * (has_t ? 7 : 5), st);
  const int grid = peer_grid(n, dtype == MSML_F32 ? 4 : 8, 1);
  MSML_DISPATCH_DTYPE(dtype, T, MSML_DISPATCH_PEER(mode, act, has_t,
      (fm_peer_mul_bwd_kernel<T, MODE, ACT, HAS_T><<<grid, kPeerThreads, 0, st>>>(
          static_cast<const T*>(dpf), static_cast<const T*>(dpt), static_cast<const T*>(src), static_cast<const T*>(yf),
          static_cast<const T*>(yt), static_cast<T*>(dsrc), static_cast<T*>(dyf), n))));
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t msml_mse_workspace(void) { return (size_t)kMseMaxCtas * sizeof(float); }

extern "C" int msml_mse_fwd(const void* a, const void* b, int64_t n, int dtype, float* out, void* ws, size_t ws_bytes, void* stream) {
  MSML_REQUIRE(n > 0, MSML_EINVAL, "mse of %lld elements", (long long)n);
  MSML_REQUIRE(a && b && out, MSML_EINVAL, "null pointer");
  MSML_REQUIRE(ws && ws_bytes >= msml_mse_workspace(), MSML_EWORKSPACE, "mse workspace too small");
  MSML_REQUIRE(aligned16(a) && aligned16(b) && aligned16(ws), MSML_EALIGN, "pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int vn = dtype == MSML_F32 ? 4 : 8;
  int grid = peer_grid(n, vn, 4);
  if (grid > kMseMaxCtas) grid = kMseMaxCtas;
  float* partial = static_cast<float*>(ws);
  {
    MSML_PROF("mse_fwd", (double)n * (dtype == MSML_F32 ? 4.0 : 2.0) * 2, st);
    MSML_DISPATCH_DTYPE(dtype, T, (mse_partial_kernel<T><<<grid, kPeerThreads, 0, st>>>(static_cast<const T*>(a), static_cast<const T*>(b), n, partial)));
    MSML_LAUNCH_CHECK();
  }
  mse_finish_kernel<<<1, kPeerThreads, 0, st>>>(partial, grid, n, out);
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" int msml_mse_bwd(const void* a, const void* b, const float* gout, void* da, void* db, int64_t n, int dtype, void* stream) {
  MSML_REQUIRE(n > 0, MSML_EINVAL, "mse of %lld elements", (long long)n);
  MSML_REQUIRE(a && b && gout && da, MSML_EINVAL, "null pointer");
  MSML_REQUIRE(aligned16(a) && aligned16(b) && aligned16(da) && aligned16(db), MSML_EALIGN, "pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int vn = dtype == MSML_F32 ? 4 : 8;
  const int grid = peer_grid(n, vn, 1);
  MSML_PROF("mse_bwd", (double)n * (dtype == MSML_F32 ? 4.0 : 2.0) * (db ? 4 : 3), st);
  MSML_DISPATCH_DTYPE(dtype, T, (mse_bwd_kernel<T><<<grid, kPeerThreads, 0, st>>>(static_cast<const T*>(a), static_cast<const T*>(b), gout,
                                                                                static_cast<T*>(da), static_cast<T*>(db), n)));
  MSML_LAUNCH_CHECK();
  return 0;
}
