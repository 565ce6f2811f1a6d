// Host build of csrc/bn_act_kernels.cuh under the CPU emulation (tests/emu/cuda_emu.h).  TEST INFRASTRUCTURE ONLY.
// Mirrors the three-launch sequence of msml_bn_fwd / msml_bn_bwd (csrc/bn_act.cu: slab statistics on G1 CTAs, per-channel
// finalize on C CTAs, apply on G3 CTAs) on host memory.  These kernels are verified on a B200; here they get a logic
// re-run plus AddressSanitizer / ThreadSanitizer coverage (sanitize_main.cpp).
#define MSML_CPU_EMU 1
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>

#include "../../msml_b200/csrc/bn_act_kernels.cuh"

#ifndef MSML_EMU_NO_ERR
namespace msml {
static char g_err[512];
char* err_buf() { return g_err; }
int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace msml
#endif

using namespace msml;

static bool emu_bn_geom(int64_t P, int64_t C, int dtype, BnGeom* g) {
  const int vn = dtype == MSML_F32 ? 4 : 8;
  if (P <= 0 || C <= 0 || C % vn) return false;
  const int vpr = (int)(C / vn);
  if (vpr > kBnThreads || kBnThreads % vpr) return false;
  g->P = P; g->C = (int)C; g->vpr = vpr; g->rows_per_pass = kBnThreads / vpr; g->skip = 0; g->G = 1;
  return true;
}
static size_t emu_bn_ws_floats(int C) { return (size_t)kBnMaxCtas * 3 * C + kBnMaxCtas + 3 * (size_t)C + 16; }

template <typename T, bool RES, bool PRELU>
static void fwd3(const void* x, const void* res, void* y, const float* gamma, const float* beta, const float* prelu, float* rm, float* rv,
                 long long* nbt, float momentum, float eps, float* mean, float* invstd, float* ws, BnGeom g, int G1, int G3,
                 float* next_ws = nullptr, bool stats_ready = false) {
  float* part = ws;
  float* part_n = part + (size_t)kBnMaxCtas * 3 * g.C;
  float* coef = part_n + kBnMaxCtas;
  float* npart = next_ws;                                        // msml_bn_fwd_ex: statistics of y for the BN that follows
  float* npart_n = next_ws ? next_ws + (size_t)kBnMaxCtas * 3 * g.C : nullptr;
  const T* xp = static_cast<const T*>(x);
  const T* rp = static_cast<const T*>(res);
  T* yp = static_cast<T*>(y);
  if (stats_ready) {                                             // the producer's apply pass left the slab statistics in ws
    g.G = kBnMaxCtas;
  } else {
    g.G = G1;
    emu_launch(dim3(G1), kBnThreads, [&] { bn_fwd_fused_kernel<T, RES, PRELU, 1>(xp, rp, yp, gamma, beta, prelu, rm, rv, nbt, momentum, eps, mean, invstd, part, part_n, coef, g, npart, npart_n); });
  }
  emu_launch(dim3(g.C), kBnThreads, [&] { bn_fwd_fused_kernel<T, RES, PRELU, 2>(xp, rp, yp, gamma, beta, prelu, rm, rv, nbt, momentum, eps, mean, invstd, part, part_n, coef, g, npart, npart_n); });
  g.G = G3;
  if (next_ws)
    emu_launch(dim3(G3), kBnThreads, [&] { bn_fwd_fused_kernel<T, RES, PRELU, 3, true>(xp, rp, yp, gamma, beta, prelu, rm, rv, nbt, momentum, eps, mean, invstd, part, part_n, coef, g, npart, npart_n); });
  else
    emu_launch(dim3(G3), kBnThreads, [&] { bn_fwd_fused_kernel<T, RES, PRELU, 3>(xp, rp, yp, gamma, beta, prelu, rm, rv, nbt, momentum, eps, mean, invstd, part, part_n, coef, g, npart, npart_n); });
}

template <typename T, bool RES, bool PRELU>
static void bwd3(const void* dy, const void* x, const void* res, const float* mean, const float* invstd, const float* gamma, const float* beta,
                 const float* prelu, void* dx, void* dres, const void* dadd, float* dgamma, float* dbeta, float* dprelu, int training,
                 int accumulate, float* ws, BnGeom g, int G1, int G3) {
  float* part = ws;
  float* coef = part + (size_t)kBnMaxCtas * 3 * g.C + kBnMaxCtas;
  const T* dyp = static_cast<const T*>(dy);
  const T* xp = static_cast<const T*>(x);
  const T* rp = static_cast<const T*>(res);
  T* dxp = static_cast<T*>(dx);
  T* drp = static_cast<T*>(dres);
  const T* dap = static_cast<const T*>(dadd);
  g.G = G1;
  emu_launch(dim3(G1), kBnThreads, [&] { bn_bwd_fused_kernel<T, RES, PRELU, 1>(dyp, xp, rp, mean, invstd, gamma, beta, prelu, dxp, drp, dap, dgamma, dbeta, dprelu, training, accumulate, part, coef, g); });
  emu_launch(dim3(g.C), kBnThreads, [&] { bn_bwd_fused_kernel<T, RES, PRELU, 2>(dyp, xp, rp, mean, invstd, gamma, beta, prelu, dxp, drp, dap, dgamma, dbeta, dprelu, training, accumulate, part, coef, g); });
  g.G = G3;
  emu_launch(dim3(G3), kBnThreads, [&] { bn_bwd_fused_kernel<T, RES, PRELU, 3>(dyp, xp, rp, mean, invstd, gamma, beta, prelu, dxp, drp, dap, dgamma, dbeta, dprelu, training, accumulate, part, coef, g); });
}

#ifndef MSML_EMU_TEMPLATES_ONLY   // sanitize_main.cpp instantiates a few variants itself instead of all of them
#define BN_DISPATCH(dtype, has_res, has_prelu, CALL)                                                            \
  if (dtype == MSML_F32) { using T = float;                                                                     \
    if (has_res) { if (has_prelu) { constexpr bool RES = true, PRELU = true; CALL; } else { constexpr bool RES = true, PRELU = false; CALL; } } \
    else { if (has_prelu) { constexpr bool RES = false, PRELU = true; CALL; } else { constexpr bool RES = false, PRELU = false; CALL; } } }      \
  else { using T = __nv_bfloat16;                                                                               \
    if (has_res) { if (has_prelu) { constexpr bool RES = true, PRELU = true; CALL; } else { constexpr bool RES = true, PRELU = false; CALL; } } \
    else { if (has_prelu) { constexpr bool RES = false, PRELU = true; CALL; } else { constexpr bool RES = false, PRELU = false; CALL; } } }

extern "C" int emu_bn_fwd(const void* x, const void* res, void* y, const float* gamma, const float* beta, const float* prelu, float* rm, float* rv,
                          long long* nbt, float* mean, float* invstd, int64_t P, int64_t C, int dtype, float momentum, float eps, int G1, int G3) {
  BnGeom g;
  if (!emu_bn_geom(P, C, dtype, &g) || G1 < 1 || G3 < 1 || G1 > kBnMaxCtas || G3 > kBnMaxCtas) return 1;
  std::vector<float> ws(emu_bn_ws_floats((int)C));
  BN_DISPATCH(dtype, res != nullptr, prelu != nullptr,
              (fwd3<T, RES, PRELU>(x, res, y, gamma, beta, prelu, rm, rv, nbt, momentum, eps, mean, invstd, ws.data(), g, G1, G3)));
  return 0;
}

// Two chained ops as the residual trunk runs them: y1 = bn_a(x) + res (its apply pass emits the statistics of y1), then
// y2 = prelu(bn_b(y1)) starting at the merge.  G3a is the producer's apply grid, G3b the consumer's.
extern "C" int emu_bn_fwd_chain(const void* x, const void* res, void* y1, void* y2, const float* gamma_a, const float* beta_a,
                                const float* gamma_b, const float* beta_b, const float* prelu_b, float* rm_b, float* rv_b,
                                long long* nbt_b, float* mean_b, float* invstd_b, int64_t P, int64_t C, int dtype, float momentum,
                                float eps, int G1, int G3a, int G3b) {
  BnGeom g;
  if (!emu_bn_geom(P, C, dtype, &g) || G1 < 1 || G3a < 1 || G3b < 1 || G1 > kBnMaxCtas || G3a > kBnMaxCtas || G3b > kBnMaxCtas) return 1;
  std::vector<float> ws_a(emu_bn_ws_floats((int)C)), ws_b(emu_bn_ws_floats((int)C), std::nanf(""));   // stale slots must not matter
  std::vector<float> mean_a(C), invstd_a(C);
  if (dtype == MSML_F32) {
    fwd3<float, true, false>(x, res, y1, gamma_a, beta_a, nullptr, nullptr, nullptr, nullptr, momentum, eps, mean_a.data(), invstd_a.data(),
                             ws_a.data(), g, G1, G3a, ws_b.data());
    if (prelu_b) fwd3<float, false, true>(y1, nullptr, y2, gamma_b, beta_b, prelu_b, rm_b, rv_b, nbt_b, momentum, eps, mean_b, invstd_b, ws_b.data(), g, G1, G3b, nullptr, true);
    else fwd3<float, false, false>(y1, nullptr, y2, gamma_b, beta_b, nullptr, rm_b, rv_b, nbt_b, momentum, eps, mean_b, invstd_b, ws_b.data(), g, G1, G3b, nullptr, true);
  } else {
    using T = __nv_bfloat16;
    fwd3<T, true, false>(x, res, y1, gamma_a, beta_a, nullptr, nullptr, nullptr, nullptr, momentum, eps, mean_a.data(), invstd_a.data(),
                         ws_a.data(), g, G1, G3a, ws_b.data());
    if (prelu_b) fwd3<T, false, true>(y1, nullptr, y2, gamma_b, beta_b, prelu_b, rm_b, rv_b, nbt_b, momentum, eps, mean_b, invstd_b, ws_b.data(), g, G1, G3b, nullptr, true);
    else fwd3<T, false, false>(y1, nullptr, y2, gamma_b, beta_b, nullptr, rm_b, rv_b, nbt_b, momentum, eps, mean_b, invstd_b, ws_b.data(), g, G1, G3b, nullptr, true);
  }
  return 0;
}

extern "C" int emu_bn_bwd(const void* dy, const void* x, const void* res, const float* gamma, const float* beta, const float* prelu,
                          const float* mean, const float* invstd, void* dx, void* dres, const void* dadd, float* dgamma, float* dbeta,
                          float* dprelu, int64_t P, int64_t C, int dtype, int training, int accumulate, int G1, int G3) {
  BnGeom g;
  if (!emu_bn_geom(P, C, dtype, &g) || G1 < 1 || G3 < 1 || G1 > kBnMaxCtas || G3 > kBnMaxCtas) return 1;
  std::vector<float> ws(emu_bn_ws_floats((int)C));
  const bool has_res = res != nullptr, has_prelu = prelu != nullptr;
  BN_DISPATCH(dtype, has_res, has_prelu,
              (bwd3<T, RES, PRELU>(dy, x, res, mean, invstd, gamma, beta, prelu, dx, dres, dadd, dgamma, dbeta, dprelu, training, accumulate,
                                   ws.data(), g, G1, G3)));
  return 0;
}
#endif
