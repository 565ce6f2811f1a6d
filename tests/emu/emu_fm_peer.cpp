// Host build of csrc/fm_peer_kernels.cuh (peer-branch products + MSE) under the CPU emulation.  TEST INFRASTRUCTURE ONLY.
// Mirrors the launchers of csrc/fm_peer.cu with the grid given by the test.
#define MSML_CPU_EMU 1
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>

#include "../../msml_b200/csrc/fm_peer_kernels.cuh"

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

#define PEER_DISPATCH(dtype, mode, act, has_t, CALL)                                                                        \
  if (dtype == MSML_F32) { using T = float; PEER_MODES(mode, act, has_t, CALL) } else { using T = __nv_bfloat16; PEER_MODES(mode, act, has_t, CALL) }
#define PEER_MODES(mode, act, has_t, CALL)                                                                                  \
  if (mode == 0) { constexpr int MODE = 0, ACT = MSML_ACT_SIGMOID; PEER_T(has_t, CALL) }                                    \
  else if (act == MSML_ACT_SIGMOID) { constexpr int MODE = 1, ACT = MSML_ACT_SIGMOID; PEER_T(has_t, CALL) }                 \
  else { constexpr int MODE = 1, ACT = MSML_ACT_TANH; PEER_T(has_t, CALL) }
#define PEER_T(has_t, CALL) if (has_t) { constexpr bool HAS_T = true; CALL; } else { constexpr bool HAS_T = false; CALL; }

extern "C" int emu_fm_peer_mul_fwd(const void* src, const void* yf, const void* yt, void* pf, void* pt, int64_t n, int dtype, int mode,
                                   int act, int blocks) {
  const bool has_t = yt != nullptr;
  PEER_DISPATCH(dtype, mode, act, has_t,
                emu_launch(dim3(blocks), kPeerThreads, [&] {
                  fm_peer_mul_fwd_kernel<T, MODE, ACT, HAS_T>(static_cast<const T*>(src), static_cast<const T*>(yf), static_cast<const T*>(yt),
                                                             static_cast<T*>(pf), static_cast<T*>(pt), n);
                }));
  return 0;
}

extern "C" int emu_fm_peer_mul_bwd(const void* dpf, const void* dpt, const void* src, const void* yf, const void* yt, void* dsrc, void* dyf,
                                   int64_t n, int dtype, int mode, int act, int blocks) {
  const bool has_t = yt != nullptr;
  PEER_DISPATCH(dtype, mode, act, has_t,
                emu_launch(dim3(blocks), kPeerThreads, [&] {
                  fm_peer_mul_bwd_kernel<T, MODE, ACT, HAS_T>(static_cast<const T*>(dpf), static_cast<const T*>(dpt), static_cast<const T*>(src),
                                                             static_cast<const T*>(yf), static_cast<const T*>(yt), static_cast<T*>(dsrc),
                                                             static_cast<T*>(dyf), n);
                }));
  return 0;
}

extern "C" int emu_mse(const void* a, const void* b, int64_t n, int dtype, float* out, float gout, void* da, void* db, int blocks) {
  if (blocks < 1 || blocks > kMseMaxCtas) return 1;
  std::vector<float> partial(kMseMaxCtas, std::nanf(""));
  float* pp = partial.data();
  const float* gp = &gout;
  if (dtype == MSML_F32) {
    using T = float;
    emu_launch(dim3(blocks), kPeerThreads, [&] { mse_partial_kernel<T>(static_cast<const T*>(a), static_cast<const T*>(b), n, pp); });
    emu_launch(dim3(1), kPeerThreads, [&] { mse_finish_kernel(pp, blocks, n, out); });
    emu_launch(dim3(blocks), kPeerThreads, [&] { mse_bwd_kernel<T>(static_cast<const T*>(a), static_cast<const T*>(b), gp, static_cast<T*>(da), static_cast<T*>(db), n); });
  } else {
    using T = __nv_bfloat16;
    emu_launch(dim3(blocks), kPeerThreads, [&] { mse_partial_kernel<T>(static_cast<const T*>(a), static_cast<const T*>(b), n, pp); });
    emu_launch(dim3(1), kPeerThreads, [&] { mse_finish_kernel(pp, blocks, n, out); });
    emu_launch(dim3(blocks), kPeerThreads, [&] { mse_bwd_kernel<T>(static_cast<const T*>(a), static_cast<const T*>(b), gp, static_cast<T*>(da), static_cast<T*>(db), n); });
  }
  return 0;
}
