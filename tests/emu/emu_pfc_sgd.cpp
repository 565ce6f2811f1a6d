// Host build of csrc/pfc_sgd_kernels.cuh under the CPU emulation (tests/emu/cuda_emu.h).  TEST INFRASTRUCTURE ONLY.
// Mirrors the launch of msml_pfc_sgd_update (csrc/optim.cu) on host memory.
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>

#include "../../msml_b200/csrc/pfc_sgd_kernels.cuh"

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

using namespace msml;

template <int V>
static void run(float* w, float* m, const float* dw, const int64_t* index, int64_t n_s, int64_t num_local, const float* lr_dev, SgdParams p,
                __nv_bfloat16* wn, float* inv) {
  const unsigned grid = (unsigned)((n_s + kSgdThreads / 32 - 1) / (kSgdThreads / 32));
  emu_launch(dim3(grid), kSgdThreads, [&] { pfc_sgd_kernel<V>(w, m, dw, index, n_s, num_local, lr_dev, p, wn, inv); });
}

extern "C" int emu_pfc_sgd_update(float* weight, float* weight_mom, const float* dw, const int64_t* index, int64_t n_s, int64_t num_local,
                                  int64_t D, const float* lr_dev, float lr, float momentum, float weight_decay, float dampening, int nesterov,
                                  void* wn_bf16, float* inv_norm) {
  if (int e = pfc_sgd_check(n_s, num_local, D)) return e;
  if (n_s == 0) return 0;
  SgdParams p{lr, momentum, weight_decay, dampening, nesterov};
  __nv_bfloat16* wn = static_cast<__nv_bfloat16*>(wn_bf16);
  switch (D / 128) {
    case 1: run<1>(weight, weight_mom, dw, index, n_s, num_local, lr_dev, p, wn, inv_norm); break;
    case 2: run<2>(weight, weight_mom, dw, index, n_s, num_local, lr_dev, p, wn, inv_norm); break;
    case 4: run<4>(weight, weight_mom, dw, index, n_s, num_local, lr_dev, p, wn, inv_norm); break;
    default: return set_error(MSML_EUNSUPPORTED, "emulation harness covers D in {128, 256, 512}");
  }
  return 0;
}

extern "C" const char* emu_sgd_last_error() { return msml::g_err; }
