// Host build of csrc/sgd_flat_kernels.cuh (flat momentum SGD + bf16 shadow) under the CPU emulation.  TEST INFRASTRUCTURE ONLY.
// Mirrors msml_sgd_flat (csrc/optim.cu) with the grid given by the test.
#define MSML_CPU_EMU 1
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>

#include "../../msml_b200/csrc/sgd_flat_kernels.cuh"

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

extern "C" int emu_sgd_flat(float* w, float* m, const float* g, void* shadow, int64_t n, float lr, const float* grad_scale, float momentum,
                            float weight_decay, int nesterov, int blocks) {
  if (n % 4 || blocks < 1) return 1;
  FlatSgdParams p{momentum, weight_decay, nesterov};
  const float* lr_p = &lr;
  emu_launch(dim3(blocks), kFlatSgdThreads, [&] { sgd_flat_kernel(w, m, g, static_cast<__nv_bfloat16*>(shadow), n / 4, lr_p, grad_scale, p); });
  return 0;
}
