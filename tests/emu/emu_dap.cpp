// Host build of csrc/dap_kernels.cuh (K-B: DAP + fused argmax mask) under the CPU emulation.  TEST INFRASTRUCTURE ONLY.
#define MSML_CPU_EMU 1
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>

#include "../../msml_b200/csrc/dap_kernels.cuh"

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

template <typename T>
static void dap_fwd(const void* x, void* y, int64_t* mask, int64_t B, int G, int kk, int64_t HW, int cl, int grid) {
  const T* xi = static_cast<const T*>(x);
  T* yo = static_cast<T*>(y);
  if (cl) {
    if (mask) emu_launch(dim3(grid), 256, [&] { dap_fwd_kernel<T, true, true>(xi, yo, mask, B, G, kk, HW); });
    else emu_launch(dim3(grid), 256, [&] { dap_fwd_kernel<T, true, false>(xi, yo, mask, B, G, kk, HW); });
  } else {
    if (mask) emu_launch(dim3(grid), 256, [&] { dap_fwd_kernel<T, false, true>(xi, yo, mask, B, G, kk, HW); });
    else emu_launch(dim3(grid), 256, [&] { dap_fwd_kernel<T, false, false>(xi, yo, mask, B, G, kk, HW); });
  }
}
template <typename T>
static void dap_bwd(const void* dy, void* dx, int64_t B, int G, int kk, int64_t HW, int cl, int grid) {
  if (cl) emu_launch(dim3(grid), 256, [&] { dap_bwd_kernel<T, true>(static_cast<const T*>(dy), static_cast<T*>(dx), B, G, kk, HW); });
  else emu_launch(dim3(grid), 256, [&] { dap_bwd_kernel<T, false>(static_cast<const T*>(dy), static_cast<T*>(dx), B, G, kk, HW); });
}

// `grid` CTAs of 256 threads (a small grid makes the grid-stride loop take several trips)
extern "C" void emu_dap_fwd(const void* x, void* y, int64_t* mask, int64_t B, int64_t G, int64_t kk, int64_t HW, int cl, int dtype, int grid) {
  if (dtype == MSML_F32) dap_fwd<float>(x, y, mask, B, (int)G, (int)kk, HW, cl, grid);
  else dap_fwd<__nv_bfloat16>(x, y, mask, B, (int)G, (int)kk, HW, cl, grid);
}
extern "C" void emu_dap_bwd(const void* dy, void* dx, int64_t B, int64_t G, int64_t kk, int64_t HW, int cl, int dtype, int grid) {
  if (dtype == MSML_F32) dap_bwd<float>(dy, dx, B, (int)G, (int)kk, HW, cl, grid);
  else dap_bwd<__nv_bfloat16>(dy, dx, B, (int)G, (int)kk, HW, cl, grid);
}
