// Host build of csrc/fm_mask_kernels.cuh (K-A extension: resized / broadcast mask staged in shared memory, dMask reduced by
// warp shuffles + shared and global float atomics) under the CPU emulation.  TEST INFRASTRUCTURE ONLY.
#define MSML_CPU_EMU 1
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>

#include "../../msml_b200/csrc/fm_mask_kernels.cuh"

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

// sigmoid / mul and tanh / add, fp32 and bf16: the combinations the tests use
template <typename T, int ACT, int ARITH>
static void mask_fwd(const void* yf, const void* m, void* out, const MaskGeom& g) {
  const unsigned grid = (unsigned)((g.B * g.H * g.W + kPix - 1) / kPix);
  emu_launch(dim3(grid), kMaskThreads,
             [&] { fm_mask_fwd_kernel<T, ACT, ARITH>(static_cast<const T*>(yf), static_cast<const T*>(m), static_cast<T*>(out), g); });
}
template <typename T, int ACT, int ARITH>
static void mask_bwd(const void* dout, const void* yf, const void* m, void* dyf, float* dm, const MaskGeom& g) {
  const unsigned grid = (unsigned)((g.B * g.H * g.W + kPix - 1) / kPix);
  emu_launch(dim3(grid), kMaskThreads, [&] {
    fm_mask_bwd_kernel<T, ACT, ARITH>(static_cast<const T*>(dout), static_cast<const T*>(yf), static_cast<const T*>(m), static_cast<T*>(dyf), dm, g);
  });
}

#ifndef MSML_EMU_TEMPLATES_ONLY
extern "C" int emu_fm_mask(const void* dout, const void* yf, const void* m, void* out, void* dyf, float* dm, int64_t B, int64_t H, int64_t W,
                           int64_t C, int64_t Hm, int64_t Wm, int64_t Cm, int dtype, int sigmoid_mul) {
  const MaskGeom g{B, H, W, C, Hm, Wm, Cm};
  if (dtype == MSML_F32) {
    if (sigmoid_mul) { mask_fwd<float, 1, 3>(yf, m, out, g); mask_bwd<float, 1, 3>(dout, yf, m, dyf, dm, g); }
    else { mask_fwd<float, 0, 0>(yf, m, out, g); mask_bwd<float, 0, 0>(dout, yf, m, dyf, dm, g); }
  } else {
    if (sigmoid_mul) { mask_fwd<__nv_bfloat16, 1, 3>(yf, m, out, g); mask_bwd<__nv_bfloat16, 1, 3>(dout, yf, m, dyf, dm, g); }
    else { mask_fwd<__nv_bfloat16, 0, 0>(yf, m, out, g); mask_bwd<__nv_bfloat16, 0, 0>(dout, yf, m, dyf, dm, g); }
  }
  return 0;
}
#endif
