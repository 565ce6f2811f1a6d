// Host build of csrc/fm_cat_kernels.cuh under the CPU emulation (tests/emu/cuda_emu.h).  TEST INFRASTRUCTURE ONLY.
// These kernels ARE verified on a B200 (tests/test_gpu_fusion.py): running them here as well cross-checks the emulation
// itself against a kernel whose GPU behaviour is known.
#define MSML_CPU_EMU 1
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>

#include "../../msml_b200/csrc/fm_cat_kernels.cuh"

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

template <typename T>
static void fwd(const void* yf, const void* yo, void* cat, const CatGeom& g, int sms) {
  emu_launch(dim3((unsigned)cat_grid(g.nvec, sms)), kCatThreads,
             [&] { fm_cat_fwd_kernel<T>(static_cast<const T*>(yf), static_cast<const T*>(yo), static_cast<T*>(cat), g); });
}
template <typename T>
static void bwd(const void* dcat, const void* dadd, void* dyf, void* dyo, const CatGeom& g, int64_t walk, int sms) {
  const dim3 grid((unsigned)cat_grid(walk, sms));
  if (dadd)
    emu_launch(grid, kCatThreads, [&] { fm_cat_bwd_kernel<T, true>(static_cast<const T*>(dcat), static_cast<const T*>(dadd), static_cast<T*>(dyf), static_cast<T*>(dyo), g); });
  else
    emu_launch(grid, kCatThreads, [&] { fm_cat_bwd_kernel<T, false>(static_cast<const T*>(dcat), nullptr, static_cast<T*>(dyf), static_cast<T*>(dyo), g); });
}

// `sms` plays the role of the SM count: a small value forces the grid-stride loops to take several trips
extern "C" int emu_fm_cat_fwd(const void* yf, const void* yo, void* cat, int64_t P, int64_t C, int64_t Co, int64_t Ct, int dtype, int sms) {
  CatGeom g;
  if (int e = cat_geom(P, C, Co, Ct, dtype, &g)) return e;
  g.nvec = P * g.vt;
  if (dtype == MSML_F32) fwd<float>(yf, yo, cat, g, sms); else fwd<__nv_bfloat16>(yf, yo, cat, g, sms);
  return 0;
}
extern "C" int emu_fm_cat_bwd(const void* dcat, const void* dadd, void* dyf, void* dyo, int64_t P, int64_t C, int64_t Co, int64_t Ct, int dtype,
                              int sms) {
  CatGeom g;
  if (int e = cat_geom(P, C, Co, Ct, dtype, &g)) return e;
  g.nvec = P * g.vf;
  const int64_t walk = g.nvec > 0 ? g.nvec : (dyo ? (P * Co + 7) / 8 : 1);
  if (dtype == MSML_F32) bwd<float>(dcat, dadd, dyf, dyo, g, walk, sms); else bwd<__nv_bfloat16>(dcat, dadd, dyf, dyo, g, walk, sms);
  return 0;
}
