// Host build of csrc/accum_kernels.cuh (multi-tensor bf16 -> fp32 gradient accumulate) under the CPU emulation.  TEST INFRASTRUCTURE ONLY.
// Mirrors msml_accum_bf16_multi (csrc/optim.cu): segments are packed into block ranges, one launch per MSML_ACCUM_MAX_SEGMENTS.
#define MSML_CPU_EMU 1
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>

#include "../../msml_b200/csrc/accum_kernels.cuh"

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

extern "C" int emu_accum_bf16_multi(int nseg, float* const* dst, const void* const* src, const int64_t* n) {
  for (int s0 = 0; s0 < nseg; s0 += MSML_ACCUM_MAX_SEGMENTS) {
    AccumSegs segs;
    int cnt = 0, blocks = 0;
    for (int i = s0; i < nseg && cnt < MSML_ACCUM_MAX_SEGMENTS; ++i) {
      if (n[i] <= 0) continue;
      segs.dst[cnt] = dst[i];
      segs.src[cnt] = static_cast<const __nv_bfloat16*>(src[i]);
      segs.n[cnt] = n[i];
      blocks += (int)((n[i] + kAccElemsPerBlock - 1) / kAccElemsPerBlock);
      segs.block_end[cnt] = blocks;
      ++cnt;
    }
    if (cnt == 0) continue;
    segs.nseg = cnt;
    emu_launch(dim3(blocks), kAccThreads, [&] { accum_bf16_multi_kernel(segs); });
  }
  return 0;
}
