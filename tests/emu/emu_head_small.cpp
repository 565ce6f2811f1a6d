// Host build of csrc/head_small_kernels.cuh (margin math, unit-norm bf16 cast, softmax-statistics merges / loss, margin
// forward and backward on a cosine matrix) under the CPU emulation.  TEST INFRASTRUCTURE ONLY.  Mirrors the launches of
// msml_wnorm_cast / msml_head_merge_stats / msml_margin_fwd / msml_margin_bwd (csrc/head.cu); the kernels are verified on a B200.
#define MSML_CPU_EMU 1
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>

#include "../../msml_b200/csrc/head_small_kernels.cuh"

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
using namespace msml::tc;

extern "C" void emu_wnorm_cast(const float* w, void* wn, float* inv_norm, int64_t n, int64_t D, int normalize) {
  emu_launch(dim3((unsigned)((n + 7) / 8)), 256, [&] { wnorm_cast_kernel(w, static_cast<__nv_bfloat16*>(wn), inv_norm, n, (int)D, normalize != 0); });
}
extern "C" void emu_transpose_bf16(const void* src, void* dst, int64_t rows, int64_t cols, int64_t ld_t) {
  emu_launch(dim3((unsigned)((rows + 63) / 64), (unsigned)((cols + 63) / 64)), 256, [&] {
    transpose_bf16_kernel(static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), rows, (int)cols, ld_t);
  });
}
extern "C" void emu_head_local_stats(const float* part_max, const float* part_sum, const float* tgt, const int64_t* tl, int n_blocks, int B_tot,
                                     float* stats) {
  emu_launch(dim3((unsigned)((B_tot + 7) / 8)), 256, [&] { head_local_stats_kernel(part_max, part_sum, tgt, tl, n_blocks, B_tot, stats); });
}
extern "C" void emu_head_merge_stats(const float* gathered, int W, int B_tot, float* gstats, float* loss) {
  emu_launch(dim3(1), 256, [&] { head_merge_kernel(gathered, W, B_tot, gstats, loss); });
}
extern "C" void emu_margin_fwd(float* cosm, const int64_t* label, int64_t B, int64_t C, int64_t ld, int kind, float s, float m, float a, float k) {
  Margin mg{kind, s, m, a, k};
  emu_launch(dim3((unsigned)((B * C + 255) / 256)), 256, [&] { margin_fwd_kernel(cosm, label, B, C, ld, mg); });
}
extern "C" void emu_margin_bwd(float* dl, const float* cosm, const int64_t* label, int64_t B, int64_t C, int64_t ld, int kind, float s, float m,
                               float a, float k) {
  Margin mg{kind, s, m, a, k};
  emu_launch(dim3((unsigned)((B * C + 255) / 256)), 256, [&] { margin_bwd_kernel(dl, cosm, label, B, C, ld, mg); });
}
