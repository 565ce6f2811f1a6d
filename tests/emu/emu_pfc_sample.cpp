// Host build of csrc/pfc_sample_kernels.cuh (K-D: label remap, positive marking, exact radix select with ordered compaction,
// searchsorted, row gather / scatter) under the CPU emulation.  TEST INFRASTRUCTURE ONLY.
// Mirrors the launch sequences of msml_pfc_* / msml_{gather,scatter}_rows_f32 (csrc/pfc_sample.cu) on host memory, so the
// bit-exact integer path of PartialFC.sample runs in the CPU test tier too (the kernels are verified on a B200).
#define MSML_CPU_EMU 1
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../msml_b200/csrc/pfc_sample_kernels.cuh"

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

extern "C" void emu_pfc_remap(int64_t* tl, int64_t n, int64_t class_start, int64_t num_local) {
  emu_launch(dim3((unsigned)((n + 255) / 256)), 256, [&] { pfc_remap_kernel(tl, n, class_start, num_local); });
}
extern "C" void emu_pfc_mark_positive(float* perm, const int64_t* tl, int64_t n, int64_t num_local) {
  emu_launch(dim3((unsigned)((n + 255) / 256)), 256, [&] { pfc_mark_kernel(perm, tl, n, num_local); });
}
extern "C" void emu_pfc_select(const float* perm, int64_t num_local, int64_t num_sample, int64_t* index, int64_t* n_index) {
  const unsigned nb = (unsigned)((num_local + kSelTile - 1) / kSelTile);
  std::vector<SelState> state(1);
  std::memset(state.data(), 0, sizeof(SelState));
  std::vector<long long> blk_gt(nb), blk_eq(nb);
  SelState* st = state.data();
  emu_launch(dim3(nb), kSelThreads, [&] { pfc_hist_kernel<0>(perm, num_local, num_sample, st); });
  emu_launch(dim3(nb), kSelThreads, [&] { pfc_hist_kernel<1>(perm, num_local, num_sample, st); });
  emu_launch(dim3(nb), kSelThreads, [&] { pfc_hist_kernel<2>(perm, num_local, num_sample, st); });
  emu_launch(dim3(nb), kSelThreads, [&] { pfc_count_kernel(perm, num_local, st, blk_gt.data(), blk_eq.data()); });
  emu_launch(dim3(nb), kSelThreads, [&] { pfc_write_kernel(perm, num_local, st, blk_gt.data(), blk_eq.data(), index, n_index); });
}
extern "C" void emu_pfc_searchsorted(int64_t* tl, int64_t n, const int64_t* index, const int64_t* n_index) {
  emu_launch(dim3((unsigned)((n + 255) / 256)), 256, [&] { pfc_searchsorted_kernel(tl, n, index, n_index); });
}
extern "C" void emu_gather_rows(const float* src, const int64_t* index, float* dst, int64_t n_rows, int64_t d) {
  emu_launch(dim3((unsigned)((n_rows + 7) / 8)), 256, [&] { rows_copy_kernel<false>(src, index, dst, n_rows, d / 4); });
}
extern "C" void emu_scatter_rows(float* dst, const int64_t* index, const float* src, int64_t n_rows, int64_t d) {
  emu_launch(dim3((unsigned)((n_rows + 7) / 8)), 256, [&] { rows_copy_kernel<true>(src, index, dst, n_rows, d / 4); });
}
