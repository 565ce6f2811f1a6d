// K-D  PartialFC label remap + negative-class sampling: C-ABI launchers.  Kernels and algorithm: pfc_sample_kernels.cuh
//   ref headers/partial_fc.py:77-94
#include "pfc_sample_kernels.cuh"

using namespace msml;

extern "C" int msml_pfc_remap(int64_t* tl, int64_t n, int64_t class_start, int64_t num_local, void* stream) {
  MSML_REQUIRE(tl && n > 0 && num_local > 0 && class_start >= 0, MSML_EINVAL, "bad remap arguments");
  MSML_PROF("pfc_remap", (double)n * 16, (cudaStream_t)stream);
  pfc_remap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(tl, n, class_start, num_local);
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" int msml_pfc_mark_positive(float* perm, const int64_t* tl, int64_t n, int64_t num_local, void* stream) {
  MSML_REQUIRE(perm && tl && n > 0 && num_local > 0, MSML_EINVAL, "bad mark_positive arguments");
  MSML_PROF("pfc_mark_positive", (double)n * 12, (cudaStream_t)stream);
  pfc_mark_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(perm, tl, n, num_local);
  MSML_LAUNCH_CHECK();
  return 0;
}

static int64_t sel_blocks(int64_t num_local) { return (num_local + kSelTile - 1) / kSelTile; }

extern "C" size_t msml_pfc_select_workspace(int64_t num_local) {
  if (num_local <= 0) return 0;
  return ((sizeof(SelState) + 255) / 256) * 256 + 2 * (size_t)sel_blocks(num_local) * sizeof(long long);
}

extern "C" int msml_pfc_select(const float* perm, int64_t num_local, int64_t num_sample, int64_t* index,
                               int64_t* n_index, void* workspace, size_t workspace_bytes, void* stream) {
  MSML_REQUIRE(perm && index && n_index && workspace, MSML_EINVAL, "null pointer");
  MSML_REQUIRE(num_local > 0 && num_sample >= 0 && num_sample <= num_local, MSML_EINVAL,
               "bad select sizes num_local=%lld num_sample=%lld", (long long)num_local, (long long)num_sample);
  MSML_REQUIRE(workspace_bytes >= msml_pfc_select_workspace(num_local), MSML_EWORKSPACE,
               "select workspace too small: %zu < %zu", workspace_bytes, msml_pfc_select_workspace(num_local));
  MSML_REQUIRE(aligned16(workspace), MSML_EALIGN, "workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  SelState* state = static_cast<SelState*>(workspace);
  long long* blk_gt = reinterpret_cast<long long*>(static_cast<char*>(workspace) + ((sizeof(SelState) + 255) / 256) * 256);
  const unsigned nb = (unsigned)sel_blocks(num_local);
  long long* blk_eq = blk_gt + nb;
  // one scope over the memset + 5 launches: 5 passes over perm (3 digit histograms, count, ordered compaction) + the index
  MSML_PROF("pfc_select", (double)num_local * 4 * 5 + (double)num_sample * 8, st);
  MSML_CUDA(cudaMemsetAsync(state, 0, sizeof(SelState), st));
  pfc_hist_kernel<0><<<nb, kSelThreads, 0, st>>>(perm, num_local, num_sample, state);
  MSML_LAUNCH_CHECK();
  pfc_hist_kernel<1><<<nb, kSelThreads, 0, st>>>(perm, num_local, num_sample, state);
  MSML_LAUNCH_CHECK();
  pfc_hist_kernel<2><<<nb, kSelThreads, 0, st>>>(perm, num_local, num_sample, state);
  MSML_LAUNCH_CHECK();
  pfc_count_kernel<<<nb, kSelThreads, 0, st>>>(perm, num_local, state, blk_gt, blk_eq);
  MSML_LAUNCH_CHECK();
  pfc_write_kernel<<<nb, kSelThreads, 0, st>>>(perm, num_local, state, blk_gt, blk_eq, index, n_index);
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" int msml_pfc_searchsorted(int64_t* tl, int64_t n, const int64_t* index, const int64_t* n_index, void* stream) {
  MSML_REQUIRE(tl && index && n_index && n > 0, MSML_EINVAL, "bad searchsorted arguments");
  MSML_PROF("pfc_searchsorted", (double)n * 16, (cudaStream_t)stream);
  pfc_searchsorted_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(tl, n, index, n_index);
  MSML_LAUNCH_CHECK();
  return 0;
}

static int rows_check(const void* a, const void* b, const void* c, int64_t n_rows, int64_t d) {
  MSML_REQUIRE(a && b && c, MSML_EINVAL, "null pointer");
  MSML_REQUIRE(n_rows > 0 && d > 0 && d % 4 == 0, MSML_EINVAL, "bad row copy shape n_rows=%lld d=%lld (d %% 4 != 0?)",
               (long long)n_rows, (long long)d);
  MSML_REQUIRE(aligned16(a) && aligned16(c), MSML_EALIGN, "row buffers must be 16-byte aligned");
  return 0;
}

extern "C" int msml_gather_rows_f32(const float* src, const int64_t* index, float* dst, int64_t n_rows, int64_t d, void* stream) {
  if (int e = rows_check(src, index, dst, n_rows, d)) return e;
  MSML_PROF("pfc_gather_rows", (double)n_rows * (d * 8 + 8), (cudaStream_t)stream);
  rows_copy_kernel<false><<<(unsigned)((n_rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(src, index, dst, n_rows, d / 4);
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" int msml_scatter_rows_f32(float* dst, const int64_t* index, const float* src, int64_t n_rows, int64_t d, void* stream) {
  if (int e = rows_check(src, index, dst, n_rows, d)) return e;
  MSML_PROF("pfc_scatter_rows", (double)n_rows * (d * 8 + 8), (cudaStream_t)stream);
  rows_copy_kernel<true><<<(unsigned)((n_rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(src, index, dst, n_rows, d / 4);
  MSML_LAUNCH_CHECK();
  return 0;
}
