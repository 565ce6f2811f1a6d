// K-S  structure-via-consensus segmentation criterion: C-ABI launchers.  Kernels and the algorithm: seg_loss_kernels.cuh
//   ref tricks/consensus_loss.py:63-178; SURVEY.md 8f-4.
// The kernels live in a header of their own so that tests/emu can compile them for the host and execute them under a
// CPU emulation of the CUDA execution model (logic check without a GPU).
#include "seg_loss_kernels.cuh"

namespace msml {

#define MSML_SEG_DISPATCH_C(Cv, ...)                                                  \
  switch (Cv) {                                                                       \
    case 2: { constexpr int C = 2; __VA_ARGS__; } break;                              \
    case 3: { constexpr int C = 3; __VA_ARGS__; } break;                              \
    default: { constexpr int C = 4; __VA_ARGS__; } break;                             \
  }

}  // namespace msml

using namespace msml;

extern "C" size_t msml_consensus_workspace(int64_t N, int64_t C, int64_t HW, int64_t K) {
  SegGeom g;
  if (seg_geom(N, C, HW, K, 0, MSML_F32, &g)) return 0;
  return (seg_partf(g) + seg_acc(g) + seg_parti(g) + seg_partbad(g)) * 4;
}

extern "C" int msml_consensus_fwd(const void* logit, const int64_t* blobs, const int64_t* target, int64_t N, int64_t C, int64_t HW,
                                  int64_t K, int channels_last, int dtype, float alpha, float beta, int reduce_pixel_all,
                                  int reduce_kl_all, float* loss, float* coef, void* ws, size_t ws_bytes, void* stream) {
  SegGeom g;
  if (int e = seg_geom(N, C, HW, K, channels_last, dtype, &g)) return e;
  MSML_REQUIRE(logit && blobs && target && loss && coef, MSML_EINVAL, "null pointer");
  MSML_REQUIRE(ws && ws_bytes >= msml_consensus_workspace(N, C, HW, K), MSML_EWORKSPACE, "consensus workspace too small");
  MSML_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 3u) == 0, MSML_EALIGN, "workspace must be 4-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  float* partf = static_cast<float*>(ws);
  float* acc = partf + seg_partf(g);
  int* parti = reinterpret_cast<int*>(acc + seg_acc(g));
  int* partbad = parti + seg_parti(g);
  const double elem = dtype == MSML_F32 ? 4.0 : 2.0;
  MSML_PROF("consensus_fwd", (double)N * HW * (C * elem + 16.0), st);
  const dim3 grid((unsigned)g.chunks, (unsigned)N);
  MSML_DISPATCH_DTYPE(dtype, T, MSML_SEG_DISPATCH_C(g.C, (seg_stats_kernel<T, C><<<grid, kSegThreads, 0, st>>>(
                                                             static_cast<const T*>(logit), blobs, target, partf, parti, partbad, g))));
  MSML_LAUNCH_CHECK();
  MSML_SEG_DISPATCH_C(g.C, (seg_finalize_kernel<C><<<1, kSegThreads, 0, st>>>(partf, parti, partbad, target, acc, coef, loss, alpha, beta,
                                                                             reduce_pixel_all, reduce_kl_all, g)));
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" int msml_consensus_bwd(const void* logit, const int64_t* blobs, const float* coef, const float* gout, void* dlogit,
                                  int64_t N, int64_t C, int64_t HW, int64_t K, int channels_last, int dtype, void* stream) {
  SegGeom g;
  if (int e = seg_geom(N, C, HW, K, channels_last, dtype, &g)) return e;
  MSML_REQUIRE(logit && blobs && coef && dlogit, MSML_EINVAL, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const double elem = dtype == MSML_F32 ? 4.0 : 2.0;
  MSML_PROF("consensus_bwd", (double)N * HW * (2 * C * elem + 8.0), st);
  const dim3 grid((unsigned)g.chunks, (unsigned)N);
  MSML_DISPATCH_DTYPE(dtype, T, MSML_SEG_DISPATCH_C(g.C, (seg_bwd_kernel<T, C><<<grid, kSegThreads, 0, st>>>(
                                                             static_cast<const T*>(logit), blobs, coef, gout, static_cast<T*>(dlogit), g))));
  MSML_LAUNCH_CHECK();
  return 0;
}
