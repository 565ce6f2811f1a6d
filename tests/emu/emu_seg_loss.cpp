// Host build of csrc/seg_loss_kernels.cuh under the CPU emulation (tests/emu/cuda_emu.h).  TEST INFRASTRUCTURE ONLY.
// Mirrors the launch sequence of msml_consensus_fwd / msml_consensus_bwd (csrc/seg_loss.cu) on host memory.
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>

#include "../../msml_b200/csrc/seg_loss_kernels.cuh"

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

template <typename T, int C>
static void run_fwd(const void* logit, const int64_t* blobs, const int64_t* target, float alpha, float beta, int pixel_all, int kl_all,
                    float* loss, float* coef, float* ws, const SegGeom& g) {
  float* partf = ws;
  float* acc = partf + seg_partf(g);
  int* parti = reinterpret_cast<int*>(acc + seg_acc(g));
  int* partbad = parti + seg_parti(g);
  emu_launch(dim3((unsigned)g.chunks, (unsigned)g.N), kSegThreads,
             [&] { seg_stats_kernel<T, C>(static_cast<const T*>(logit), blobs, target, partf, parti, partbad, g); });
  emu_launch(dim3(1), kSegThreads,
             [&] { seg_finalize_kernel<C>(partf, parti, partbad, target, acc, coef, loss, alpha, beta, pixel_all, kl_all, g); });
}

template <typename T, int C>
static void run_bwd(const void* logit, const int64_t* blobs, const float* coef, const float* gout, void* dlogit, const SegGeom& g) {
  emu_launch(dim3((unsigned)g.chunks, (unsigned)g.N), kSegThreads,
             [&] { seg_bwd_kernel<T, C>(static_cast<const T*>(logit), blobs, coef, gout, static_cast<T*>(dlogit), g); });
}

#define DISPATCH(dtype, Cv, CALL)                                                                   \
  if (dtype == MSML_F32) { using T = float; if (Cv == 2) { constexpr int C = 2; CALL; } else if (Cv == 3) { constexpr int C = 3; CALL; } else { constexpr int C = 4; CALL; } } \
  else { using T = __nv_bfloat16; if (Cv == 2) { constexpr int C = 2; CALL; } else if (Cv == 3) { constexpr int C = 3; CALL; } else { constexpr int C = 4; CALL; } }

extern "C" size_t emu_consensus_workspace(int64_t N, int64_t C, int64_t HW, int64_t K) {
  SegGeom g;
  if (seg_geom(N, C, HW, K, 0, MSML_F32, &g)) return 0;
  return (seg_partf(g) + seg_acc(g) + seg_parti(g) + seg_partbad(g)) * 4;
}

extern "C" int emu_consensus_fwd(const void* logit, const int64_t* blobs, const int64_t* target, int64_t N, int64_t C, int64_t HW, int64_t K,
                                 int channels_last, int dtype, float alpha, float beta, int pixel_all, int kl_all, float* loss, float* coef,
                                 void* ws) {
  SegGeom g;
  if (int e = seg_geom(N, C, HW, K, channels_last, dtype, &g)) return e;
  DISPATCH(dtype, g.C, (run_fwd<T, C>(logit, blobs, target, alpha, beta, pixel_all, kl_all, loss, coef, static_cast<float*>(ws), g)));
  return 0;
}

extern "C" int emu_consensus_bwd(const void* logit, const int64_t* blobs, const float* coef, const float* gout, void* dlogit, int64_t N,
                                 int64_t C, int64_t HW, int64_t K, int channels_last, int dtype) {
  SegGeom g;
  if (int e = seg_geom(N, C, HW, K, channels_last, dtype, &g)) return e;
  DISPATCH(dtype, g.C, (run_bwd<T, C>(logit, blobs, coef, gout, dlogit, g)));
  return 0;
}

extern "C" const char* emu_last_error() { return msml::g_err; }
