// Host build of csrc/fm_gate_kernels.cuh (K-A, the mask-fusion tail) under the CPU emulation.  TEST INFRASTRUCTURE ONLY.
// Mirrors msml_fm_gate_fwd[_multi] / msml_fm_gate_bwd[_multi] (csrc/fm_gate.cu): deal the CTAs to the segments, one launch.
// The kernels are verified on a B200; here the north-star kernel also runs in the CPU test tier and under the sanitizers.
#define MSML_CPU_EMU 1
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>

#include "../../msml_b200/csrc/fm_gate_kernels.cuh"

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

#define GATE_DISPATCH(dtype, act, arith, CALL)                                                         \
  do {                                                                                                 \
    const int _k = (act) * 4 + (arith);                                                                \
    if (dtype == MSML_F32) { using T = float;                                                          \
      switch (_k) { case 0: { constexpr int ACT = 0, ARITH = 0; CALL; } break; case 1: { constexpr int ACT = 0, ARITH = 1; CALL; } break; \
                    case 2: { constexpr int ACT = 0, ARITH = 2; CALL; } break; case 3: { constexpr int ACT = 0, ARITH = 3; CALL; } break; \
                    case 4: { constexpr int ACT = 1, ARITH = 0; CALL; } break; case 5: { constexpr int ACT = 1, ARITH = 1; CALL; } break; \
                    case 6: { constexpr int ACT = 1, ARITH = 2; CALL; } break; default: { constexpr int ACT = 1, ARITH = 3; CALL; } break; } } \
    else { using T = __nv_bfloat16;                                                                    \
      switch (_k) { case 0: { constexpr int ACT = 0, ARITH = 0; CALL; } break; case 1: { constexpr int ACT = 0, ARITH = 1; CALL; } break; \
                    case 2: { constexpr int ACT = 0, ARITH = 2; CALL; } break; case 3: { constexpr int ACT = 0, ARITH = 3; CALL; } break; \
                    case 4: { constexpr int ACT = 1, ARITH = 0; CALL; } break; case 5: { constexpr int ACT = 1, ARITH = 1; CALL; } break; \
                    case 6: { constexpr int ACT = 1, ARITH = 2; CALL; } break; default: { constexpr int ACT = 1, ARITH = 3; CALL; } break; } } \
  } while (0)

template <typename T, int ACT, int ARITH>
static void run_fwd(const FwdSegs& s, bool has_fout) {
  const int grid = s.block_end[MSML_MAX_SEGMENTS - 1];
  if (has_fout) emu_launch(dim3(grid), kThreads, [&] { fm_gate_fwd_kernel<T, ACT, ARITH, true>(s); });
  else emu_launch(dim3(grid), kThreads, [&] { fm_gate_fwd_kernel<T, ACT, ARITH, false>(s); });
}
template <typename T, int ACT, int ARITH>
static void run_bwd(const BwdSegs& s) {
  emu_launch(dim3(s.block_end[MSML_MAX_SEGMENTS - 1]), kThreads, [&] { fm_gate_bwd_kernel<T, ACT, ARITH>(s); });
}

#ifndef MSML_EMU_TEMPLATES_ONLY   // sanitize_main.cpp instantiates a few variants itself instead of all of them
// `sms` plays the SM count: small values force several trips of the persistent loops and uneven CTA dealing
extern "C" int emu_fm_gate_fwd_multi(int nseg, const void* const* yf, const void* const* z, const void* const* f_out, void* const* out,
                                     const int64_t* n, int dtype, int act, int arith, int sms) {
  if (nseg < 1 || nseg > MSML_MAX_SEGMENTS) return 1;
  FwdSegs s{};
  s.nseg = nseg;
  bool has_fout = f_out && f_out[0];
  for (int i = 0; i < nseg; ++i) { s.yf[i] = yf[i]; s.z[i] = z[i]; s.out[i] = out[i]; s.f_out[i] = has_fout ? f_out[i] : nullptr; s.n[i] = n[i]; }
  deal_blocks(nseg, n, dtype == MSML_F32 ? 4 : 8, has_fout ? kUnroll3 : kUnroll, kCtasPerSm, s.block_end, sms);
  GATE_DISPATCH(dtype, act, arith, (run_fwd<T, ACT, ARITH>(s, has_fout)));
  return 0;
}

extern "C" int emu_fm_gate_bwd_multi(int nseg, const void* const* dout, const void* const* yf, const void* const* z, void* const* dyf,
                                     void* const* dz, const int64_t* n, int dtype, int act, int arith, int sms) {
  if (nseg < 1 || nseg > MSML_MAX_SEGMENTS) return 1;
  BwdSegs s{};
  s.nseg = nseg;
  for (int i = 0; i < nseg; ++i) { s.dout[i] = dout[i]; s.yf[i] = yf[i]; s.z[i] = z[i]; s.dyf[i] = dyf[i]; s.dz[i] = dz[i]; s.n[i] = n[i]; }
  deal_blocks(nseg, n, dtype == MSML_F32 ? 4 : 8, kUnroll3, kCtasPerSmBwd, s.block_end, sms);
  GATE_DISPATCH(dtype, act, arith, (run_bwd<T, ACT, ARITH>(s)));
  return 0;
}
#endif
