// K-A  mask fusion: C-ABI launchers.  Kernels, CTA dealing and the algorithm: fm_gate_kernels.cuh
//   ref backbones/fm/fmoperator.py:288 (act), :304 (arith), :307-308 (+ f_out), :310 (+ identity)
#include "fm_gate_kernels.cuh"

namespace msml {

template <typename T, int ACT, int ARITH>
static int launch_fwd(const FwdSegs& s, bool has_fout, cudaStream_t st) {
  const int grid = s.block_end[MSML_MAX_SEGMENTS - 1];
  double bytes = 0;
  for (int i = 0; i < s.nseg; ++i) bytes += (double)s.n[i] * sizeof(T) * (has_fout ? 4 : 3);
  MSML_PROF("fm_gate_fwd", bytes, st);
  if (has_fout) fm_gate_fwd_kernel<T, ACT, ARITH, true><<<grid, kThreads, 0, st>>>(s);
  else fm_gate_fwd_kernel<T, ACT, ARITH, false><<<grid, kThreads, 0, st>>>(s);
  MSML_LAUNCH_CHECK();
  return 0;
}
template <typename T, int ACT, int ARITH>
static int launch_bwd(const BwdSegs& s, cudaStream_t st) {
  const int grid = s.block_end[MSML_MAX_SEGMENTS - 1];
  double bytes = 0;
  for (int i = 0; i < s.nseg; ++i) bytes += (double)s.n[i] * sizeof(T) * 5;
  MSML_PROF("fm_gate_bwd", bytes, st);
  fm_gate_bwd_kernel<T, ACT, ARITH><<<grid, kThreads, 0, st>>>(s);
  MSML_LAUNCH_CHECK();
  return 0;
}

#define MSML_DISPATCH_ACT_ARITH(act, arith, ACT, ARITH, ...)                                        \
  do {                                                                                              \
    const int _k = (act) * 4 + (arith);                                                             \
    switch (_k) {                                                                                   \
      case 0: { constexpr int ACT = 0, ARITH = 0; __VA_ARGS__; } break;                             \
      case 1: { constexpr int ACT = 0, ARITH = 1; __VA_ARGS__; } break;                             \
      case 2: { constexpr int ACT = 0, ARITH = 2; __VA_ARGS__; } break;                             \
      case 3: { constexpr int ACT = 0, ARITH = 3; __VA_ARGS__; } break;                             \
      case 4: { constexpr int ACT = 1, ARITH = 0; __VA_ARGS__; } break;                             \
      case 5: { constexpr int ACT = 1, ARITH = 1; __VA_ARGS__; } break;                             \
      case 6: { constexpr int ACT = 1, ARITH = 2; __VA_ARGS__; } break;                             \
      case 7: { constexpr int ACT = 1, ARITH = 3; __VA_ARGS__; } break;                             \
      default: return ::msml::set_error(MSML_EINVAL, "bad act/arith %d/%d", (int)(act), (int)(arith)); \
    }                                                                                               \
  } while (0)

static int check_common(int nseg, const int64_t* n, int dtype, int act, int arith) {
  MSML_REQUIRE(nseg >= 1 && nseg <= MSML_MAX_SEGMENTS, MSML_EINVAL, "nseg %d out of [1,%d]", nseg, MSML_MAX_SEGMENTS);
  MSML_REQUIRE(dtype >= 0 && dtype <= 2, MSML_EINVAL, "unknown dtype %d", dtype);
  MSML_REQUIRE(act == MSML_ACT_TANH || act == MSML_ACT_SIGMOID, MSML_EINVAL, "activation type error (%d)", act);
  MSML_REQUIRE(arith >= 0 && arith <= 3, MSML_EINVAL, "arith type error (%d)", arith);
  for (int i = 0; i < nseg; ++i) MSML_REQUIRE(n[i] > 0, MSML_EINVAL, "segment %d has n=%lld", i, (long long)n[i]);
  return 0;
}

}  // namespace msml

using namespace msml;

extern "C" int msml_fm_gate_fwd_multi(int nseg, const void* const* yf, const void* const* z,
                                      void* const* out, const int64_t* n, int dtype, int act, int arith,
                                      void* stream) {
  if (int e = check_common(nseg, n, dtype, act, arith)) return e;
  FwdSegs s{};
  s.nseg = nseg;
  for (int i = 0; i < nseg; ++i) {
    MSML_REQUIRE(yf[i] && z[i] && out[i], MSML_EINVAL, "null pointer in segment %d", i);
    MSML_REQUIRE(aligned16(yf[i]) && aligned16(z[i]) && aligned16(out[i]), MSML_EALIGN,
                 "segment %d: pointers must be 16-byte aligned", i);
    s.yf[i] = yf[i]; s.z[i] = z[i]; s.out[i] = out[i]; s.f_out[i] = nullptr; s.n[i] = n[i];
  }
  deal_blocks(nseg, n, dtype == MSML_F32 ? 4 : 8, kUnroll, kCtasPerSm, s.block_end, num_sms());
  MSML_DISPATCH_DTYPE(dtype, T, MSML_DISPATCH_ACT_ARITH(act, arith, ACT, ARITH,
      return (launch_fwd<T, ACT, ARITH>(s, false, (cudaStream_t)stream))));
  return 0;
}

extern "C" int msml_fm_gate_fwd(const void* yf, const void* z, const void* f_out, void* out, int64_t n,
                                int dtype, int act, int arith, void* stream) {
  if (int e = check_common(1, &n, dtype, act, arith)) return e;
  MSML_REQUIRE(yf && z && out, MSML_EINVAL, "null pointer");
  MSML_REQUIRE(aligned16(yf) && aligned16(z) && aligned16(out) && aligned16(f_out), MSML_EALIGN,
               "pointers must be 16-byte aligned");
  FwdSegs s{};
  s.nseg = 1;
  s.yf[0] = yf; s.z[0] = z; s.f_out[0] = f_out; s.out[0] = out; s.n[0] = n;
  deal_blocks(1, &n, dtype == MSML_F32 ? 4 : 8, f_out ? kUnroll3 : kUnroll, kCtasPerSm, s.block_end, num_sms());
  MSML_DISPATCH_DTYPE(dtype, T, MSML_DISPATCH_ACT_ARITH(act, arith, ACT, ARITH,
      return (launch_fwd<T, ACT, ARITH>(s, f_out != nullptr, (cudaStream_t)stream))));
  return 0;
}

extern "C" int msml_fm_gate_bwd_multi(int nseg, const void* const* dout, const void* const* yf,
                                      const void* const* z, void* const* dyf, void* const* dz,
                                      const int64_t* n, int dtype, int act, int arith, void* stream) {
  if (int e = check_common(nseg, n, dtype, act, arith)) return e;
  BwdSegs s{};
  s.nseg = nseg;
  for (int i = 0; i < nseg; ++i) {
    MSML_REQUIRE(dout[i] && yf[i] && z[i] && dyf[i] && dz[i], MSML_EINVAL, "null pointer in segment %d", i);
    MSML_REQUIRE(aligned16(dout[i]) && aligned16(yf[i]) && aligned16(z[i]) && aligned16(dyf[i]) && aligned16(dz[i]),
                 MSML_EALIGN, "segment %d: pointers must be 16-byte aligned", i);
    s.dout[i] = dout[i]; s.yf[i] = yf[i]; s.z[i] = z[i]; s.dyf[i] = dyf[i]; s.dz[i] = dz[i]; s.n[i] = n[i];
  }
  deal_blocks(nseg, n, dtype == MSML_F32 ? 4 : 8, kUnroll3, kCtasPerSmBwd, s.block_end, num_sms());
  MSML_DISPATCH_DTYPE(dtype, T, MSML_DISPATCH_ACT_ARITH(act, arith, ACT, ARITH,
      return (launch_bwd<T, ACT, ARITH>(s, (cudaStream_t)stream))));
  return 0;
}

extern "C" int msml_fm_gate_bwd(const void* dout, const void* yf, const void* z, void* dyf, void* dz,
                                int64_t n, int dtype, int act, int arith, void* stream) {
  const void* d[1] = {dout};
  const void* y[1] = {yf};
  const void* zz[1] = {z};
  void* o1[1] = {dyf};
  void* o2[1] = {dz};
  return msml_fm_gate_bwd_multi(1, d, y, zz, o1, o2, &n, dtype, act, arith, stream);
}
