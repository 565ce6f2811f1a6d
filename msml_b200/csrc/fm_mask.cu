// K-A (extension)  resized / broadcast mask fusion: C-ABI launchers.  Kernels and algorithm: fm_mask_kernels.cuh
#include "fm_mask_kernels.cuh"

namespace msml {

static int mask_check(const void* a, const void* b, const void* c, const MaskGeom& g, int dtype, int act, int arith) {
  MSML_REQUIRE(a && b && c, MSML_EINVAL, "null pointer");
  MSML_REQUIRE(g.B > 0 && g.H > 0 && g.W > 0 && g.C > 0 && g.Hm > 0 && g.Wm > 0, MSML_EINVAL, "bad mask-fusion shape");
  MSML_REQUIRE(g.Cm == 1 || g.Cm == g.C, MSML_EINVAL, "mask channels must be 1 or C (got %lld)", (long long)g.Cm);
  const int vn = dtype == MSML_F32 ? 4 : 8;
  const int64_t vpp = g.C / vn;
  MSML_REQUIRE(g.C % vn == 0 && (vpp & (vpp - 1)) == 0, MSML_EUNSUPPORTED,
               "C=%lld must be a power-of-two multiple of the 16-byte vector width", (long long)g.C);
  MSML_REQUIRE(aligned16(a) && aligned16(b) && aligned16(c), MSML_EALIGN, "pointers must be 16-byte aligned");
  MSML_REQUIRE(act == MSML_ACT_TANH || act == MSML_ACT_SIGMOID, MSML_EINVAL, "activation type error (%d)", act);
  MSML_REQUIRE(arith >= 0 && arith <= 3, MSML_EINVAL, "arith type error (%d)", arith);
  return 0;
}

}  // namespace msml

using namespace msml;

#define MSML_DISPATCH_AA(act, arith, ...)                                   \
  switch ((act) * 4 + (arith)) {                                            \
    case 0: { constexpr int ACT = 0, ARITH = 0; __VA_ARGS__; } break;       \
    case 1: { constexpr int ACT = 0, ARITH = 1; __VA_ARGS__; } break;       \
    case 2: { constexpr int ACT = 0, ARITH = 2; __VA_ARGS__; } break;       \
    case 3: { constexpr int ACT = 0, ARITH = 3; __VA_ARGS__; } break;       \
    case 4: { constexpr int ACT = 1, ARITH = 0; __VA_ARGS__; } break;       \
    case 5: { constexpr int ACT = 1, ARITH = 1; __VA_ARGS__; } break;       \
    case 6: { constexpr int ACT = 1, ARITH = 2; __VA_ARGS__; } break;       \
    default: { constexpr int ACT = 1, ARITH = 3; __VA_ARGS__; } break;      \
  }

extern "C" int msml_fm_mask_fwd(const void* yf, const void* m, void* out, int64_t B, int64_t H, int64_t W, int64_t C,
                                int64_t Hm, int64_t Wm, int64_t Cm, int dtype, int act, int arith, void* stream) {
  const MaskGeom g{B, H, W, C, Hm, Wm, Cm};
  if (int e = mask_check(yf, m, out, g, dtype, act, arith)) return e;
  const unsigned grid = (unsigned)((B * H * W + kPix - 1) / kPix);
  MSML_DISPATCH_DTYPE(dtype, T, MSML_DISPATCH_AA(act, arith,
      (fm_mask_fwd_kernel<T, ACT, ARITH><<<grid, kMaskThreads, 0, (cudaStream_t)stream>>>(
          static_cast<const T*>(yf), static_cast<const T*>(m), static_cast<T*>(out), g))));
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" int msml_fm_mask_bwd(const void* dout, const void* yf, const void* m, void* dyf, float* dm, int64_t B, int64_t H,
                                int64_t W, int64_t C, int64_t Hm, int64_t Wm, int64_t Cm, int dtype, int act, int arith,
                                void* stream) {
  const MaskGeom g{B, H, W, C, Hm, Wm, Cm};
  if (int e = mask_check(dout, yf, dyf, g, dtype, act, arith)) return e;
  MSML_REQUIRE(m && dm, MSML_EINVAL, "null pointer");
  const unsigned grid = (unsigned)((B * H * W + kPix - 1) / kPix);
  MSML_DISPATCH_DTYPE(dtype, T, MSML_DISPATCH_AA(act, arith,
      (fm_mask_bwd_kernel<T, ACT, ARITH><<<grid, kMaskThreads, 0, (cudaStream_t)stream>>>(
          static_cast<const T*>(dout), static_cast<const T*>(yf), static_cast<const T*>(m), static_cast<T*>(dyf), dm, g))));
  MSML_LAUNCH_CHECK();
  return 0;
}
