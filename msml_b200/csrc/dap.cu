// K-B  DAP head (+ fused argmax mask): C-ABI launchers.  Kernels and algorithm: dap_kernels.cuh
//   ref backbones/osb/unet.py:158-161,223; train.py:357
#include "dap_kernels.cuh"

namespace msml {

static int dap_grid(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace msml

using namespace msml;

static int dap_check(const void* a, const void* b, int64_t B, int64_t G, int64_t kk, int64_t H, int64_t W) {
  MSML_REQUIRE(a && b, MSML_EINVAL, "null pointer");
  MSML_REQUIRE(B > 0 && G > 0 && kk > 0 && H > 0 && W > 0, MSML_EINVAL, "bad DAP shape B=%lld G=%lld kk=%lld H=%lld W=%lld",
               (long long)B, (long long)G, (long long)kk, (long long)H, (long long)W);
  MSML_REQUIRE(G <= 1024 && kk <= 1024, MSML_EINVAL, "DAP group/kernel too large");
  return 0;
}

extern "C" int msml_dap_fwd(const void* x, void* y, int64_t* mask, int64_t B, int64_t G, int64_t kk, int64_t H,
                            int64_t W, int channels_last, int dtype, void* stream) {
  if (int e = dap_check(x, y, B, G, kk, H, W)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = dap_grid(B * H * W);
  MSML_PROF("dap_fwd", (double)B * H * W * (G * kk + G) * (dtype == MSML_F32 ? 4 : 2), st);
  MSML_DISPATCH_DTYPE(dtype, T, {
    const T* xi = static_cast<const T*>(x);
    T* yo = static_cast<T*>(y);
    if (channels_last) {
      if (mask) dap_fwd_kernel<T, true, true><<<grid, 256, 0, st>>>(xi, yo, mask, B, (int)G, (int)kk, H * W);
      else dap_fwd_kernel<T, true, false><<<grid, 256, 0, st>>>(xi, yo, mask, B, (int)G, (int)kk, H * W);
    } else {
      if (mask) dap_fwd_kernel<T, false, true><<<grid, 256, 0, st>>>(xi, yo, mask, B, (int)G, (int)kk, H * W);
      else dap_fwd_kernel<T, false, false><<<grid, 256, 0, st>>>(xi, yo, mask, B, (int)G, (int)kk, H * W);
    }
  });
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" int msml_dap_bwd(const void* dy, void* dx, int64_t B, int64_t G, int64_t kk, int64_t H, int64_t W,
                            int channels_last, int dtype, void* stream) {
  if (int e = dap_check(dy, dx, B, G, kk, H, W)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = dap_grid(B * H * W);
  MSML_PROF("dap_bwd", (double)B * H * W * (G * kk + G) * (dtype == MSML_F32 ? 4 : 2), st);
  MSML_DISPATCH_DTYPE(dtype, T, {
    if (channels_last) dap_bwd_kernel<T, true><<<grid, 256, 0, st>>>(static_cast<const T*>(dy), static_cast<T*>(dx), B, (int)G, (int)kk, H * W);
    else dap_bwd_kernel<T, false><<<grid, 256, 0, st>>>(static_cast<const T*>(dy), static_cast<T*>(dx), B, (int)G, (int)kk, H * W);
  });
  MSML_LAUNCH_CHECK();
  return 0;
}
