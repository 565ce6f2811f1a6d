// K-B  DAP head of the occlusion-segmentation branch (+ fused 2-way argmax mask).
//   ref backbones/osb/unet.py:158-161,223: PixelShuffle(k) -> AvgPool2d(k)  ==  mean over the k*k
//   consecutive input channels of each output channel (the (B,G,kH,kW) intermediate the reference
//   materialises never exists here);  train.py:357: final_seg[b].max(0)[1].
// One thread per pixel (per pixel pair for 16-bit NCHW): reads G*kk values, writes G (+1 int64).
// The mean and the argmax are taken on the fp32 accumulator, before rounding to the output dtype.
#include "common.cuh"

namespace msml {

template <typename T, bool CL, bool MASK>
__global__ void __launch_bounds__(256)
dap_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t* __restrict__ mask,
               int64_t B, int G, int kk, int64_t HW) {
  const int64_t total = B * HW;
  const float inv = 1.0f / (float)kk;
  (void)inv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / HW, p = i - b * HW;
    float best = 0.f;
    int64_t arg = 0;
    for (int g = 0; g < G; ++g) {
      float acc = 0.f;
      for (int j = 0; j < kk; ++j) {
        const int64_t c = (int64_t)g * kk + j;
        const int64_t idx = CL ? (i * G * kk + c) : ((b * G * kk + c) * HW + p);
        acc += to_f32(x[idx]);
      }
      acc = acc / (float)kk;   // ATen avg_pool2d: sum / divide_factor
      const int64_t oidx = CL ? (i * G + g) : ((b * G + g) * HW + p);
      y[oidx] = from_f32<T>(acc);
      if (MASK && (g == 0 || acc > best)) { best = acc; arg = g; }   // strict '>' : first index on ties
    }
    if (MASK) mask[i] = arg;
  }
}

template <typename T, bool CL>
__global__ void __launch_bounds__(256)
dap_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int64_t B, int G, int kk, int64_t HW) {
  const int64_t total = B * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / HW, p = i - b * HW;
    for (int g = 0; g < G; ++g) {
      const int64_t oidx = CL ? (i * G + g) : ((b * G + g) * HW + p);
      const T v = from_f32<T>(to_f32(dy[oidx]) / (float)kk);
      for (int j = 0; j < kk; ++j) {
        const int64_t c = (int64_t)g * kk + j;
        const int64_t idx = CL ? (i * G * kk + c) : ((b * G * kk + c) * HW + p);
        dx[idx] = v;
      }
    }
  }
}

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
