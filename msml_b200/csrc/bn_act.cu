// K-N  fused BatchNorm (+ residual add) (+ PReLU), NHWC, forward and backward.
//   SURVEY.md 8(f)-1 "next" row: the BN / PReLU chains of the iResNet unit
//   (ref backbones/frb/iresnet.py:56-67, backbones/osb/unet.py:80-91) and of the FM bottlenecks
//   (ref backbones/fm/fmoperator.py:52-68):      y = prelu( bn(x) [+ res] )
//
// In the reference these are 3-6 separate ATen kernels per layer (batch_norm statistics, transform,
// prelu, add; and five more in backward) and make up ~60 % of the training step on B200.  Here:
//   forward   stats  : one read of x   -> per-CTA (mean, M2) partials -> finalize (Chan merge in fp64,
//                                         running stats, scale/shift)             [training only]
//             apply  : read x [, res], write y = prelu(x*scale + shift [+ res])
//   backward  reduce : read dy, x [, res] -> per-CTA partials of (sum du, sum du*xhat, sum dy*u*[u<=0])
//                      -> finalize (dgamma, dbeta, dprelu, per-channel coefficients)
//             apply  : read dy, x [, res], write dx [, dres = du]
// All passes are HBM-bound streams over a (P = N*H*W) x C matrix with C contiguous: a thread owns
// one 16-byte channel vector for its whole life (grid strides are multiples of the row), so the
// per-channel coefficients live in registers; loads are 128-bit, several rows in flight per thread.
#include "common.cuh"

namespace msml {

constexpr int kBnThreads = 256;
constexpr int kBnMaxCtas = 148 * 4;

struct BnGeom {
  int64_t P;       // rows (N*H*W)
  int C;           // channels (contiguous)
  int vpr;         // 16-byte vectors per row
  int rows_per_pass;   // kBnThreads / vpr
  int grid;
};

// ------------------------------------------------------------------------------------------- forward
template <typename T>
__global__ void __launch_bounds__(kBnThreads)
bn_stats_partial_kernel(const T* __restrict__ x, BnGeom g, float* __restrict__ part /* [grid][2][C]: mean, M2 */,
                        float* __restrict__ part_n /* [grid] */) {
  constexpr int VN = Vec<T>::N;
  __shared__ float red[2][kBnThreads * 8 / 8 * 8];   // [2][rows_per_pass * C] <= [2][256*8]
  const int cv = threadIdx.x % g.vpr, rl = threadIdx.x / g.vpr;
  const int64_t rows_per_cta = (g.P + g.grid - 1) / g.grid;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  int64_t r1 = r0 + rows_per_cta;
  if (r1 > g.P) r1 = g.P;
  float s[VN], q[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) { s[i] = 0.f; q[i] = 0.f; }
  const uint4* xv = reinterpret_cast<const uint4*>(x) + cv;
  int64_t r = r0 + rl;
  // 4 rows in flight per thread
  for (; r + 3 * g.rows_per_pass < r1; r += 4 * g.rows_per_pass) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ld_stream(xv + (r + (int64_t)u * g.rows_per_pass) * g.vpr);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[VN];
      Vec<T>::unpack(v[u], f);
#pragma unroll
      for (int i = 0; i < VN; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
    }
  }
  for (; r < r1; r += g.rows_per_pass) {
    float f[VN];
    Vec<T>::unpack(ld_stream(xv + r * g.vpr), f);
#pragma unroll
    for (int i = 0; i < VN; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
  }
  // reduce over the row lanes of the CTA
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    red[0][rl * g.C + cv * VN + i] = s[i];
    red[1][rl * g.C + cv * VN + i] = q[i];
  }
  __syncthreads();
  const float n = (float)(r1 > r0 ? r1 - r0 : 0);
  for (int c = threadIdx.x; c < g.C; c += kBnThreads) {
    float ss = 0.f, qq = 0.f;
    for (int k = 0; k < g.rows_per_pass; ++k) { ss += red[0][k * g.C + c]; qq += red[1][k * g.C + c]; }
    const float mean = n > 0.f ? ss / n : 0.f;
    part[((int64_t)blockIdx.x * 2 + 0) * g.C + c] = mean;
    part[((int64_t)blockIdx.x * 2 + 1) * g.C + c] = fmaxf(qq - ss * mean, 0.f);   // M2 of this slab
  }
  if (threadIdx.x == 0) part_n[blockIdx.x] = n;
}

// Chan et al. parallel merge of the slab statistics; running stats; per-channel scale/shift.
// One warp per channel: lanes stride over the slabs, then a 5-step shuffle merge (all fp32: the merge of
// (n, mean, M2) triples is well conditioned).
__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float mb, float m2b) {
  if (nb <= 0.f) return;
  const float tot = n + nb, delta = mb - mean;
  const float f = nb / tot;
  mean = fmaf(delta, f, mean);
  m2 += m2b + delta * delta * n * f;
  n = tot;
}

__global__ void __launch_bounds__(256)
bn_stats_finalize_kernel(const float* __restrict__ part, const float* __restrict__ part_n, int grid, int C, float P,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ running_mean,
                         float* __restrict__ running_var, long long* __restrict__ nbt, float momentum, float eps,
                         float* __restrict__ save_mean, float* __restrict__ save_invstd, float* __restrict__ coef /* [2][C] */) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int b = lane; b < grid; b += 32)
    chan_merge(n, mean, m2, part_n[b], part[((int64_t)b * 2 + 0) * C + c], part[((int64_t)b * 2 + 1) * C + c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mean, o),
                m2b = __shfl_xor_sync(0xffffffffu, m2, o);
    chan_merge(n, mean, m2, nb, mb, m2b);
  }
  if (lane != 0) return;
  const float var = m2 / P;                       // biased: normalisation
  const float invstd = rsqrtf(var + eps);
  save_mean[c] = mean;
  save_invstd[c] = invstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
  if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (P > 1.f ? m2 / (P - 1.f) : var);
  const float sc = (gamma ? gamma[c] : 1.f) * invstd;
  coef[c] = sc;
  coef[C + c] = (beta ? beta[c] : 0.f) - mean * sc;
}

// eval mode: coefficients from the running statistics
__global__ void __launch_bounds__(128)
bn_eval_coef_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ running_mean,
                    const float* __restrict__ running_var, float eps, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                    float* __restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = rsqrtf(running_var[c] + eps);
  save_mean[c] = running_mean[c];
  save_invstd[c] = invstd;
  const float sc = (gamma ? gamma[c] : 1.f) * invstd;
  coef[c] = sc;
  coef[C + c] = (beta ? beta[c] : 0.f) - running_mean[c] * sc;
}

template <typename T, bool RES, bool PRELU>
__global__ void __launch_bounds__(kBnThreads)
bn_apply_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y, const float* __restrict__ coef,
                    const float* __restrict__ prelu, BnGeom g) {
  constexpr int VN = Vec<T>::N;
  const int cv = threadIdx.x % g.vpr;
  float sc[VN], sh[VN], pa[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    sc[i] = coef[cv * VN + i];
    sh[i] = coef[g.C + cv * VN + i];
    pa[i] = PRELU ? prelu[cv * VN + i] : 0.f;
  }
  const int64_t total = g.P * g.vpr;
  const int64_t stride = (int64_t)gridDim.x * kBnThreads;        // multiple of vpr: channel vector is loop-invariant
  constexpr int U = RES ? 2 : 4;
  for (int64_t base = (int64_t)blockIdx.x * kBnThreads + threadIdx.x; base < total; base += stride * U) {
    uint4 a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = base + u * stride;
      if (v < total) {
        a[u] = ld_stream(reinterpret_cast<const uint4*>(x) + v);
        if (RES) b[u] = ld_stream(reinterpret_cast<const uint4*>(res) + v);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = base + u * stride;
      if (v >= total) break;
      float f[VN], r[VN], o[VN];
      Vec<T>::unpack(a[u], f);
      if (RES) Vec<T>::unpack(b[u], r);
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        float t = fmaf(f[i], sc[i], sh[i]);
        if (RES) t += r[i];
        o[i] = PRELU ? (t > 0.f ? t : t * pa[i]) : t;
      }
      *(reinterpret_cast<uint4*>(y) + v) = Vec<T>::pack(o);      // y is re-read by the next conv: default policy
    }
  }
}

// ------------------------------------------------------------------------------------------- backward
// u = xhat*gamma + beta [+ res];  du = dy * (u > 0 ? 1 : a);  partial sums per slab:
//   [0] sum du   [1] sum du * xhat   [2] sum dy * u * [u <= 0]
template <typename T, bool RES, bool PRELU>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ res,
                     const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ prelu, BnGeom g,
                     float* __restrict__ part /* [grid][3][C] */) {
  constexpr int VN = Vec<T>::N;
  __shared__ float red[3][kBnThreads * 8];
  const int cv = threadIdx.x % g.vpr, rl = threadIdx.x / g.vpr;
  float mu[VN], is[VN], ga[VN], be[VN], pa[VN], s0[VN], s1[VN], s2[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    const int c = cv * VN + i;
    mu[i] = mean[c]; is[i] = invstd[c];
    ga[i] = gamma ? gamma[c] : 1.f; be[i] = beta ? beta[c] : 0.f;
    pa[i] = PRELU ? prelu[c] : 1.f;
    s0[i] = s1[i] = s2[i] = 0.f;
  }
  const int64_t rows_per_cta = (g.P + g.grid - 1) / g.grid;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  int64_t r1 = r0 + rows_per_cta;
  if (r1 > g.P) r1 = g.P;
  constexpr int U = 2;
  for (int64_t r = r0 + rl; r < r1; r += (int64_t)U * g.rows_per_pass) {
    uint4 a[U], b[U], c4[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + (int64_t)u * g.rows_per_pass;
      if (rr < r1) {
        const int64_t v = rr * g.vpr + cv;
        a[u] = ld_stream(reinterpret_cast<const uint4*>(dy) + v);
        b[u] = ld_stream(reinterpret_cast<const uint4*>(x) + v);
        if (RES && PRELU) c4[u] = ld_stream(reinterpret_cast<const uint4*>(res) + v);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + (int64_t)u * g.rows_per_pass;
      if (rr >= r1) break;
      float d[VN], f[VN], rs[VN];
      Vec<T>::unpack(a[u], d);
      Vec<T>::unpack(b[u], f);
      if (RES && PRELU) Vec<T>::unpack(c4[u], rs);
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        const float xh = (f[i] - mu[i]) * is[i];
        float du = d[i];
        if (PRELU) {
          float uu = fmaf(xh, ga[i], be[i]);
          if (RES) uu += rs[i];
          const bool neg = !(uu > 0.f);
          if (neg) { s2[i] = fmaf(d[i], uu, s2[i]); du *= pa[i]; }
        }
        s0[i] += du;
        s1[i] = fmaf(du, xh, s1[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    red[0][rl * g.C + cv * VN + i] = s0[i];
    red[1][rl * g.C + cv * VN + i] = s1[i];
    red[2][rl * g.C + cv * VN + i] = s2[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < g.C; c += kBnThreads) {
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int k = 0; k < g.rows_per_pass; ++k) { t0 += red[0][k * g.C + c]; t1 += red[1][k * g.C + c]; t2 += red[2][k * g.C + c]; }
    part[((int64_t)blockIdx.x * 3 + 0) * g.C + c] = t0;
    part[((int64_t)blockIdx.x * 3 + 1) * g.C + c] = t1;
    part[((int64_t)blockIdx.x * 3 + 2) * g.C + c] = t2;
  }
}

// dgamma, dbeta, dprelu and the coefficients of  dx = A * (du - B - xhat * G); one warp per channel
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const float* __restrict__ part, int grid, int C, float P, const float* __restrict__ gamma,
                       const float* __restrict__ invstd, int training, float* __restrict__ dgamma, float* __restrict__ dbeta,
                       float* __restrict__ dprelu, float* __restrict__ coef /* [3][C]: A, B, G */) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  float t0 = 0.f, t1 = 0.f, t2 = 0.f;
  for (int b = lane; b < grid; b += 32) {
    t0 += part[((int64_t)b * 3 + 0) * C + c];
    t1 += part[((int64_t)b * 3 + 1) * C + c];
    t2 += part[((int64_t)b * 3 + 2) * C + c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    t0 += __shfl_xor_sync(0xffffffffu, t0, o);
    t1 += __shfl_xor_sync(0xffffffffu, t1, o);
    t2 += __shfl_xor_sync(0xffffffffu, t2, o);
  }
  if (lane != 0) return;
  if (dbeta) dbeta[c] = t0;
  if (dgamma) dgamma[c] = t1;
  if (dprelu) dprelu[c] = t2;
  coef[c] = (gamma ? gamma[c] : 1.f) * invstd[c];
  coef[C + c] = training ? t0 / P : 0.f;      // eval: statistics are constants
  coef[2 * C + c] = training ? t1 / P : 0.f;
}

template <typename T, bool RES, bool PRELU>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ res,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ prelu, const float* __restrict__ coef,
                    T* __restrict__ dx, T* __restrict__ dres, BnGeom g) {
  constexpr int VN = Vec<T>::N;
  const int cv = threadIdx.x % g.vpr;
  float mu[VN], is[VN], ga[VN], be[VN], pa[VN], A[VN], B[VN], G[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    const int c = cv * VN + i;
    mu[i] = mean[c]; is[i] = invstd[c];
    ga[i] = gamma ? gamma[c] : 1.f; be[i] = beta ? beta[c] : 0.f;
    pa[i] = PRELU ? prelu[c] : 1.f;
    A[i] = coef[c]; B[i] = coef[g.C + c]; G[i] = coef[2 * g.C + c];
  }
  const int64_t total = g.P * g.vpr;
  const int64_t stride = (int64_t)gridDim.x * kBnThreads;
  constexpr int U = 2;
  for (int64_t base = (int64_t)blockIdx.x * kBnThreads + threadIdx.x; base < total; base += stride * U) {
    uint4 a[U], b[U], c4[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = base + u * stride;
      if (v < total) {
        a[u] = ld_stream(reinterpret_cast<const uint4*>(dy) + v);
        b[u] = ld_stream(reinterpret_cast<const uint4*>(x) + v);
        if (RES && PRELU) c4[u] = ld_stream(reinterpret_cast<const uint4*>(res) + v);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = base + u * stride;
      if (v >= total) break;
      float d[VN], f[VN], rs[VN], o[VN], dr[VN];
      Vec<T>::unpack(a[u], d);
      Vec<T>::unpack(b[u], f);
      if (RES && PRELU) Vec<T>::unpack(c4[u], rs);
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        const float xh = (f[i] - mu[i]) * is[i];
        float du = d[i];
        if (PRELU) {
          float uu = fmaf(xh, ga[i], be[i]);
          if (RES) uu += rs[i];
          if (!(uu > 0.f)) du *= pa[i];
        }
        dr[i] = du;
        o[i] = A[i] * (du - B[i] - xh * G[i]);
      }
      *(reinterpret_cast<uint4*>(dx) + v) = Vec<T>::pack(o);
      if (RES && PRELU) *(reinterpret_cast<uint4*>(dres) + v) = Vec<T>::pack(dr);
    }
  }
}

static int bn_geom(int64_t P, int64_t C, int dtype, BnGeom* g) {
  MSML_REQUIRE(P > 0 && C > 0, MSML_EINVAL, "bad BN shape P=%lld C=%lld", (long long)P, (long long)C);
  MSML_REQUIRE(dtype == MSML_F32 || dtype == MSML_BF16 || dtype == MSML_F16, MSML_EINVAL, "unknown dtype %d", dtype);
  const int vn = dtype == MSML_F32 ? 4 : 8;
  MSML_REQUIRE(C % vn == 0, MSML_EUNSUPPORTED, "C=%lld must be a multiple of %d", (long long)C, vn);
  const int vpr = (int)(C / vn);
  MSML_REQUIRE(vpr <= kBnThreads && kBnThreads % vpr == 0, MSML_EUNSUPPORTED,
               "C=%lld: vectors per row (%d) must divide %d", (long long)C, vpr, kBnThreads);
  g->P = P; g->C = (int)C; g->vpr = vpr; g->rows_per_pass = kBnThreads / vpr;
  // slabs of >= 8 passes per CTA, at most kBnMaxCtas CTAs
  int64_t want = (P + (int64_t)g->rows_per_pass * 8 - 1) / ((int64_t)g->rows_per_pass * 8);
  const int64_t cap = (int64_t)num_sms() * 4 < kBnMaxCtas ? (int64_t)num_sms() * 4 : kBnMaxCtas;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  g->grid = (int)want;
  return 0;
}

static size_t bn_ws_floats(int C) { return (size_t)kBnMaxCtas * 3 * C + kBnMaxCtas + 3 * (size_t)C; }

}  // namespace msml

using namespace msml;

extern "C" size_t msml_bn_workspace(int64_t P, int64_t C) {
  (void)P;
  return C > 0 ? bn_ws_floats((int)C) * sizeof(float) : 0;
}

#define MSML_BN_DISPATCH(dtype, has_res, has_prelu, ...)                                        \
  MSML_DISPATCH_DTYPE(dtype, T, {                                                               \
    if (has_res) { if (has_prelu) { constexpr bool RES = true, PRELU = true; __VA_ARGS__; }     \
                   else { constexpr bool RES = true, PRELU = false; __VA_ARGS__; } }            \
    else { if (has_prelu) { constexpr bool RES = false, PRELU = true; __VA_ARGS__; }            \
           else { constexpr bool RES = false, PRELU = false; __VA_ARGS__; } }                   \
  })

extern "C" int msml_bn_fwd(const void* x, const void* res, void* y, const float* gamma, const float* beta, const float* prelu,
                           float* running_mean, float* running_var, int64_t* num_batches_tracked, float* save_mean,
                           float* save_invstd, int64_t P, int64_t C, int dtype, int training, float momentum, float eps,
                           void* ws, size_t ws_bytes, void* stream) {
  BnGeom g;
  if (int e = bn_geom(P, C, dtype, &g)) return e;
  MSML_REQUIRE(x && y && save_mean && save_invstd && ws, MSML_EINVAL, "null pointer");
  MSML_REQUIRE(training || (running_mean && running_var), MSML_EINVAL, "eval mode needs running statistics");
  MSML_REQUIRE(aligned16(x) && aligned16(y) && aligned16(res) && aligned16(ws), MSML_EALIGN, "pointers must be 16-byte aligned");
  MSML_REQUIRE(ws_bytes >= msml_bn_workspace(P, C), MSML_EWORKSPACE, "BN workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* part = static_cast<float*>(ws);
  float* part_n = part + (size_t)kBnMaxCtas * 3 * C;
  float* coef = part_n + kBnMaxCtas;
  const double elem = dtype == MSML_F32 ? 4.0 : 2.0;
  if (training) {
    {
      MSML_PROF("bn_stats", (double)P * C * elem, st);
      MSML_DISPATCH_DTYPE(dtype, T, (bn_stats_partial_kernel<T><<<g.grid, kBnThreads, 0, st>>>(static_cast<const T*>(x), g, part, part_n)));
      MSML_LAUNCH_CHECK();
    }
    MSML_PROF("bn_stats_finalize", (double)g.grid * 2 * C * 4, st);
    bn_stats_finalize_kernel<<<(unsigned)((C + 7) / 8), 256, 0, st>>>(part, part_n, g.grid, (int)C, (float)P, gamma, beta,
                                                                        running_mean, running_var,
                                                                        reinterpret_cast<long long*>(num_batches_tracked),
                                                                        momentum, eps, save_mean, save_invstd, coef);
    MSML_LAUNCH_CHECK();
  } else {
    bn_eval_coef_kernel<<<(unsigned)((C + 127) / 128), 128, 0, st>>>((int)C, gamma, beta, running_mean, running_var, eps, save_mean,
                                                                   save_invstd, coef);
    MSML_LAUNCH_CHECK();
  }
  const int64_t total = P * g.vpr;
  int64_t blocks = (total + (int64_t)kBnThreads * 4 - 1) / ((int64_t)kBnThreads * 4);
  const int64_t cap = (int64_t)num_sms() * 4;
  const int grid = (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
  {
    MSML_PROF("bn_apply_fwd", (double)P * C * elem * (res ? 3 : 2), st);
    MSML_BN_DISPATCH(dtype, res != nullptr, prelu != nullptr,
                     (bn_apply_fwd_kernel<T, RES, PRELU><<<grid, kBnThreads, 0, st>>>(static_cast<const T*>(x), static_cast<const T*>(res),
                                                                                     static_cast<T*>(y), coef, prelu, g)));
    MSML_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int msml_bn_bwd(const void* dy, const void* x, const void* res, const float* gamma, const float* beta,
                           const float* prelu, const float* save_mean, const float* save_invstd, void* dx, void* dres,
                           float* dgamma, float* dbeta, float* dprelu, int64_t P, int64_t C, int dtype, int training, void* ws,
                           size_t ws_bytes, void* stream) {
  BnGeom g;
  if (int e = bn_geom(P, C, dtype, &g)) return e;
  MSML_REQUIRE(dy && x && dx && save_mean && save_invstd && ws, MSML_EINVAL, "null pointer");
  const bool has_prelu = prelu != nullptr, has_res = res != nullptr;
  MSML_REQUIRE(!(has_prelu && has_res) || dres, MSML_EINVAL, "dres is required when both a residual and PReLU are fused");
  MSML_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(res) && aligned16(dx) && aligned16(dres) && aligned16(ws), MSML_EALIGN,
               "pointers must be 16-byte aligned");
  MSML_REQUIRE(ws_bytes >= msml_bn_workspace(P, C), MSML_EWORKSPACE, "BN workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* part = static_cast<float*>(ws);
  float* coef = part + (size_t)kBnMaxCtas * 3 * C + kBnMaxCtas;
  const double elem = dtype == MSML_F32 ? 4.0 : 2.0;
  const int streams_in = 2 + (has_prelu && has_res ? 1 : 0);
  {
    MSML_PROF("bn_bwd_reduce", (double)P * C * elem * streams_in, st);
    MSML_BN_DISPATCH(dtype, has_res, has_prelu,
                     (bn_bwd_reduce_kernel<T, RES, PRELU><<<g.grid, kBnThreads, 0, st>>>(
                         static_cast<const T*>(dy), static_cast<const T*>(x), static_cast<const T*>(res), save_mean, save_invstd, gamma,
                         beta, prelu, g, part)));
    MSML_LAUNCH_CHECK();
  }
  {
    MSML_PROF("bn_bwd_finalize", (double)g.grid * 3 * C * 4, st);
    bn_bwd_finalize_kernel<<<(unsigned)((C + 7) / 8), 256, 0, st>>>(part, g.grid, (int)C, (float)P, gamma, save_invstd, training,
                                                                  dgamma, dbeta, has_prelu ? dprelu : nullptr, coef);
    MSML_LAUNCH_CHECK();
  }
  const int64_t total = P * g.vpr;
  int64_t blocks = (total + (int64_t)kBnThreads * 2 - 1) / ((int64_t)kBnThreads * 2);
  const int64_t cap = (int64_t)num_sms() * 4;
  const int grid = (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
  {
    MSML_PROF("bn_bwd_apply", (double)P * C * elem * (streams_in + 1 + (has_prelu && has_res ? 1 : 0)), st);
    MSML_BN_DISPATCH(dtype, has_res, has_prelu,
                     (bn_bwd_apply_kernel<T, RES, PRELU><<<grid, kBnThreads, 0, st>>>(
                         static_cast<const T*>(dy), static_cast<const T*>(x), static_cast<const T*>(res), save_mean, save_invstd, gamma,
                         beta, prelu, coef, static_cast<T*>(dx), static_cast<T*>(dres), g)));
    MSML_LAUNCH_CHECK();
  }
  return 0;
}
