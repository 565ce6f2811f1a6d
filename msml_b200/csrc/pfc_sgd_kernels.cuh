// K-O  sharded-weight SGD for the PartialFC head: momentum SGD + weight decay on the sampled class centres, written
//      straight back into the shard, optionally emitting the next step's normalised bf16 centres.
//   ref train.py:188-191,299-300 (opt_pfc = SGD(momentum 0.9, weight_decay 5e-4); opt_pfc.step(); module_partial_fc.update())
//   ref headers/partial_fc.py:93-94 (gather of weight[index] / weight_mom[index]), :101-104 (scatter back), :112-115
//   (optimizer-state surgery, normalize(sub_weight)).
// In the reference one step moves every sampled row ten times: gather w and mom (2 reads + 2 writes), torch SGD
// (3 reads + 2 writes), scatter back (2 + 2), normalize (1 + 1).  Here one warp owns one row r of the sample:
//     row = index ? index[r] : r                     (sample_rate 1: the sample is the shard itself)
//     d   = dw[r] + wd * w[row]
//     m'  = mu * mom[row] + (1 - dampening) * d      (torch.optim.SGD with an existing momentum buffer: PartialFC always
//     w'  = w[row] - lr * (nesterov ? d + mu*m' : m')  supplies one, ref partial_fc.py:114)
//     w[row] = w', mom[row] = m'                     in place in the shard: no gather, no scatter
//     momentum == 0: w' = w - lr * d and the momentum buffer is left untouched (torch.optim.SGD ignores dampening and
//     never creates or updates a buffer in that case)
//     [wn[r] = bf16(w' / max(||w'||, 1e-12)), inv_norm[r] = 1 / max(||w'||, 1e-12)]     (valid for the next step when its
//                                                                                        sample has the same rows: sample_rate 1)
// = 3 reads + 2 writes (+ a half-width write).  lr is read from device memory when lr_dev is given (captured graphs
// with LR schedules), else passed by value.  HBM-bound: 20 (+2) bytes per element.
#pragma once
#include "common.cuh"

namespace msml {

constexpr int kSgdThreads = 256;
constexpr int kSgdMaxD = 1024;         // a lane keeps its slice of the row in registers: D / 128 float4 per lane

struct SgdParams {
  float lr, momentum, weight_decay, dampening;
  int nesterov;
};

// PROJECT: `dw` holds the RAW gradient of the normalised centres, dWn = dcos^T X (msml_head_bwd_raw); the backward of
// ref partial_fc.py:115 `normalize(sub_weight)` is applied here, on the row the warp already holds in registers:
//     n = max(||w||, 1e-12);  g = dWn / n - w * <w, dWn> / n^3        (= (dWn - wn <wn, dWn>) / n with wn = w / n)
// so the two GEMM epilogues that used to do it (a <Wn, dWn> column reduction in dcos, a Wn stream in dW) have nothing to do.
// It uses the fp32 master row, as the reference's autograd does (the unfused path projects with the bf16-rounded wn).
template <int VPL, bool PROJECT = false>   // float4 vectors per lane: D = 128 * VPL
__global__ void __launch_bounds__(kSgdThreads)
pfc_sgd_kernel(float* __restrict__ weight, float* __restrict__ weight_mom, const float* __restrict__ dw,
               const int64_t* __restrict__ index, int64_t n_s, int64_t num_local, const float* __restrict__ lr_dev, SgdParams p,
               __nv_bfloat16* __restrict__ wn, float* __restrict__ inv_norm) {
  constexpr int D = 128 * VPL;
  const int64_t r = (int64_t)blockIdx.x * (kSgdThreads / 32) + (threadIdx.x >> 5);
  if (r >= n_s) return;
  const int lane = threadIdx.x & 31;
  const int64_t row = index ? index[r] : r;
  if (row < 0 || row >= num_local) return;                 // an index outside the shard is never dereferenced
  const float lr = lr_dev ? lr_dev[0] : p.lr;
  float4* w4 = reinterpret_cast<float4*>(weight + row * D);
  float4* m4 = reinterpret_cast<float4*>(weight_mom + row * D);
  const float4* g4 = reinterpret_cast<const float4*>(dw + r * D);
  float4 w[VPL], m[VPL], g[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {                          // all loads first: 3 * VPL independent 16-byte loads per lane
    g[v] = __ldg(g4 + v * 32 + lane);
    w[v] = w4[v * 32 + lane];
    m[v] = m4[v * 32 + lane];
  }
  float ga = 1.f, gb = 0.f;              // g_true = ga * g_raw - gb * w
  if (PROJECT) {
    float s2 = 0.f, dot = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const float* wf = reinterpret_cast<const float*>(&w[v]);
      const float* gf = reinterpret_cast<const float*>(&g[v]);
#pragma unroll
      for (int i = 0; i < 4; ++i) { s2 = fmaf(wf[i], wf[i], s2); dot = fmaf(wf[i], gf[i], dot); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s2 += __shfl_xor_sync(0xffffffffu, s2, o); dot += __shfl_xor_sync(0xffffffffu, dot, o); }
    const float inv_n = 1.0f / fmaxf(sqrtf(s2), 1e-12f);
    ga = inv_n;
    gb = dot * inv_n * inv_n * inv_n;
  }
  float ss = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    float* wf = reinterpret_cast<float*>(&w[v]);
    float* mf = reinterpret_cast<float*>(&m[v]);
    const float* gf = reinterpret_cast<const float*>(&g[v]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float gt = PROJECT ? fmaf(ga, gf[i], -gb * wf[i]) : gf[i];
      const float d = fmaf(p.weight_decay, wf[i], gt);
      const float mn = fmaf(p.momentum, mf[i], (1.f - p.dampening) * d);
      const float upd = p.momentum == 0.f ? d : (p.nesterov ? fmaf(p.momentum, mn, d) : mn);
      const float wnew = fmaf(-lr, upd, wf[i]);
      mf[i] = mn;
      wf[i] = wnew;
      ss = fmaf(wnew, wnew, ss);
    }
    w4[v * 32 + lane] = w[v];
    if (p.momentum != 0.f) m4[v * 32 + lane] = m[v];
  }
  if (wn == nullptr) return;                               // uniform over the grid
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  if (lane == 0 && inv_norm) inv_norm[r] = inv;
  __nv_bfloat16* out = wn + r * D;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const float* wf = reinterpret_cast<const float*>(&w[v]);
    __nv_bfloat162 lo = __floats2bfloat162_rn(wf[0] * inv, wf[1] * inv), hi = __floats2bfloat162_rn(wf[2] * inv, wf[3] * inv);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(out + (v * 32 + lane) * 4) = pk;
  }
}

inline int pfc_sgd_check(int64_t n_s, int64_t num_local, int64_t D) {
  MSML_REQUIRE(n_s >= 0 && num_local > 0 && n_s <= num_local, MSML_EINVAL, "bad sample size n_s=%lld of %lld rows", (long long)n_s,
               (long long)num_local);
  MSML_REQUIRE(D > 0 && D % 128 == 0 && D <= kSgdMaxD, MSML_EUNSUPPORTED, "D=%lld: multiples of 128 up to %d are supported",
               (long long)D, kSgdMaxD);
  return 0;
}

}  // namespace msml
