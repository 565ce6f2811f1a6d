// Gradient plumbing kernels of the training step (SURVEY.md 8f-2 neighbourhood).
//   msml_pfc_sgd_update   : fused momentum SGD + weight decay + in-place write-back (+ next normalised bf16 centres) on the sampled
//     rows of the PartialFC shard; kernel and algorithm in pfc_sgd_kernels.cuh.
//   msml_sgd_flat         : momentum SGD + weight decay + gradient scale over one flat fp32 buffer, emitting the bf16 shadow weights of
//     the next step; kernel and algorithm in sgd_flat_kernels.cuh.
//   msml_accum_bf16_multi : dst_f32[i] += float(src_bf16[i]) for up to MSML_ACCUM_MAX_SEGMENTS tensors in ONE launch.
//     The bf16 weight gradients cuDNN returns are added into the fp32 flat-gradient views of engine.TrainStep after
//     the backward pass: one launch instead of one mixed-dtype ATen add (a non-vectorised kernel) per weight.
#include "common.cuh"
#include "accum_kernels.cuh"
#include "pfc_sgd_kernels.cuh"
#include "sgd_flat_kernels.cuh"

using namespace msml;

extern "C" int msml_accum_bf16_multi(int nseg, float* const* dst, const void* const* src, const int64_t* n, void* stream) {
  MSML_REQUIRE(nseg >= 0 && (nseg == 0 || (dst && src && n)), MSML_EINVAL, "bad accumulate arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int s0 = 0; s0 < nseg; s0 += MSML_ACCUM_MAX_SEGMENTS) {
    AccumSegs segs;
    int cnt = 0, blocks = 0;
    double bytes = 0.0;
    for (int i = s0; i < nseg && cnt < MSML_ACCUM_MAX_SEGMENTS; ++i) {
      if (n[i] <= 0) continue;
      MSML_REQUIRE(dst[i] && src[i], MSML_EINVAL, "null gradient pointer in segment %d", i);
      segs.dst[cnt] = dst[i];
      segs.src[cnt] = static_cast<const __nv_bfloat16*>(src[i]);
      segs.n[cnt] = n[i];
      blocks += (int)((n[i] + kAccElemsPerBlock - 1) / kAccElemsPerBlock);
      segs.block_end[cnt] = blocks;
      bytes += (double)n[i] * 10.0;                            // read bf16 + read/write fp32
      ++cnt;
    }
    if (cnt == 0) continue;
    segs.nseg = cnt;
    MSML_PROF("accum_bf16_multi", bytes, st);
    accum_bf16_multi_kernel<<<blocks, kAccThreads, 0, st>>>(segs);
    MSML_LAUNCH_CHECK();
  }
  return 0;
}

static int pfc_sgd_impl(float* weight, float* weight_mom, const float* dw, const int64_t* index, int64_t n_s,
                        int64_t num_local, int64_t D, const float* lr_dev, float lr, float momentum, float weight_decay,
                        float dampening, int nesterov, void* wn_bf16, float* inv_norm, void* stream, bool project) {
  if (int e = pfc_sgd_check(n_s, num_local, D)) return e;
  if (n_s == 0) return 0;
  MSML_REQUIRE(weight && weight_mom && dw, MSML_EINVAL, "null pointer");
  MSML_REQUIRE(!nesterov || (momentum > 0.f && dampening == 0.f), MSML_EINVAL, "nesterov needs momentum > 0 and zero dampening");
  MSML_REQUIRE(aligned16(weight) && aligned16(weight_mom) && aligned16(dw) && aligned16(wn_bf16), MSML_EALIGN,
               "weight, weight_mom, dw and wn must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  SgdParams p{lr, momentum, weight_decay, dampening, nesterov};
  const unsigned grid = (unsigned)((n_s + kSgdThreads / 32 - 1) / (kSgdThreads / 32));
  __nv_bfloat16* wn = static_cast<__nv_bfloat16*>(wn_bf16);
  MSML_PROF("pfc_sgd_update", (double)n_s * D * (20.0 + (wn ? 2.0 : 0.0)), st);
  switch (D / 128) {
#define MSML_SGD_CASE(V) case V: \
      if (project) pfc_sgd_kernel<V, true><<<grid, kSgdThreads, 0, st>>>(weight, weight_mom, dw, index, n_s, num_local, lr_dev, p, wn, inv_norm); \
      else pfc_sgd_kernel<V, false><<<grid, kSgdThreads, 0, st>>>(weight, weight_mom, dw, index, n_s, num_local, lr_dev, p, wn, inv_norm); \
      break;
    MSML_SGD_CASE(1) MSML_SGD_CASE(2) MSML_SGD_CASE(3) MSML_SGD_CASE(4) MSML_SGD_CASE(5) MSML_SGD_CASE(6) MSML_SGD_CASE(7) MSML_SGD_CASE(8)
#undef MSML_SGD_CASE
    default: return set_error(MSML_EUNSUPPORTED, "D=%lld unsupported", (long long)D);
  }
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" int msml_pfc_sgd_update(float* weight, float* weight_mom, const float* dw, const int64_t* index, int64_t n_s,
                                   int64_t num_local, int64_t D, const float* lr_dev, float lr, float momentum, float weight_decay,
                                   float dampening, int nesterov, void* wn_bf16, float* inv_norm, void* stream) {
  return pfc_sgd_impl(weight, weight_mom, dw, index, n_s, num_local, D, lr_dev, lr, momentum, weight_decay, dampening, nesterov,
                      wn_bf16, inv_norm, stream, false);
}

extern "C" int msml_pfc_sgd_update_raw(float* weight, float* weight_mom, const float* dwn, const int64_t* index, int64_t n_s,
                                       int64_t num_local, int64_t D, const float* lr_dev, float lr, float momentum,
                                       float weight_decay, float dampening, int nesterov, void* wn_bf16, float* inv_norm,
                                       void* stream) {
  return pfc_sgd_impl(weight, weight_mom, dwn, index, n_s, num_local, D, lr_dev, lr, momentum, weight_decay, dampening, nesterov,
                      wn_bf16, inv_norm, stream, true);
}

extern "C" int msml_sgd_flat(float* weight, float* momentum_buf, const float* grad, void* shadow_bf16, int64_t n, const float* lr_dev,
                             const float* grad_scale_dev, float momentum, float weight_decay, int nesterov, void* stream) {
  MSML_REQUIRE(n >= 0 && n % 4 == 0, MSML_EINVAL, "flat SGD: n=%lld must be a non-negative multiple of 4", (long long)n);
  if (n == 0) return 0;
  MSML_REQUIRE(weight && grad && lr_dev && (momentum == 0.f || momentum_buf), MSML_EINVAL, "null pointer");
  MSML_REQUIRE(!nesterov || momentum > 0.f, MSML_EINVAL, "nesterov needs momentum > 0");
  MSML_REQUIRE(aligned16(weight) && aligned16(momentum_buf) && aligned16(grad) && aligned16(shadow_bf16), MSML_EALIGN,
               "weight, momentum, grad and shadow must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n4 = n / 4;
  const int64_t per_cta = (int64_t)kFlatSgdThreads * kFlatSgdUnroll;
  int64_t blocks = (n4 + per_cta - 1) / per_cta;
  const int64_t cap = (int64_t)num_sms() * 8;                 // one resident wave of 8 CTAs per SM, grid-stride beyond
  if (blocks > cap) blocks = cap;
  FlatSgdParams p{momentum, weight_decay, nesterov};
  MSML_PROF("sgd_flat", (double)n * ((momentum != 0.f ? 20.0 : 12.0) + (shadow_bf16 ? 2.0 : 0.0)), st);
  sgd_flat_kernel<<<(unsigned)blocks, kFlatSgdThreads, 0, st>>>(weight, momentum_buf, grad, static_cast<__nv_bfloat16*>(shadow_bf16), n4,
                                                             lr_dev, grad_scale_dev, p);
  MSML_LAUNCH_CHECK();
  return 0;
}
