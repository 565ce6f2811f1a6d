// K-N  fused BatchNorm (+ residual add) (+ PReLU): geometry, grid sizing and the C-ABI launchers.
//      Kernels and the algorithm: bn_act_kernels.cuh
#include <cstdlib>
#include <map>

#include "bn_act_kernels.cuh"

namespace msml {

static int bn_geom(int64_t P, int64_t C, int dtype, BnGeom* g) {
  MSML_REQUIRE(P > 0 && C > 0, MSML_EINVAL, "bad BN shape P=%lld C=%lld", (long long)P, (long long)C);
  MSML_REQUIRE(dtype == MSML_F32 || dtype == MSML_BF16 || dtype == MSML_F16, MSML_EINVAL, "unknown dtype %d", dtype);
  const int vn = dtype == MSML_F32 ? 4 : 8;
  MSML_REQUIRE(C % vn == 0, MSML_EUNSUPPORTED, "C=%lld must be a multiple of %d", (long long)C, vn);
  const int vpr = (int)(C / vn);
  MSML_REQUIRE(vpr <= kBnThreads && kBnThreads % vpr == 0, MSML_EUNSUPPORTED,
               "C=%lld: vectors per row (%d) must divide %d", (long long)C, vpr, kBnThreads);
  g->P = P; g->C = (int)C; g->vpr = vpr; g->rows_per_pass = kBnThreads / vpr;
  static const int skip = getenv("MSML_BN_SKIP_PHASES") ? atoi(getenv("MSML_BN_SKIP_PHASES")) : 0;
  g->skip = skip;
  return 0;
}

// L2 residency hints (BnGeom.skip & 32) for ops whose slab (`tensors` streams of P x C elements) can stay in the 126 MB L2
// between the statistics pass and the apply pass.  ncu (--cache-control none, every launch of one step) showed the backward
// apply pass re-reading about one tensor per op from HBM even for 13 MB layers: the convolution weight gradients on the side
// stream push the slab out between the two passes.  Measured on one B200, alternating same-box runs: 15.24 / 15.24 ms per step
// without, 15.13 / 15.13 with a 64 MB limit, 15.26 / 15.26 with 110 MB (larger slabs do not fit and only pollute); again on the
// final single-code-path kernels (every slab load carries a block-uniform policy, evict_normal when the hint is off; a
// separate unhinted path had cost registers and spilled): 14.89 -> 14.79 ms.
// MSML_BN_L2_KEEP=0 disables, MSML_BN_L2_KEEP_MB sets the limit.
static void bn_l2_hint(BnGeom* g, int dtype, int tensors) {
  static const int fused = getenv("MSML_BN_FUSED") ? atoi(getenv("MSML_BN_FUSED")) : 0;   // the single cooperative launch (comparison
  if (fused) return;                                                                      // mode) has no per-phase policy
  static const int on = getenv("MSML_BN_L2_KEEP") ? atoi(getenv("MSML_BN_L2_KEEP")) : 1;
  static const double cap_mb = getenv("MSML_BN_L2_KEEP_MB") ? atof(getenv("MSML_BN_L2_KEEP_MB")) : 64.0;
  if (!on) return;
  const double mb = (double)g->P * g->C * (dtype == MSML_F32 ? 4.0 : 2.0) * tensors / 1e6;
  if (mb <= cap_mb) g->skip |= 32;
}

// Every CTA must be co-resident (grid barrier), so the grid is capped by the occupancy of the
// instantiation; slabs of >= 4 passes per CTA.
template <typename K>
static int coop_grid(K kernel, const BnGeom& g, int* out) {
  static thread_local std::map<const void*, int> cache;     // max co-resident CTAs per kernel instantiation
  const void* key = reinterpret_cast<const void*>(kernel);
  auto it = cache.find(key);
  int cap;
  if (it == cache.end()) {
    int per_sm = 0;
    MSML_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBnThreads, 0));
    MSML_REQUIRE(per_sm > 0, MSML_EUNSUPPORTED, "BN kernel cannot be resident");
    cap = per_sm * num_sms();
    cache[key] = cap;
  } else {
    cap = it->second;
  }
  if (cap > kBnMaxCtas) cap = kBnMaxCtas;
  static const int env_cap = getenv("MSML_BN_MAX_CTAS") ? atoi(getenv("MSML_BN_MAX_CTAS")) : 0;   // debug
  if (env_cap > 0 && cap > env_cap) cap = env_cap;
  static const int env_passes = getenv("MSML_BN_PASSES") ? atoi(getenv("MSML_BN_PASSES")) : 0;   // debug
  const int passes = env_passes > 0 ? env_passes : 4;      // measured: 2..8 passes per CTA are equivalent, 16+ lose (each pass is a
                                                           // ~1.5 us memory round trip, so fewer / fatter CTAs serialise latency)
  int64_t want = (g.P + (int64_t)g.rows_per_pass * passes - 1) / ((int64_t)g.rows_per_pass * passes);
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  *out = (int)want;
  return 0;
}

}  // namespace msml

using namespace msml;

static size_t bn_ws_floats(int C) { return (size_t)kBnMaxCtas * 3 * C + kBnMaxCtas + 3 * (size_t)C + 16; }

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

static int launch_grid_synced(const void* kern, int grid, void** args, cudaStream_t st) {
  MSML_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kBnThreads), args, 0, st));
  count_launch();
  return 0;
}

// Default: the three phases as three PLAIN launches of the phase-specialised kernels (phase 2 on C CTAs); the kernel
// boundary plays the role of the grid barrier.  Measured inside CUDA graphs on B200 this beats the single cooperative
// launch at every layer size (205 MB: 118 vs 147 us forward, 259 vs 349 us backward; 6.4 MB: 16.5 vs 18.9 us): a kernel
// boundary costs ~2 us there, a grid barrier 2-5 us (arrival skew of 300-600 CTAs + same-address atomics) on top of the
// dearer cooperative launch, and each phase compiled on its own needs fewer registers.  MSML_BN_FUSED=1 selects the
// cooperative single launch (kept for that comparison).
static int bn_fused_mode() {
  static const int mode = getenv("MSML_BN_FUSED") ? atoi(getenv("MSML_BN_FUSED")) : 0;
  return mode;
}
// The FIRST launch of an op as a programmatic dependent of the kernel before it in the stream (a cuDNN convolution, usually):
// that kernel never triggers early, so the launch is released when its last CTA exits, but without the full kernel-boundary
// flush + launch latency in between; the BN kernel blocks in griddepcontrol.wait until the predecessor's writes are visible.
static bool bn_pdl_first() {
  static const int on = getenv("MSML_BN_PDL_FIRST") ? atoi(getenv("MSML_BN_PDL_FIRST")) : 0;
  return on != 0;
}
static int launch_plain(const void* kern, int blocks, void** args, cudaStream_t st, bool dependent = false) {
  static const int pdl = getenv("MSML_BN_PDL") ? atoi(getenv("MSML_BN_PDL")) : 1;     // 0 disables (comparison)
  if (dependent && pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(kBnThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MSML_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
  } else {
    MSML_CUDA(cudaLaunchKernel(kern, dim3(blocks), dim3(kBnThreads), args, 0, st));
  }
  count_launch();
  return 0;
}

template <typename T, bool RES, bool PRELU>
static int launch_bn_fwd_fused(const void* x, const void* res, void* y, const float* gamma, const float* beta, const float* prelu,
                               float* running_mean, float* running_var, int64_t* nbt, float momentum, float eps, float* save_mean,
                               float* save_invstd, float* part, float* part_n, float* coef, BnGeom g, cudaStream_t st,
                               float* next_part, float* next_part_n, bool stats_ready) {
  const T* xp = static_cast<const T*>(x);
  const T* rp = static_cast<const T*>(res);
  T* yp = static_cast<T*>(y);
  long long* nb = reinterpret_cast<long long*>(nbt);
  void* args[] = {&xp, &rp, &yp, &gamma, &beta, &prelu, &running_mean, &running_var, &nb, &momentum, &eps,
                  &save_mean, &save_invstd, &part, &part_n, &coef, &g, &next_part, &next_part_n};
  if (bn_fused_mode()) {
    MSML_REQUIRE(!next_part && !stats_ready, MSML_EUNSUPPORTED, "chained BN statistics need the split launches (MSML_BN_FUSED=0)");
    auto kern = bn_fwd_fused_kernel<T, RES, PRELU, 0>;
    int grid = 0;
    if (int e = coop_grid(kern, g, &grid)) return e;
    g.G = grid;
    return launch_grid_synced(reinterpret_cast<const void*>(kern), grid, args, st);
  }
  // each streaming phase gets a grid of at most ONE resident wave of its own kernel (a second, partly filled wave of the
  // statically partitioned slabs would only add a tail); phase 3 may partition the rows differently from phase 1
  int g1 = 0, g3 = 0;
  const void* k3 = next_part ? reinterpret_cast<const void*>(bn_fwd_fused_kernel<T, RES, PRELU, 3, true>)
                             : reinterpret_cast<const void*>(bn_fwd_fused_kernel<T, RES, PRELU, 3>);
  if (int e = coop_grid(bn_fwd_fused_kernel<T, RES, PRELU, 1>, g, &g1)) return e;
  if (next_part) { if (int e = coop_grid(bn_fwd_fused_kernel<T, RES, PRELU, 3, true>, g, &g3)) return e; }
  else if (int e = coop_grid(bn_fwd_fused_kernel<T, RES, PRELU, 3>, g, &g3)) return e;
  if (stats_ready) {
    // the producer's phase 3 left the slab statistics in this workspace: partial stride kBnMaxCtas, zero counts behind its grid
    g.G = kBnMaxCtas;
    if (int e = launch_plain(reinterpret_cast<const void*>(bn_fwd_fused_kernel<T, RES, PRELU, 2>), g.C, args, st, bn_pdl_first())) return e;
  } else {
    g.G = g1;
    if (int e = launch_plain(reinterpret_cast<const void*>(bn_fwd_fused_kernel<T, RES, PRELU, 1>), g1, args, st, bn_pdl_first())) return e;
    if (int e = launch_plain(reinterpret_cast<const void*>(bn_fwd_fused_kernel<T, RES, PRELU, 2>), g.C, args, st, true)) return e;
  }
  g.G = g3;
  return launch_plain(k3, g3, args, st, true);
}

template <typename T, bool RES, bool PRELU>
static int launch_bn_bwd_fused(const void* dy, const void* x, const void* res, const float* mean, const float* invstd,
                               const float* gamma, const float* beta, const float* prelu, void* dx, void* dres, const void* dadd, float* dgamma,
                               float* dbeta, float* dprelu, int training, int accumulate, float* part, float* coef, BnGeom g,
                               cudaStream_t st) {
  const T* dyp = static_cast<const T*>(dy);
  const T* xp = static_cast<const T*>(x);
  const T* rp = static_cast<const T*>(res);
  T* dxp = static_cast<T*>(dx);
  T* drp = static_cast<T*>(dres);
  const T* dap = static_cast<const T*>(dadd);
  void* args[] = {&dyp, &xp, &rp, &mean, &invstd, &gamma, &beta, &prelu, &dxp, &drp, &dap, &dgamma, &dbeta, &dprelu,
                  &training, &accumulate, &part, &coef, &g};
  if (bn_fused_mode()) {
    auto kern = bn_bwd_fused_kernel<T, RES, PRELU, 0>;
    int grid = 0;
    if (int e = coop_grid(kern, g, &grid)) return e;
    g.G = grid;
    return launch_grid_synced(reinterpret_cast<const void*>(kern), grid, args, st);
  }
  int g1 = 0, g3 = 0;
  if (int e = coop_grid(bn_bwd_fused_kernel<T, RES, PRELU, 1>, g, &g1)) return e;
  if (int e = coop_grid(bn_bwd_fused_kernel<T, RES, PRELU, 3>, g, &g3)) return e;
  g.G = g1;
  if (int e = launch_plain(reinterpret_cast<const void*>(bn_bwd_fused_kernel<T, RES, PRELU, 1>), g1, args, st, bn_pdl_first())) return e;
  if (int e = launch_plain(reinterpret_cast<const void*>(bn_bwd_fused_kernel<T, RES, PRELU, 2>), g.C, args, st, true)) return e;
  g.G = g3;
  return launch_plain(reinterpret_cast<const void*>(bn_bwd_fused_kernel<T, RES, PRELU, 3>), g3, args, st, true);
}

static int bn_fwd_impl(const void* x, const void* res, void* y, const float* gamma, const float* beta, const float* prelu,
                       float* running_mean, float* running_var, int64_t* num_batches_tracked, float* save_mean,
                       float* save_invstd, int64_t P, int64_t C, int dtype, int training, float momentum, float eps,
                       void* ws, size_t ws_bytes, void* next_ws, size_t next_ws_bytes, int stats_ready, void* stream) {
  BnGeom g;
  MSML_REQUIRE(training || (!next_ws && !stats_ready), MSML_EINVAL, "chained BN statistics exist in training mode only");
  MSML_REQUIRE(!next_ws || (aligned16(next_ws) && next_ws_bytes >= msml_bn_workspace(P, C)), MSML_EWORKSPACE,
               "next-op BN workspace too small or misaligned");
  MSML_REQUIRE(next_ws != ws || !next_ws, MSML_EINVAL, "next_ws must not alias ws");
  if (int e = bn_geom(P, C, dtype, &g)) return e;
  MSML_REQUIRE(x && y && save_mean && save_invstd, MSML_EINVAL, "null pointer");
  MSML_REQUIRE(training || (running_mean && running_var), MSML_EINVAL, "eval mode needs running statistics");
  MSML_REQUIRE(aligned16(x) && aligned16(y) && aligned16(res) && aligned16(ws) && aligned16(gamma) && aligned16(beta) &&
                   aligned16(prelu),
               MSML_EALIGN, "pointers must be 16-byte aligned");
  MSML_REQUIRE(ws && ws_bytes >= msml_bn_workspace(P, C), MSML_EWORKSPACE, "BN workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* part = static_cast<float*>(ws);
  float* part_n = part + (size_t)kBnMaxCtas * 3 * C;
  float* coef = part_n + kBnMaxCtas;
  float* next_part = static_cast<float*>(next_ws);
  float* next_part_n = next_part ? next_part + (size_t)kBnMaxCtas * 3 * C : nullptr;
  const double elem = dtype == MSML_F32 ? 4.0 : 2.0;
  if (training) {
    bn_l2_hint(&g, dtype, 1);
    MSML_PROF("bn_fwd", (double)P * C * elem * (res ? 3 : 2), st);
    int rc = 0;
    MSML_BN_DISPATCH(dtype, res != nullptr, prelu != nullptr,
                     (rc = launch_bn_fwd_fused<T, RES, PRELU>(x, res, y, gamma, beta, prelu, running_mean, running_var,
                                                              num_batches_tracked, momentum, eps, save_mean, save_invstd,
                                                              part, part_n, coef, g, st, next_part, next_part_n,
                                                              stats_ready != 0)));
    return rc;
  }
  bn_eval_coef_kernel<<<(unsigned)((C + 127) / 128), 128, 0, st>>>((int)C, gamma, beta, running_mean, running_var, eps, save_mean,
                                                                 save_invstd, coef);
  MSML_LAUNCH_CHECK();
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

extern "C" int msml_bn_fwd(const void* x, const void* res, void* y, const float* gamma, const float* beta, const float* prelu,
                           float* running_mean, float* running_var, int64_t* num_batches_tracked, float* save_mean,
                           float* save_invstd, int64_t P, int64_t C, int dtype, int training, float momentum, float eps,
                           void* ws, size_t ws_bytes, void* stream) {
  return bn_fwd_impl(x, res, y, gamma, beta, prelu, running_mean, running_var, num_batches_tracked, save_mean, save_invstd, P, C,
                     dtype, training, momentum, eps, ws, ws_bytes, nullptr, 0, 0, stream);
}

extern "C" int msml_bn_fwd_ex(const void* x, const void* res, void* y, const float* gamma, const float* beta, const float* prelu,
                              float* running_mean, float* running_var, int64_t* num_batches_tracked, float* save_mean,
                              float* save_invstd, int64_t P, int64_t C, int dtype, int training, float momentum, float eps,
                              void* ws, size_t ws_bytes, void* next_ws, size_t next_ws_bytes, int stats_ready, void* stream) {
  return bn_fwd_impl(x, res, y, gamma, beta, prelu, running_mean, running_var, num_batches_tracked, save_mean, save_invstd, P, C,
                     dtype, training, momentum, eps, ws, ws_bytes, next_ws, next_ws_bytes, stats_ready, stream);
}

extern "C" int msml_bn_bwd(const void* dy, const void* x, const void* res, const float* gamma, const float* beta,
                           const float* prelu, const float* save_mean, const float* save_invstd, void* dx, void* dres,
                           const void* dadd, float* dgamma, float* dbeta, float* dprelu, int64_t P, int64_t C, int dtype, int training,
                           int accumulate_param_grads, void* ws, size_t ws_bytes, void* stream) {
  BnGeom g;
  if (int e = bn_geom(P, C, dtype, &g)) return e;
  MSML_REQUIRE(dy && x && dx && save_mean && save_invstd, MSML_EINVAL, "null pointer");
  const bool has_prelu = prelu != nullptr, has_res = res != nullptr;
  MSML_REQUIRE(!(has_prelu && has_res) || dres, MSML_EINVAL, "dres is required when both a residual and PReLU are fused");
  MSML_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(res) && aligned16(dx) && aligned16(dres) && aligned16(ws) && aligned16(dadd) &&
                   aligned16(gamma) && aligned16(beta) && aligned16(prelu) && aligned16(save_mean) && aligned16(save_invstd),
               MSML_EALIGN, "pointers must be 16-byte aligned");
  MSML_REQUIRE(ws && ws_bytes >= msml_bn_workspace(P, C), MSML_EWORKSPACE, "BN workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* part = static_cast<float*>(ws);
  float* coef = part + (size_t)kBnMaxCtas * 3 * C + kBnMaxCtas;
  const double elem = dtype == MSML_F32 ? 4.0 : 2.0;
  const int both = has_prelu && has_res ? 1 : 0;
  bn_l2_hint(&g, dtype, 2 + both);
  MSML_PROF("bn_bwd", (double)P * C * elem * (3 + 2 * both + (dadd ? 1 : 0)), st);
  int rc = 0;
  MSML_BN_DISPATCH(dtype, has_res, has_prelu,
                   (rc = launch_bn_bwd_fused<T, RES, PRELU>(dy, x, res, save_mean, save_invstd, gamma, beta, prelu, dx, dres, dadd, dgamma,
                                                            dbeta, dprelu, training, accumulate_param_grads, part, coef, g, st)));
  return rc;
}
