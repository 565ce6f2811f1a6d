// K-C  input assembly of the Feature-Masking operator: C-ABI launchers.  Kernels and algorithm: fm_cat_kernels.cuh
//   ref backbones/fm/fmoperator.py:277-279  `x = torch.cat((yf, yo), dim=1)` feeding same_conv
#include "fm_cat_kernels.cuh"

using namespace msml;

extern "C" int msml_fm_cat_fwd(const void* yf, const void* yo, void* cat, int64_t P, int64_t C, int64_t Co, int64_t Ct, int dtype,
                               void* stream) {
  CatGeom g;
  if (int e = cat_geom(P, C, Co, Ct, dtype, &g)) return e;
  MSML_REQUIRE(cat && (yf || C == 0) && (yo || Co == 0), MSML_EINVAL, "null pointer");
  MSML_REQUIRE(aligned16(yf) && aligned16(cat), MSML_EALIGN, "yf and cat must be 16-byte aligned");
  g.nvec = P * g.vt;
  cudaStream_t st = (cudaStream_t)stream;
  const double elem = dtype == MSML_F32 ? 4.0 : 2.0;
  MSML_PROF("fm_cat_fwd", (double)P * (C + Co + Ct) * elem, st);
  MSML_DISPATCH_DTYPE(dtype, T, (fm_cat_fwd_kernel<T><<<cat_grid(g.nvec, num_sms()), kCatThreads, 0, st>>>(
                                    static_cast<const T*>(yf), static_cast<const T*>(yo), static_cast<T*>(cat), g)));
  MSML_LAUNCH_CHECK();
  return 0;
}

extern "C" int msml_fm_cat_bwd(const void* dcat, const void* dadd, void* dyf, void* dyo, int64_t P, int64_t C, int64_t Co,
                               int64_t Ct, int dtype, void* stream) {
  CatGeom g;
  if (int e = cat_geom(P, C, Co, Ct, dtype, &g)) return e;
  MSML_REQUIRE(dcat && (dyf || C == 0), MSML_EINVAL, "null pointer");
  MSML_REQUIRE(aligned16(dcat) && aligned16(dadd) && aligned16(dyf), MSML_EALIGN, "dcat, dadd and dyf must be 16-byte aligned");
  g.nvec = P * g.vf;
  cudaStream_t st = (cudaStream_t)stream;
  const double elem = dtype == MSML_F32 ? 4.0 : 2.0;
  MSML_PROF("fm_cat_bwd", (double)P * (C * (dadd ? 3 : 2) + (dyo ? 2 * Co : 0)) * elem, st);
  const int64_t walk = g.nvec > 0 ? g.nvec : (dyo ? (P * Co + 7) / 8 : 1);
  const int grid = cat_grid(walk, num_sms());
  MSML_DISPATCH_DTYPE(dtype, T, {
    if (dadd)
      fm_cat_bwd_kernel<T, true><<<grid, kCatThreads, 0, st>>>(static_cast<const T*>(dcat), static_cast<const T*>(dadd),
                                                               static_cast<T*>(dyf), static_cast<T*>(dyo), g);
    else
      fm_cat_bwd_kernel<T, false><<<grid, kCatThreads, 0, st>>>(static_cast<const T*>(dcat), nullptr, static_cast<T*>(dyf),
                                                                static_cast<T*>(dyo), g);
  });
  MSML_LAUNCH_CHECK();
  return 0;
}
