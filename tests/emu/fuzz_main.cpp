// Randomised edge-geometry run of the emulated kernels under AddressSanitizer / UBSan.  TEST INFRASTRUCTURE ONLY.
// Small random shapes (rows fewer than CTAs, one vector per row, sizes below one vector, ragged everything) on exact-size heap
// buffers: an out-of-bounds or misaligned access in a corner of the index arithmetic aborts the run.
//     g++ -std=c++20 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -pthread -I/usr/local/cuda/include fuzz_main.cpp
//     ./a.out [iterations] [seed]
#define MSML_CPU_EMU 1
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>
#include <random>
#include <string>

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

#define MSML_EMU_NO_ERR 1
#define MSML_EMU_TEMPLATES_ONLY 1
#include "../../msml_b200/csrc/fm_cat_kernels.cuh"
#include "../../msml_b200/csrc/seg_loss_kernels.cuh"
#include "../../msml_b200/csrc/sgd_flat_kernels.cuh"
#include "../../msml_b200/csrc/fm_peer_kernels.cuh"
#include "emu_bn.cpp"
#include "emu_fm_gate.cpp"
#include "emu_pfc_sample.cpp"

using namespace msml;

static std::mt19937 rng;
static int ri(int lo, int hi) { return lo + (int)(rng() % (unsigned)(hi - lo + 1)); }
static float frand() { return std::normal_distribution<float>(0.f, 1.f)(rng); }
template <typename T> static T conv(float v);
template <> float conv<float>(float v) { return v; }
template <> __nv_bfloat16 conv<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T, bool RES, bool PRELU>
static void fuzz_bn() {
  const int vn = sizeof(T) == 4 ? 4 : 8;
  static const int vprs[] = {1, 2, 4, 8, 16, 32, 64};
  const int64_t C = (int64_t)vprs[ri(0, 6)] * vn, P = ri(1, 300);
  const int G1 = ri(1, 9), G3 = ri(1, 9);
  BnGeom g;
  if (!emu_bn_geom(P, C, sizeof(T) == 4 ? MSML_F32 : MSML_BF16, &g)) return;
  std::vector<T> x(P * C), r(P * C), y(P * C), dy(P * C), dx(P * C), dres(P * C), dadd(P * C);
  for (auto& v : x) v = conv<T>(frand());
  for (auto& v : r) v = conv<T>(frand());
  for (auto& v : dy) v = conv<T>(frand());
  for (auto& v : dadd) v = conv<T>(frand());
  std::vector<float> gamma(C, 1.f), beta(C, 0.f), a(C, 0.25f), rm(C, 0.f), rv(C, 1.f), mean(C), invstd(C), dg(C), db(C), dp(C);
  std::vector<float> ws(emu_bn_ws_floats((int)C));
  long long nbt = 0;
  constexpr bool both = RES && PRELU;
  fwd3<T, RES, PRELU>(x.data(), RES ? r.data() : nullptr, y.data(), gamma.data(), beta.data(), PRELU ? a.data() : nullptr, rm.data(), rv.data(), &nbt,
                      0.1f, 1e-5f, mean.data(), invstd.data(), ws.data(), g, G1, G3);
  bwd3<T, both, PRELU>(dy.data(), x.data(), both ? r.data() : nullptr, mean.data(), invstd.data(), gamma.data(), beta.data(),
                       PRELU ? a.data() : nullptr, dx.data(), both ? dres.data() : nullptr, ri(0, 1) ? dadd.data() : nullptr, dg.data(), db.data(),
                       PRELU ? dp.data() : nullptr, 1, ri(0, 1), ws.data(), g, G1, G3);
}

// msml_bn_fwd_ex: a producer whose apply pass emits the statistics of its output into the consumer's exact-size workspace
// (partial stride kBnMaxCtas), then the consumer started at the merge.
template <typename T, bool PRELU_B>
static void fuzz_bn_chain() {
  const int vn = sizeof(T) == 4 ? 4 : 8;
  static const int vprs[] = {1, 2, 4, 8, 16, 32, 64};
  const int64_t C = (int64_t)vprs[ri(0, 6)] * vn, P = ri(1, 300);
  const int G1 = ri(1, 9), G3a = ri(1, 12), G3b = ri(1, 9);
  BnGeom g;
  if (!emu_bn_geom(P, C, sizeof(T) == 4 ? MSML_F32 : MSML_BF16, &g)) return;
  std::vector<T> x(P * C), r(P * C), y1(P * C), y2(P * C);
  for (auto& v : x) v = conv<T>(frand());
  for (auto& v : r) v = conv<T>(frand());
  std::vector<float> gamma(C, 1.f), beta(C, 0.f), a(C, 0.25f), rm(C, 0.f), rv(C, 1.f), mean(C), invstd(C);
  std::vector<float> ws_a(emu_bn_ws_floats((int)C)), ws_b(emu_bn_ws_floats((int)C));
  long long nbt = 0;
  fwd3<T, true, false>(x.data(), r.data(), y1.data(), gamma.data(), beta.data(), nullptr, nullptr, nullptr, nullptr, 0.1f, 1e-5f, mean.data(),
                       invstd.data(), ws_a.data(), g, G1, G3a, ws_b.data());
  fwd3<T, false, PRELU_B>(y1.data(), nullptr, y2.data(), gamma.data(), beta.data(), PRELU_B ? a.data() : nullptr, rm.data(), rv.data(), &nbt, 0.1f,
                          1e-5f, mean.data(), invstd.data(), ws_b.data(), g, G1, G3b, nullptr, true);
  for (float v : invstd)
    if (!(v > 0.f)) { printf("fuzz_bn_chain: bad invstd\n"); exit(3); }
}

static void fuzz_sgd_flat() {
  const int64_t n = 4 * (int64_t)ri(1, 3000);
  const int blocks = ri(1, 7);
  std::vector<float> w(n), m(n, 0.f), gr(n);
  for (auto& v : w) v = frand();
  for (auto& v : gr) v = frand();
  const bool with_shadow = ri(0, 1), with_scale = ri(0, 1);
  std::vector<__nv_bfloat16> shadow(with_shadow ? n : 0);
  float lr = 0.05f, scale = 1.5f;
  FlatSgdParams p{ri(0, 1) ? 0.9f : 0.f, 5e-4f, 0};
  p.nesterov = p.momentum != 0.f && ri(0, 1);
  const float* lr_p = &lr;
  const float* sc_p = with_scale ? &scale : nullptr;
  float* mp = p.momentum != 0.f ? m.data() : nullptr;
  emu_launch(dim3(blocks), kFlatSgdThreads, [&] { sgd_flat_kernel(w.data(), mp, gr.data(), with_shadow ? shadow.data() : nullptr, n / 4, lr_p, sc_p, p); });
}

template <typename T, int MODE, int ACT>
static void fuzz_peer() {
  const int64_t n = ri(1, 5000);
  const int blocks = ri(1, 6);
  std::vector<T> src(n), yf(n), yt(n), pf(n), pt(n), dpf(n), dpt(n), dsrc(n), dyf(n), da(n), db(n);
  for (auto* v : {&src, &yf, &yt, &dpf, &dpt})
    for (auto& e : *v) e = conv<T>(frand());
  if (ri(0, 1)) {
    emu_launch(dim3(blocks), kPeerThreads, [&] { fm_peer_mul_fwd_kernel<T, MODE, ACT, true>(src.data(), yf.data(), yt.data(), pf.data(), pt.data(), n); });
    emu_launch(dim3(blocks), kPeerThreads,
               [&] { fm_peer_mul_bwd_kernel<T, MODE, ACT, true>(dpf.data(), dpt.data(), src.data(), yf.data(), yt.data(), dsrc.data(), dyf.data(), n); });
  } else {
    emu_launch(dim3(blocks), kPeerThreads, [&] { fm_peer_mul_fwd_kernel<T, MODE, ACT, false>(src.data(), yf.data(), nullptr, pf.data(), nullptr, n); });
    emu_launch(dim3(blocks), kPeerThreads,
               [&] { fm_peer_mul_bwd_kernel<T, MODE, ACT, false>(dpf.data(), nullptr, src.data(), yf.data(), nullptr, dsrc.data(), dyf.data(), n); });
  }
  std::vector<float> partial(blocks);
  float out = 0.f, gout = 0.7f;
  float* pp = partial.data();
  float* op = &out;
  const float* gp = &gout;
  const bool both = ri(0, 1);
  emu_launch(dim3(blocks), kPeerThreads, [&] { mse_partial_kernel<T>(pf.data(), yf.data(), n, pp); });
  emu_launch(dim3(1), kPeerThreads, [&] { mse_finish_kernel(pp, blocks, n, op); });
  emu_launch(dim3(blocks), kPeerThreads, [&] { mse_bwd_kernel<T>(pf.data(), yf.data(), gp, da.data(), both ? db.data() : nullptr, n); });
  if (!(out >= 0.f)) { printf("fuzz_peer: bad mse\n"); exit(3); }
}

template <typename T>
static void fuzz_cat() {
  const int vn = sizeof(T) == 4 ? 4 : 8;
  const int64_t P = ri(1, 400), C = (int64_t)ri(0, 6) * vn, Co = ri(C == 0 ? 1 : 0, 20);
  const int64_t Ct = (C + Co + 7) / 8 * 8 + 8 * ri(0, 1);
  CatGeom g;
  if (Ct == 0 || cat_geom(P, C, Co, Ct, sizeof(T) == 4 ? MSML_F32 : MSML_BF16, &g)) return;
  std::vector<T> yf(P * C + 1), yo(P * Co + 1), cat(P * Ct), dcat(P * Ct), dadd(P * C + 1), dyf(P * C + 1), dyo(P * Co + 1);
  const int sms = ri(1, 3);
  g.nvec = P * g.vt;
  emu_launch(dim3((unsigned)cat_grid(g.nvec, sms)), kCatThreads, [&] { fm_cat_fwd_kernel<T>(yf.data(), yo.data(), cat.data(), g); });
  g.nvec = P * g.vf;
  const int64_t walk = g.nvec > 0 ? g.nvec : (P * Co + 7) / 8;
  emu_launch(dim3((unsigned)cat_grid(walk, sms)), kCatThreads,
             [&] { fm_cat_bwd_kernel<T, true>(dcat.data(), dadd.data(), dyf.data(), Co ? dyo.data() : nullptr, g); });
}

template <typename T>
static void fuzz_gate() {
  const int nseg = ri(1, 4);
  const int vn = sizeof(T) == 4 ? 4 : 8;
  std::vector<std::vector<T>> bufs(7 * nseg);
  FwdSegs fs{};
  BwdSegs bs{};
  fs.nseg = bs.nseg = nseg;
  int64_t n[MSML_MAX_SEGMENTS] = {};
  for (int i = 0; i < nseg; ++i) {
    n[i] = ri(0, 3) == 0 ? ri(1, 7) : ri(1, 5000);
    for (int k = 0; k < 7; ++k) bufs[7 * i + k].resize(n[i]);
    for (auto& v : bufs[7 * i + 1]) v = conv<T>(frand());
    fs.yf[i] = bufs[7 * i].data(); fs.z[i] = bufs[7 * i + 1].data(); fs.f_out[i] = nullptr; fs.out[i] = bufs[7 * i + 2].data(); fs.n[i] = n[i];
    bs.dout[i] = bufs[7 * i + 3].data(); bs.yf[i] = bufs[7 * i].data(); bs.z[i] = bufs[7 * i + 1].data();
    bs.dyf[i] = bufs[7 * i + 4].data(); bs.dz[i] = bufs[7 * i + 5].data(); bs.n[i] = n[i];
  }
  const int sms = ri(1, 3);
  deal_blocks(nseg, n, vn, kUnroll, kCtasPerSm, fs.block_end, sms);
  deal_blocks(nseg, n, vn, kUnroll3, kCtasPerSmBwd, bs.block_end, sms);
  run_fwd<T, 1, 3>(fs, false);
  run_bwd<T, 1, 3>(bs);
}

static void fuzz_seg() {
  const int64_t N = ri(1, 4), HW = ri(1, 2500);
  const int K = ri(1, 5), cl = ri(0, 1);
  SegGeom g;
  if (seg_geom(N, 2, HW, K, cl, MSML_F32, &g)) return;
  std::vector<float> logit(N * 2 * HW), dlogit(N * 2 * HW);
  for (auto& v : logit) v = 3.f * frand();
  std::vector<int64_t> blobs(N * HW), target(N * HW);
  for (int64_t i = 0; i < N * HW; ++i) { blobs[i] = ri(-1, K - 1); target[i] = blobs[i] < 0 ? 0 : blobs[i] % 2; }
  std::vector<float> partf(seg_partf(g)), acc(seg_acc(g)), coef(2 * (size_t)K * N * 2), loss(1);
  std::vector<int> parti(seg_parti(g)), partbad(seg_partbad(g));
  emu_launch(dim3((unsigned)g.chunks, (unsigned)N), kSegThreads,
             [&] { seg_stats_kernel<float, 2>(logit.data(), blobs.data(), target.data(), partf.data(), parti.data(), partbad.data(), g); });
  const int pixel_all = ri(0, 1), kl_all = ri(0, 1);            // drawn once, outside the emulated threads
  emu_launch(dim3(1), kSegThreads, [&] {
    seg_finalize_kernel<2>(partf.data(), parti.data(), partbad.data(), target.data(), acc.data(), coef.data(), loss.data(), 10.f, 5.f, pixel_all, kl_all, g);
  });
  emu_launch(dim3((unsigned)g.chunks, (unsigned)N), kSegThreads,
             [&] { seg_bwd_kernel<float, 2>(logit.data(), blobs.data(), coef.data(), nullptr, dlogit.data(), g); });
}

static void fuzz_pfc() {
  const int64_t num_local = ri(1, 9000), num_sample = ri(0, (int)num_local), n_labels = ri(1, 300);
  std::vector<float> perm(num_local);
  for (auto& v : perm) v = (float)(rng() % 53) / 53.f;
  std::vector<int64_t> tl(n_labels);
  for (auto& v : tl) v = (int64_t)(rng() % (unsigned)(2 * num_local + 1));
  emu_pfc_remap(tl.data(), n_labels, 7, num_local);
  emu_pfc_mark_positive(perm.data(), tl.data(), n_labels, num_local);
  std::vector<int64_t> index(std::max<int64_t>(std::max(num_sample, n_labels), 1)), n_index(1);
  emu_pfc_select(perm.data(), num_local, num_sample, index.data(), n_index.data());
  if (n_index[0] < num_sample || n_index[0] > (int64_t)index.size()) { fprintf(stderr, "select returned %lld of %lld\n", (long long)n_index[0], (long long)num_sample); std::abort(); }
  emu_pfc_searchsorted(tl.data(), n_labels, index.data(), n_index.data());
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? std::stoi(argv[1]) : 6;
  rng.seed(argc > 2 ? std::stoul(argv[2]) : 1234u);
  for (int it = 0; it < iters; ++it) {
    fuzz_bn<float, false, false>();
    fuzz_bn<__nv_bfloat16, true, true>();
    fuzz_bn<__nv_bfloat16, false, true>();
    fuzz_cat<float>();
    fuzz_cat<__nv_bfloat16>();
    fuzz_gate<float>();
    fuzz_gate<__nv_bfloat16>();
    fuzz_seg();
    fuzz_pfc();
    fuzz_bn_chain<__nv_bfloat16, true>();
    fuzz_bn_chain<float, false>();
    fuzz_sgd_flat();
    fuzz_peer<__nv_bfloat16, 1, MSML_ACT_SIGMOID>();
    fuzz_peer<float, 0, MSML_ACT_SIGMOID>();
    fuzz_peer<float, 1, MSML_ACT_TANH>();
  }
  printf("fuzz: %d iterations clean\n", iters);
  return 0;
}
