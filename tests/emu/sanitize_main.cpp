// Standalone sanitizer run of the emulated kernels.  TEST INFRASTRUCTURE ONLY.
// Built twice by tests/test_emu_kernels.py: with -fsanitize=address,undefined (every buffer is an exact-size heap
// allocation, so an out-of-bounds access of a kernel is reported) and with -fsanitize=thread (the threads of a CTA are
// real threads: a missing __syncthreads or two threads writing one location shows up as a data race) — the CPU
// counterparts of compute-sanitizer's memcheck and racecheck.  Exit code 0 = ran to completion; the sanitizers abort or
// print a report otherwise.
#define MSML_CPU_EMU 1
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>
#include <random>
#include <string>

#include "../../msml_b200/csrc/fm_cat_kernels.cuh"
#include "../../msml_b200/csrc/pfc_sgd_kernels.cuh"
#include "../../msml_b200/csrc/seg_loss_kernels.cuh"
#include "../../msml_b200/csrc/sgd_flat_kernels.cuh"
#include "../../msml_b200/csrc/fm_peer_kernels.cuh"

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
#include "emu_bn.cpp"      // the three-launch BN harness (emu_bn_fwd / emu_bn_bwd)
#include "emu_fm_gate.cpp" // the K-A harness (emu_fm_gate_fwd_multi / emu_fm_gate_bwd_multi)
#include "emu_dap.cpp"         // K-B: DAP + argmax mask
#include "emu_fm_mask.cpp"     // K-A extension: shared-memory gate tile, float atomics
#include "emu_head_small.cpp"  // margin math, weight normalisation, statistics merge, bf16 transpose
#include "emu_pfc_sample.cpp"  // the K-D harness (remap, mark, radix select with tickets and atomics, searchsorted, row copies)

using namespace msml;

static std::mt19937 rng(7);
static float frand() { return std::normal_distribution<float>(0.f, 1.f)(rng); }

template <typename T> static T conv(float v);
template <> float conv<float>(float v) { return v; }
template <> __nv_bfloat16 conv<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T, int C>
static int run_seg(int64_t N, int64_t HW, int K, int cl) {
  SegGeom g;
  if (seg_geom(N, C, HW, K, cl, sizeof(T) == 4 ? MSML_F32 : MSML_BF16, &g)) return 1;
  std::vector<T> logit(N * C * HW), dlogit(N * C * HW);
  for (auto& v : logit) v = conv<T>(frand());
  std::vector<int64_t> blobs(N * HW), target(N * HW);
  for (int64_t i = 0; i < N * HW; ++i) {
    int b = (int)(rng() % (K + 1)) - 1;                       // -1 .. K-1: ignored pixels included
    if (i >= (N - 1) * HW && b == 1) b = 0;                   // blob 1 is missing from the last sample
    blobs[i] = b;
    target[i] = b < 0 ? 0 : (b + 1) % C;
  }
  std::vector<float> partf(seg_partf(g)), acc(seg_acc(g)), coef(2 * (size_t)K * N * C), loss(1), gout(1, 1.5f);
  std::vector<int> parti(seg_parti(g)), partbad(seg_partbad(g));
  emu_launch(dim3((unsigned)g.chunks, (unsigned)N), kSegThreads,
             [&] { seg_stats_kernel<T, C>(logit.data(), blobs.data(), target.data(), partf.data(), parti.data(), partbad.data(), g); });
  emu_launch(dim3(1), kSegThreads, [&] {
    seg_finalize_kernel<C>(partf.data(), parti.data(), partbad.data(), target.data(), acc.data(), coef.data(), loss.data(), 10.f, 5.f, 0, 0, g);
  });
  emu_launch(dim3((unsigned)g.chunks, (unsigned)N), kSegThreads,
             [&] { seg_bwd_kernel<T, C>(logit.data(), blobs.data(), coef.data(), gout.data(), dlogit.data(), g); });
  if (!(loss[0] == loss[0])) { fprintf(stderr, "consensus loss is NaN\n"); return 1; }
  return 0;
}

template <int V>
static int run_sgd(int64_t num_local, int64_t n_s, bool sampled) {
  constexpr int D = 128 * V;
  std::vector<float> w(num_local * D), m(num_local * D), dw(n_s * D), inv(n_s);
  for (auto& v : w) v = 0.01f * frand();
  for (auto& v : m) v = 0.001f * frand();
  for (auto& v : dw) v = 0.1f * frand();
  std::vector<int64_t> index(n_s);
  for (int64_t r = 0; r < n_s; ++r) index[r] = r * num_local / n_s;     // distinct, sorted
  std::vector<__nv_bfloat16> wn(n_s * D);
  SgdParams p{0.1f, 0.9f, 5e-4f, 0.f, 0};
  const unsigned grid = (unsigned)((n_s + kSgdThreads / 32 - 1) / (kSgdThreads / 32));
  emu_launch(dim3(grid), kSgdThreads, [&] {
    pfc_sgd_kernel<V>(w.data(), m.data(), dw.data(), sampled ? index.data() : nullptr, n_s, num_local, nullptr, p, wn.data(), inv.data());
  });
  return 0;
}

template <typename T>
static int run_cat(int64_t P, int64_t C, int64_t Co, int sms) {
  const int vn = sizeof(T) == 4 ? 4 : 8;
  const int64_t Ct = (C + Co + 7) / 8 * 8;
  CatGeom g;
  if (cat_geom(P, C, Co, Ct, sizeof(T) == 4 ? MSML_F32 : MSML_BF16, &g)) return 1;
  (void)vn;
  std::vector<T> yf(P * C), yo(P * Co), cat(P * Ct), dcat(P * Ct), dadd(P * C), dyf(P * C), dyo(P * Co);
  for (auto& v : yf) v = conv<T>(frand());
  for (auto& v : yo) v = conv<T>(frand());
  for (auto& v : dcat) v = conv<T>(frand());
  for (auto& v : dadd) v = conv<T>(frand());
  g.nvec = P * g.vt;
  emu_launch(dim3((unsigned)cat_grid(g.nvec, sms)), kCatThreads, [&] { fm_cat_fwd_kernel<T>(yf.data(), yo.data(), cat.data(), g); });
  g.nvec = P * g.vf;
  emu_launch(dim3((unsigned)cat_grid(g.nvec, sms)), kCatThreads,
             [&] { fm_cat_bwd_kernel<T, true>(dcat.data(), dadd.data(), dyf.data(), dyo.data(), g); });
  return 0;
}

template <typename T, bool RES, bool PRELU>
static int run_bn(int64_t P, int64_t C, int G1, int G3) {
  const int dtype = sizeof(T) == 4 ? MSML_F32 : MSML_BF16;
  BnGeom g;
  if (!emu_bn_geom(P, C, dtype, &g)) return 1;
  std::vector<T> x(P * C), r(P * C), y(P * C), dy(P * C), dx(P * C), dres(P * C), dadd(P * C);
  for (auto& v : x) v = conv<T>(1.f + 2.f * frand());
  for (auto& v : r) v = conv<T>(frand());
  for (auto& v : dy) v = conv<T>(frand());
  for (auto& v : dadd) v = conv<T>(frand());
  std::vector<float> gamma(C, 1.1f), beta(C, -0.2f), a(C, 0.25f), rm(C, 0.f), rv(C, 1.f), mean(C), invstd(C), dg(C, 0.f), db(C, 0.f), dp(C, 0.f);
  std::vector<float> ws(emu_bn_ws_floats((int)C));
  long long nbt = 0;
  constexpr bool both = PRELU && RES;
  fwd3<T, RES, PRELU>(x.data(), RES ? r.data() : nullptr, y.data(), gamma.data(), beta.data(), PRELU ? a.data() : nullptr, rm.data(), rv.data(),
                      &nbt, 0.1f, 1e-5f, mean.data(), invstd.data(), ws.data(), g, G1, G3);
  bwd3<T, both, PRELU>(dy.data(), x.data(), both ? r.data() : nullptr, mean.data(), invstd.data(), gamma.data(), beta.data(),
                       PRELU ? a.data() : nullptr, dx.data(), both ? dres.data() : nullptr, dadd.data(), dg.data(), db.data(),
                       PRELU ? dp.data() : nullptr, 1, 1, ws.data(), g, G1, G3);
  return nbt != 1;
}

// msml_bn_fwd_ex: producer (bn + res, emits the statistics of its output) -> consumer (starts at the merge); the consumer's
// workspace is an exact-size allocation the producer writes with the kBnMaxCtas partial stride.
template <typename T>
static int run_bn_chain(int64_t P, int64_t C, int G1, int G3a, int G3b) {
  const int dtype = sizeof(T) == 4 ? MSML_F32 : MSML_BF16;
  BnGeom g;
  if (!emu_bn_geom(P, C, dtype, &g)) return 1;
  std::vector<T> x(P * C), r(P * C), y1(P * C), y2(P * C);
  for (auto& v : x) v = conv<T>(1.f + 2.f * frand());
  for (auto& v : r) v = conv<T>(frand());
  std::vector<float> gamma(C, 1.1f), beta(C, -0.2f), a(C, 0.25f), rm(C, 0.f), rv(C, 1.f), mean(C), invstd(C);
  std::vector<float> ws_a(emu_bn_ws_floats((int)C)), ws_b(emu_bn_ws_floats((int)C));
  long long nbt = 0;
  fwd3<T, true, false>(x.data(), r.data(), y1.data(), gamma.data(), beta.data(), nullptr, nullptr, nullptr, nullptr, 0.1f, 1e-5f, mean.data(),
                       invstd.data(), ws_a.data(), g, G1, G3a, ws_b.data());
  fwd3<T, false, true>(y1.data(), nullptr, y2.data(), gamma.data(), beta.data(), a.data(), rm.data(), rv.data(), &nbt, 0.1f, 1e-5f, mean.data(),
                       invstd.data(), ws_b.data(), g, G1, G3b, nullptr, true);
  return nbt != 1;
}

static int run_sgd_flat(int64_t n, int blocks, bool with_shadow, bool with_scale) {
  std::vector<float> w(n), m(n, 0.f), gr(n);
  for (auto& v : w) v = frand();
  for (auto& v : gr) v = frand();
  std::vector<__nv_bfloat16> shadow(with_shadow ? n : 0);
  float lr = 0.05f, scale = 1.5f;
  FlatSgdParams p{0.9f, 5e-4f, 0};
  const float* lr_p = &lr;
  const float* sc_p = with_scale ? &scale : nullptr;
  emu_launch(dim3(blocks), kFlatSgdThreads,
             [&] { sgd_flat_kernel(w.data(), m.data(), gr.data(), with_shadow ? shadow.data() : nullptr, n / 4, lr_p, sc_p, p); });
  return 0;
}

template <typename T, int MODE>
static int run_peer(int64_t n, int blocks) {
  std::vector<T> src(n), yf(n), yt(n), pf(n), pt(n), dpf(n), dpt(n), dsrc(n), dyf(n), da(n), db(n);
  for (auto* v : {&src, &yf, &yt, &dpf, &dpt})
    for (auto& e : *v) e = conv<T>(frand());
  emu_launch(dim3(blocks), kPeerThreads, [&] { fm_peer_mul_fwd_kernel<T, MODE, MSML_ACT_SIGMOID, true>(src.data(), yf.data(), yt.data(), pf.data(), pt.data(), n); });
  emu_launch(dim3(blocks), kPeerThreads, [&] { fm_peer_mul_fwd_kernel<T, MODE, MSML_ACT_TANH, false>(src.data(), yf.data(), nullptr, pf.data(), nullptr, n); });
  emu_launch(dim3(blocks), kPeerThreads,
             [&] { fm_peer_mul_bwd_kernel<T, MODE, MSML_ACT_SIGMOID, true>(dpf.data(), dpt.data(), src.data(), yf.data(), yt.data(), dsrc.data(), dyf.data(), n); });
  std::vector<float> partial(blocks);
  float out = 0.f, gout = 0.7f;
  float* pp = partial.data();
  float* op = &out;
  const float* gp = &gout;
  emu_launch(dim3(blocks), kPeerThreads, [&] { mse_partial_kernel<T>(pf.data(), pt.data(), n, pp); });
  emu_launch(dim3(1), kPeerThreads, [&] { mse_finish_kernel(pp, blocks, n, op); });
  emu_launch(dim3(blocks), kPeerThreads, [&] { mse_bwd_kernel<T>(pf.data(), pt.data(), gp, da.data(), db.data(), n); });
  return !(out >= 0.f);
}

template <typename T, int ACT, int ARITH>
static int run_gate(bool with_fout) {
  const int64_t n[3] = {3077, 19, 1203};                       // scalar tails, a segment smaller than one vector pass
  const int vn = sizeof(T) == 4 ? 4 : 8;
  std::vector<T> yf[3], z[3], fo[3], out[3], d[3], dyf[3], dz[3];
  FwdSegs fs{};
  BwdSegs bs{};
  fs.nseg = with_fout ? 1 : 3;
  bs.nseg = 3;
  for (int i = 0; i < 3; ++i) {
    for (auto* v : {&yf[i], &z[i], &fo[i], &out[i], &d[i], &dyf[i], &dz[i]}) v->resize(n[i]);
    for (auto& v : yf[i]) v = conv<T>(frand());
    for (auto& v : z[i]) v = conv<T>(frand());
    for (auto& v : fo[i]) v = conv<T>(frand());
    for (auto& v : d[i]) v = conv<T>(frand());
    fs.yf[i] = yf[i].data(); fs.z[i] = z[i].data(); fs.f_out[i] = with_fout ? fo[i].data() : nullptr; fs.out[i] = out[i].data(); fs.n[i] = n[i];
    bs.dout[i] = d[i].data(); bs.yf[i] = yf[i].data(); bs.z[i] = z[i].data(); bs.dyf[i] = dyf[i].data(); bs.dz[i] = dz[i].data(); bs.n[i] = n[i];
  }
  deal_blocks(fs.nseg, n, vn, with_fout ? kUnroll3 : kUnroll, kCtasPerSm, fs.block_end, 2);
  deal_blocks(3, n, vn, kUnroll3, kCtasPerSmBwd, bs.block_end, 2);
  run_fwd<T, ACT, ARITH>(fs, with_fout);
  run_bwd<T, ACT, ARITH>(bs);
  return 0;
}

static int run_pfc(int64_t num_local, int64_t num_sample, int64_t n_labels) {
  std::vector<float> perm(num_local);
  for (auto& v : perm) v = (float)(rng() % 997) / 997.f;        // ties at the threshold
  std::vector<int64_t> tl(n_labels);
  for (auto& v : tl) v = (int64_t)(rng() % (2 * num_local));
  emu_pfc_remap(tl.data(), n_labels, 100, num_local);
  emu_pfc_mark_positive(perm.data(), tl.data(), n_labels, num_local);
  std::vector<int64_t> index(num_sample > n_labels ? num_sample : n_labels), n_index(1);
  emu_pfc_select(perm.data(), num_local, num_sample, index.data(), n_index.data());
  if (n_index[0] < num_sample || n_index[0] > (int64_t)index.size()) return 1;
  for (int64_t i = 1; i < n_index[0]; ++i) if (index[i] <= index[i - 1]) return 1;     // sorted, distinct
  emu_pfc_searchsorted(tl.data(), n_labels, index.data(), n_index.data());
  const int64_t D = 32, rows = n_index[0];
  std::vector<float> w(num_local * D, 1.f), sub(rows * D);
  emu_gather_rows(w.data(), index.data(), sub.data(), rows, D);
  emu_scatter_rows(w.data(), index.data(), sub.data(), rows, D);
  return 0;
}

template <typename T>
static int run_mask(int64_t Cm, int64_t Hm, int64_t Wm) {
  const MaskGeom g{2, 7, 5, 64, Hm, Wm, Cm};                 // 70 pixels: a ragged last CTA of 32
  const int64_t n = g.B * g.H * g.W * g.C, nm = g.B * Hm * Wm * Cm;
  std::vector<T> yf(n), out(n), dout(n), dyf(n), m(nm);
  for (auto& v : yf) v = conv<T>(frand());
  for (auto& v : dout) v = conv<T>(frand());
  for (auto& v : m) v = conv<T>(frand());
  std::vector<float> dm(nm, 0.f);
  mask_fwd<T, 1, 3>(yf.data(), m.data(), out.data(), g);
  mask_bwd<T, 1, 3>(dout.data(), yf.data(), m.data(), dyf.data(), dm.data(), g);
  return 0;
}

static int run_small() {
  // K-B
  const int64_t B = 2, G = 2, kk = 9, HW = 131;
  std::vector<__nv_bfloat16> x(B * G * kk * HW), y(B * G * HW), dy(B * G * HW), dx(B * G * kk * HW);
  for (auto& v : x) v = conv<__nv_bfloat16>(frand());
  for (auto& v : dy) v = conv<__nv_bfloat16>(frand());
  std::vector<int64_t> mask(B * HW);
  emu_dap_fwd(x.data(), y.data(), mask.data(), B, G, kk, HW, 1, MSML_BF16, 2);
  emu_dap_bwd(dy.data(), dx.data(), B, G, kk, HW, 0, MSML_BF16, 1);
  // head
  const int64_t n = 19, D = 128;
  std::vector<float> w(n * D), inv(n);
  for (auto& v : w) v = 0.01f * frand();
  std::vector<__nv_bfloat16> wn(n * D), wt(D * 24);
  emu_wnorm_cast(w.data(), wn.data(), inv.data(), n, D, 1);
  emu_transpose_bf16(wn.data(), wt.data(), n, D, 24);
  const int W = 2, B_tot = 70, n_blocks = 5;
  std::vector<float> pmax(n_blocks * B_tot), psum(n_blocks * B_tot), tgt(B_tot), stats(3 * B_tot * W), gstats(2 * B_tot), loss(1);
  for (auto& v : pmax) v = 8.f * frand();
  for (auto& v : psum) v = 1.f + (float)(rng() % 7);
  for (auto& v : tgt) v = frand();
  std::vector<int64_t> tl(B_tot);
  for (int r = 0; r < W; ++r) {
    for (int i = 0; i < B_tot; ++i) tl[i] = (i % W == r) ? i : -1;
    emu_head_local_stats(pmax.data(), psum.data(), tgt.data(), tl.data(), n_blocks, B_tot, stats.data() + (size_t)r * 3 * B_tot);
  }
  emu_head_merge_stats(stats.data(), W, B_tot, gstats.data(), loss.data());
  std::vector<float> cosm(6 * 11), dl(6 * 11, 1.f);
  for (auto& v : cosm) v = 0.9f * std::tanh(frand());
  std::vector<float> cos0 = cosm;
  std::vector<int64_t> lab = {3, -1, 10, 0, -1, 7};
  emu_margin_fwd(cosm.data(), lab.data(), 6, 11, 11, 0, 64.f, 0.5f, 1.2f, 0.1f);
  emu_margin_bwd(dl.data(), cos0.data(), lab.data(), 6, 11, 11, 1, 64.f, 0.4f, 0.f, 0.f);
  return !(loss[0] == loss[0]);
}

// Seeded defects: the test suite checks that the sanitizers DO report them (a detector that never fires proves nothing).
static void racy_kernel(float* out) {            // a reduction that forgot its __syncthreads
  __shared__ float buf[64];
  buf[threadIdx.x] = (float)threadIdx.x;
  if (threadIdx.x == 0) { float s = 0.f; for (int i = 0; i < 64; ++i) s += buf[i]; out[0] = s; }
}
static void oob_kernel(const float* in, float* out, int n) {      // the classic missing tail guard
  out[threadIdx.x] = in[threadIdx.x + 1];
  (void)n;
}

static void misaligned_kernel(const float* in, uint4* out) {      // a 128-bit access on a pointer that is only 4-byte aligned
  if (threadIdx.x == 0) out[0] = ld_stream(in + 1);
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--seed-misaligned") {
    std::vector<float> in(16);
    std::vector<uint4> out(1);
    emu_launch(dim3(1), 32, [&] { misaligned_kernel(in.data(), out.data()); });
    return 0;
  }
  if (argc > 1 && std::string(argv[1]) == "--seed-race") {
    std::vector<float> out(1);
    emu_launch(dim3(1), 64, [&] { racy_kernel(out.data()); });
    return 0;
  }
  if (argc > 1 && std::string(argv[1]) == "--seed-oob") {
    std::vector<float> in(64), out(64);
    emu_launch(dim3(1), 64, [&] { oob_kernel(in.data(), out.data(), 64); });
    return 0;
  }
  int rc = 0;
  rc |= run_seg<float, 2>(2, 1353, 3, 0);
  rc |= run_seg<__nv_bfloat16, 3>(2, 1030, 4, 1);
  rc |= run_seg<float, 4>(2, 77, 2, 1);
  rc |= run_sgd<1>(20, 9, true);
  rc |= run_sgd<4>(11, 11, false);
  rc |= run_bn<float, false, false>(162, 32, 3, 5);
  rc |= run_bn<__nv_bfloat16, true, true>(162, 32, 3, 5);
  rc |= run_bn<__nv_bfloat16, false, true>(40, 64, 2, 3);
  rc |= run_bn<float, true, false>(53, 64, 4, 2);
  rc |= run_bn_chain<__nv_bfloat16>(70, 32, 2, 3, 2);
  rc |= run_sgd_flat(4 * 600, 2, true, true);
  rc |= run_peer<__nv_bfloat16, 1>(8 * 300 + 5, 2);
  rc |= run_peer<float, 0>(4 * 129 + 3, 1);
  rc |= run_gate<float, 1, 3>(false);
  rc |= run_gate<__nv_bfloat16, 0, 0>(false);
  rc |= run_gate<__nv_bfloat16, 1, 2>(true);
  rc |= run_mask<float>(1, 4, 3);
  rc |= run_mask<__nv_bfloat16>(64, 7, 5);
  rc |= run_small();
  rc |= run_pfc(4100, 1200, 200);                                // two CTAs, ragged second tile
  rc |= run_pfc(4096, 40, 200);                                   // positives outnumber num_sample
  rc |= run_cat<float>(301, 64, 18, 1);
  rc |= run_cat<__nv_bfloat16>(130, 128, 18, 2);
  printf("emulated kernels ran to completion, rc=%d\n", rc);
  return rc;
}
