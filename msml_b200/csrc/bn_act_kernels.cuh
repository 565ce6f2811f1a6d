// K-N  fused BatchNorm (+ residual add) (+ PReLU), NHWC, forward and backward: the kernels (launchers: bn_act.cu;
//      tests/emu runs this header under the CPU emulation and its sanitizers).
//   SURVEY.md 8(f)-1 "next" row: the BN / PReLU chains of the iResNet unit
//   (ref backbones/frb/iresnet.py:56-67, backbones/osb/unet.py:80-91) and of the FM bottlenecks
//   (ref backbones/fm/fmoperator.py:52-68):      y = prelu( bn(x) [+ res] )
//
// In the reference these are 3-6 separate ATen kernels per layer (batch_norm statistics, transform,
// prelu, add; and five more in backward) and make up ~60 % of the training step on B200.  A layer's
// tensors are 6-50 MB, i.e. 1-8 us of HBM time, so the chain is bound by launch / dependency latency, not
// bandwidth.  Each direction is three phases:
//   forward   phase 1  read x -> per-CTA (mean, M2) slab statistics
//             phase 2  Chan merge of the slabs (one CTA per channel), running stats, scale / shift
//             phase 3  re-read the SAME slab newest-first (an activation of <= 51 MB is still in the
//                      126 MB L2) [+ res], write y = prelu(x*scale + shift [+ res])
//   backward  phase 1  read dy, x [, res] -> per-CTA partials of (sum du, sum du*x, sum dy*u*[u<=0])
//             phase 2  dgamma, dbeta, dprelu and the coefficients of dx
//             phase 3  re-read the slab, write dx [, dres = du] [+ the skip branch's gradient]
// launched as THREE PLAIN KERNELS compiled per phase (template PHASE = 1, 2, 3), phases 2 and 3 with
// programmatic dependent launch so that their launch latency and prologue hide under the predecessor's tail.
// Measured inside CUDA graphs this beats both alternatives that were built and timed:
//   * ONE cooperative launch with two grid barriers (PHASE = 0, still selectable with MSML_BN_FUSED=1): a kernel
//     boundary costs ~2 us in a graph, a grid barrier 2-5 us (arrival skew of 300-600 CTAs + same-address
//     atomics) on top of the dearer cooperative launch, and the fused kernel carries the register budget of
//     its heaviest phase (205 MB layer: 147 vs 116 us forward, 349 vs 197 us backward);
//   * fp64 atomics into per-channel accumulators with a single barrier: same-address atomics from 300-600
//     CTAs cost more than the second barrier.
// Minimum HBM traffic is one read of each input + one write of each output (the re-reads hit L2: for slabs of <= 64 MB the
// statistics pass loads with an L2 evict_last policy and the apply pass with evict_first (BnGeom.skip & 32), so that other
// streams' data — the weight-gradient convolutions on the side stream — are evicted before the slab between the two passes).
// All phases stream a (P = N*H*W) x C matrix with C contiguous: a thread owns one 16-byte channel
// vector for its whole life, so per-channel coefficients live in registers; loads are 128-bit, several
// rows in flight per thread; each streaming phase runs exactly one resident wave of its own kernel.
// Eval mode (running statistics) is a coefficient kernel + one apply pass.
#pragma once
#ifndef MSML_CPU_EMU
#include <cooperative_groups.h>
#endif

#include "common.cuh"

namespace cg = cooperative_groups;      // under MSML_CPU_EMU tests/emu/cuda_emu.h supplies a stub (phases run as separate launches there)

namespace msml {

constexpr int kBnThreads = 256;
constexpr int kBnMaxCtas = 148 * 4;

struct BnGeom {
  int64_t P;       // rows (N*H*W)
  int C;           // channels (contiguous)
  int vpr;         // 16-byte vectors per row
  int rows_per_pass;   // kBnThreads / vpr
  int skip;            // bitmask: 1, 2, 4 skip the work of phase 1, 2, 3 (split launches / debug); 8 = timestamps (debug);
                       // 16 = no grid barriers (the phases run as separate plain launches);
                       // 32 = L2 residency hints: the statistics pass tags the slab evict_last, the apply pass evict_first
  int G;               // CTAs of the streaming phases (1 and 3): partial layout and slab geometry
};

// VN consecutive per-channel values as 16-byte loads (a strided scalar load per element costs one 32-byte sector each)
template <int VN>
__device__ __forceinline__ void ld_coef(const float* p, float* out) {
#pragma unroll
  for (int i = 0; i < VN; i += 4) {
    const float4 v = *reinterpret_cast<const float4*>(p + i);
    out[i] = v.x; out[i + 1] = v.y; out[i + 2] = v.z; out[i + 3] = v.w;
  }
}
template <int VN>
__device__ __forceinline__ void st_coef(float* p, const float* v, bool accumulate) {
#pragma unroll
  for (int i = 0; i < VN; i += 4) {
    float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    if (accumulate) {
      const float4 old = *reinterpret_cast<const float4*>(p + i);
      o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
    }
    *reinterpret_cast<float4*>(p + i) = o;
  }
}

// Lanes of a warp that own the same channel vector (vpr < 32) are folded with shuffles.
template <int VN>
__device__ __forceinline__ void fold_lanes(float* v, int vpr) {
  for (int o = vpr; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < VN; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
}
// After fold_lanes, a CTA holds `nslots` partial vectors per channel: one per warp (vpr < 32) or one per row lane.
struct FoldSlots {
  int slot, nslots;
  bool writer;
};
__device__ __forceinline__ FoldSlots fold_slots(int vpr, int rl) {
  FoldSlots f;
  if (vpr < 32) { f.slot = threadIdx.x >> 5; f.nslots = kBnThreads / 32; f.writer = (int)(threadIdx.x & 31) < vpr; }
  else { f.slot = rl; f.nslots = kBnThreads / vpr; f.writer = true; }
  return f;
}

// Chan et al. parallel merge of (n, mean, M2) triples (fp32: the merge is well conditioned).
__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float mb, float m2b) {
  if (nb <= 0.f) return;
  const float tot = n + nb, delta = mb - mean;
  const float f = nb / tot;
  mean = fmaf(delta, f, mean);
  m2 += m2b + delta * delta * n * f;
  n = tot;
}
// CTAs without phase-2 work would otherwise spin on the barrier's L2 line while the few busy CTAs read their
// partials through the same L2: let them sleep first.
__device__ __forceinline__ void idle_before_barrier(bool idle) {
  if (idle) __nanosleep(1500);
}

// debug attribution (MSML_BN_SKIP_PHASES & 8): CTA 0 stamps %globaltimer at the pass boundaries into the state tail
__device__ __forceinline__ void dbg_stamp(const int skip, float* coef, int C, int idx) {
  if ((skip & 8) && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t = 0;
#ifndef MSML_CPU_EMU
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
#endif
    reinterpret_cast<unsigned long long*>(coef + 3 * C)[idx] = t;
  }
}

// Programmatic dependent launch: phases 2 and 3 are launched while their predecessor still runs (launch latency and the
// prologue overlap its tail) and block here until it has completed and its writes are visible.
#ifndef MSML_CPU_EMU
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else   // emulated launches are already serialised
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}
#endif

// rows [r0, r1) of this CTA; thread-local row index k maps to row r0 + rl + k * rows_per_pass
struct Slab {
  int64_t r0, r1;
  int n_it;        // rows owned by this thread
};
__device__ __forceinline__ Slab slab_of(const BnGeom& g, int rl) {
  Slab s;
  const int64_t rows_per_cta = (g.P + g.G - 1) / g.G;
  s.r0 = (int64_t)blockIdx.x * rows_per_cta;
  s.r1 = s.r0 + rows_per_cta;
  if (s.r1 > g.P) s.r1 = g.P;
  const int64_t span = s.r1 - s.r0 - rl;
  s.n_it = span > 0 ? (int)((span + g.rows_per_pass - 1) / g.rows_per_pass) : 0;
  return s;
}

// Tail of a statistics pass: fold a thread's (sum, sum of squares) over its rows into the CTA's slab (mean, M2) and store
// them as partial `blockIdx.x` of the [k][C][stride] layout that phase 2 merges.
template <int VN>
__device__ __forceinline__ void slab_stats_store(float* s, float* q, float (*red)[kBnThreads * 8], const BnGeom& g, const Slab& sl,
                                                 int cv, int rl, float* part, float* part_n, int stride) {
  fold_lanes<VN>(s, g.vpr);
  fold_lanes<VN>(q, g.vpr);
  const FoldSlots fs = fold_slots(g.vpr, rl);
  if (fs.writer) {
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      red[0][fs.slot * g.C + cv * VN + i] = s[i];
      red[1][fs.slot * g.C + cv * VN + i] = q[i];
    }
  }
  __syncthreads();
  const float n = (float)(sl.r1 > sl.r0 ? sl.r1 - sl.r0 : 0);
  for (int c = threadIdx.x; c < g.C; c += kBnThreads) {
    float ss = 0.f, qq = 0.f;
    for (int r = 0; r < fs.nslots; ++r) { ss += red[0][r * g.C + c]; qq += red[1][r * g.C + c]; }
    const float mean = n > 0.f ? ss / n : 0.f;
    part[(size_t)c * stride + blockIdx.x] = mean;
    part[(size_t)(g.C + c) * stride + blockIdx.x] = fmaxf(qq - ss * mean, 0.f);   // M2 of this slab
  }
  if (threadIdx.x == 0) part_n[blockIdx.x] = n;
}

// Slab loads of the two streaming passes always carry an L2 policy (ONE code path: a second, unhinted path cost registers and
// spilled): evict_normal — the default behaviour — unless the residency hint (BnGeom.skip & 32) is on, in which case the
// statistics pass keeps (evict_last) and the apply pass drops (evict_first).  The policy is block-uniform.
template <int PHASE>
__device__ __forceinline__ uint64_t slab_policy(int skip) {
  const uint64_t on = PHASE == 1 ? l2_policy_keep() : l2_policy_drop();
  return (skip & 32) ? on : l2_policy_normal();
}
__device__ __forceinline__ uint4 ld_slab(const uint4* p, uint64_t policy) { return ld_stream_hint(p, policy); }

// ------------------------------------------------------------------------------------------- forward (training)
// part layout: [k][C][grid] (channel-major so that phase 2 reads the slabs of a channel coalesced)
// NEXT (phase 3 of the split launches only): y feeds another BatchNorm next (bn3 + skip of one residual unit -> bn1 of the
// following one, ref iresnet.py:56-67), so the apply pass also leaves the slab statistics of the y it writes (as rounded to T,
// i.e. exactly what that BN would read) in the NEXT op's workspace, with partial stride kBnMaxCtas and zero counts behind
// its own grid; that op then starts at phase 2 (msml_bn_fwd_ex stats_ready) and never reads y for statistics.
template <typename T, bool RES, bool PRELU, int PHASE, bool NEXT = false>   // PHASE 0: all three phases with grid barriers (cooperative launch);
__global__ void __launch_bounds__(kBnThreads, 4)            // 1 / 2 / 3: that phase only (plain launch, compiled on its own);
                                                            // 4 CTAs / SM: the 592-CTA grid is exactly one wave
bn_fwd_fused_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ prelu, float* __restrict__ running_mean,
                    float* __restrict__ running_var, long long* __restrict__ nbt, float momentum, float eps,
                    float* __restrict__ save_mean, float* __restrict__ save_invstd, float* part, float* part_n, float* coef,
                    BnGeom g, float* next_part, float* next_part_n) {
  static_assert(!NEXT || PHASE == 3, "next-op statistics are emitted by the split phase-3 kernel only");
  constexpr int VN = Vec<T>::N;
  __shared__ float red[2][kBnThreads * 8];   // [2][nslots * C] <= [2][256*8]
  cg::grid_group grid = cg::this_grid();
  (void)grid;
  const int G = g.G;
  const int cv = threadIdx.x % g.vpr, rl = threadIdx.x / g.vpr;
  const Slab sl = slab_of(g, rl);
  const uint4* xv = reinterpret_cast<const uint4*>(x) + cv;
  const int64_t row0 = sl.r0 + rl;
  const uint64_t pol = slab_policy<PHASE>(g.skip);
  dbg_stamp(g.skip, coef, g.C, 0);
  if (PHASE == 1 || PHASE == 2) pdl_launch_dependents();
  if (PHASE != 0) pdl_wait();      // phase 1 too: it may be launched as a programmatic dependent of whatever kernel precedes it in the
                                   // stream (a no-op when it was not), so nothing below may be read before this point

  // ---- phase 1: slab statistics
  if (PHASE == 0 ? !(g.skip & 1) : PHASE == 1) {
    float s[VN], q[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) { s[i] = 0.f; q[i] = 0.f; }
    int k = 0;
    for (; k + 4 <= sl.n_it; k += 4) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld_slab(xv + (row0 + (int64_t)(k + u) * g.rows_per_pass) * g.vpr, pol);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[VN];
        Vec<T>::unpack(v[u], f);
#pragma unroll
        for (int i = 0; i < VN; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
      }
    }
    for (; k < sl.n_it; ++k) {
      float f[VN];
      Vec<T>::unpack(ld_slab(xv + (row0 + (int64_t)k * g.rows_per_pass) * g.vpr, pol), f);
#pragma unroll
      for (int i = 0; i < VN; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
    }
    slab_stats_store<VN>(s, q, red, g, sl, cv, rl, part, part_n, G);
  }
  dbg_stamp(g.skip, coef, g.C, 1);
  if (PHASE == 0) grid.sync();
  dbg_stamp(g.skip, coef, g.C, 2);

  // ---- phase 2: merge the slabs, one CTA per channel (all loads issued up front, then a shuffle / smem tree)
  if (PHASE == 0 ? !(g.skip & 2) : PHASE == 2) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0 && nbt) *nbt += 1;
    const float Pf = (float)g.P;
    constexpr int NL = (kBnMaxCtas + kBnThreads - 1) / kBnThreads;
    for (int c = blockIdx.x; c < g.C; c += gridDim.x) {
      // thread 0 issues its (cold, HBM) parameter loads first so that their latency overlaps the merge
      float p_g = 1.f, p_b = 0.f, p_rm = 0.f, p_rv = 0.f;
      if (threadIdx.x == 0) {
        if (gamma) p_g = gamma[c];
        if (beta) p_b = beta[c];
        if (running_mean) p_rm = running_mean[c];
        if (running_var) p_rv = running_var[c];
      }
      const float* pm = part + (size_t)c * G;
      const float* pq = part + (size_t)(g.C + c) * G;
      float ln[NL], lm[NL], lq[NL];
#pragma unroll
      for (int j = 0; j < NL; ++j) {
        const int b = threadIdx.x + j * kBnThreads;
        ln[j] = b < G ? part_n[b] : 0.f;
        lm[j] = b < G ? pm[b] : 0.f;
        lq[j] = b < G ? pq[b] : 0.f;
      }
      float n = 0.f, mean = 0.f, m2 = 0.f;
#pragma unroll
      for (int j = 0; j < NL; ++j) chan_merge(n, mean, m2, ln[j], lm[j], lq[j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mean, o),
                    m2b = __shfl_xor_sync(0xffffffffu, m2, o);
        chan_merge(n, mean, m2, nb, mb, m2b);
      }
      __syncthreads();                      // red[] is free (phase 1 / previous channel done)
      if (lane == 0) { red[0][warp * 3] = n; red[0][warp * 3 + 1] = mean; red[0][warp * 3 + 2] = m2; }
      __syncthreads();
      if (threadIdx.x == 0) {
        n = 0.f; mean = 0.f; m2 = 0.f;
#pragma unroll
        for (int w = 0; w < kBnThreads / 32; ++w) chan_merge(n, mean, m2, red[0][w * 3], red[0][w * 3 + 1], red[0][w * 3 + 2]);
        const float var = m2 / Pf;                       // biased: normalisation
        const float invstd = rsqrtf(var + eps);
        save_mean[c] = mean;
        save_invstd[c] = invstd;
        if (running_mean) running_mean[c] = (1.f - momentum) * p_rm + momentum * mean;
        if (running_var) running_var[c] = (1.f - momentum) * p_rv + momentum * (Pf > 1.f ? m2 / (Pf - 1.f) : var);
        const float sc = p_g * invstd;
        coef[c] = sc;
        coef[g.C + c] = p_b - mean * sc;
      }
    }
    idle_before_barrier(blockIdx.x >= g.C);
  }
  dbg_stamp(g.skip, coef, g.C, 3);
  if (PHASE == 0) grid.sync();
  dbg_stamp(g.skip, coef, g.C, 4);

  // ---- phase 3: apply over the same slab, newest rows first (they are the likeliest L2 hits)
  if (PHASE == 0 ? !(g.skip & 4) : PHASE == 3) {
    float sc[VN], sh[VN], pa[VN];
    ld_coef<VN>(coef + cv * VN, sc);
    ld_coef<VN>(coef + g.C + cv * VN, sh);
    if (PRELU) ld_coef<VN>(prelu + cv * VN, pa);
    constexpr int U = RES ? 2 : 4;
    const uint4* rv = reinterpret_cast<const uint4*>(res) + cv;
    uint4* yv = reinterpret_cast<uint4*>(y) + cv;
    float ns[VN], nq[VN];                    // NEXT: sum / sum of squares of the y this thread writes
#pragma unroll
    for (int i = 0; i < VN; ++i) { ns[i] = 0.f; nq[i] = 0.f; }
    for (int k = sl.n_it - 1; k >= 0; k -= U) {
      uint4 a[U], b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k - u >= 0) {
          const int64_t v = (row0 + (int64_t)(k - u) * g.rows_per_pass) * g.vpr;
          a[u] = ld_slab(xv + v, pol);
          if (RES) b[u] = ld_stream(rv + v);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k - u < 0) break;
        float f[VN], r[VN], o[VN];
        Vec<T>::unpack(a[u], f);
        if (RES) Vec<T>::unpack(b[u], r);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          float t = fmaf(f[i], sc[i], sh[i]);
          if (RES) t += r[i];
          o[i] = PRELU ? (t > 0.f ? t : t * pa[i]) : t;
        }
        const uint4 packed = Vec<T>::pack(o);
        yv[(row0 + (int64_t)(k - u) * g.rows_per_pass) * g.vpr] = packed;   // re-read by the next conv: default policy
        if (NEXT) {
          Vec<T>::unpack(packed, o);
#pragma unroll
          for (int i = 0; i < VN; ++i) { ns[i] += o[i]; nq[i] = fmaf(o[i], o[i], nq[i]); }
        }
      }
    }
    if (NEXT) {
      slab_stats_store<VN>(ns, nq, red, g, sl, cv, rl, next_part, next_part_n, kBnMaxCtas);
      if (blockIdx.x == 0)
        for (int b = (int)gridDim.x + (int)threadIdx.x; b < kBnMaxCtas; b += kBnThreads) next_part_n[b] = 0.f;
    }
  }
  dbg_stamp(g.skip, coef, g.C, 5);
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
// u = x*sc + sh [+ res] (sc = gamma*invstd, sh = beta - mean*sc);  du = dy * (u > 0 ? 1 : a);  xhat = (x - mean)*invstd
// accumulators [3][C]:  [0] T0 = sum du   [1] Q1 = sum du * x   [2] sum dy * u * [u <= 0]
//   sum du * xhat = invstd * (Q1 - mean * T0): the streaming loops carry as few per-channel vectors as possible
//   (registers decide how many CTAs an SM holds, i.e. how many loads are in flight)
// dx = A*(du - B - xhat*G) = sc*du + x*K1 + K0   with A = sc, B = T0 / P, G = sum du*xhat / P,
//   K1 = -A*G*invstd, K0 = -A*B + A*G*mean*invstd      (B = G = 0 in eval mode)
template <typename T, bool RES, bool PRELU, int PHASE>
__global__ void __launch_bounds__(kBnThreads, PHASE == 0 ? 1 : ((RES && PRELU) ? 2 : 3))   // <= 80 registers without spilling
bn_bwd_fused_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ res,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ prelu, T* __restrict__ dx, T* __restrict__ dres,
                    const T* __restrict__ dadd, float* dgamma, float* dbeta, float* dprelu, int training, int accumulate,
                    float* part, float* coef, BnGeom g) {
  constexpr int VN = Vec<T>::N;
  constexpr bool R3 = RES && PRELU;          // the residual is only needed to recover the sign of u
  constexpr int NA = PRELU ? 3 : 2;
  __shared__ float red[NA][kBnThreads * 8];
  cg::grid_group grid = cg::this_grid();
  (void)grid;
  const int G = g.G;
  const int cv = threadIdx.x % g.vpr, rl = threadIdx.x / g.vpr;
  const Slab sl = slab_of(g, rl);
  const int64_t row0 = sl.r0 + rl;
  const uint4* dyv = reinterpret_cast<const uint4*>(dy) + cv;
  const uint4* xv = reinterpret_cast<const uint4*>(x) + cv;
  const uint4* rv = reinterpret_cast<const uint4*>(res) + cv;

  const uint64_t pol = slab_policy<PHASE>(g.skip);
  float sc[VN], sh[VN], pa[VN];
  if (PHASE == 1) pdl_wait();        // as a programmatic dependent of an arbitrary predecessor: wait before the first global read.
                                     // (Phase 3 loads the coefficients below BEFORE its wait: they were written by the forward pass /
                                     // the optimizer, never by this op's phases 1-2, and the loads hide under phase 2.)
  if (PHASE != 2) {
    float is[VN], mu[VN], ga[VN], be[VN];
    ld_coef<VN>(invstd + cv * VN, is);
    ld_coef<VN>(mean + cv * VN, mu);
    if (gamma) ld_coef<VN>(gamma + cv * VN, ga);
    if (beta) ld_coef<VN>(beta + cv * VN, be);
    if (PRELU) ld_coef<VN>(prelu + cv * VN, pa);
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      sc[i] = (gamma ? ga[i] : 1.f) * is[i];
      sh[i] = (beta ? be[i] : 0.f) - mu[i] * sc[i];
      if (!PRELU) pa[i] = 1.f;
    }
  }
  dbg_stamp(g.skip, coef, g.C, 0);
  if (PHASE == 1 || PHASE == 2) pdl_launch_dependents();
  if (PHASE != 0) pdl_wait();      // phase 1 too: it may be launched as a programmatic dependent of whatever kernel precedes it in the
                                   // stream (a no-op when it was not), so nothing below may be read before this point

  // ---- phase 1: slab reductions
  if (PHASE == 0 ? !(g.skip & 1) : PHASE == 1) {
    float s0[VN], s1[VN], s2[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) s0[i] = s1[i] = s2[i] = 0.f;
    constexpr int U = 2;
    for (int k = 0; k < sl.n_it; k += U) {
      uint4 a[U], b[U], c4[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k + u < sl.n_it) {
          const int64_t v = (row0 + (int64_t)(k + u) * g.rows_per_pass) * g.vpr;
          a[u] = ld_slab(dyv + v, pol);
          b[u] = ld_slab(xv + v, pol);
          if (R3) c4[u] = ld_slab(rv + v, pol);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k + u >= sl.n_it) break;
        float d[VN], f[VN], rs[VN];
        Vec<T>::unpack(a[u], d);
        Vec<T>::unpack(b[u], f);
        if (R3) Vec<T>::unpack(c4[u], rs);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          float du = d[i];
          if (PRELU) {
            float uu = fmaf(f[i], sc[i], sh[i]);
            if (RES) uu += rs[i];
            if (!(uu > 0.f)) { s2[i] = fmaf(d[i], uu, s2[i]); du *= pa[i]; }
          }
          s0[i] += du;
          s1[i] = fmaf(du, f[i], s1[i]);           // Q1 = sum du * x
        }
      }
    }
    fold_lanes<VN>(s0, g.vpr);
    fold_lanes<VN>(s1, g.vpr);
    if (PRELU) fold_lanes<VN>(s2, g.vpr);
    const FoldSlots fs = fold_slots(g.vpr, rl);
    if (fs.writer) {
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        red[0][fs.slot * g.C + cv * VN + i] = s0[i];
        red[1][fs.slot * g.C + cv * VN + i] = s1[i];
        if (PRELU) red[NA - 1][fs.slot * g.C + cv * VN + i] = s2[i];
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < g.C; c += kBnThreads) {
      float t0 = 0.f, t1 = 0.f, t2 = 0.f;
      for (int r = 0; r < fs.nslots; ++r) {
        t0 += red[0][r * g.C + c]; t1 += red[1][r * g.C + c];
        if (PRELU) t2 += red[NA - 1][r * g.C + c];
      }
      part[(size_t)c * G + blockIdx.x] = t0;
      part[(size_t)(g.C + c) * G + blockIdx.x] = t1;
      if (PRELU) part[(size_t)(2 * g.C + c) * G + blockIdx.x] = t2;
    }
  }
  dbg_stamp(g.skip, coef, g.C, 1);
  if (PHASE == 0) grid.sync();
  dbg_stamp(g.skip, coef, g.C, 2);

  // ---- phase 2: dgamma, dbeta, dprelu, coefficients; one CTA per channel
  if (PHASE == 0 ? !(g.skip & 2) : PHASE == 2) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float Pf = (float)g.P;
    for (int c = blockIdx.x; c < g.C; c += gridDim.x) {
      float t0 = 0.f, t1 = 0.f, t2 = 0.f;
      for (int b = threadIdx.x; b < G; b += kBnThreads) {
        t0 += part[(size_t)c * G + b];
        t1 += part[(size_t)(g.C + c) * G + b];
        if (PRELU) t2 += part[(size_t)(2 * g.C + c) * G + b];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        t0 += __shfl_xor_sync(0xffffffffu, t0, o);
        t1 += __shfl_xor_sync(0xffffffffu, t1, o);
        t2 += __shfl_xor_sync(0xffffffffu, t2, o);
      }
      __syncthreads();
      if (lane == 0) { red[0][warp * 3] = t0; red[0][warp * 3 + 1] = t1; red[0][warp * 3 + 2] = t2; }
      __syncthreads();
      if (threadIdx.x == 0) {
        t0 = t1 = t2 = 0.f;
#pragma unroll
        for (int w = 0; w < kBnThreads / 32; ++w) { t0 += red[0][w * 3]; t1 += red[0][w * 3 + 1]; t2 += red[0][w * 3 + 2]; }
        const float is_c = invstd[c], mu_c = mean[c];
        const float s1 = is_c * (t1 - mu_c * t0);          // sum du * xhat
        if (accumulate) {       // write straight into the parameters' .grad (flat gradient buffer)
          if (dbeta) dbeta[c] += t0;
          if (dgamma) dgamma[c] += s1;
          if (PRELU && dprelu) dprelu[c] += t2;
        } else {
          if (dbeta) dbeta[c] = t0;
          if (dgamma) dgamma[c] = s1;
          if (PRELU && dprelu) dprelu[c] = t2;
        }
        const float A = (gamma ? gamma[c] : 1.f) * is_c;
        const float Bc = training ? t0 / Pf : 0.f;         // eval: statistics are constants
        const float Gc = training ? s1 / Pf : 0.f;
        coef[c] = -A * Gc * is_c;                          // K1
        coef[g.C + c] = A * (Gc * mu_c * is_c - Bc);       // K0
      }
    }
    idle_before_barrier(blockIdx.x >= g.C);
  }
  dbg_stamp(g.skip, coef, g.C, 3);
  if (PHASE == 0) grid.sync();
  dbg_stamp(g.skip, coef, g.C, 4);

  // ---- phase 3: dx [, dres] over the same slab, newest rows first
  if (PHASE == 0 ? !(g.skip & 4) : PHASE == 3) {
    float K1[VN], K0[VN];
    ld_coef<VN>(coef + cv * VN, K1);
    ld_coef<VN>(coef + g.C + cv * VN, K0);
    uint4* dxv = reinterpret_cast<uint4*>(dx) + cv;
    uint4* drv = reinterpret_cast<uint4*>(dres) + cv;
    const uint4* dav = reinterpret_cast<const uint4*>(dadd) + cv;     // gradient of the other consumer of x (skip branch)
    const bool has_add = dadd != nullptr;
    constexpr int U = 2;
    for (int k = sl.n_it - 1; k >= 0; k -= U) {
      uint4 a[U], b[U], c4[U], e[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k - u >= 0) {
          const int64_t v = (row0 + (int64_t)(k - u) * g.rows_per_pass) * g.vpr;
          a[u] = ld_slab(dyv + v, pol);
          b[u] = ld_slab(xv + v, pol);
          if (R3) c4[u] = ld_slab(rv + v, pol);
          if (has_add) e[u] = ld_stream(dav + v);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k - u < 0) break;
        const int64_t v = (row0 + (int64_t)(k - u) * g.rows_per_pass) * g.vpr;
        float d[VN], f[VN], rs[VN], o[VN], dr[VN];
        Vec<T>::unpack(a[u], d);
        Vec<T>::unpack(b[u], f);
        if (R3) Vec<T>::unpack(c4[u], rs);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          float du = d[i];
          if (PRELU) {
            float uu = fmaf(f[i], sc[i], sh[i]);
            if (RES) uu += rs[i];
            if (!(uu > 0.f)) du *= pa[i];
          }
          dr[i] = du;
          o[i] = fmaf(sc[i], du, fmaf(f[i], K1[i], K0[i]));
        }
        if (has_add) {
          float ad[VN];
          Vec<T>::unpack(e[u], ad);
#pragma unroll
          for (int i = 0; i < VN; ++i) o[i] += ad[i];
        }
        dxv[v] = Vec<T>::pack(o);
        if (R3) drv[v] = Vec<T>::pack(dr);
      }
    }
  }
  dbg_stamp(g.skip, coef, g.C, 5);
}

}  // namespace msml
