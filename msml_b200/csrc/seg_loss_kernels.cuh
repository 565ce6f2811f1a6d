// K-S  structure-via-consensus segmentation criterion, forward and backward: the kernels (launchers: seg_loss.cu).
//   ref tricks/consensus_loss.py:63-178 (StructureConsensuLossFunction), called as
//   `seg_criterion(final_seg, msk, msk)` with (alpha, beta) = (10, 5), 'idx', 'idx' (ref train.py:228-229,258); SURVEY.md 8f-4.
//
// For every blob id b with at least one pixel (M = [blobs == b], p = softmax over channels, cnt_n = #pixels of b in sample n):
//     t[n,c]   = sum_{x in b} p[n,c,x] / cnt_n            ('idx';  / (H*W) for 'all';  0 where cnt_n == 0)
//     y        = target at the first pixel of b in (n, h, w) order
//     loss_avg = mean_n [cnt_n > 0] * -log t[n,y]
//     kl       = sum_{n,c} ( t log t * nz[n,c] - t * sum_{x in b, p != 0} log p[n,c,x] ),   nz = #{x in b : p != 0}
//     L_b      = alpha * loss_avg + beta * kl / D,    D = sum nz ('idx')  or  N*H*W ('all');        loss = mean_b L_b
// The reference loops over blobs in Python with ~40 ATen kernels per blob over the (N,C,H,W) map.  Every quantity above is
// a function of three per-(b,n,c) sums, so the forward is ONE pass over the logits (each pixel contributes to the sums of
// its own blob) + a one-CTA finalize, and the backward is one elementwise pass:
//     dlogit[n,c,x] = gout * ( w_c - p_c * sum_k w_k ),   w_c = p_c * A[b,n,c] - [p_c != 0] * Tc[b,n,c]
// with A = (dL/dt)/den/nb and Tc = beta/D * t / nb written by the finalize (b = the pixel's blob, nb = number of blobs).
// Blob ids must be integers in [0, K), K <= 32 (-1 = pixel of no blob, skipped); any other id, labels that differ inside a
// blob (the reference asserts on that, ref :103) or a label outside [0, C) turn the loss into NaN (no host synchronisation).
// HBM-bound and tiny (final_seg is N x 2 x 112 x 112): minimum traffic = logits + blob ids read once per direction,
// gradient written once.  Partial sums are combined in a fixed order: the result is deterministic.
#include <cfloat>
#include <climits>

#pragma once
#include "common.cuh"

namespace msml {

constexpr int kSegThreads = 256;
constexpr int kSegPix = 4;                       // pixels per thread
constexpr int kSegChunk = kSegThreads * kSegPix; // pixels per CTA
constexpr int kSegMaxBlobs = 32;
constexpr int kSegMaxC = 4;

struct SegGeom {
  int64_t N, HW;
  int C, K;
  int chunks;            // CTAs per sample
  int64_t s_n, s_c, s_x; // element strides of the logits
};

template <typename T, int C>
__device__ __forceinline__ void seg_softmax(const T* __restrict__ z, const SegGeom& g, int64_t n, int64_t x, float* p) {
  float v[C], m = -FLT_MAX;
#pragma unroll
  for (int c = 0; c < C; ++c) { v[c] = to_f32(z[n * g.s_n + c * g.s_c + x * g.s_x]); m = fmaxf(m, v[c]); }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) { v[c] = expf(v[c] - m); s += v[c]; }     // expf, not __expf: denormals decide `p != 0`
#pragma unroll
  for (int c = 0; c < C; ++c) p[c] = v[c] / s;
}

// float slots per (n, chunk, b): S1[C] S2[C] nz[C] cnt;   int slots: first pixel, min label, max label
template <int C> struct SegSlots { static constexpr int F = 3 * C + 1, I = 3; };

template <typename T, int C>
__global__ void __launch_bounds__(kSegThreads)
seg_stats_kernel(const T* __restrict__ logit, const int64_t* __restrict__ blobs, const int64_t* __restrict__ target,
                 float* __restrict__ partf, int* __restrict__ parti, int* __restrict__ partbad, SegGeom g) {
  constexpr int F = SegSlots<C>::F;
  __shared__ float redf[kSegThreads / 32][F];
  __shared__ int redi[kSegThreads / 32][3];
  const int64_t n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float p[kSegPix][C];
  int bid[kSegPix], lab[kSegPix], px[kSegPix];
  int bad = 0;
#pragma unroll
  for (int j = 0; j < kSegPix; ++j) {
    const int64_t x = (int64_t)blockIdx.x * kSegChunk + j * kSegThreads + threadIdx.x;
    bid[j] = -1; lab[j] = 0; px[j] = INT_MAX;
    if (x < g.HW) {
      const int64_t b = blobs[n * g.HW + x];
      if (b == -1) { /* ignored pixel: member of no blob */ }
      else if (b < 0 || b >= g.K) { bad = 1; }
      else {
        bid[j] = (int)b;
        const int64_t t = target[n * g.HW + x];
        lab[j] = t < INT_MIN ? INT_MIN : (t > INT_MAX ? INT_MAX : (int)t);
        px[j] = (int)x;
        seg_softmax<T, C>(logit, g, n, x, p[j]);
      }
    }
  }
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) partbad[n * g.chunks + blockIdx.x] = bad;
  for (int b = 0; b < g.K; ++b) {
    float acc[F];
#pragma unroll
    for (int i = 0; i < F; ++i) acc[i] = 0.f;
    int first = INT_MAX, lmin = INT_MAX, lmax = INT_MIN;
#pragma unroll
    for (int j = 0; j < kSegPix; ++j) {
      if (bid[j] == b) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          acc[c] += p[j][c];
          if (p[j][c] != 0.f) { acc[C + c] += logf(p[j][c]); acc[2 * C + c] += 1.f; }
        }
        acc[3 * C] += 1.f;
        first = min(first, px[j]);
        lmin = min(lmin, lab[j]);
        lmax = max(lmax, lab[j]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < F; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
      first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
      lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    }
    __syncthreads();                 // the previous blob's slots have been consumed
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < F; ++i) redf[warp][i] = acc[i];
      redi[warp][0] = first; redi[warp][1] = lmin; redi[warp][2] = lmax;
    }
    __syncthreads();
    const size_t slot = ((size_t)n * g.chunks + blockIdx.x) * g.K + b;
    if (threadIdx.x < F) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kSegThreads / 32; ++w) s += redf[w][threadIdx.x];
      partf[slot * F + threadIdx.x] = s;
    } else if (threadIdx.x >= 32 && threadIdx.x < 35) {
      const int k = threadIdx.x - 32;
      int v = redi[0][k];
      for (int w = 1; w < kSegThreads / 32; ++w) v = (k == 2) ? max(v, redi[w][k]) : min(v, redi[w][k]);
      parti[slot * 3 + k] = v;
    }
  }
}

// deterministic block sum (double): result valid in every thread
__device__ __forceinline__ double seg_block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < kSegThreads / 32; ++w) s += sh[w];
  return s;
}
__device__ __forceinline__ long long seg_block_min(long long v, long long* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const long long u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  long long s = sh[0];
#pragma unroll
  for (int w = 1; w < kSegThreads / 32; ++w) s = sh[w] < s ? sh[w] : s;
  return s;
}

// acc layout [b][n][F]: S1 -> t, S2, nz, cnt.   coef layout [2][b][n][C]: A, Tc.
template <int C>
__global__ void __launch_bounds__(kSegThreads)
seg_finalize_kernel(const float* __restrict__ partf, const int* __restrict__ parti, const int* __restrict__ partbad,
                    const int64_t* __restrict__ target, float* __restrict__ acc, float* __restrict__ coef,
                    float* __restrict__ loss, float alpha, float beta, int pixel_all, int kl_all, SegGeom g) {
  constexpr int F = SegSlots<C>::F;
  __shared__ double shd[kSegThreads / 32];
  __shared__ long long shl[kSegThreads / 32];
  __shared__ double s_D[kSegMaxBlobs];
  __shared__ int s_y[kSegMaxBlobs], s_present[kSegMaxBlobs];
  const int K = g.K;
  const int64_t N = g.N;
  const float hw = (float)g.HW;
  // ---- combine the chunks of every (b, n) in a fixed order; t = blob mean
  long long poison = 0;
  for (int64_t i = threadIdx.x; i < (int64_t)K * N; i += kSegThreads) {
    const int b = (int)(i / N);
    const int64_t n = i - (int64_t)b * N;
    float a[F];
#pragma unroll
    for (int k = 0; k < F; ++k) a[k] = 0.f;
    for (int ch = 0; ch < g.chunks; ++ch) {
      const float* src = partf + (((size_t)n * g.chunks + ch) * K + b) * F;
#pragma unroll
      for (int k = 0; k < F; ++k) a[k] += src[k];
    }
    const float cnt = a[3 * C];
    const float den = pixel_all ? hw : cnt;
#pragma unroll
    for (int c = 0; c < C; ++c) a[c] = cnt > 0.f ? a[c] / den : 0.f;
    float* dst = acc + ((size_t)b * N + n) * F;
#pragma unroll
    for (int k = 0; k < F; ++k) dst[k] = a[k];
  }
  for (int64_t i = threadIdx.x; i < N * g.chunks; i += kSegThreads) poison |= partbad[i];
  __syncthreads();      // acc is read back below by other threads of this (single) CTA

  double total = 0.0;
  int nb = 0;
  for (int b = 0; b < K; ++b) {
    // first pixel of the blob in (n, x) order, label uniformity
    long long first = LLONG_MAX;
    int lmin = INT_MAX, lmax = INT_MIN;
    for (int64_t i = threadIdx.x; i < N * g.chunks; i += kSegThreads) {
      const int* pi = parti + ((size_t)i * K + b) * 3;
      if (pi[0] != INT_MAX) {
        const int64_t n = i / g.chunks;
        const long long key = (long long)n * g.HW + pi[0];
        first = key < first ? key : first;
        lmin = min(lmin, pi[1]);
        lmax = max(lmax, pi[2]);
      }
    }
    first = seg_block_min(first, shl);
    const long long lmin_all = seg_block_min((long long)lmin, shl);
    const long long lmax_all = -seg_block_min(-(long long)lmax, shl);
    const int present = first != LLONG_MAX;
    int y = 0;
    if (present) {
      const int64_t t = target[first];
      if (lmin_all != lmax_all || t < 0 || t >= C) poison = 1; else y = (int)t;
    }
    double s_avg = 0.0, s_nz = 0.0, s_kl = 0.0;
    if (present) {
      for (int64_t n = threadIdx.x; n < N; n += kSegThreads) {
        const float* a = acc + ((size_t)b * N + n) * F;
        if (a[3 * C] > 0.f) {
          float ty = a[0];
#pragma unroll
          for (int c = 1; c < C; ++c) if (c == y) ty = a[c];
          s_avg += (double)(-logf(ty));
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float t = a[c], s2 = a[C + c], nz = a[2 * C + c];
          s_nz += (double)nz;
          if (nz > 0.f) s_kl += (double)(t * logf(t) * nz - t * s2);
        }
      }
    }
    s_avg = seg_block_sum(s_avg, shd);
    s_nz = seg_block_sum(s_nz, shd);
    s_kl = seg_block_sum(s_kl, shd);
    const double D = kl_all ? (double)N * (double)g.HW : s_nz;
    if (present) { total += (double)alpha * s_avg / (double)N + (double)beta * s_kl / D; nb += 1; }
    if (threadIdx.x == 0) { s_D[b] = D; s_y[b] = y; s_present[b] = present; }
  }
  poison = -seg_block_min(-(poison != 0 ? 1LL : 0LL), shl);      // any thread poisoned -> all
  __syncthreads();
  if (threadIdx.x == 0) loss[0] = poison ? __int_as_float(0x7fc00000) : (float)(total / (double)nb);

  // ---- backward coefficients
  const float inv_nb = nb > 0 ? 1.f / (float)nb : 0.f;
  for (int64_t i = threadIdx.x; i < (int64_t)K * N; i += kSegThreads) {
    const int b = (int)(i / N);
    const int64_t n = i - (int64_t)b * N;
    const float* a = acc + ((size_t)b * N + n) * F;
    const float cnt = a[3 * C];
    const float den = pixel_all ? hw : cnt;
    const float bD = s_present[b] ? (float)((double)beta / s_D[b]) : 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float t = a[c], s2 = a[C + c], nz = a[2 * C + c];
      float dLdt = 0.f;
      if (s_present[b] && cnt > 0.f) {
        if (c == s_y[b]) dLdt += alpha / (float)N * (-1.f / t);
        if (nz > 0.f) dLdt += bD * ((logf(t) + 1.f) * nz - s2);
      }
      coef[((size_t)b * N + n) * C + c] = cnt > 0.f ? dLdt / den * inv_nb : 0.f;
      coef[((size_t)(K + b) * N + n) * C + c] = bD * t * inv_nb;
    }
  }
}

template <typename T, int C>
__global__ void __launch_bounds__(kSegThreads)
seg_bwd_kernel(const T* __restrict__ logit, const int64_t* __restrict__ blobs, const float* __restrict__ coef,
               const float* __restrict__ gout, T* __restrict__ dlogit, SegGeom g) {
  const int64_t n = blockIdx.y;
  const float go = gout ? gout[0] : 1.f;
#pragma unroll
  for (int j = 0; j < kSegPix; ++j) {
    const int64_t x = (int64_t)blockIdx.x * kSegChunk + j * kSegThreads + threadIdx.x;
    if (x >= g.HW) continue;
    const int64_t b = blobs[n * g.HW + x];
    float d[C];
    if (b < 0 || b >= g.K) {
#pragma unroll
      for (int c = 0; c < C; ++c) d[c] = 0.f;
    } else {
      float p[C], w[C], sw = 0.f;
      seg_softmax<T, C>(logit, g, n, x, p);
      const float* A = coef + ((size_t)b * g.N + n) * C;
      const float* Tc = coef + ((size_t)(g.K + b) * g.N + n) * C;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        w[c] = p[c] * __ldg(A + c) - (p[c] != 0.f ? __ldg(Tc + c) : 0.f);
        sw += w[c];
      }
#pragma unroll
      for (int c = 0; c < C; ++c) d[c] = go * (w[c] - p[c] * sw);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) dlogit[n * g.s_n + c * g.s_c + x * g.s_x] = from_f32<T>(d[c]);
  }
}

// ---- host-side geometry and workspace layout (shared by the launchers and the CPU emulation harness)
inline int seg_geom(int64_t N, int64_t C, int64_t HW, int64_t K, int channels_last, int dtype, SegGeom* g) {
  MSML_REQUIRE(N > 0 && HW > 0, MSML_EINVAL, "bad shape N=%lld HW=%lld", (long long)N, (long long)HW);
  MSML_REQUIRE(C >= 2 && C <= kSegMaxC, MSML_EUNSUPPORTED, "C=%lld: 2..%d classes are supported", (long long)C, kSegMaxC);
  MSML_REQUIRE(K >= 1 && K <= kSegMaxBlobs, MSML_EUNSUPPORTED, "K=%lld: 1..%d blob ids are supported", (long long)K, kSegMaxBlobs);
  MSML_REQUIRE(HW < INT_MAX && N < 65536, MSML_EUNSUPPORTED, "N=%lld HW=%lld too large", (long long)N, (long long)HW);
  MSML_REQUIRE(dtype == MSML_F32 || dtype == MSML_BF16 || dtype == MSML_F16, MSML_EINVAL, "unknown dtype %d", dtype);
  g->N = N; g->HW = HW; g->C = (int)C; g->K = (int)K;
  g->chunks = (int)((HW + kSegChunk - 1) / kSegChunk);
  g->s_n = C * HW;
  g->s_c = channels_last ? 1 : HW;
  g->s_x = channels_last ? C : 1;
  return 0;
}

// workspace: partf | acc | parti | partbad   (floats first: every region starts 4-byte aligned)
inline size_t seg_partf(const SegGeom& g) { return (size_t)g.N * g.chunks * g.K * (3 * g.C + 1); }
inline size_t seg_acc(const SegGeom& g) { return (size_t)g.K * g.N * (3 * g.C + 1); }
inline size_t seg_parti(const SegGeom& g) { return (size_t)g.N * g.chunks * g.K * 3; }
inline size_t seg_partbad(const SegGeom& g) { return (size_t)g.N * g.chunks; }

}  // namespace msml
