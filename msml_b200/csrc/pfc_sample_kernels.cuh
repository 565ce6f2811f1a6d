// K-D  PartialFC label remap + negative-class sampling: the kernels.      ref headers/partial_fc.py:77-94
//      (launchers: pfc_sample.cu; tests/emu runs this header under the CPU emulation: bit-exact sampling on the CPU test tier)
//
//   :79-81  remap           off-shard -> -1, on-shard -= class_start
//   :84     positive        (no torch.unique needed: the scatter below is idempotent)
//   :85-86  perm = rand(num_local); perm[positive] = 2.0      (rand stays in torch: same generator
//           consumption as the reference; the draw is an INPUT here)
//   :87-88  index = sort(topk(perm, num_sample).indices)
//   :89-90  else index = positive                              (n_pos > num_sample)
//   :92     labels = searchsorted(index, labels)
//
// topk+sort is replaced by an EXACT radix select of the k-th largest key (3 digit passes of
// 11/11/10 bits over an order-preserving uint32 image of the float) followed by an ordered
// compaction, which yields the sorted index list directly.  k = max(num_sample, n_pos) covers the
// :89-90 branch because positives are the only keys equal to 2.0 and rand() < 1.  Ties at the
// k-th value go to the lowest class index.  Everything is integer/compare work: bit-exact.
// Kernels are small multi-CTA passes; the digit pick / block-offset scan runs in the last CTA to
// finish (threadfence + ticket), so no kernel ever waits on another CTA.  Both are block-wide scans over all 512 threads:
// round 1 walked the 2048 bins (and the per-CTA counts) from ONE thread with dependent global loads, ~0.6 us each —
// 477 us of a 0.84 ms sampled step at 125,000 classes (gpurun r02_head_125k_sr0.1.json), the whole "sampled-path cliff".
#pragma once
#include "common.cuh"

namespace msml {

constexpr int kSelThreads = 512;
constexpr int kSelItems = 8;                       // consecutive elements per thread
constexpr int kSelTile = kSelThreads * kSelItems;  // 4096 elements per CTA
constexpr int kBins = 2048;

struct SelState {
  unsigned int hist[3][kBins];
  unsigned int ticket[4];
  unsigned int prefix;        // key bits decided so far
  unsigned int n_pos;         // #keys == key(2.0f)
  long long k;                // elements to select
  long long k_rem;            // still to take among keys matching prefix
  long long need_eq;          // after pass 2: how many keys == threshold to take
  unsigned int threshold;     // after pass 2: the k-th largest key
};

__device__ __forceinline__ unsigned int f2key(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// inclusive block scan of one long long per thread (Hillis-Steele in shared memory; every thread of the CTA must call it)
__device__ __forceinline__ long long block_incl_scan(long long v, long long* sh) {
  const int t = threadIdx.x;
  sh[t] = v;
  __syncthreads();
  for (int o = 1; o < kSelThreads; o <<= 1) {
    const long long add = t >= o ? sh[t - o] : 0;
    __syncthreads();
    sh[t] += add;
    __syncthreads();
  }
  const long long r = sh[t];
  __syncthreads();
  return r;
}

__device__ __forceinline__ int digit_shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }
__device__ __forceinline__ int digit_bits(int pass) { return pass == 2 ? 10 : 11; }

__global__ void pfc_remap_kernel(int64_t* __restrict__ tl, int64_t n, int64_t class_start, int64_t num_local) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t v = tl[i];
  tl[i] = (v >= class_start && v < class_start + num_local) ? v - class_start : -1;
}

__global__ void pfc_mark_kernel(float* __restrict__ perm, const int64_t* __restrict__ tl, int64_t n, int64_t num_local) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t v = tl[i];
  if (v >= 0 && v < num_local) perm[v] = 2.0f;
}

// One digit pass.  PASS 0 also counts positives and fixes k.
template <int PASS>
__global__ void __launch_bounds__(kSelThreads)
pfc_hist_kernel(const float* __restrict__ perm, int64_t n, int64_t num_sample, SelState* st) {
  __shared__ unsigned int sh[kBins];
  __shared__ unsigned int s_pos;
  __shared__ bool s_last;
  for (int i = threadIdx.x; i < kBins; i += kSelThreads) sh[i] = 0;
  if (threadIdx.x == 0) s_pos = 0;
  __syncthreads();
  const unsigned int prefix = PASS == 0 ? 0u : st->prefix;
  const int shift = digit_shift(PASS);
  const unsigned int hi_mask = PASS == 0 ? 0u : (0xffffffffu << (shift + digit_bits(PASS)));
  const unsigned int key_two = f2key(2.0f);
  unsigned int pos = 0;
  const int64_t base = (int64_t)blockIdx.x * kSelTile;
#pragma unroll
  for (int j = 0; j < kSelItems; ++j) {
    const int64_t i = base + (int64_t)j * kSelThreads + threadIdx.x;   // coalesced
    if (i < n) {
      const unsigned int key = f2key(perm[i]);
      if (PASS == 0 && key == key_two) ++pos;
      if ((key & hi_mask) == prefix) atomicAdd(&sh[(key >> shift) & ((1u << digit_bits(PASS)) - 1u)], 1u);
    }
  }
  if (PASS == 0 && pos) atomicAdd(&s_pos, pos);
  __syncthreads();
  for (int i = threadIdx.x; i < kBins; i += kSelThreads)
    if (sh[i]) atomicAdd(&st->hist[PASS][i], sh[i]);
  if (PASS == 0 && threadIdx.x == 0 && s_pos) atomicAdd(&st->n_pos, s_pos);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket[PASS], 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA: pick the digit d that contains the k_rem-th largest key = the first d, walking down from the top bin, with
  // (#keys in bins above d) + hist[d] >= k_rem; d = 0 if no bin above 0 gets there.  Thread t owns `per` consecutive bins,
  // descending from nb - 1 - t * per; a block scan gives every thread the count above its range.
  __shared__ long long s_scan[kSelThreads];
  __shared__ long long s_krem;
  __shared__ int s_found;
  volatile SelState* v = st;
  if (threadIdx.x == 0) {
    long long k_rem;
    if (PASS == 0) {
      long long k = num_sample;
      if ((long long)v->n_pos > k) k = v->n_pos;
      if (k > n) k = n;
      v->k = k;
      k_rem = k;
    } else {
      k_rem = v->k_rem;
    }
    s_krem = k_rem;
    s_found = 0;
  }
  constexpr int nb = 1 << (PASS == 2 ? 10 : 11);
  constexpr int per = nb / kSelThreads;
  const int top = nb - 1 - (int)threadIdx.x * per;
  long long c[per], loc = 0;
#pragma unroll
  for (int j = 0; j < per; ++j) { c[j] = v->hist[PASS][top - j]; loc += c[j]; }
  long long above = block_incl_scan(loc, s_scan) - loc;       // keys in the bins above this thread's range (syncs: s_krem visible)
  const long long k_rem = s_krem;
#pragma unroll
  for (int j = 0; j < per; ++j) {
    const int d = top - j;
    if (d > 0 && above + c[j] >= k_rem && (d == nb - 1 || above < k_rem)) {      // exactly one (thread, j) can satisfy this
      s_found = 1;
      const unsigned int p = prefix | ((unsigned int)d << shift);
      v->prefix = p;
      v->k_rem = k_rem - above;
      if (PASS == 2) { v->threshold = p; v->need_eq = k_rem - above; }
    }
    if (d > 0) above += c[j];
  }
  __syncthreads();
  if (threadIdx.x == kSelThreads - 1 && !s_found) {            // owner of bin 0; `above` = all keys in bins 1 .. nb-1
    v->prefix = prefix;
    v->k_rem = k_rem - above;
    if (PASS == 2) { v->threshold = prefix; v->need_eq = k_rem - above; }
  }
}

// per-CTA counts of keys > T and keys == T; last CTA turns them into exclusive offsets
__global__ void __launch_bounds__(kSelThreads)
pfc_count_kernel(const float* __restrict__ perm, int64_t n, SelState* st, long long* __restrict__ blk_gt,
                 long long* __restrict__ blk_eq) {
  __shared__ unsigned int s_gt, s_eq;
  __shared__ bool s_last;
  if (threadIdx.x == 0) { s_gt = 0; s_eq = 0; }
  __syncthreads();
  const unsigned int T = st->threshold;
  unsigned int gt = 0, eq = 0;
  const int64_t base = (int64_t)blockIdx.x * kSelTile;
#pragma unroll
  for (int j = 0; j < kSelItems; ++j) {
    const int64_t i = base + (int64_t)j * kSelThreads + threadIdx.x;
    if (i < n) {
      const unsigned int key = f2key(perm[i]);
      gt += key > T;
      eq += key == T;
    }
  }
  gt = __reduce_add_sync(0xffffffffu, gt);
  eq = __reduce_add_sync(0xffffffffu, eq);
  if ((threadIdx.x & 31) == 0) { if (gt) atomicAdd(&s_gt, gt); if (eq) atomicAdd(&s_eq, eq); }
  __syncthreads();
  if (threadIdx.x == 0) {
    blk_gt[blockIdx.x] = s_gt;
    blk_eq[blockIdx.x] = s_eq;
    __threadfence();
    s_last = (atomicAdd(&st->ticket[3], 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // exclusive scan of the per-CTA counts (31 CTAs at 125,000 classes, 245 at 1,000,000), 512 at a time
  __shared__ long long s_scan[kSelThreads];
  __shared__ long long s_tot[2];
  volatile long long* g = blk_gt;
  volatile long long* e = blk_eq;
  long long carry_g = 0, carry_e = 0;
  for (unsigned int b0 = 0; b0 < gridDim.x; b0 += kSelThreads) {
    const unsigned int b = b0 + threadIdx.x;
    const long long cg = b < gridDim.x ? g[b] : 0, ce = b < gridDim.x ? e[b] : 0;
    const long long ig = block_incl_scan(cg, s_scan), ie = block_incl_scan(ce, s_scan);
    if (b < gridDim.x) { g[b] = carry_g + ig - cg; e[b] = carry_e + ie - ce; }
    if (threadIdx.x == kSelThreads - 1) { s_tot[0] = ig; s_tot[1] = ie; }     // chunk totals
    __syncthreads();
    carry_g += s_tot[0];
    carry_e += s_tot[1];
    __syncthreads();
  }
}

// ordered compaction: element i is taken iff key > T, or key == T and it is among the first
// need_eq such keys; its slot = (#taken before i).
__global__ void __launch_bounds__(kSelThreads)
pfc_write_kernel(const float* __restrict__ perm, int64_t n, const SelState* __restrict__ st,
                 const long long* __restrict__ blk_gt, const long long* __restrict__ blk_eq,
                 int64_t* __restrict__ index, int64_t* __restrict__ n_index) {
  __shared__ unsigned int w_gt[kSelThreads / 32], w_eq[kSelThreads / 32];
  const unsigned int T = st->threshold;
  const long long need_eq = st->need_eq;
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_index = st->k;
  // thread t owns kSelItems CONSECUTIVE elements so that order is preserved
  const int64_t base = (int64_t)blockIdx.x * kSelTile + (int64_t)threadIdx.x * kSelItems;
  unsigned int key[kSelItems];
  unsigned int gt = 0, eq = 0;
#pragma unroll
  for (int j = 0; j < kSelItems; ++j) {
    const int64_t i = base + j;
    key[j] = (i < n) ? f2key(perm[i]) : 0u;   // key 0 is below every real key (f2key never yields 0 for non-NaN)
    if (i < n) { gt += key[j] > T; eq += key[j] == T; }
  }
  // block exclusive scan of (gt, eq)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int sg = gt, se = eq;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int tg = __shfl_up_sync(0xffffffffu, sg, o), te = __shfl_up_sync(0xffffffffu, se, o);
    if (lane >= o) { sg += tg; se += te; }
  }
  if (lane == 31) { w_gt[warp] = sg; w_eq[warp] = se; }
  __syncthreads();
  unsigned int og = 0, oe = 0;
  for (int w = 0; w < warp; ++w) { og += w_gt[w]; oe += w_eq[w]; }
  long long pre_gt = blk_gt[blockIdx.x] + og + (sg - gt);
  long long pre_eq = blk_eq[blockIdx.x] + oe + (se - eq);
#pragma unroll
  for (int j = 0; j < kSelItems; ++j) {
    const int64_t i = base + j;
    if (i >= n) break;
    const bool is_gt = key[j] > T, is_eq = key[j] == T;
    const long long slot = pre_gt + (pre_eq < need_eq ? pre_eq : need_eq);
    if (is_gt || (is_eq && pre_eq < need_eq)) index[slot] = i;
    pre_gt += is_gt;
    pre_eq += is_eq;
  }
}

__global__ void pfc_searchsorted_kernel(int64_t* __restrict__ tl, int64_t n, const int64_t* __restrict__ index,
                                        const int64_t* __restrict__ n_index) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t v = tl[i];
  if (v < 0) return;
  int64_t lo = 0, hi = *n_index;   // lower_bound
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (index[mid] < v) lo = mid + 1; else hi = mid;
  }
  tl[i] = lo;
}

// one warp per row, 128-bit accesses
template <bool SCATTER>
__global__ void __launch_bounds__(256)
rows_copy_kernel(const float* __restrict__ src, const int64_t* __restrict__ index, float* __restrict__ dst,
                 int64_t n_rows, int64_t d4) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int64_t other = index[row];
  const float4* s = reinterpret_cast<const float4*>(src) + (SCATTER ? row : other) * d4;
  float4* o = reinterpret_cast<float4*>(dst) + (SCATTER ? other : row) * d4;
  for (int64_t j = threadIdx.x & 31; j < d4; j += 32) o[j] = s[j];
}

}  // namespace msml
