// K-A  mask fusion: fused gate / arith / skip of the Feature-Masking operator, forward + backward: the kernels and the
//      CTA dealing (launchers: fm_gate.cu; tests/emu runs this header under the CPU emulation and its sanitizers).
//   ref backbones/fm/fmoperator.py:288 (act), :304 (arith), :307-308 (+ f_out), :310 (+ identity)
//
// HBM-bound streaming kernels: every tensor is read or written exactly once.
//   fwd: read yf, z [, f_out]  write out          -> 3 passes   (eager: cat-free tail = 9)
//   bwd: read dout, yf, z      write dyf, dz      -> 5 passes   (gate recomputed, not saved)
// Layout: all operands share one physical layout (NHWC in the drop-in), so the element index is
// flat; each thread moves UNROLL independent 128-bit vectors per stream per iteration
// (ld.global.nc.L1::no_allocate / st.global.cs), persistent grid = SMs x CTAS_PER_SM.
// Up to MSML_MAX_SEGMENTS tensors (the four feature scales) go through ONE launch.
#pragma once
#include "common.cuh"

namespace msml {

constexpr int kThreads = 256;
constexpr int kCtasPerSm = 4;     // fwd: 2-3 input streams, 64 regs
constexpr int kCtasPerSmBwd = 4;
constexpr int kUnroll = 4;         // 2 input streams
constexpr int kUnroll3 = 2;        // 3 input streams (fwd with f_out, bwd): same bytes in flight, no spills

// Each CTA works inside ONE segment (blocks are dealt to segments in proportion to their size by
// the host), so the per-CTA pointers are loop invariants and the streaming loop stays lean.
struct FwdSegs {
  const void* yf[MSML_MAX_SEGMENTS];
  const void* z[MSML_MAX_SEGMENTS];
  const void* f_out[MSML_MAX_SEGMENTS];
  void* out[MSML_MAX_SEGMENTS];
  int64_t n[MSML_MAX_SEGMENTS];          // elements
  int block_end[MSML_MAX_SEGMENTS];      // exclusive prefix end of the CTAs dealt to each segment
  int nseg;
};
struct BwdSegs {
  const void* dout[MSML_MAX_SEGMENTS];
  const void* yf[MSML_MAX_SEGMENTS];
  const void* z[MSML_MAX_SEGMENTS];
  void* dyf[MSML_MAX_SEGMENTS];
  void* dz[MSML_MAX_SEGMENTS];
  int64_t n[MSML_MAX_SEGMENTS];
  int block_end[MSML_MAX_SEGMENTS];
  int nseg;
};

// Static-index select chain (no dynamic indexing of kernel parameters => no local-memory copy).
#define MSML_PICK_SEG(s, field, seg, out)                              \
  {                                                                    \
    out = s.field[0];                                                  \
    _Pragma("unroll") for (int _i = 1; _i < MSML_MAX_SEGMENTS; ++_i)   \
      if (seg == _i) out = s.field[_i];                                \
  }

template <typename Segs>
__device__ __forceinline__ int block_segment(const Segs& s, int& first_block, int& nblocks) {
  int seg = 0, begin = 0, end = s.block_end[0];
#pragma unroll
  for (int i = 1; i < MSML_MAX_SEGMENTS; ++i) {
    if (i < s.nseg && (int)blockIdx.x >= s.block_end[i - 1]) { seg = i; begin = s.block_end[i - 1]; end = s.block_end[i]; }
  }
  first_block = begin;
  nblocks = end - begin;
  return seg;
}

template <typename T, int ACT, int ARITH, bool HAS_FOUT>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
fm_gate_fwd_kernel(const __grid_constant__ FwdSegs s) {
  constexpr int VN = Vec<T>::N;
  constexpr int U = HAS_FOUT ? kUnroll3 : kUnroll;
  int first_block, nblocks;
  const int seg = block_segment(s, first_block, nblocks);
  const void *p_yf, *p_z, *p_fo;
  void* p_out;
  int64_t n;
  MSML_PICK_SEG(s, yf, seg, p_yf);
  MSML_PICK_SEG(s, z, seg, p_z);
  MSML_PICK_SEG(s, f_out, seg, p_fo);
  MSML_PICK_SEG(s, out, seg, p_out);
  MSML_PICK_SEG(s, n, seg, n);
  const uint4* yf4 = static_cast<const uint4*>(p_yf);
  const uint4* z4 = static_cast<const uint4*>(p_z);
  const uint4* fo4 = static_cast<const uint4*>(p_fo);
  uint4* out4 = static_cast<uint4*>(p_out);
  const int64_t nvec = n / VN;
  const int64_t stride = (int64_t)nblocks * kThreads;
  // one iteration = kUnroll vectors spaced `stride` apart per stream: coalesced per instruction,
  // all loads issued before the first use
  for (int64_t base = (int64_t)(blockIdx.x - first_block) * kThreads + threadIdx.x; base < nvec;
       base += stride * U) {
    uint4 a[U], b[U], c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = base + u * stride;
      if (v < nvec) {
        a[u] = ld_stream(yf4 + v);
        b[u] = ld_stream(z4 + v);
        if (HAS_FOUT) c[u] = ld_stream(fo4 + v);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = base + u * stride;
      if (v >= nvec) break;
      float yf[VN], z[VN], fo[VN], o[VN];
      Vec<T>::unpack(a[u], yf);
      Vec<T>::unpack(b[u], z);
      if (HAS_FOUT) Vec<T>::unpack(c[u], fo);
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        const float g = gate_act<ACT>(z[i]);
        o[i] = gate_fuse<ARITH>(yf[i], g);
        if (HAS_FOUT) o[i] += fo[i];
      }
      st_stream(out4 + v, Vec<T>::pack(o));
    }
  }
  // scalar tail (n % VN elements); first CTA of the segment only
  if ((int)blockIdx.x == first_block) {
    for (int64_t i = nvec * VN + threadIdx.x; i < n; i += kThreads) {
      const float yf = to_f32(static_cast<const T*>(p_yf)[i]);
      const float g = gate_act<ACT>(to_f32(static_cast<const T*>(p_z)[i]));
      float o = gate_fuse<ARITH>(yf, g);
      if (HAS_FOUT) o += to_f32(static_cast<const T*>(p_fo)[i]);
      static_cast<T*>(p_out)[i] = from_f32<T>(o);
    }
  }
}

template <typename T, int ACT, int ARITH>
__global__ void __launch_bounds__(kThreads, kCtasPerSmBwd)
fm_gate_bwd_kernel(const __grid_constant__ BwdSegs s) {
  constexpr int VN = Vec<T>::N;
  constexpr int U = kUnroll3;
  int first_block, nblocks;
  const int seg = block_segment(s, first_block, nblocks);
  const void *p_d, *p_yf, *p_z;
  void *p_dyf, *p_dz;
  int64_t n;
  MSML_PICK_SEG(s, dout, seg, p_d);
  MSML_PICK_SEG(s, yf, seg, p_yf);
  MSML_PICK_SEG(s, z, seg, p_z);
  MSML_PICK_SEG(s, dyf, seg, p_dyf);
  MSML_PICK_SEG(s, dz, seg, p_dz);
  MSML_PICK_SEG(s, n, seg, n);
  const uint4* d4 = static_cast<const uint4*>(p_d);
  const uint4* yf4 = static_cast<const uint4*>(p_yf);
  const uint4* z4 = static_cast<const uint4*>(p_z);
  uint4* dyf4 = static_cast<uint4*>(p_dyf);
  uint4* dz4 = static_cast<uint4*>(p_dz);
  const int64_t nvec = n / VN;
  const int64_t stride = (int64_t)nblocks * kThreads;
  for (int64_t base = (int64_t)(blockIdx.x - first_block) * kThreads + threadIdx.x; base < nvec;
       base += stride * U) {
    uint4 a[U], b[U], c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = base + u * stride;
      if (v < nvec) {
        a[u] = ld_stream(d4 + v);
        b[u] = ld_stream(yf4 + v);
        c[u] = ld_stream(z4 + v);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = base + u * stride;
      if (v >= nvec) break;
      float d[VN], yf[VN], z[VN], dyf[VN], dz[VN];
      Vec<T>::unpack(a[u], d);
      Vec<T>::unpack(b[u], yf);
      Vec<T>::unpack(c[u], z);
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        const float g = gate_act<ACT>(z[i]);
        float dg;
        gate_fuse_grad<ARITH>(d[i], yf[i], g, dyf[i], dg);
        dz[i] = dg * gate_act_grad<ACT>(g);
      }
      st_stream(dyf4 + v, Vec<T>::pack(dyf));
      st_stream(dz4 + v, Vec<T>::pack(dz));
    }
  }
  if ((int)blockIdx.x == first_block) {
    for (int64_t i = nvec * VN + threadIdx.x; i < n; i += kThreads) {
      const float d = to_f32(static_cast<const T*>(p_d)[i]);
      const float yf = to_f32(static_cast<const T*>(p_yf)[i]);
      const float g = gate_act<ACT>(to_f32(static_cast<const T*>(p_z)[i]));
      float dyf, dg;
      gate_fuse_grad<ARITH>(d, yf, g, dyf, dg);
      static_cast<T*>(p_dyf)[i] = from_f32<T>(dyf);
      static_cast<T*>(p_dz)[i] = from_f32<T>(dg * gate_act_grad<ACT>(g));
    }
  }
}

// Deal a persistent grid (SMs x ctas_per_sm, or fewer for small inputs) to the segments in
// proportion to their vector counts; every segment gets at least one CTA.
inline int deal_blocks(int nseg, const int64_t* n, int vn, int unroll, int ctas_per_sm, int* block_end, int sms) {
  int64_t total_vec = 0;
  for (int i = 0; i < nseg; ++i) total_vec += n[i] / vn;
  const int64_t per_block = (int64_t)kThreads * unroll;
  int64_t want = (total_vec + per_block - 1) / per_block;
  const int64_t cap = (int64_t)sms * ctas_per_sm;
  if (want > cap) want = cap;
  if (want < nseg) want = nseg;
  int acc = 0;
  int64_t vec_acc = 0;
  for (int i = 0; i < nseg; ++i) {
    vec_acc += n[i] / vn;
    int end = total_vec > 0 ? (int)((vec_acc * want + total_vec - 1) / total_vec) : i + 1;
    if (end < acc + 1) end = acc + 1;           // at least one CTA per segment
    const int remaining = nseg - 1 - i;         // leave one for each later segment
    if (end > want - remaining) end = (int)(want - remaining) > acc ? (int)(want - remaining) : acc + 1;
    block_end[i] = end;
    acc = end;
  }
  for (int i = nseg; i < MSML_MAX_SEGMENTS; ++i) block_end[i] = acc;
  return acc;
}

}  // namespace msml
