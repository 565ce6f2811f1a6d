// Minimal CPU emulation of the CUDA execution model, for LOGIC checks of msml_b200 kernels without a GPU.
// TEST INFRASTRUCTURE ONLY.  A kernel header is compiled for the host with g++ (this file included first) and each
// CTA is executed by blockDim.x OS threads: __syncthreads is a barrier over the CTA, warp shuffles exchange values
// through a per-warp buffer, __shared__ variables are function-local statics (CTAs run one after another).
// What this does NOT check: memory-model races, alignment rules of vector accesses, launch limits, performance.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <new>
#include <thread>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __launch_bounds__
#undef __shared__
#undef __grid_constant__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __grid_constant__

namespace cuda_emu {
struct Cta {
  std::barrier<> all;
  std::vector<std::unique_ptr<std::barrier<>>> warp;
  std::vector<uint64_t> xchg;          // one 8-byte slot per thread
  std::atomic<int> vote{0};
  explicit Cta(int threads) : all(threads), xchg(threads) {
    for (int w = 0; w < (threads + 31) / 32; ++w) {
      const int n = threads - w * 32 < 32 ? threads - w * 32 : 32;
      warp.emplace_back(new std::barrier<>(n));
    }
  }
};
inline thread_local Cta* cta = nullptr;
}  // namespace cuda_emu

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { cuda_emu::cta->all.arrive_and_wait(); }
inline int __syncthreads_or(int pred) {
  cuda_emu::Cta* c = cuda_emu::cta;
  c->all.arrive_and_wait();
  if (threadIdx.x == 0) c->vote.store(0);
  c->all.arrive_and_wait();
  if (pred) c->vote.fetch_or(1);
  c->all.arrive_and_wait();
  return c->vote.load();
}
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <typename V>
inline V __shfl_xor_sync(unsigned, V v, int lane_mask) {
  static_assert(sizeof(V) <= 8, "shuffle of up to 8 bytes");
  cuda_emu::Cta* c = cuda_emu::cta;
  const int t = (int)threadIdx.x, w = t >> 5;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(V));
  c->xchg[t] = bits;
  c->warp[w]->arrive_and_wait();
  const int src = (t & ~31) | ((t & 31) ^ lane_mask);
  const uint64_t got = src < (int)blockDim.x ? c->xchg[src] : bits;
  c->warp[w]->arrive_and_wait();
  V out;
  std::memcpy(&out, &got, sizeof(V));
  return out;
}

template <typename V>
inline V __shfl_up_sync(unsigned, V v, int delta) {
  static_assert(sizeof(V) <= 8, "shuffle of up to 8 bytes");
  cuda_emu::Cta* c = cuda_emu::cta;
  const int t = (int)threadIdx.x, w = t >> 5, lane = t & 31;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(V));
  c->xchg[t] = bits;
  c->warp[w]->arrive_and_wait();
  const uint64_t got = lane >= delta ? c->xchg[t - delta] : bits;
  c->warp[w]->arrive_and_wait();
  V out;
  std::memcpy(&out, &got, sizeof(V));
  return out;
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
// atomics on shared (function-local statics) and global memory: real atomic read-modify-writes between the CTA's threads
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline float atomicAdd(float* p, float v) {                      // red.add.f32: a compare-and-swap loop on the bit pattern
  unsigned* u = reinterpret_cast<unsigned*>(p);
  unsigned old = __atomic_load_n(u, __ATOMIC_RELAXED), want;
  float f;
  do {
    std::memcpy(&f, &old, 4);
    f += v;
    std::memcpy(&want, &f, 4);
  } while (!__atomic_compare_exchange_n(u, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  std::memcpy(&f, &old, 4);
  return f;
}

template <typename V> inline V __ldg(const V* p) { return *p; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
inline float __uint_as_float(unsigned v) { float f; std::memcpy(&f, &v, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned v; std::memcpy(&v, &f, 4); return v; }
inline float __expf(float x) { return std::exp(x); }
inline float __fdividef(float a, float b) { return a / b; }
inline void __nanosleep(unsigned) {}

inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }

// cooperative groups: only what kernels that ALSO have a multi-launch form need to compile; a grid-wide barrier cannot be
// emulated with CTAs that run one after another, so grid.sync() aborts if it is ever reached
namespace cooperative_groups {
struct grid_group {
  void sync() const { std::abort(); }
};
inline grid_group this_grid() { return {}; }
}  // namespace cooperative_groups

// launch<<<grid, block>>>: CTAs one after another, the threads of a CTA concurrently.  The OS threads are created once
// per launch and walk through the CTAs together; a thread that returns from the kernel early leaves the CTA barrier
// (as an exited CUDA thread does) and the barrier is rebuilt for the next CTA.
template <typename F>
inline void emu_launch(dim3 grid, int threads, F&& body) {
  cuda_emu::Cta cta(threads);
  std::barrier<> cta_end(threads);
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&, t] {
      cuda_emu::cta = &cta;
      threadIdx = {(unsigned)t, 0, 0};
      blockDim = dim3(threads);
      gridDim = grid;
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            blockIdx = {bx, by, bz};
            body();
            cta.all.arrive_and_drop();                 // exited threads no longer take part in __syncthreads
            cta_end.arrive_and_wait();                 // the whole CTA has finished
            if (t == 0) {
              cta.all.~barrier();
              new (&cta.all) std::barrier<>(threads);
            }
            cta_end.arrive_and_wait();                 // the barrier is fresh before anyone enters the next CTA
          }
    });
  for (auto& th : pool) th.join();
}
