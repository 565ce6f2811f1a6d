// Class-sharded head over NCCL, owned by the library.
//   ref headers/partial_fc.py issues SIX collectives per step from Python: all_gather(labels) :110, all_gather(features)
//   :126, all_reduce(MAX) :136, all_reduce(SUM) :141, all_reduce(SUM) of the target probability :162, reduce_scatter :174.
// Here a step is THREE collectives, enqueued from C back to back with the kernels on ONE stream:
//   msml_head_gather  pack [bf16 embeddings | int64 labels] of this rank into one message -> ncclAllGather -> unpack into the
//                     contiguous (B_tot, D) operand of the tcgen05 GEMM and the shard-local labels (the remap of ref :79-81 is
//                     done by the unpack kernel)                                                   = ref :110 + :126 + :79-81
//   msml_head_step    msml_head_fwd -> ncclAllGather of the per-row (max, sum-exp, target logit) -> merge (global max, sum,
//                     loss) -> msml_head_bwd -> ncclReduceScatter(dX) -> x world_size               = ref :132-175
// NCCL is bound at run time (dlopen "libnccl.so.2": inside a PyTorch process that is the copy torch already loaded, so the
// library has no link-time dependency and never brings a second NCCL into the process).  The communicator is created from a
// 128-byte unique id that the host side broadcasts through whatever it already has (torch.distributed in this repo).
// With comm == NULL (world size 1) the same entry points run without any collective.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace msml {
namespace {

// the handful of NCCL types / enums used here (ABI-stable since NCCL 2.0; values as in nccl.h)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclSuccess = 0, kNcclInt8 = 0, kNcclFloat32 = 7, kNcclSum = 0 };

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok;
};

NcclApi* nccl_api() {
  static NcclApi api = {};
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.ReduceScatter = reinterpret_cast<decltype(api.ReduceScatter)>(dlsym(h, "ncclReduceScatter"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.ReduceScatter && api.GetErrorString;
  });
  return &api;
}

#define MSML_NCCL(expr)                                                                                        \
  do {                                                                                                         \
    ncclResult_t _r = (expr);                                                                                  \
    if (_r != kNcclSuccess)                                                                                    \
      return ::msml::set_error(MSML_ENCCL, "%s failed: %s (%s:%d)", #expr, nccl_api()->GetErrorString(_r), __FILE__, __LINE__); \
  } while (0)

// ---- kernels -----------------------------------------------------------------------------------------------------------
// message of one rank: [B x D bf16 | B x int64], padded to 16 bytes
__host__ __device__ inline size_t msg_bytes(int64_t B, int64_t D) { return ((size_t)B * D * 2 + (size_t)B * 8 + 15) / 16 * 16; }

// one warp per embedding row: fp32 -> bf16 (128-bit stores); the first warps also copy the labels behind the embeddings
__global__ void __launch_bounds__(256)
head_pack_kernel(const float* __restrict__ feat, const int64_t* __restrict__ label, uint8_t* __restrict__ msg, int64_t B, int D) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float4* src = reinterpret_cast<const float4*>(feat + row * D);
  uint4* dst = reinterpret_cast<uint4*>(msg + (size_t)row * D * 2);
  for (int j = lane; j < D / 8; j += 32) {
    const float4 a = __ldg(src + 2 * j), b = __ldg(src + 2 * j + 1);
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    dst[j] = Vec<__nv_bfloat16>::pack(f);
  }
  if (lane == 0) reinterpret_cast<int64_t*>(msg + (size_t)B * D * 2)[row] = label[row];
}

// gathered messages (W x msg) -> X (W*B, D) bf16 contiguous, total_label (W*B) remapped to the shard (ref :79-81)
__global__ void __launch_bounds__(256)
head_unpack_kernel(const uint8_t* __restrict__ gathered, size_t msg, __nv_bfloat16* __restrict__ x, int64_t* __restrict__ tl,
                   int64_t B, int D, int64_t W, int64_t class_start, int64_t num_local) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);     // global row in [0, W*B)
  const int lane = threadIdx.x & 31;
  if (row >= W * B) return;
  const int64_t r = row / B, i = row - r * B;
  const uint8_t* m = gathered + (size_t)r * msg;
  const uint4* src = reinterpret_cast<const uint4*>(m + (size_t)i * D * 2);
  uint4* dst = reinterpret_cast<uint4*>(x + row * D);
  for (int j = lane; j < D / 8; j += 32) dst[j] = src[j];
  if (lane == 0) {
    const int64_t v = reinterpret_cast<const int64_t*>(m + (size_t)B * D * 2)[i];
    tl[row] = (v >= class_start && v < class_start + num_local) ? v - class_start : -1;
  }
}

__global__ void scale_kernel(float* __restrict__ p, int64_t n, float s) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] *= s;
}

}  // namespace
}  // namespace msml

using namespace msml;

struct msml_comm {
  ncclComm_t nccl;
  int rank, world;
};

extern "C" int msml_nccl_unique_id(void* out128) {
  MSML_REQUIRE(out128, MSML_EINVAL, "null pointer");
  NcclApi* api = nccl_api();
  MSML_REQUIRE(api->ok, MSML_ENCCL, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "symbols missing");
  ncclUniqueId id;
  MSML_NCCL(api->GetUniqueId(&id));
  memcpy(out128, id.internal, 128);
  return 0;
}

extern "C" int msml_nccl_init(const void* unique_id128, int rank, int world, msml_comm** out) {
  MSML_REQUIRE(unique_id128 && out && world >= 1 && rank >= 0 && rank < world, MSML_EINVAL, "bad communicator arguments");
  NcclApi* api = nccl_api();
  MSML_REQUIRE(api->ok, MSML_ENCCL, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  memcpy(id.internal, unique_id128, 128);
  ncclComm_t c = nullptr;
  MSML_NCCL(api->CommInitRank(&c, world, id, rank));          // synchronous: the only blocking call of this file
  msml_comm* m = new msml_comm{c, rank, world};
  *out = m;
  return 0;
}

extern "C" int msml_nccl_destroy(msml_comm* comm) {
  if (!comm) return 0;
  NcclApi* api = nccl_api();
  ncclResult_t r = api->ok ? api->CommDestroy(comm->nccl) : kNcclSuccess;
  delete comm;
  MSML_REQUIRE(r == kNcclSuccess, MSML_ENCCL, "ncclCommDestroy failed: %s", api->GetErrorString(r));
  return 0;
}

extern "C" int msml_comm_world(const msml_comm* comm) { return comm ? comm->world : 1; }
extern "C" int msml_comm_rank(const msml_comm* comm) { return comm ? comm->rank : 0; }

extern "C" size_t msml_head_gather_workspace(int64_t B, int64_t W, int64_t D) {
  if (B <= 0 || W <= 0 || D <= 0) return 0;
  return (size_t)(W + 1) * msg_bytes(B, D);                  // this rank's message + the W gathered ones
}

extern "C" int msml_head_gather(msml_comm* comm, const float* feat, const int64_t* label, int64_t B, int64_t D,
                                int64_t class_start, int64_t num_local, void* x_bf16, int64_t* total_label, void* ws,
                                size_t ws_bytes, void* stream) {
  const int64_t W = comm ? comm->world : 1;
  MSML_REQUIRE(feat && label && x_bf16 && total_label && B > 0 && D > 0 && D % 8 == 0 && num_local > 0 && class_start >= 0,
               MSML_EINVAL, "bad gather arguments (D %% 8 must be 0)");
  MSML_REQUIRE(aligned16(feat) && aligned16(x_bf16) && ws && aligned16(ws), MSML_EALIGN, "gather buffers must be 16-byte aligned");
  MSML_REQUIRE(ws_bytes >= msml_head_gather_workspace(B, W, D), MSML_EWORKSPACE, "gather workspace too small: %zu < %zu", ws_bytes,
               msml_head_gather_workspace(B, W, D));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t msg = msg_bytes(B, D);
  uint8_t* send = static_cast<uint8_t*>(ws);
  uint8_t* recv = send + msg;
  {
    MSML_PROF("head_pack", (double)B * D * 6, st);
    head_pack_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(feat, label, W > 1 ? send : recv, B, (int)D);
    MSML_LAUNCH_CHECK();
  }
  if (W > 1) {
    MSML_PROF("nccl_allgather_embeddings_labels", (double)W * msg, st);     // per-collective device time (includes waiting for peers)
    MSML_NCCL(nccl_api()->AllGather(send, recv, msg, kNcclInt8, comm->nccl, st));
  }
  {
    MSML_PROF("head_unpack", (double)W * B * D * 4, st);
    head_unpack_kernel<<<(unsigned)((W * B + 7) / 8), 256, 0, st>>>(recv, msg, static_cast<__nv_bfloat16*>(x_bf16), total_label, B,
                                                                   (int)D, W, class_start, num_local);
    MSML_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" size_t msml_head_step_workspace(int64_t B, int64_t W, int64_t n_s, int64_t D) {
  if (B <= 0 || W <= 0 || n_s <= 0) return 0;
  const int64_t B_tot = B * W;
  size_t off = (msml_head_workspace(B_tot, n_s, D) + 255) / 256 * 256;
  off += ((size_t)3 * B_tot * 4 + 255) / 256 * 256;          // stats
  off += ((size_t)W * 3 * B_tot * 4 + 255) / 256 * 256;      // gathered stats
  off += ((size_t)2 * B_tot * 4 + 255) / 256 * 256;          // merged (max, sum)
  off += ((size_t)B_tot * D * 4 + 255) / 256 * 256;          // dX over the gathered batch
  return off;
}

static int head_step_impl(msml_comm* comm, const void* x, const void* wn, const float* inv_norm, const int64_t* tl, int64_t B,
                          int64_t n_s, int64_t D, const msml_margin_params* margin, float* x_grad, float* dw, float* loss,
                          void* ws, size_t ws_bytes, void* stream, bool raw) {
  const int64_t W = comm ? comm->world : 1, B_tot = B * W;
  MSML_REQUIRE(x && wn && (raw || inv_norm) && tl && x_grad && dw && loss && B > 0, MSML_EINVAL, "null pointer");
  MSML_REQUIRE(ws && aligned16(ws) && ws_bytes >= msml_head_step_workspace(B, W, n_s, D), MSML_EWORKSPACE,
               "head step workspace missing, misaligned or too small (%zu < %zu)", ws_bytes, msml_head_step_workspace(B, W, n_s, D));
  cudaStream_t st = (cudaStream_t)stream;
  char* base = static_cast<char*>(ws);
  size_t off = 0;
  auto carve = [&](size_t bytes) { void* p = base + off; off += (bytes + 255) / 256 * 256; return p; };
  const size_t hw = msml_head_workspace(B_tot, n_s, D);
  void* head_ws = carve(hw);
  float* stats = static_cast<float*>(carve((size_t)3 * B_tot * 4));
  float* gathered = static_cast<float*>(carve((size_t)W * 3 * B_tot * 4));
  float* gstats = static_cast<float*>(carve((size_t)2 * B_tot * 4));
  float* dx_full = static_cast<float*>(carve((size_t)B_tot * D * 4));

  if (int e = msml_head_fwd(x, wn, tl, B_tot, n_s, D, margin, stats, head_ws, hw, stream)) return e;
  // ONE all-gather of (max, sum-exp, target logit) per row replaces all_reduce(MAX), all_reduce(SUM) and the loss all_reduce
  if (W > 1) {
    MSML_PROF("nccl_allgather_row_stats", (double)W * 3 * B_tot * 4, st);
    MSML_NCCL(nccl_api()->AllGather(stats, gathered, (size_t)3 * B_tot, kNcclFloat32, comm->nccl, st));
  }
  if (int e = msml_head_merge_stats(W > 1 ? gathered : stats, W, B_tot, gstats, loss, stream)) return e;
  if (raw) {
    if (int e = msml_head_bwd_raw(x, wn, tl, B_tot, n_s, D, margin, gstats, W > 1 ? dx_full : x_grad, dw, head_ws, hw, stream)) return e;
  } else {
    if (int e = msml_head_bwd(x, wn, inv_norm, tl, B_tot, n_s, D, margin, gstats, W > 1 ? dx_full : x_grad, dw, head_ws, hw, stream)) return e;
  }
  if (W > 1) {
    {
      MSML_PROF("nccl_reduce_scatter_dx", (double)B_tot * D * 4, st);
      MSML_NCCL(nccl_api()->ReduceScatter(dx_full, x_grad, (size_t)B * D, kNcclFloat32, kNcclSum, comm->nccl, st));
    }
    scale_kernel<<<(unsigned)((B * D + 255) / 256), 256, 0, st>>>(x_grad, B * D, (float)W);      // ref :175
    MSML_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int msml_head_step(msml_comm* comm, const void* x, const void* wn, const float* inv_norm, const int64_t* tl, int64_t B,
                              int64_t n_s, int64_t D, const msml_margin_params* margin, float* x_grad, float* dw, float* loss,
                              void* ws, size_t ws_bytes, void* stream) {
  return head_step_impl(comm, x, wn, inv_norm, tl, B, n_s, D, margin, x_grad, dw, loss, ws, ws_bytes, stream, false);
}

extern "C" int msml_head_step_raw(msml_comm* comm, const void* x, const void* wn, const int64_t* tl, int64_t B, int64_t n_s,
                                  int64_t D, const msml_margin_params* margin, float* x_grad, float* dwn, float* loss, void* ws,
                                  size_t ws_bytes, void* stream) {
  return head_step_impl(comm, x, wn, nullptr, tl, B, n_s, D, margin, x_grad, dwn, loss, ws, ws_bytes, stream, true);
}
