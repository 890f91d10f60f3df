// Context object behind the opaque void* of include/mqb200.h.
// Conventions of capp/src/libllmod.cpp:23-65 (magic + version + refcount) re-designed for a CUDA library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <atomic>
#include <mutex>
#include <vector>
#include <utility>
#include "mqb200.h"

namespace mq {

constexpr uint32_t kMagic = 0x4d514232u;  // "MQB2"

struct Ctx {
  uint32_t magic = kMagic;
  std::atomic<int> refs{1};
  int device = 0;
  int sm_count = 148;
  void* ws = nullptr;      // device scratch of the first stream seen (deterministic two-stage reductions, split-K partials)
  size_t ws_bytes = 0;     // size of EVERY per-stream scratch buffer
  // One scratch buffer per stream: kernels of different streams (the calibration step runs its weight pass on a side stream)
  // must not share reduction partials.  Buffers are created on first use of a stream (never during a graph capture: the
  // eager warm-up step of the calibration loops touches every stream first) and live as long as the context.
  std::mutex ws_mutex;
  std::vector<std::pair<cudaStream_t, void*>> ws_by_stream;
  int* counters = nullptr; // zero-initialised arrival counters of the fused skinny-GEMM epilogue (self-resetting)
  int n_counters = 0;
  // Destination of the latest mq_unpack4: weight codes inside this range are produced by a kernel of the same stream (packed
  // 4-bit weights expanded into scratch right before the GEMM), so the skinny GEMM must not request them ahead of its grid
  // dependency (decode.cu: weight-tile prefetch before pdl_wait).
  const char* unpack_lo = nullptr;
  const char* unpack_hi = nullptr;
  std::string last_error[6];
};

inline Ctx* as_ctx(void* p) {
  Ctx* c = reinterpret_cast<Ctx*>(p);
  return (c && c->magic == kMagic) ? c : nullptr;
}

int fail(Ctx* c, int code, const std::string& what);
void* stream_ws(Ctx* c, cudaStream_t st);      // scratch buffer (ws_bytes) owned by `st`; nullptr when it cannot be allocated
int check_launch(Ctx* c, const char* what);

// Launch with the programmatic-dependent-launch attribute (common.cuh: pdl_wait / pdl_trigger).  MQB200_PDL=0 launches plainly.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

}  // namespace mq

#define MQ_CTX(c, p)                               \
  mq::Ctx* c = mq::as_ctx(p);                      \
  if (!c) return MQ_INVALID_CONTEXT;
#define MQ_REQUIRE(c, cond, msg)                                         \
  do {                                                                   \
    if (!(cond)) return mq::fail(c, MQ_INVALID_ARGUMENT, std::string(__func__) + ": " + (msg)); \
  } while (0)
