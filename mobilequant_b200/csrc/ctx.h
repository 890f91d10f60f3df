// Context object behind the opaque void* of include/mqb200.h.
// Conventions of capp/src/libllmod.cpp:23-65 (magic + version + refcount) re-designed for a CUDA library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <atomic>
#include "mqb200.h"

namespace mq {

constexpr uint32_t kMagic = 0x4d514232u;  // "MQB2"

struct Ctx {
  uint32_t magic = kMagic;
  std::atomic<int> refs{1};
  int device = 0;
  int sm_count = 148;
  void* ws = nullptr;      // device scratch (deterministic two-stage reductions, split-K partials)
  size_t ws_bytes = 0;
  int* counters = nullptr; // zero-initialised arrival counters of the fused skinny-GEMM epilogue (self-resetting)
  int n_counters = 0;
  std::string last_error[6];
};

inline Ctx* as_ctx(void* p) {
  Ctx* c = reinterpret_cast<Ctx*>(p);
  return (c && c->magic == kMagic) ? c : nullptr;
}

int fail(Ctx* c, int code, const std::string& what);
int check_launch(Ctx* c, const char* what);

}  // namespace mq

#define MQ_CTX(c, p)                               \
  mq::Ctx* c = mq::as_ctx(p);                      \
  if (!c) return MQ_INVALID_CONTEXT;
#define MQ_REQUIRE(c, cond, msg)                                         \
  do {                                                                   \
    if (!(cond)) return mq::fail(c, MQ_INVALID_ARGUMENT, std::string(__func__) + ": " + (msg)); \
  } while (0)
