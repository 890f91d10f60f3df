// Context management + error reporting of libmqb200 (include/mqb200.h).
#include "ctx.h"
#include <cstdlib>
#include <new>

namespace mq {

static std::string g_setup_error[6];  // errors raised before a context exists (libllmod.h:126-128)

int fail(Ctx* c, int code, const std::string& what) {
  if (code < 0 || code > MQ_INTERNAL_ERROR) code = MQ_INTERNAL_ERROR;
  if (c) c->last_error[code] = what; else g_setup_error[code] = what;
  return code;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MQB200_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
  return on != 0;
}

void* stream_ws(Ctx* c, cudaStream_t st) {
  std::lock_guard<std::mutex> lock(c->ws_mutex);
  for (auto& e : c->ws_by_stream)
    if (e.first == st) return e.second;
  void* p = nullptr;
  if (c->ws_by_stream.empty()) {
    p = c->ws;                                  // the buffer mq_setup allocated goes to the first stream
  } else {
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(c->device);
    cudaError_t e = cudaMalloc(&p, c->ws_bytes);
    cudaSetDevice(cur);
    if (e != cudaSuccess) {
      cudaGetLastError();
      fail(c, MQ_FAILED_ALLOCATION, std::string("per-stream workspace cudaMalloc: ") + cudaGetErrorString(e) +
                                        " (a stream must be used once outside graph capture before it is captured)");
      return nullptr;
    }
  }
  // the last 64 bytes of every scratch buffer are self-resetting arrival counters (zero between launches)
  cudaMemsetAsync(static_cast<char*>(p) + c->ws_bytes - 64, 0, 64, st);
  c->ws_by_stream.emplace_back(st, p);
  return p;
}

int check_launch(Ctx* c, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(c, MQ_RUNTIME_ERROR, std::string(what) + ": " + cudaGetErrorString(e));
  return MQ_NO_ERROR;
}

}  // namespace mq

extern "C" {

int mq_version(void) { return MQB200_VERSION; }

const char* mq_get_error_description(int errorcode) {
  switch (errorcode) {
    case MQ_NO_ERROR: return "no error";
    case MQ_INVALID_CONTEXT: return "invalid context";
    case MQ_INVALID_ARGUMENT: return "invalid argument";
    case MQ_FAILED_ALLOCATION: return "failed allocation";
    case MQ_RUNTIME_ERROR: return "CUDA runtime error";
    case MQ_INTERNAL_ERROR: return "internal error";
    default: return nullptr;
  }
}

const char* mq_get_last_error_extra_info(int errorcode, void* ctx) {
  if (errorcode < 0 || errorcode > MQ_INTERNAL_ERROR) return nullptr;
  mq::Ctx* c = mq::as_ctx(ctx);
  const std::string& s = c ? c->last_error[errorcode] : mq::g_setup_error[errorcode];
  return s.empty() ? nullptr : s.c_str();
}

int mq_setup(void** ctx, int device) {
  if (!ctx) return mq::fail(nullptr, MQ_INVALID_ARGUMENT, "mq_setup: ctx is NULL");
  *ctx = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    return mq::fail(nullptr, MQ_RUNTIME_ERROR,
                    std::string("mq_setup: no usable CUDA device ") + std::to_string(device) + " (" +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "index out of range") + ")");
  }
  mq::Ctx* c = new (std::nothrow) mq::Ctx();
  if (!c) return mq::fail(nullptr, MQ_FAILED_ALLOCATION, "mq_setup: host allocation failed");
  *ctx = c;
  c->device = device;
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return mq::fail(c, MQ_RUNTIME_ERROR, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
  if (prop.major != 10)
    return mq::fail(c, MQ_RUNTIME_ERROR, "libmqb200 is built for sm_100a only; device is sm_" +
                                             std::to_string(prop.major * 10 + prop.minor));
  c->sm_count = prop.multiProcessorCount;
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(device);
  c->ws_bytes = size_t(64) << 20;
  e = cudaMalloc(&c->ws, c->ws_bytes);
  cudaSetDevice(cur);
  if (e != cudaSuccess) {
    cudaSetDevice(cur);
    c->ws = nullptr;
    return mq::fail(c, MQ_FAILED_ALLOCATION, std::string("workspace cudaMalloc: ") + cudaGetErrorString(e));
  }
  cudaSetDevice(device);
  c->n_counters = 8192;
  e = cudaMalloc(reinterpret_cast<void**>(&c->counters), c->n_counters * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(c->counters, 0, c->n_counters * sizeof(int));
  cudaSetDevice(cur);
  if (e != cudaSuccess) {
    c->counters = nullptr;
    return mq::fail(c, MQ_FAILED_ALLOCATION, std::string("counter cudaMalloc: ") + cudaGetErrorString(e));
  }
  return MQ_NO_ERROR;
}

int mq_ref_context(void* ctx) {
  MQ_CTX(c, ctx);
  c->refs.fetch_add(1);
  return MQ_NO_ERROR;
}

int mq_release(void* ctx) {
  MQ_CTX(c, ctx);
  if (c->refs.fetch_sub(1) == 1) {
    for (auto& e : c->ws_by_stream)
      if (e.second && e.second != c->ws) cudaFree(e.second);
    if (c->ws) cudaFree(c->ws);
    if (c->counters) cudaFree(c->counters);
    c->magic = 0;
    delete c;
  }
  return MQ_NO_ERROR;
}

int mq_device_sm_count(void* ctx, int* out) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, out != nullptr, "out is NULL");
  *out = c->sm_count;
  return MQ_NO_ERROR;
}

}  // extern "C"
