// Optimiser step of the calibration loops: global L2 gradient norm + skip-on-non-finite + AdamW over ONE flat parameter
// buffer (every LET / LWC / LRL learnable of the model is a view into it), two launches per step.
//
// Replaces, per step, NativeScalerWithGradNormCount.__call__ (mobilellm/utils/optim.py:28-41: unscale -> grad norm ->
// GradScaler.step, which skips the update when a gradient is inf / nan) and torch.optim.AdamW.step as the reference
// constructs it (alg:513, 716-722: betas (0.9, 0.999), eps 1e-8, decoupled weight decay).  The library optimiser walks
// ~1500 tiny tensors (0-d scales / offsets, [H] LET vectors): 38 ms of a 94 ms TinyLlama e2e step on a B200; this is ~10 us.
//
//   norm       = sqrt(sum g^2)                              (double accumulation, fixed order: deterministic)
//   if !finite(norm): nothing changes (step counter included)
//   step += 1;  p -= lr_g*wd*p;  m += (1-b1)(g-m);  v = b2 v + (1-b2) g^2
//   p -= (lr_g / (1-b1^step)) * m / (sqrt(v)/sqrt(1-b2^step) + eps)          lr_g: learning rate of the element's group
#include "common.cuh"
#include "ctx.h"
#include <string>

namespace mq {

constexpr int kOptBlocks = 296;        // 2 per SM
constexpr int kOptThreads = 256;
constexpr int kOptMaxGroups = 8;

struct AdamArgs {
  float* p; const float* g; float* m; float* v;
  int64_t n;
  int ngroups;
  int64_t seg_end[kOptMaxGroups];      // element index one past the end of each learning-rate group (ascending)
  const float* lr;                     // [ngroups] device
  float beta1, beta2, eps, wd;
  float* state;                        // device [8]: 0 step, 1 norm, 2 found_inf, 3 1-b1^step, 4 sqrt(1-b2^step), 5 skipped steps
  double* partial;                     // [kOptBlocks] workspace
  unsigned* ticket;                    // self-resetting arrival counter
};

__global__ void __launch_bounds__(kOptThreads) adam_norm_kernel(const AdamArgs a) {
  __shared__ double s_red[kOptThreads / 32];
  __shared__ bool s_last;
  double acc = 0.0;
  const int64_t n4 = a.n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(a.g);
  for (int64_t i = int64_t(blockIdx.x) * kOptThreads + threadIdx.x; i < n4; i += int64_t(gridDim.x) * kOptThreads) {
    const float4 x = __ldg(g4 + i);
    acc += (double)x.x * x.x + (double)x.y * x.y + (double)x.z * x.z + (double)x.w * x.w;
  }
  if (blockIdx.x == 0) for (int64_t i = (n4 << 2) + threadIdx.x; i < a.n; i += kOptThreads) acc += (double)a.g[i] * a.g[i];
  acc = warp_reduce(acc, OpSum{});
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kOptThreads / 32; ++i) t += s_red[i];
    a.partial[blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  // last block: fold the partials in block order, finish the step bookkeeping
  if (threadIdx.x == 0) {
    __threadfence();
    double t = 0.0;
    for (unsigned i = 0; i < gridDim.x; ++i) t += reinterpret_cast<volatile double*>(a.partial)[i];
    const float norm = (float)sqrt(t);
    const bool bad = !isfinite(norm);
    a.state[1] = norm;
    a.state[2] = bad ? 1.f : 0.f;
    if (bad) {
      a.state[5] += 1.f;
    } else {
      const float step = a.state[0] + 1.f;
      a.state[0] = step;
      a.state[3] = (float)(1.0 - pow((double)a.beta1, (double)step));
      a.state[4] = (float)sqrt(1.0 - pow((double)a.beta2, (double)step));
    }
    *a.ticket = 0u;
  }
}

__global__ void __launch_bounds__(kOptThreads) adam_update_kernel(const AdamArgs a) {
  if (a.state[2] != 0.f) return;                          // GradScaler semantics: inf / nan gradient -> the step is skipped
  const float bc1 = a.state[3], bc2s = a.state[4];
  const float omb1 = 1.f - a.beta1, omb2 = 1.f - a.beta2;
  float lr[kOptMaxGroups];
#pragma unroll
  for (int i = 0; i < kOptMaxGroups; ++i) lr[i] = i < a.ngroups ? __ldg(a.lr + i) : 0.f;
  for (int64_t i = int64_t(blockIdx.x) * kOptThreads + threadIdx.x; i < a.n; i += int64_t(gridDim.x) * kOptThreads) {
    float l = lr[0];
#pragma unroll
    for (int k = 1; k < kOptMaxGroups; ++k) if (k < a.ngroups && i >= a.seg_end[k - 1]) l = lr[k];
    const float g = a.g[i];
    float p = a.p[i], m = a.m[i], v = a.v[i];
    if (a.wd != 0.f) p = fsub(p, fmul(fmul(l, a.wd), p));
    m = fadd(m, fmul(omb1, fsub(g, m)));                   // exp_avg.lerp_(grad, 1 - beta1)
    v = fadd(fmul(a.beta2, v), fmul(omb2, fmul(g, g)));    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    const float denom = fadd(fdiv(__fsqrt_rn(v), bc2s), a.eps);
    p = fsub(p, fmul(fdiv(l, bc1), fdiv(m, denom)));       // param.addcdiv_(exp_avg, denom, value = -lr / bias_correction1)
    a.p[i] = p; a.m[i] = m; a.v[i] = v;
  }
}

}  // namespace mq

using namespace mq;

extern "C" int mq_adamw_step(void* ctx, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int ngroups,
                             const int64_t* seg_end, const float* lr, float beta1, float beta2, float eps, float weight_decay,
                             float* state, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, params && grads && exp_avg && exp_avg_sq && lr && state && seg_end && n > 0, "null pointer or empty parameter buffer");
  MQ_REQUIRE(c, ngroups >= 1 && ngroups <= kOptMaxGroups, "1..8 learning-rate groups");
  MQ_REQUIRE(c, (reinterpret_cast<uintptr_t>(grads) & 15) == 0, "the gradient buffer must be 16-byte aligned");
  MQ_REQUIRE(c, size_t(kOptBlocks) * sizeof(double) + 16 <= c->ws_bytes, "workspace too small");
  AdamArgs a;
  a.p = params; a.g = grads; a.m = exp_avg; a.v = exp_avg_sq; a.n = n; a.ngroups = ngroups;
  int64_t prev = 0;
  for (int i = 0; i < kOptMaxGroups; ++i) {
    a.seg_end[i] = i < ngroups ? seg_end[i] : n;
    MQ_REQUIRE(c, i >= ngroups || (seg_end[i] >= prev && seg_end[i] <= n), "seg_end must be ascending and <= n");
    if (i < ngroups) prev = seg_end[i];
  }
  MQ_REQUIRE(c, seg_end[ngroups - 1] == n, "the last group must end at n");
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay; a.state = state;
  // the optimiser owns the tail of the context's counter array (the skinny-GEMM epilogue uses the head) and the workspace
  // for the duration of its two launches (single stream per context, include/mqb200.h "Threading")
  a.partial = reinterpret_cast<double*>(stream_ws(c, (cudaStream_t)stream));
  if (!a.partial) return MQ_FAILED_ALLOCATION;
  a.ticket = reinterpret_cast<unsigned*>(c->counters + c->n_counters - 1);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t want = (n / 4 + kOptThreads - 1) / kOptThreads;
  const int blocks = (int)(want < 1 ? 1 : (want > kOptBlocks ? kOptBlocks : want));
  adam_norm_kernel<<<blocks, kOptThreads, 0, st>>>(a);
  adam_update_kernel<<<blocks, kOptThreads, 0, st>>>(a);
  return check_launch(c, "mq_adamw_step");
}
