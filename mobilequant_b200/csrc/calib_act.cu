// Element-wise pieces of a decoder block in calibration mode, each ONE kernel forward and ONE backward instead of the chain of
// ATen launches the module graph produces (the quantizers' scale / offset gradients come out of the same backward pass,
// folded deterministically; arithmetic as in fq_math.cuh):
//   * gated SiLU MLP core (hm:1042-1062 with QSiLU qm:691-753 and w2.input_quantizer):
//       out = fq_w( fq_o( a * fq_s(sigmoid(a)) ) * b ),   a = w1(x) (already quantised by w1), b = w3(x)
#include "common.cuh"
#include "ctx.h"
#include "fq_math.cuh"

namespace mq {

static inline int grid_for_elems(Ctx* c, int64_t n4, int waves) {
  int64_t need = (n4 + 255) / 256, cap = int64_t(c->sm_count) * waves;
  if (need < 1) need = 1;
  return int(need < cap ? need : cap);
}

struct GateArgs {
  const float *a, *b; float* out; int64_t n;
  const float *s_s, *o_s; float qmin_s, qmax_s;      // fq_s: QSiLU.input2_quantizer (on sigmoid(a))
  const float *s_o, *o_o; float qmin_o, qmax_o;      // fq_o: QSiLU.output_quantizer
  const float *s_w, *o_w; float qmin_w, qmax_w;      // fq_w: w2.input_quantizer
  const float* g; float *da, *db;
  double* partial; unsigned* ticket; float* gout;    // gout[6] = d/d(scale, offset) of fq_s, fq_o, fq_w
};

// ATen: sigmoid(x) = 1 / (1 + exp(-x)) in fp32
__device__ __forceinline__ float sigmoid_rn(float x) { return __frcp_rn(fadd(1.f, expf(-x))); }

__global__ void __launch_bounds__(256) silu_gate_fwd_kernel(const GateArgs p) {
  const FqP qs = load_fqp(p.s_s, p.o_s, p.qmin_s, p.qmax_s), qo = load_fqp(p.s_o, p.o_o, p.qmin_o, p.qmax_o),
            qw = load_fqp(p.s_w, p.o_w, p.qmin_w, p.qmax_w);
  const int64_t n4 = p.n >> 2, nthr = int64_t(gridDim.x) * blockDim.x;
  auto body = [&](auto five_tag) {
    constexpr bool FIVE = decltype(five_tag)::value;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += nthr) {
      const float4 a4 = ldg4_stream(p.a + (i << 2)), b4 = ldg4_stream(p.b + (i << 2));
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float y = fq_apply<FIVE>(sigmoid_rn(av[e]), qs);
        const float h = fq_apply<FIVE>(fmul(av[e], y), qo);
        o[e] = fq_apply<FIVE>(fmul(h, bv[e]), qw);
      }
      *reinterpret_cast<float4*>(p.out + (i << 2)) = make_float4(o[0], o[1], o[2], o[3]);
    }
  };
  dispatch_five(qs.five | qo.five | qw.five, body);
}

__global__ void __launch_bounds__(256) silu_gate_bwd_kernel(const GateArgs p) {
  __shared__ float red[32];
  __shared__ bool s_last;
  const FqP qs = load_fqp(p.s_s, p.o_s, p.qmin_s, p.qmax_s), qo = load_fqp(p.s_o, p.o_o, p.qmin_o, p.qmax_o),
            qw = load_fqp(p.s_w, p.o_w, p.qmin_w, p.qmax_w);
  const int64_t n4 = p.n >> 2, nthr = int64_t(gridDim.x) * blockDim.x;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto body = [&](auto five_tag) {
    constexpr bool FIVE = decltype(five_tag)::value;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += nthr) {
      const float4 a4 = ldg4_stream(p.a + (i << 2)), b4 = ldg4_stream(p.b + (i << 2)), g4 = ldg4_stream(p.g + (i << 2));
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w};
      float da[4], db[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float sg = sigmoid_rn(av[e]);
        const float y = fq_apply<FIVE>(sg, qs);
        const float hh = fmul(av[e], y);
        const float h = fq_apply<FIVE>(hh, qo);
        const FqGrad rw = fq_grad<FIVE>(fmul(h, bv[e]), gv[e], qw);          // through w2.input_quantizer
        db[e] = fmul(rw.gx, h);
        const FqGrad ro = fq_grad<FIVE>(hh, fmul(rw.gx, bv[e]), qo);         // through QSiLU.output_quantizer
        const FqGrad rs = fq_grad<FIVE>(sg, fmul(ro.gx, av[e]), qs);         // through QSiLU.input2_quantizer
        // a feeds the product directly and through the sigmoid (ATen sigmoid_backward: g * (1 - y) * y)
        da[e] = fadd(fmul(ro.gx, y), fmul(fmul(rs.gx, fsub(1.f, sg)), sg));
        acc[0] += rs.gs; acc[1] += rs.go; acc[2] += ro.gs; acc[3] += ro.go; acc[4] += rw.gs; acc[5] += rw.go;
      }
      *reinterpret_cast<float4*>(p.da + (i << 2)) = make_float4(da[0], da[1], da[2], da[3]);
      *reinterpret_cast<float4*>(p.db + (i << 2)) = make_float4(db[0], db[1], db[2], db[3]);
    }
  };
  dispatch_five(qs.five | qo.five | qw.five, body);
  if (p.gout) grid_fold<6>(acc, p.partial, p.ticket, p.gout, red, &s_last);
}

}  // namespace mq

using namespace mq;

extern "C" {

static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

int mq_silu_gate_fwd(void* ctx, const float* a, const float* b, float* out, int64_t n, const float* const* scales,
                     const float* const* offsets, const float* qmins, const float* qmaxs, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, a && b && out && n >= 0 && n % 4 == 0 && scales && offsets && qmins && qmaxs, "null pointer or n % 4 != 0");
  MQ_REQUIRE(c, al16(a) && al16(b) && al16(out), "a / b / out must be 16-byte aligned");
  for (int i = 0; i < 3; ++i) MQ_REQUIRE(c, (scales[i] == nullptr) == (offsets[i] == nullptr), "scale and offset come in pairs");
  if (n == 0) return MQ_NO_ERROR;
  GateArgs p{};
  p.a = a; p.b = b; p.out = out; p.n = n;
  p.s_s = scales[0]; p.o_s = offsets[0]; p.qmin_s = qmins[0]; p.qmax_s = qmaxs[0];
  p.s_o = scales[1]; p.o_o = offsets[1]; p.qmin_o = qmins[1]; p.qmax_o = qmaxs[1];
  p.s_w = scales[2]; p.o_w = offsets[2]; p.qmin_w = qmins[2]; p.qmax_w = qmaxs[2];
  silu_gate_fwd_kernel<<<grid_for_elems(c, n / 4, 8), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch(c, "mq_silu_gate_fwd");
}

int mq_silu_gate_bwd(void* ctx, const float* a, const float* b, const float* g, float* da, float* db, int64_t n,
                     const float* const* scales, const float* const* offsets, const float* qmins, const float* qmaxs,
                     float* gparams, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, a && b && g && da && db && n >= 0 && n % 4 == 0 && scales && offsets && qmins && qmaxs, "null pointer or n % 4 != 0");
  MQ_REQUIRE(c, al16(a) && al16(b) && al16(g) && al16(da) && al16(db), "tensors must be 16-byte aligned");
  for (int i = 0; i < 3; ++i) MQ_REQUIRE(c, (scales[i] == nullptr) == (offsets[i] == nullptr), "scale and offset come in pairs");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    if (gparams) cudaMemsetAsync(gparams, 0, 6 * sizeof(float), st);
    return MQ_NO_ERROR;
  }
  GateArgs p{};
  p.a = a; p.b = b; p.g = g; p.da = da; p.db = db; p.n = n; p.gout = gparams;
  p.s_s = scales[0]; p.o_s = offsets[0]; p.qmin_s = qmins[0]; p.qmax_s = qmaxs[0];
  p.s_o = scales[1]; p.o_o = offsets[1]; p.qmin_o = qmins[1]; p.qmax_o = qmaxs[1];
  p.s_w = scales[2]; p.o_w = offsets[2]; p.qmin_w = qmins[2]; p.qmax_w = qmaxs[2];
  if (gparams) {
    void* wsp = stream_ws(c, st);
    if (!wsp) return MQ_FAILED_ALLOCATION;
    p.partial = reinterpret_cast<double*>(wsp);
    p.ticket = reinterpret_cast<unsigned*>(static_cast<char*>(wsp) + c->ws_bytes - 64);
  }
  silu_gate_bwd_kernel<<<grid_for_elems(c, n / 4, 4), 256, 0, st>>>(p);
  return check_launch(c, "mq_silu_gate_bwd");
}

}  // extern "C"
