// Calibration attention core between the two batched matmuls (hm:514-534 with the QMatMul quantizers of qm:453-466):
//   A = softmax( fq_1(S) / sqrt(hd) + causal_mask ),  P^ = fq_2(A)
// fq_1 = qk_bmm.output_quantizer (on the raw scores), fq_2 = pv_bmm.input_quantizer (on the probabilities).  The reference
// runs this as ~8 element-wise launches forward and ~12 backward over [B, nh, T, T] fp32 tensors (134 MB each at T 1024);
// here the forward is ONE pass (read S below the diagonal, write P^) and the backward ONE pass (read dP^ and S below the
// diagonal, write dS), the row statistics (max, sum) being the only thing kept in between.  One warp per row, the row
// lives in registers.  Arithmetic is the reference's op for op (explicit _rn intrinsics; divisions by a launch- or
// row-uniform divisor through the exact div_rn of common.cuh); the two quantizers' scale / offset gradients (LRL) are
// accumulated in the same pass and folded in fixed block order by the last block to arrive (deterministic).
// Columns above the diagonal: A == 0 exactly (exp(finfo.min - max) underflows to 0 as in the reference), so P^ = fq_2(0),
// dS == 0, and their fq_2 gradient terms vanish whenever 0 is inside fq_2's code range -- they are skipped then, and
// processed like every other column otherwise.
#include "common.cuh"
#include "ctx.h"
#include "fq_math.cuh"

namespace mq {

struct FqP {
  float s, rs, o, qmin, qmax;
  bool five, on;
};
__device__ __forceinline__ FqP load_fqp(const float* scale, const float* offset, float qmin, float qmax) {
  FqP q;
  q.on = scale != nullptr;
  q.s = q.on ? __ldg(scale) : 1.f;
  q.o = q.on ? __ldg(offset) : 0.f;
  q.rs = __frcp_rn(q.s);
  q.five = mantissa_all_ones(q.s);
  q.qmin = qmin; q.qmax = qmax;
  return q;
}
__device__ __forceinline__ float fq_apply(float x, const FqP& q) {
  if (!q.on) return x;
  const float c = q.five ? quant_code_v<true>(x, q.s, q.rs, q.o, q.qmin, q.qmax) : quant_code_v<false>(x, q.s, q.rs, q.o, q.qmin, q.qmax);
  return dequant(c, q.s, q.o);
}
__device__ __forceinline__ FqGrad fq_grad(float x, float g, const FqP& q) {
  if (!q.on) { FqGrad r; r.gx = g; r.gs = 0.f; r.go = 0.f; return r; }
  return q.five ? fq_bwd_elem_v<true>(x, g, q.s, q.rs, q.o, q.qmin, q.qmax) : fq_bwd_elem_v<false>(x, g, q.s, q.rs, q.o, q.qmin, q.qmax);
}

struct ProbArgs {
  const float* S; float* P; float* stats;        // stats[rows][2] = (row max of the scaled scores, sum of exp)
  int64_t rows; int T, Tq, causal;
  const float *s1, *o1; float qmin1, qmax1;
  const float *s2, *o2; float qmin2, qmax2;
  float mul;                                     // fp32 reciprocal of sqrt(hd): ATen divides by a scalar as x * (1/d)
  // backward only
  const float* g; float* dS;
  double* partial; unsigned* ticket; float* gout;  // gout[4] = d/ds1, d/do1, d/ds2, d/do2
};

constexpr int kRowsPerCta = 8;

template <int NV>
__global__ void __launch_bounds__(256) attn_probs_fwd_kernel(const ProbArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * kRowsPerCta + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const int T = a.T;
  const int ncol = a.causal ? int(row % a.Tq) + 1 + (T - a.Tq) : T;
  const FqP q1 = load_fqp(a.s1, a.o1, a.qmin1, a.qmax1), q2 = load_fqp(a.s2, a.o2, a.qmin2, a.qmax2);
  const float* sr = a.S + row * T;
  float v[NV][4];
  float mx = -FLT_MAX;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int k0 = j * 128 + lane * 4;
    if (k0 < ncol) {
      const float4 x = ldg4_stream(sr + k0);
      const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[j][e] = fmul(fq_apply(xv[e], q1), a.mul);
        if (k0 + e < ncol) mx = fmaxf(mx, v[j][e]);
      }
    }
  }
  mx = warp_reduce(mx, OpFMax());
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int k0 = j * 128 + lane * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[j][e] = (k0 + e < ncol) ? expf(fsub(v[j][e], mx)) : 0.f;
      sum += v[j][e];
    }
  }
  sum = warp_reduce(sum, OpSum());
  const RowDiv rd = make_rowdiv(sum);
  const float pz = fq_apply(0.f, q2);
  float* pr = a.P + row * T;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int k0 = j * 128 + lane * 4;
    if (k0 >= T) continue;
    float o4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o4[e] = (k0 + e < ncol) ? fq_apply(div_any(v[j][e], rd), q2) : pz;
    *reinterpret_cast<float4*>(pr + k0) = make_float4(o4[0], o4[1], o4[2], o4[3]);
  }
  if (lane == 0) { a.stats[2 * row] = mx; a.stats[2 * row + 1] = sum; }
}

template <int NV>
__global__ void __launch_bounds__(256) attn_probs_bwd_kernel(const ProbArgs a) {
  __shared__ float red[32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31;
  const int T = a.T;
  const FqP q1 = load_fqp(a.s1, a.o1, a.qmin1, a.qmax1), q2 = load_fqp(a.s2, a.o2, a.qmin2, a.qmax2);
  // fq_2 at a masked column: x = 0 -> t3 = o2; its (gs, go) terms are identically 0 iff o2 is inside the code range
  const bool neutral = !q2.on || (q2.o >= q2.qmin && q2.o <= q2.qmax);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};            // gs1, go1, gs2, go2
  for (int64_t row = int64_t(blockIdx.x) * kRowsPerCta + (threadIdx.x >> 5); row < a.rows; row += int64_t(gridDim.x) * kRowsPerCta) {
    const int ncol = a.causal ? int(row % a.Tq) + 1 + (T - a.Tq) : T;
    const int nproc = neutral ? ncol : T;
    const float* sr = a.S + row * T;
    const float* gr = a.g + row * T;
    const float mx = __ldg(a.stats + 2 * row);
    const RowDiv rd = make_rowdiv(__ldg(a.stats + 2 * row + 1));
    float p[NV][4], gp[NV][4];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int k0 = j * 128 + lane * 4;
      if (k0 < nproc) {
        const float4 g4 = ldg4_stream(gr + k0);
        const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
        float xv[4] = {0.f, 0.f, 0.f, 0.f};
        if (k0 < ncol) { const float4 x = ldg4(sr + k0); xv[0] = x.x; xv[1] = x.y; xv[2] = x.z; xv[3] = x.w; }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float pe = 0.f;
          if (k0 + e < ncol) pe = div_any(expf(fsub(fmul(fq_apply(xv[e], q1), a.mul), mx)), rd);
          const FqGrad r = fq_grad(pe, gv[e], q2);
          p[j][e] = pe; gp[j][e] = r.gx;
          acc[2] += r.gs; acc[3] += r.go;
          dot = fmaf(pe, r.gx, dot);
        }
      }
    }
    dot = warp_reduce(dot, OpSum());
    float* dr = a.dS + row * T;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int k0 = j * 128 + lane * 4;
      if (k0 >= T) continue;
      float d4[4] = {0.f, 0.f, 0.f, 0.f};
      if (k0 < ncol) {
        const float4 x = ldg4(sr + k0);
        const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (k0 + e < ncol) {
            const float da = fmul(fsub(gp[j][e], dot), p[j][e]);          // softmax backward: (g - sum(g*y)) * y
            const FqGrad r = fq_grad(xv[e], fmul(da, a.mul), q1);
            d4[e] = r.gx; acc[0] += r.gs; acc[1] += r.go;
          }
        }
      }
      *reinterpret_cast<float4*>(dr + k0) = make_float4(d4[0], d4[1], d4[2], d4[3]);
    }
  }
  if (!a.gout) return;
  float b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = block_reduce(acc[i], OpSum(), red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) a.partial[4 * blockIdx.x + i] = b[i];
    __threadfence();
    s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence();
    const volatile double* vp = a.partial;
    double t[4] = {0., 0., 0., 0.};
    for (unsigned i = threadIdx.x; i < gridDim.x; i += 32)
#pragma unroll
      for (int k = 0; k < 4; ++k) t[k] += vp[4 * i + k];
#pragma unroll
    for (int k = 0; k < 4; ++k) t[k] = warp_reduce(t[k], OpSum());
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) a.gout[k] = (float)t[k];
      *a.ticket = 0u;
    }
  }
}

static int probs_nv(int T) {
  if (T % 4 != 0 || T < 4) return 0;
  if (T <= 256) return 2;
  if (T <= 512) return 4;
  if (T <= 1024) return 8;
  if (T <= 2048) return 16;
  return 0;
}

}  // namespace mq

using namespace mq;

extern "C" {

int mq_attn_probs_supported(int T) { return probs_nv(T) != 0; }

int mq_attn_probs_fwd(void* ctx, const float* S, float* P, float* stats, int64_t rows, int T, int Tq, int causal, float mul,
                      const float* s1, const float* o1, float qmin1, float qmax1, const float* s2, const float* o2, float qmin2,
                      float qmax2, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, S && P && stats && rows >= 0 && Tq > 0 && Tq <= T, "null pointer or bad shape");
  MQ_REQUIRE(c, (s1 == nullptr) == (o1 == nullptr) && (s2 == nullptr) == (o2 == nullptr), "scale and offset come in pairs");
  const int nv = probs_nv(T);
  MQ_REQUIRE(c, nv != 0, "T must be a multiple of 4 and <= 2048");
  MQ_REQUIRE(c, ((reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(P)) & 15) == 0, "S / P must be 16-byte aligned");
  if (rows == 0) return MQ_NO_ERROR;
  ProbArgs a{};
  a.S = S; a.P = P; a.stats = stats; a.rows = rows; a.T = T; a.Tq = Tq; a.causal = causal; a.mul = mul;
  a.s1 = s1; a.o1 = o1; a.qmin1 = qmin1; a.qmax1 = qmax1; a.s2 = s2; a.o2 = o2; a.qmin2 = qmin2; a.qmax2 = qmax2;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((rows + kRowsPerCta - 1) / kRowsPerCta);
  switch (nv) {
    case 2: attn_probs_fwd_kernel<2><<<grid, 256, 0, st>>>(a); break;
    case 4: attn_probs_fwd_kernel<4><<<grid, 256, 0, st>>>(a); break;
    case 8: attn_probs_fwd_kernel<8><<<grid, 256, 0, st>>>(a); break;
    default: attn_probs_fwd_kernel<16><<<grid, 256, 0, st>>>(a); break;
  }
  return check_launch(c, "mq_attn_probs_fwd");
}

int mq_attn_probs_bwd(void* ctx, const float* S, const float* stats, const float* g, float* dS, int64_t rows, int T, int Tq,
                      int causal, float mul, const float* s1, const float* o1, float qmin1, float qmax1, const float* s2,
                      const float* o2, float qmin2, float qmax2, float* gparams, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, S && stats && g && dS && rows >= 0 && Tq > 0 && Tq <= T, "null pointer or bad shape");
  MQ_REQUIRE(c, (s1 == nullptr) == (o1 == nullptr) && (s2 == nullptr) == (o2 == nullptr), "scale and offset come in pairs");
  const int nv = probs_nv(T);
  MQ_REQUIRE(c, nv != 0, "T must be a multiple of 4 and <= 2048");
  MQ_REQUIRE(c, ((reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(dS)) & 15) == 0,
             "S / g / dS must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    if (gparams) cudaMemsetAsync(gparams, 0, 4 * sizeof(float), st);
    return MQ_NO_ERROR;
  }
  ProbArgs a{};
  a.S = S; a.stats = const_cast<float*>(stats); a.rows = rows; a.T = T; a.Tq = Tq; a.causal = causal; a.mul = mul;
  a.s1 = s1; a.o1 = o1; a.qmin1 = qmin1; a.qmax1 = qmax1; a.s2 = s2; a.o2 = o2; a.qmin2 = qmin2; a.qmax2 = qmax2;
  a.g = g; a.dS = dS; a.gout = gparams;
  const int64_t need = (rows + kRowsPerCta - 1) / kRowsPerCta;
  const int64_t cap = int64_t(c->sm_count) * (nv <= 8 ? 2 : 1) * 4;
  const unsigned grid = (unsigned)(need < cap ? need : cap);
  if (gparams) {
    void* wsp = stream_ws(c, st);
    if (!wsp) return MQ_FAILED_ALLOCATION;
    a.partial = reinterpret_cast<double*>(wsp);
    a.ticket = reinterpret_cast<unsigned*>(static_cast<char*>(wsp) + c->ws_bytes - 64);
  }
  switch (nv) {
    case 2: attn_probs_bwd_kernel<2><<<grid, 256, 0, st>>>(a); break;
    case 4: attn_probs_bwd_kernel<4><<<grid, 256, 0, st>>>(a); break;
    case 8: attn_probs_bwd_kernel<8><<<grid, 256, 0, st>>>(a); break;
    default: attn_probs_bwd_kernel<16><<<grid, 256, 0, st>>>(a); break;
  }
  return check_launch(c, "mq_attn_probs_bwd");
}

}  // extern "C"
