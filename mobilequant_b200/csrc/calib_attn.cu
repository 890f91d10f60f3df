// Calibration attention core between the two batched matmuls (hm:514-534 with the QMatMul quantizers of qm:453-466):
//   A = softmax( fq_1(S) / sqrt(hd) + causal_mask ),  P^ = fq_2(A)
// fq_1 = qk_bmm.output_quantizer (on the raw scores), fq_2 = pv_bmm.input_quantizer (on the probabilities).  The reference
// runs this as ~8 element-wise launches forward and ~12 backward over [B, nh, T, T] fp32 tensors (134 MB each at T 1024);
// here the forward is ONE pass (read S below the diagonal, write P^) and the backward ONE pass (read dP^ and S below the
// diagonal, write dS), the row statistics (max, sum) being the only thing kept in between.  One warp per row, the row
// lives in registers.  The forward arithmetic is the reference's op for op with explicit _rn intrinsics (x / scale through the
// exact div_rn of common.cuh), except that a probability is exp * RN(1 / sum) instead of exp / sum (<= 1 ulp apart; forward
// and backward use the same value).  The backward evaluates the quantizer gradients in the cancellation-free form of
// fq_math.cuh (fq_bwd_elem_c); both quantizers' scale / offset gradients (LRL) are accumulated in the same pass and folded in
// fixed block order by the last block to arrive (deterministic).
// Columns above the diagonal: A == 0 exactly (exp(finfo.min - max) underflows to 0 as in the reference), so P^ = fq_2(0),
// dS == 0, and their fq_2 gradient terms vanish whenever 0 is inside fq_2's code range -- they are skipped then, and
// processed like every other column otherwise.
#include "common.cuh"
#include "ctx.h"
#include "fq_math.cuh"

namespace mq {

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
constexpr float kNegInf = -INFINITY;               // masked score: exp(-inf - max) == 0 exactly, like exp(finfo.min - max)

template <int NV>
__global__ void __launch_bounds__(256, NV <= 8 ? 4 : 2) attn_probs_fwd_kernel(const ProbArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * kRowsPerCta + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const int T = a.T;
  const int ncol = a.causal ? int(row % a.Tq) + 1 + (T - a.Tq) : T;
  const FqP q1 = load_fqp(a.s1, a.o1, a.qmin1, a.qmax1), q2 = load_fqp(a.s2, a.o2, a.qmin2, a.qmax2);
  const float* sr = a.S + row * T;
  float* pr = a.P + row * T;
  float mx, sum;
  auto body = [&](auto five_tag) {
    constexpr bool FIVE = decltype(five_tag)::value;
    float v[NV][4];
    mx = kNegInf;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int k0 = j * 128 + lane * 4;
      v[j][0] = v[j][1] = v[j][2] = v[j][3] = kNegInf;
      if (k0 < ncol) {
        const float4 x = ldg4_stream(sr + k0);
        const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float t = fmul(fq_apply<FIVE>(xv[e], q1), a.mul);
          v[j][e] = (k0 + e < ncol) ? t : kNegInf;
          mx = fmaxf(mx, v[j][e]);
        }
      }
    }
    mx = warp_reduce(mx, OpFMax());
    sum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (j * 128 + lane * 4 >= ncol) continue;                     // fully masked chunk: exp(-inf) == 0
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[j][e] = expf(fsub(v[j][e], mx));
        sum += v[j][e];
      }
    }
    sum = warp_reduce(sum, OpSum());
    const float rsum = __frcp_rn(sum);
    const float pz = fq_apply<FIVE>(0.f, q2);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int k0 = j * 128 + lane * 4;
      if (k0 >= T) continue;
      float o4[4] = {pz, pz, pz, pz};
      if (k0 < ncol) {
#pragma unroll
        for (int e = 0; e < 4; ++e) o4[e] = fq_apply<FIVE>(fmul(v[j][e], rsum), q2);
      }
      *reinterpret_cast<float4*>(pr + k0) = make_float4(o4[0], o4[1], o4[2], o4[3]);
    }
  };
  dispatch_five(q1.five | q2.five, body);
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
  auto body = [&](auto five_tag) {
    constexpr bool FIVE = decltype(five_tag)::value;
    for (int64_t row = int64_t(blockIdx.x) * kRowsPerCta + (threadIdx.x >> 5); row < a.rows; row += int64_t(gridDim.x) * kRowsPerCta) {
      const int ncol = a.causal ? int(row % a.Tq) + 1 + (T - a.Tq) : T;
      const int nproc = neutral ? ncol : T;
      const float* sr = a.S + row * T;
      const float* gr = a.g + row * T;
      const float mx = __ldg(a.stats + 2 * row);
      const float rsum = __frcp_rn(__ldg(a.stats + 2 * row + 1));
      // With y = softmax row, h = dL/dy (after fq_2's backward) and dot = sum(y*h), the gradient that reaches fq_1's output is
      // g1_i = mul * (y_i*h_i - dot*y_i).  Everything fq_1's backward needs from element i besides g1_i is known in this
      // pass -- c_i = dy/dscale = t5 - m*u and the in-range bit m -- so its row sums are accumulated as
      //   sum g1_i*c_i = mul * (A - dot*B),  A = sum y*h*c, B = sum y*c     and   sum_{!m} g1_i = mul * (C - dot*D)
      // and the second pass only scales: dS_i = m ? g1_i : 0 from the stashed y_i*h_i and y_i (both zeroed where !m).
      float pm[NV][4], wm[NV][4];
      float dot = 0.f, A = 0.f, Bq = 0.f, C = 0.f, D = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int k0 = j * 128 + lane * 4;
        pm[j][0] = pm[j][1] = pm[j][2] = pm[j][3] = 0.f;
        wm[j][0] = wm[j][1] = wm[j][2] = wm[j][3] = 0.f;
        if (k0 < nproc) {
          const float4 g4 = ldg4_stream(gr + k0);
          const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
          float xv[4] = {0.f, 0.f, 0.f, 0.f};
          if (k0 < ncol) { const float4 x = ldg4_stream(sr + k0); xv[0] = x.x; xv[1] = x.y; xv[2] = x.z; xv[3] = x.w; }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float sq = xv[e], c1 = 0.f;
            bool m1 = true;
            if (q1.on) {
              const float u = div_rn<FIVE>(xv[e], q1.s, q1.rs);
              const float t3 = fadd(rne_magic(u), q1.o);
              m1 = (t3 >= q1.qmin) && (t3 <= q1.qmax);
              const float t5 = fsub(fminf(fmaxf(t3, q1.qmin), q1.qmax), q1.o);
              sq = fmul(t5, q1.s);
              c1 = fsub(t5, m1 ? u : 0.f);
            }
            const float t = (k0 + e < ncol) ? fmul(sq, a.mul) : kNegInf;
            const float pe = fmul(expf(fsub(t, mx)), rsum);
            const FqGrad r = fq_grad<FIVE>(pe, gv[e], q2);
            acc[2] += r.gs; acc[3] += r.go;
            const float w = fmul(pe, r.gx);
            dot += w;
            A = fmaf(w, c1, A); Bq = fmaf(pe, c1, Bq);
            C += m1 ? 0.f : w; D += m1 ? 0.f : pe;
            pm[j][e] = m1 ? pe : 0.f; wm[j][e] = m1 ? w : 0.f;
          }
        }
      }
      dot = warp_reduce(dot, OpSum());
      acc[0] += fmul(a.mul, fmaf(-dot, Bq, A));
      acc[1] -= fmul(fmul(q1.s, a.mul), fmaf(-dot, D, C));
      float* dr = a.dS + row * T;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int k0 = j * 128 + lane * 4;
        if (k0 >= T) continue;
        float d4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) d4[e] = fmul(a.mul, fmaf(-dot, pm[j][e], wm[j][e]));
        *reinterpret_cast<float4*>(dr + k0) = make_float4(d4[0], d4[1], d4[2], d4[3]);
      }
    }
  };
  dispatch_five(q1.five | q2.five, body);
  if (a.gout) grid_fold<4>(acc, a.partial, a.ticket, a.gout, red, &s_last);
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
