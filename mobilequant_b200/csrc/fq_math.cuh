// Element-wise fake-quant arithmetic shared by the calibration kernels (quant.cu, calib_attn.cu).
#pragma once
#include "common.cuh"

namespace mq {

// ================================================================================================================
// K1 backward.  Element-wise terms exactly as autograd evaluates them for qm:286-290:
//   t5 = clamp(rne(x/s)+o) - o ; g_t1 = g*s*m ; gx = g_t1 / s ; gs = g*t5 - g_t1*((x/s)/s) ; go = g_t1 - g*s
// ================================================================================================================
struct FqGrad { float gx, gs, go; };
__device__ __forceinline__ FqGrad fq_bwd_elem(float x, float g, float s, float o, float qmin, float qmax) {
  float u = fdiv(x, s);
  float t3 = fadd(rintf(u), o);
  bool m = (t3 >= qmin) && (t3 <= qmax);
  float t5 = fsub(fminf(fmaxf(t3, qmin), qmax), o);
  float gs5 = fmul(g, s);
  float gt1 = m ? gs5 : 0.f;
  FqGrad r;
  r.gx = fdiv(gt1, s);
  r.gs = fsub(fmul(g, t5), fmul(gt1, fdiv(u, s)));
  r.go = fsub(gt1, gs5);
  return r;
}

// Division by a value that is uniform over a row / CTA / launch: reciprocal and FIVE variant chosen once (common.cuh div_rn).
struct RowDiv { float r, rr; bool five; };
__device__ __forceinline__ RowDiv make_rowdiv(float r) { RowDiv d; d.r = r; d.rr = __frcp_rn(r); d.five = mantissa_all_ones(r); return d; }
__device__ __forceinline__ float div_any(float a, const RowDiv& d) { return d.five ? div_rn<true>(a, d.r, d.rr) : div_rn<false>(a, d.r, d.rr); }

// rintf(u) for the purposes of a quantizer whose code range lies inside +-2^22: |u| is clamped first, which cannot change
// clamp(rne(u) + o, qmin, qmax) nor the in-range test
__device__ __forceinline__ float rne_magic(float u) {
  const float uc = fminf(fmaxf(u, -4194303.f), 4194303.f);
  return __fsub_rn(__fadd_rn(uc, kRoundMagic), kRoundMagic);
}
template <bool FIVE>
__device__ __forceinline__ float quant_code_v(float x, float s, float rs, float o, float qmin, float qmax) {
  return fminf(fmaxf(fadd(rne_magic(div_rn<FIVE>(x, s, rs)), o), qmin), qmax);
}
template <bool FIVE>
__device__ __forceinline__ FqGrad fq_bwd_elem_v(float x, float g, float s, float rs, float o, float qmin, float qmax) {
  const float u = div_rn<FIVE>(x, s, rs);
  const float t3 = fadd(rne_magic(u), o);
  const bool m = (t3 >= qmin) && (t3 <= qmax);
  const float t5 = fsub(fminf(fmaxf(t3, qmin), qmax), o);
  const float gs5 = fmul(g, s);
  const float gt1 = m ? gs5 : 0.f;
  FqGrad r;
  r.gx = div_rn<FIVE>(gt1, s, rs);
  r.gs = fsub(fmul(g, t5), fmul(gt1, div_rn<FIVE>(u, s, rs)));
  r.go = fsub(gt1, gs5);
  return r;
}

}  // namespace mq
