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

// The same gradients in the cancellation-free form  dy/dx = m,  dy/ds = t5 - m*u,  dy/do = (m - 1)*s   (u = x/s,
// t5 = clamp(rne(u)+o) - o, m = in-range).  The op-by-op form above evaluates dy/ds as g*t5 - (g*s*m)*(u/s): two products of
// magnitude |g|*|u| whose difference is at most |g|/2, each carrying an fp32 rounding error of |g|*|u|*2^-24 (about 1 % of
// the result for 16-bit codes); here t5 - u is exact (both are within 1/2 of each other), there is one exact division instead
// of three, and the two forms differ by that rounding noise only (and gx by the <= 1 ulp of (g*s)/s versus g).
template <bool FIVE>
__device__ __forceinline__ FqGrad fq_bwd_elem_c(float x, float g, float s, float rs, float o, float qmin, float qmax) {
  const float u = div_rn<FIVE>(x, s, rs);
  const float t3 = fadd(rne_magic(u), o);
  const bool m = (t3 >= qmin) && (t3 <= qmax);
  const float t5 = fsub(fminf(fmaxf(t3, qmin), qmax), o);
  FqGrad r;
  r.gx = m ? g : 0.f;
  r.gs = fmul(g, fsub(t5, m ? u : 0.f));
  r.go = m ? 0.f : -fmul(g, s);
  return r;
}

// ---- a per-tensor static quantizer as the fused calibration kernels (calib_attn.cu, calib_act.cu) carry it ---------------
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
// FIVE is chosen once per launch (true when any of the launch's scales needs the second Newton step; the longer variant is
// exact for every divisor).  fq_grad is the cancellation-free form (fq_bwd_elem_c).
template <bool FIVE>
__device__ __forceinline__ float fq_apply(float x, const FqP& q) {
  if (!q.on) return x;
  return dequant(quant_code_v<FIVE>(x, q.s, q.rs, q.o, q.qmin, q.qmax), q.s, q.o);
}
template <bool FIVE>
__device__ __forceinline__ FqGrad fq_grad(float x, float g, const FqP& q) {
  if (!q.on) { FqGrad r; r.gx = g; r.gs = 0.f; r.go = 0.f; return r; }
  return fq_bwd_elem_c<FIVE>(x, g, q.s, q.rs, q.o, q.qmin, q.qmax);
}


// Deterministic grid-wide fold of NACC per-thread float accumulators (quantizer scale / offset gradients): block sums go to
// double partial[NACC * nblk]; the block whose ticket shows it arrived last adds them -- every thread of that block takes the
// partials bid = tid, tid + blockDim, ... and the threads meet in a fixed shuffle / shared-memory tree, so the order of the
// additions depends only on the launch geometry -- and writes gout[NACC]; the ticket resets itself.
// `red` = 32 floats of shared memory, `s_last` one shared bool.  blockDim.x: a multiple of 32, <= 256.
template <int NACC>
__device__ __forceinline__ void grid_fold_to(const float (&acc)[NACC], double* partial, unsigned* ticket, float* const (&outs)[NACC],
                                             float* red, bool* s_last, unsigned bid = blockIdx.x, unsigned nblk = gridDim.x) {
  __shared__ double s_fold[8][NACC];
  float b[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) b[i] = block_reduce(acc[i], OpSum(), red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) partial[NACC * bid + i] = b[i];
    __threadfence();
    *s_last = atomicAdd(ticket, 1u) == nblk - 1;
  }
  __syncthreads();
  if (!*s_last) return;                                   // block-uniform
  __threadfence();
  const volatile double* vp = partial;
  double t[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) t[k] = 0.;
  for (unsigned i = threadIdx.x; i < nblk; i += blockDim.x)
#pragma unroll
    for (int k = 0; k < NACC; ++k) t[k] += vp[NACC * i + k];
#pragma unroll
  for (int k = 0; k < NACC; ++k) t[k] = warp_reduce(t[k], OpSum());
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NACC; ++k) s_fold[wid][k] = t[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
      double s = s_fold[0][k];
      for (int w = 1; w < nw; ++w) s += s_fold[w][k];
      if (outs[k]) *outs[k] = (float)s;
    }
    *ticket = 0u;
  }
}
template <int NACC>
__device__ __forceinline__ void grid_fold(const float (&acc)[NACC], double* partial, unsigned* ticket, float* gout, float* red,
                                          bool* s_last, unsigned bid = blockIdx.x, unsigned nblk = gridDim.x) {
  float* outs[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) outs[k] = gout + k;
  grid_fold_to<NACC>(acc, partial, ticket, outs, red, s_last, bid, nblk);
}

}  // namespace mq
