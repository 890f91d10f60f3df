// Shared device/host helpers for libmqb200 (sm_100a only).
//
// Numerics contract (DESIGN.md "Arithmetic"): every parity-critical fp32 operation is written with an
// explicit round-to-nearest intrinsic (__fdiv_rn, __fmul_rn, __fadd_rn, rintf) so that ptxas can never
// contract it into an FMA or replace a division by a reciprocal multiply.  The reference computes
// clamp(round(x / scale) + offset, qmin, qmax) with torch fp32 ops (mobilellm/quantization/qmodule.py:286-290);
// torch.round is round-half-to-even == rintf == cvt.rni.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include <type_traits>

#define MQ_CLIPMIN 1e-5f   // qmodule.py:11
#define MQ_CLIPMAX 1e6f    // qmodule.py:12

namespace mq {

constexpr int kWarp = 32;

struct Ctx;  // defined in api.cu

// ---- exact fp32 primitives -------------------------------------------------------------------------------------
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

// q = clamp(rne(x / s) + o, qmin, qmax)       (qmodule.py:286-287)
__device__ __forceinline__ float quant_code(float x, float s, float o, float qmin, float qmax) {
  float q = fadd(rintf(fdiv(x, s)), o);
  return fminf(fmaxf(q, qmin), qmax);
}
// x^ = (q - o) * s                            (qmodule.py:290)
__device__ __forceinline__ float dequant(float q, float s, float o) { return fmul(fsub(q, o), s); }

__device__ __forceinline__ float fake_quant(float x, float s, float o, float qmin, float qmax) {
  return dequant(quant_code(x, s, o, qmin, qmax), s, o);
}

// scale/offset from (min,max)                  (qmodule.py:40-61)
__device__ __forceinline__ void scale_offset_from_minmax(float mn, float mx, int bits, bool symmetric, float& s,
                                                         float& o, float& qmin, float& qmax) {
  float alpha, beta;
  if (symmetric) {
    alpha = fmaxf(fabsf(mn), fabsf(mx));
    beta = 0.f;
    qmin = -(float)(1 << (bits - 1));
    qmax = (float)((1 << (bits - 1)) - 1);
  } else {
    alpha = fsub(mx, mn);
    beta = mn;
    qmin = 0.f;
    qmax = (float)((1 << bits) - 1);
  }
  s = fdiv(alpha, qmax);
  s = fminf(fmaxf(s, MQ_CLIPMIN), MQ_CLIPMAX);
  // offset = -(beta / scale).round(); symmetric: -(0/scale).round() = -0 -> 0
  o = symmetric ? 0.f : -rintf(fdiv(beta, s));
}

// ---- branch-free exact requantisation (integer engine kernels) ----------------------------------------------------
// RN(a / b) without the XU pipe or a slow-path branch, given rb = RN(1/b) (host: 1.0f/b in IEEE fp32, device:
// __frcp_rn).  q0 = a*rb is within 1 ulp of a/b; one Newton step on the exact FMA residual gives the correctly rounded
// quotient (Markstein) -- except when b's significand is all ones, where RN(1/b) is not accurate enough and a second
// step is needed (`five`, decided once per scale on the host / per kernel).  Valid for normal-range operands
// (activation and score magnitudes); tests/test_quant_kernels_gpu.py::test_div_rn_exact checks it against __fdiv_rn.
template <bool FIVE>
__device__ __forceinline__ float div_rn(float a, float b, float rb) {
  float q = __fmul_rn(a, rb);
  float r = __fmaf_rn(-q, b, a);
  q = __fmaf_rn(r, rb, q);
  if (FIVE) {
    r = __fmaf_rn(-q, b, a);
    q = __fmaf_rn(r, rb, q);
  }
  return q;
}
__host__ __device__ __forceinline__ bool mantissa_all_ones(float b) {
#ifdef __CUDA_ARCH__
  return (__float_as_uint(b) & 0x7fffffu) == 0x7fffffu;
#else
  union { float f; uint32_t u; } v; v.f = b; return (v.u & 0x7fffffu) == 0x7fffffu;
#endif
}
constexpr float kRoundMagic = 12582912.f;        // 1.5 * 2^23: x + magic rounds x to an integer (RNE) for |x| < 2^22
constexpr int kRoundMagicBits = 0x4B400000;

// Static quantizer with an INTEGRAL offset, prepared once: code = clamp(rne(x/s)+o, 0, qmax) computed as
// rne(clamp(x/s, -o, qmax-o)) + o (identical because rne is monotone and the bounds are integers).
struct QParam {
  float s, rs, lo, hi;
  int ioff;                                       // int(o) - kRoundMagicBits
  bool five;                                      // s has an all-ones significand: callers take the FIVE=true path
};
__device__ __forceinline__ QParam make_qparam(float s, float o, float qmax) {
  QParam p;
  p.s = s; p.rs = __frcp_rn(s); p.lo = -o; p.hi = __fsub_rn(qmax, o); p.ioff = (int)o - kRoundMagicBits;
  p.five = mantissa_all_ones(s);
  return p;
}
// rounded, clamped quotient as (magic + n): __float_as_int(m) + ioff is the code, __fsub_rn(m, magic) is n = code - o
template <bool FIVE>
__device__ __forceinline__ float quant_magic(float x, const QParam& p) {
  float q = div_rn<FIVE>(x, p.s, p.rs);
  q = fminf(fmaxf(q, p.lo), p.hi);
  return __fadd_rn(q, kRoundMagic);
}
template <bool FIVE>
__device__ __forceinline__ int quant_int(float x, const QParam& p) { return __float_as_int(quant_magic<FIVE>(x, p)) + p.ioff; }
// run body(std::true_type / std::false_type) on the exact-division variant the scales need (uniform branch)
template <typename F>
__device__ __forceinline__ void dispatch_five(bool five, F&& body) {
  if (five) body(std::true_type{}); else body(std::false_type{});
}

// ---- ordered-int encoding so float min/max can use integer atomics ---------------------------------------------
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// ---- warp / block reductions -----------------------------------------------------------------------------------
template <typename T, typename Op>
__device__ __forceinline__ T warp_reduce(T v, Op op) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
struct OpMin { template <typename T> __device__ T operator()(T a, T b) const { return a < b ? a : b; } };
struct OpMax { template <typename T> __device__ T operator()(T a, T b) const { return a > b ? a : b; } };
struct OpSum { template <typename T> __device__ T operator()(T a, T b) const { return a + b; } };
struct OpFMin { __device__ float operator()(float a, float b) const { return fminf(a, b); } };
struct OpFMax { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };

// Block reduce over blockDim.x threads (multiple of 32, <= 1024). Result valid in every thread.
template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T* smem /* >= 32 entries */) {
  v = warp_reduce(v, op);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  T r = smem[0];
  for (int i = 1; i < nw; ++i) r = op(r, smem[i]);
  return r;
}

// ---- programmatic dependent launch (decode step: ~200 dependent kernels of 3-15 us per token) -------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the stream is
// still running: everything before pdl_wait() (barrier / TMEM set-up, loads of CONSTANT data such as weights and lookup tables)
// overlaps the predecessor; pdl_wait() returns once the predecessor grid has completed and its writes are visible.
// pdl_trigger() lets the NEXT kernel in the stream be scheduled (it is issued at the top of every decode kernel: a dependent
// can only start once every CTA of its predecessor has started, so a running kernel never waits for a slot held by a
// kernel that waits for it).  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// hints for constant data touched before pdl_wait()
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- vector ld/st ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

}  // namespace mq
