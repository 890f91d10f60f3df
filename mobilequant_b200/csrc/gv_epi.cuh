// Requantisation epilogue of the skinny (decode-step) GEMM, shared by decode.cu (stand-alone / last-CTA-fused epilogue)
// and engine.cu (RESID epilogue folded into the following row norm).  Same arithmetic as qgemm_kernel's epilogue (qgemm.cu).
#pragma once
#include "common.cuh"

namespace mq {

enum { GV_NONE = -1, GV_QUANT = 0, GV_ACTMUL = 1, GV_RESID = 2 };

struct GvEpiArgs {
  int B, N;
  int32_t* acc; int ldacc;
  const int32_t* rowsum; const float* sxw; const int32_t* ow; const int32_t* c0; const float* bias;
  const float* so; const float* oo; int qgroup; float qmax;
  uint8_t* out; int64_t ldo; int32_t* rowsum_out;
  const float* lut; float s2, o2, qmax2;
  float* resid;
  int32_t* zero_out;       // [B] or null: cleared here so that the NEXT accumulation into it starts from zero
};

__device__ __forceinline__ float gv_y(const GvEpiArgs& a, int acc, int rs, int n) {
  const int I = acc - __ldg(a.ow + n) * rs + __ldg(a.c0 + n);
  float y = __fmul_rn(__int2float_rn(I), __ldg(a.sxw + n));
  if (a.bias) y = __fadd_rn(y, __ldg(a.bias + n));
  return y;
}

// row m, output columns j0..j0+3 (j0 % 4 == 0, j0 < number of output columns); returns the sum of the emitted codes.
// The accumulator words are read from L2 (they were produced by red.add of other CTAs) and handed back zeroed.
template <int MODE>
__device__ __forceinline__ int gv_epi_quad(const GvEpiArgs& a, int m, int j0) {
  const int rs = __ldg(a.rowsum + m);
  int32_t* accm = a.acc + int64_t(m) * a.ldacc;
  const int gmax = (a.N - 1) / a.qgroup;
  int csum = 0;
  if (MODE == GV_ACTMUL) {
    // output column j <-> w1 accumulator column (j / 128) * 256 + j % 128, w3 column 128 further
    const int n1 = (j0 >> 7) * 256 + (j0 & 127), n3 = n1 + 128;
    const int4 a1 = __ldcg(reinterpret_cast<const int4*>(accm + n1)), a3 = __ldcg(reinterpret_cast<const int4*>(accm + n3));
    *reinterpret_cast<int4*>(accm + n1) = make_int4(0, 0, 0, 0);
    *reinterpret_cast<int4*>(accm + n3) = make_int4(0, 0, 0, 0);
    const int g1 = min(n1 / a.qgroup, gmax), g3 = min(n3 / a.qgroup, gmax);
    const QParam q1 = make_qparam(__ldg(a.so + g1), __ldg(a.oo + g1), a.qmax), q3 = make_qparam(__ldg(a.so + g3), __ldg(a.oo + g3), a.qmax);
    const QParam q2 = make_qparam(a.s2, a.o2, a.qmax2);
    const int v1[4] = {a1.x, a1.y, a1.z, a1.w}, v3[4] = {a3.x, a3.y, a3.z, a3.w};
    uint32_t w = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float y1 = gv_y(a, v1[e], rs, n1 + e), y3 = gv_y(a, v3[e], rs, n3 + e);
      const float act = __ldg(a.lut + quant_int<true>(y1, q1));
      const float u = __fmul_rn(__fsub_rn(quant_magic<true>(y3, q3), kRoundMagic), q3.s);
      w |= (uint32_t)quant_int<true>(__fmul_rn(act, u), q2) << (8 * e);
    }
    csum = (int)__dp4a(w, 0x01010101u, 0u);
    *reinterpret_cast<uint32_t*>(a.out + int64_t(m) * a.ldo + j0) = w;
  } else {
    const int4 av = __ldcg(reinterpret_cast<const int4*>(accm + j0));
    *reinterpret_cast<int4*>(accm + j0) = make_int4(0, 0, 0, 0);
    const int g = min(j0 / a.qgroup, gmax);
    const QParam q = make_qparam(__ldg(a.so + g), __ldg(a.oo + g), a.qmax);
    const int v[4] = {av.x, av.y, av.z, av.w};
    if (MODE == GV_QUANT) {
      uint32_t w = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) w |= (uint32_t)quant_int<true>(gv_y(a, v[e], rs, j0 + e), q) << (8 * e);
      csum = (int)__dp4a(w, 0x01010101u, 0u);
      *reinterpret_cast<uint32_t*>(a.out + int64_t(m) * a.ldo + j0) = w;
    } else {
      float4* dst = reinterpret_cast<float4*>(a.resid + int64_t(m) * a.ldo + j0);
      float4 h = *dst;
      float d[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) d[e] = __fmul_rn(__fsub_rn(quant_magic<true>(gv_y(a, v[e], rs, j0 + e), q), kRoundMagic), q.s);
      h.x = __fadd_rn(h.x, d[0]); h.y = __fadd_rn(h.y, d[1]); h.z = __fadd_rn(h.z, d[2]); h.w = __fadd_rn(h.w, d[3]);
      *dst = h;
    }
  }
  return csum;
}


}  // namespace mq
