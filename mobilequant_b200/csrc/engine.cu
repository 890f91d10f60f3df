// K4/K5/K6 of the statically-quantised integer forward: norm, RoPE + requant, exact quantised causal attention.
// Every kernel consumes and produces integer codes; the arithmetic is restated 1:1 in oracle/int_ref.py.
#include "common.cuh"
#include "ctx.h"
#include "gv_epi.cuh"
#include "attn.cuh"
#include "tc_common.cuh"
#include <string>
#include <algorithm>
#include <cstdlib>
#include <cstdio>

namespace mq {

// =====================================================================================================================
// K4: QRMSNorm.forward (qm:515-531) with l2norm_as_rmsnorm (hm:187-195):
//   r = clamp(rne(x/s_in)+o_in, 0, qmax_in) - o_in ; x^ = r*s_in ; nrm = sqrt(float(sum r^2)) * s_in  (sum exact, u64)
//   t = w_fq * (alpha * (x^ / max(nrm, 1e-12))) (+ bias) ; code = clamp(rne(t/s_out)+o_out, 0, qmax_out)
// QLayerNorm.forward (qm:625-642): mean/var from the exact integer sums (double), y = ((x^-mean)*rstd)*w + b.
// One warp per row.  NV > 0: the row's de-offset integers r stay in registers (H == 128*NV) so x is read once;
// NV == 0: generic H, second pass re-reads the row through L1/L2.  All requantisation is the branch-free exact
// division of common.cuh (the previous version was bound by the XU pipe: MUFU.RCP + FRND + F2I per element).
// =====================================================================================================================
struct NormArgs {
  const float* x; int64_t rows; int H;
  float s_in, o_in, qmax_in;
  const float* w_fq; const float* bias;
  float alpha, eps, s_out, o_out, qmax_out;
  uint8_t* codes; int32_t* rowsum;
};

template <bool kLayerNorm, int NV>
__global__ void __launch_bounds__(128) qnorm_kernel(const NormArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const int H = a.H;
  const float* xr = a.x + row * H;
  const QParam qi = make_qparam(a.s_in, a.o_in, a.qmax_in);
  const QParam qo = make_qparam(a.s_out, a.o_out, a.qmax_out);
  constexpr int NR = NV > 0 ? NV : 1;
  float rr[NR][4];                               // r = code - o_in as exact fp32 integers
  // ---- pass 1: integer statistics
  unsigned long long s2 = 0; long long s1 = 0;
  auto stats = [&](auto five_tag) {
    constexpr bool FIVE = decltype(five_tag)::value;
    auto one = [&](float4 v, float (&r)[4]) {
      r[0] = __fsub_rn(quant_magic<FIVE>(v.x, qi), kRoundMagic); r[1] = __fsub_rn(quant_magic<FIVE>(v.y, qi), kRoundMagic);
      r[2] = __fsub_rn(quant_magic<FIVE>(v.z, qi), kRoundMagic); r[3] = __fsub_rn(quant_magic<FIVE>(v.w, qi), kRoundMagic);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = __float2int_rn(r[j]);
        s2 += (unsigned long long)((long long)i * i);
        if (kLayerNorm) s1 += i;
      }
    };
    if (NV > 0) {
#pragma unroll
      for (int it = 0; it < NR; ++it) one(ldg4_stream(xr + it * 128 + lane * 4), rr[it]);
    } else {
      float tmp[4];
      for (int k = lane * 4; k < H; k += 128) one(ldg4(xr + k), tmp);
    }
  };
  dispatch_five(qi.five, stats);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    if (kLayerNorm) s1 += __shfl_xor_sync(0xffffffffu, s1, d);
  }
  float denom = 1.f, rdenom = 1.f, mean = 0.f, rstd = 1.f;
  bool five = qo.five | qi.five;
  if (kLayerNorm) {
    const double m = (double)s1 / (double)H;
    const double var = (double)s2 / (double)H - m * m;
    mean = (float)(m * (double)a.s_in);
    rstd = (float)(1.0 / sqrt(var * (double)a.s_in * (double)a.s_in + (double)a.eps));
  } else {
    denom = fmaxf(fmul(__fsqrt_rn(__ull2float_rn(s2)), a.s_in), 1e-12f);
    rdenom = __frcp_rn(denom);
    five |= mantissa_all_ones(denom);           // warp-uniform: one row per warp
  }
  // ---- pass 2: normalise, requantise, emit codes + row sum
  int csum = 0;
  auto emit = [&](auto five_tag) {
    constexpr bool FIVE = decltype(five_tag)::value;
    auto one = [&](const float (&r)[4], int k) {
      const float4 w = ldg4(a.w_fq + k);
      const float wv[4] = {w.x, w.y, w.z, w.w};
      float bv[4] = {0.f, 0.f, 0.f, 0.f};
      if (a.bias) { const float4 b = ldg4(a.bias + k); bv[0] = b.x; bv[1] = b.y; bv[2] = b.z; bv[3] = b.w; }
      uint32_t packed = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = fmul(r[j], a.s_in);
        float t;
        if (kLayerNorm) t = fmul(fmul(fsub(xh, mean), rstd), wv[j]);
        else t = fmul(wv[j], fmul(a.alpha, div_rn<FIVE>(xh, denom, rdenom)));
        if (a.bias) t = fadd(t, bv[j]);
        packed |= (uint32_t)quant_int<FIVE>(t, qo) << (8 * j);
      }
      csum = (int)__dp4a(packed, 0x01010101u, (unsigned)csum);
      *reinterpret_cast<uint32_t*>(a.codes + row * H + k) = packed;
    };
    if (NV > 0) {
#pragma unroll
      for (int it = 0; it < NR; ++it) one(rr[it], it * 128 + lane * 4);
    } else {
      for (int k = lane * 4; k < H; k += 128) {
        const float4 v = ldg4(xr + k);
        float r[4];
        r[0] = __fsub_rn(quant_magic<FIVE>(v.x, qi), kRoundMagic); r[1] = __fsub_rn(quant_magic<FIVE>(v.y, qi), kRoundMagic);
        r[2] = __fsub_rn(quant_magic<FIVE>(v.z, qi), kRoundMagic); r[3] = __fsub_rn(quant_magic<FIVE>(v.w, qi), kRoundMagic);
        one(r, k);
      }
    }
  };
  dispatch_five(five, emit);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, d);
  if (lane == 0 && a.rowsum) a.rowsum[row] = csum;
}

// Pipelined variant for the prefill (thousands of rows, H == 128*NV): persistent CTAs, one warp per row, the row's 4*H bytes
// arrive by ONE bulk async copy (cp.async.bulk + mbarrier) into the warp's shared-memory slot; pass 1 pulls the row into
// registers, after which the slot is refilled with the warp's next row underneath pass 2 (the longer pass), so the exact
// requantisation chain (~25 instructions per element: this kernel is ALU-bound, not HBM-bound) no longer
// waits for its loads.  Same arithmetic as qnorm_kernel; w_fq / bias are staged once per CTA.
template <bool kLayerNorm, int NV>
__global__ void __launch_bounds__(128) qnorm_pipe_kernel(const NormArgs a) {
  constexpr int H = NV * 128;
  constexpr int kRowBytes = H * 4;
  extern __shared__ __align__(128) uint8_t qn_smem[];
  float* s_w = reinterpret_cast<float*>(qn_smem);                           // [H]
  float* s_b = s_w + H;                                                     // [H] (zeros without a bias)
  float* s_x = s_b + H;                                                     // [4 warps][H]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_x + 4 * H);               // [4 warps]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x * 4; k < H; k += 128 * 4) {
    *reinterpret_cast<float4*>(s_w + k) = ldg4(a.w_fq + k);
    *reinterpret_cast<float4*>(s_b + k) = a.bias ? ldg4(a.bias + k) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float* xb = s_x + warp * H;
  uint64_t* bar = s_bar + warp;
  if (lane == 0) {
    tc::mbar_init(bar, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  const QParam qi = make_qparam(a.s_in, a.o_in, a.qmax_in);
  const QParam qo = make_qparam(a.s_out, a.o_out, a.qmax_out);
  const int64_t nwarps = int64_t(gridDim.x) * 4;
  int64_t row = int64_t(blockIdx.x) * 4 + warp;
  auto issue = [&](int64_t r) {                                             // lane 0 only
    tc::mbar_expect_tx(bar, kRowBytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(xb)), "l"(a.x + r * H), "r"(kRowBytes), "r"(tc::smem_u32(bar)) : "memory");
  };
  if (row < a.rows && lane == 0) issue(row);
  for (int it = 0; row < a.rows; ++it, row += nwarps) {
    tc::mbar_wait(bar, it & 1);
    const float* xr = xb;
    float rr[NV][4];
    unsigned long long s2 = 0; long long s1 = 0;
    auto stats = [&](auto five_tag) {
      constexpr bool FIVE = decltype(five_tag)::value;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(xr + j * 128 + lane * 4);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float m = quant_magic<FIVE>(vv[e], qi);
          rr[j][e] = __fsub_rn(m, kRoundMagic);
          const int i = __float_as_int(m) - kRoundMagicBits;        // == int(rr): no F2I
          s2 += (unsigned long long)((long long)i * i);
          if (kLayerNorm) s1 += i;
        }
      }
    };
    dispatch_five(qi.five, stats);
    __syncwarp();                                     // the row is in registers: refill the slot underneath pass 2
    if (row + nwarps < a.rows && lane == 0) issue(row + nwarps);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      s2 += __shfl_xor_sync(0xffffffffu, s2, d);
      if (kLayerNorm) s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    }
    float denom = 1.f, rdenom = 1.f, mean = 0.f, rstd = 1.f;
    bool five = qo.five | qi.five;
    if (kLayerNorm) {
      const double m = (double)s1 / (double)H;
      const double var = (double)s2 / (double)H - m * m;
      mean = (float)(m * (double)a.s_in);
      rstd = (float)(1.0 / sqrt(var * (double)a.s_in * (double)a.s_in + (double)a.eps));
    } else {
      denom = fmaxf(fmul(__fsqrt_rn(__ull2float_rn(s2)), a.s_in), 1e-12f);
      rdenom = __frcp_rn(denom);
      five |= mantissa_all_ones(denom);
    }
    int csum = 0;
    const bool has_bias = a.bias != nullptr;
    auto emit = [&](auto five_tag) {
      constexpr bool FIVE = decltype(five_tag)::value;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int k = j * 128 + lane * 4;
        const float4 w = *reinterpret_cast<const float4*>(s_w + k);
        const float4 bb = *reinterpret_cast<const float4*>(s_b + k);
        const float wv[4] = {w.x, w.y, w.z, w.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
        uint32_t packed = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float xh = fmul(rr[j][e], a.s_in);
          float t;
          if (kLayerNorm) t = fmul(fmul(fsub(xh, mean), rstd), wv[e]);
          else t = fmul(wv[e], fmul(a.alpha, div_rn<FIVE>(xh, denom, rdenom)));
          if (has_bias) t = fadd(t, bv[e]);
          packed |= (uint32_t)quant_int<FIVE>(t, qo) << (8 * e);
        }
        csum = (int)__dp4a(packed, 0x01010101u, (unsigned)csum);
        *reinterpret_cast<uint32_t*>(a.codes + row * H + k) = packed;
      }
    };
    dispatch_five(five, emit);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, d);
    if (lane == 0 && a.rowsum) a.rowsum[row] = csum;
  }
}

// Split-row variant for the prefill (H == 256*NV): TWO warps per row, each keeping half of the row's de-offset integers in
// registers (half the registers of qnorm_kernel -> twice the resident warps, half the dependent chain per warp); the exact
// integer statistics and the emitted-code sum meet through shared memory behind a 64-thread named barrier.  Same arithmetic.
template <bool kLayerNorm, int NV>
__global__ void __launch_bounds__(128) qnorm_split_kernel(const NormArgs a) {
  __shared__ unsigned long long s_s2[2][2];
  __shared__ long long s_s1[2][2];
  __shared__ int s_cs[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, slot = warp >> 1, hw = warp & 1;
  const int64_t row = int64_t(blockIdx.x) * 2 + slot;
  if (row >= a.rows) return;                       // both warps of a row leave together
  constexpr int H = NV * 256;
  const float* xr = a.x + row * H + hw * (H / 2);
  const QParam qi = make_qparam(a.s_in, a.o_in, a.qmax_in);
  const QParam qo = make_qparam(a.s_out, a.o_out, a.qmax_out);
  float rr[NV][4];
  unsigned long long s2 = 0; long long s1 = 0;
  auto stats = [&](auto five_tag) {
    constexpr bool FIVE = decltype(five_tag)::value;
#pragma unroll
    for (int it = 0; it < NV; ++it) {
      const float4 v = ldg4_stream(xr + it * 128 + lane * 4);
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float m = quant_magic<FIVE>(vv[e], qi);
        rr[it][e] = __fsub_rn(m, kRoundMagic);
        const int i = __float_as_int(m) - kRoundMagicBits;          // == int(rr)
        s2 += (unsigned long long)((long long)i * i);
        if (kLayerNorm) s1 += i;
      }
    }
  };
  dispatch_five(qi.five, stats);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    if (kLayerNorm) s1 += __shfl_xor_sync(0xffffffffu, s1, d);
  }
  if (lane == 0) { s_s2[slot][hw] = s2; if (kLayerNorm) s_s1[slot][hw] = s1; }
  asm volatile("bar.sync %0, 64;" ::"r"(1 + slot) : "memory");
  s2 = s_s2[slot][0] + s_s2[slot][1];
  if (kLayerNorm) s1 = s_s1[slot][0] + s_s1[slot][1];
  float denom = 1.f, rdenom = 1.f, mean = 0.f, rstd = 1.f;
  bool five = qo.five | qi.five;
  if (kLayerNorm) {
    const double m = (double)s1 / (double)H;
    const double var = (double)s2 / (double)H - m * m;
    mean = (float)(m * (double)a.s_in);
    rstd = (float)(1.0 / sqrt(var * (double)a.s_in * (double)a.s_in + (double)a.eps));
  } else {
    denom = fmaxf(fmul(__fsqrt_rn(__ull2float_rn(s2)), a.s_in), 1e-12f);
    rdenom = __frcp_rn(denom);
    five |= mantissa_all_ones(denom);
  }
  int csum = 0;
  const bool has_bias = a.bias != nullptr;
  auto emit = [&](auto five_tag) {
    constexpr bool FIVE = decltype(five_tag)::value;
#pragma unroll
    for (int it = 0; it < NV; ++it) {
      const int k = hw * (H / 2) + it * 128 + lane * 4;
      const float4 w = ldg4(a.w_fq + k);
      const float wv[4] = {w.x, w.y, w.z, w.w};
      float bv[4] = {0.f, 0.f, 0.f, 0.f};
      if (has_bias) { const float4 b = ldg4(a.bias + k); bv[0] = b.x; bv[1] = b.y; bv[2] = b.z; bv[3] = b.w; }
      uint32_t packed = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xh = fmul(rr[it][e], a.s_in);
        float t;
        if (kLayerNorm) t = fmul(fmul(fsub(xh, mean), rstd), wv[e]);
        else t = fmul(wv[e], fmul(a.alpha, div_rn<FIVE>(xh, denom, rdenom)));
        if (has_bias) t = fadd(t, bv[e]);
        packed |= (uint32_t)quant_int<FIVE>(t, qo) << (8 * e);
      }
      csum = (int)__dp4a(packed, 0x01010101u, (unsigned)csum);
      *reinterpret_cast<uint32_t*>(a.codes + row * H + k) = packed;
    }
  };
  dispatch_five(five, emit);
  if (a.rowsum) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, d);
    if (hw == 1 && lane == 0) s_cs[slot] = csum;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + slot) : "memory");
    if (hw == 0 && lane == 0) a.rowsum[row] = csum + s_cs[slot];
  }
}

// Few-row variant (decode step: rows == batch): one warp per row leaves a ~3 k-instruction dependent chain on a single
// warp (17 us per launch for 8 rows); here one 256-thread CTA owns a row, statistics meet through shared memory.  Same
// arithmetic: the integer sums are exact, so the reduction order does not matter.
// FUSE: the residual-add epilogue of the preceding skinny GEMM (o_proj / w2 of the decode step) is applied to the row first
// -- x[m, :] += dequant(Q_out(y)), accumulator handed back zeroed, exactly gv_epi_quad<GV_RESID> -- by the thread that
// then normalises those columns: two launches per layer fewer in the decode chain.
template <bool kLayerNorm, bool FUSE>
__global__ void __launch_bounds__(256) qnorm_row_kernel(const NormArgs a, const GvEpiArgs e) {
  pdl_trigger();                                 // decode chain: let the next kernel set itself up (common.cuh)
  for (int k = threadIdx.x * 4; k < a.H; k += 1024) {       // constants of the second pass: into L1 while the predecessor runs
    prefetch_l1(a.w_fq + k);
    if (a.bias) prefetch_l1(a.bias + k);
  }
  pdl_wait();                                    // x / the GEMV accumulator come from the predecessor
  __shared__ unsigned long long s_s2[8];
  __shared__ long long s_s1[8];
  __shared__ int s_cs[8];
  const int64_t row = blockIdx.x;
  const int H = a.H, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* xr = a.x + row * H;
  const QParam qi = make_qparam(a.s_in, a.o_in, a.qmax_in);
  const QParam qo = make_qparam(a.s_out, a.o_out, a.qmax_out);
  constexpr int NR = 8;                           // float4 per thread: H <= 8192
  float rr[NR][4];
  unsigned long long s2 = 0; long long s1 = 0;
#pragma unroll
  for (int it = 0; it < NR; ++it) {
    const int k = (it * 256 + tid) * 4;
    if (k < H) {
      float4 v;
      if (FUSE) {
        gv_epi_quad<GV_RESID>(e, (int)row, k);            // updates x[row, k..k+3] in place (e.resid == a.x)
        v = *reinterpret_cast<const float4*>(xr + k);      // this thread's own stores
      } else {
        v = ldg4_stream(xr + k);
      }
      rr[it][0] = __fsub_rn(quant_magic<true>(v.x, qi), kRoundMagic); rr[it][1] = __fsub_rn(quant_magic<true>(v.y, qi), kRoundMagic);
      rr[it][2] = __fsub_rn(quant_magic<true>(v.z, qi), kRoundMagic); rr[it][3] = __fsub_rn(quant_magic<true>(v.w, qi), kRoundMagic);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = __float2int_rn(rr[it][j]);
        s2 += (unsigned long long)((long long)i * i);
        if (kLayerNorm) s1 += i;
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    if (kLayerNorm) s1 += __shfl_xor_sync(0xffffffffu, s1, d);
  }
  if (lane == 0) { s_s2[warp] = s2; s_s1[warp] = s1; }
  __syncthreads();
  s2 = 0; s1 = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s2 += s_s2[i]; s1 += s_s1[i]; }
  float denom = 1.f, rdenom = 1.f, mean = 0.f, rstd = 1.f;
  if (kLayerNorm) {
    const double m = (double)s1 / (double)H;
    const double var = (double)s2 / (double)H - m * m;
    mean = (float)(m * (double)a.s_in);
    rstd = (float)(1.0 / sqrt(var * (double)a.s_in * (double)a.s_in + (double)a.eps));
  } else {
    denom = fmaxf(fmul(__fsqrt_rn(__ull2float_rn(s2)), a.s_in), 1e-12f);
    rdenom = __frcp_rn(denom);
  }
  int csum = 0;
#pragma unroll
  for (int it = 0; it < NR; ++it) {
    const int k = (it * 256 + tid) * 4;
    if (k < H) {
      const float4 w = ldg4(a.w_fq + k);
      const float wv[4] = {w.x, w.y, w.z, w.w};
      float bv[4] = {0.f, 0.f, 0.f, 0.f};
      if (a.bias) { const float4 b = ldg4(a.bias + k); bv[0] = b.x; bv[1] = b.y; bv[2] = b.z; bv[3] = b.w; }
      uint32_t packed = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = fmul(rr[it][j], a.s_in);
        float t;
        if (kLayerNorm) t = fmul(fmul(fsub(xh, mean), rstd), wv[j]);
        else t = fmul(wv[j], fmul(a.alpha, div_rn<true>(xh, denom, rdenom)));
        if (a.bias) t = fadd(t, bv[j]);
        packed |= (uint32_t)quant_int<true>(t, qo) << (8 * j);
      }
      csum = (int)__dp4a(packed, 0x01010101u, (unsigned)csum);
      *reinterpret_cast<uint32_t*>(a.codes + row * H + k) = packed;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, d);
  if (lane == 0) s_cs[warp] = csum;
  __syncthreads();
  if (tid == 0 && a.rowsum) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s_cs[i];
    a.rowsum[row] = t;
  }
}

// =====================================================================================================================
// K5: de-quantise q/k/v projection codes, RoPE (hm:338-367, partial rotary hm:489-501), re-quantise with the qk_bmm /
// pv_bmm input quantizers (qm:455-459) and write the attention layouts:
//   q  [B, nh, T, hd] u8      k [B, nkv, T, hd] u8      vT [B, nkv, hd, T] u8   (+ per-token code sums of q and k rows)
// One CTA per 16 tokens.  A thread owns 4 consecutive head dims (one 32-bit word of codes) of one (token, head) and
// reads the partner word for rotate_half; the lanes of one (token, head) are adjacent so the code sum is a shuffle
// reduction.  V goes through a shared-memory transpose so that vT rows are written 16 bytes at a time.
// =====================================================================================================================
constexpr int kRopeTok = 16;

struct RopeArgs {
  const uint8_t* qkv;     // [M, ldq] codes of the fused q|k|v projection
  int ldq;
  int B, T, nh, nkv, hd, rot;
  float sq_in, oq_in, sk_in, ok_in, sv_in, ov_in;       // projection output quantizers (dequant)
  float sq, oq, sk, ok, sv, ov;                         // qk_bmm.input, qk_bmm.input2, pv_bmm.input2 (requant), 8 bit
  const float* cos;       // [T, rot]
  const float* sin;       // [T, rot]
  uint8_t *q, *k, *vt;
  int32_t *rsq, *rsk;     // [B, nh, T], [B, nkv, T] code sums over hd
};

__device__ __forceinline__ void unpack4(uint32_t w, float s, float o, float (&x)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) x[j] = dequant(__uint2float_rn((w >> (8 * j)) & 255u), s, o);
}

template <int HD, bool FIVE>
__device__ __forceinline__ void qrope_body(const RopeArgs& a, uint8_t* vs) {
  constexpr int WORDS = HD / 4;                          // 32-bit words per head
  constexpr int WPT = WORDS > 32 ? WORDS / 32 : 1;       // words per thread (hd = 256 -> 2)
  constexpr int LPI = WORDS / WPT;                       // lanes per (token, head): 8, 16 or 32
  constexpr int HPW = 32 / LPI;                          // heads per warp iteration
  const int tok0 = blockIdx.x * kRopeTok;
  const int M = a.B * a.T;
  const int half = a.rot / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane % LPI, grp = lane / LPI;
  const QParam qq = make_qparam(a.sq, a.oq, 255.f), qk = make_qparam(a.sk, a.ok, 255.f), qv = make_qparam(a.sv, a.ov, 255.f);
  const int heads_qk = a.nh + a.nkv;
  const int vw = a.nkv * HD;
  const int vstride = vw + 4;
  // ---- each warp owns tokens warp, warp + 8 of the CTA's 16
  for (int tl = warp; tl < kRopeTok; tl += 8) {
    const int tok = tok0 + tl;
    if (tok >= M) {                                      // keep the V tile defined for the transpose below
      for (int c = lane * 4; c < vw; c += 128) *reinterpret_cast<uint32_t*>(vs + tl * vstride + c) = 0u;
      continue;
    }
    const int b = tok / a.T, t = tok % a.T;
    const uint8_t* row = a.qkv + int64_t(tok) * a.ldq;
    if (a.rot == HD) {
      // ---- full rotary width: pair-wise path.  A lane owns 8 dims of the lower half of a head AND their 8 partners in the upper
      // half (HD / 16 lanes per head), so every code is loaded, unpacked and de-quantised once and both rotated halves come out
      // of the same registers: out[d] = x[d]*cos[d] + x[d+h]*(-sin[d]),  out[d+h] = x[d+h]*cos[d+h] + x[d]*sin[d+h]
      // (hm:338-367 with rotate_half = cat(-x2, x1); operation for operation what the generic path below computes).
      constexpr int LPH = HD / 16, HPW2 = 32 / LPH;
      const int sub2 = lane % LPH, grp2 = lane / LPH;
      const int dl = sub2 * 8, dh = dl + HD / 2;
      float cl[8], ch[8], nsl[8], sh[8];
      {
        const float* cr = a.cos + int64_t(t) * a.rot; const float* sr = a.sin + int64_t(t) * a.rot;
        const float4 c0 = ldg4(cr + dl), c1 = ldg4(cr + dl + 4), c2 = ldg4(cr + dh), c3 = ldg4(cr + dh + 4);
        const float4 s0 = ldg4(sr + dl), s1 = ldg4(sr + dl + 4), s2 = ldg4(sr + dh), s3 = ldg4(sr + dh + 4);
        cl[0] = c0.x; cl[1] = c0.y; cl[2] = c0.z; cl[3] = c0.w; cl[4] = c1.x; cl[5] = c1.y; cl[6] = c1.z; cl[7] = c1.w;
        ch[0] = c2.x; ch[1] = c2.y; ch[2] = c2.z; ch[3] = c2.w; ch[4] = c3.x; ch[5] = c3.y; ch[6] = c3.z; ch[7] = c3.w;
        nsl[0] = -s0.x; nsl[1] = -s0.y; nsl[2] = -s0.z; nsl[3] = -s0.w; nsl[4] = -s1.x; nsl[5] = -s1.y; nsl[6] = -s1.z; nsl[7] = -s1.w;
        sh[0] = s2.x; sh[1] = s2.y; sh[2] = s2.z; sh[3] = s2.w; sh[4] = s3.x; sh[5] = s3.y; sh[6] = s3.z; sh[7] = s3.w;
      }
      const int64_t strideh2 = int64_t(a.T) * HD;
      auto heads2 = [&](int nheads, const uint8_t* src0, float s_in, float o_in, const QParam& qo, uint8_t* dst0, int32_t* rs0) {
        constexpr int U = 1;                             // head groups in flight per lane (2 doubles the registers: 128, two CTAs per SM)
        for (int hh0 = 0; hh0 < nheads; hh0 += U * HPW2) {
          uint2 wl[U], wh[U];
          bool ok[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int hh = hh0 + u * HPW2 + grp2;
            ok[u] = hh < nheads;
            const uint8_t* src = src0 + (ok[u] ? hh : 0) * HD;
            wl[u] = __ldg(reinterpret_cast<const uint2*>(src + dl));
            wh[u] = __ldg(reinterpret_cast<const uint2*>(src + dh));
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int hh = hh0 + u * HPW2 + grp2;
            float x[8], y[8];
            { float t4[4]; unpack4(wl[u].x, s_in, o_in, t4); x[0] = t4[0]; x[1] = t4[1]; x[2] = t4[2]; x[3] = t4[3];
              unpack4(wl[u].y, s_in, o_in, t4); x[4] = t4[0]; x[5] = t4[1]; x[6] = t4[2]; x[7] = t4[3];
              unpack4(wh[u].x, s_in, o_in, t4); y[0] = t4[0]; y[1] = t4[1]; y[2] = t4[2]; y[3] = t4[3];
              unpack4(wh[u].y, s_in, o_in, t4); y[4] = t4[0]; y[5] = t4[1]; y[6] = t4[2]; y[7] = t4[3]; }
            uint32_t pl[2] = {0u, 0u}, ph[2] = {0u, 0u};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float lo = fadd(fmul(x[e], cl[e]), fmul(y[e], nsl[e]));
              const float hi = fadd(fmul(y[e], ch[e]), fmul(x[e], sh[e]));
              pl[e >> 2] |= (uint32_t)quant_int<FIVE>(lo, qo) << (8 * (e & 3));
              ph[e >> 2] |= (uint32_t)quant_int<FIVE>(hi, qo) << (8 * (e & 3));
            }
            uint8_t* dst = dst0 + hh * strideh2;
            if (ok[u]) {
              *reinterpret_cast<uint2*>(dst + dl) = make_uint2(pl[0], pl[1]);
              *reinterpret_cast<uint2*>(dst + dh) = make_uint2(ph[0], ph[1]);
            }
            int csum = (int)__dp4a(pl[0], 0x01010101u, 0u);
            csum = (int)__dp4a(pl[1], 0x01010101u, (unsigned)csum);
            csum = (int)__dp4a(ph[0], 0x01010101u, (unsigned)csum);
            csum = (int)__dp4a(ph[1], 0x01010101u, (unsigned)csum);
#pragma unroll
            for (int dd = LPH >> 1; dd > 0; dd >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, dd);
            if (ok[u] && sub2 == 0) rs0[int64_t(hh) * a.T] = csum;
          }
        }
      };
      heads2(a.nh, row, a.sq_in, a.oq_in, qq, a.q + (int64_t(b) * a.nh * a.T + t) * HD, a.rsq + int64_t(b) * a.nh * a.T + t);
      heads2(a.nkv, row + a.nh * HD, a.sk_in, a.ok_in, qk, a.k + (int64_t(b) * a.nkv * a.T + t) * HD, a.rsk + int64_t(b) * a.nkv * a.T + t);
    } else {
      // cos / sin of this lane's head dims (shared by every head of the token).  Dims beyond the rotary width (partial rotary,
      // hm:489-501) use cos 1, sin 0 and themselves as partner -- x*1 + x*0 == x exactly -- so the head loops are branch
      // free; the sign of rotate_half (hm:338-344: cat(-x2, x1)) is folded into sin (fmul(-y, s) == fmul(y, -s)).
      float cv[WPT][4], sv[WPT][4];
      int dpart[WPT];
  #pragma unroll
      for (int wi = 0; wi < WPT; ++wi) {
        const int d = (sub + wi * LPI) * 4;
        if (d < a.rot) {
          const bool neg = d < half;
          dpart[wi] = neg ? d + half : d - half;
          const float4 c = ldg4(a.cos + int64_t(t) * a.rot + d), sn = ldg4(a.sin + int64_t(t) * a.rot + d);
          cv[wi][0] = c.x; cv[wi][1] = c.y; cv[wi][2] = c.z; cv[wi][3] = c.w;
          sv[wi][0] = neg ? -sn.x : sn.x; sv[wi][1] = neg ? -sn.y : sn.y; sv[wi][2] = neg ? -sn.z : sn.z; sv[wi][3] = neg ? -sn.w : sn.w;
        } else {
          dpart[wi] = d;
          cv[wi][0] = cv[wi][1] = cv[wi][2] = cv[wi][3] = 1.f;
          sv[wi][0] = sv[wi][1] = sv[wi][2] = sv[wi][3] = 0.f;
        }
      }
      // ---- q heads, then k heads: HPW heads per iteration, LPI adjacent lanes per head.  src0: codes of head 0 of the kind in
      // the token's row; dst0 / rs0: head 0 of this token in the output layout (heads are T*HD / T elements apart)
      const int64_t strideh = int64_t(a.T) * HD;
      auto heads = [&](int nheads, const uint8_t* src0, float s_in, float o_in, const QParam& qo, uint8_t* dst0, int32_t* rs0) {
        constexpr int U = 2;                               // heads in flight per lane group: the loop is load-latency bound
        for (int hh0 = 0; hh0 < nheads; hh0 += U * HPW) {
          uint32_t wx[U][WPT], wy[U][WPT];
          bool ok[U];
  #pragma unroll
          for (int u = 0; u < U; ++u) {
            const int hh = hh0 + u * HPW + grp;
            ok[u] = hh < nheads;
            const uint8_t* src = src0 + (ok[u] ? hh : 0) * HD;
  #pragma unroll
            for (int wi = 0; wi < WPT; ++wi) {
              wx[u][wi] = __ldg(reinterpret_cast<const uint32_t*>(src + (sub + wi * LPI) * 4));
              wy[u][wi] = __ldg(reinterpret_cast<const uint32_t*>(src + dpart[wi]));
            }
          }
  #pragma unroll
          for (int u = 0; u < U; ++u) {
            const int hh = hh0 + u * HPW + grp;
            int csum = 0;
            uint8_t* dst = dst0 + hh * strideh;
  #pragma unroll
            for (int wi = 0; wi < WPT; ++wi) {
              const int d = (sub + wi * LPI) * 4;
              float x[4], y[4];
              unpack4(wx[u][wi], s_in, o_in, x);
              unpack4(wy[u][wi], s_in, o_in, y);
              uint32_t packed = 0;
  #pragma unroll
              for (int j = 0; j < 4; ++j)   // q_embed = q * cos + rotate_half(q) * sin
                packed |= (uint32_t)quant_int<FIVE>(fadd(fmul(x[j], cv[wi][j]), fmul(y[j], sv[wi][j])), qo) << (8 * j);
              if (ok[u]) *reinterpret_cast<uint32_t*>(dst + d) = packed;
              csum = (int)__dp4a(packed, 0x01010101u, (unsigned)csum);
            }
  #pragma unroll
            for (int dd = LPI >> 1; dd > 0; dd >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, dd);
            if (ok[u] && sub == 0) rs0[int64_t(hh) * a.T] = csum;
          }
        }
      };
      heads(a.nh, row, a.sq_in, a.oq_in, qq, a.q + (int64_t(b) * a.nh * a.T + t) * HD, a.rsq + int64_t(b) * a.nh * a.T + t);
      heads(a.nkv, row + a.nh * HD, a.sk_in, a.ok_in, qk, a.k + (int64_t(b) * a.nkv * a.T + t) * HD, a.rsk + int64_t(b) * a.nkv * a.T + t);
    }
    // ---- v: requant into the smem tile
    const uint8_t* vsrc = row + heads_qk * HD;
    for (int c = lane * 4; c < vw; c += 128) {
      float x[4];
      unpack4(__ldg(reinterpret_cast<const uint32_t*>(vsrc + c)), a.sv_in, a.ov_in, x);
      uint32_t packed = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) packed |= (uint32_t)quant_int<FIVE>(x[j], qv) << (8 * j);
      *reinterpret_cast<uint32_t*>(vs + tl * vstride + c) = packed;
    }
  }
  __syncthreads();
  // each thread writes one vT row segment: (kv head, d) x 16 tokens; segments that straddle a sequence boundary or are
  // not 16-byte aligned fall back to byte stores.
  const int b0 = tok0 / a.T, t0 = tok0 % a.T;
  const bool fast = tok0 + kRopeTok <= M && t0 + kRopeTok <= a.T && (t0 & 15) == 0 && (a.T & 15) == 0;
  for (int c = threadIdx.x; c < vw; c += blockDim.x) {
    const int kvh = c / HD, d = c % HD;
    if (fast) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        w[j] = vs[(4 * j) * vstride + c] | (vs[(4 * j + 1) * vstride + c] << 8) | (vs[(4 * j + 2) * vstride + c] << 16) |
               (vs[(4 * j + 3) * vstride + c] << 24);
      *reinterpret_cast<uint4*>(a.vt + ((int64_t(b0) * a.nkv + kvh) * HD + d) * a.T + t0) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      for (int tl = 0; tl < kRopeTok; ++tl) {
        const int tok = tok0 + tl;
        if (tok >= M) break;
        const int b = tok / a.T, t = tok % a.T;
        a.vt[((int64_t(b) * a.nkv + kvh) * HD + d) * a.T + t] = vs[tl * vstride + c];
      }
    }
  }
}

template <int HD>
__global__ void __launch_bounds__(256) qrope_kernel(const RopeArgs a, const int five) {
  extern __shared__ __align__(16) uint8_t vs[];          // [16 tokens][nkv*hd + 4]
  if (five) qrope_body<HD, true>(a, vs); else qrope_body<HD, false>(a, vs);
}

// =====================================================================================================================
// K6: exact quantised causal attention == HFAttention.forward hm:510-534 with QMatMul qk_bmm / pv_bmm (qm:453-466):
//   I_ij = sum_d (q-oq)(k-ok)                            int8 tensor-core MMA (mma.sync m16n8k32, s32 accumulate)
//   c_ij = clamp(rne((float(I)*sq*sk)/s_s)+o_s, 0, qmax_s)        qk_bmm.output_quantizer (16 bit)
//   E_ij = (A[k>>8] * B[k&255]) >> 31,  k = cmax_i - c_ij          two-level exp table, A[i] = rne(2^31 exp(-256 i a)),
//                                                                  B[j] = rne(2^31 exp(-j a)), a = s_s/sqrt(hd), float64 host
//   p_ij = float(E_ij)/float(sum_j E_ij)                 fp32 softmax of (c-o_s)*s_s/sqrt(hd) + causal mask, exact sum (u64)
//   cp_ij = clamp(rne(p/s_p)+o_p, 0, qmax_p)             pv_bmm.input_quantizer (16 bit, o_p == 0)
//   A_id = sum_j cp_ij*(v_jd - ov)                       two u8 MMAs on the hi/lo bytes of cp
//   out  = clamp(rne((float(A)*s_p*s_v)/s_out)+o_out, 0, 255)     pv_bmm.output_quantizer, written token-major [M, nh*hd]
// Three passes over the key tiles (row max, row sum, P.V): the [T,T] score matrix never touches HBM (the reference
// materialises it ~9x per layer in fp32).  One CTA = 64 query rows of one head; each warp owns 16 rows.  K/V tiles are
// double buffered with cp.async; every requantisation is the branch-free exact division of common.cuh.
// =====================================================================================================================

__device__ __forceinline__ void mma_u8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src, bool valid) {
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int HD, int DV, bool FIVE>
__global__ void __launch_bounds__(128, (HD <= 64 ? 4 : 2)) qattn_kernel(const AttnArgs a) {
  constexpr int KT = 64;                      // keys per tile (two 32-key halves, processed by a rolled loop)
  constexpr int QT = 64;                      // queries per CTA
  constexpr int KSTR = HD + 16;               // padded row strides (bank-conflict free fragment loads)
  constexpr int VSTR = KT + 16;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  uint8_t* sk = smem_attn;                                   // [2][KT * KSTR]
  uint8_t* sv = sk + 2 * KT * KSTR;                          // [2][DV * VSTR]
  int* s_rsk = reinterpret_cast<int*>(sv + 2 * DV * VSTR);   // [2][KT]
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(s_rsk + 2 * KT);   // [512]

  // head dims above 128 are split into HD/DV output chunks (each CTA still contracts the full HD for the scores)
  constexpr int NCH = HD / DV;
  const int d0 = (blockIdx.x % NCH) * DV;
  const int qt = (gridDim.x / NCH) - 1 - (blockIdx.x / NCH);  // heavy (late) query tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (a.nh / a.nkv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int q0 = qt * QT;
  const uint8_t* qbase = a.q + ((int64_t(b) * a.nh + h) * a.T) * HD;
  const uint8_t* kbase = a.k + ((int64_t(b) * a.nkv + kvh) * a.T) * HD;
  const uint8_t* vbase = a.vt + ((int64_t(b) * a.nkv + kvh) * HD + d0) * a.T;
  const int32_t* rskb = a.rsk + (int64_t(b) * a.nkv + kvh) * a.T;
  const bool v_aligned = (a.T & 15) == 0 && ((reinterpret_cast<uintptr_t>(a.vt) & 15) == 0);

  for (int i = threadIdx.x; i < 512; i += 128) s_tab[i] = __ldg(a.lut + i);

  // ---- Q fragments of this warp's 16 rows, straight from global (rows beyond T read as zero)
  uint32_t qa[HD / 32][4];
  const int r_lo = warp * 16 + g, r_hi = r_lo + 8;
  const int qi_lo = q0 + r_lo, qi_hi = q0 + r_hi;      // absolute query positions of this thread's two rows
#pragma unroll
  for (int ks = 0; ks < HD / 32; ++ks) {
    qa[ks][0] = qi_lo < a.T ? __ldg(reinterpret_cast<const uint32_t*>(qbase + int64_t(qi_lo) * HD + ks * 32 + 4 * t4)) : 0u;
    qa[ks][1] = qi_hi < a.T ? __ldg(reinterpret_cast<const uint32_t*>(qbase + int64_t(qi_hi) * HD + ks * 32 + 4 * t4)) : 0u;
    qa[ks][2] = qi_lo < a.T ? __ldg(reinterpret_cast<const uint32_t*>(qbase + int64_t(qi_lo) * HD + ks * 32 + 16 + 4 * t4)) : 0u;
    qa[ks][3] = qi_hi < a.T ? __ldg(reinterpret_cast<const uint32_t*>(qbase + int64_t(qi_hi) * HD + ks * 32 + 16 + 4 * t4)) : 0u;
  }
  const int32_t* rsqb = a.rsq + (int64_t(b) * a.nh + h) * a.T;
  const int ioq = (int)a.oq, iok = (int)a.ok, iov = (int)a.ov;
  // I = acc - iok*rsq - ioq*rsk + HD*ioq*iok = acc + colc[key] + rc[row]
  const int rc_lo = HD * ioq * iok - iok * (qi_lo < a.T ? __ldg(rsqb + qi_lo) : 0);
  const int rc_hi = HD * ioq * iok - iok * (qi_hi < a.T ? __ldg(rsqb + qi_hi) : 0);
  const int n_ktiles = (min(q0 + QT, a.T) + KT - 1) / KT;   // causal: keys <= last query of the tile
  const int total_steps = 3 * n_ktiles;

  auto issue_loads = [&](int step) {
    const int buf = step & 1;
    const int kt = step % n_ktiles;
    const bool with_v = step >= 2 * n_ktiles;
    const int k0 = kt * KT;
    uint8_t* skb = sk + buf * KT * KSTR;
    for (int i = threadIdx.x; i < KT * (HD / 16); i += 128) {
      const int r = i / (HD / 16), c = i % (HD / 16);
      const bool ok = k0 + r < a.T;
      cp_async16(skb + r * KSTR + c * 16, kbase + int64_t(ok ? k0 + r : 0) * HD + c * 16, ok);
    }
    if (threadIdx.x < KT) {
      const bool ok = k0 + threadIdx.x < a.T;
      cp_async4(s_rsk + buf * KT + threadIdx.x, rskb + (ok ? k0 + threadIdx.x : 0), ok);
    }
    if (with_v) {
      uint8_t* svb = sv + buf * DV * VSTR;
      for (int i = threadIdx.x; i < DV * (KT / 16); i += 128) {
        const int d = i / (KT / 16), c = i % (KT / 16);
        const uint8_t* src = vbase + int64_t(d) * a.T + k0 + c * 16;
        if (v_aligned) {
          const bool ok = k0 + c * 16 + 16 <= a.T;
          cp_async16(svb + d * VSTR + c * 16, ok ? src : vbase, ok);
        } else {
          uint8_t tmp[16];
          for (int j = 0; j < 16; ++j) tmp[j] = (k0 + c * 16 + j < a.T) ? src[j] : 0;
          *reinterpret_cast<uint4*>(svb + d * VSTR + c * 16) = *reinterpret_cast<uint4*>(tmp);
        }
      }
    }
    cp_async_commit();
  };

  // per-thread state of its two rows
  const QParam qs = make_qparam(a.s_s, a.o_s, a.qmax_s);
  const QParam qp = make_qparam(a.s_p, 0.f, a.qmax_p);
  const QParam qo = make_qparam(a.s_out, a.o_out, 255.f);
  int mx_lo = INT_MIN, mx_hi = INT_MIN;          // row maxima of acc + colc
  int cm_lo = 0, cm_hi = 0;                      // bits of (magic + clamped code - o_s) of the row maximum
  unsigned long long sum_lo = 0, sum_hi = 0;
  float den_lo = 1.f, den_hi = 1.f, rden_lo = 1.f, rden_hi = 1.f;
  int oacc[DV / 8][4];
#pragma unroll
  for (int i = 0; i < DV / 8; ++i) { oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0; }
  int psum_lo = 0, psum_hi = 0;                 // sum_j cp_ij (zero-point correction of V)

  // accumulators of 32 keys (4 n-tiles) for this warp's 16 rows in the MMA C layout, key zero-point term included
  auto score_half = [&](const uint8_t* skh, const int* rkh, int (&I)[4][4]) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      int acc[4] = {0, 0, 0, 0};
#pragma unroll
      for (int ks = 0; ks < HD / 32; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(skh + (nt * 8 + g) * KSTR + ks * 32 + 4 * t4);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(skh + (nt * 8 + g) * KSTR + ks * 32 + 16 + 4 * t4);
        mma_u8(acc, qa[ks], b0, b1);
      }
      const int2 rk2 = *reinterpret_cast<const int2*>(rkh + nt * 8 + 2 * t4);
      const int c0 = -ioq * rk2.x, c1 = -ioq * rk2.y;
      I[nt][0] = acc[0] + c0; I[nt][1] = acc[1] + c1; I[nt][2] = acc[2] + c0; I[nt][3] = acc[3] + c1;
    }
  };
  // bits of magic + (clamped code - o_s); differences of these bits are differences of codes
  auto score_bits = [&](int I) -> int { return __float_as_int(quant_magic<FIVE>(fmul(__int2float_rn(I), a.sqk), qs)); };
  auto exp_tab = [&](int k) -> uint32_t {
    const uint32_t ea = s_tab[__byte_perm((uint32_t)k, 0u, 0x4441)], eb = s_tab[256 + (k & 255)];   // (k >> 8) & 255: masked lanes stay in range
    return (uint32_t)(((unsigned long long)ea * eb) >> 31);
  };

  // one 64-key tile of one pass; DIAG = the tile holds keys beyond some query of the CTA (causal / length mask)
  auto tile = [&](auto diag_tag, int pass, int kt, const uint8_t* skb, const uint8_t* svb, const int* rk) {
    constexpr bool DIAG = decltype(diag_tag)::value;
#pragma unroll 1
    for (int hf = 0; hf < 2; ++hf) {
      int I[4][4];
      score_half(skb + hf * 32 * KSTR, rk + hf * 32, I);
      const int key0 = kt * KT + hf * 32 + 2 * t4;
      if (pass == 0) {
        // ---- pass 1: row maxima (the code is monotone in I)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int key = key0 + nt * 8;
          if (!DIAG || key <= qi_lo) mx_lo = max(mx_lo, I[nt][0]);
          if (!DIAG || key + 1 <= qi_lo) mx_lo = max(mx_lo, I[nt][1]);
          if (!DIAG || key <= qi_hi) mx_hi = max(mx_hi, I[nt][2]);
          if (!DIAG || key + 1 <= qi_hi) mx_hi = max(mx_hi, I[nt][3]);
        }
      } else if (pass == 1) {
        // ---- pass 2: exact row sums of E
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int key = key0 + nt * 8;
          const uint32_t e0 = exp_tab(cm_lo - score_bits(I[nt][0] + rc_lo)), e1 = exp_tab(cm_lo - score_bits(I[nt][1] + rc_lo));
          const uint32_t e2 = exp_tab(cm_hi - score_bits(I[nt][2] + rc_hi)), e3 = exp_tab(cm_hi - score_bits(I[nt][3] + rc_hi));
          sum_lo += (!DIAG || key <= qi_lo) ? e0 : 0u; sum_lo += (!DIAG || key + 1 <= qi_lo) ? e1 : 0u;
          sum_hi += (!DIAG || key <= qi_hi) ? e2 : 0u; sum_hi += (!DIAG || key + 1 <= qi_hi) ? e3 : 0u;
        }
      } else {
        // ---- pass 3: P codes and P.V.  (magic + code) keeps the 16-bit code in its low half-word (o_p == 0).
        // The division by the row sum always takes the two-step form (the sum may have an all-ones significand).
        auto prob_bits = [&](int Iv, int cm, float den, float rden, bool valid) -> uint32_t {
          const uint32_t e = (!DIAG || valid) ? exp_tab(cm - score_bits(Iv)) : 0u;
          const float pr = div_rn<true>(__uint2float_rn(e), den, rden);
          return (uint32_t)__float_as_int(quant_magic<FIVE>(pr, qp));
        };
        // A fragments (hi and lo bytes) with the key permutation
        // slot(4t..4t+3) of the k-step  <->  keys {8*0+2t, +1, 8*1+2t, +1};  slot(16+4t..) <-> n-tiles 2, 3
        uint32_t ahi[4], alo[4];
#pragma unroll
        for (int hsel = 0; hsel < 2; ++hsel) {      // hsel 0 -> a0/a1 (slots 4t..), 1 -> a2/a3 (slots 16+4t..)
          uint32_t pl[2], ph[2];                    // [n-tile w] packed code pairs of row lo / row hi
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            const int nt = hsel * 2 + w;
            const int key = key0 + nt * 8;
            const uint32_t c0 = prob_bits(I[nt][0] + rc_lo, cm_lo, den_lo, rden_lo, key <= qi_lo);
            const uint32_t c1 = prob_bits(I[nt][1] + rc_lo, cm_lo, den_lo, rden_lo, key + 1 <= qi_lo);
            const uint32_t c2 = prob_bits(I[nt][2] + rc_hi, cm_hi, den_hi, rden_hi, key <= qi_hi);
            const uint32_t c3 = prob_bits(I[nt][3] + rc_hi, cm_hi, den_hi, rden_hi, key + 1 <= qi_hi);
            pl[w] = __byte_perm(c0, c1, 0x5410);    // code0 | code1 << 16
            ph[w] = __byte_perm(c2, c3, 0x5410);
            psum_lo = (int)__dp2a_lo(pl[w], 0x0101u, (unsigned)psum_lo);
            psum_hi = (int)__dp2a_lo(ph[w], 0x0101u, (unsigned)psum_hi);
          }
          alo[hsel * 2 + 0] = __byte_perm(pl[0], pl[1], 0x6420); ahi[hsel * 2 + 0] = __byte_perm(pl[0], pl[1], 0x7531);
          alo[hsel * 2 + 1] = __byte_perm(ph[0], ph[1], 0x6420); ahi[hsel * 2 + 1] = __byte_perm(ph[0], ph[1], 0x7531);
        }
#pragma unroll
        for (int dn = 0; dn < DV / 8; ++dn) {
          int phi[4] = {0, 0, 0, 0};
          const uint8_t* vrow = svb + (dn * 8 + g) * VSTR + hf * 32 + 2 * t4;
          const uint32_t b0 = (uint32_t)*reinterpret_cast<const uint16_t*>(vrow) | ((uint32_t)*reinterpret_cast<const uint16_t*>(vrow + 8) << 16);
          const uint32_t b1 = (uint32_t)*reinterpret_cast<const uint16_t*>(vrow + 16) | ((uint32_t)*reinterpret_cast<const uint16_t*>(vrow + 24) << 16);
          mma_u8(phi, ahi, b0, b1);
          mma_u8(oacc[dn], alo, b0, b1);            // the lo-byte products accumulate straight into the output
#pragma unroll
          for (int j = 0; j < 4; ++j) oacc[dn][j] += phi[j] * 256;
        }
      }
    }
  };

  issue_loads(0);
  for (int step = 0; step < total_steps; ++step) {
    const int buf = step & 1;
    if (step + 1 < total_steps) { issue_loads(step + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const int pass = step / n_ktiles, kt = step % n_ktiles;
    const uint8_t* skb = sk + buf * KT * KSTR;
    const uint8_t* svb = sv + buf * DV * VSTR;
    const int* rk = s_rsk + buf * KT;
    if (kt == n_ktiles - 1) {                   // the only tile that needs the causal / length mask; ends the pass
      tile(std::true_type{}, pass, kt, skb, svb, rk);
      if (pass == 0) {
        mx_lo = max(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)); mx_lo = max(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
        mx_hi = max(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)); mx_hi = max(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
        cm_lo = score_bits(mx_lo + rc_lo); cm_hi = score_bits(mx_hi + rc_hi);
      } else if (pass == 1) {
        sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1); sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
        sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1); sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
        den_lo = __ull2float_rn(sum_lo); den_hi = __ull2float_rn(sum_hi);
        rden_lo = __frcp_rn(den_lo); rden_hi = __frcp_rn(den_hi);
      }
    } else {
      tile(std::false_type{}, pass, kt, skb, svb, rk);
    }
    __syncthreads();                              // everyone is done with `buf` before step+2's loads are issued into it
  }
  psum_lo += __shfl_xor_sync(0xffffffffu, psum_lo, 1); psum_lo += __shfl_xor_sync(0xffffffffu, psum_lo, 2);
  psum_hi += __shfl_xor_sync(0xffffffffu, psum_hi, 1); psum_hi += __shfl_xor_sync(0xffffffffu, psum_hi, 2);

  // ---- epilogue: remove the V zero point, requantise, store token-major
  int csum_lo = 0, csum_hi = 0;
  const int ldo = a.nh * HD;
#pragma unroll
  for (int dn = 0; dn < DV / 8; ++dn) {
    const int d = d0 + dn * 8 + 2 * t4;
    int code[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int A = oacc[dn][j] - iov * (j < 2 ? psum_lo : psum_hi);
      code[j] = quant_int<FIVE>(fmul(__int2float_rn(A), a.spv), qo);
    }
    if (qi_lo < a.T) {
      *reinterpret_cast<uint16_t*>(a.out + (int64_t(b) * a.T + qi_lo) * ldo + h * HD + d) = (uint16_t)(code[0] | (code[1] << 8));
      csum_lo += code[0] + code[1];
    }
    if (qi_hi < a.T) {
      *reinterpret_cast<uint16_t*>(a.out + (int64_t(b) * a.T + qi_hi) * ldo + h * HD + d) = (uint16_t)(code[2] | (code[3] << 8));
      csum_hi += code[2] + code[3];
    }
  }
  csum_lo += __shfl_xor_sync(0xffffffffu, csum_lo, 1); csum_lo += __shfl_xor_sync(0xffffffffu, csum_lo, 2);
  csum_hi += __shfl_xor_sync(0xffffffffu, csum_hi, 1); csum_hi += __shfl_xor_sync(0xffffffffu, csum_hi, 2);
  if (t4 == 0 && a.rowsum_out) {
    if (qi_lo < a.T) atomicAdd(a.rowsum_out + int64_t(b) * a.T + qi_lo, csum_lo);
    if (qi_hi < a.T) atomicAdd(a.rowsum_out + int64_t(b) * a.T + qi_hi, csum_hi);
  }
}

// =====================================================================================================================
// K6, single-QK-pass variant (the one the engine runs whenever the score codes of a CTA fit in shared memory).
// Same arithmetic as qattn_kernel above, restructured so that Q.K^T and the score requantisation are done ONCE:
//   pass A  stream K in 128-key stages; I -> 16-bit score code c; codes parked in shared memory (2 B / score), row max
//   pass B  shared memory only: E = tab(cmax - c), exact u64 row sums
//   pass C  stream V in 128-key stages; E again -> p -> 16-bit prob code -> hi/lo byte MMAs against V
// One CTA = 32 query rows of one head = 2 row groups x 4 key-split warps: warp (rg, w) owns the 32-key chunks
// c = 4*stage + w of row group rg, keeps their codes in a private, lane-major (conflict-free, 16-byte vectorised)
// slab, and the four key-split warps of a row group combine max / sum / P.V partials through shared memory.
// K rows are written to shared memory in the order that makes the QK^T accumulator fragment of a thread hold keys
// 4t..4t+3 (the A-fragment slots of the P.V MMA), so V fragments are plain 32-bit loads of the natural vT layout.
// The two exp tables are replicated 4x / 8x across banks (lane & 3 / lane & 7 picks the copy): random-index LUT reads
// were the LSU hot spot of the three-pass kernel.  CTAs are persistent (tables filled once) and walk the work items
// round-robin, heaviest query tiles first; K/V stages go through a three-buffer cp.async ring with one barrier each.
// =====================================================================================================================
constexpr int kA4Stage = 128;     // keys per staged tile
constexpr int kA4RepA = 4;        // bank replication of the exp tables: A (index k >> 8), 16 B per entry
constexpr int kA4RepB = 8;        //                                     B (index k & 255), 32 B per entry
constexpr int kA4TabBytes = 256 * (kA4RepA + kA4RepB) * 4;
// staged-tile ring depth: 3 (one barrier per stage) where it fits next to the score codes, 2 (two barriers per stage) for
// head dims above 128, whose 128-key K stage is 34 KB
template <int HD>
__host__ __device__ constexpr int a4_bufs() { return HD > 128 ? 2 : 3; }

template <int HD, int DV>
__host__ __device__ constexpr int a4_stage_bytes() {
  return (kA4Stage * (HD + 16) > DV * (kA4Stage + 16)) ? kA4Stage * (HD + 16) : DV * (kA4Stage + 16);
}
template <int DV>
__host__ __device__ constexpr int a4_red_bytes() { return 8 * 32 * (DV / 2) * 4; }
template <int HD, int DV>
static size_t a4_smem_bytes(int cpw) {
  const size_t region0 = std::max<size_t>(size_t(8) * cpw * 1024, a4_red_bytes<DV>());
  return region0 + a4_bufs<HD>() * size_t(a4_stage_bytes<HD, DV>()) + a4_bufs<HD>() * kA4Stage * 4 + kA4TabBytes + 128 * 4 + 128 * 8 + 32 * 4;
}

template <int HD, int DV, bool FIVE>
__global__ void __launch_bounds__(256, (HD <= 64 ? 2 : 1)) qattn4_kernel(const AttnArgs a, const int cpw, const int n_items) {
  constexpr int ST = kA4Stage;
  constexpr int KSTR = HD + 16;               // padded row strides (bank-conflict free fragment loads)
  constexpr int VSTR = ST + 16;
  constexpr int STAGE = a4_stage_bytes<HD, DV>();
  constexpr int NCH = HD / DV;
  constexpr int NDN = DV / 8;                 // output n-tiles per warp
  constexpr int CPR = HD / 16;                // 16-byte chunks per K row
  constexpr int NKL = ST * CPR / 256;         // K cp.async per thread and stage
  constexpr int NVL = DV * (ST / 16) / 256;   // V cp.async per thread and stage
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int code_bytes = 8 * cpw * 1024;
  const int region0 = code_bytes > a4_red_bytes<DV>() ? code_bytes : a4_red_bytes<DV>();
  uint4* s_codes = reinterpret_cast<uint4*>(smem_attn);                    // [8 warps][cpw][2 rows][32 lanes] uint4
  int* s_red = reinterpret_cast<int*>(smem_attn);                          // aliases the codes after pass C
  constexpr int BUFS = a4_bufs<HD>();
  uint8_t* s_stage = smem_attn + region0;                                  // [BUFS][STAGE]
  int* s_rsk = reinterpret_cast<int*>(s_stage + BUFS * STAGE);             // [BUFS][ST]
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(s_rsk + BUFS * ST);        // A [256][kA4RepA] then B [256][kA4RepB]
  // head dims above DV: the output chunks of DV dims are produced one after the other from the parked codes (pass C runs
  // HD / DV times, passes A and B once); their cross-warp partials go through the then idle stage buffers
  static_assert(NCH == 1 || BUFS * STAGE >= a4_red_bytes<DV>(), "stage ring too small for the P.V partials");
  int* s_redp = NCH == 1 ? s_red : reinterpret_cast<int*>(s_stage);
  int* s_xi = reinterpret_cast<int*>(s_tab + kA4TabBytes / 4);             // [2 rg][4 w][16 rows]
  unsigned long long* s_xs = reinterpret_cast<unsigned long long*>(s_xi + 128);   // [2][4][16]
  int* s_cs = reinterpret_cast<int*>(s_xs + 128);                          // [2][16]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int rg = warp >> 2, w = warp & 3;
  const bool v_aligned = (a.T & 15) == 0 && ((reinterpret_cast<uintptr_t>(a.vt) & 15) == 0);
  const int nqt = (a.T + 31) >> 5;
  const int per_qt = a.B * a.nh;              // work items of one query-tile row, heavy (late) tiles first

  // ---- once per CTA: the two exp tables, entries replicated across banks (lane & 3 / lane & 7 picks the copy)
  static_assert(kA4RepA == 4 && kA4RepB == 8, "table fill writes one / two uint4 per entry");
  {
    const uint32_t va = __ldg(a.lut + threadIdx.x), vb = __ldg(a.lut + 256 + threadIdx.x);
    reinterpret_cast<uint4*>(s_tab)[threadIdx.x] = make_uint4(va, va, va, va);
    reinterpret_cast<uint4*>(s_tab)[256 + 2 * threadIdx.x] = make_uint4(vb, vb, vb, vb);
    reinterpret_cast<uint4*>(s_tab)[256 + 2 * threadIdx.x + 1] = make_uint4(vb, vb, vb, vb);
  }
  const uint8_t* tabA = reinterpret_cast<const uint8_t*>(s_tab) + (lane & (kA4RepA - 1)) * 4;
  const uint8_t* tabB = reinterpret_cast<const uint8_t*>(s_tab) + 256 * kA4RepA * 4 + (lane & (kA4RepB - 1)) * 4;

  // shared-memory row of key kk (0..127 within a stage): accumulator column (nt, j) of a 32-key chunk holds key
  // 16*(nt>>1) + 4*(j>>1) + 2*(nt&1) + (j&1), i.e. thread t4 (columns 2t4, 2t4+1) owns keys 4t4..4t4+3 of each half
  auto krow = [](int kk) -> int {
    const int r = kk & 15;
    const int nt = 2 * ((kk >> 4) & 1) + ((r >> 1) & 1), j = 2 * (r >> 2) + (r & 1);
    return (kk & ~31) + nt * 8 + j;
  };
  // per-thread copy descriptors (stage independent): K chunk l moves key kk_l, 16-byte column kc; V chunk l row vd_l
  const int kc = threadIdx.x % CPR;
  int k_kk[NKL], k_dst[NKL];
#pragma unroll
  for (int l = 0; l < NKL; ++l) {
    k_kk[l] = threadIdx.x / CPR + l * (256 / CPR);
    k_dst[l] = krow(k_kk[l]) * KSTR + kc * 16;
  }
  const int rsk_dst = krow(threadIdx.x & (ST - 1));
  const int vc = threadIdx.x & 7, vd0 = threadIdx.x >> 3;

  const QParam qs = make_qparam(a.s_s, a.o_s, a.qmax_s);
  const QParam qp = make_qparam(a.s_p, 0.f, a.qmax_p);
  const QParam qo = make_qparam(a.s_out, a.o_out, 255.f);
  const int ioq = (int)a.oq, iok = (int)a.ok, iov = (int)a.ov;
  // E(k) = (A[k >> 8] * B[k & 255]) >> 31; indices are masked so that discarded (masked) lanes stay in range
  auto exp_tab = [&](int k) -> uint32_t {
    const uint32_t ea = *reinterpret_cast<const uint32_t*>(tabA + (((uint32_t)k >> 8 & 255u) * (kA4RepA * 4)));
    const uint32_t eb = *reinterpret_cast<const uint32_t*>(tabB + (((uint32_t)k & 255u) * (kA4RepB * 4)));
    return (uint32_t)(((unsigned long long)ea * eb) >> 31);
  };
  // both codes of a packed word at once: kk = (cm | cm << 16) - word holds the two 16-bit differences (no borrow: every
  // code of a fully visible chunk is <= its row maximum); each table index is one shift + one mask (16 / 32 B entries)
  static_assert(kA4RepA * 4 == 16 && kA4RepB * 4 == 32, "exp_pair hard-codes the table entry sizes");
  auto exp_pair = [&](uint32_t kk, uint32_t& e0, uint32_t& e1) {
    const uint32_t a0 = *reinterpret_cast<const uint32_t*>(tabA + ((kk >> 4) & 0xFF0u));
    const uint32_t b0 = *reinterpret_cast<const uint32_t*>(tabB + ((kk << 5) & 0x1FE0u));
    const uint32_t a1 = *reinterpret_cast<const uint32_t*>(tabA + ((kk >> 20) & 0xFF0u));
    const uint32_t b1 = *reinterpret_cast<const uint32_t*>(tabB + ((kk >> 11) & 0x1FE0u));
    e0 = (uint32_t)(((unsigned long long)a0 * b0) >> 31);
    e1 = (uint32_t)(((unsigned long long)a1 * b1) >> 31);
  };
  // key offset (within its 32-key chunk) of accumulator element (nt, e) of this thread
  auto koff = [&](int nt, int e) -> int { return 16 * (nt >> 1) + 4 * t4 + 2 * (nt & 1) + e; };
  uint4* my_codes = s_codes + (size_t(warp) * cpw) * 64 + lane;

  // ======================= persistent loop over (query tile, batch, head[, d chunk]) work items =======================
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int qt = nqt - 1 - item / per_qt;
    const int rem = item % per_qt;
    const int h = rem % a.nh, b = rem / a.nh;
    const int kvh = h / (a.nh / a.nkv);
    const int q0 = qt * 32;
    const uint8_t* qbase = a.q + ((int64_t(b) * a.nh + h) * a.T) * HD;
    const uint8_t* kbase = a.k + ((int64_t(b) * a.nkv + kvh) * a.T) * HD;
    const uint8_t* vbase = a.vt + ((int64_t(b) * a.nkv + kvh) * HD) * a.T;       // advanced by DV rows per output chunk
    const int32_t* rskb = a.rsk + (int64_t(b) * a.nkv + kvh) * a.T;
    const int key_end = min(a.T, q0 + 32);                       // keys this item ever needs
    const int n_st = (qt + 1 + 3) >> 2;                          // stages of 4 chunks; chunk qt is the diagonal one

    auto issue_k = [&](int s, int buf) {
      uint8_t* skb = s_stage + buf * STAGE;
#pragma unroll
      for (int l = 0; l < NKL; ++l) {
        const int key = s * ST + k_kk[l];
        const bool ok = key < key_end;
        cp_async16(skb + k_dst[l], kbase + int64_t(ok ? key : 0) * HD + kc * 16, ok);
      }
      if (threadIdx.x < ST) {
        const int key = s * ST + threadIdx.x;
        const bool ok = key < key_end;
        cp_async4(s_rsk + buf * ST + rsk_dst, rskb + (ok ? key : 0), ok);
      }
      cp_async_commit();
    };
    auto issue_v = [&](int s, int buf) {
      uint8_t* svb = s_stage + buf * STAGE;
      const int k0 = s * ST + vc * 16;
#pragma unroll
      for (int l = 0; l < NVL; ++l) {
        const int d = vd0 + l * 32;
        const uint8_t* src = vbase + int64_t(d) * a.T + k0;
        if (v_aligned) {
          const bool ok = k0 + 16 <= a.T;
          cp_async16(svb + d * VSTR + vc * 16, ok ? src : vbase, ok);
        } else {
          uint8_t tmp[16];
          for (int j = 0; j < 16; ++j) tmp[j] = (k0 + j < a.T) ? src[j] : 0;
          *reinterpret_cast<uint4*>(svb + d * VSTR + vc * 16) = *reinterpret_cast<uint4*>(tmp);
        }
      }
      cp_async_commit();
    };

    issue_k(0, 0);
    if (n_st > 1) issue_k(1, 1);

    // ---- Q fragments of this row group's 16 rows, straight from global (rows beyond T read as zero)
    uint32_t qa[HD / 32][4];
    const int qi_lo = q0 + rg * 16 + g, qi_hi = qi_lo + 8;       // absolute query positions of this thread's two rows
#pragma unroll
    for (int ks = 0; ks < HD / 32; ++ks) {
      qa[ks][0] = qi_lo < a.T ? __ldg(reinterpret_cast<const uint32_t*>(qbase + int64_t(qi_lo) * HD + ks * 32 + 4 * t4)) : 0u;
      qa[ks][1] = qi_hi < a.T ? __ldg(reinterpret_cast<const uint32_t*>(qbase + int64_t(qi_hi) * HD + ks * 32 + 4 * t4)) : 0u;
      qa[ks][2] = qi_lo < a.T ? __ldg(reinterpret_cast<const uint32_t*>(qbase + int64_t(qi_lo) * HD + ks * 32 + 16 + 4 * t4)) : 0u;
      qa[ks][3] = qi_hi < a.T ? __ldg(reinterpret_cast<const uint32_t*>(qbase + int64_t(qi_hi) * HD + ks * 32 + 16 + 4 * t4)) : 0u;
    }
    const int32_t* rsqb = a.rsq + (int64_t(b) * a.nh + h) * a.T;
    // I = acc - iok*rsq - ioq*rsk + HD*ioq*iok = acc + colc[key] + rc[row]
    const int rc_lo = HD * ioq * iok - iok * (qi_lo < a.T ? __ldg(rsqb + qi_lo) : 0);
    const int rc_hi = HD * ioq * iok - iok * (qi_hi < a.T ? __ldg(rsqb + qi_hi) : 0);
    if (threadIdx.x < 32) s_cs[threadIdx.x] = 0;

    // ================================================ pass A: codes + row max ==========================================
    // three-buffer ring, one barrier per stage: the barrier of stage s also proves stage s-1 has been consumed, so its
    // buffer (== buffer of stage s+2) can be refilled right away
    int mx_lo = -1, mx_hi = -1;
    for (int s = 0; s < n_st; ++s) {
      const int buf = s % BUFS;
      if (s + 1 < n_st) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncthreads();
      if (BUFS == 3 && s + 2 < n_st) issue_k(s + 2, (s + 2) % BUFS);
      const int c = 4 * s + w;
      if (c <= qt) {
        const uint8_t* skh = s_stage + buf * STAGE + w * 32 * KSTR;
        const int* rkh = s_rsk + buf * ST + w * 32;
        const bool diag = c == qt;
        uint32_t wl[4], wh[4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          int acc[4] = {0, 0, 0, 0};
#pragma unroll
          for (int ks = 0; ks < HD / 32; ++ks) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(skh + (nt * 8 + g) * KSTR + ks * 32 + 4 * t4);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(skh + (nt * 8 + g) * KSTR + ks * 32 + 16 + 4 * t4);
            mma_u8(acc, qa[ks], b0, b1);
          }
          const int2 rk2 = *reinterpret_cast<const int2*>(rkh + nt * 8 + 2 * t4);
          const int c0 = -ioq * rk2.x, c1 = -ioq * rk2.y;
          int code[4];
          code[0] = quant_int<FIVE>(fmul(__int2float_rn(acc[0] + c0 + rc_lo), a.sqk), qs);
          code[1] = quant_int<FIVE>(fmul(__int2float_rn(acc[1] + c1 + rc_lo), a.sqk), qs);
          code[2] = quant_int<FIVE>(fmul(__int2float_rn(acc[2] + c0 + rc_hi), a.sqk), qs);
          code[3] = quant_int<FIVE>(fmul(__int2float_rn(acc[3] + c1 + rc_hi), a.sqk), qs);
          wl[nt] = (uint32_t)code[0] | ((uint32_t)code[1] << 16);
          wh[nt] = (uint32_t)code[2] | ((uint32_t)code[3] << 16);
          if (diag) {
            const int key = c * 32 + koff(nt, 0);
            if (key > qi_lo) code[0] = -1;
            if (key + 1 > qi_lo) code[1] = -1;
            if (key > qi_hi) code[2] = -1;
            if (key + 1 > qi_hi) code[3] = -1;
          }
          mx_lo = max(mx_lo, max(code[0], code[1]));
          mx_hi = max(mx_hi, max(code[2], code[3]));
        }
        my_codes[(s * 2 + 0) * 32] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        my_codes[(s * 2 + 1) * 32] = make_uint4(wh[0], wh[1], wh[2], wh[3]);
      }
      if (BUFS == 2 && s + 2 < n_st) { __syncthreads(); issue_k(s + 2, buf); }   // two-buffer ring: refill after everyone is done
    }
    mx_lo = max(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)); mx_lo = max(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = max(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)); mx_hi = max(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    if (t4 == 0) { s_xi[(rg * 4 + w) * 16 + g] = mx_lo; s_xi[(rg * 4 + w) * 16 + g + 8] = mx_hi; }
    __syncthreads();                                // row maxima published; every K buffer has been consumed
    issue_v(0, 0);                                  // V streams in underneath pass B
    if (n_st > 1) issue_v(1, 1);
    int cm_lo = -1, cm_hi = -1;                     // row maxima of the score codes
#pragma unroll
    for (int ww = 0; ww < 4; ++ww) { cm_lo = max(cm_lo, s_xi[(rg * 4 + ww) * 16 + g]); cm_hi = max(cm_hi, s_xi[(rg * 4 + ww) * 16 + g + 8]); }
    const uint32_t cmcm_lo = (uint32_t)cm_lo | ((uint32_t)cm_lo << 16), cmcm_hi = (uint32_t)cm_hi | ((uint32_t)cm_hi << 16);

    // ================================================ pass B: exact row sums of E ======================================
    unsigned long long sum_lo = 0, sum_hi = 0;
    for (int s = 0; 4 * s + w <= qt; ++s) {
      const int c = 4 * s + w;
      const uint4 vl = my_codes[(s * 2 + 0) * 32], vh = my_codes[(s * 2 + 1) * 32];
      const uint32_t wl[4] = {vl.x, vl.y, vl.z, vl.w}, wh[4] = {vh.x, vh.y, vh.z, vh.w};
      if (c < qt) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          uint32_t e0, e1, e2, e3;                  // E <= 2^31: pairs are summed in 64 bits
          exp_pair(cmcm_lo - wl[nt], e0, e1); exp_pair(cmcm_hi - wh[nt], e2, e3);
          sum_lo += (unsigned long long)e0 + e1; sum_hi += (unsigned long long)e2 + e3;
        }
      } else {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int key = c * 32 + koff(nt, 0);
          const uint32_t e0 = exp_tab(cm_lo - (int)(wl[nt] & 0xffffu)), e1 = exp_tab(cm_lo - (int)(wl[nt] >> 16));
          const uint32_t e2 = exp_tab(cm_hi - (int)(wh[nt] & 0xffffu)), e3 = exp_tab(cm_hi - (int)(wh[nt] >> 16));
          sum_lo += key <= qi_lo ? e0 : 0u; sum_lo += key + 1 <= qi_lo ? e1 : 0u;
          sum_hi += key <= qi_hi ? e2 : 0u; sum_hi += key + 1 <= qi_hi ? e3 : 0u;
        }
      }
    }
    sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1); sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
    sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1); sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
    if (t4 == 0) { s_xs[(rg * 4 + w) * 16 + g] = sum_lo; s_xs[(rg * 4 + w) * 16 + g + 8] = sum_hi; }

    // ================================================ pass C: P codes and P.V ==========================================
    // (the first barrier of the loop below also publishes the row sums)
    float den_lo = 1.f, den_hi = 1.f, rden_lo = 1.f, rden_hi = 1.f;
    bool den_five = true;
    int olo[NDN][4], ohi[NDN][4];
    int psum_lo = 0, psum_hi = 0;                 // sum_j cp_ij (zero-point correction of V)
    auto chunk_pv = [&](auto five_tag, auto diag_tag, int s, int c, const uint8_t* svb) {
      constexpr bool DF = decltype(five_tag)::value;
      constexpr bool DIAG = decltype(diag_tag)::value;
      const uint4 vl = my_codes[(s * 2 + 0) * 32], vh = my_codes[(s * 2 + 1) * 32];
      const uint32_t wl[4] = {vl.x, vl.y, vl.z, vl.w}, wh[4] = {vh.x, vh.y, vh.z, vh.w};
      // (magic + code) keeps the 16-bit prob code in its low half-word (o_p == 0); p >= 0: the lower clamp never binds
      auto prob_of = [&](uint32_t e, float den, float rden) -> uint32_t {
        const float pr = div_rn<DF>(__uint2float_rn(e), den, rden);
        return (uint32_t)__float_as_int(__fadd_rn(fminf(div_rn<FIVE>(pr, qp.s, qp.rs), qp.hi), kRoundMagic));
      };
      uint32_t ahi[4], alo[4];
#pragma unroll
      for (int hsel = 0; hsel < 2; ++hsel) {        // hsel 0 -> a0/a1 (slots 4t..), 1 -> a2/a3 (slots 16+4t..)
        uint32_t pl[2], ph[2];
#pragma unroll
        for (int ww = 0; ww < 2; ++ww) {
          const int nt = hsel * 2 + ww;
          const int key = c * 32 + koff(nt, 0);
          uint32_t e0, e1, e2, e3;
          if (DIAG) {
            e0 = key <= qi_lo ? exp_tab(cm_lo - (int)(wl[nt] & 0xffffu)) : 0u;
            e1 = key + 1 <= qi_lo ? exp_tab(cm_lo - (int)(wl[nt] >> 16)) : 0u;
            e2 = key <= qi_hi ? exp_tab(cm_hi - (int)(wh[nt] & 0xffffu)) : 0u;
            e3 = key + 1 <= qi_hi ? exp_tab(cm_hi - (int)(wh[nt] >> 16)) : 0u;
          } else {
            exp_pair(cmcm_lo - wl[nt], e0, e1); exp_pair(cmcm_hi - wh[nt], e2, e3);
          }
          const uint32_t c0 = prob_of(e0, den_lo, rden_lo), c1 = prob_of(e1, den_lo, rden_lo);
          const uint32_t c2 = prob_of(e2, den_hi, rden_hi), c3 = prob_of(e3, den_hi, rden_hi);
          pl[ww] = __byte_perm(c0, c1, 0x5410);     // code0 | code1 << 16
          ph[ww] = __byte_perm(c2, c3, 0x5410);
          psum_lo = (int)__dp2a_lo(pl[ww], 0x0101u, (unsigned)psum_lo);
          psum_hi = (int)__dp2a_lo(ph[ww], 0x0101u, (unsigned)psum_hi);
        }
        alo[hsel * 2 + 0] = __byte_perm(pl[0], pl[1], 0x6420); ahi[hsel * 2 + 0] = __byte_perm(pl[0], pl[1], 0x7531);
        alo[hsel * 2 + 1] = __byte_perm(ph[0], ph[1], 0x6420); ahi[hsel * 2 + 1] = __byte_perm(ph[0], ph[1], 0x7531);
      }
#pragma unroll
      for (int dn = 0; dn < NDN; ++dn) {
        const uint8_t* vrow = svb + (dn * 8 + g) * VSTR + w * 32 + 4 * t4;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(vrow), b1 = *reinterpret_cast<const uint32_t*>(vrow + 16);
        mma_u8(ohi[dn], ahi, b0, b1);
        mma_u8(olo[dn], alo, b0, b1);
      }
    };
#pragma unroll 1
    for (int ch = 0; ch < NCH; ++ch) {
    const int d0 = ch * DV;
    if (ch > 0) {                                   // (the first chunk's V stages were issued before pass B)
      vbase += int64_t(DV) * a.T;
      issue_v(0, 0);
      if (n_st > 1) issue_v(1, 1);
    }
#pragma unroll
    for (int i = 0; i < NDN; ++i) { olo[i][0] = olo[i][1] = olo[i][2] = olo[i][3] = 0; ohi[i][0] = ohi[i][1] = ohi[i][2] = ohi[i][3] = 0; }
    psum_lo = 0; psum_hi = 0;
    for (int s = 0; s < n_st; ++s) {
      const int buf = s % BUFS;
      if (s + 1 < n_st) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncthreads();
      if (BUFS == 3 && s + 2 < n_st) issue_v(s + 2, (s + 2) % BUFS);
      if (s == 0 && ch == 0) {
        unsigned long long t_lo = 0, t_hi = 0;
#pragma unroll
        for (int ww = 0; ww < 4; ++ww) { t_lo += s_xs[(rg * 4 + ww) * 16 + g]; t_hi += s_xs[(rg * 4 + ww) * 16 + g + 8]; }
        den_lo = __ull2float_rn(t_lo); den_hi = __ull2float_rn(t_hi);      // >= 1: the row maximum contributes E(0) = 2^31
        rden_lo = __frcp_rn(den_lo); rden_hi = __frcp_rn(den_hi);
        // the Markstein division needs its second step only for an all-ones significand of the divisor (common.cuh)
        den_five = __any_sync(0xffffffffu, mantissa_all_ones(den_lo) || mantissa_all_ones(den_hi));
      }
      const int c = 4 * s + w;
      const uint8_t* svb = s_stage + buf * STAGE;
      if (c < qt) {
        if (den_five) chunk_pv(std::true_type{}, std::false_type{}, s, c, svb);
        else chunk_pv(std::false_type{}, std::false_type{}, s, c, svb);
      } else if (c == qt) {
        chunk_pv(std::true_type{}, std::true_type{}, s, c, svb);
      }
      if (BUFS == 2 && s + 2 < n_st) { __syncthreads(); issue_v(s + 2, buf); }
    }
    psum_lo += __shfl_xor_sync(0xffffffffu, psum_lo, 1); psum_lo += __shfl_xor_sync(0xffffffffu, psum_lo, 2);
    psum_hi += __shfl_xor_sync(0xffffffffu, psum_hi, 1); psum_hi += __shfl_xor_sync(0xffffffffu, psum_hi, 2);
    __syncthreads();                               // every warp is done with this chunk's V stages (and, NCH == 1, its codes)
    if (t4 == 0) { s_xi[(rg * 4 + w) * 16 + g] = psum_lo; s_xi[(rg * 4 + w) * 16 + g + 8] = psum_hi; }
    // ---- combine the four key-split partials of each row group (over the dead codes, or the idle stage ring)
    int* my_red = s_redp + (size_t(rg * 4 + w) * (DV / 2)) * 32 + lane;
#pragma unroll
    for (int dn = 0; dn < NDN; ++dn)
#pragma unroll
      for (int j = 0; j < 4; ++j) my_red[(dn * 4 + j) * 32] = olo[dn][j] + ohi[dn][j] * 256;
    __syncthreads();
    psum_lo = 0; psum_hi = 0;
#pragma unroll
    for (int ww = 0; ww < 4; ++ww) { psum_lo += s_xi[(rg * 4 + ww) * 16 + g]; psum_hi += s_xi[(rg * 4 + ww) * 16 + g + 8]; }

    // ---- epilogue: warp w finalises output n-tiles [w*NDN/4, (w+1)*NDN/4): remove the V zero point, requantise, store
    int csum_lo = 0, csum_hi = 0;
    const int ldo = a.nh * HD;
#pragma unroll
    for (int i = 0; i < NDN / 4; ++i) {
      const int dn = w * (NDN / 4) + i;
      const int d = d0 + dn * 8 + 2 * t4;
      int code[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int A = 0;
#pragma unroll
        for (int ww = 0; ww < 4; ++ww) A += s_redp[(size_t(rg * 4 + ww) * (DV / 2) + dn * 4 + j) * 32 + lane];
        A -= iov * (j < 2 ? psum_lo : psum_hi);
        code[j] = quant_int<FIVE>(fmul(__int2float_rn(A), a.spv), qo);
      }
      if (qi_lo < a.T) {
        *reinterpret_cast<uint16_t*>(a.out + (int64_t(b) * a.T + qi_lo) * ldo + h * HD + d) = (uint16_t)(code[0] | (code[1] << 8));
        csum_lo += code[0] + code[1];
      }
      if (qi_hi < a.T) {
        *reinterpret_cast<uint16_t*>(a.out + (int64_t(b) * a.T + qi_hi) * ldo + h * HD + d) = (uint16_t)(code[2] | (code[3] << 8));
        csum_hi += code[2] + code[3];
      }
    }
    if (a.rowsum_out) {
      csum_lo += __shfl_xor_sync(0xffffffffu, csum_lo, 1); csum_lo += __shfl_xor_sync(0xffffffffu, csum_lo, 2);
      csum_hi += __shfl_xor_sync(0xffffffffu, csum_hi, 1); csum_hi += __shfl_xor_sync(0xffffffffu, csum_hi, 2);
      if (t4 == 0) { atomicAdd(s_cs + rg * 16 + g, csum_lo); atomicAdd(s_cs + rg * 16 + g + 8, csum_hi); }
    }
    __syncthreads();                               // partials consumed: the next chunk / item may overwrite them, s_xi, s_cs
    }                                              // output chunks
    if (a.rowsum_out && threadIdx.x < 32) {
      const int qi = q0 + threadIdx.x;
      if (qi < a.T) atomicAdd(a.rowsum_out + int64_t(b) * a.T + qi, s_cs[threadIdx.x]);
    }
  }
}

template <int HD, int DV>
static size_t attn_smem_bytes() { return size_t(2) * 64 * (HD + 16) + size_t(2) * DV * 80 + 2 * 64 * 4 + 512 * 4; }

template <int HD, int DV, bool FIVE>
static int launch_qattn2(Ctx* c, const AttnArgs& a, dim3 grid, cudaStream_t st) {
  const size_t smem = attn_smem_bytes<HD, DV>();
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(qattn_kernel<HD, DV, FIVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(c, MQ_RUNTIME_ERROR, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    attr_set = true;
  }
  qattn_kernel<HD, DV, FIVE><<<grid, 128, smem, st>>>(a);
  return check_launch(c, "mq_qattn");
}
template <int HD, int DV, bool FIVE>
static int launch_qattn4(Ctx* c, const AttnArgs& a, int cpw, size_t smem, cudaStream_t st) {
  static size_t attr_smem = 0;                   // opt-in grows monotonically with the longest sequence seen
  static int ctas_per_sm = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(qattn4_kernel<HD, DV, FIVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(c, MQ_RUNTIME_ERROR, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    attr_smem = smem;
    ctas_per_sm = 0;
  }
  if (ctas_per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, qattn4_kernel<HD, DV, FIVE>, 256, attr_smem) != cudaSuccess || ctas_per_sm < 1) {
      cudaGetLastError();
      ctas_per_sm = 1;
    }
    if (getenv("MQB200_DEBUG")) fprintf(stderr, "[mqb200] qattn<%d,%d>: %zu B smem, %d CTAs/SM\n", HD, DV, attr_smem, ctas_per_sm);
  }
  // persistent CTAs walk the (query tile, batch, head) items round-robin, heaviest query tiles first
  const long long n_items = (long long)((a.T + 31) / 32) * a.B * a.nh;
  if (n_items > 0x7fffffffLL) return fail(c, MQ_INVALID_ARGUMENT, "mq_qattn: too many work items");
  const int grid = (int)std::min<long long>(n_items, (long long)c->sm_count * ctas_per_sm);
  qattn4_kernel<HD, DV, FIVE><<<grid, 256, attr_smem, st>>>(a, cpw, (int)n_items);
  return check_launch(c, "mq_qattn");
}
// MQB200_QATTN picks the kernel (read per call so that the tests can exercise all of them in one process):
//   unset / "tc"  tcgen05 + TMA + TMEM kernel (qattn_tc.cu) where the shape is covered (hd 64 / 128, T % 16 == 0)
//   "smem"        mma.sync kernel with the score codes parked in shared memory (falls back to 3pass when they do not fit)
//   "3pass"       streaming three-pass mma.sync kernel (any T)
static int qattn_choice() {
  const char* e = getenv("MQB200_QATTN");
  if (e && e[0] == '3') return 2;
  if (e && e[0] == 's') return 1;
  if (e && e[0] == 't' && e[1] == 'c' && e[2] == '!') return 3;       // "tc!": fail instead of falling back (tests)
  return 0;
}
template <int HD, int DV>
static int launch_qattn(Ctx* c, const AttnArgs& a, dim3 grid, cudaStream_t st) {
  // scales with an all-ones significand need the two-step exact division everywhere (common.cuh: div_rn)
  const bool five = mantissa_all_ones(a.s_s) || mantissa_all_ones(a.s_p) || mantissa_all_ones(a.s_out);
  const int cpw = ((a.T + 31) / 32 + 3) / 4;     // 32-key chunks per key-split warp
  const size_t smem4 = a4_smem_bytes<HD, DV>(cpw);
  if (smem4 <= 227 * 1024 && qattn_choice() != 2)
    return five ? launch_qattn4<HD, DV, true>(c, a, cpw, smem4, st) : launch_qattn4<HD, DV, false>(c, a, cpw, smem4, st);
  return five ? launch_qattn2<HD, DV, true>(c, a, grid, st) : launch_qattn2<HD, DV, false>(c, a, grid, st);
}

}  // namespace mq

using namespace mq;

extern "C" {

int mq_qnorm(void* ctx, const float* x, int64_t rows, int H, int is_layernorm, float s_in, float o_in, float qmax_in,
             const float* w_fq, const float* bias, float alpha, float eps, float s_out, float o_out, float qmax_out,
             uint8_t* codes, int32_t* rowsum, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, x && w_fq && codes && rows > 0 && H > 0, "null pointer or empty input");
  MQ_REQUIRE(c, H % 4 == 0, "hidden size must be a multiple of 4");
  MQ_REQUIRE(c, qmax_out <= 255.f && qmax_in <= 65535.f, "norm output codes are 8 bit, input codes at most 16 bit");
  MQ_REQUIRE(c, o_in == rintf(o_in) && o_out == rintf(o_out), "integer engine kernels need integral offsets (qm:60)");
  MQ_REQUIRE(c, (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_fq) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(codes) & 3) == 0,
             "x / w_fq / bias must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  NormArgs a{x, rows, H, s_in, o_in, qmax_in, w_fq, bias, alpha, eps, s_out, o_out, qmax_out, codes, rowsum};
  if (rows <= 256 && H <= 8192) {                 // decode-sized inputs: one CTA per row
    const GvEpiArgs none{};
    if (is_layernorm) launch_pdl(qnorm_row_kernel<true, false>, dim3((unsigned)rows), dim3(256), 0, st, a, none);
    else launch_pdl(qnorm_row_kernel<false, false>, dim3((unsigned)rows), dim3(256), 0, st, a, none);
    return check_launch(c, "mq_qnorm");
  }
  // prefill-sized inputs with H 1024 / 2048: two warps per row (qnorm_split_kernel).  MQB200_QNORM=simple keeps the one-warp-per-
  // row kernel, =pipe the persistent bulk-copy pipeline (both measured slower on B200: profiles/r2_qnorm_variants.md)
  const char* env = getenv("MQB200_QNORM");
  const char mode = env ? env[0] : 'd';
  if ((H == 2048 || H == 1024) && rows >= 1024 && mode == 'p') {
    const size_t smem = size_t(H) * 4 * (2 + 4) + 64;
    const int per_sm = (int)std::min<size_t>(8, (220 * 1024) / (smem + 1024));
    const unsigned g = (unsigned)std::min<int64_t>((rows + 3) / 4, int64_t(c->sm_count) * per_sm);
#define MQ_NORMP(LN, NV)                                                                                                   \
  do {                                                                                                                     \
    static bool attr = false;                                                                                              \
    if (!attr) { cudaFuncSetAttribute(qnorm_pipe_kernel<LN, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; } \
    qnorm_pipe_kernel<LN, NV><<<g, 128, smem, st>>>(a);                                                                    \
  } while (0)
    if (is_layernorm) { if (H == 2048) MQ_NORMP(true, 16); else MQ_NORMP(true, 8); }
    else { if (H == 2048) MQ_NORMP(false, 16); else MQ_NORMP(false, 8); }
#undef MQ_NORMP
    return check_launch(c, "mq_qnorm");
  }
  if ((H == 2048 || H == 1024) && rows >= 1024 && mode != 's') {
    const unsigned g2 = (unsigned)((rows + 1) / 2);
    if (is_layernorm) { if (H == 2048) qnorm_split_kernel<true, 8><<<g2, 128, 0, st>>>(a); else qnorm_split_kernel<true, 4><<<g2, 128, 0, st>>>(a); }
    else { if (H == 2048) qnorm_split_kernel<false, 8><<<g2, 128, 0, st>>>(a); else qnorm_split_kernel<false, 4><<<g2, 128, 0, st>>>(a); }
    return check_launch(c, "mq_qnorm");
  }
  unsigned grid = (unsigned)((rows + 3) / 4);
#define MQ_NORM(LN, NV) qnorm_kernel<LN, NV><<<grid, 128, 0, st>>>(a)
  if (is_layernorm) {
    if (H == 2048) MQ_NORM(true, 16); else if (H == 1024) MQ_NORM(true, 8); else MQ_NORM(true, 0);
  } else {
    if (H == 2048) MQ_NORM(false, 16); else if (H == 1024) MQ_NORM(false, 8); else MQ_NORM(false, 0);
  }
#undef MQ_NORM
  return check_launch(c, "mq_qnorm");
}

int mq_qnorm_resid(void* ctx, float* x, int rows, int H, int is_layernorm, float s_in, float o_in, float qmax_in, const float* w_fq,
                   const float* bias, float alpha, float eps, float s_out, float o_out, float qmax_out, uint8_t* codes, int32_t* rowsum,
                   int32_t* acc, int ldacc, const int32_t* g_rowsum, const float* g_sxw, const int32_t* g_ow, const int32_t* g_c0,
                   const float* g_bias, const float* g_so, const float* g_oo, float g_qmax, int g_qgroup, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, x && w_fq && codes && rows > 0 && H > 0, "null pointer or empty input");
  MQ_REQUIRE(c, rows <= 256 && H <= 8192 && H % 4 == 0, "the fused form covers decode-sized inputs (rows <= 256, H <= 8192, H % 4 == 0)");
  MQ_REQUIRE(c, qmax_out <= 255.f && qmax_in <= 65535.f, "norm output codes are 8 bit, input codes at most 16 bit");
  MQ_REQUIRE(c, o_in == rintf(o_in) && o_out == rintf(o_out), "integer engine kernels need integral offsets (qm:60)");
  MQ_REQUIRE(c, acc && g_rowsum && g_sxw && g_ow && g_c0 && g_so && g_oo, "the fused residual epilogue needs acc / rowsum / sxw / ow / c0 / so / oo");
  MQ_REQUIRE(c, ldacc >= H && ldacc % 4 == 0 && g_qgroup > 0 && g_qgroup % 4 == 0 && g_qmax < 4194304.f, "bad accumulator stride / qgroup / qmax");
  MQ_REQUIRE(c, (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_fq) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(codes) & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(acc) & 15) == 0, "x / w_fq / bias / acc must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  NormArgs a{x, rows, H, s_in, o_in, qmax_in, w_fq, bias, alpha, eps, s_out, o_out, qmax_out, codes, rowsum};
  GvEpiArgs e{rows, H, acc, ldacc, g_rowsum, g_sxw, g_ow, g_c0, g_bias, g_so, g_oo, g_qgroup, g_qmax, nullptr, (int64_t)H, nullptr,
              nullptr, 1.f, 0.f, 255.f, x, nullptr};
  if (is_layernorm) launch_pdl(qnorm_row_kernel<true, true>, dim3((unsigned)rows), dim3(256), 0, st, a, e);
  else launch_pdl(qnorm_row_kernel<false, true>, dim3((unsigned)rows), dim3(256), 0, st, a, e);
  return check_launch(c, "mq_qnorm_resid");
}

int mq_qrope(void* ctx, const uint8_t* qkv, int ldq, int B, int T, int nh, int nkv, int hd, int rot, const float* in_qparams,
             const float* out_qparams, const float* cos, const float* sin, uint8_t* q, uint8_t* k, uint8_t* vt, int32_t* rsq,
             int32_t* rsk, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, qkv && in_qparams && out_qparams && cos && sin && q && k && vt && rsq && rsk, "null pointer");
  MQ_REQUIRE(c, B > 0 && T > 0 && nh > 0 && nkv > 0 && hd > 0 && rot >= 0 && rot <= hd, "bad shape");
  MQ_REQUIRE(c, (hd == 32 || hd == 64 || hd == 128 || hd == 256) && rot % 8 == 0, "head_dim must be 32/64/128/256 and the rotary width a multiple of 8");
  MQ_REQUIRE(c, ldq % 4 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 3) == 0 && (reinterpret_cast<uintptr_t>(cos) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(sin) & 15) == 0 && (reinterpret_cast<uintptr_t>(q) & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(k) & 3) == 0, "qkv/q/k must be 4-byte aligned (ldq % 4 == 0), cos/sin 16-byte aligned");
  RopeArgs a;
  a.qkv = qkv; a.ldq = ldq; a.B = B; a.T = T; a.nh = nh; a.nkv = nkv; a.hd = hd; a.rot = rot;
  a.sq_in = in_qparams[0]; a.oq_in = in_qparams[1]; a.sk_in = in_qparams[2]; a.ok_in = in_qparams[3];
  a.sv_in = in_qparams[4]; a.ov_in = in_qparams[5];
  a.sq = out_qparams[0]; a.oq = out_qparams[1]; a.sk = out_qparams[2]; a.ok = out_qparams[3]; a.sv = out_qparams[4];
  a.ov = out_qparams[5];
  MQ_REQUIRE(c, a.oq == rintf(a.oq) && a.ok == rintf(a.ok) && a.ov == rintf(a.ov), "integer engine kernels need integral offsets (qm:60)");
  a.cos = cos; a.sin = sin; a.q = q; a.k = k; a.vt = vt; a.rsq = rsq; a.rsk = rsk;
  const int M = B * T;
  size_t smem = size_t(kRopeTok) * (nkv * hd + 4);
  MQ_REQUIRE(c, smem <= 48 * 1024, "nkv*hd too large for the V transpose tile");
  const int five = mantissa_all_ones(a.sq) || mantissa_all_ones(a.sk) || mantissa_all_ones(a.sv);
  const unsigned grid = (unsigned)((M + kRopeTok - 1) / kRopeTok);
  cudaStream_t st = (cudaStream_t)stream;
  if (hd == 32) qrope_kernel<32><<<grid, 256, smem, st>>>(a, five);
  else if (hd == 64) qrope_kernel<64><<<grid, 256, smem, st>>>(a, five);
  else if (hd == 128) qrope_kernel<128><<<grid, 256, smem, st>>>(a, five);
  else qrope_kernel<256><<<grid, 256, smem, st>>>(a, five);
  return check_launch(c, "mq_qrope");
}

static int fill_attn_args(Ctx* c, AttnArgs& a, const float* qparams) {
  a.oq = qparams[0]; a.ok = qparams[1]; a.ov = qparams[2]; a.sqk = qparams[3]; a.s_s = qparams[4]; a.o_s = qparams[5];
  a.qmax_s = qparams[6]; a.s_p = qparams[7]; a.qmax_p = qparams[8]; a.spv = qparams[9]; a.s_out = qparams[10];
  a.o_out = qparams[11];
  MQ_REQUIRE(c, a.qmax_s <= 65535.f && a.qmax_p <= 65535.f, "score / probability codes are at most 16 bit");
  MQ_REQUIRE(c, a.o_s == rintf(a.o_s) && a.o_out == rintf(a.o_out), "integer engine kernels need integral offsets (qm:60)");
  return MQ_NO_ERROR;
}

int mq_qattn_shard(void* ctx, const uint8_t* q, const uint8_t* k, const uint8_t* vt, const int32_t* rsq, const int32_t* rsk, int B,
                   int Tq, int T, int q_start, int nh, int nkv, int hd, const float* qparams, const uint32_t* lut, uint8_t* out,
                   int32_t* rowsum_out, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, q && k && vt && rsq && rsk && qparams && lut && out, "null pointer");
  MQ_REQUIRE(c, B > 0 && T > 0 && Tq > 0 && q_start >= 0 && q_start + Tq <= T && nh > 0 && nkv > 0 && nh % nkv == 0, "bad shape");
  AttnArgs a;
  a.q = q; a.k = k; a.vt = vt; a.rsq = rsq; a.rsk = rsk; a.B = B; a.T = T; a.nh = nh; a.nkv = nkv; a.hd = hd;
  if (int rc = fill_attn_args(c, a, qparams)) return rc;
  a.lut = lut; a.out = out; a.rowsum_out = rowsum_out; a.q_start = q_start; a.Tq = Tq;
  MQ_REQUIRE(c, qattn_tc_supported(a), "mq_qattn_shard needs hd 64 / 128, T % 16 == 0, q_start % 128 == 0 and 16-byte aligned buffers");
  return launch_qattn_tc(c, a, (cudaStream_t)stream);
}

int mq_qattn(void* ctx, const uint8_t* q, const uint8_t* k, const uint8_t* vt, const int32_t* rsq, const int32_t* rsk, int B,
             int T, int nh, int nkv, int hd, const float* qparams, const uint32_t* lut, uint8_t* out, int32_t* rowsum_out,
             void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, q && k && vt && rsq && rsk && qparams && lut && out, "null pointer");
  MQ_REQUIRE(c, B > 0 && T > 0 && nh > 0 && nkv > 0 && nh % nkv == 0, "bad shape");
  MQ_REQUIRE(c, hd == 32 || hd == 64 || hd == 128 || hd == 256, "head_dim must be 32, 64, 128 or 256");
  MQ_REQUIRE(c, (reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0, "q/k must be 16-byte aligned");
  AttnArgs a;
  a.q = q; a.k = k; a.vt = vt; a.rsq = rsq; a.rsk = rsk; a.B = B; a.T = T; a.nh = nh; a.nkv = nkv; a.hd = hd;
  if (int rc = fill_attn_args(c, a, qparams)) return rc;
  a.lut = lut; a.out = out; a.rowsum_out = rowsum_out; a.q_start = 0; a.Tq = T;
  dim3 grid((T + 63) / 64 * (hd == 256 ? 2 : 1), nh, B);
  cudaStream_t st = (cudaStream_t)stream;
  const int choice = qattn_choice();
  if ((choice == 0 || choice == 3) && qattn_tc_supported(a)) return launch_qattn_tc(c, a, st);
  MQ_REQUIRE(c, choice != 3, "MQB200_QATTN=tc! but the shape is not covered by the tcgen05 kernel (hd 64/128, T % 16 == 0)");
  if (hd == 32) return launch_qattn<32, 32>(c, a, grid, st);
  if (hd == 64) return launch_qattn<64, 64>(c, a, grid, st);
  if (hd == 128) return launch_qattn<128, 128>(c, a, grid, st);
  return launch_qattn<256, 128>(c, a, grid, st);
}

}  // extern "C"
