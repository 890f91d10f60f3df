// K4/K5/K6 of the statically-quantised integer forward: norm, RoPE + requant, exact quantised causal attention.
// Every kernel consumes and produces integer codes; the arithmetic is restated 1:1 in oracle/int_ref.py.
#include "common.cuh"
#include "ctx.h"
#include <string>

namespace mq {

// =====================================================================================================================
// K4: QRMSNorm.forward (qm:515-531) with l2norm_as_rmsnorm (hm:187-195):
//   r = clamp(rne(x/s_in)+o_in, 0, qmax_in) - o_in ; x^ = r*s_in ; nrm = sqrt(float(sum r^2)) * s_in  (sum exact, u64)
//   t = w_fq * (alpha * (x^ / max(nrm, 1e-12))) (+ bias) ; code = clamp(rne(t/s_out)+o_out, 0, qmax_out)
// QLayerNorm.forward (qm:625-642): mean/var from the exact integer sums (double), y = ((x^-mean)*rstd)*w + b.
// One warp per row, the row stays in registers (H <= 32*kMaxPerLane); emits u8 codes + their row sum.
// =====================================================================================================================
constexpr int kNormMaxVec = 16;   // float4 per lane -> H <= 2048*... 32 lanes * 16 * 4 = 2048... extended by loop below

template <bool kLayerNorm>
__global__ void __launch_bounds__(128) qnorm_kernel(const float* __restrict__ x, int64_t rows, int H, float s_in, float o_in,
                                                     float qmax_in, const float* __restrict__ w_fq,
                                                     const float* __restrict__ bias, float alpha, float eps, float s_out,
                                                     float o_out, float qmax_out, uint8_t* __restrict__ codes,
                                                     int32_t* __restrict__ rowsum) {
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * H;
  // pass 1: integer statistics (row is re-read from L1/L2 in pass 2; H*4 bytes per warp)
  unsigned long long s2 = 0; long long s1 = 0;
  for (int k = lane * 4; k < H; k += 128) {
    float4 v = *reinterpret_cast<const float4*>(xr + k);
    float r0 = fsub(quant_code(v.x, s_in, o_in, 0.f, qmax_in), o_in), r1 = fsub(quant_code(v.y, s_in, o_in, 0.f, qmax_in), o_in);
    float r2 = fsub(quant_code(v.z, s_in, o_in, 0.f, qmax_in), o_in), r3 = fsub(quant_code(v.w, s_in, o_in, 0.f, qmax_in), o_in);
    long long i0 = (long long)r0, i1 = (long long)r1, i2 = (long long)r2, i3 = (long long)r3;
    s2 += (unsigned long long)(i0 * i0) + (unsigned long long)(i1 * i1) + (unsigned long long)(i2 * i2) + (unsigned long long)(i3 * i3);
    s1 += i0 + i1 + i2 + i3;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    s1 += __shfl_xor_sync(0xffffffffu, s1, d);
  }
  float denom = 1.f, mean = 0.f, rstd = 1.f;
  if (kLayerNorm) {
    const double m = (double)s1 / (double)H;
    const double var = (double)s2 / (double)H - m * m;
    mean = (float)(m * (double)s_in);
    rstd = (float)(1.0 / sqrt(var * (double)s_in * (double)s_in + (double)eps));
  } else {
    denom = fmaxf(fmul(__fsqrt_rn(__ull2float_rn(s2)), s_in), 1e-12f);
  }
  int csum = 0;
  for (int k = lane * 4; k < H; k += 128) {
    float4 v = *reinterpret_cast<const float4*>(xr + k);
    float4 w = __ldg(reinterpret_cast<const float4*>(w_fq + k));
    float xv[4] = {v.x, v.y, v.z, v.w}, wv[4] = {w.x, w.y, w.z, w.w};
    uint32_t packed = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xh = dequant(quant_code(xv[j], s_in, o_in, 0.f, qmax_in), s_in, o_in);
      float t;
      if (kLayerNorm) {
        t = fmul(fmul(fsub(xh, mean), rstd), wv[j]);
        if (bias) t = fadd(t, __ldg(bias + k + j));
      } else {
        t = fmul(wv[j], fmul(alpha, fdiv(xh, denom)));
        if (bias) t = fadd(t, __ldg(bias + k + j));
      }
      const int c = (int)quant_code(t, s_out, o_out, 0.f, qmax_out);
      csum += c;
      packed |= (uint32_t)c << (8 * j);
    }
    *reinterpret_cast<uint32_t*>(codes + row * H + k) = packed;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, d);
  if (lane == 0 && rowsum) rowsum[row] = csum;
}

// =====================================================================================================================
// K5: de-quantise q/k/v projection codes, RoPE (hm:338-367, partial rotary hm:489-501), re-quantise with the qk_bmm /
// pv_bmm input quantizers (qm:455-459) and write the attention layouts:
//   q  [B, nh, T, hd] u8      k [B, nkv, T, hd] u8      vT [B, nkv, hd, T] u8   (+ per-token code sums of q and k rows)
// One CTA per 32 tokens; V goes through a shared-memory transpose so that vT rows are written 32 bytes at a time.
// =====================================================================================================================
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

__global__ void __launch_bounds__(256) qrope_kernel(const RopeArgs a) {
  extern __shared__ uint8_t vs[];                       // [32 tokens][nkv*hd + 4]
  const int tok0 = blockIdx.x * 32;
  const int M = a.B * a.T;
  const int half = a.rot / 2;
  const int vstride = a.nkv * a.hd + 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- q and k: one warp per (token, head), lanes over the head dim
  const int heads_qk = a.nh + a.nkv;
  for (int item = warp; item < 32 * heads_qk; item += 8) {
    const int tl = item / heads_qk, hh = item % heads_qk;
    const int tok = tok0 + tl;
    if (tok >= M) continue;
    const int b = tok / a.T, t = tok % a.T;
    const bool is_q = hh < a.nh;
    const int h = is_q ? hh : hh - a.nh;
    const uint8_t* src = a.qkv + int64_t(tok) * a.ldq + (is_q ? h * a.hd : a.nh * a.hd + h * a.hd);
    const float s_in = is_q ? a.sq_in : a.sk_in, o_in = is_q ? a.oq_in : a.ok_in;
    const float s_o = is_q ? a.sq : a.sk, o_o = is_q ? a.oq : a.ok;
    uint8_t* dst = is_q ? a.q + ((int64_t(b) * a.nh + h) * a.T + t) * a.hd : a.k + ((int64_t(b) * a.nkv + h) * a.T + t) * a.hd;
    int csum = 0;
    for (int d = lane; d < a.hd; d += 32) {
      float out;
      const float xd = dequant((float)src[d], s_in, o_in);
      if (d < a.rot) {
        const float c = __ldg(a.cos + int64_t(t) * a.rot + d), s = __ldg(a.sin + int64_t(t) * a.rot + d);
        // q_embed = (q * cos) + (rotate_half(q) * sin); rotate_half = cat(-x2, x1)
        const float other = dequant((float)src[d < half ? d + half : d - half], s_in, o_in);
        const float rh = d < half ? -other : other;
        out = fadd(fmul(xd, c), fmul(rh, s));
      } else {
        out = xd;
      }
      const int code = (int)quant_code(out, s_o, o_o, 0.f, 255.f);
      dst[d] = (uint8_t)code;
      csum += code;
    }
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, dd);
    if (lane == 0) {
      if (is_q) a.rsq[(int64_t(b) * a.nh + h) * a.T + t] = csum;
      else a.rsk[(int64_t(b) * a.nkv + h) * a.T + t] = csum;
    }
  }
  // ---- v: requant into smem, then transposed store
  const int vw = a.nkv * a.hd;
  for (int idx = threadIdx.x; idx < 32 * vw; idx += 256) {
    const int tl = idx / vw, c = idx % vw;
    const int tok = tok0 + tl;
    uint8_t code = 0;
    if (tok < M) {
      const float xv = dequant((float)a.qkv[int64_t(tok) * a.ldq + (a.nh + a.nkv) * a.hd + c], a.sv_in, a.ov_in);
      code = (uint8_t)quant_code(xv, a.sv, a.ov, 0.f, 255.f);
    }
    vs[tl * vstride + c] = code;
  }
  __syncthreads();
  // each thread writes one vT row segment: (kv head, d) x 32 tokens.  tok0 % 32 == 0 and T % 32 == 0 is not required:
  // segments that straddle a sequence boundary fall back to byte stores.
  for (int c = threadIdx.x; c < vw; c += 256) {
    const int kvh = c / a.hd, d = c % a.hd;
    const int b0 = tok0 / a.T, t0 = tok0 % a.T;
    if (tok0 + 32 <= M && t0 + 32 <= a.T && (t0 & 15) == 0 && (a.T & 15) == 0) {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        w[j] = vs[(4 * j) * vstride + c] | (vs[(4 * j + 1) * vstride + c] << 8) | (vs[(4 * j + 2) * vstride + c] << 16) |
               (vs[(4 * j + 3) * vstride + c] << 24);
      uint4* dst = reinterpret_cast<uint4*>(a.vt + ((int64_t(b0) * a.nkv + kvh) * a.hd + d) * a.T + t0);
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
      for (int tl = 0; tl < 32; ++tl) {
        const int tok = tok0 + tl;
        if (tok >= M) break;
        const int b = tok / a.T, t = tok % a.T;
        a.vt[((int64_t(b) * a.nkv + kvh) * a.hd + d) * a.T + t] = vs[tl * vstride + c];
      }
    }
  }
}

// =====================================================================================================================
// K6: exact quantised causal attention == HFAttention.forward hm:510-534 with QMatMul qk_bmm / pv_bmm (qm:453-466):
//   I_ij = sum_d (q-oq)(k-ok)                            int8 tensor-core MMA (mma.sync m16n8k32, s32 accumulate)
//   c_ij = clamp(rne((float(I)*sq*sk)/s_s)+o_s, 0, qmax_s)        qk_bmm.output_quantizer (16 bit)
//   E_ij = LUT[cmax_i - c_ij]                            LUT[k] = rne(2^31 * exp(-k*s_s/sqrt(hd))), host float64
//   p_ij = float(E_ij)/float(sum_j E_ij)                 fp32 softmax of (c-o_s)*s_s/sqrt(hd) + causal mask, exact sum (u64)
//   cp_ij = clamp(rne(p/s_p)+o_p, 0, qmax_p)             pv_bmm.input_quantizer (16 bit, o_p == 0)
//   A_id = sum_j cp_ij*(v_jd - ov)                       two u8 MMAs on the hi/lo bytes of cp, folded in s32 per key tile
//   out  = clamp(rne((float(A)*s_p*s_v)/s_out)+o_out, 0, 255)     pv_bmm.output_quantizer, written token-major [M, nh*hd]
// Three passes over the key tiles (row max, row sum, P.V): the [T,T] score matrix never touches HBM (the reference
// materialises it ~9x per layer in fp32).  One CTA = 64 query rows of one head; each warp owns 16 rows.
// =====================================================================================================================
struct AttnArgs {
  const uint8_t *q, *k, *vt;
  const int32_t *rsq, *rsk;
  int B, T, nh, nkv, hd;
  float oq, ok, ov;            // integer zero points
  float sqk;                   // sq*sk
  float s_s, o_s, qmax_s;      // score quantizer
  const uint32_t* lut;         // [qmax_s+1]
  float s_p, qmax_p;           // prob quantizer (offset 0)
  float spv;                   // s_p*s_v
  float s_out, o_out;          // output quantizer (8 bit)
  uint8_t* out;                // [B*T, nh*hd]
  int32_t* rowsum_out;         // [B*T] atomically accumulated
};

__device__ __forceinline__ void mma_u8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int HD, int DV>
__global__ void __launch_bounds__(128) qattn_kernel(const AttnArgs a) {
  constexpr int KT = 64;                      // keys per tile
  constexpr int QT = 64;                      // queries per CTA
  constexpr int KSTR = HD + 16;               // padded row strides (bank-conflict free fragment loads)
  constexpr int VSTR = KT + 16;
  __shared__ __align__(16) uint8_t sq[QT * KSTR];
  __shared__ __align__(16) uint8_t sk[KT * KSTR];
  __shared__ __align__(16) uint8_t sv[DV * VSTR];
  __shared__ int s_rsk[KT];

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

  // ---- load the Q tile (zero rows beyond T)
  for (int i = threadIdx.x; i < QT * (HD / 16); i += 128) {
    const int r = i / (HD / 16), c = i % (HD / 16);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q0 + r < a.T) v = *reinterpret_cast<const uint4*>(qbase + int64_t(q0 + r) * HD + c * 16);
    *reinterpret_cast<uint4*>(sq + r * KSTR + c * 16) = v;
  }
  __syncthreads();
  // Q fragments of this warp's 16 rows stay in registers
  uint32_t qa[HD / 32][4];
  const int r_lo = warp * 16 + g, r_hi = r_lo + 8;
#pragma unroll
  for (int ks = 0; ks < HD / 32; ++ks) {
    qa[ks][0] = *reinterpret_cast<const uint32_t*>(sq + r_lo * KSTR + ks * 32 + 4 * t4);
    qa[ks][1] = *reinterpret_cast<const uint32_t*>(sq + r_hi * KSTR + ks * 32 + 4 * t4);
    qa[ks][2] = *reinterpret_cast<const uint32_t*>(sq + r_lo * KSTR + ks * 32 + 16 + 4 * t4);
    qa[ks][3] = *reinterpret_cast<const uint32_t*>(sq + r_hi * KSTR + ks * 32 + 16 + 4 * t4);
  }
  const int qi_lo = q0 + r_lo, qi_hi = q0 + r_hi;      // absolute query positions of this thread's two rows
  const int32_t* rsqb = a.rsq + (int64_t(b) * a.nh + h) * a.T;
  const int rsq_lo = qi_lo < a.T ? rsqb[qi_lo] : 0, rsq_hi = qi_hi < a.T ? rsqb[qi_hi] : 0;
  const int ioq = (int)a.oq, iok = (int)a.ok, iov = (int)a.ov;
  const int kconst = HD * ioq * iok;
  const int n_ktiles = (min(q0 + QT, a.T) + KT - 1) / KT;   // causal: keys <= last query of the tile

  auto load_k_tile = [&](int kt, bool with_v) {
    __syncthreads();
    const int k0 = kt * KT;
    for (int i = threadIdx.x; i < KT * (HD / 16); i += 128) {
      const int r = i / (HD / 16), c = i % (HD / 16);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (k0 + r < a.T) v = *reinterpret_cast<const uint4*>(kbase + int64_t(k0 + r) * HD + c * 16);
      *reinterpret_cast<uint4*>(sk + r * KSTR + c * 16) = v;
    }
    if (threadIdx.x < KT) s_rsk[threadIdx.x] = (k0 + threadIdx.x < a.T) ? rskb[k0 + threadIdx.x] : 0;
    if (with_v) {
      for (int i = threadIdx.x; i < DV * (KT / 16); i += 128) {
        const int d = i / (KT / 16), c = i % (KT / 16);
        uint4 v = make_uint4(0, 0, 0, 0);
        const uint8_t* src = vbase + int64_t(d) * a.T + k0 + c * 16;
        if (k0 + c * 16 + 16 <= a.T && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
          v = *reinterpret_cast<const uint4*>(src);
        } else {
          uint8_t tmp[16];
          for (int j = 0; j < 16; ++j) tmp[j] = (k0 + c * 16 + j < a.T) ? src[j] : 0;
          v = *reinterpret_cast<uint4*>(tmp);
        }
        *reinterpret_cast<uint4*>(sv + d * VSTR + c * 16) = v;
      }
    }
    __syncthreads();
  };

  // scores of one key tile for this warp's 16 rows: I[nt][0..3] in the mma C layout (zero points removed)
  auto score_tile = [&](int (&I)[KT / 8][4]) {
#pragma unroll
    for (int nt = 0; nt < KT / 8; ++nt) {
      int acc[4] = {0, 0, 0, 0};
#pragma unroll
      for (int ks = 0; ks < HD / 32; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(sk + (nt * 8 + g) * KSTR + ks * 32 + 4 * t4);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(sk + (nt * 8 + g) * KSTR + ks * 32 + 16 + 4 * t4);
        mma_u8(acc, qa[ks], b0, b1);
      }
      const int c0 = nt * 8 + 2 * t4;
      const int rk0 = s_rsk[c0], rk1 = s_rsk[c0 + 1];
      I[nt][0] = acc[0] - iok * rsq_lo - ioq * rk0 + kconst;
      I[nt][1] = acc[1] - iok * rsq_lo - ioq * rk1 + kconst;
      I[nt][2] = acc[2] - iok * rsq_hi - ioq * rk0 + kconst;
      I[nt][3] = acc[3] - iok * rsq_hi - ioq * rk1 + kconst;
    }
  };
  auto score_code = [&](int I) -> int {
    return (int)quant_code(fmul(__int2float_rn(I), a.sqk), a.s_s, a.o_s, 0.f, a.qmax_s);
  };

  // ---- pass 1: row maxima of I (the code is monotone in I)
  int mx_lo = INT_MIN, mx_hi = INT_MIN;
  for (int kt = 0; kt < n_ktiles; ++kt) {
    load_k_tile(kt, false);
    int I[KT / 8][4];
    score_tile(I);
#pragma unroll
    for (int nt = 0; nt < KT / 8; ++nt) {
      const int key = kt * KT + nt * 8 + 2 * t4;
      if (key <= qi_lo) mx_lo = max(mx_lo, I[nt][0]);
      if (key + 1 <= qi_lo) mx_lo = max(mx_lo, I[nt][1]);
      if (key <= qi_hi) mx_hi = max(mx_hi, I[nt][2]);
      if (key + 1 <= qi_hi) mx_hi = max(mx_hi, I[nt][3]);
    }
  }
  mx_lo = max(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)); mx_lo = max(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
  mx_hi = max(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)); mx_hi = max(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
  const int cmax_lo = score_code(mx_lo), cmax_hi = score_code(mx_hi);

  // ---- pass 2: exact row sums of E
  unsigned long long sum_lo = 0, sum_hi = 0;
  for (int kt = 0; kt < n_ktiles; ++kt) {
    load_k_tile(kt, false);
    int I[KT / 8][4];
    score_tile(I);
#pragma unroll
    for (int nt = 0; nt < KT / 8; ++nt) {
      const int key = kt * KT + nt * 8 + 2 * t4;
      if (key <= qi_lo) sum_lo += __ldg(a.lut + (cmax_lo - score_code(I[nt][0])));
      if (key + 1 <= qi_lo) sum_lo += __ldg(a.lut + (cmax_lo - score_code(I[nt][1])));
      if (key <= qi_hi) sum_hi += __ldg(a.lut + (cmax_hi - score_code(I[nt][2])));
      if (key + 1 <= qi_hi) sum_hi += __ldg(a.lut + (cmax_hi - score_code(I[nt][3])));
    }
  }
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1); sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1); sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
  const float den_lo = __ull2float_rn(sum_lo), den_hi = __ull2float_rn(sum_hi);

  // ---- pass 3: P codes and P.V
  int oacc[DV / 8][4];
#pragma unroll
  for (int i = 0; i < DV / 8; ++i) { oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0; }
  int psum_lo = 0, psum_hi = 0;                 // sum_j cp_ij (zero-point correction of V)
  auto prob_code = [&](int I, int cmax, float den, bool valid) -> int {
    if (!valid) return 0;
    const float e = __uint2float_rn(__ldg(a.lut + (cmax - score_code(I))));
    return (int)quant_code(fdiv(e, den), a.s_p, 0.f, 0.f, a.qmax_p);
  };
  for (int kt = 0; kt < n_ktiles; ++kt) {
    load_k_tile(kt, true);
    int I[KT / 8][4];
    score_tile(I);
    // P codes in the C layout, then packed straight into A fragments (hi and lo bytes) with the key permutation
    // slot(4t..4t+3) of k-step ks  <->  keys {8(4ks)+2t, +1, 8(4ks+1)+2t, +1};  slot(16+4t..) <-> n-tiles 4ks+2, 4ks+3
    uint32_t ahi[KT / 32][4], alo[KT / 32][4];
#pragma unroll
    for (int ks = 0; ks < KT / 32; ++ks) {
#pragma unroll
      for (int hsel = 0; hsel < 2; ++hsel) {      // hsel 0 -> a0/a1 (slots 4t..), 1 -> a2/a3 (slots 16+4t..)
        uint32_t hi_lo_row[2] = {0, 0}, lo_lo_row[2] = {0, 0};   // [row lo/hi]
#pragma unroll
        for (int w = 0; w < 2; ++w) {             // two n-tiles feed the four bytes
          const int nt = ks * 4 + hsel * 2 + w;
          const int key = kt * KT + nt * 8 + 2 * t4;
          const int c0 = prob_code(I[nt][0], cmax_lo, den_lo, key <= qi_lo);
          const int c1 = prob_code(I[nt][1], cmax_lo, den_lo, key + 1 <= qi_lo);
          const int c2 = prob_code(I[nt][2], cmax_hi, den_hi, key <= qi_hi);
          const int c3 = prob_code(I[nt][3], cmax_hi, den_hi, key + 1 <= qi_hi);
          psum_lo += c0 + c1; psum_hi += c2 + c3;
          hi_lo_row[0] |= ((uint32_t)(c0 >> 8) | ((uint32_t)(c1 >> 8) << 8)) << (16 * w);
          lo_lo_row[0] |= ((uint32_t)(c0 & 255) | ((uint32_t)(c1 & 255) << 8)) << (16 * w);
          hi_lo_row[1] |= ((uint32_t)(c2 >> 8) | ((uint32_t)(c3 >> 8) << 8)) << (16 * w);
          lo_lo_row[1] |= ((uint32_t)(c2 & 255) | ((uint32_t)(c3 & 255) << 8)) << (16 * w);
        }
        ahi[ks][hsel * 2 + 0] = hi_lo_row[0]; ahi[ks][hsel * 2 + 1] = hi_lo_row[1];
        alo[ks][hsel * 2 + 0] = lo_lo_row[0]; alo[ks][hsel * 2 + 1] = lo_lo_row[1];
      }
    }
#pragma unroll
    for (int dn = 0; dn < DV / 8; ++dn) {
      int phi[4] = {0, 0, 0, 0}, plo[4] = {0, 0, 0, 0};
#pragma unroll
      for (int ks = 0; ks < KT / 32; ++ks) {
        const uint8_t* vrow = sv + (dn * 8 + g) * VSTR + ks * 32 + 2 * t4;
        const uint32_t b0 = (uint32_t)*reinterpret_cast<const uint16_t*>(vrow) | ((uint32_t)*reinterpret_cast<const uint16_t*>(vrow + 8) << 16);
        const uint32_t b1 = (uint32_t)*reinterpret_cast<const uint16_t*>(vrow + 16) | ((uint32_t)*reinterpret_cast<const uint16_t*>(vrow + 24) << 16);
        mma_u8(phi, ahi[ks], b0, b1);
        mma_u8(plo, alo[ks], b0, b1);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) oacc[dn][j] += phi[j] * 256 + plo[j];
    }
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
      code[j] = (int)quant_code(fmul(__int2float_rn(A), a.spv), a.s_out, a.o_out, 0.f, 255.f);
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

}  // namespace mq

using namespace mq;

extern "C" {

int mq_qnorm(void* ctx, const float* x, int64_t rows, int H, int is_layernorm, float s_in, float o_in, float qmax_in,
             const float* w_fq, const float* bias, float alpha, float eps, float s_out, float o_out, float qmax_out,
             uint8_t* codes, int32_t* rowsum, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, x && w_fq && codes && rows > 0 && H > 0, "null pointer or empty input");
  MQ_REQUIRE(c, H % 4 == 0, "hidden size must be a multiple of 4");
  MQ_REQUIRE(c, qmax_out <= 255.f, "norm output codes are 8 bit");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned grid = (unsigned)((rows + 3) / 4);
  if (is_layernorm) qnorm_kernel<true><<<grid, 128, 0, st>>>(x, rows, H, s_in, o_in, qmax_in, w_fq, bias, alpha, eps, s_out, o_out, qmax_out, codes, rowsum);
  else qnorm_kernel<false><<<grid, 128, 0, st>>>(x, rows, H, s_in, o_in, qmax_in, w_fq, bias, alpha, eps, s_out, o_out, qmax_out, codes, rowsum);
  return check_launch(c, "mq_qnorm");
}

int mq_qrope(void* ctx, const uint8_t* qkv, int ldq, int B, int T, int nh, int nkv, int hd, int rot, const float* in_qparams,
             const float* out_qparams, const float* cos, const float* sin, uint8_t* q, uint8_t* k, uint8_t* vt, int32_t* rsq,
             int32_t* rsk, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, qkv && in_qparams && out_qparams && cos && sin && q && k && vt && rsq && rsk, "null pointer");
  MQ_REQUIRE(c, B > 0 && T > 0 && nh > 0 && nkv > 0 && hd > 0 && rot >= 0 && rot <= hd && rot % 2 == 0, "bad shape");
  RopeArgs a;
  a.qkv = qkv; a.ldq = ldq; a.B = B; a.T = T; a.nh = nh; a.nkv = nkv; a.hd = hd; a.rot = rot;
  a.sq_in = in_qparams[0]; a.oq_in = in_qparams[1]; a.sk_in = in_qparams[2]; a.ok_in = in_qparams[3];
  a.sv_in = in_qparams[4]; a.ov_in = in_qparams[5];
  a.sq = out_qparams[0]; a.oq = out_qparams[1]; a.sk = out_qparams[2]; a.ok = out_qparams[3]; a.sv = out_qparams[4];
  a.ov = out_qparams[5];
  a.cos = cos; a.sin = sin; a.q = q; a.k = k; a.vt = vt; a.rsq = rsq; a.rsk = rsk;
  const int M = B * T;
  size_t smem = size_t(32) * (nkv * hd + 4);
  MQ_REQUIRE(c, smem <= 48 * 1024, "nkv*hd too large for the V transpose tile");
  qrope_kernel<<<(M + 31) / 32, 256, smem, (cudaStream_t)stream>>>(a);
  return check_launch(c, "mq_qrope");
}

int mq_qattn(void* ctx, const uint8_t* q, const uint8_t* k, const uint8_t* vt, const int32_t* rsq, const int32_t* rsk, int B,
             int T, int nh, int nkv, int hd, const float* qparams, const uint32_t* lut, uint8_t* out, int32_t* rowsum_out,
             void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, q && k && vt && rsq && rsk && qparams && lut && out, "null pointer");
  MQ_REQUIRE(c, B > 0 && T > 0 && nh > 0 && nkv > 0 && nh % nkv == 0, "bad shape");
  MQ_REQUIRE(c, hd == 32 || hd == 64 || hd == 128 || hd == 256, "head_dim must be 32, 64, 128 or 256");
  AttnArgs a;
  a.q = q; a.k = k; a.vt = vt; a.rsq = rsq; a.rsk = rsk; a.B = B; a.T = T; a.nh = nh; a.nkv = nkv; a.hd = hd;
  a.oq = qparams[0]; a.ok = qparams[1]; a.ov = qparams[2]; a.sqk = qparams[3]; a.s_s = qparams[4]; a.o_s = qparams[5];
  a.qmax_s = qparams[6]; a.s_p = qparams[7]; a.qmax_p = qparams[8]; a.spv = qparams[9]; a.s_out = qparams[10];
  a.o_out = qparams[11];
  a.lut = lut; a.out = out; a.rowsum_out = rowsum_out;
  dim3 grid((T + 63) / 64 * (hd == 256 ? 2 : 1), nh, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (hd == 32) qattn_kernel<32, 32><<<grid, 128, 0, st>>>(a);
  else if (hd == 64) qattn_kernel<64, 64><<<grid, 128, 0, st>>>(a);
  else if (hd == 128) qattn_kernel<128, 128><<<grid, 128, 0, st>>>(a);
  else qattn_kernel<256, 128><<<grid, 128, 0, st>>>(a);
  return check_launch(c, "mq_qattn");
}

}  // extern "C"
