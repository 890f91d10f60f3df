// Element-wise pieces of a decoder block in calibration mode, each ONE kernel forward and ONE backward instead of the chain of
// ATen launches the module graph produces (the quantizers' scale / offset gradients come out of the same backward pass,
// folded deterministically; arithmetic as in fq_math.cuh):
//   * gated SiLU MLP core (hm:1042-1062 with QSiLU qm:691-753 and w2.input_quantizer):
//       out = fq_w( fq_o( a * fq_s(sigmoid(a)) ) * b ),   a = fq_a(w1 x + b1), b = fq_b(w3 x + b3)
//     fq_a / fq_b = the output quantizers of w1 / w3; a and b may be the two halves of one [rows, 2*I] GEMM result (row stride)
//   * QRMSNorm in its L2-norm form (qm:515-531 over hm:187-195, F.normalize):
//       out = fq_out( w * (alpha * xq / max(||xq||_2, eps)) + bias ),   xq = fq_in(x)
//     one CTA per row (row statistics through shared memory), dL/dw and dL/dbias accumulated per thread over the rows of a
//     CTA and folded over CTAs in fixed order by a second tiny kernel.
//   * Q/K/V post-processing between the (concatenated) projection GEMM and the two attention matmuls (hm:470-512 with the
//     QLinear output quantizers, qm:356-358, and the QMatMul input quantizers, qm:455-458):
//       q = fq_qk.in( rope( fq_qproj.out(yq) ) ),  k = fq_qk.in2( rope( fq_kproj.out(yk) ) ),  v = fq_pv.in2( fq_vproj.out(yv) )
//     read from the [B*T, (nh + 2 nkv) hd] GEMM result, written head-major ([B, nh, T, hd] / [B, nkv, T, hd]): the transposes,
//     rotate_half / cat and six fake-quant passes of the module graph in one pass each way.
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
  const float *a, *b; float* out; int64_t rows; int cols; int64_t ld;   // a, b: [rows, cols] with row stride ld; out: [rows, cols]
  const float *sc[5], *of[5]; float qmin[5], qmax[5];  // fq_a (w1.output), fq_b (w3.output), fq_s, fq_o (QSiLU), fq_w (w2.input)
  const float* g; float *da, *db; int64_t ldd;         // backward: da, db with row stride ldd
  double* partial; unsigned* ticket; float* gout;      // gout[10] = (d/dscale, d/doffset) of the five quantizers
};

// ATen: sigmoid(x) = 1 / (1 + exp(-x)) in fp32
__device__ __forceinline__ float sigmoid_rn(float x) { return __frcp_rn(fadd(1.f, expf(-x))); }

__global__ void __launch_bounds__(256) silu_gate_fwd_kernel(const GateArgs p) {
  FqP q[5];
  bool five = false;
#pragma unroll
  for (int i = 0; i < 5; ++i) { q[i] = load_fqp(p.sc[i], p.of[i], p.qmin[i], p.qmax[i]); five |= q[i].five; }
  const int c4 = p.cols >> 2;
  const int64_t n4 = p.rows * c4, nthr = int64_t(gridDim.x) * blockDim.x;
  auto body = [&](auto five_tag) {
    constexpr bool FIVE = decltype(five_tag)::value;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += nthr) {
      const int64_t r = i / c4;
      const int c = int(i - r * c4) << 2;
      const float4 a4 = ldg4_stream(p.a + r * p.ld + c), b4 = ldg4_stream(p.b + r * p.ld + c);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = fq_apply<FIVE>(av[e], q[0]), b = fq_apply<FIVE>(bv[e], q[1]);
        const float y = fq_apply<FIVE>(sigmoid_rn(a), q[2]);
        const float h = fq_apply<FIVE>(fmul(a, y), q[3]);
        o[e] = fq_apply<FIVE>(fmul(h, b), q[4]);
      }
      *reinterpret_cast<float4*>(p.out + r * p.cols + c) = make_float4(o[0], o[1], o[2], o[3]);
    }
  };
  dispatch_five(five, body);
}

__global__ void __launch_bounds__(256) silu_gate_bwd_kernel(const GateArgs p) {
  __shared__ float red[32];
  __shared__ bool s_last;
  FqP q[5];
  bool five = false;
#pragma unroll
  for (int i = 0; i < 5; ++i) { q[i] = load_fqp(p.sc[i], p.of[i], p.qmin[i], p.qmax[i]); five |= q[i].five; }
  const int c4 = p.cols >> 2;
  const int64_t n4 = p.rows * c4, nthr = int64_t(gridDim.x) * blockDim.x;
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.f;
  auto body = [&](auto five_tag) {
    constexpr bool FIVE = decltype(five_tag)::value;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += nthr) {
      const int64_t r = i / c4;
      const int c = int(i - r * c4) << 2;
      const float4 a4 = ldg4_stream(p.a + r * p.ld + c), b4 = ldg4_stream(p.b + r * p.ld + c), g4 = ldg4_stream(p.g + r * p.cols + c);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w};
      float da[4], db[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = fq_apply<FIVE>(av[e], q[0]), b = fq_apply<FIVE>(bv[e], q[1]);
        const float sg = sigmoid_rn(a);
        const float y = fq_apply<FIVE>(sg, q[2]);
        const float hh = fmul(a, y);
        const float h = fq_apply<FIVE>(hh, q[3]);
        const FqGrad rw = fq_grad<FIVE>(fmul(h, b), gv[e], q[4]);           // through w2.input_quantizer
        const FqGrad rb = fq_grad<FIVE>(bv[e], fmul(rw.gx, h), q[1]);        // gate operand, through w3.output_quantizer
        const FqGrad ro = fq_grad<FIVE>(hh, fmul(rw.gx, b), q[3]);          // through QSiLU.output_quantizer
        const FqGrad rs = fq_grad<FIVE>(sg, fmul(ro.gx, a), q[2]);          // through QSiLU.input2_quantizer
        // a feeds the product directly and through the sigmoid (ATen sigmoid_backward: g * (1 - y) * y)
        const float ga = fadd(fmul(ro.gx, y), fmul(fmul(rs.gx, fsub(1.f, sg)), sg));
        const FqGrad ra = fq_grad<FIVE>(av[e], ga, q[0]);                    // through w1.output_quantizer
        da[e] = ra.gx; db[e] = rb.gx;
        acc[0] += ra.gs; acc[1] += ra.go; acc[2] += rb.gs; acc[3] += rb.go; acc[4] += rs.gs; acc[5] += rs.go;
        acc[6] += ro.gs; acc[7] += ro.go; acc[8] += rw.gs; acc[9] += rw.go;
      }
      *reinterpret_cast<float4*>(p.da + r * p.ldd + c) = make_float4(da[0], da[1], da[2], da[3]);
      *reinterpret_cast<float4*>(p.db + r * p.ldd + c) = make_float4(db[0], db[1], db[2], db[3]);
    }
  };
  dispatch_five(five, body);
  if (p.gout) grid_fold<10>(acc, p.partial, p.ticket, p.gout, red, &s_last);
}

struct NormCArgs {
  const float *x, *w, *b; float *out, *nrm; int64_t rows; int H; float alpha, eps;
  const float *s_i, *o_i; float qmin_i, qmax_i;      // input quantizer
  const float *s_o, *o_o; float qmin_o, qmax_o;      // output quantizer
  const float* g; float *dx, *pdw, *pdb;             // backward: pdw / pdb = per-CTA column partials [gridDim.x][H]
  double* partial; unsigned* ticket; float* gout;    // gout[4] = d/d(scale, offset) of the input, output quantizer
};

// thread t of the CTA owns columns (k*256 + t)*4 .. +3, k < NVT
template <int NVT>
__global__ void __launch_bounds__(256) rmsnorm_l2_fwd_kernel(const NormCArgs p) {
  __shared__ float red[32];
  const FqP qi = load_fqp(p.s_i, p.o_i, p.qmin_i, p.qmax_i), qo = load_fqp(p.s_o, p.o_o, p.qmin_o, p.qmax_o);
  const int H = p.H;
  auto body = [&](auto five_tag) {
    constexpr bool FIVE = decltype(five_tag)::value;
    for (int64_t row = blockIdx.x; row < p.rows; row += gridDim.x) {
      const float* xr = p.x + row * H;
      float xq[NVT][4];
      float ss = 0.f;
#pragma unroll
      for (int k = 0; k < NVT; ++k) {
        const int c = (k * 256 + threadIdx.x) * 4;
        xq[k][0] = xq[k][1] = xq[k][2] = xq[k][3] = 0.f;
        if (c < H) {
          const float4 v = ldg4_stream(xr + c);
          xq[k][0] = fq_apply<FIVE>(v.x, qi); xq[k][1] = fq_apply<FIVE>(v.y, qi);
          xq[k][2] = fq_apply<FIVE>(v.z, qi); xq[k][3] = fq_apply<FIVE>(v.w, qi);
#pragma unroll
          for (int e = 0; e < 4; ++e) ss = fmaf(xq[k][e], xq[k][e], ss);
        }
      }
      ss = block_reduce(ss, OpSum(), red);
      const float raw = __fsqrt_rn(ss), nrm = fmaxf(raw, p.eps), rn = __frcp_rn(nrm);
      if (threadIdx.x == 0) p.nrm[row] = raw;
#pragma unroll
      for (int k = 0; k < NVT; ++k) {
        const int c = (k * 256 + threadIdx.x) * 4;
        if (c >= H) continue;
        const float4 w4 = ldg4(p.w + c);
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.b) b4 = ldg4(p.b + c);
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float y = fmul(wv[e], fmul(p.alpha, div_rn<true>(xq[k][e], nrm, rn)));
          if (p.b) y = fadd(y, bv[e]);
          o[e] = fq_apply<FIVE>(y, qo);
        }
        *reinterpret_cast<float4*>(p.out + row * H + c) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  };
  dispatch_five(qi.five | qo.five, body);
}

template <int NVT>
__global__ void __launch_bounds__(256) rmsnorm_l2_bwd_kernel(const NormCArgs p) {
  __shared__ float red[32];
  __shared__ bool s_last;
  const FqP qi = load_fqp(p.s_i, p.o_i, p.qmin_i, p.qmax_i), qo = load_fqp(p.s_o, p.o_o, p.qmin_o, p.qmax_o);
  const int H = p.H;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float dw[NVT][4], db[NVT][4];
#pragma unroll
  for (int k = 0; k < NVT; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) dw[k][e] = db[k][e] = 0.f;
  auto body = [&](auto five_tag) {
    constexpr bool FIVE = decltype(five_tag)::value;
    for (int64_t row = blockIdx.x; row < p.rows; row += gridDim.x) {
      const float* xr = p.x + row * H;
      const float* gr = p.g + row * H;
      const float raw = __ldg(p.nrm + row), nrm = fmaxf(raw, p.eps), rn = __frcp_rn(nrm);
      const bool clamped = raw < p.eps;                       // clamp_min passes no gradient to the norm then
      float nv[NVT][4], dn[NVT][4];
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < NVT; ++k) {
        const int c = (k * 256 + threadIdx.x) * 4;
        nv[k][0] = nv[k][1] = nv[k][2] = nv[k][3] = 0.f;
        dn[k][0] = dn[k][1] = dn[k][2] = dn[k][3] = 0.f;
        if (c < H) {
          const float4 v = ldg4(xr + c), g4 = ldg4_stream(gr + c), w4 = ldg4(p.w + c);
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.b) b4 = ldg4(p.b + c);
          const float xv[4] = {v.x, v.y, v.z, v.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w},
                      bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float n = div_rn<true>(fq_apply<FIVE>(xv[e], qi), nrm, rn);
            const float t = fmul(p.alpha, n);
            float y = fmul(wv[e], t);
            if (p.b) y = fadd(y, bv[e]);
            const FqGrad ro = fq_grad<FIVE>(y, gv[e], qo);
            acc[2] += ro.gs; acc[3] += ro.go;
            dw[k][e] = fmaf(ro.gx, t, dw[k][e]);
            db[k][e] += ro.gx;
            nv[k][e] = n;
            dn[k][e] = fmul(p.alpha, fmul(ro.gx, wv[e]));
            dot = fmaf(dn[k][e], n, dot);
          }
        }
      }
      dot = block_reduce(dot, OpSum(), red);
      if (clamped) dot = 0.f;
#pragma unroll
      for (int k = 0; k < NVT; ++k) {
        const int c = (k * 256 + threadIdx.x) * 4;
        if (c >= H) continue;
        const float4 v = ldg4(xr + c);
        const float xv[4] = {v.x, v.y, v.z, v.w};
        float d[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float dxq = fmul(fmaf(-nv[k][e], dot, dn[k][e]), rn);
          const FqGrad ri = fq_grad<FIVE>(xv[e], dxq, qi);
          d[e] = ri.gx; acc[0] += ri.gs; acc[1] += ri.go;
        }
        *reinterpret_cast<float4*>(p.dx + row * H + c) = make_float4(d[0], d[1], d[2], d[3]);
      }
    }
  };
  dispatch_five(qi.five | qo.five, body);
#pragma unroll
  for (int k = 0; k < NVT; ++k) {
    const int c = (k * 256 + threadIdx.x) * 4;
    if (c >= H) continue;
    *reinterpret_cast<float4*>(p.pdw + int64_t(blockIdx.x) * H + c) = make_float4(dw[k][0], dw[k][1], dw[k][2], dw[k][3]);
    if (p.pdb) *reinterpret_cast<float4*>(p.pdb + int64_t(blockIdx.x) * H + c) = make_float4(db[k][0], db[k][1], db[k][2], db[k][3]);
  }
  if (p.gout) grid_fold<4>(acc, p.partial, p.ticket, p.gout, red, &s_last);
}

// out[c] = sum over blocks of part[blk][c]: 32 columns x 8 block slices per CTA (slice y takes blocks y, y + 8, ...), the slices
// meet in shared memory in slice order -- a fixed order, and 8x shorter dependent chains than one thread per column
__global__ void __launch_bounds__(256) fold_cols_kernel(const float* __restrict__ part, int nblk, int H, float* __restrict__ out) {
  __shared__ float s_p[8][33];
  const int cx = threadIdx.x & 31, sy = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < H)
    for (int b = sy; b < nblk; b += 8) s += part[int64_t(b) * H + c];
  s_p[sy][cx] = s;
  __syncthreads();
  if (sy == 0 && c < H) {
    float t = s_p[0][cx];
#pragma unroll
    for (int y = 1; y < 8; ++y) t += s_p[y][cx];
    out[c] = t;
  }
}

static int norm_nvt(int H) {
  if (H % 4 != 0 || H < 4 || H > 8192) return 0;
  const int chunks = (H / 4 + 255) / 256;
  return chunks <= 1 ? 1 : chunks <= 2 ? 2 : chunks <= 4 ? 4 : 8;
}


struct QkvArgs {
  const float* y; int64_t rows; int T, nh, nkv, hd, rot;   // y: [rows = B*T, (nh + 2 nkv) * hd]
  const float *cos, *sin; int cs_batched;                  // [B or 1, T, rot]
  float *q, *k, *v;                                        // [B, nh, T, hd], [B, nkv, T, hd] x 2  (backward: the incoming gradients)
  const float *sc[6], *of[6]; float qmin[6], qmax[6];      // q_proj.out, k_proj.out, v_proj.out, qk.input, qk.input2, pv.input2
  float* dy;                                               // backward: [rows, (nh + 2 nkv) * hd]
  double* partial; unsigned* ticket; float* gout;          // gout[12] = (d/dscale, d/doffset) of the six quantizers in that order
};

// One thread per 8-element unit of a head: a rotary unit is 4 dims d_lo.. and their partners d_lo + rot/2.., a pass-through
// unit 8 consecutive dims (beyond rot, or anywhere in a V head).  Thread (blockIdx.x, threadIdx.x) keeps its unit column and
// walks rows blockIdx.y, + gridDim.y, ...: its pair of quantizers is fixed.
struct QkvUnit {
  int seg, head, hl, nheads, d_lo, d_hi; bool rotary, valid;
};
__device__ __forceinline__ QkvUnit qkv_unit(const QkvArgs& p) {
  QkvUnit n;
  const int upr = p.hd >> 3;
  const int u = blockIdx.x * 128 + threadIdx.x;
  n.valid = u < (p.nh + 2 * p.nkv) * upr;
  const int uu = n.valid ? u : 0;
  n.head = uu / upr;
  const int w = uu - n.head * upr;
  n.seg = n.head < p.nh ? 0 : (n.head < p.nh + p.nkv ? 1 : 2);
  n.hl = n.seg == 0 ? n.head : (n.seg == 1 ? n.head - p.nh : n.head - p.nh - p.nkv);
  n.nheads = n.seg == 0 ? p.nh : p.nkv;
  n.rotary = n.seg < 2 && w < (p.rot >> 3);
  n.d_lo = n.rotary ? w * 4 : (n.seg < 2 ? p.rot + (w - (p.rot >> 3)) * 8 : w * 8);
  n.d_hi = n.rotary ? n.d_lo + (p.rot >> 1) : n.d_lo + 4;
  return n;
}
struct QkvTrig { float clo[4], chi[4], slo[4], shi[4]; };
__device__ __forceinline__ void qkv_trig(const QkvArgs& p, const QkvUnit& n, int64_t b, int t, QkvTrig& g) {
  const int64_t cb = ((p.cs_batched ? b : 0) * p.T + t) * p.rot;
  const float4 c0 = ldg4(p.cos + cb + n.d_lo), c1 = ldg4(p.cos + cb + n.d_hi), s0 = ldg4(p.sin + cb + n.d_lo), s1 = ldg4(p.sin + cb + n.d_hi);
  g.clo[0] = c0.x; g.clo[1] = c0.y; g.clo[2] = c0.z; g.clo[3] = c0.w; g.chi[0] = c1.x; g.chi[1] = c1.y; g.chi[2] = c1.z; g.chi[3] = c1.w;
  g.slo[0] = s0.x; g.slo[1] = s0.y; g.slo[2] = s0.z; g.slo[3] = s0.w; g.shi[0] = s1.x; g.shi[1] = s1.y; g.shi[2] = s1.z; g.shi[3] = s1.w;
}
__device__ __forceinline__ bool qkv_any_five(const QkvArgs& p) {
  bool five = false;
#pragma unroll
  for (int i = 0; i < 6; ++i) five |= (p.sc[i] != nullptr) && mantissa_all_ones(__ldg(p.sc[i]));
  return five;
}

__global__ void __launch_bounds__(128) qkv_rope_fwd_kernel(const QkvArgs p) {
  const QkvUnit n = qkv_unit(p);
  if (!n.valid) return;
  const FqP ql = load_fqp(p.sc[n.seg], p.of[n.seg], p.qmin[n.seg], p.qmax[n.seg]);
  const FqP qm = load_fqp(p.sc[3 + n.seg], p.of[3 + n.seg], p.qmin[3 + n.seg], p.qmax[3 + n.seg]);
  float* dst = n.seg == 0 ? p.q : (n.seg == 1 ? p.k : p.v);
  const int C = (p.nh + 2 * p.nkv) * p.hd;
  auto body = [&](auto five_tag) {
    constexpr bool FIVE = decltype(five_tag)::value;
    for (int64_t r = blockIdx.y; r < p.rows; r += gridDim.y) {
      const int64_t b = r / p.T;
      const int t = int(r - b * p.T);
      const float* yr = p.y + r * C + n.head * p.hd;
      const float4 lo4 = ldg4_stream(yr + n.d_lo), hi4 = ldg4_stream(yr + n.d_hi);
      const float xlo[4] = {lo4.x, lo4.y, lo4.z, lo4.w}, xhi[4] = {hi4.x, hi4.y, hi4.z, hi4.w};
      QkvTrig g;
      if (n.rotary) qkv_trig(p, n, b, t, g);
      float olo[4], ohi[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = fq_apply<FIVE>(xlo[e], ql), c = fq_apply<FIVE>(xhi[e], ql);
        float rl = a, rh = c;
        if (n.rotary) {                                        // (x * cos) + (rotate_half(x) * sin), hm:338-367
          rl = fadd(fmul(a, g.clo[e]), fmul(-c, g.slo[e]));
          rh = fadd(fmul(c, g.chi[e]), fmul(a, g.shi[e]));
        }
        olo[e] = fq_apply<FIVE>(rl, qm); ohi[e] = fq_apply<FIVE>(rh, qm);
      }
      float* dr = dst + ((b * n.nheads + n.hl) * p.T + t) * p.hd;
      *reinterpret_cast<float4*>(dr + n.d_lo) = make_float4(olo[0], olo[1], olo[2], olo[3]);
      *reinterpret_cast<float4*>(dr + n.d_hi) = make_float4(ohi[0], ohi[1], ohi[2], ohi[3]);
    }
  };
  dispatch_five(qkv_any_five(p), body);
}

__global__ void __launch_bounds__(128) qkv_rope_bwd_kernel(const QkvArgs p) {
  __shared__ float red[32];
  __shared__ bool s_last;
  const QkvUnit n = qkv_unit(p);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};                        // lin scale, lin offset, mm scale, mm offset of this thread's segment
  if (n.valid) {
    const FqP ql = load_fqp(p.sc[n.seg], p.of[n.seg], p.qmin[n.seg], p.qmax[n.seg]);
    const FqP qm = load_fqp(p.sc[3 + n.seg], p.of[3 + n.seg], p.qmin[3 + n.seg], p.qmax[3 + n.seg]);
    const float* gsrc = n.seg == 0 ? p.q : (n.seg == 1 ? p.k : p.v);
    const int C = (p.nh + 2 * p.nkv) * p.hd;
    auto body = [&](auto five_tag) {
      constexpr bool FIVE = decltype(five_tag)::value;
      for (int64_t r = blockIdx.y; r < p.rows; r += gridDim.y) {
        const int64_t b = r / p.T;
        const int t = int(r - b * p.T);
        const float* yr = p.y + r * C + n.head * p.hd;
        const float* gr = gsrc + ((b * n.nheads + n.hl) * p.T + t) * p.hd;
        const float4 lo4 = ldg4_stream(yr + n.d_lo), hi4 = ldg4_stream(yr + n.d_hi);
        const float4 gl4 = ldg4_stream(gr + n.d_lo), gh4 = ldg4_stream(gr + n.d_hi);
        const float xlo[4] = {lo4.x, lo4.y, lo4.z, lo4.w}, xhi[4] = {hi4.x, hi4.y, hi4.z, hi4.w};
        const float glo[4] = {gl4.x, gl4.y, gl4.z, gl4.w}, ghi[4] = {gh4.x, gh4.y, gh4.z, gh4.w};
        QkvTrig g;
        if (n.rotary) qkv_trig(p, n, b, t, g);
        float dlo[4], dhi[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = fq_apply<FIVE>(xlo[e], ql), c = fq_apply<FIVE>(xhi[e], ql);
          float rl = a, rh = c;
          if (n.rotary) {
            rl = fadd(fmul(a, g.clo[e]), fmul(-c, g.slo[e]));
            rh = fadd(fmul(c, g.chi[e]), fmul(a, g.shi[e]));
          }
          const FqGrad ml = fq_grad<FIVE>(rl, glo[e], qm), mh = fq_grad<FIVE>(rh, ghi[e], qm);
          acc[2] += ml.gs + mh.gs; acc[3] += ml.go + mh.go;
          float ga = ml.gx, gc = mh.gx;
          if (n.rotary) {                                      // d/da = g_lo*cos_lo + g_hi*sin_hi ; d/dc = g_hi*cos_hi - g_lo*sin_lo
            ga = fadd(fmul(ml.gx, g.clo[e]), fmul(mh.gx, g.shi[e]));
            gc = fadd(fmul(mh.gx, g.chi[e]), -fmul(ml.gx, g.slo[e]));
          }
          const FqGrad ll = fq_grad<FIVE>(xlo[e], ga, ql), lh = fq_grad<FIVE>(xhi[e], gc, ql);
          acc[0] += ll.gs + lh.gs; acc[1] += ll.go + lh.go;
          dlo[e] = ll.gx; dhi[e] = lh.gx;
        }
        float* dr = p.dy + r * C + n.head * p.hd;
        *reinterpret_cast<float4*>(dr + n.d_lo) = make_float4(dlo[0], dlo[1], dlo[2], dlo[3]);
        *reinterpret_cast<float4*>(dr + n.d_hi) = make_float4(dhi[0], dhi[1], dhi[2], dhi[3]);
      }
    };
    dispatch_five(qkv_any_five(p), body);
  }
  if (!p.gout) return;
  float a12[12];
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const bool mine = n.valid && n.seg == s;
    a12[2 * s] = mine ? acc[0] : 0.f; a12[2 * s + 1] = mine ? acc[1] : 0.f;
    a12[6 + 2 * s] = mine ? acc[2] : 0.f; a12[6 + 2 * s + 1] = mine ? acc[3] : 0.f;
  }
  grid_fold<12>(a12, p.partial, p.ticket, p.gout, red, &s_last, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y);
}

}  // namespace mq

using namespace mq;

extern "C" {

static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

static void gate_fill(GateArgs& p, const float* const* scales, const float* const* offsets, const float* qmins, const float* qmaxs) {
  for (int i = 0; i < 5; ++i) { p.sc[i] = scales[i]; p.of[i] = offsets[i]; p.qmin[i] = qmins[i]; p.qmax[i] = qmaxs[i]; }
}

int mq_silu_gate_fwd(void* ctx, const float* a, const float* b, int64_t ld, float* out, int64_t rows, int cols,
                     const float* const* scales, const float* const* offsets, const float* qmins, const float* qmaxs, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, a && b && out && rows >= 0 && cols > 0 && cols % 4 == 0 && ld >= cols && ld % 4 == 0 && scales && offsets && qmins && qmaxs,
             "null pointer or cols / ld not a multiple of 4");
  MQ_REQUIRE(c, al16(a) && al16(b) && al16(out), "a / b / out must be 16-byte aligned");
  for (int i = 0; i < 5; ++i) MQ_REQUIRE(c, (scales[i] == nullptr) == (offsets[i] == nullptr), "scale and offset come in pairs");
  if (rows == 0) return MQ_NO_ERROR;
  GateArgs p{};
  p.a = a; p.b = b; p.ld = ld; p.out = out; p.rows = rows; p.cols = cols;
  gate_fill(p, scales, offsets, qmins, qmaxs);
  silu_gate_fwd_kernel<<<grid_for_elems(c, rows * (cols / 4), 8), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch(c, "mq_silu_gate_fwd");
}

int mq_silu_gate_bwd(void* ctx, const float* a, const float* b, int64_t ld, const float* g, float* da, float* db, int64_t ldd,
                     int64_t rows, int cols, const float* const* scales, const float* const* offsets, const float* qmins,
                     const float* qmaxs, float* gparams, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, a && b && g && da && db && rows >= 0 && cols > 0 && cols % 4 == 0 && ld >= cols && ld % 4 == 0 && ldd >= cols && ldd % 4 == 0 &&
                 scales && offsets && qmins && qmaxs, "null pointer or cols / ld / ldd not a multiple of 4");
  MQ_REQUIRE(c, al16(a) && al16(b) && al16(g) && al16(da) && al16(db), "tensors must be 16-byte aligned");
  for (int i = 0; i < 5; ++i) MQ_REQUIRE(c, (scales[i] == nullptr) == (offsets[i] == nullptr), "scale and offset come in pairs");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    if (gparams) cudaMemsetAsync(gparams, 0, 10 * sizeof(float), st);
    return MQ_NO_ERROR;
  }
  GateArgs p{};
  p.a = a; p.b = b; p.ld = ld; p.g = g; p.da = da; p.db = db; p.ldd = ldd; p.rows = rows; p.cols = cols; p.gout = gparams;
  gate_fill(p, scales, offsets, qmins, qmaxs);
  if (gparams) {
    void* wsp = stream_ws(c, st);
    if (!wsp) return MQ_FAILED_ALLOCATION;
    p.partial = reinterpret_cast<double*>(wsp);
    p.ticket = reinterpret_cast<unsigned*>(static_cast<char*>(wsp) + c->ws_bytes - 64);
  }
  silu_gate_bwd_kernel<<<grid_for_elems(c, rows * (cols / 4), 4), 256, 0, st>>>(p);
  return check_launch(c, "mq_silu_gate_bwd");
}

int mq_rmsnorm_l2_supported(int H) { return norm_nvt(H) != 0; }

static void norm_fill(NormCArgs& p, const float* x, const float* w, const float* b, int64_t rows, int H, float alpha, float eps,
                      const float* const* scales, const float* const* offsets, const float* qmins, const float* qmaxs) {
  p.x = x; p.w = w; p.b = b; p.rows = rows; p.H = H; p.alpha = alpha; p.eps = eps;
  p.s_i = scales[0]; p.o_i = offsets[0]; p.qmin_i = qmins[0]; p.qmax_i = qmaxs[0];
  p.s_o = scales[1]; p.o_o = offsets[1]; p.qmin_o = qmins[1]; p.qmax_o = qmaxs[1];
}

int mq_rmsnorm_l2_fwd(void* ctx, const float* x, const float* w, const float* bias, float* out, float* nrm, int64_t rows, int H,
                      float alpha, float eps, const float* const* scales, const float* const* offsets, const float* qmins,
                      const float* qmaxs, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, x && w && out && nrm && rows >= 0 && scales && offsets && qmins && qmaxs, "null pointer or negative size");
  const int nvt = norm_nvt(H);
  MQ_REQUIRE(c, nvt != 0, "H must be a multiple of 4 and <= 8192");
  MQ_REQUIRE(c, al16(x) && al16(w) && al16(out) && (!bias || al16(bias)), "tensors must be 16-byte aligned");
  for (int i = 0; i < 2; ++i) MQ_REQUIRE(c, (scales[i] == nullptr) == (offsets[i] == nullptr), "scale and offset come in pairs");
  if (rows == 0) return MQ_NO_ERROR;
  NormCArgs p{};
  norm_fill(p, x, w, bias, rows, H, alpha, eps, scales, offsets, qmins, qmaxs);
  p.out = out; p.nrm = nrm;
  const int64_t cap = int64_t(c->sm_count) * 8;
  const unsigned grid = (unsigned)(rows < cap ? rows : cap);
  cudaStream_t st = (cudaStream_t)stream;
  switch (nvt) {
    case 1: rmsnorm_l2_fwd_kernel<1><<<grid, 256, 0, st>>>(p); break;
    case 2: rmsnorm_l2_fwd_kernel<2><<<grid, 256, 0, st>>>(p); break;
    case 4: rmsnorm_l2_fwd_kernel<4><<<grid, 256, 0, st>>>(p); break;
    default: rmsnorm_l2_fwd_kernel<8><<<grid, 256, 0, st>>>(p); break;
  }
  return check_launch(c, "mq_rmsnorm_l2_fwd");
}

int mq_rmsnorm_l2_bwd(void* ctx, const float* x, const float* w, const float* bias, const float* nrm, const float* g, float* dx,
                      float* dw, float* dbias, int64_t rows, int H, float alpha, float eps, const float* const* scales,
                      const float* const* offsets, const float* qmins, const float* qmaxs, float* gparams, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, x && w && nrm && g && dx && dw && rows > 0 && scales && offsets && qmins && qmaxs, "null pointer or empty input");
  MQ_REQUIRE(c, !dbias || bias, "dbias without a bias");
  const int nvt = norm_nvt(H);
  MQ_REQUIRE(c, nvt != 0, "H must be a multiple of 4 and <= 8192");
  MQ_REQUIRE(c, al16(x) && al16(w) && al16(g) && al16(dx) && (!bias || al16(bias)), "tensors must be 16-byte aligned");
  for (int i = 0; i < 2; ++i) MQ_REQUIRE(c, (scales[i] == nullptr) == (offsets[i] == nullptr), "scale and offset come in pairs");
  cudaStream_t st = (cudaStream_t)stream;
  NormCArgs p{};
  norm_fill(p, x, w, bias, rows, H, alpha, eps, scales, offsets, qmins, qmaxs);
  p.nrm = const_cast<float*>(nrm); p.g = g; p.dx = dx; p.gout = gparams;
  const int64_t cap = int64_t(c->sm_count) * 2;
  const unsigned grid = (unsigned)(rows < cap ? rows : cap);
  // workspace of this stream: [grid_fold partials (64 KB)] [pdw: grid*H floats] [pdb: grid*H floats] ... [tickets: last 64 B]
  char* wsp = static_cast<char*>(stream_ws(c, st));
  if (!wsp) return MQ_FAILED_ALLOCATION;
  const size_t part_bytes = size_t(grid) * H * sizeof(float);
  MQ_REQUIRE(c, 65536 + 2 * part_bytes + 64 <= c->ws_bytes, "workspace too small");
  p.partial = reinterpret_cast<double*>(wsp);
  p.ticket = reinterpret_cast<unsigned*>(wsp + c->ws_bytes - 64);
  p.pdw = reinterpret_cast<float*>(wsp + 65536);
  p.pdb = dbias ? reinterpret_cast<float*>(wsp + 65536 + part_bytes) : nullptr;
  switch (nvt) {
    case 1: rmsnorm_l2_bwd_kernel<1><<<grid, 256, 0, st>>>(p); break;
    case 2: rmsnorm_l2_bwd_kernel<2><<<grid, 256, 0, st>>>(p); break;
    case 4: rmsnorm_l2_bwd_kernel<4><<<grid, 256, 0, st>>>(p); break;
    default: rmsnorm_l2_bwd_kernel<8><<<grid, 256, 0, st>>>(p); break;
  }
  fold_cols_kernel<<<(H + 31) / 32, 256, 0, st>>>(p.pdw, (int)grid, H, dw);
  if (dbias) fold_cols_kernel<<<(H + 31) / 32, 256, 0, st>>>(p.pdb, (int)grid, H, dbias);
  return check_launch(c, "mq_rmsnorm_l2_bwd");
}

static int qkv_check(Ctx* c, const float* y, int64_t rows, int T, int nh, int nkv, int hd, int rot, const float* cosp, const float* sinp,
                     const float* const* scales, const float* const* offsets) {
  MQ_REQUIRE(c, y && rows >= 0 && T > 0 && rows % T == 0 && nh > 0 && nkv > 0 && scales && offsets, "null pointer or bad shape");
  MQ_REQUIRE(c, hd % 8 == 0 && rot >= 0 && rot <= hd && rot % 8 == 0, "head_dim and the rotary width must be multiples of 8");
  MQ_REQUIRE(c, rot == 0 || (cosp && sinp && al16(cosp) && al16(sinp)), "cos / sin tables missing or misaligned");
  for (int i = 0; i < 6; ++i) MQ_REQUIRE(c, (scales[i] == nullptr) == (offsets[i] == nullptr), "scale and offset come in pairs");
  return MQ_NO_ERROR;
}
static void qkv_fill(QkvArgs& p, const float* y, int64_t rows, int T, int nh, int nkv, int hd, int rot, const float* cosp, const float* sinp,
                     int cs_batched, const float* const* scales, const float* const* offsets, const float* qmins, const float* qmaxs) {
  p.y = y; p.rows = rows; p.T = T; p.nh = nh; p.nkv = nkv; p.hd = hd; p.rot = rot; p.cos = cosp; p.sin = sinp; p.cs_batched = cs_batched;
  for (int i = 0; i < 6; ++i) { p.sc[i] = scales[i]; p.of[i] = offsets[i]; p.qmin[i] = qmins[i]; p.qmax[i] = qmaxs[i]; }
}
static dim3 qkv_grid(Ctx* c, const QkvArgs& p, int waves) {
  const int U = (p.nh + 2 * p.nkv) * (p.hd >> 3);
  const unsigned gx = (unsigned)((U + 127) / 128);
  int64_t gy = (int64_t(c->sm_count) * waves + gx - 1) / gx;
  if (gy > p.rows) gy = p.rows;
  if (gy > 65535) gy = 65535;
  return dim3(gx, (unsigned)(gy < 1 ? 1 : gy), 1);
}

int mq_qkv_rope_fwd(void* ctx, const float* y, int64_t rows, int T, int nh, int nkv, int hd, int rot, const float* cosp, const float* sinp,
                    int cs_batched, float* q, float* k, float* v, const float* const* scales, const float* const* offsets,
                    const float* qmins, const float* qmaxs, void* stream) {
  MQ_CTX(c, ctx);
  if (int rc = qkv_check(c, y, rows, T, nh, nkv, hd, rot, cosp, sinp, scales, offsets)) return rc;
  MQ_REQUIRE(c, q && k && v && qmins && qmaxs && al16(y) && al16(q) && al16(k) && al16(v), "null or misaligned tensor");
  if (rows == 0) return MQ_NO_ERROR;
  QkvArgs p{};
  qkv_fill(p, y, rows, T, nh, nkv, hd, rot, cosp, sinp, cs_batched, scales, offsets, qmins, qmaxs);
  p.q = q; p.k = k; p.v = v;
  qkv_rope_fwd_kernel<<<qkv_grid(c, p, 16), 128, 0, (cudaStream_t)stream>>>(p);
  return check_launch(c, "mq_qkv_rope_fwd");
}

int mq_qkv_rope_bwd(void* ctx, const float* y, int64_t rows, int T, int nh, int nkv, int hd, int rot, const float* cosp, const float* sinp,
                    int cs_batched, const float* dq, const float* dk, const float* dv, float* dy, const float* const* scales,
                    const float* const* offsets, const float* qmins, const float* qmaxs, float* gparams, void* stream) {
  MQ_CTX(c, ctx);
  if (int rc = qkv_check(c, y, rows, T, nh, nkv, hd, rot, cosp, sinp, scales, offsets)) return rc;
  MQ_REQUIRE(c, dq && dk && dv && dy && qmins && qmaxs && al16(y) && al16(dq) && al16(dk) && al16(dv) && al16(dy), "null or misaligned tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    if (gparams) cudaMemsetAsync(gparams, 0, 12 * sizeof(float), st);
    return MQ_NO_ERROR;
  }
  QkvArgs p{};
  qkv_fill(p, y, rows, T, nh, nkv, hd, rot, cosp, sinp, cs_batched, scales, offsets, qmins, qmaxs);
  p.q = const_cast<float*>(dq); p.k = const_cast<float*>(dk); p.v = const_cast<float*>(dv); p.dy = dy; p.gout = gparams;
  if (gparams) {
    void* wsp = stream_ws(c, st);
    if (!wsp) return MQ_FAILED_ALLOCATION;
    p.partial = reinterpret_cast<double*>(wsp);
    p.ticket = reinterpret_cast<unsigned*>(static_cast<char*>(wsp) + c->ws_bytes - 64);
  }
  qkv_rope_bwd_kernel<<<qkv_grid(c, p, 8), 128, 0, st>>>(p);
  return check_launch(c, "mq_qkv_rope_bwd");
}

}  // extern "C"
