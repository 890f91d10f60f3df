// K1 (static fake-quant fwd/bwd), K8 (range statistics) and K2 (LET + LWC weight prep fwd/bwd).
// All kernels are HBM-bound streaming kernels: 128-bit loads, grids sized as multiples of the SM count,
// deterministic two-stage reductions (no float atomics) so that repeated runs give identical bits.
#include "common.cuh"
#include "ctx.h"
#include "fq_math.cuh"
#include <cstdlib>

namespace mq {

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int grid_for(Ctx* c, int64_t work_items, int per_block, int waves = 4) {
  int64_t need = (work_items + per_block - 1) / per_block;
  int64_t cap = int64_t(c->sm_count) * waves;
  if (need < 1) need = 1;
  return int(need < cap ? need : cap);
}

// Threads per row-CTA of the weight-pass kernels.  A row is only 8-22 KB: with 256-thread CTAs an SM holds 8 rows (64 KB in flight)
// and every thread issues one or two loads before the block reduction; narrower CTAs keep up to 32 rows per SM in flight.
// MQ_WPREP_THREADS overrides (A/B measurements).
static unsigned wprep_threads(int64_t cols) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("MQ_WPREP_THREADS"); forced = e ? atoi(e) : 0; }
  if (forced == 32 || forced == 64 || forced == 128 || forced == 256) return (unsigned)forced;
  return cols >= 4096 ? 128u : 64u;
}

// ================================================================================================================
// K1 forward
// ================================================================================================================
template <bool kVec>
__global__ void __launch_bounds__(256) fq_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                      int32_t* __restrict__ codes, int64_t n,
                                                      const float* __restrict__ scale,
                                                      const float* __restrict__ offset, int64_t group, float qmin,
                                                      float qmax) {
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = int64_t(gridDim.x) * blockDim.x;
  float s = 0.f, o = 0.f;
  if (group == 0) { s = __ldg(scale); o = __ldg(offset); }
  if (kVec) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += nthr) {
      if (group) { int64_t g = (i << 2) / group; s = __ldg(scale + g); o = __ldg(offset + g); }
      float4 v = ldg4_stream(x + (i << 2));
      float q0 = quant_code(v.x, s, o, qmin, qmax), q1 = quant_code(v.y, s, o, qmin, qmax);
      float q2 = quant_code(v.z, s, o, qmin, qmax), q3 = quant_code(v.w, s, o, qmin, qmax);
      if (y) {
        float4 r = make_float4(dequant(q0, s, o), dequant(q1, s, o), dequant(q2, s, o), dequant(q3, s, o));
        *reinterpret_cast<float4*>(y + (i << 2)) = r;
      }
      if (codes) *reinterpret_cast<int4*>(codes + (i << 2)) = make_int4((int)q0, (int)q1, (int)q2, (int)q3);
    }
  } else {
    for (int64_t i = tid; i < n; i += nthr) {
      if (group) { int64_t g = i / group; s = __ldg(scale + g); o = __ldg(offset + g); }
      float q = quant_code(__ldg(x + i), s, o, qmin, qmax);
      if (y) y[i] = dequant(q, s, o);
      if (codes) codes[i] = (int)q;
    }
  }
}

// ================================================================================================================
// K1 backward.  Element-wise terms exactly as autograd evaluates them for qm:286-290:
//   t5 = clamp(rne(x/s)+o) - o ; g_t1 = g*s*m ; gx = g_t1 / s ; gs = g*t5 - g_t1*((x/s)/s) ; go = g_t1 - g*s
// ================================================================================================================
template <bool kVec>
__global__ void __launch_bounds__(256) fq_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                      float* __restrict__ gx, int64_t n,
                                                      const float* __restrict__ scale,
                                                      const float* __restrict__ offset, int64_t group, float qmin,
                                                      float qmax, double* __restrict__ partial /*[2*grid] or NULL*/,
                                                      unsigned* __restrict__ ticket, float* __restrict__ gscale,
                                                      float* __restrict__ goffset) {
  __shared__ float red[32];
  __shared__ bool s_last;
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = int64_t(gridDim.x) * blockDim.x;
  float s = 0.f, o = 0.f;
  if (group == 0) { s = __ldg(scale); o = __ldg(offset); }
  float acc_s = 0.f, acc_o = 0.f;
  // per-tensor ranges (every activation quantizer): the same ops with the three IEEE divisions by the launch-uniform scale done
  // through the exact reciprocal-based div_rn (bit-identical, ~3x fewer instructions); per-group ranges keep __fdiv_rn
  const float rs = __frcp_rn(s);
  auto run = [&](auto five_tag) {
    constexpr bool FIVE = decltype(five_tag)::value;
    auto elem = [&](float xv, float gv) { return group ? fq_bwd_elem(xv, gv, s, o, qmin, qmax) : fq_bwd_elem_v<FIVE>(xv, gv, s, rs, o, qmin, qmax); };
    if (kVec) {
      const int64_t n4 = n >> 2;
      for (int64_t i = tid; i < n4; i += nthr) {
        if (group) { int64_t gi = (i << 2) / group; s = __ldg(scale + gi); o = __ldg(offset + gi); }
        float4 xv = ldg4_stream(x + (i << 2)), gv = ldg4_stream(g + (i << 2));
        FqGrad a = elem(xv.x, gv.x), b = elem(xv.y, gv.y), c = elem(xv.z, gv.z), d = elem(xv.w, gv.w);
        if (gx) *reinterpret_cast<float4*>(gx + (i << 2)) = make_float4(a.gx, b.gx, c.gx, d.gx);
        acc_s += (a.gs + b.gs) + (c.gs + d.gs);
        acc_o += (a.go + b.go) + (c.go + d.go);
      }
    } else {
      for (int64_t i = tid; i < n; i += nthr) {
        if (group) { int64_t gi = i / group; s = __ldg(scale + gi); o = __ldg(offset + gi); }
        FqGrad a = elem(__ldg(x + i), __ldg(g + i));
        if (gx) gx[i] = a.gx;
        acc_s += a.gs;
        acc_o += a.go;
      }
    }
  };
  dispatch_five(group == 0 && mantissa_all_ones(s), run);
  if (partial) {            // the block that arrives last folds the per-block partials in a fixed order -- no second launch
    const float acc2[2] = {acc_s, acc_o};
    float* const outs[2] = {gscale, goffset};
    grid_fold_to<2>(acc2, partial, ticket, outs, red, &s_last);
  }
}

// ================================================================================================================
// K8 range statistics
// ================================================================================================================
template <bool kVec>
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ x, int64_t n,
                                                      float* __restrict__ partial /*[2*grid]*/) {
  __shared__ float red[32];
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = int64_t(gridDim.x) * blockDim.x;
  float mn = FLT_MAX, mx = -FLT_MAX;
  if (kVec) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += nthr) {
      float4 v = ldg4_stream(x + (i << 2));
      mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
      mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
    }
  } else {
    for (int64_t i = tid; i < n; i += nthr) { float v = __ldg(x + i); mn = fminf(mn, v); mx = fmaxf(mx, v); }
  }
  mn = block_reduce(mn, OpFMin(), red);
  mx = block_reduce(mx, OpFMax(), red);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = mn; partial[2 * blockIdx.x + 1] = mx; }
}

__global__ void minmax_final_kernel(const float* __restrict__ partial, int nblocks, float* minmax, int accumulate) {
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (int i = threadIdx.x; i < nblocks; i += 32) { mn = fminf(mn, partial[2 * i]); mx = fmaxf(mx, partial[2 * i + 1]); }
  mn = warp_reduce(mn, OpFMin());
  mx = warp_reduce(mx, OpFMax());
  if (threadIdx.x == 0) {
    if (accumulate) { mn = fminf(mn, minmax[0]); mx = fmaxf(mx, minmax[1]); }
    minmax[0] = mn; minmax[1] = mx;
  }
}

// per-row min/max: one CTA per row
__global__ void __launch_bounds__(256) minmax_rows_kernel(const float* __restrict__ x, int64_t cols,
                                                           float* __restrict__ out_min, float* __restrict__ out_max,
                                                           int accumulate) {
  __shared__ float red[32];
  const float* row = x + int64_t(blockIdx.x) * cols;
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (int64_t k = threadIdx.x; k < cols; k += blockDim.x) { float v = __ldg(row + k); mn = fminf(mn, v); mx = fmaxf(mx, v); }
  mn = block_reduce(mn, OpFMin(), red);
  mx = block_reduce(mx, OpFMax(), red);
  if (threadIdx.x == 0) {
    if (accumulate) { mn = fminf(mn, out_min[blockIdx.x]); mx = fmaxf(mx, out_max[blockIdx.x]); }
    out_min[blockIdx.x] = mn; out_max[blockIdx.x] = mx;
  }
}

// per-column min/max: thread per column, grid.y row segments, ordered-int atomics into ws
__global__ void minmax_cols_init_kernel(int* ws, int64_t cols) {
  int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c < cols) { ws[c] = float_to_ordered(FLT_MAX); ws[cols + c] = float_to_ordered(-FLT_MAX); }
}
__global__ void __launch_bounds__(256) minmax_cols_kernel(const float* __restrict__ x, int64_t rows, int64_t cols,
                                                           int* __restrict__ ws) {
  int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  int64_t r0 = rows * blockIdx.y / gridDim.y, r1 = rows * (blockIdx.y + 1) / gridDim.y;
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (int64_t r = r0; r < r1; ++r) { float v = __ldg(x + r * cols + c); mn = fminf(mn, v); mx = fmaxf(mx, v); }
  atomicMin(ws + c, float_to_ordered(mn));
  atomicMax(ws + cols + c, float_to_ordered(mx));
}
__global__ void minmax_cols_final_kernel(const int* __restrict__ ws, int64_t cols, float* out_min, float* out_max,
                                         int accumulate) {
  int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float mn = ordered_to_float(ws[c]), mx = ordered_to_float(ws[cols + c]);
  if (accumulate) { mn = fminf(mn, out_min[c]); mx = fmaxf(mx, out_max[c]); }
  out_min[c] = mn; out_max[c] = mx;
}

// ================================================================================================================
// K2 weight prep
// ================================================================================================================
struct LetArgs {
  const float* col_fac;  // [cols] or NULL
  const float* row_fac;  // [rows] or NULL
  int col_mode;          // 0 none, 1 divide, 2 multiply
  int row_mode;
};
__device__ __forceinline__ float let_apply(float w, float c, float r, int col_mode, int row_mode) {
  float t = w;
  if (col_mode == 2) t = fmul(t, c); else if (col_mode == 1) t = fdiv(t, c);
  if (row_mode == 1) t = fdiv(t, r); else if (row_mode == 2) t = fmul(t, r);
  return t;
}

// pass 1: per-row min/max of W'
__global__ void __launch_bounds__(256) wprep_rowminmax_kernel(const float* __restrict__ w, int64_t cols, LetArgs la,
                                                               float* __restrict__ row_mn, float* __restrict__ row_mx) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const float* wr = w + row * cols;
  const float r = la.row_mode ? __ldg(la.row_fac + row) : 1.f;
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (int64_t k = threadIdx.x; k < cols; k += blockDim.x) {
    float c = la.col_mode ? __ldg(la.col_fac + k) : 1.f;
    float t = let_apply(__ldg(wr + k), c, r, la.col_mode, la.row_mode);
    mn = fminf(mn, t); mx = fmaxf(mx, t);
  }
  mn = block_reduce(mn, OpFMin(), red);
  mx = block_reduce(mx, OpFMax(), red);
  if (threadIdx.x == 0) { row_mn[row] = mn; row_mx[row] = mx; }
}
// per-tensor: fold the row results into entry 0 (single block, fixed order)
__global__ void __launch_bounds__(256) wprep_fold_kernel(float* row_mn, float* row_mx, int64_t rows) {
  __shared__ float red[32];
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) { mn = fminf(mn, row_mn[i]); mx = fmaxf(mx, row_mx[i]); }
  mn = block_reduce(mn, OpFMin(), red);
  mx = block_reduce(mx, OpFMax(), red);
  __syncthreads();
  if (threadIdx.x == 0) { row_mn[0] = mn; row_mx[0] = mx; }
}

struct GroupQ { float s, o, qmin, qmax, mnp, mxp; };
__device__ __forceinline__ GroupQ group_quant(float mn, float mx, const float* sig_up, const float* sig_low, int64_t g,
                                              int bits, bool sym) {
  GroupQ q;
  q.mxp = sig_up ? fmul(__ldg(sig_up + g), mx) : mx;     // qm:271
  q.mnp = sig_low ? fmul(__ldg(sig_low + g), mn) : mn;   // qm:272
  scale_offset_from_minmax(q.mnp, q.mxp, bits, sym, q.s, q.o, q.qmin, q.qmax);
  return q;
}

// pass 2: quantise. one CTA per row.
__global__ void __launch_bounds__(256) wprep_quant_kernel(const float* __restrict__ w, int64_t cols, LetArgs la,
                                                           const float* __restrict__ row_mn,
                                                           const float* __restrict__ row_mx,
                                                           const float* __restrict__ sig_up,
                                                           const float* __restrict__ sig_low, int per_channel, int bits,
                                                           int sym, float* __restrict__ w_fq, uint8_t* __restrict__ codes,
                                                           int pack4, float* __restrict__ scale_out,
                                                           float* __restrict__ offset_out, int32_t* __restrict__ colsum,
                                                           float* __restrict__ wt_out, float* __restrict__ mm_out, int64_t groups) {
  __shared__ int redi[32];
  const int64_t row = blockIdx.x;
  const int64_t g = per_channel ? row : 0;
  const GroupQ q = group_quant(row_mn[g], row_mx[g], sig_up, sig_low, g, bits, sym != 0);
  const float* wr = w + row * cols;
  const float r = la.row_mode ? __ldg(la.row_fac + row) : 1.f;
  int csum = 0;
  // two elements per thread-iteration so that 4-bit packing writes whole bytes
  for (int64_t k = 2 * threadIdx.x; k < cols; k += 2 * blockDim.x) {
    float t[2]; int code[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int64_t kk = k + j;
      if (kk < cols) {
        float c = la.col_mode ? __ldg(la.col_fac + kk) : 1.f;
        t[j] = let_apply(__ldg(wr + kk), c, r, la.col_mode, la.row_mode);
        float qc = quant_code(t[j], q.s, q.o, q.qmin, q.qmax);
        code[j] = (int)qc;
        csum += code[j];
        if (w_fq) w_fq[row * cols + kk] = dequant(qc, q.s, q.o);
        if (wt_out) wt_out[row * cols + kk] = t[j];
      } else { t[j] = 0.f; code[j] = 0; }
    }
    if (codes) {
      if (pack4) {
        codes[(row * cols + k) >> 1] = (uint8_t)((code[0] & 0xF) | ((code[1] & 0xF) << 4));
      } else {
        codes[row * cols + k] = (uint8_t)(code[0] & 0xFF);
        if (k + 1 < cols) codes[row * cols + k + 1] = (uint8_t)(code[1] & 0xFF);
      }
    }
  }
  if (colsum) {
    int tot = block_reduce(csum, OpSum(), redi);
    if (threadIdx.x == 0) colsum[row] = tot;
  }
  if (threadIdx.x == 0 && (per_channel || row == 0)) {
    if (scale_out) scale_out[g] = q.s;
    if (offset_out) offset_out[g] = q.o;
    if (mm_out) { mm_out[g] = row_mn[g]; mm_out[groups + g] = row_mx[g]; }
  }
}

// ---- backward ---------------------------------------------------------------------------------------------------
// stats pass: per row  gs = sum_k dL/dscale terms, number of elements tied with the group min / max
__global__ void __launch_bounds__(256) wprep_bwd_stats_kernel(const float* __restrict__ w, const float* __restrict__ g,
                                                               int64_t cols, LetArgs la,
                                                               const float* __restrict__ row_mn,
                                                               const float* __restrict__ row_mx,
                                                               const float* __restrict__ sig_up,
                                                               const float* __restrict__ sig_low, int per_channel,
                                                               int bits, int sym, double* __restrict__ row_gs,
                                                               int* __restrict__ row_cmn, int* __restrict__ row_cmx) {
  __shared__ float redf[32];
  __shared__ int redi[32];
  const int64_t row = blockIdx.x;
  const int64_t gi = per_channel ? row : 0;
  const float gmn = row_mn[gi], gmx = row_mx[gi];
  const GroupQ q = group_quant(gmn, gmx, sig_up, sig_low, gi, bits, sym != 0);
  const float* wr = w + row * cols;
  const float* gr = g + row * cols;
  const float r = la.row_mode ? __ldg(la.row_fac + row) : 1.f;
  float acc = 0.f; int cmn = 0, cmx = 0;
  for (int64_t k = threadIdx.x; k < cols; k += blockDim.x) {
    float c = la.col_mode ? __ldg(la.col_fac + k) : 1.f;
    float t = let_apply(__ldg(wr + k), c, r, la.col_mode, la.row_mode);
    FqGrad e = fq_bwd_elem(t, __ldg(gr + k), q.s, q.o, q.qmin, q.qmax);
    acc += e.gs;
    cmn += (t == gmn); cmx += (t == gmx);
  }
  float tot = block_reduce(acc, OpSum(), redf);
  int tmn = block_reduce(cmn, OpSum(), redi);
  int tmx = block_reduce(cmx, OpSum(), redi);
  if (threadIdx.x == 0) { row_gs[row] = tot; row_cmn[row] = tmn; row_cmx[row] = tmx; }
}

// group pass: turn (gs, counts) into the per-tied-element gradient shares and the sigmoid-factor gradients.
// per-tensor: one block folds all rows first.
struct GroupGrad { float share_mn, share_mx; };
__global__ void __launch_bounds__(256) wprep_bwd_group_kernel(const float* __restrict__ row_mn,
                                                               const float* __restrict__ row_mx,
                                                               const float* __restrict__ sig_up,
                                                               const float* __restrict__ sig_low, int per_channel,
                                                               int bits, int sym, int64_t rows,
                                                               const double* __restrict__ row_gs,
                                                               const int* __restrict__ row_cmn,
                                                               const int* __restrict__ row_cmx,
                                                               GroupGrad* __restrict__ gg, float* __restrict__ g_sig_up,
                                                               float* __restrict__ g_sig_low) {
  __shared__ double redd[32];
  __shared__ int redi[32];
  int64_t gi; double gs; int cmn, cmx;
  if (per_channel) {
    gi = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gi >= rows) return;
    gs = row_gs[gi]; cmn = row_cmn[gi]; cmx = row_cmx[gi];
  } else {
    double a = 0.; int b = 0, c = 0;
    for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) { a += row_gs[i]; b += row_cmn[i]; c += row_cmx[i]; }
    gs = block_reduce(a, OpSum(), redd);
    cmn = block_reduce(b, OpSum(), redi);
    cmx = block_reduce(c, OpSum(), redi);
    if (threadIdx.x != 0) return;
    gi = 0;
  }
  const float mn = row_mn[gi], mx = row_mx[gi];
  const float su = sig_up ? __ldg(sig_up + gi) : 1.f, sl = sig_low ? __ldg(sig_low + gi) : 1.f;
  const float mxp = fmul(su, mx), mnp = fmul(sl, mn);
  float qmax_f, alpha;
  if (sym) { qmax_f = (float)((1 << (bits - 1)) - 1); alpha = fmaxf(fabsf(mnp), fabsf(mxp)); }
  else     { qmax_f = (float)((1 << bits) - 1);       alpha = fsub(mxp, mnp); }
  const float s_raw = fdiv(alpha, qmax_f);
  // clamp(min,max) passes gradient only inside [CLIPMIN, CLIPMAX] (inclusive, as torch.clamp does)
  float g_alpha = (s_raw >= MQ_CLIPMIN && s_raw <= MQ_CLIPMAX) ? fdiv((float)gs, qmax_f) : 0.f;
  float g_mxp, g_mnp;
  if (sym) {
    float a = fabsf(mnp), b = fabsf(mxp);
    float wmn = a > b ? 1.f : (a == b ? 0.5f : 0.f), wmx = 1.f - wmn;   // torch.maximum splits ties evenly
    float sgn_mn = (mnp > 0.f) - (mnp < 0.f), sgn_mx = (mxp > 0.f) - (mxp < 0.f);
    g_mnp = g_alpha * wmn * sgn_mn; g_mxp = g_alpha * wmx * sgn_mx;
  } else { g_mxp = g_alpha; g_mnp = -g_alpha; }
  if (g_sig_up) g_sig_up[gi] = sig_up ? fmul(g_mxp, mx) : 0.f;
  if (g_sig_low) g_sig_low[gi] = sig_low ? fmul(g_mnp, mn) : 0.f;
  gg[gi].share_mx = fdiv(fmul(g_mxp, su), (float)(cmx > 0 ? cmx : 1));
  gg[gi].share_mn = fdiv(fmul(g_mnp, sl), (float)(cmn > 0 ? cmn : 1));
}

// apply pass: dL/dW' per element, chain into the LET factors.
__global__ void __launch_bounds__(256) wprep_bwd_apply_kernel(const float* __restrict__ w, const float* __restrict__ g,
                                                               int64_t cols, LetArgs la,
                                                               const float* __restrict__ row_mn,
                                                               const float* __restrict__ row_mx,
                                                               const float* __restrict__ sig_up,
                                                               const float* __restrict__ sig_low, int per_channel,
                                                               int bits, int sym, const GroupGrad* __restrict__ gg,
                                                               float* __restrict__ col_contrib /*[rows,cols] or NULL*/,
                                                               float* __restrict__ g_row_fac /*[rows] or NULL*/,
                                                               float* __restrict__ g_wt /*[rows,cols] or NULL*/) {
  __shared__ float redf[32];
  const int64_t row = blockIdx.x;
  const int64_t gi = per_channel ? row : 0;
  const float gmn = row_mn[gi], gmx = row_mx[gi];
  const GroupQ q = group_quant(gmn, gmx, sig_up, sig_low, gi, bits, sym != 0);
  const GroupGrad sh = gg[gi];
  const float* wr = w + row * cols;
  const float* gr = g + row * cols;
  const float r = la.row_mode ? __ldg(la.row_fac + row) : 1.f;
  float acc_r = 0.f;
  for (int64_t k = threadIdx.x; k < cols; k += blockDim.x) {
    const float wv = __ldg(wr + k);
    const float c = la.col_mode ? __ldg(la.col_fac + k) : 1.f;
    float t = wv;                                   // after the column op
    if (la.col_mode == 2) t = fmul(wv, c); else if (la.col_mode == 1) t = fdiv(wv, c);
    float wp = t;                                   // after the row op == W'
    if (la.row_mode == 1) wp = fdiv(t, r); else if (la.row_mode == 2) wp = fmul(t, r);
    FqGrad e = fq_bwd_elem(wp, __ldg(gr + k), q.s, q.o, q.qmin, q.qmax);
    float dwp = e.gx;
    if (wp == gmx) dwp += sh.share_mx;
    if (wp == gmn) dwp += sh.share_mn;
    if (g_wt) g_wt[row * cols + k] = dwp;
    float gt = dwp;                                 // dL/dt
    if (la.row_mode == 1) { gt = fdiv(dwp, r); acc_r -= fmul(dwp, fdiv(wp, r)); }
    else if (la.row_mode == 2) { gt = fmul(dwp, r); acc_r += fmul(dwp, t); }
    if (col_contrib) {
      float cc = 0.f;
      if (la.col_mode == 2) cc = fmul(gt, wv); else if (la.col_mode == 1) cc = -fmul(gt, fdiv(t, c));
      col_contrib[row * cols + k] = cc;
    }
  }
  if (g_row_fac) {
    float tot = block_reduce(acc_r, OpSum(), redf);
    if (threadIdx.x == 0) g_row_fac[row] = tot;
  }
}

// ---- vectorised variants (cols % 4 == 0, 16-byte aligned rows) -------------------------------------------------------
// Same arithmetic, bit for bit, with 128-bit loads/stores and the branch-free exact division of common.cuh for every
// division by the group scale / the row factor (the divisor is uniform over a CTA, so its reciprocal and the FIVE
// variant are chosen once).  Divisions by a per-column LET factor (only norm weights, rows == 1) keep __fdiv_rn.
__device__ __forceinline__ float let_apply_v(float w, float c, const RowDiv& rd, int col_mode, int row_mode) {
  float t = w;
  if (col_mode == 2) t = fmul(t, c); else if (col_mode == 1) t = fdiv(t, c);
  if (row_mode == 1) t = div_any(t, rd); else if (row_mode == 2) t = fmul(t, rd.r);
  return t;
}
__device__ __forceinline__ float4 let_apply4(float4 w, const float* col_fac, int64_t k, const RowDiv& rd, const LetArgs& la) {
  float4 c = make_float4(1.f, 1.f, 1.f, 1.f);
  if (la.col_mode) c = ldg4(col_fac + k);
  return make_float4(let_apply_v(w.x, c.x, rd, la.col_mode, la.row_mode), let_apply_v(w.y, c.y, rd, la.col_mode, la.row_mode),
                     let_apply_v(w.z, c.z, rd, la.col_mode, la.row_mode), let_apply_v(w.w, c.w, rd, la.col_mode, la.row_mode));
}
__global__ void __launch_bounds__(256) wprep_rowminmax_v_kernel(const float* __restrict__ w, int64_t cols, LetArgs la,
                                                                 float* __restrict__ row_mn, float* __restrict__ row_mx) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const float* wr = w + row * cols;
  const RowDiv rd = make_rowdiv(la.row_mode ? __ldg(la.row_fac + row) : 1.f);
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (int64_t k = 4 * threadIdx.x; k < cols; k += 4 * blockDim.x) {
    const float4 t = let_apply4(ldg4_stream(wr + k), la.col_fac, k, rd, la);
    mn = fminf(fminf(mn, t.x), fminf(t.y, fminf(t.z, t.w)));
    mx = fmaxf(fmaxf(mx, t.x), fmaxf(t.y, fmaxf(t.z, t.w)));
  }
  mn = block_reduce(mn, OpFMin(), red);
  mx = block_reduce(mx, OpFMax(), red);
  if (threadIdx.x == 0) { row_mn[row] = mn; row_mx[row] = mx; }
}

template <bool FIVE>
__device__ __forceinline__ void wprep_quant_v_body(const float* __restrict__ w, int64_t cols, const LetArgs& la, const GroupQ& q,
                                                    int64_t row, float* __restrict__ w_fq, uint8_t* __restrict__ codes, int pack4,
                                                    int32_t* __restrict__ colsum, float* __restrict__ wt_out, int* redi) {
  const float* wr = w + row * cols;
  const RowDiv rd = make_rowdiv(la.row_mode ? __ldg(la.row_fac + row) : 1.f);
  const float rs = __frcp_rn(q.s);
  int csum = 0;
  for (int64_t k = 4 * threadIdx.x; k < cols; k += 4 * blockDim.x) {
    const float4 t = let_apply4(ldg4_stream(wr + k), la.col_fac, k, rd, la);
    const float tv[4] = {t.x, t.y, t.z, t.w};
    float qc[4]; int code[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { qc[j] = quant_code_v<FIVE>(tv[j], q.s, rs, q.o, q.qmin, q.qmax); code[j] = (int)qc[j]; csum += code[j]; }
    if (w_fq) *reinterpret_cast<float4*>(w_fq + row * cols + k) =
        make_float4(dequant(qc[0], q.s, q.o), dequant(qc[1], q.s, q.o), dequant(qc[2], q.s, q.o), dequant(qc[3], q.s, q.o));
    if (wt_out) *reinterpret_cast<float4*>(wt_out + row * cols + k) = t;
    if (codes) {
      if (pack4) {
        *reinterpret_cast<uint16_t*>(codes + ((row * cols + k) >> 1)) =
            (uint16_t)((code[0] & 0xF) | ((code[1] & 0xF) << 4) | ((code[2] & 0xF) << 8) | ((code[3] & 0xF) << 12));
      } else {
        *reinterpret_cast<uint32_t*>(codes + row * cols + k) =
            (uint32_t)(code[0] & 0xFF) | ((uint32_t)(code[1] & 0xFF) << 8) | ((uint32_t)(code[2] & 0xFF) << 16) | ((uint32_t)(code[3] & 0xFF) << 24);
      }
    }
  }
  if (colsum) {
    int tot = block_reduce(csum, OpSum(), redi);
    if (threadIdx.x == 0) colsum[row] = tot;
  }
}
__global__ void __launch_bounds__(256) wprep_quant_v_kernel(const float* __restrict__ w, int64_t cols, LetArgs la,
                                                             const float* __restrict__ row_mn, const float* __restrict__ row_mx,
                                                             const float* __restrict__ sig_up, const float* __restrict__ sig_low,
                                                             int per_channel, int bits, int sym, float* __restrict__ w_fq,
                                                             uint8_t* __restrict__ codes, int pack4, float* __restrict__ scale_out,
                                                             float* __restrict__ offset_out, int32_t* __restrict__ colsum,
                                                             float* __restrict__ wt_out, float* __restrict__ mm_out, int64_t groups) {
  __shared__ int redi[32];
  const int64_t row = blockIdx.x;
  const int64_t g = per_channel ? row : 0;
  const GroupQ q = group_quant(row_mn[g], row_mx[g], sig_up, sig_low, g, bits, sym != 0);
  if (mantissa_all_ones(q.s)) wprep_quant_v_body<true>(w, cols, la, q, row, w_fq, codes, pack4, colsum, wt_out, redi);
  else wprep_quant_v_body<false>(w, cols, la, q, row, w_fq, codes, pack4, colsum, wt_out, redi);
  if (threadIdx.x == 0 && (per_channel || row == 0)) {
    if (scale_out) scale_out[g] = q.s;
    if (offset_out) offset_out[g] = q.o;
    if (mm_out) { mm_out[g] = row_mn[g]; mm_out[groups + g] = row_mx[g]; }
  }
}

template <bool FIVE>
__device__ __forceinline__ void wprep_bwd_stats_v_body(const float* __restrict__ wr, const float* __restrict__ gr, int64_t cols,
                                                        const LetArgs& la, const RowDiv& rd, const GroupQ& q, float gmn, float gmx,
                                                        float& acc, int& cmn, int& cmx) {
  const float rs = __frcp_rn(q.s);
  for (int64_t k = 4 * threadIdx.x; k < cols; k += 4 * blockDim.x) {
    const float4 t = let_apply4(ldg4_stream(wr + k), la.col_fac, k, rd, la);
    const float4 gv = ldg4_stream(gr + k);
    const float tv[4] = {t.x, t.y, t.z, t.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const FqGrad e = fq_bwd_elem_v<FIVE>(tv[j], gg[j], q.s, rs, q.o, q.qmin, q.qmax);
      acc += e.gs;
      cmn += (tv[j] == gmn); cmx += (tv[j] == gmx);
    }
  }
}
__global__ void __launch_bounds__(256) wprep_bwd_stats_v_kernel(const float* __restrict__ w, const float* __restrict__ g,
                                                                 int64_t cols, LetArgs la, const float* __restrict__ row_mn,
                                                                 const float* __restrict__ row_mx, const float* __restrict__ sig_up,
                                                                 const float* __restrict__ sig_low, int per_channel, int bits, int sym,
                                                                 double* __restrict__ row_gs, int* __restrict__ row_cmn,
                                                                 int* __restrict__ row_cmx) {
  __shared__ float redf[32];
  __shared__ int redi[32];
  const int64_t row = blockIdx.x;
  const int64_t gi = per_channel ? row : 0;
  const float gmn = row_mn[gi], gmx = row_mx[gi];
  const GroupQ q = group_quant(gmn, gmx, sig_up, sig_low, gi, bits, sym != 0);
  const RowDiv rd = make_rowdiv(la.row_mode ? __ldg(la.row_fac + row) : 1.f);
  float acc = 0.f; int cmn = 0, cmx = 0;
  if (mantissa_all_ones(q.s)) wprep_bwd_stats_v_body<true>(w + row * cols, g + row * cols, cols, la, rd, q, gmn, gmx, acc, cmn, cmx);
  else wprep_bwd_stats_v_body<false>(w + row * cols, g + row * cols, cols, la, rd, q, gmn, gmx, acc, cmn, cmx);
  float tot = block_reduce(acc, OpSum(), redf);
  int tmn = block_reduce(cmn, OpSum(), redi);
  int tmx = block_reduce(cmx, OpSum(), redi);
  if (threadIdx.x == 0) { row_gs[row] = tot; row_cmn[row] = tmn; row_cmx[row] = tmx; }
}

template <bool FIVE>
__device__ __forceinline__ float wprep_bwd_apply_v_body(const float* __restrict__ wr, const float* __restrict__ gr, int64_t cols,
                                                         const LetArgs& la, const RowDiv& rd, const GroupQ& q, float gmn, float gmx,
                                                         const GroupGrad& sh, float* __restrict__ col_contrib, float* __restrict__ g_wt) {
  const float rs = __frcp_rn(q.s);
  float acc_r = 0.f;
  for (int64_t k = 4 * threadIdx.x; k < cols; k += 4 * blockDim.x) {
    const float4 w4 = ldg4_stream(wr + k), g4 = ldg4_stream(gr + k);
    float4 c4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (la.col_mode) c4 = ldg4(la.col_fac + k);
    const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w}, cv[4] = {c4.x, c4.y, c4.z, c4.w};
    float dw[4], cc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = wv[j];                                   // after the column op
      if (la.col_mode == 2) t = fmul(wv[j], cv[j]); else if (la.col_mode == 1) t = fdiv(wv[j], cv[j]);
      float wp = t;                                      // after the row op == W'
      if (la.row_mode == 1) wp = div_any(t, rd); else if (la.row_mode == 2) wp = fmul(t, rd.r);
      const FqGrad e = fq_bwd_elem_v<FIVE>(wp, gv[j], q.s, rs, q.o, q.qmin, q.qmax);
      float dwp = e.gx;
      if (wp == gmx) dwp += sh.share_mx;
      if (wp == gmn) dwp += sh.share_mn;
      dw[j] = dwp;
      float gt = dwp;                                    // dL/dt
      if (la.row_mode == 1) { gt = div_any(dwp, rd); acc_r -= fmul(dwp, div_any(wp, rd)); }
      else if (la.row_mode == 2) { gt = fmul(dwp, rd.r); acc_r += fmul(dwp, t); }
      cc[j] = 0.f;
      if (la.col_mode == 2) cc[j] = fmul(gt, wv[j]); else if (la.col_mode == 1) cc[j] = -fmul(gt, fdiv(t, cv[j]));
    }
    if (g_wt) *reinterpret_cast<float4*>(g_wt + k) = make_float4(dw[0], dw[1], dw[2], dw[3]);
    if (col_contrib) *reinterpret_cast<float4*>(col_contrib + k) = make_float4(cc[0], cc[1], cc[2], cc[3]);
  }
  return acc_r;
}
__global__ void __launch_bounds__(256) wprep_bwd_apply_v_kernel(const float* __restrict__ w, const float* __restrict__ g,
                                                                 int64_t cols, LetArgs la, const float* __restrict__ row_mn,
                                                                 const float* __restrict__ row_mx, const float* __restrict__ sig_up,
                                                                 const float* __restrict__ sig_low, int per_channel, int bits, int sym,
                                                                 const GroupGrad* __restrict__ gg, float* __restrict__ col_contrib,
                                                                 float* __restrict__ g_row_fac, float* __restrict__ g_wt) {
  __shared__ float redf[32];
  const int64_t row = blockIdx.x;
  const int64_t gi = per_channel ? row : 0;
  const float gmn = row_mn[gi], gmx = row_mx[gi];
  const GroupQ q = group_quant(gmn, gmx, sig_up, sig_low, gi, bits, sym != 0);
  const GroupGrad sh = gg[gi];
  const RowDiv rd = make_rowdiv(la.row_mode ? __ldg(la.row_fac + row) : 1.f);
  float* cc = col_contrib ? col_contrib + row * cols : nullptr;
  float* gw = g_wt ? g_wt + row * cols : nullptr;
  float acc_r = mantissa_all_ones(q.s) ? wprep_bwd_apply_v_body<true>(w + row * cols, g + row * cols, cols, la, rd, q, gmn, gmx, sh, cc, gw)
                                       : wprep_bwd_apply_v_body<false>(w + row * cols, g + row * cols, cols, la, rd, q, gmn, gmx, sh, cc, gw);
  if (g_row_fac) {
    float tot = block_reduce(acc_r, OpSum(), redf);
    if (threadIdx.x == 0) g_row_fac[row] = tot;
  }
}

// deterministic column sums of a [rows, cols] matrix: grid (col tiles of 128, S row segments) -> partial[S, cols]
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ a, int64_t rows, int64_t cols,
                                                              float* __restrict__ partial) {
  __shared__ float sm[8][128];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t c0 = int64_t(blockIdx.x) * 128 + lane * 4;
  const int64_t r0 = rows * blockIdx.y / gridDim.y, r1 = rows * (blockIdx.y + 1) / gridDim.y;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t r = r0 + wid; r < r1; r += 8) {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (c0 + j < cols) acc[j] += __ldg(a + r * cols + c0 + j);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) sm[wid][lane * 4 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 128) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sm[i][threadIdx.x];
    int64_t c = int64_t(blockIdx.x) * 128 + threadIdx.x;
    if (c < cols) partial[int64_t(blockIdx.y) * cols + c] = s;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int S, int64_t cols, float* __restrict__ out) {
  int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double s = 0.;
  for (int i = 0; i < S; ++i) s += partial[int64_t(i) * cols + c];
  out[c] = (float)s;
}


// ================================================================================================================
// Self-test of the branch-free exact requantisation (common.cuh: div_rn / quant_int) against the IEEE reference
// sequence rintf(__fdiv_rn(x, s)) on pseudo-random operands.  mode 0: random scales; 1: one fixed scale, integer-valued
// numerators times a random factor (the GEMM / attention epilogue pattern); 2: scales with an all-ones significand.
// ================================================================================================================
__device__ __forceinline__ uint32_t xs32(uint64_t& st) {
  st ^= st << 13; st ^= st >> 7; st ^= st << 17;
  return (uint32_t)(st >> 16);
}
__global__ void __launch_bounds__(256) selftest_div_kernel(int64_t n, uint64_t seed, int mode, float fixed_scale,
                                                            unsigned long long* mism) {
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = int64_t(gridDim.x) * blockDim.x;
  uint64_t st = seed * 0x9E3779B97F4A7C15ull + (uint64_t)tid * 0xD1342543DE82EF95ull + 1;
  unsigned long long bad = 0;
  for (int64_t i = tid; i < n; i += nthr) {
    float a, b;
    if (mode == 1) {
      const int I = (int)(xs32(st) & 0xffffff) - 0x800000;
      a = __fmul_rn(__int2float_rn(I), __uint_as_float(0x38000000u + (xs32(st) & 0x03ffffffu)));   // * [3e-5, 0.5)
      b = fixed_scale;
    } else {
      a = __uint_as_float((xs32(st) & 0x807fffffu) | ((100u + (xs32(st) % 50u)) << 23));           // |a| in 2^[-27, 23)
      b = __uint_as_float((xs32(st) & 0x007fffffu) | ((106u + (xs32(st) % 30u)) << 23));           // b in 2^[-21, 9)
      if (mode == 2) b = __uint_as_float(__float_as_uint(b) | 0x007fffffu);
    }
    const float ref = __fdiv_rn(a, b);
    const float rb = __frcp_rn(b);
    const bool five = mantissa_all_ones(b);
    const float got = five ? div_rn<true>(a, b, rb) : div_rn<false>(a, b, rb);
    bad += (__float_as_uint(ref) != __float_as_uint(got)) && !(ref == 0.f && got == 0.f);
    // code path: clamp(rne(a/b)+o, 0, qmax) for a 16-bit quantizer with an arbitrary integral offset
    const float o = (float)(xs32(st) & 0xffff), qmax = 65535.f;
    const float cref = fminf(fmaxf(__fadd_rn(rintf(ref), o), 0.f), qmax);
    const QParam qp = make_qparam(b, o, qmax);
    const int cgot = five ? quant_int<true>(a, qp) : quant_int<false>(a, qp);
    bad += ((int)cref != cgot);
  }
  bad = warp_reduce(bad, OpSum());
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mism, bad);
}

}  // namespace mq

using namespace mq;

extern "C" {

int mq_fq_fwd(void* ctx, const float* x, float* y, int32_t* codes, int64_t n, const float* scale, const float* offset,
              int64_t group, float qmin, float qmax, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, x && scale && offset && n >= 0 && group >= 0, "null pointer or negative size");
  if (n == 0) return MQ_NO_ERROR;
  cudaStream_t st = (cudaStream_t)stream;
  bool vec = (n % 4 == 0) && aligned16(x) && (!y || aligned16(y)) && (!codes || aligned16(codes)) && (group % 4 == 0);
  int grid = grid_for(c, vec ? n / 4 : n, 256, 8);
  if (vec) fq_fwd_kernel<true><<<grid, 256, 0, st>>>(x, y, codes, n, scale, offset, group, qmin, qmax);
  else fq_fwd_kernel<false><<<grid, 256, 0, st>>>(x, y, codes, n, scale, offset, group, qmin, qmax);
  return check_launch(c, "mq_fq_fwd");
}

int mq_fq_bwd(void* ctx, const float* x, const float* g, float* gx, int64_t n, const float* scale, const float* offset,
              int64_t group, float qmin, float qmax, float* gscale, float* goffset, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, x && g && scale && offset && n >= 0 && group >= 0, "null pointer or negative size");
  MQ_REQUIRE(c, group == 0 || (!gscale && !goffset), "gscale/goffset need group == 0");
  cudaStream_t st = (cudaStream_t)stream;
  bool want = gscale || goffset;
  if (n == 0) {
    if (gscale) cudaMemsetAsync(gscale, 0, sizeof(float), st);
    if (goffset) cudaMemsetAsync(goffset, 0, sizeof(float), st);
    return MQ_NO_ERROR;
  }
  bool vec = (n % 4 == 0) && aligned16(x) && aligned16(g) && (!gx || aligned16(gx)) && (group % 4 == 0);
  int grid = grid_for(c, vec ? n / 4 : n, 256, 8);
  void* wsp = want ? stream_ws(c, st) : nullptr;
  if (want && !wsp) return MQ_FAILED_ALLOCATION;
  double* partial = reinterpret_cast<double*>(wsp);
  unsigned* ticket = want ? reinterpret_cast<unsigned*>(static_cast<char*>(wsp) + c->ws_bytes - 64) : nullptr;
  if (vec) fq_bwd_kernel<true><<<grid, 256, 0, st>>>(x, g, gx, n, scale, offset, group, qmin, qmax, partial, ticket, gscale, goffset);
  else fq_bwd_kernel<false><<<grid, 256, 0, st>>>(x, g, gx, n, scale, offset, group, qmin, qmax, partial, ticket, gscale, goffset);
  return check_launch(c, "mq_fq_bwd");
}

int mq_minmax(void* ctx, const float* x, int64_t n, float* minmax, int accumulate, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, x && minmax && n > 0, "null pointer or empty tensor");
  cudaStream_t st = (cudaStream_t)stream;
  bool vec = (n % 4 == 0) && aligned16(x);
  int grid = grid_for(c, vec ? n / 4 : n, 256 * 4, 8);
  float* partial = reinterpret_cast<float*>(stream_ws(c, st));
  if (!partial) return MQ_FAILED_ALLOCATION;
  if (vec) minmax_kernel<true><<<grid, 256, 0, st>>>(x, n, partial);
  else minmax_kernel<false><<<grid, 256, 0, st>>>(x, n, partial);
  minmax_final_kernel<<<1, 32, 0, st>>>(partial, grid, minmax, accumulate);
  return check_launch(c, "mq_minmax");
}

int mq_minmax_2d(void* ctx, const float* x, int64_t rows, int64_t cols, int per_row, float* out_min, float* out_max,
                 int accumulate, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, x && out_min && out_max && rows > 0 && cols > 0, "null pointer or empty tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (per_row) {
    minmax_rows_kernel<<<(unsigned)rows, wprep_threads(cols), 0, st>>>(x, cols, out_min, out_max, accumulate);
  } else {
    MQ_REQUIRE(c, size_t(cols) * 2 * sizeof(int) + 64 <= c->ws_bytes, "too many columns for the workspace");
    int* ws = reinterpret_cast<int*>(stream_ws(c, st));
    if (!ws) return MQ_FAILED_ALLOCATION;
    unsigned gx = (unsigned)((cols + 255) / 256);
    int seg = (int)((int64_t(c->sm_count) * 4 + gx - 1) / gx);
    if (seg > rows) seg = (int)rows;
    if (seg < 1) seg = 1;
    minmax_cols_init_kernel<<<gx, 256, 0, st>>>(ws, cols);
    minmax_cols_kernel<<<dim3(gx, seg), 256, 0, st>>>(x, rows, cols, ws);
    minmax_cols_final_kernel<<<gx, 256, 0, st>>>(ws, cols, out_min, out_max, accumulate);
  }
  return check_launch(c, "mq_minmax_2d");
}

// the vectorised kernels need 16-byte aligned rows of every fp32 operand (and whole words of packed codes)
static bool wprep_vec_ok(int64_t cols, const void* a, const void* b, const void* c, const void* d, const void* codes, int /*pack4*/) {
  auto al = [](const void* p, uintptr_t m) { return (reinterpret_cast<uintptr_t>(p) & m) == 0; };
  return cols % 4 == 0 && al(a, 15) && al(b, 15) && al(c, 15) && al(d, 15) && al(codes, 3);
}

static int wprep_check(Ctx* c, const float* w, int64_t rows, int64_t cols, const float* col_fac, int col_mode,
                       const float* row_fac, int row_mode, mq_qcfg cfg) {
  MQ_REQUIRE(c, w && rows > 0 && cols > 0, "null pointer or empty weight");
  MQ_REQUIRE(c, col_mode >= 0 && col_mode <= 2 && row_mode >= 0 && row_mode <= 2, "bad LET mode");
  MQ_REQUIRE(c, (col_mode == 0) || col_fac, "col_mode set but col_fac is NULL");
  MQ_REQUIRE(c, (row_mode == 0) || row_fac, "row_mode set but row_fac is NULL");
  MQ_REQUIRE(c, cfg.bitwidth >= 2 && cfg.bitwidth <= 16, "bitwidth must be in [2,16]");
  MQ_REQUIRE(c, size_t(rows) * 32 + 128 <= c->ws_bytes, "too many rows for the workspace");
  return MQ_NO_ERROR;
}

int mq_wprep_fwd(void* ctx, const float* w, int64_t rows, int64_t cols, const float* col_fac, int col_mode,
                 const float* row_fac, int row_mode, const float* sig_up, const float* sig_low, int per_channel,
                 mq_qcfg cfg, float* w_fq, void* codes, int pack4, float* scale_out, float* offset_out, int32_t* colsum,
                 float* wt_out, float* minmax_out, void* stream) {
  MQ_CTX(c, ctx);
  if (int e = wprep_check(c, w, rows, cols, col_fac, col_mode, row_fac, row_mode, cfg)) return e;
  const int64_t groups = per_channel ? rows : 1;
  MQ_REQUIRE(c, !codes || cfg.bitwidth <= 8, "integer codes are stored in 8 bits");
  MQ_REQUIRE(c, !pack4 || (cfg.bitwidth <= 4 && cols % 2 == 0), "pack4 needs bitwidth <= 4 and even cols");
  cudaStream_t st = (cudaStream_t)stream;
  LetArgs la{col_fac, row_fac, col_mode, row_mode};
  float* row_mn = reinterpret_cast<float*>(stream_ws(c, st));
  if (!row_mn) return MQ_FAILED_ALLOCATION;
  float* row_mx = row_mn + rows;
  const bool vec = wprep_vec_ok(cols, w, col_fac, w_fq, wt_out, codes, pack4);
  if (vec) wprep_rowminmax_v_kernel<<<(unsigned)rows, wprep_threads(cols), 0, st>>>(w, cols, la, row_mn, row_mx);
  else wprep_rowminmax_kernel<<<(unsigned)rows, wprep_threads(cols), 0, st>>>(w, cols, la, row_mn, row_mx);
  if (!per_channel) wprep_fold_kernel<<<1, 256, 0, st>>>(row_mn, row_mx, rows);
  if (vec)
    wprep_quant_v_kernel<<<(unsigned)rows, wprep_threads(cols), 0, st>>>(w, cols, la, row_mn, row_mx, sig_up, sig_low, per_channel,
                                                         cfg.bitwidth, cfg.is_symmetric, w_fq, (uint8_t*)codes, pack4,
                                                         scale_out, offset_out, colsum, wt_out, minmax_out, groups);
  else
    wprep_quant_kernel<<<(unsigned)rows, wprep_threads(cols), 0, st>>>(w, cols, la, row_mn, row_mx, sig_up, sig_low, per_channel,
                                                       cfg.bitwidth, cfg.is_symmetric, w_fq, (uint8_t*)codes, pack4,
                                                       scale_out, offset_out, colsum, wt_out, minmax_out, groups);
  return check_launch(c, "mq_wprep_fwd");
}

int mq_wprep_bwd(void* ctx, const float* w, const float* g, int64_t rows, int64_t cols, const float* col_fac,
                 int col_mode, const float* row_fac, int row_mode, const float* sig_up, const float* sig_low,
                 int per_channel, mq_qcfg cfg, float* g_col_fac, float* g_row_fac, float* g_sig_up, float* g_sig_low,
                 float* g_wt, float* scratch, const float* minmax_in, void* stream) {
  MQ_CTX(c, ctx);
  if (int e = wprep_check(c, w, rows, cols, col_fac, col_mode, row_fac, row_mode, cfg)) return e;
  MQ_REQUIRE(c, g != nullptr, "g is NULL");
  MQ_REQUIRE(c, !g_col_fac || (scratch && col_mode), "g_col_fac needs scratch and col_mode != 0");
  MQ_REQUIRE(c, !g_row_fac || row_mode, "g_row_fac needs row_mode != 0");
  cudaStream_t st = (cudaStream_t)stream;
  LetArgs la{col_fac, row_fac, col_mode, row_mode};
  // workspace carve-up (all sized by rows): mn, mx | gs (double) | cmn, cmx | GroupGrad
  char* const ws0 = reinterpret_cast<char*>(stream_ws(c, st));
  if (!ws0) return MQ_FAILED_ALLOCATION;
  char* p = ws0;
  float* row_mn = reinterpret_cast<float*>(p); p += rows * sizeof(float);
  float* row_mx = reinterpret_cast<float*>(p); p += rows * sizeof(float);
  p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
  double* row_gs = reinterpret_cast<double*>(p); p += rows * sizeof(double);
  int* row_cmn = reinterpret_cast<int*>(p); p += rows * sizeof(int);
  int* row_cmx = reinterpret_cast<int*>(p); p += rows * sizeof(int);
  GroupGrad* gg = reinterpret_cast<GroupGrad*>(p); p += rows * sizeof(GroupGrad);
  p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
  float* partial = reinterpret_cast<float*>(p);
  // (wprep_check bounds rows * 32 bytes; the carve-up above adds at most 32 bytes of alignment padding per context -- checked)
  MQ_REQUIRE(c, size_t(p - ws0) + 80 <= c->ws_bytes, "too many rows for the workspace");
  size_t partial_cap = (c->ws_bytes - 64 - size_t(p - ws0)) / sizeof(float);      // (the last 64 bytes are arrival counters)

  const bool vec = wprep_vec_ok(cols, w, col_fac, g, g_wt, scratch, 0);
  if (minmax_in) {                      // group min / max as the forward left them (mq_wprep_fwd minmax_out): no second pass over w
    row_mn = const_cast<float*>(minmax_in);
    row_mx = row_mn + (per_channel ? rows : 1);
  } else {
    if (vec) wprep_rowminmax_v_kernel<<<(unsigned)rows, wprep_threads(cols), 0, st>>>(w, cols, la, row_mn, row_mx);
    else wprep_rowminmax_kernel<<<(unsigned)rows, wprep_threads(cols), 0, st>>>(w, cols, la, row_mn, row_mx);
    if (!per_channel) wprep_fold_kernel<<<1, 256, 0, st>>>(row_mn, row_mx, rows);
  }
  if (vec)
    wprep_bwd_stats_v_kernel<<<(unsigned)rows, wprep_threads(cols), 0, st>>>(w, g, cols, la, row_mn, row_mx, sig_up, sig_low, per_channel,
                                                             cfg.bitwidth, cfg.is_symmetric, row_gs, row_cmn, row_cmx);
  else
    wprep_bwd_stats_kernel<<<(unsigned)rows, wprep_threads(cols), 0, st>>>(w, g, cols, la, row_mn, row_mx, sig_up, sig_low, per_channel,
                                                           cfg.bitwidth, cfg.is_symmetric, row_gs, row_cmn, row_cmx);
  unsigned ggrid = per_channel ? (unsigned)((rows + 255) / 256) : 1u;
  wprep_bwd_group_kernel<<<ggrid, 256, 0, st>>>(row_mn, row_mx, sig_up, sig_low, per_channel, cfg.bitwidth,
                                                cfg.is_symmetric, rows, row_gs, row_cmn, row_cmx, gg, g_sig_up,
                                                g_sig_low);
  if (g_col_fac || g_row_fac || g_wt) {
    if (vec)
      wprep_bwd_apply_v_kernel<<<(unsigned)rows, wprep_threads(cols), 0, st>>>(w, g, cols, la, row_mn, row_mx, sig_up, sig_low,
                                                               per_channel, cfg.bitwidth, cfg.is_symmetric, gg,
                                                               g_col_fac ? scratch : nullptr, g_row_fac, g_wt);
    else
      wprep_bwd_apply_kernel<<<(unsigned)rows, wprep_threads(cols), 0, st>>>(w, g, cols, la, row_mn, row_mx, sig_up, sig_low,
                                                             per_channel, cfg.bitwidth, cfg.is_symmetric, gg,
                                                             g_col_fac ? scratch : nullptr, g_row_fac, g_wt);
  }
  if (g_col_fac) {
    unsigned gx = (unsigned)((cols + 127) / 128);
    int S = (int)((int64_t(c->sm_count) * 2 + gx - 1) / gx);
    if (S > rows) S = (int)rows;
    if (S < 1) S = 1;
    while (size_t(S) * cols > partial_cap && S > 1) --S;
    MQ_REQUIRE(c, size_t(S) * cols <= partial_cap, "workspace too small for column sums");
    colsum_partial_kernel<<<dim3(gx, S), 256, 0, st>>>(scratch, rows, cols, partial);
    colsum_final_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(partial, S, cols, g_col_fac);
  }
  return check_launch(c, "mq_wprep_bwd");
}

int mq_selftest_div(void* ctx, int64_t n, uint64_t seed, int mode, float fixed_scale, uint64_t* mismatches, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, n > 0 && mismatches && mode >= 0 && mode <= 2, "bad arguments");
  MQ_REQUIRE(c, mode != 1 || fixed_scale > 0.f, "mode 1 needs a positive scale");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(mismatches, 0, sizeof(uint64_t), st);
  selftest_div_kernel<<<c->sm_count * 8, 256, 0, st>>>(n, seed, mode, fixed_scale, reinterpret_cast<unsigned long long*>(mismatches));
  return check_launch(c, "mq_selftest_div");
}

}  // extern "C"
