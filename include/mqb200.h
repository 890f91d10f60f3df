/* libmqb200 -- C ABI of the B200-native MobileQuant hot path.
 *
 * The reference (saic-fi/MobileQuant) has no FFI on this path: its boundary is the Python module surface of
 * mobilellm/quantization/qmodule.py and algorithm.py.  This header is the thin C ABI that sits *behind* that
 * surface (mobilequant_b200/quantization/*.py binds it with ctypes).  Conventions follow the reference's only
 * C ABI, capp/api/libllmod.h:10-17,42-133: extern "C", every call returns an int status (0 = OK), an opaque
 * context created by mq_setup / freed by mq_release (magic + refcount validated, capp/src/libllmod.cpp:23-65),
 * per-context last-error text, and no C++ exception ever crosses the boundary.
 *
 * Ownership: every device buffer is allocated and owned by the caller (PyTorch); the library receives raw device
 * pointers, element counts and a cudaStream_t (passed as void*).  All calls are asynchronous on that stream and
 * never synchronise.  One host thread per context.
 *
 * Each entry point cites the reference code whose arithmetic it replaces ("qm" = mobilellm/quantization/qmodule.py,
 * "alg" = mobilellm/quantization/algorithm.py, "hm" = mobilellm/model/hf_model.py).
 */
#ifndef MQB200_H
#define MQB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MQB200_VERSION 100 /* major*100 + minor */

enum mq_status_code { /* mirrors libllmod_status_code, capp/api/libllmod.h:10-17 */
  MQ_NO_ERROR = 0,
  MQ_INVALID_CONTEXT,
  MQ_INVALID_ARGUMENT,
  MQ_FAILED_ALLOCATION,
  MQ_RUNTIME_ERROR,
  MQ_INTERNAL_ERROR,
};

/* Quantizer description == QuantConfig (qm:81-107) restricted to what reaches a kernel. */
typedef struct mq_qcfg {
  int32_t bitwidth;     /* 2..16; >16 means pass-through and never reaches the library (qm:252) */
  int32_t is_symmetric; /* qm:45-54 */
} mq_qcfg;

int mq_version(void);
const char* mq_get_error_description(int errorcode);                 /* libllmod.h:119 */
const char* mq_get_last_error_extra_info(int errorcode, void* ctx);  /* libllmod.h:133 */
int mq_setup(void** ctx, int device);                                /* libllmod.h:42  */
int mq_ref_context(void* ctx);                                       /* libllmod.h:67  */
int mq_release(void* ctx);                                           /* libllmod.h:73  */
int mq_device_sm_count(void* ctx, int* out);

/* ---- K1: static fake-quant, Quantizer.forward with cached scale/offset (qm:279-295) -------------------------
 * y = (clamp(rne(x/scale)+offset, qmin, qmax) - offset) * scale.  scale/offset are DEVICE pointers:
 *   group == 0 : one (scale, offset) for the whole tensor (per-tensor activations, LRL parameters)
 *   group  > 0 : element i uses scale[i / group] (per-channel weights: group = in_features)
 * y and/or codes may be NULL.  codes receives the integer code as int32 (exact when offset is an integer).     */
int mq_fq_fwd(void* ctx, const float* x, float* y, int32_t* codes, int64_t n, const float* scale,
              const float* offset, int64_t group, float qmin, float qmax, void* stream);

/* Backward of the above as autograd derives it from qm:17-21,286-290:
 *   gx = g*m;  gscale = sum g*(m*(rne(u)-u) + (1-m)*(qc-o));  goffset = -sum g*(1-m)*scale,   u = x/scale.
 * gscale/goffset (may be NULL) are per-group DEVICE accumulators that are OVERWRITTEN (deterministic two-stage
 * reduction in ctx workspace).  Only group == 0 supports gscale/goffset.                                        */
int mq_fq_bwd(void* ctx, const float* x, const float* g, float* gx, int64_t n, const float* scale,
              const float* offset, int64_t group, float qmin, float qmax, float* gscale, float* goffset,
              void* stream);

/* ---- Calibration attention core, the element-wise chain between the two batched matmuls of HFAttention.forward
 * (hm:514-534) with the QMatMul quantizers around it (qm:453-466), fused:
 *   P = fq2( softmax( fq1(S) * mul + causal_mask ) )      S, P: [rows, T] fp32, rows = B*nh*Tq
 * fq1 = qk_bmm.output_quantizer, fq2 = pv_bmm.input_quantizer (per-tensor static; a NULL scale/offset pair disables one),
 * mul = fp32 1/sqrt(head_dim), causal != 0: row r sees columns <= (r % Tq) + (T - Tq) (finfo.min elsewhere, as
 * _prepare_4d_causal_attention_mask builds it, hm:1548-1555).  stats[rows][2] receives (row max, sum of exp) for the
 * backward.  T % 4 == 0 and T <= 2048 (mq_attn_probs_supported); one launch each way, nothing else materialised.
 * Backward: dS from g = dL/dP; gparams (may be NULL) = DEVICE float[4] OVERWRITTEN with d/dscale1, d/doffset1,
 * d/dscale2, d/doffset2 (deterministic fixed-order reduction in the ctx workspace of `stream`).                  */
int mq_attn_probs_supported(int T);
int mq_attn_probs_fwd(void* ctx, const float* S, float* P, float* stats, int64_t rows, int T, int Tq, int causal, float mul,
                      const float* scale1, const float* offset1, float qmin1, float qmax1, const float* scale2,
                      const float* offset2, float qmin2, float qmax2, void* stream);
int mq_attn_probs_bwd(void* ctx, const float* S, const float* stats, const float* g, float* dS, int64_t rows, int T, int Tq,
                      int causal, float mul, const float* scale1, const float* offset1, float qmin1, float qmax1,
                      const float* scale2, const float* offset2, float qmin2, float qmax2, float* gparams, void* stream);

/* ---- Calibration MLP core of a gated-SiLU block (hm:1042-1062 with QSiLU, qm:691-753, the output quantizers of w1 / w3 and
 * w2.input_quantizer), fused:
 *   out = fq_w( fq_o( a * fq_s(sigmoid(a)) ) * b ),   a = fq_a(ya), b = fq_b(yb)      ya = w1 x + b1, yb = w3 x + b3
 * ya, yb: [rows, cols] fp32 with row stride ld (the two halves of ONE [rows, 2*cols] GEMM result, or two tensors with ld == cols);
 * out: [rows, cols] contiguous; cols % 4 == 0.  scales / offsets: HOST arrays of 5 DEVICE pointers in the order fq_a
 * (w1.output_quantizer), fq_b (w3.output_quantizer), fq_s (QSiLU.input2_quantizer), fq_o (QSiLU.output_quantizer), fq_w
 * (w2.input_quantizer); a NULL pair disables that quantizer; qmins / qmaxs: HOST float[5].
 * Backward: dya, dyb (row stride ldd) from g = dL/dout; gparams (may be NULL) = DEVICE float[10] OVERWRITTEN with
 * (d/dscale, d/doffset) of the five quantizers in that order (deterministic fixed-order reduction, ctx workspace of `stream`). */
int mq_silu_gate_fwd(void* ctx, const float* ya, const float* yb, int64_t ld, float* out, int64_t rows, int cols,
                     const float* const* scales, const float* const* offsets, const float* qmins, const float* qmaxs, void* stream);
int mq_silu_gate_bwd(void* ctx, const float* ya, const float* yb, int64_t ld, const float* g, float* dya, float* dyb, int64_t ldd,
                     int64_t rows, int cols, const float* const* scales, const float* const* offsets, const float* qmins,
                     const float* qmaxs, float* gparams, void* stream);

/* ---- Calibration QRMSNorm in its L2-norm form (qm:515-531 over hm:187-195 / F.normalize), fused with both quantizers:
 *   out = fq_out( w * (alpha * xq / max(||xq||_2, eps)) + bias ),  xq = fq_in(x)      x, out: [rows, H] fp32, H % 4 == 0, H <= 8192
 * w = the (already fake-quantised) norm weight [H], bias may be NULL; nrm[rows] receives ||xq||_2 for the backward.
 * scales / offsets: HOST arrays of 2 DEVICE pointers (input, output quantizer; NULL pair = disabled), qmins / qmaxs HOST float[2].
 * Backward: dx [rows, H], dw [H], dbias [H] (may be NULL) from g = dL/dout; gparams (may be NULL) = DEVICE float[4] OVERWRITTEN
 * with (d/dscale, d/doffset) of the input and the output quantizer.  All reductions are fixed-order (deterministic).         */
int mq_rmsnorm_l2_supported(int H);
int mq_rmsnorm_l2_fwd(void* ctx, const float* x, const float* w, const float* bias, float* out, float* nrm, int64_t rows, int H,
                      float alpha, float eps, const float* const* scales, const float* const* offsets, const float* qmins,
                      const float* qmaxs, void* stream);
int mq_rmsnorm_l2_bwd(void* ctx, const float* x, const float* w, const float* bias, const float* nrm, const float* g, float* dx,
                      float* dw, float* dbias, int64_t rows, int H, float alpha, float eps, const float* const* scales,
                      const float* const* offsets, const float* qmins, const float* qmaxs, float* gparams, void* stream);

/* ---- Calibration Q/K/V post-processing between the projection GEMM and the attention matmuls (hm:470-512 with the QLinear
 * output quantizers, qm:356-358, and the QMatMul input quantizers, qm:455-458), fused:
 *   q = fq_3( rope( fq_0(yq) ) ) -> [B, nh, T, hd]    k = fq_4( rope( fq_1(yk) ) ) -> [B, nkv, T, hd]    v = fq_5( fq_2(yv) ) -> [B, nkv, T, hd]
 * y = [yq | yk | yv]: [rows = B*T, (nh + 2 nkv) * hd] fp32, the result of ONE GEMM over the concatenated projection weights.
 * rope: (x * cos) + (rotate_half(x) * sin) on the first `rot` dims of a head (hm:338-367; rot % 8 == 0, (hd - rot) % 8 == 0),
 * cos / sin: [B or 1, T, rot] fp32 (cs_batched != 0: one table per batch row).  scales / offsets: HOST arrays of 6 DEVICE
 * pointers in the order q_proj / k_proj / v_proj.output_quantizer, qk_bmm.input / input2_quantizer, pv_bmm.input2_quantizer
 * (NULL pair = disabled); qmins / qmaxs: HOST float[6].
 * Backward: dy [rows, (nh + 2 nkv) * hd] from dq / dk / dv; gparams (may be NULL) = DEVICE float[12] OVERWRITTEN with
 * (d/dscale, d/doffset) of the six quantizers in that order (deterministic fixed-order reduction, ctx workspace of `stream`). */
int mq_qkv_rope_fwd(void* ctx, const float* y, int64_t rows, int T, int nh, int nkv, int hd, int rot, const float* cos,
                    const float* sin, int cs_batched, float* q, float* k, float* v, const float* const* scales,
                    const float* const* offsets, const float* qmins, const float* qmaxs, void* stream);
int mq_qkv_rope_bwd(void* ctx, const float* y, int64_t rows, int T, int nh, int nkv, int hd, int rot, const float* cos,
                    const float* sin, int cs_batched, const float* dq, const float* dk, const float* dv, float* dy,
                    const float* const* scales, const float* const* offsets, const float* qmins, const float* qmaxs,
                    float* gparams, void* stream);

/* ---- K8: range statistics, generate_act_range.py:55-69 (per tensor) / :57-63 (per channel) -------------------
 * minmax[0] = min(minmax[0], min x), minmax[1] = max(minmax[1], max x) when accumulate != 0, else overwritten.
 * rows variant: x is [rows, cols]; per_row != 0 reduces over cols (weights, qm:30) else over rows (per-channel
 * activation stats); out_min/out_max have rows or cols entries.                                                 */
int mq_minmax(void* ctx, const float* x, int64_t n, float* minmax, int accumulate, void* stream);
int mq_minmax_2d(void* ctx, const float* x, int64_t rows, int64_t cols, int per_row, float* out_min,
                 float* out_max, int accumulate, void* stream);

/* ---- K2: weight prep = LET transform (alg:60-96) + dynamic/LWC Quantizer.forward (qm:262-290) -----------------
 * W' = (W {* or /} col_fac[k]) {/ or *} row_fac[n]   (modes: 0 none, 1 divide, 2 multiply; NULL when mode 0).
 *   ln.weight / s (alg:60) is rows=1, col_mode=1;  fc.weight * s (alg:68) col_mode=2;  fc1/q: row_mode=1 (alg:77,93);
 *   k_proj: row_mode=2 (alg:95).
 * (mn,mx) = per tensor (per_channel == 0) or per row; mx *= sig_up[g]; mn *= sig_low[g] (NULL = no LWC)
 * scale/offset per group from qm:40-61; outputs (any may be NULL):
 *   w_fq   fp32 fake-quantised W'            (what QLinear.forward feeds F.linear, qm:347-353)
 *   codes  int8 storage: asymmetric -> uint8 codes, symmetric -> int8 codes; bitwidth 4 with pack4 != 0 packs two
 *          codes per byte (low nibble = even k)
 *   scale_out/offset_out [groups]; colsum int32 [rows] = sum_k code  (zero-point correction of the int GEMM)
 *   wt_out fp32 W' before quantisation (the temp_weight of alg:68, kept for the backward)
 *   minmax_out [2 * groups]: the group minima then maxima of W' (before LWC) -- hand it to mq_wprep_bwd as minmax_in and the
 *          backward skips its own min / max pass over w                                                          */
int mq_wprep_fwd(void* ctx, const float* w, int64_t rows, int64_t cols, const float* col_fac, int col_mode,
                 const float* row_fac, int row_mode, const float* sig_up, const float* sig_low, int per_channel,
                 mq_qcfg cfg, float* w_fq, void* codes, int pack4, float* scale_out, float* offset_out,
                 int32_t* colsum, float* wt_out, float* minmax_out, void* stream);

/* Backward of mq_wprep_fwd: given g = dL/dw_fq produce dL/dcol_fac [cols], dL/drow_fac [rows], dL/dsig_up,
 * dL/dsig_low [groups] (any may be NULL), including the amin/amax paths of qm:264-275 (gradient of a min/max is
 * split evenly among tied elements, as torch.amin/amax do).  g_wt (may be NULL) receives dL/dW' [rows, cols] for
 * callers that built W' themselves (Quantizer.forward on an arbitrary tensor).  scratch: rows*cols floats owned by
 * the caller, needed only with g_col_fac.                                                                       */
int mq_wprep_bwd(void* ctx, const float* w, const float* g, int64_t rows, int64_t cols, const float* col_fac,
                 int col_mode, const float* row_fac, int row_mode, const float* sig_up, const float* sig_low, int per_channel,
                 mq_qcfg cfg, float* g_col_fac, float* g_row_fac, float* g_sig_up, float* g_sig_low,
                 float* g_wt, float* scratch, const float* minmax_in, void* stream);

/* ---- K3/K7: integer GEMM with fused requantisation epilogue == QLinear.forward on codes (qm:341-358) ------------
 * acc = A[M,K] (u8/s8 codes, row-major) x B[N,K]^T (u8/s8 weight codes from mq_wprep_fwd), s32 accumulate on
 * tcgen05 tensor cores.  Zero points are removed exactly in the epilogue:
 *   I = acc - ow[n]*rowsum[m] + c0[n],  c0[n] = K*ox*ow[n] - ox*colsum[n];   y = float(I)*sxw[n] (+ bias[n])
 * so/oo describe the (per-tensor, qm:216-245) output quantizers of the fused column segments: one (scale, integral offset)
 * entry per group of `qgroup` consecutive columns (qgroup % 32 == 0; e.g. the fused q|k|v projection has three segments).
 * mode 0 QUANT : out = clamp(rne(y/so[g]) + oo[g], 0, qmax) as u8 (out_bits 8) or u16 (16), ld = ldo elements;
 *                rowsum_out[m] (optional, zero-initialised by the caller) += sum_n code  (for the next GEMM)
 * mode 1 ACTMUL: B rows are interleaved per 256-row tile as [128 rows of w1 | the same 128 rows of w3];
 *                out[m, j] = Q_w2in( lut[Q_w1out(y1)] * dequant(Q_w3out(y3)) )  u8, N/2 columns (HFMLP, hm:1057-1060 with
 *                QSiLU/QGELU qm:739-753,790-799 folded into the 256-entry lut); qgroup must divide 128
 * mode 2 RESID : resid[m,n] += dequant(Q_out(y))   fp32 in place (o_proj / w2 + residual add, hm:1257,1270); the add is
 *                performed by the L2 (TMA reduce-add, round-to-nearest; subnormal sums flush to zero)
 * mode 3 F32   : out = y (fp32);  mode 4 I32: out = I (int32)                                                   */
int mq_qgemm(void* ctx, const void* a_codes, int a_signed, const void* b_codes, int b_signed, int M, int N, int K,
             const int32_t* rowsum, const float* sxw, const int32_t* ow, const int32_t* c0, const float* bias, int mode,
             const float* so, const float* oo, float qmax, int out_bits, void* out, int64_t ldo, int32_t* rowsum_out,
             const float* lut, float s2, float o2, float qmax2, float* resid, int qgroup, void* stream);

/* mq_qgemm_w4a8: the same GEMM with the weight operand PACKED, two 4-bit codes per byte (b_packed [N, K/2] bytes, the code
 * of column 2j in the low nibble of byte j, UNSIGNED nibbles 0..15; K % 32 == 0).  Symmetric 4-bit weights (codes -8..7,
 * experiments/w4a8/main/e2e_llama-s1024-ep60-sym.sh:26) are passed in offset-binary form code + 8 with ow[n] = 8: the zero-
 * point algebra above is unchanged (c0 is invariant under that shift).  The packed k-slices are staged by TMA and expanded
 * to one code per byte INSIDE the kernel (four warps write the 128B-swizzled operand rows the tensor core reads), so the
 * weights stream from HBM at 4 bits per code and no unpacked copy ever exists in global memory (BASELINE config 3).       */
int mq_qgemm_w4a8(void* ctx, const void* a_codes, int a_signed, const void* b_packed, int M, int N, int K,
                  const int32_t* rowsum, const float* sxw, const int32_t* ow, const int32_t* c0, const float* bias,
                  int mode, const float* so, const float* oo, float qmax, int out_bits, void* out, int64_t ldo,
                  int32_t* rowsum_out, const float* lut, float s2, float o2, float qmax2, float* resid, int qgroup,
                  void* stream);

/* ---- K4: QRMSNorm.forward (qm:515-531, L2-norm form hm:187-195) / QLayerNorm.forward (qm:625-642) on codes ---------
 * x fp32 residual stream [rows, H] -> 16-bit input quantizer -> norm with the fake-quantised weight w_fq (from
 * mq_wprep_fwd) -> 8-bit output quantizer: codes u8 [rows, H] and rowsum[rows] = sum of the codes.               */
int mq_qnorm(void* ctx, const float* x, int64_t rows, int H, int is_layernorm, float s_in, float o_in, float qmax_in,
             const float* w_fq, const float* bias, float alpha, float eps, float s_out, float o_out, float qmax_out,
             uint8_t* codes, int32_t* rowsum, void* stream);

/* mq_qnorm_resid (decode step): mq_qgemv_epilogue mode 2 (RESID) of the preceding skinny GEMM folded into the norm of the
 * same rows: x[m, :] += dequant(Q_out(y)) from the s32 accumulator acc[rows, ldacc] (handed back zeroed; g_* are that
 * GEMM's rowsum / sxw / ow / c0 / bias / output quantizer, N == H), then mq_qnorm of the updated rows.  rows <= 256.     */
int mq_qnorm_resid(void* ctx, float* x, int rows, int H, int is_layernorm, float s_in, float o_in, float qmax_in, const float* w_fq,
                   const float* bias, float alpha, float eps, float s_out, float o_out, float qmax_out, uint8_t* codes, int32_t* rowsum,
                   int32_t* acc, int ldacc, const int32_t* g_rowsum, const float* g_sxw, const int32_t* g_ow, const int32_t* g_c0,
                   const float* g_bias, const float* g_so, const float* g_oo, float g_qmax, int g_qgroup, void* stream);

/* ---- K5: RoPE between two quantizers (hm:486-501 + qm:455-459) ------------------------------------------------------
 * qkv: u8 codes [B*T, ldq] of the fused q|k|v projection.  in_qparams / out_qparams are HOST arrays
 * {s_q,o_q,s_k,o_k,s_v,o_v}: projection output quantizers, then qk_bmm.input / qk_bmm.input2 / pv_bmm.input2.
 * cos/sin: device [T, rot] (hm:308-318).  Outputs: q [B,nh,T,hd], k [B,nkv,T,hd], vt [B,nkv,hd,T] u8 codes and the
 * per-row code sums rsq [B,nh,T], rsk [B,nkv,T].                                                                 */
int mq_qrope(void* ctx, const uint8_t* qkv, int ldq, int B, int T, int nh, int nkv, int hd, int rot, const float* in_qparams,
             const float* out_qparams, const float* cos, const float* sin, uint8_t* q, uint8_t* k, uint8_t* vt, int32_t* rsq,
             int32_t* rsk, void* stream);

/* ---- K6: quantised causal attention (HFAttention.forward hm:510-534 with QMatMul qm:453-466) -----------------------
 * qparams (HOST) = {o_q, o_k, o_v, s_q*s_k, s_s, o_s, qmax_s, s_p, qmax_p, s_p*s_v, s_out, o_out}; lut (device u32[512])
 * is the two-level exp table A[256] | B[256], A[i] = rne(2^31 exp(-256 i a)), B[j] = rne(2^31 exp(-j a)),
 * a = s_s / sqrt(hd): exp(-k a) for a 16-bit code distance k is evaluated as (A[k >> 8] * B[k & 255]) >> 31.
 * out: u8 codes [B*T, nh*hd]; rowsum_out[B*T] (optional, zero-initialised by the caller) accumulates the emitted codes
 * for the o_proj zero-point correction.  Offsets must be integral (ranges reloaded through act_dict.json, qm:60).   */
int mq_qattn(void* ctx, const uint8_t* q, const uint8_t* k, const uint8_t* vt, const int32_t* rsq, const int32_t* rsk, int B,
             int T, int nh, int nkv, int hd, const float* qparams, const uint32_t* lut, uint8_t* out, int32_t* rowsum_out,
             void* stream);

/* mq_qattn_shard: the same attention for a SHARD of the queries (sequence-sharded prefill): q / rsq / out /
 * rowsum_out hold Tq query rows per (batch, head) whose absolute positions are q_start .. q_start + Tq - 1; k / vt / rsk
 * hold all T keys of the sequence (the shards' K/V codes are exchanged by the caller, 2 * hd + 4 bytes per token and kv
 * head); query i sees keys 0 .. q_start + i.  q_start must be a multiple of 128 and the shape must be covered by the
 * tcgen05 kernel (hd 64 / 128, T % 16 == 0); mq_qattn is the case q_start = 0, Tq = T.                                  */
int mq_qattn_shard(void* ctx, const uint8_t* q, const uint8_t* k, const uint8_t* vt, const int32_t* rsq, const int32_t* rsk, int B,
                   int Tq, int T, int q_start, int nh, int nkv, int hd, const float* qparams, const uint32_t* lut, uint8_t* out,
                   int32_t* rowsum_out, void* stream);

/* ---- decode step: one new token per sequence against an int8 KV cache ---------------------------------------------
 * Reference: SimModel.generate / SimModel.forward with k_cache, v_cache (mobilellm/model/sim_model.py:105-132,181-235) and
 * the on-device loop capp/src/llm.cpp:545-653 (uint8 caches [L, n_heads, T-1, head_dim]).  The static quantizers are the
 * prefill's, so every integer tensor of a decode step equals row `pos` of the full-sequence forward (mq_qgemm / mq_qattn).
 *
 * mq_qgemv: acc[b, n] += sum_k x[b, k] * w[n, k] for B <= 128 rows (u8/s8 codes, s32 accumulate on tcgen05 with the
 * weights as the 128-row operand); the K loop is split over `ksplit` CTAs per 128-row weight tile (<= 0: chosen so that
 * every SM streams) and the partial sums are combined with integer red.add, so acc must be zero on entry (exact,
 * order-independent).  mq_qgemv_epilogue applies mq_qgemm's zero-point removal and epilogue `mode` (0 QUANT 8-bit,
 * 1 ACTMUL, 2 RESID; same arguments, same arithmetic) to acc[B, N] and leaves acc zeroed for the next call.        */
int mq_qgemv(void* ctx, const void* x_codes, int x_signed, const void* w_codes, int w_signed, int B, int N, int K, int32_t* acc,
             int ldacc, int ksplit, void* stream);
int mq_qgemv_epilogue(void* ctx, int32_t* acc, int ldacc, int B, int N, const int32_t* rowsum, const float* sxw, const int32_t* ow,
                      const int32_t* c0, const float* bias, int mode, const float* so, const float* oo, float qmax, uint8_t* out,
                      int64_t ldo, int32_t* rowsum_out, const float* lut, float s2, float o2, float qmax2, float* resid, int qgroup,
                      int32_t* zero_out, void* stream);
/* mq_qgemv_fused = mq_qgemv + mq_qgemv_epilogue in one launch: the last CTA to arrive at a group of 128 output columns
 * (arrival counters inside the context, so one stream per context) finds all partial sums in L2 and runs the epilogue. */
int mq_qgemv_fused(void* ctx, const void* x_codes, int x_signed, const void* w_codes, int w_signed, int B, int N, int K, int32_t* acc,
                   int ldacc, int ksplit, const int32_t* rowsum, const float* sxw, const int32_t* ow, const int32_t* c0, const float* bias,
                   int mode, const float* so, const float* oo, float qmax, uint8_t* out, int64_t ldo, int32_t* rowsum_out, const float* lut,
                   float s2, float o2, float qmax2, float* resid, int qgroup, int32_t* zero_out, void* stream);
/* zero_out (optional, [B]): a code-sum buffer some LATER kernel of the step accumulates into; cleared here so that the
 * step needs no separate memset launches (it must not alias rowsum / rowsum_out of this call).
 *
 * mq_fgemv: out[b, v] = sum_k x[b, k] * w[v, k] in fp32 for B <= 16 rows -- the unquantised lm_head (qm:843-845) of the
 * decode step, one pass over w at HBM speed (fixed summation order, run-to-run deterministic).                      */
int mq_fgemv(void* ctx, const float* x, const float* w, float* out, int B, int V, int K, void* stream);

/* mq_qattn_decode: qkv u8 codes [B, ldq] of the new token's fused q|k|v projection -> RoPE at position pos between the
 * projection output quantizers and the bmm input quantizers (rope_in/out_qparams as in mq_qrope, HOST arrays; cos/sin
 * device [> pos, rot]) -> k / v codes and the k code sum appended at row pos of the caches
 *   k_cache, v_cache u8 [B, nkv, Tmax, hd],  rsk_cache s32 [B, nkv, Tmax]
 * -> exact quantised softmax attention of the new row over keys 0..pos (qparams / lut as in mq_qattn) -> out u8 [B, nh*hd]
 * and rowsum_out[B] += sum of the emitted codes.  pos is read from *pos_dev when pos_dev != NULL (CUDA-graph replay of
 * the step; pos_bound is then the largest position the launch must accommodate), else from `pos`.                   */
int mq_qattn_decode(void* ctx, const uint8_t* qkv, int ldq, int B, int nh, int nkv, int hd, int rot, int Tmax, int pos,
                    const int* pos_dev, int pos_bound, const float* rope_in_qparams, const float* rope_out_qparams, const float* cos,
                    const float* sin, uint8_t* k_cache, uint8_t* v_cache, int32_t* rsk_cache, const float* qparams,
                    const uint32_t* lut, uint8_t* out, int32_t* rowsum_out, void* stream);

/* mq_unpack4: n_codes packed 4-bit weight codes (two per byte, low nibble first, as written by mq_wprep_fwd with pack4)
 * -> one code per byte (int8 when is_signed, i.e. symmetric weights, else uint8): W4A8 weights stay packed in HBM and a
 * layer is expanded into an L2-sized scratch buffer right before its mq_qgemm / mq_qgemv.  n_codes % 32 == 0.        */
int mq_unpack4(void* ctx, const uint8_t* packed, int64_t n_codes, int is_signed, void* out, void* stream);

/* ---- test hook ---------------------------------------------------------------------------------------------------
 * The integer-engine kernels requantise with a branch-free exact division (RN(a/b) from RN(1/b) and two FMAs, a
 * third/fourth for scales whose significand is all ones) instead of the IEEE division + rint of qm:286.  This entry
 * point checks that path against __fdiv_rn / rintf on n pseudo-random operand pairs and writes the number of
 * disagreements to *mismatches (device u64).  mode 0: random scales; 1: `fixed_scale` with integer-valued numerators;
 * 2: all-ones significands.                                                                                         */
int mq_selftest_div(void* ctx, int64_t n, uint64_t seed, int mode, float fixed_scale, uint64_t* mismatches, void* stream);

/* ---- optimiser step of the calibration loops (alg:513,716-722 torch.optim.AdamW + mobilellm/utils/optim.py:28-41) ----
 * One flat fp32 buffer holds every learnable (LET scales, LWC bound factors, LRL scales / offsets); `grads`,
 * `exp_avg`, `exp_avg_sq` are buffers of the same length n.  Elements [seg_end[i-1], seg_end[i]) use learning rate
 * lr[i] (device array, ngroups <= 8; seg_end is a HOST array).  Two launches: (1) global L2 norm of grads (double
 * accumulation, fixed order) -> state[1]; a non-finite norm sets state[2] = 1, counts a skipped step in state[5] and
 * leaves parameters, moments and the step counter untouched (GradScaler.step semantics); (2) AdamW update with
 * betas / eps / decoupled weight decay, bias corrections from the step counter state[0] (kept on the device so that
 * the whole training step can be replayed as a CUDA graph).  state: device float[8], zero-initialised by the caller. */
int mq_adamw_step(void* ctx, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int ngroups,
                  const int64_t* seg_end, const float* lr, float beta1, float beta2, float eps, float weight_decay,
                  float* state, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MQB200_H */
