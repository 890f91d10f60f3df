// Decode step of the statically-quantised integer forward: one new token per sequence against an int8 KV cache.
//
// The reference decodes with SimModel.generate (mobilellm/model/sim_model.py:181-235: context encoding, then one-token
// steps that append to k_cache / v_cache [L, n_heads, T-1, head_dim]) and, on the phone, with the capp loop
// (capp/src/llm.cpp:545-653, uint8 caches); both run the same static quantizers as the prefill, so the integer tensors
// of a decode step equal row `pos` of a full-sequence forward.  Kernels of the step:
//
//   mq_qgemv          skinny GEMM  acc[b, n] += sum_k X[b, k] W[n, k]  for B <= 128 rows: the WEIGHTS are the 128-row
//                     UMMA operand (tcgen05.mma kind::i8, M = 128, N = padded batch), every weight byte leaves HBM exactly
//                     once, the K loop is split across CTAs so that all SMs stream, partial sums meet in an s32
//                     accumulator with red.global.add (integer adds commute: the result is exact and deterministic)
//   mq_qgemv_epilogue the requantisation epilogues of mq_qgemm (QUANT / ACTMUL / RESID) on that accumulator, same
//                     arithmetic (gv_epi.cuh), and the accumulator is handed back zeroed; mq_qgemv_fused runs it inside
//                     the GEMV (last CTA of a column group); mq_qnorm_resid (engine.cu) folds RESID into the next norm
//   mq_qattn_decode   RoPE + requant of the new token's q / k / v codes, append to the cache, exact quantised softmax
//                     attention of the new row over keys 0..pos (same integer arithmetic as mq_qattn); a cluster of CTAs
//                     splits the keys and exchanges exact integer partials through distributed shared memory
//   mq_fgemv          the unquantised fp32 lm_head for <= 16 rows, one pass over the weights
//   mq_unpack4        packed 4-bit weight codes -> one code per byte right before a GEMM (W4A8 weights stay packed in HBM)
// All are HBM- or latency-bound (weights / cache streamed once); oracle: oracle/int_ref.py (IntModel.decode == row pos of
// the full forward, tests/test_decode_oracle_cpu.py).
#include "common.cuh"
#include "tc_common.cuh"
#include "ctx.h"
#include "gv_epi.cuh"
#include <string>
#include <algorithm>
#include <cstdlib>

namespace mq {
using namespace tc;

// =====================================================================================================================
// epilogue of the skinny GEMM: same arithmetic as qgemm_kernel's epilogue (qgemm.cu); the unit of work is one row m and
// four consecutive output columns.  Used by the standalone epilogue kernel and by the last CTA of a column group inside
// qgemv_kernel (fused path).
// =====================================================================================================================
// standalone epilogue: blockIdx.y = m so that the emitted-code sum of a block belongs to one row
template <int MODE>
__global__ void __launch_bounds__(128) qgemv_epi_kernel(const GvEpiArgs a) {
  __shared__ int s_sum[4];
  pdl_trigger();
  pdl_wait();                                                       // the accumulator comes from the GEMV before this launch
  const int m = blockIdx.y;
  const int NO = MODE == GV_ACTMUL ? a.N / 2 : a.N;                 // output columns
  const int j0 = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (a.zero_out && blockIdx.x == 0 && threadIdx.x == 0) a.zero_out[m] = 0;
  int csum = j0 < NO ? gv_epi_quad<MODE>(a, m, j0) : 0;
  if (MODE != GV_RESID && a.rowsum_out) {
    csum = warp_reduce(csum, OpSum());
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = csum;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(a.rowsum_out + m, s_sum[0] + s_sum[1] + s_sum[2] + s_sum[3]);
  }
}

// =====================================================================================================================
// qgemv
// =====================================================================================================================
constexpr int kGvBM = 128;        // weight rows per tile == UMMA M
constexpr int kGvBK = 128;        // bytes of K per stage == one 128B swizzle row
constexpr int kGvStages = 4;      // 4 x (16 KB + BP x 128 B): two CTAs per SM stay resident up to BP = 64

struct QGemvArgs {
  int B, N, K;
  int32_t* acc;       // [B, ldacc] s32, accumulated with red.add
  int ldacc;
  int ksplit;
  int mode;           // GV_NONE: raw accumulator only; else the last CTA of each column group runs that epilogue
  int* counters;      // [column groups] arrival counters (zero on entry, reset by the last CTA)
  int w_const;        // the weight codes are not written by the preceding kernels of the stream: tiles may be requested early
  GvEpiArgs epi;
};

template <int BP>
struct GvSmem {
  static constexpr int kWBytes = kGvBM * kGvBK;
  static constexpr int kXBytes = BP * kGvBK;
  static constexpr int kXOff = kGvStages * kWBytes;
  static constexpr int kBarOff = kXOff + kGvStages * kXBytes;
  static constexpr int kTotal = kBarOff + 128;
  static constexpr int kTmemCols = BP < 32 ? 32 : BP;
};

template <int BP>
__global__ void __launch_bounds__(192, 1)
qgemv_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, const QGemvArgs p, const uint32_t idesc) {
  using L = GvSmem<BP>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_w = smem;
  uint8_t* smem_x = smem + L::kXOff;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + kGvStages;
  uint64_t* tfull_bar = empty_bar + kGvStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  volatile int* epi_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / p.ksplit, ks = blockIdx.x % p.ksplit;
  const int k_iters = (p.K + kGvBK - 1) / kGvBK;
  const int k0 = (int)((long long)ks * k_iters / p.ksplit), k1 = (int)((long long)(ks + 1) * k_iters / p.ksplit);
  if (k0 >= k1) return;                         // uniform per CTA: nothing allocated yet
  const int n0 = tile * kGvBM;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_x);
    for (int i = 0; i < kGvStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, L::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // The weights are constants: the first ring of weight tiles is requested BEFORE the grid dependency resolves, so a
      // dependent launch streams them while the kernel that produces x is still running; x follows after pdl_wait().
      const int npre = p.w_const ? min(kGvStages, k1 - k0) : 0;
      for (int i = 0; i < npre; ++i) {
        mbar_expect_tx(&full_bar[i], L::kWBytes + L::kXBytes);
        tma_load_2d(smem_w + i * L::kWBytes, &tmap_w, &full_bar[i], (k0 + i) * kGvBK, n0);
      }
      pdl_wait();
      for (int i = 0; i < npre; ++i)
        tma_load_2d(smem_x + i * L::kXBytes, &tmap_x, &full_bar[i], (k0 + i) * kGvBK, 0);      // rows >= B are zero-filled
      int stage = npre == kGvStages ? 0 : npre; uint32_t phase = npre == kGvStages ? 1 : 0;
      for (int k = k0 + npre; k < k1; ++k) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], L::kWBytes + L::kXBytes);
        tma_load_2d(smem_w + stage * L::kWBytes, &tmap_w, &full_bar[stage], k * kGvBK, n0);
        tma_load_2d(smem_x + stage * L::kXBytes, &tmap_x, &full_bar[stage], k * kGvBK, 0);
        if (++stage == kGvStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    int stage = 0; uint32_t phase = 0;
    for (int k = k0; k < k1; ++k) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t adesc = smem_desc_k128(smem_u32(smem_w + stage * L::kWBytes));
        const uint64_t bdesc = smem_desc_k128(smem_u32(smem_x + stage * L::kXBytes));
#pragma unroll
        for (int kk = 0; kk < kGvBK / 32; ++kk)
          mma_i8(tmem_base, adesc + uint64_t(kk * 2), bdesc + uint64_t(kk * 2), idesc, (k > k0) | (kk != 0));
        tc_commit(&empty_bar[stage]);
        if (k == k1 - 1) tc_commit(tfull_bar);
      }
      __syncwarp();
      if (++stage == kGvStages) { stage = 0; phase ^= 1; }
    }
  } else {
    // warps 2..5: TMEM lane quarter (warp & 3) == 32 weight rows; thread = weight row n, registers = batch columns
    const int quarter = warp & 3;
    pdl_wait();                                 // the accumulator these warps add into is zeroed by an earlier kernel of the chain
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const int n = n0 + quarter * 32 + lane;
    const uint32_t trow = tmem_base + (uint32_t(quarter * 32) << 16);
#pragma unroll
    for (int c = 0; c < BP; c += 16) {
      if (c >= p.B) break;
      uint32_t r[16];
      tmem_ld16(trow + c, r);
      tc_wait_ld();
      if (n < p.N) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c + j < p.B) atomicAdd(p.acc + int64_t(c + j) * p.ldacc + n, (int)r[j]);   // result unused -> RED.ADD
      }
    }
    if (p.mode != GV_NONE) {
      // ---- fused epilogue: the last CTA to arrive at a column group (one 128-row weight tile; a w1|w3 tile pair for
      // ACTMUL) finds every partial sum in L2 and requantises the group for all B rows
      const int et = threadIdx.x - 64;                       // 0..127
      const int gr = p.mode == GV_ACTMUL ? 2 : 1;
      const int group = tile / gr;
      __threadfence();                                        // this thread's red.adds before the arrival
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        const int total = gr * p.ksplit;
        const int old = atomicAdd(p.counters + group, 1);
        *epi_flag = (old == total - 1);
        if (old == total - 1) p.counters[group] = 0;          // self-cleaning for the next GEMV on this stream
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (*epi_flag) {
        __threadfence();
        const int NO = p.mode == GV_ACTMUL ? p.N / 2 : p.N;
        const int jbase = group * 128;                        // 128 output columns per group in every mode
        if (group == 0 && p.epi.zero_out) for (int m = et; m < p.B; m += 128) p.epi.zero_out[m] = 0;
        for (int item = et; item < p.B * 32; item += 128) {   // a warp's 32 items share the row m
          const int m = item >> 5, j0 = jbase + (item & 31) * 4;
          int csum = 0;
          if (j0 < NO) {
            if (p.mode == GV_QUANT) csum = gv_epi_quad<GV_QUANT>(p.epi, m, j0);
            else if (p.mode == GV_ACTMUL) csum = gv_epi_quad<GV_ACTMUL>(p.epi, m, j0);
            else csum = gv_epi_quad<GV_RESID>(p.epi, m, j0);
          }
          if (p.mode != GV_RESID && p.epi.rowsum_out) {
            csum = warp_reduce(csum, OpSum());
            if (lane == 0) atomicAdd(p.epi.rowsum_out + m, csum);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, L::kTmemCols);
}

// =====================================================================================================================
// mq_qattn_decode: one CTA per (sequence, kv head); the R = nh / nkv query heads that share the kv head ride together.
//   1. new token: de-quantise q/k/v projection codes, RoPE at position pos, re-quantise (== qrope_kernel, engine.cu);
//      k / v codes and the k code sum are appended to the cache at row pos
//   2. scores of the R query rows against keys 0..pos (dp4a on the u8 codes, exact zero-point removal), 16-bit codes
//   3. row max -> E = tab(cmax - c) -> exact u64 row sum -> p = E / sum -> 16-bit prob codes   (== qattn kernels)
//   4. A = P.V on the codes, zero-point removal, 8-bit output quantizer, token-major store, code sum for o_proj
// Cache layouts: K, V  u8 [B, nkv, Tmax, hd];  rsk  s32 [B, nkv, Tmax].
// =====================================================================================================================
struct AttnDecArgs {
  const uint8_t* qkv; int ldq;
  int B, nh, nkv, hd, rot, Tmax, pos;
  const int* pos_dev;
  int pos_max;                                        // largest position this launch was sized for (< Tmax)
  float sq_in, oq_in, sk_in, ok_in, sv_in, ov_in;     // projection output quantizers
  float sq, oq, sk, ok, sv, ov;                       // qk_bmm.input / input2, pv_bmm.input2
  const float* cos; const float* sin;                 // [>= pos + 1, rot]
  uint8_t* kc; uint8_t* vc; int32_t* rskc;
  float sqk, s_s, o_s, qmax_s, s_p, qmax_p, spv, s_out, o_out;
  const uint32_t* lut;
  uint8_t* out; int32_t* rowsum_out;
  int CS;                                             // CTAs per cluster == key slices per (sequence, kv head, row group)
  int RG;                                             // query heads of the kv head handled by one cluster (divides nh / nkv)
  int Tslice;                                         // score slab stride (multiple of 8 >= ceil((pos + 1) / CS))
};

// cross-CTA exchange block (one per CTA, read by the cluster peers through distributed shared memory)
struct DecXchg {
  unsigned long long sum[8];
  int mx[8];
  int ps[8];
};

__device__ __forceinline__ uint32_t map_peer(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ int ld_peer_s32(uint32_t addr) {
  int v;
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_peer_u64(uint32_t addr) {
  unsigned long long v;
  asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}

// reduce R (<= 8) per-thread values over the 256 threads of the block at once; result in every thread
template <int N, typename T, typename Op>
__device__ __forceinline__ void block_reduce256_n(T (&v)[N], int R, Op op, T* scratch /* [8 warps][8] */) {
#pragma unroll
  for (int r = 0; r < N; ++r) if (r < R) v[r] = warp_reduce(v[r], op);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int r = 0; r < N; ++r) if (r < R) scratch[(threadIdx.x >> 5) * 8 + r] = v[r];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < N; ++r) {
    if (r < R) {
      T t = scratch[r];
#pragma unroll
      for (int wi = 1; wi < 8; ++wi) t = op(t, scratch[wi * 8 + r]);
      v[r] = t;
    }
  }
}

// RMAX = query heads per cluster (a.RG <= RMAX).  The per-thread instruction stream of this kernel is a serial chain whose length
// grows with the number of query heads a thread carries (every loop below is unrolled over them); splitting the heads of a kv
// head over several clusters shortens the chain and multiplies the CTAs (the K / V slices are then read once per row group,
// from L2).
template <int HD, int RMAX>
__global__ void __launch_bounds__(256) qattn_decode_kernel(const AttnDecArgs a) {
  constexpr int WPR = HD / 4;                 // 32-bit words per head row
  constexpr int KS = 256 / WPR;               // key sub-slices of the P.V pass inside a CTA
  extern __shared__ __align__(16) uint8_t smem_dec[];
  const int Rtot = a.nh / a.nkv;
  const int R = a.RG;
  DecXchg* s_x = reinterpret_cast<DecXchg*>(smem_dec);                    // first: same offset in every CTA of the cluster
  unsigned long long* s_scr = reinterpret_cast<unsigned long long*>(s_x + 1);   // [8][8] reduction scratch
  uint32_t* s_q = reinterpret_cast<uint32_t*>(s_scr + 64);                // [RMAX][WPR] q codes of the new token
  uint32_t* s_knew = s_q + RMAX * WPR;                                    // [WPR] new k codes
  uint32_t* s_vnew = s_knew + WPR;                                        // [WPR] new v codes
  uint32_t* s_tab = s_vnew + WPR;                                         // [512]
  int* s_red = reinterpret_cast<int*>(s_tab + 512);                       // [KS][RMAX][HD] P.V partials
  int* s_rsq = s_red + KS * RMAX * HD;                                    // [RMAX] (+ [RMAX] = code sum of the new k row)
  int* s_cm = s_rsq + 2 * RMAX;                                           // [RMAX]
  float* s_den = reinterpret_cast<float*>(s_cm + RMAX);                   // [RMAX]
  uint16_t* s_c = reinterpret_cast<uint16_t*>(s_den + RMAX);              // [R][Tslice] score codes, then prob codes

  pdl_trigger();                                          // decode chain (common.cuh)
  for (int i = threadIdx.x; i < 512; i += 256) s_tab[i] = __ldg(a.lut + i);   // constant table: before the grid dependency resolves
  const int CS = a.CS;
  const int cr = (int)cluster_ctarank();
  const int grp = blockIdx.x / CS;
  const int ngrp = Rtot / R, rg = grp % ngrp;           // row group: query heads kvh * Rtot + rg * R + (0 .. R-1)
  const int b = (grp / ngrp) / a.nkv, kvh = (grp / ngrp) % a.nkv;
  const bool appender = cr == 0 && rg == 0;             // exactly one CTA per (sequence, kv head) appends the new k / v row
  // *pos_dev is advanced by the LAST kernel of a decode step, which has completed before any kernel of this step started
  const int pos = a.pos_dev ? *a.pos_dev : a.pos;
  // a replayed graph increments *pos_dev on the device: a step beyond the position the launch was sized for (cache capacity,
  // score slab, cos / sin tables) must not touch memory.  Every CTA of every cluster sees the same pos: uniform exit.
  if (pos < 0 || pos > a.pos_max) { pdl_wait(); return; }   // (a grid that skips the wait would let ITS dependents overtake the chain)
  const int Tk = pos + 1;
  const int per = (Tk + CS - 1) / CS;
  const int j_lo = min(Tk, cr * per), j_hi = min(Tk, j_lo + per);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* kcache = a.kc + (int64_t(b) * a.nkv + kvh) * a.Tmax * HD;
  uint8_t* vcache = a.vc + (int64_t(b) * a.nkv + kvh) * a.Tmax * HD;
  // the cache rows of earlier positions were written by earlier decode steps: pull this CTA's K and V slices towards L2 while
  // the predecessor (the QKV epilogue) is still running
  for (int64_t off = int64_t(j_lo) * HD + tid * 128; off < int64_t(min(j_hi, pos)) * HD; off += 256 * 128) {
    prefetch_l2(kcache + off);
    prefetch_l2(vcache + off);
  }
  pdl_wait();                                             // the qkv codes of this step come from earlier kernels of the chain
  const uint8_t* row = a.qkv + int64_t(b) * a.ldq;
  int32_t* rsk = a.rskc + (int64_t(b) * a.nkv + kvh) * a.Tmax;
  const int half = a.rot / 2;


  // ---- 1. RoPE + requant of the new token (every CTA of the cluster, redundantly: it is a few hundred elements);
  //         heads 0..R-1 = q, R = k, R+1 = v; one thread = 4 head dims; rank 0 appends k / v to the cache
  const QParam qq = make_qparam(a.sq, a.oq, 255.f), qk = make_qparam(a.sk, a.ok, 255.f), qv = make_qparam(a.sv, a.ov, 255.f);
  for (int i = tid; i < (R + 2) * WPR; i += 256) {
    const int hh = i / WPR, d = (i % WPR) * 4;
    const bool is_q = hh < R, is_k = hh == R;
    const uint8_t* src = is_q ? row + (kvh * Rtot + rg * R + hh) * HD : (is_k ? row + (a.nh + kvh) * HD : row + (a.nh + a.nkv + kvh) * HD);
    const float s_in = is_q ? a.sq_in : (is_k ? a.sk_in : a.sv_in), o_in = is_q ? a.oq_in : (is_k ? a.ok_in : a.ov_in);
    const QParam& qo = is_q ? qq : (is_k ? qk : qv);
    const uint32_t wx = *reinterpret_cast<const uint32_t*>(src + d);
    float x[4], o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = dequant(__uint2float_rn((wx >> (8 * j)) & 255u), s_in, o_in);
    if (hh <= R && d < a.rot) {
      const int dp = d < half ? d + half : d - half;
      const bool neg = d < half;
      const uint32_t wy = *reinterpret_cast<const uint32_t*>(src + dp);
      const float4 c = ldg4(a.cos + int64_t(pos) * a.rot + d), sn = ldg4(a.sin + int64_t(pos) * a.rot + d);
      const float cv[4] = {c.x, c.y, c.z, c.w}, sv[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float y = dequant(__uint2float_rn((wy >> (8 * j)) & 255u), s_in, o_in);
        o[j] = fadd(fmul(x[j], cv[j]), fmul(neg ? -y : y, sv[j]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = x[j];
    }
    uint32_t packed = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) packed |= (uint32_t)quant_int<true>(o[j], qo) << (8 * j);
    if (is_q) s_q[hh * WPR + (d >> 2)] = packed;
    else if (is_k) { s_knew[d >> 2] = packed; if (appender) *reinterpret_cast<uint32_t*>(kcache + int64_t(pos) * HD + d) = packed; }
    else { s_vnew[d >> 2] = packed; if (appender) *reinterpret_cast<uint32_t*>(vcache + int64_t(pos) * HD + d) = packed; }
  }
  __syncthreads();
  // code sums: q head r -> s_rsq[r], new k row -> s_rsq[RMAX] (and the cache, rank 0)
  for (int hh = warp; hh <= R; hh += 8) {
    int sum = 0;
    for (int wd = lane; wd < WPR; wd += 32) sum = (int)__dp4a(hh < R ? s_q[hh * WPR + wd] : s_knew[wd], 0x01010101u, (unsigned)sum);
    sum = warp_reduce(sum, OpSum());
    if (lane == 0) { if (hh < R) s_rsq[hh] = sum; else { s_rsq[RMAX] = sum; if (appender) rsk[pos] = sum; } }
  }
  __syncthreads();

  // ---- 2. scores of this CTA's key slice: thread = key, all R query rows at once
  const QParam qs = make_qparam(a.s_s, a.o_s, a.qmax_s);
  const QParam qp = make_qparam(a.s_p, 0.f, a.qmax_p);
  const QParam qo = make_qparam(a.s_out, a.o_out, 255.f);
  const int ioq = (int)a.oq, iok = (int)a.ok, iov = (int)a.ov;
  int mx[RMAX];
#pragma unroll
  for (int r = 0; r < RMAX; ++r) mx[r] = -1;
  for (int j = j_lo + tid; j < j_hi; j += 256) {
    const uint4* krow = j == pos ? reinterpret_cast<const uint4*>(s_knew) : reinterpret_cast<const uint4*>(kcache + int64_t(j) * HD);
    const int rskj = j == pos ? s_rsq[RMAX] : rsk[j];  // issued with the key row's loads, not after the dot products
    int acc[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) acc[r] = 0;
#pragma unroll 4
    for (int w4 = 0; w4 < WPR / 4; ++w4) {
      const uint4 kv = krow[w4];
      const uint32_t kw[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int r = 0; r < RMAX; ++r)
          if (r < R) acc[r] = (int)__dp4a(kw[e], s_q[r * WPR + w4 * 4 + e], (unsigned)acc[r]);
    }
    const int colc = -ioq * rskj + HD * ioq * iok;
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      if (r < R) {
        const int I = acc[r] + colc - iok * s_rsq[r];
        const int code = quant_int<true>(fmul(__int2float_rn(I), a.sqk), qs);
        s_c[r * a.Tslice + (j - j_lo)] = (uint16_t)code;
        mx[r] = max(mx[r], code);
      }
    }
  }
  block_reduce256_n(mx, R, OpMax(), reinterpret_cast<int*>(s_scr));
  if (tid < R) s_x->mx[tid] = mx[tid];
  cluster_sync_all();                                   // #1: slice maxima visible across the cluster
  if (tid < R) {
    int pm[8];                                          // all peers' loads in flight before the first use (DSMEM latency once)
#pragma unroll
    for (int p = 0; p < 8; ++p) pm[p] = p < CS ? ld_peer_s32(map_peer(&s_x->mx[tid], p)) : -1;
    int m = -1;
#pragma unroll
    for (int p = 0; p < 8; ++p) m = max(m, pm[p]);
    s_cm[tid] = m;
  }
  __syncthreads();

  // ---- 3. exact softmax on the codes
  auto exp_tab = [&](int k) -> uint32_t {
    const uint32_t ea = s_tab[(k >> 8) & 255], eb = s_tab[256 + (k & 255)];
    return (uint32_t)(((unsigned long long)ea * eb) >> 31);
  };
  const int nloc = j_hi - j_lo;
  {
    unsigned long long sum[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) sum[r] = 0;
    for (int jj = tid; jj < nloc; jj += 256) {
#pragma unroll
      for (int r = 0; r < RMAX; ++r) if (r < R) sum[r] += exp_tab(s_cm[r] - (int)s_c[r * a.Tslice + jj]);
    }
    block_reduce256_n(sum, R, OpSum(), s_scr);
    if (tid < R) s_x->sum[tid] = sum[tid];
  }
  cluster_sync_all();                                   // #2: slice sums visible
  if (tid < R) {
    unsigned long long pt[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) pt[p] = p < CS ? ld_peer_u64(map_peer(&s_x->sum[tid], p)) : 0ull;
    unsigned long long t = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p) t += pt[p];
    s_den[tid] = __ull2float_rn(t);
  }
  __syncthreads();
  {
    int ps[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) ps[r] = 0;
    for (int jj = tid; jj < nloc; jj += 256) {
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        if (r < R) {
          const float den = s_den[r];
          const uint32_t e = exp_tab(s_cm[r] - (int)s_c[r * a.Tslice + jj]);
          const float pr = div_rn<true>(__uint2float_rn(e), den, __frcp_rn(den));
          const int cp = quant_int<true>(pr, qp);
          s_c[r * a.Tslice + jj] = (uint16_t)cp;
          ps[r] += cp;
        }
      }
    }
    block_reduce256_n(ps, R, OpSum(), reinterpret_cast<int*>(s_scr));   // its barriers also publish the prob codes
    if (tid < R) s_x->ps[tid] = ps[tid];
  }

  // ---- 4. P.V over the slice: thread = (key sub-slice, 4 head dims); four keys in flight per thread
  {
    const int slice = tid / WPR, dq = tid % WPR;
    int pv[RMAX][4];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) { pv[r][0] = pv[r][1] = pv[r][2] = pv[r][3] = 0; }
    auto vword = [&](int jj) -> uint32_t {
      const int j = j_lo + jj;
      return j == pos ? s_vnew[dq] : *reinterpret_cast<const uint32_t*>(vcache + int64_t(j) * HD + dq * 4);
    };
    auto fma_key = [&](int jj, uint32_t vw) {
      const int v0 = vw & 255u, v1 = (vw >> 8) & 255u, v2 = (vw >> 16) & 255u, v3 = vw >> 24;
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        if (r < R) {
          const int c = s_c[r * a.Tslice + jj];
          pv[r][0] += c * v0; pv[r][1] += c * v1; pv[r][2] += c * v2; pv[r][3] += c * v3;
        }
      }
    };
    int jj = slice;
    for (; jj + 3 * KS < nloc; jj += 4 * KS) {
      const uint32_t w0 = vword(jj), w1 = vword(jj + KS), w2 = vword(jj + 2 * KS), w3 = vword(jj + 3 * KS);
      fma_key(jj, w0); fma_key(jj + KS, w1); fma_key(jj + 2 * KS, w2); fma_key(jj + 3 * KS, w3);
    }
    for (; jj < nloc; jj += KS) fma_key(jj, vword(jj));
#pragma unroll
    for (int r = 0; r < RMAX; ++r)
      if (r < R) *reinterpret_cast<int4*>(s_red + (slice * RMAX + r) * HD + dq * 4) = make_int4(pv[r][0], pv[r][1], pv[r][2], pv[r][3]);
  }
  __syncthreads();
  // fold the KS sub-slices into sub-slice 0 (what the cluster leader reads)
  for (int i = tid; i < R * HD; i += 256) {
    const int r = i / HD, d = i % HD;
    int A = 0;
    for (int sl = 0; sl < KS; ++sl) A += s_red[(sl * RMAX + r) * HD + d];
    s_red[r * HD + d] = A;                               // (sub-slice 0, row r) is only read by this thread
  }
  cluster_sync_all();                                   // #3: slice partials and prob-code sums visible
  if (cr == 0) {
    int csum = 0;
    const int ldo = a.nh * HD;
    for (int i = tid; i < R * HD; i += 256) {
      const int r = i / HD, d = i % HD;
      int pa[8], pp[8];                                 // 2 * CS remote loads in flight, then the integer sums
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        pa[p] = p < CS ? ld_peer_s32(map_peer(s_red + r * HD + d, p)) : 0;
        pp[p] = p < CS ? ld_peer_s32(map_peer(&s_x->ps[r], p)) : 0;
      }
      int A = 0, psum = 0;
#pragma unroll
      for (int p = 0; p < 8; ++p) { A += pa[p]; psum += pp[p]; }
      A -= iov * psum;
      const int code = quant_int<true>(fmul(__int2float_rn(A), a.spv), qo);
      a.out[int64_t(b) * ldo + (kvh * Rtot + rg * R + r) * HD + d] = (uint8_t)code;
      csum += code;
    }
    if (a.rowsum_out) {
      int one[1] = {csum};
      block_reduce256_n(one, 1, OpSum(), reinterpret_cast<int*>(s_scr));
      if (tid == 0) atomicAdd(a.rowsum_out + b, one[0]);
    }
  }
  cluster_sync_all();                                   // #4: peers keep their shared memory alive until the leader has read it
}

static size_t attn_dec_smem(int hd, int rmax, int R, int Tslice) {
  const int WPR = hd / 4, KS = 256 / WPR;
  return sizeof(DecXchg) + 64 * 8 + size_t(rmax + 2) * WPR * 4 + 512 * 4 + size_t(KS) * rmax * hd * 4 + 4 * rmax * 4 + size_t(R) * Tslice * 2 + 16;
}

// =====================================================================================================================
// fp32 lm_head for the decode step: logits[b, v] = sum_k x[b, k] * W[v, k] for B <= 16 rows.  The reference keeps lm_head
// unquantised (qm:843-845); a library SGEMM moves its 0.26-2.1 GB of weights at 2.4-2.6 TB/s for these shapes, this kernel
// streams every weight row once with 16-byte loads (one warp per vocabulary row, x rows via L1) and is HBM-bound.
// Summation order is fixed (lane-strided partials, then a shuffle tree): run-to-run deterministic.
// =====================================================================================================================
// One warp owns kFgRows consecutive vocabulary rows so that every x vector fetched from L1 feeds kFgRows x 4 FMAs
// (a single row per warp re-reads x once per weight vector and is L1-bandwidth bound at 3.0-3.5 TB/s).
constexpr int kFgRows = 4;
template <int BMAX>
__global__ void __launch_bounds__(256) fgemv_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ out,
                                                    int B, int V, int K) {
  pdl_trigger();
  pdl_wait();                                  // x comes from the final norm before this launch
  const int lane = threadIdx.x & 31;
  const int v0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * kFgRows;
  if (v0 >= V) return;
  float acc[kFgRows][BMAX];
#pragma unroll
  for (int r = 0; r < kFgRows; ++r)
#pragma unroll
    for (int b = 0; b < BMAX; ++b) acc[r][b] = 0.f;
  const float* wr[kFgRows];
#pragma unroll
  for (int r = 0; r < kFgRows; ++r) wr[r] = w + int64_t(min(v0 + r, V - 1)) * K;      // tail rows alias the last row (not stored)
  // two k-steps (2 x kFgRows 16-byte weight loads per lane) in flight: ~65 KB outstanding per SM, what 6.5 TB/s needs
  auto fma_step = [&](const float4 (&wv)[kFgRows], int k) {
    float4 xv[BMAX];                             // all x vectors of the step first: one L1 round trip, not BMAX in a chain
#pragma unroll
    for (int b = 0; b < BMAX; ++b) xv[b] = ldg4(x + int64_t(min(b, B - 1)) * K + k);
#pragma unroll
    for (int b = 0; b < BMAX; ++b) {
#pragma unroll
      for (int r = 0; r < kFgRows; ++r) {
        acc[r][b] = __fmaf_rn(wv[r].x, xv[b].x, acc[r][b]); acc[r][b] = __fmaf_rn(wv[r].y, xv[b].y, acc[r][b]);
        acc[r][b] = __fmaf_rn(wv[r].z, xv[b].z, acc[r][b]); acc[r][b] = __fmaf_rn(wv[r].w, xv[b].w, acc[r][b]);
      }
    }
  };
  int k = lane * 4;
  for (; k + 128 < K; k += 256) {
    float4 w0[kFgRows], w1[kFgRows];
#pragma unroll
    for (int r = 0; r < kFgRows; ++r) { w0[r] = ldg4_stream(wr[r] + k); w1[r] = ldg4_stream(wr[r] + k + 128); }
    fma_step(w0, k);
    fma_step(w1, k + 128);
  }
  for (; k < K; k += 128) {
    float4 w0[kFgRows];
#pragma unroll
    for (int r = 0; r < kFgRows; ++r) w0[r] = ldg4_stream(wr[r] + k);
    fma_step(w0, k);
  }
#pragma unroll
  for (int r = 0; r < kFgRows; ++r) {
#pragma unroll
    for (int b = 0; b < BMAX; ++b) {
      if (b < B) {
        float t = acc[r][b];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (lane == 0 && v0 + r < V) out[int64_t(b) * V + v0 + r] = t;
      }
    }
  }
}

// =====================================================================================================================
// mq_unpack4: packed 4-bit weight codes (two per byte, low nibble = even column; mq_wprep_fwd pack4) -> one int8 / uint8
// code per byte, the operand format of the tcgen05 kind::i8 GEMMs.  W4A8 weights stay packed in HBM (0.48 GB for
// TinyLlama-1.1B); a layer's matrices are expanded right before their GEMM into a scratch buffer that fits the 126 MB L2,
// so the GEMM's operand traffic is served from L2 and the expansion costs one pass over the packed bytes.
// =====================================================================================================================
__global__ void __launch_bounds__(256) unpack4_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t n16, int is_signed) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n16) return;
  const uint4 p = __ldg(src + i);
  const uint32_t pw[4] = {p.x, p.y, p.z, p.w};
  uint32_t o[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t lo = pw[j] & 0x0F0F0F0Fu, hi = (pw[j] >> 4) & 0x0F0F0F0Fu;
    uint32_t a = __byte_perm(lo, hi, 0x5140), b = __byte_perm(lo, hi, 0x7362);     // l0 h0 l1 h1 | l2 h2 l3 h3
    if (is_signed) {                                                              // 4-bit two's complement -> int8
      a |= (a & 0x08080808u) * 0x1Eu;
      b |= (b & 0x08080808u) * 0x1Eu;
    }
    o[2 * j] = a; o[2 * j + 1] = b;
  }
  dst[2 * i] = make_uint4(o[0], o[1], o[2], o[3]);
  dst[2 * i + 1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (same helper as qgemm.cu, local copy of the lookup)
typedef CUresult (*PFN_encodeTiledGv)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiledGv gv_encode() {
  static PFN_encodeTiledGv fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiledGv>(p);
  }
  return fn;
}
static bool gv_tmap(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int box_rows) {
  PFN_encodeTiledGv enc = gv_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols};
  cuuint32_t box[2] = {128u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BP>
static int launch_qgemv(Ctx* c, const void* x, int x_signed, const void* w, int w_signed, const QGemvArgs& args, cudaStream_t st) {
  using L = GvSmem<BP>;
  CUtensorMap tw, tx;
  if (!gv_tmap(&tw, w, args.N, args.K, kGvBM) || !gv_tmap(&tx, x, args.B, args.K, BP))
    return fail(c, MQ_RUNTIME_ERROR, "cuTensorMapEncodeTiled failed (pointers must be 16B aligned, K a multiple of 16)");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(qgemv_kernel<BP>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return fail(c, MQ_RUNTIME_ERROR, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    attr_set = true;
  }
  const uint32_t idesc = make_idesc(2u, w_signed ? 1u : 0u, x_signed ? 1u : 0u, 0u, 0u, kGvBM, BP);
  const int n_tiles = (args.N + kGvBM - 1) / kGvBM;
  QGemvArgs args2 = args;
  const char* wlo = static_cast<const char*>(w);
  args2.w_const = !(c->unpack_lo && wlo < c->unpack_hi && wlo + size_t(args.N) * args.K > c->unpack_lo);
  cudaError_t le = launch_pdl(qgemv_kernel<BP>, dim3(n_tiles * args.ksplit), dim3(192), (size_t)L::kTotal, st, tw, tx, args2, idesc);
  if (le != cudaSuccess) return fail(c, MQ_RUNTIME_ERROR, std::string("mq_qgemv launch: ") + cudaGetErrorString(le));
  return check_launch(c, "mq_qgemv");
}

}  // namespace mq

using namespace mq;

extern "C" {

static int qgemv_dispatch(Ctx* c, const void* x_codes, int x_signed, const void* w_codes, int w_signed, QGemvArgs& args, cudaStream_t st) {
  const int k_iters = (args.K + kGvBK - 1) / kGvBK;
  const int n_tiles = (args.N + kGvBM - 1) / kGvBM;
  if (args.ksplit <= 0) {
    // fill the machine (two CTAs per SM) but keep at least two 128-byte K slices per CTA
    // (MQB200_GEMV_FILL / MQB200_GEMV_MINK override the two numbers: A/B measurements)
    static int fill = -1, mink = -1;
    if (fill < 0) { const char* e = getenv("MQB200_GEMV_FILL"); fill = e && atoi(e) > 0 ? atoi(e) : 2; }
    if (mink < 0) { const char* e = getenv("MQB200_GEMV_MINK"); mink = e && atoi(e) > 0 ? atoi(e) : 2; }
    args.ksplit = std::max(1, std::min(fill * c->sm_count / n_tiles, std::max(1, k_iters / mink)));
  }
  args.ksplit = std::min(args.ksplit, k_iters);
  if (args.B <= 16) return launch_qgemv<16>(c, x_codes, x_signed, w_codes, w_signed, args, st);
  if (args.B <= 32) return launch_qgemv<32>(c, x_codes, x_signed, w_codes, w_signed, args, st);
  if (args.B <= 64) return launch_qgemv<64>(c, x_codes, x_signed, w_codes, w_signed, args, st);
  return launch_qgemv<128>(c, x_codes, x_signed, w_codes, w_signed, args, st);
}

static int qgemv_check(Ctx* c, const void* x_codes, const void* w_codes, const int32_t* acc, int B, int N, int K, int ldacc) {
  MQ_REQUIRE(c, x_codes && w_codes && acc && B > 0 && N > 0 && K > 0, "null operand or empty problem");
  MQ_REQUIRE(c, B <= 128, "the skinny GEMM covers up to 128 rows; use mq_qgemm above");
  MQ_REQUIRE(c, K % 16 == 0 && ldacc >= N, "K must be a multiple of 16 and ldacc >= N");
  MQ_REQUIRE(c, (reinterpret_cast<uintptr_t>(x_codes) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_codes) & 15) == 0, "operands must be 16-byte aligned");
  return MQ_NO_ERROR;
}

static int qgemv_epi_check(Ctx* c, const int32_t* acc, int ldacc, int B, int N, const int32_t* rowsum, const float* sxw, const int32_t* ow,
                           const int32_t* c0, int mode, const float* so, const float* oo, float qmax, const uint8_t* out, int64_t ldo,
                           const int32_t* rowsum_out, const float* lut, float qmax2, const float* resid, int qgroup, const int32_t* zero_out) {
  MQ_REQUIRE(c, zero_out != rowsum && (zero_out == nullptr || zero_out != rowsum_out), "zero_out must not alias rowsum / rowsum_out");
  MQ_REQUIRE(c, acc && rowsum && sxw && ow && c0 && so && oo && B > 0 && N > 0, "null pointer or empty problem");
  MQ_REQUIRE(c, mode >= GV_QUANT && mode <= GV_RESID, "mode must be 0 (QUANT), 1 (ACTMUL) or 2 (RESID)");
  MQ_REQUIRE(c, N % 4 == 0 && ldacc % 4 == 0 && (reinterpret_cast<uintptr_t>(acc) & 15) == 0, "N, ldacc must be multiples of 4, acc 16-byte aligned");
  MQ_REQUIRE(c, qgroup > 0 && qgroup % 4 == 0, "qgroup must be a positive multiple of 4");
  MQ_REQUIRE(c, mode != GV_QUANT || (out && ldo % 4 == 0 && qmax <= 255.f), "QUANT needs out, ldo % 4 == 0, 8-bit codes");
  MQ_REQUIRE(c, mode != GV_ACTMUL || (out && lut && N % 256 == 0 && 128 % qgroup == 0 && ldo % 4 == 0), "ACTMUL needs out/lut, N % 256 == 0, qgroup | 128");
  MQ_REQUIRE(c, mode != GV_RESID || (resid && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(resid) & 15) == 0), "RESID needs a 16-byte aligned resid, ldo % 4 == 0");
  MQ_REQUIRE(c, qmax < 4194304.f && qmax2 < 4194304.f, "qmax must be below 2^22");
  return MQ_NO_ERROR;
}

int mq_qgemv(void* ctx, const void* x_codes, int x_signed, const void* w_codes, int w_signed, int B, int N, int K, int32_t* acc,
             int ldacc, int ksplit, void* stream) {
  MQ_CTX(c, ctx);
  if (int rc = qgemv_check(c, x_codes, w_codes, acc, B, N, K, ldacc)) return rc;
  QGemvArgs args{};
  args.B = B; args.N = N; args.K = K; args.acc = acc; args.ldacc = ldacc; args.ksplit = ksplit; args.mode = GV_NONE; args.counters = nullptr;
  return qgemv_dispatch(c, x_codes, x_signed, w_codes, w_signed, args, (cudaStream_t)stream);
}

int mq_qgemv_fused(void* ctx, const void* x_codes, int x_signed, const void* w_codes, int w_signed, int B, int N, int K, int32_t* acc,
                   int ldacc, int ksplit, const int32_t* rowsum, const float* sxw, const int32_t* ow, const int32_t* c0, const float* bias,
                   int mode, const float* so, const float* oo, float qmax, uint8_t* out, int64_t ldo, int32_t* rowsum_out, const float* lut,
                   float s2, float o2, float qmax2, float* resid, int qgroup, int32_t* zero_out, void* stream) {
  MQ_CTX(c, ctx);
  if (int rc = qgemv_check(c, x_codes, w_codes, acc, B, N, K, ldacc)) return rc;
  if (int rc = qgemv_epi_check(c, acc, ldacc, B, N, rowsum, sxw, ow, c0, mode, so, oo, qmax, out, ldo, rowsum_out, lut, qmax2, resid, qgroup, zero_out)) return rc;
  MQ_REQUIRE(c, c->counters && (N + 127) / 128 <= c->n_counters, "too many column groups for the arrival counters");
  QGemvArgs args{};
  args.B = B; args.N = N; args.K = K; args.acc = acc; args.ldacc = ldacc; args.ksplit = ksplit; args.mode = mode; args.counters = c->counters;
  args.epi = GvEpiArgs{B, N, acc, ldacc, rowsum, sxw, ow, c0, bias, so, oo, qgroup, qmax, out, ldo, rowsum_out, lut, s2, o2, qmax2, resid, zero_out};
  return qgemv_dispatch(c, x_codes, x_signed, w_codes, w_signed, args, (cudaStream_t)stream);
}

int mq_qgemv_epilogue(void* ctx, int32_t* acc, int ldacc, int B, int N, const int32_t* rowsum, const float* sxw, const int32_t* ow,
                      const int32_t* c0, const float* bias, int mode, const float* so, const float* oo, float qmax, uint8_t* out,
                      int64_t ldo, int32_t* rowsum_out, const float* lut, float s2, float o2, float qmax2, float* resid, int qgroup,
                      int32_t* zero_out, void* stream) {
  MQ_CTX(c, ctx);
  if (int rc = qgemv_epi_check(c, acc, ldacc, B, N, rowsum, sxw, ow, c0, mode, so, oo, qmax, out, ldo, rowsum_out, lut, qmax2, resid, qgroup, zero_out)) return rc;
  GvEpiArgs a{B, N, acc, ldacc, rowsum, sxw, ow, c0, bias, so, oo, qgroup, qmax, out, ldo, rowsum_out, lut, s2, o2, qmax2, resid, zero_out};
  const int NO = mode == GV_ACTMUL ? N / 2 : N;
  dim3 grid((NO / 4 + 127) / 128, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == GV_QUANT) launch_pdl(qgemv_epi_kernel<GV_QUANT>, grid, dim3(128), 0, st, a);
  else if (mode == GV_ACTMUL) launch_pdl(qgemv_epi_kernel<GV_ACTMUL>, grid, dim3(128), 0, st, a);
  else launch_pdl(qgemv_epi_kernel<GV_RESID>, grid, dim3(128), 0, st, a);
  return check_launch(c, "mq_qgemv_epilogue");
}

int mq_unpack4(void* ctx, const uint8_t* packed, int64_t n_codes, int is_signed, void* out, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, packed && out && n_codes > 0, "null pointer or empty input");
  MQ_REQUIRE(c, n_codes % 32 == 0, "the number of codes must be a multiple of 32");
  MQ_REQUIRE(c, (reinterpret_cast<uintptr_t>(packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "buffers must be 16-byte aligned");
  const int64_t n16 = n_codes / 32;
  c->unpack_lo = static_cast<const char*>(out);
  c->unpack_hi = c->unpack_lo + n_codes;
  unpack4_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(packed),
                                                                                 reinterpret_cast<uint4*>(out), n16, is_signed);
  return check_launch(c, "mq_unpack4");
}

int mq_fgemv(void* ctx, const float* x, const float* w, float* out, int B, int V, int K, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, x && w && out && B > 0 && V > 0 && K > 0, "null pointer or empty problem");
  MQ_REQUIRE(c, B <= 16 && K % 4 == 0, "the decode lm_head kernel covers up to 16 rows, K a multiple of 4");
  MQ_REQUIRE(c, (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0, "x and w must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((V + 8 * kFgRows - 1) / (8 * kFgRows));
  if (B <= 8) launch_pdl(fgemv_kernel<8>, dim3(grid), dim3(256), 0, st, x, w, out, B, V, K);
  else launch_pdl(fgemv_kernel<16>, dim3(grid), dim3(256), 0, st, x, w, out, B, V, K);
  return check_launch(c, "mq_fgemv");
}

int mq_qattn_decode(void* ctx, const uint8_t* qkv, int ldq, int B, int nh, int nkv, int hd, int rot, int Tmax, int pos,
                    const int* pos_dev, int pos_bound, const float* rope_in_qparams, const float* rope_out_qparams, const float* cos,
                    const float* sin, uint8_t* k_cache, uint8_t* v_cache, int32_t* rsk_cache, const float* qparams,
                    const uint32_t* lut, uint8_t* out, int32_t* rowsum_out, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, qkv && rope_in_qparams && rope_out_qparams && cos && sin && k_cache && v_cache && rsk_cache && qparams && lut && out, "null pointer");
  MQ_REQUIRE(c, B > 0 && nh > 0 && nkv > 0 && nh % nkv == 0 && nh / nkv <= 8, "bad head counts (at most 8 query heads per kv head)");
  MQ_REQUIRE(c, hd == 32 || hd == 64 || hd == 128 || hd == 256, "head_dim must be 32, 64, 128 or 256");
  MQ_REQUIRE(c, rot >= 0 && rot <= hd && rot % 8 == 0 && ldq % 4 == 0, "rotary width must be a multiple of 8, ldq of 4");
  if (!pos_dev) pos_bound = pos;
  MQ_REQUIRE(c, pos_bound >= 0 && pos_bound < Tmax && (pos_dev || (pos >= 0 && pos < Tmax)), "position outside the cache");
  MQ_REQUIRE(c, (reinterpret_cast<uintptr_t>(qkv) & 3) == 0 && (reinterpret_cast<uintptr_t>(k_cache) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(v_cache) & 15) == 0 && (reinterpret_cast<uintptr_t>(cos) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(sin) & 15) == 0, "qkv must be 4-byte, caches and cos/sin 16-byte aligned");
  AttnDecArgs a;
  a.qkv = qkv; a.ldq = ldq; a.B = B; a.nh = nh; a.nkv = nkv; a.hd = hd; a.rot = rot; a.Tmax = Tmax; a.pos = pos; a.pos_dev = pos_dev; a.pos_max = pos_bound;
  a.sq_in = rope_in_qparams[0]; a.oq_in = rope_in_qparams[1]; a.sk_in = rope_in_qparams[2]; a.ok_in = rope_in_qparams[3];
  a.sv_in = rope_in_qparams[4]; a.ov_in = rope_in_qparams[5];
  a.sq = rope_out_qparams[0]; a.oq = rope_out_qparams[1]; a.sk = rope_out_qparams[2]; a.ok = rope_out_qparams[3];
  a.sv = rope_out_qparams[4]; a.ov = rope_out_qparams[5];
  MQ_REQUIRE(c, a.oq == qparams[0] && a.ok == qparams[1] && a.ov == qparams[2], "rope output offsets must match the attention zero points");
  a.cos = cos; a.sin = sin; a.kc = k_cache; a.vc = v_cache; a.rskc = rsk_cache;
  a.sqk = qparams[3]; a.s_s = qparams[4]; a.o_s = qparams[5]; a.qmax_s = qparams[6]; a.s_p = qparams[7]; a.qmax_p = qparams[8];
  a.spv = qparams[9]; a.s_out = qparams[10]; a.o_out = qparams[11];
  MQ_REQUIRE(c, a.qmax_s <= 65535.f && a.qmax_p <= 65535.f, "score / probability codes are at most 16 bit");
  MQ_REQUIRE(c, a.oq == rintf(a.oq) && a.ok == rintf(a.ok) && a.ov == rintf(a.ov) && a.o_s == rintf(a.o_s) && a.o_out == rintf(a.o_out),
             "integer engine kernels need integral offsets (qm:60)");
  a.lut = lut; a.out = out; a.rowsum_out = rowsum_out;
  // row groups: RG query heads of a kv head per cluster (MQB200_DEC_RG overrides; measured in profiles/r2_decode_kernels.txt)
  const int Rtot = nh / nkv;
  int RG = Rtot >= 2 ? 2 : 1;
  { const char* e = getenv("MQB200_DEC_RG"); if (e && atoi(e) >= 1 && atoi(e) <= Rtot && Rtot % atoi(e) == 0) RG = atoi(e); }
  while (Rtot % RG) --RG;
  const int rmax = RG <= 1 ? 1 : (RG <= 2 ? 2 : (RG <= 4 ? 4 : 8));
  a.RG = RG;
  // key slices: one cluster per (sequence, kv head, row group); at least 256 keys per slice
  int CS = 1;
  int cs_max = 8;
  { const char* e = getenv("MQB200_DEC_CS"); if (e && atoi(e) >= 1 && atoi(e) <= 8) cs_max = atoi(e); }   // A/B measurements
  // measured (profiles/r1d_decode_kernels.md): slices of >= 256 keys; 1024 keys -> 4 CTAs (22.8 us), 2048 keys -> 8 (50.3 us)
  while (CS < cs_max && B * nkv * CS * 2 <= 2 * c->sm_count && (pos_bound + 1) / (CS * 2) >= 256) CS *= 2;
  a.CS = CS;
  a.Tslice = (((pos_bound + 1) + CS - 1) / CS + 7) / 8 * 8;
  const size_t smem = attn_dec_smem(hd, rmax, RG, a.Tslice);
  MQ_REQUIRE(c, smem <= 227 * 1024, "sequence too long for the decode attention score slab");
  cudaStream_t st = (cudaStream_t)stream;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B * nkv * (Rtot / RG) * CS); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t le = cudaSuccess;
#define MQ_DEC2(HDV, RM)                                                                                                    \
  {                                                                                                                         \
    static size_t attr_smem = 48 * 1024;                                                                                    \
    if (smem > attr_smem) {                                                                                                 \
      cudaError_t e = cudaFuncSetAttribute(qattn_decode_kernel<HDV, RM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return fail(c, MQ_RUNTIME_ERROR, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));  \
      attr_smem = smem;                                                                                                     \
    }                                                                                                                       \
    le = cudaLaunchKernelEx(&cfg, qattn_decode_kernel<HDV, RM>, a);                                                         \
  }
#define MQ_DEC(HDV) { if (rmax == 1) MQ_DEC2(HDV, 1) else if (rmax == 2) MQ_DEC2(HDV, 2) else if (rmax == 4) MQ_DEC2(HDV, 4) else MQ_DEC2(HDV, 8) }
  if (hd == 32) MQ_DEC(32) else if (hd == 64) MQ_DEC(64) else if (hd == 128) MQ_DEC(128) else MQ_DEC(256)
#undef MQ_DEC2
#undef MQ_DEC
  if (le != cudaSuccess) return fail(c, MQ_RUNTIME_ERROR, std::string("mq_qattn_decode launch: ") + cudaGetErrorString(le));
  return check_launch(c, "mq_qattn_decode");
}

}  // extern "C"
