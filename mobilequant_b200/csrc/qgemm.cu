// K3/K7: statically-quantised integer GEMM on the 5th-gen tensor cores.
//
//   acc[m,n] = sum_k A[m,k] * B[n,k]          A: activation codes [M,K] (u8/s8), B: weight codes [N,K] (u8/s8), s32 in TMEM
//   I        = acc - ow[n]*rowsum[m] + c0[n]   == sum_k (A-ox)(B-ow[n])  with c0[n] = K*ox*ow[n] - ox*colsum[n]   (exact)
//   y        = float(I) * sxw[n] (+ bias[n])   sxw[n] = s_x * s_w[n]
//   epilogue = output Quantizer (qm:286-287) -> u8/u16 codes | SiLU/GELU-LUT * gate -> u8 | 16-bit requant + residual add
//
// This is what QLinear.forward (qm:341-358) computes on de-quantised fp32 tensors, restated on the integer codes; the
// CPU restatement the kernel is bit-exact against is oracle/int_ref.py:qlinear_int.
//
// Structure (one CTA per SM, persistent over 128 x 256 output tiles):
//   warp 0      TMA producer   : cp.async.bulk.tensor 128B-swizzled A/B k-slices into a kStages-deep smem ring
//   warp 1      MMA issuer     : one elected lane issues tcgen05.mma.kind::i8 (M=128, N=256, K=32) into TMEM; owns TMEM alloc
//   warps 2..   epilogue       : tcgen05.ld the s32 tile (thread = output row, 32 columns at a time), requantise with the
//                                branch-free exact division of common.cuh (no XU pipe, no slow path), store
//   TMEM holds two 256-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
// The residual epilogue never loads the fp32 stream into the SM: each warp stages a 32x32 tile of de-quantised values in
// 128B-swizzled smem and hands it to the TMA as an L2-side reduce-add (cp.reduce.async.bulk.tensor .add).
#include "common.cuh"
#include "tc_common.cuh"
#include "ctx.h"
#include "tc_host.h"
#include <string>
#include <cstdlib>

namespace mq {
using namespace tc;

constexpr int kBM = 128;          // UMMA M
constexpr int kBN = 256;          // UMMA N
constexpr int kBK = 128;          // bytes of K per stage == one 128B swizzle row
constexpr int kUmmaK = 32;        // 8-bit operands: 32 elements per MMA
// NE = epilogue warps (multiple of 4: one TMEM lane quarter each): 8 (32-column chunks) or 16 (16-column chunks so
// that 576 threads fit the register file).

enum { EPI_QUANT = 0, EPI_ACTMUL = 1, EPI_RESID = 2, EPI_F32 = 3, EPI_I32 = 4 };
enum { CP_NEGOW = 0, CP_C0, CP_SXW, CP_BIAS, CP_Q, CP_COUNT };   // CP_Q: output quantizer (scale [0..8), offset [8..16)) per 32-column chunk

struct QGemmArgs {
  int M, N, K;
  const int32_t* rowsum;   // [M]
  const float* sxw;        // [N]
  const int32_t* ow;       // [N]
  const int32_t* c0;       // [N]
  const float* bias;       // [N] or null
  const float* so;         // [ceil(N/qgroup)] output quantizer scale per group of qgroup columns
  const float* oo;         // [ceil(N/qgroup)] output quantizer offset (integral)
  int qgroup;              // multiple of 32 (the epilogue's column chunk); ACTMUL: divides 128
  float qmax;              // output quantizer qmax (qmin = 0: activations are asymmetric)
  int out_bits;            // 8 or 16 (EPI_QUANT)
  void* out;               // codes / fp32 / int32
  int64_t ldo;
  int32_t* rowsum_out;     // [M] atomically accumulated sum of the emitted codes (or null)
  const float* lut;        // [256] EPI_ACTMUL: act(w1 code) as fp32 (QSiLU/QGELU folded, qm:739-753)
  float s2, o2, qmax2;     // EPI_ACTMUL: w2.input_quantizer
  int split_f;             // tail wave: the tiles of the last, partial wave are split into split_f column slices (1, 2 or 4)
  int dbg;                 // measurement only (MQ_QGEMM_DBG): 1 = epilogue releases the accumulator without reading it
};

// CL = CTAs per cluster.
// CL 2 (PAIR): two CTAs of a cluster (one TPC) share a 256 x 256 output tile: each loads its own 128 A rows and HALF of
// the B rows, one tcgen05.mma.cta_group::2 (M = 256) issued by the leader consumes both halves -> B traffic from L2 and
// the B shared-memory footprint per CTA are halved, which buys two more pipeline stages.
// CL 4 (QUAD): two such pairs stacked along M (a 512 x 256 super-tile) read the SAME B tile: each CTA fetches only a
// quarter of it (64 rows) and the TMA multicasts the box to the CTA of the other pair that needs the same half -> per
// CTA and k-slice 16 KB of A + 8 KB of B leave the L2 instead of 16 + 32 (single) / 16 + 16 (pair).  At 128 MMA cycles
// per 32 bytes of K the single-CTA kernel asks the L2 for 96 B/clk/SM, more than twice what it sustains chip-wide
// (~43-50 B/clk/SM): the operand traffic, not the tensor pipe, was the ceiling of the first two variants.
// W4: the weight operand arrives PACKED (two 4-bit codes per byte, [N, K/2], unsigned nibbles): the TMA stages 64-byte packed
// rows into a small ring, four extra warps expand the nibbles into the 128B-swizzled operand rows the MMA reads (the fused
// int4 x int8 matmul of BASELINE config 3: the weights never exist one-code-per-byte outside shared memory).  CL 1 only.
template <int MODE, int CL, int NE, bool W4 = false>
struct SmemLayout {
  static constexpr bool PAIR = CL >= 2;
  static_assert(!W4 || CL == 1, "the packed-weight path is built for the single-CTA variant");
  static constexpr int kPStages = W4 ? 2 : 0;                // packed-B ring
  static constexpr int kPBytes = 256 * 64;                   // 256 rows x 64 packed bytes (one k-slice of 128 codes)
  static constexpr int kCW = NE == 16 ? 16 : 32;             // accumulator columns per tcgen05.ld / per staging tile
  static constexpr int kOutTile = 32 * kCW * 4;              // RESID: one fp32 staging tile (32 rows x kCW columns)
  static constexpr int kStages = W4 ? (MODE == EPI_RESID ? 2 : 3) : (PAIR ? (MODE == EPI_RESID ? 4 : 6) : (MODE == EPI_RESID ? 3 : 4));
  static constexpr int kABytes = kBM * kBK;
  static constexpr int kBRows = PAIR ? kBN / 2 : kBN;       // B rows resident per CTA
  static constexpr int kBBoxRows = CL == 4 ? kBRows / 2 : kBRows;   // B rows per TMA box
  static constexpr int kBBytes = kBRows * kBK;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kPackOff = kStages * kStageBytes;                 // [kPStages][kPBytes] packed weight k-slices (W4)
  static constexpr int kColpOff = kPackOff + kPStages * kPBytes;         // 2 x CP_COUNT x 256 x 4 B
  static constexpr int kLutOff = kColpOff + 2 * CP_COUNT * kBN * 4;      // 256 floats
  static constexpr int kOutOff = kLutOff + 1024;                         // RESID: per-warp 32x32 fp32 staging tiles
  static constexpr int kOutBufs = 2;                                     // staging tiles per warp
  static constexpr int kOutBytes = MODE == EPI_RESID ? NE * kOutBufs * kOutTile : 0;
  static constexpr int kBarOff = kOutOff + kOutBytes;
  static constexpr int kTotal = kBarOff + 256;
  static_assert(kOutOff % 1024 == 0 && kPackOff % 1024 == 0, "staging tiles must keep the swizzle phase");
  static_assert(kTotal <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
};

constexpr int kUnpackWarps = 4;

template <int MODE, int CL, int NE, bool W4 = false>
__global__ void __launch_bounds__(64 + NE * 32 + (W4 ? kUnpackWarps * 32 : 0), 1)
qgemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
             const __grid_constant__ CUtensorMap tmap_r, const __grid_constant__ CUtensorMap tmap_b2, const QGemmArgs p,
             const uint32_t idesc, const uint32_t idesc2) {
  using L = SmemLayout<MODE, CL, NE, W4>;
  constexpr bool PAIR = CL >= 2, QUAD = CL == 4;
  constexpr int kNE = NE, kParts = NE / 4, CW = L::kCW;        // kParts = column parts per TMEM lane quarter
  constexpr int kCtas = CL;                                     // CTAs (128-row blocks) per scheduling unit
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;      // rank in the cluster; even ranks lead their pair
  const uint32_t pair_rank = cta_rank & 1u, pair_id = cta_rank >> 1;
  const int unit = blockIdx.x / CL;                             // persistent scheduling unit: a CTA, a pair or two pairs
  const int num_units = gridDim.x / CL;
  constexpr int BN = kBN;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + L::kStages * L::kABytes;
  float* colp = reinterpret_cast<float*>(smem + L::kColpOff);
  float* lut_s = reinterpret_cast<float*>(smem + L::kLutOff);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + L::kStages;
  uint64_t* tfull_bar = empty_bar + L::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* pfull_bar = tempty_bar + 2;                      // W4: packed k-slice landed / consumed by the unpack warps
  uint64_t* pempty_bar = pfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pempty_bar + 2);
  uint8_t* smem_p = smem + L::kPackOff;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + kCtas * kBM - 1) / (kCtas * kBM), n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_iters = (p.K + kBK - 1) / kBK;
  // Tail-wave splitting: the persistent units walk whole 128(x CL) x 256 tiles for every complete wave; the tiles of the last,
  // partial wave are cut into split_f column slices (own TMA box / UMMA N), so that e.g. 68 left-over tiles keep 136 of 148
  // SMs busy for half a tile time instead of 68 SMs for a whole one (o_proj at batch 8 x seq 1024: 4 -> 3.5 waves).
  const int SF = p.split_f;
  const int full_tiles = SF > 1 ? (num_tiles / num_units) * num_units : num_tiles;
  const int num_items = full_tiles + (num_tiles - full_tiles) * SF;
  const int sub_cols = BN / SF;
  // work item -> (tile, first column inside the tile, columns)
  auto item_tile = [&](int w, int& t, int& c0, int& nc) {
    if (w < full_tiles) { t = w; c0 = 0; nc = BN; }
    else { const int s2 = w - full_tiles; t = full_tiles + s2 / SF; c0 = (s2 % SF) * sub_cols; nc = sub_cols; }
  };

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();                 // the swizzle math below assumes a 1024B-aligned window
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_b2);
    if (MODE == EPI_RESID) prefetch_tmap(&tmap_r);
    // QUAD: a stage is free once BOTH pairs have retired the MMAs that read it (the B quarters are written across pairs)
    // W4: a stage is full once the A tile has landed (1 arrival + tx bytes) AND the four unpack warps have written the B tile
    for (int i = 0; i < L::kStages; ++i) { mbar_init(&full_bar[i], W4 ? 1 + kUnpackWarps : 1); mbar_init(&empty_bar[i], QUAD ? 2 : 1); }
    if (W4) for (int i = 0; i < L::kPStages; ++i) { mbar_init(&pfull_bar[i], 1); mbar_init(&pempty_bar[i], kUnpackWarps); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], kNE * (PAIR ? 2 : 1)); }
    fence_barrier_init();
  }
  if (warp == 1) { if (PAIR) tmem_alloc_pair(tmem_slot, 2 * BN); else tmem_alloc(tmem_slot, 2 * BN); }
  if (MODE == EPI_ACTMUL && threadIdx.x >= 64 && threadIdx.x < 64 + kNE * 32) {
    for (int i = threadIdx.x - 64; i < 256; i += kNE * 32) lut_s[i] = __ldg(p.lut + i);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                 // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int pstage = 0; uint32_t pphase = 0;
      (void)pstage; (void)pphase;
      for (int w = unit; w < num_items; w += num_units) {
        int t, c0, nc; item_tile(w, t, c0, nc);
        const bool sub = nc != BN;                           // column slice of a tail-wave tile: smaller B box (tmap_b2)
        const int brows = PAIR ? nc / 2 : nc;                // B rows this CTA loads
        const int m0 = (t / n_tiles) * (kCtas * kBM) + cta_rank * kBM, n0 = (t % n_tiles) * BN + c0 + pair_rank * brows;
        for (int k = 0; k < k_iters; ++k) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (PAIR) {
            // the pair leader's barrier collects the bytes of both CTAs
            if (pair_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (L::kABytes + brows * kBK));
            tma_load_2d_pair(smem_a + stage * L::kABytes, &tmap_a, &full_bar[stage], k * kBK, m0);
            if (QUAD)   // this CTA's quarter of the B tile, delivered to the same half of both pairs
              tma_load_2d_pair_mc(smem_b + stage * L::kBBytes + pair_id * (L::kBBoxRows * kBK), &tmap_b, &full_bar[stage], k * kBK,
                                  n0 + pair_id * L::kBBoxRows, (uint16_t)(0x5u << pair_rank));
            else
              tma_load_2d_pair(smem_b + stage * L::kBBytes, sub ? &tmap_b2 : &tmap_b, &full_bar[stage], k * kBK, n0);
          } else if (W4) {
            // A into the operand stage; the packed B k-slice (64 bytes per row) into the packed ring for the unpack warps
            mbar_expect_tx(&full_bar[stage], L::kABytes);
            tma_load_2d(smem_a + stage * L::kABytes, &tmap_a, &full_bar[stage], k * kBK, m0);
            mbar_wait(&pempty_bar[pstage], pphase ^ 1);
            mbar_expect_tx(&pfull_bar[pstage], L::kPBytes);
            tma_load_2d(smem_p + pstage * L::kPBytes, &tmap_b, &pfull_bar[pstage], k * (kBK / 2), n0);
            if (++pstage == L::kPStages) { pstage = 0; pphase ^= 1; }
          } else {
            mbar_expect_tx(&full_bar[stage], L::kABytes + brows * kBK);
            tma_load_2d(smem_a + stage * L::kABytes, &tmap_a, &full_bar[stage], k * kBK, m0);
            tma_load_2d(smem_b + stage * L::kBBytes, sub ? &tmap_b2 : &tmap_b, &full_bar[stage], k * kBK, n0);
          }
          if (++stage == L::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && pair_rank == 0) {
    // ===================== MMA issuer (the leader CTA of a pair) =====================
    const uint16_t pair_mask = (uint16_t)(0x3u << (2 * pair_id)), all_mask = (uint16_t)((1u << CL) - 1u);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int w = unit; w < num_items; w += num_units) {
      int t, c0, nc; item_tile(w, t, c0, nc);
      const uint32_t idesc_w = nc != BN ? idesc2 : idesc;  // UMMA N of a column slice
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);          // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int k = 0; k < k_iters; ++k) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = smem_desc_k128(smem_u32(smem_a + stage * L::kABytes));
          const uint64_t bdesc = smem_desc_k128(smem_u32(smem_b + stage * L::kBBytes));
#pragma unroll
          for (int kk = 0; kk < kBK / kUmmaK; ++kk) {
            // advance both descriptors by kk*32 bytes inside the 128B swizzle row (address field is in 16B units)
            if (PAIR) mma_i8_pair(tmem_d, adesc + uint64_t(kk * (kUmmaK >> 4)), bdesc + uint64_t(kk * (kUmmaK >> 4)), idesc_w, (k | kk) != 0);
            else mma_i8(tmem_d, adesc + uint64_t(kk * (kUmmaK >> 4)), bdesc + uint64_t(kk * (kUmmaK >> 4)), idesc_w, (k | kk) != 0);
          }
          if (PAIR) {
            tc_commit_pair(&empty_bar[stage], all_mask);                       // smem slot free once these MMAs retire
            if (k == k_iters - 1) tc_commit_pair(&tfull_bar[acc], pair_mask);  // accumulator complete (both halves)
          } else {
            tc_commit(&empty_bar[stage]);
            if (k == k_iters - 1) tc_commit(&tfull_bar[acc]);
          }
        }
        __syncwarp();
        if (++stage == L::kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (W4 && warp >= 2 + kNE) {
    // ===================== nibble expansion (W4) =====================
    // packed ring slot: 256 rows x 64 B, TMA SWIZZLE_64B (16-byte chunk c of row r at chunk c ^ ((r >> 1) & 3)); operand
    // slot: 256 rows x 128 B, SWIZZLE_128B (chunk c at c ^ (r & 7)).  A thread expands 16 packed bytes (32 codes, low nibble
    // first) of one row per step: 8 consecutive lanes take 8 consecutive rows, so loads and stores are bank-conflict free.
    const int ut = threadIdx.x - (64 + kNE * 32);
    int stage = 0; uint32_t phase = 0;
    int pstage = 0; uint32_t pphase = 0;
    for (int w = unit; w < num_items; w += num_units) {
      for (int k = 0; k < k_iters; ++k) {
        mbar_wait(&empty_bar[stage], phase ^ 1);            // the MMAs that read this operand slot have retired
        mbar_wait(&pfull_bar[pstage], pphase);              // the packed k-slice has landed
        const uint8_t* src = smem_p + pstage * L::kPBytes;
        uint8_t* dst = smem_b + stage * L::kBBytes;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int task = it * (kUnpackWarps * 32) + ut;
          const int r = task & 255, c = task >> 8;          // row, packed 16-byte chunk (4 per row)
          const uint4 pk = *reinterpret_cast<const uint4*>(src + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
          const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t lo = w[j] & 0x0F0F0F0Fu, hi = (w[j] >> 4) & 0x0F0F0F0Fu;
            o[2 * j] = __byte_perm(lo, hi, 0x5140);         // codes 0..3 of the word: lo0 hi0 lo1 hi1
            o[2 * j + 1] = __byte_perm(lo, hi, 0x7362);     // codes 4..7:             lo2 hi2 lo3 hi3
          }
          uint8_t* drow = dst + r * 128;
          *reinterpret_cast<uint4*>(drow + (((2 * c) ^ (r & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(drow + (((2 * c + 1) ^ (r & 7)) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
        }
        fence_proxy_async();                                // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) { mbar_arrive(&full_bar[stage]); mbar_arrive(&pempty_bar[pstage]); }
        if (++stage == L::kStages) { stage = 0; phase ^= 1; }
        if (++pstage == L::kPStages) { pstage = 0; pphase ^= 1; }
      }
    }
  } else if (warp >= 2 && warp < 2 + kNE) {
    // ===================== epilogue =====================
    const int ew = warp - 2;                 // 0..kNE-1
    const int quarter = warp & 3;            // TMEM lane quarter this warp may read
    const int part = ew >> 2;                // column part
    const int etid = threadIdx.x - 64;
    const bool has_bias = p.bias != nullptr;
    uint8_t* stage_out = smem + L::kOutOff + ew * L::kOutBufs * L::kOutTile;
    int out_buf = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int w = unit; w < num_items; w += num_units) {
      int t, c0s, nc; item_tile(w, t, c0s, nc);
      const int m0 = (t / n_tiles) * (kCtas * kBM) + cta_rank * kBM, n0 = (t % n_tiles) * BN + c0s;
      // stage the per-column parameters of this tile (double buffered with the accumulator)
      float* cp = colp + acc * CP_COUNT * BN;
      int* cpi = reinterpret_cast<int*>(cp);
      for (int c = etid; c < BN; c += kNE * 32) {
        const int n = n0 + c;
        const bool ok = n < p.N;
        cpi[CP_NEGOW * BN + c] = ok ? -__ldg(p.ow + n) : 0;
        cpi[CP_C0 * BN + c] = ok ? __ldg(p.c0 + n) : 0;
        cp[CP_SXW * BN + c] = ok ? __ldg(p.sxw + n) : 0.f;
        cp[CP_BIAS * BN + c] = (ok && has_bias) ? __ldg(p.bias + n) : 0.f;
      }
      // the output quantizer of every 32-column chunk of the tile (so / oo are per column group): staged with the column
      // parameters so that no global load sits between the accumulator's arrival and the first requantisation
      if (etid < BN / 32 && MODE != EPI_F32 && MODE != EPI_I32) {
        const int g = min((n0 + etid * 32) / p.qgroup, (p.N - 1) / p.qgroup);
        cp[CP_Q * BN + etid] = __ldg(p.so + g);
        cp[CP_Q * BN + 8 + etid] = __ldg(p.oo + g);
      }
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M && p.dbg != 1;
      const int rs = row_ok ? __ldg(p.rowsum + row) : 0;      // (issued before the wait: its latency hides behind the MMAs)
      asm volatile("bar.sync 1, %0;" ::"n"(kNE * 32) : "memory");
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();

      if (p.dbg == 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR && pair_rank != 0) mbar_arrive_remote(&tempty_bar[acc], cta_rank & ~1u); else mbar_arrive(&tempty_bar[acc]); }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      const uint32_t trow = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN;
      const int4* v_negow = reinterpret_cast<const int4*>(cpi + CP_NEGOW * BN);
      const int4* v_c0 = reinterpret_cast<const int4*>(cpi + CP_C0 * BN);
      const float4* v_sxw = reinterpret_cast<const float4*>(cp + CP_SXW * BN);
      const float4* v_bias = reinterpret_cast<const float4*>(cp + CP_BIAS * BN);
      int code_sum = 0;

      // y[j] of 4 consecutive columns starting at tile column c (c % 4 == 0)
      auto y4 = [&](const uint32_t* r, int c, float (&y)[4]) {
        const int4 no = v_negow[c >> 2], cz = v_c0[c >> 2];
        const float4 sx = v_sxw[c >> 2];
        const int i0 = (int)r[0] + no.x * rs + cz.x, i1 = (int)r[1] + no.y * rs + cz.y;
        const int i2 = (int)r[2] + no.z * rs + cz.z, i3 = (int)r[3] + no.w * rs + cz.w;
        if (MODE == EPI_I32) {
          y[0] = __int_as_float(i0); y[1] = __int_as_float(i1); y[2] = __int_as_float(i2); y[3] = __int_as_float(i3);
          return;
        }
        y[0] = __fmul_rn(__int2float_rn(i0), sx.x); y[1] = __fmul_rn(__int2float_rn(i1), sx.y);
        y[2] = __fmul_rn(__int2float_rn(i2), sx.z); y[3] = __fmul_rn(__int2float_rn(i3), sx.w);
        if (has_bias) {
          const float4 b = v_bias[c >> 2];
          y[0] = __fadd_rn(y[0], b.x); y[1] = __fadd_rn(y[1], b.y); y[2] = __fadd_rn(y[2], b.z); y[3] = __fadd_rn(y[3], b.w);
        }
      };
      auto group_q = [&](int col /* tile column, multiple of 32 */, float qmax) {
        return make_qparam(cp[CP_Q * BN + (col >> 5)], cp[CP_Q * BN + 8 + (col >> 5)], qmax);
      };

      if (MODE == EPI_ACTMUL) {
        // columns [0,128) of the tile are w1 rows, [128,256) the matching w3 rows
        constexpr int H = BN / 2;
        const QParam q1 = group_q(0, p.qmax), q3 = group_q(H, p.qmax);
        const QParam q2 = make_qparam(p.s2, p.o2, p.qmax2);
        dispatch_five(q1.five | q2.five | q3.five, [&](auto five_tag) {
        constexpr bool FIVE = decltype(five_tag)::value;
        for (int cc = part * (H / kParts); cc < (part + 1) * (H / kParts); cc += CW) {
          uint32_t r1[CW], r3[CW];
          tmem_ld(trow + cc, r1);
          tmem_ld(trow + H + cc, r3);
          tc_wait_ld();
          uint32_t packed[CW / 4];
#pragma unroll
          for (int j4 = 0; j4 < CW / 4; ++j4) {
            float y1[4], y3[4];
            y4(r1 + 4 * j4, cc + 4 * j4, y1);
            y4(r3 + 4 * j4, H + cc + 4 * j4, y3);
            uint32_t w = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = lut_s[quant_int<FIVE>(y1[e], q1)];                                      // fq_out(act(w1x))
              const float u = __fmul_rn(__fsub_rn(quant_magic<FIVE>(y3[e], q3), kRoundMagic), q3.s);  // fq(w3x)
              w |= (uint32_t)quant_int<FIVE>(__fmul_rn(a, u), q2) << (8 * e);                         // w2.input_quantizer
            }
            packed[j4] = w;
            code_sum = (int)__dp4a(w, 0x01010101u, (unsigned)code_sum);
          }
          const int ncol = n0 / 2 + cc;       // output column (N/2 wide)
          if (row_ok) {
            uint8_t* dst = reinterpret_cast<uint8_t*>(p.out) + int64_t(row) * p.ldo + ncol;
            if (ncol + CW <= p.N / 2 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
              for (int v = 0; v < CW / 16; ++v)
                reinterpret_cast<uint4*>(dst)[v] = make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
            } else {
              for (int j = 0; j < CW; ++j)
                if (ncol + j < p.N / 2) dst[j] = (uint8_t)(packed[j >> 2] >> (8 * (j & 3)));
            }
          }
        }
        });
      } else {
        const int W = nc / kParts;                       // a column slice of a tail-wave tile is narrower (nc a multiple of kParts * CW)
        for (int cc = part * W; cc < (part + 1) * W; cc += CW) {
          if (n0 + cc >= p.N) break;
          uint32_t r[CW];
          tmem_ld(trow + cc, r);
          tc_wait_ld();
          const int ncol = n0 + cc;
          const bool full = (ncol + CW <= p.N);
          if (MODE == EPI_QUANT) {
            const QParam q = group_q(cc, p.qmax);
            uint32_t code[CW];
            dispatch_five(q.five, [&](auto five_tag) {
              constexpr bool FIVE = decltype(five_tag)::value;
#pragma unroll
              for (int j4 = 0; j4 < CW / 4; ++j4) {
                float y[4];
                y4(r + 4 * j4, cc + 4 * j4, y);
#pragma unroll
                for (int e = 0; e < 4; ++e) code[4 * j4 + e] = (uint32_t)quant_int<FIVE>(y[e], q);
              }
            });
            if (full) {
#pragma unroll
              for (int j = 0; j < CW; ++j) code_sum += (int)code[j];
            } else {
              for (int j = 0; j < CW; ++j) if (ncol + j < p.N) code_sum += (int)code[j];
            }
            if (row_ok) {
              if (p.out_bits == 8) {
                uint8_t* dst = reinterpret_cast<uint8_t*>(p.out) + int64_t(row) * p.ldo + ncol;
                if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                  uint32_t w[CW / 4];
#pragma unroll
                  for (int j = 0; j < CW / 4; ++j) w[j] = code[4 * j] | (code[4 * j + 1] << 8) | (code[4 * j + 2] << 16) | (code[4 * j + 3] << 24);
#pragma unroll
                  for (int v = 0; v < CW / 16; ++v) reinterpret_cast<uint4*>(dst)[v] = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
                } else {
                  for (int j = 0; j < CW; ++j) if (ncol + j < p.N) dst[j] = (uint8_t)code[j];
                }
              } else {
                uint16_t* dst = reinterpret_cast<uint16_t*>(p.out) + int64_t(row) * p.ldo + ncol;
                if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                  for (int v = 0; v < CW / 8; ++v) {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) w[j] = code[8 * v + 2 * j] | (code[8 * v + 2 * j + 1] << 16);
                    reinterpret_cast<uint4*>(dst)[v] = make_uint4(w[0], w[1], w[2], w[3]);
                  }
                } else {
                  for (int j = 0; j < CW; ++j) if (ncol + j < p.N) dst[j] = (uint16_t)code[j];
                }
              }
            }
          } else if (MODE == EPI_RESID) {
            // de-quantised output-quantizer values (hm:1257,1270 add them to the fp32 stream) -> swizzled smem tile ->
            // TMA reduce-add.  CW 32: row r of the tile is 128 B, its 16-byte chunk c lives at chunk (c ^ (r & 7))
            // (SWIZZLE_128B); CW 16: rows of 64 B, chunk (c ^ ((r >> 1) & 3)) (SWIZZLE_64B).
            const QParam q = group_q(cc, p.qmax);
            uint8_t* tile = stage_out + out_buf * L::kOutTile;
            if (lane == 0) bulk_wait_read<L::kOutBufs - 1>();      // the tile's previous reduce has been read out
            __syncwarp();
            dispatch_five(q.five, [&](auto five_tag) {
              constexpr bool FIVE = decltype(five_tag)::value;
#pragma unroll
              for (int j4 = 0; j4 < CW / 4; ++j4) {
                float y[4], v[4];
                y4(r + 4 * j4, cc + 4 * j4, y);
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = __fmul_rn(__fsub_rn(quant_magic<FIVE>(y[e], q), kRoundMagic), q.s);
                const int sw = CW == 32 ? (lane & 7) : ((lane >> 1) & 3);
                *reinterpret_cast<float4*>(tile + lane * (CW * 4) + ((j4 ^ sw) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
              }
            });
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_2d(&tmap_r, tile, ncol, m0 + quarter * 32);   // rows >= M / columns >= N are clipped by the TMA
              bulk_commit();
            }
            if (++out_buf == L::kOutBufs) out_buf = 0;
          } else {   // EPI_F32 / EPI_I32
            float y[CW];
#pragma unroll
            for (int j4 = 0; j4 < CW / 4; ++j4) {
              float yy[4];
              y4(r + 4 * j4, cc + 4 * j4, yy);
              y[4 * j4] = yy[0]; y[4 * j4 + 1] = yy[1]; y[4 * j4 + 2] = yy[2]; y[4 * j4 + 3] = yy[3];
            }
            if (row_ok) {
              float* dst = reinterpret_cast<float*>(p.out) + int64_t(row) * p.ldo + ncol;
              if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                for (int v = 0; v < CW / 4; ++v)
                  reinterpret_cast<float4*>(dst)[v] = make_float4(y[4 * v], y[4 * v + 1], y[4 * v + 2], y[4 * v + 3]);
              } else {
                for (int j = 0; j < CW; ++j) if (ncol + j < p.N) dst[j] = y[j];
              }
            }
          }
        }
      }
      if (p.rowsum_out && row_ok && (MODE == EPI_QUANT || MODE == EPI_ACTMUL)) atomicAdd(p.rowsum_out + row, code_sum);
      // release the accumulator back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR && pair_rank != 0) mbar_arrive_remote(&tempty_bar[acc], cta_rank & ~1u); else mbar_arrive(&tempty_bar[acc]); }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (MODE == EPI_RESID && lane == 0) bulk_wait<0>();      // all reduce-adds have landed before the grid retires
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                 // the peer may still be reading this CTA's B half / signalling its barriers
  if (warp == 1) { if (PAIR) tmem_dealloc_pair(tmem_base, 2 * BN); else tmem_dealloc(tmem_base, 2 * BN); }
}

// ---- host -----------------------------------------------------------------------------------------------------------
template <int MODE, int CL, int NE, bool W4 = false>
static int launch_qgemm2(Ctx* c, const void* a, const void* b, const QGemmArgs& args, float* resid, int a_signed, int b_signed,
                         cudaStream_t st) {
  using L = SmemLayout<MODE, CL, NE, W4>;
  constexpr bool PAIR = CL >= 2;
  constexpr int kThreads = 64 + NE * 32 + (W4 ? kUnpackWarps * 32 : 0);
  CUtensorMap ta, tb, tr;
  // W4: b is the packed matrix [N, K/2] (64-byte boxes, SWIZZLE_64B) -- it is expanded inside the kernel
  if (!make_tmap_2d(&ta, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, a, args.M, args.K, args.K, kBM) ||
      !(W4 ? make_tmap_2d(&tb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, b, args.N, args.K / 2, args.K / 2, kBN, 64)
           : make_tmap_2d(&tb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, b, args.N, args.K, args.K, L::kBBoxRows)))
    return fail(c, MQ_RUNTIME_ERROR, "cuTensorMapEncodeTiled failed (pointers must be 16B aligned, K a multiple of 16)");
  tr = ta;
  CUtensorMap tb2 = tb;
  if (MODE == EPI_RESID && !make_tmap_2d(&tr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, resid, args.M, args.N, args.ldo * 4, 32, L::kCW * 4))
    return fail(c, MQ_RUNTIME_ERROR, "cuTensorMapEncodeTiled failed for the residual stream (16B-aligned pointer, ldo % 4 == 0)");
  static int max_units = -1;                    // co-resident clusters of this variant (GPC boundaries can cost a few)
  if (max_units < 0) {
    cudaError_t e = cudaFuncSetAttribute(qgemm_kernel<MODE, CL, NE, W4>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return fail(c, MQ_RUNTIME_ERROR, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    if (CL > 2 && (e = cudaFuncSetAttribute(qgemm_kernel<MODE, CL, NE, W4>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)) != cudaSuccess)
      return fail(c, MQ_RUNTIME_ERROR, std::string("cudaFuncSetAttribute(cluster): ") + cudaGetErrorString(e));
    max_units = c->sm_count / CL;
    if (PAIR) {
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(c->sm_count / CL * CL); q.blockDim = dim3(kThreads); q.dynamicSmemBytes = L::kTotal;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, qgemm_kernel<MODE, CL, NE, W4>, &q) == cudaSuccess && n > 0 && n < max_units) max_units = n;
      cudaGetLastError();
    }
  }
  const int m_tiles = (args.M + CL * kBM - 1) / (CL * kBM), n_tiles = (args.N + kBN - 1) / kBN;
  int units = m_tiles * n_tiles;
  if (units > max_units) units = max_units;
  // tail-wave splitting (see the kernel): column slices for the tiles of the last, partial wave.  MQ_QGEMM_SPLIT=0 disables it.
  QGemmArgs args2 = args;
  args2.split_f = 1;
  {
    const char* e = getenv("MQ_QGEMM_SPLIT");
    const int rem = (m_tiles * n_tiles) % units;
    if (!(e && e[0] == '0') && MODE != EPI_ACTMUL && !W4 && CL <= 2 && rem > 0) {
      if (rem * 4 <= units) args2.split_f = 4; else if (rem * 2 <= units) args2.split_f = 2;
    }
    if (args2.split_f > 1 && !make_tmap_2d(&tb2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, b, args.N, args.K, args.K, L::kBBoxRows / args2.split_f))
      return fail(c, MQ_RUNTIME_ERROR, "cuTensorMapEncodeTiled failed for the tail-wave column slices");
  }
  const uint32_t idesc = make_idesc(2u, a_signed ? 1u : 0u, b_signed ? 1u : 0u, 0u, 0u, (PAIR ? 2 : 1) * kBM, kBN);
  const uint32_t idesc2 = make_idesc(2u, a_signed ? 1u : 0u, b_signed ? 1u : 0u, 0u, 0u, (PAIR ? 2 : 1) * kBM, kBN / args2.split_f);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(units * CL); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = L::kTotal; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = PAIR ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, qgemm_kernel<MODE, CL, NE, W4>, ta, tb, tr, tb2, args2, idesc, idesc2);
  if (e != cudaSuccess) return fail(c, MQ_RUNTIME_ERROR, std::string("mq_qgemm launch: ") + cudaGetErrorString(e));
  return check_launch(c, "mq_qgemm");
}

// Kernel choice (measured on B200, profiles/r1_qgemm_variants.md): the operand traffic from L2 is the ceiling of the
// single-CTA kernel; the CTA-pair kernel halves the B traffic, the two-pair cluster with B multicast halves it again.
// MQ_QGEMM_CL=1/2/4 forces one variant (A/B measurements; read per call); MQ_QGEMM_PAIR=0/1 is the older spelling.
static int pick_cluster(int M, int K) {
  const char* e = getenv("MQ_QGEMM_CL");
  if (e && (e[0] == '1' || e[0] == '2' || e[0] == '4')) return e[0] - '0';
  e = getenv("MQ_QGEMM_PAIR");
  if (e) return e[0] == '0' ? 1 : 2;
  (void)M;
  return K >= 4096 ? 2 : 1;
}

// MQ_QGEMM_NE=8/16 picks the number of epilogue warps (A/B measurements; read per call).  Measured on B200
// (profiles/r1d_qgemm_variants.md): 16 warps do not shorten the tile -- under a sustained load the GEMM runs into the
// 1000 W power cap (SM clock 1.1-1.6 GHz), not into epilogue latency -- so 8 stays the default.
static int pick_epi_warps() {
  const char* e = getenv("MQ_QGEMM_NE");
  return (e && e[0] == '1') ? 16 : 8;
}

template <int MODE, int NE>
static int launch_qgemm1(Ctx* c, const void* a, const void* b, const QGemmArgs& args, float* resid, int a_signed, int b_signed,
                         cudaStream_t st) {
  switch (pick_cluster(args.M, args.K)) {
    case 4: return launch_qgemm2<MODE, 4, NE>(c, a, b, args, resid, a_signed, b_signed, st);
    case 2: return launch_qgemm2<MODE, 2, NE>(c, a, b, args, resid, a_signed, b_signed, st);
    default: return launch_qgemm2<MODE, 1, NE>(c, a, b, args, resid, a_signed, b_signed, st);
  }
}
template <int MODE>
static int launch_qgemm(Ctx* c, const void* a, const void* b, const QGemmArgs& args, float* resid, int a_signed, int b_signed,
                        cudaStream_t st) {
  return pick_epi_warps() == 8 ? launch_qgemm1<MODE, 8>(c, a, b, args, resid, a_signed, b_signed, st)
                               : launch_qgemm1<MODE, 16>(c, a, b, args, resid, a_signed, b_signed, st);
}

}  // namespace mq

using namespace mq;

static int qgemm_entry(void* ctx, bool w4, const void* a_codes, int a_signed, const void* b_codes, int b_signed, int M, int N, int K,
                       const int32_t* rowsum, const float* sxw, const int32_t* ow, const int32_t* c0, const float* bias,
                       int mode, const float* so, const float* oo, float qmax, int out_bits, void* out, int64_t ldo,
                       int32_t* rowsum_out, const float* lut, float s2, float o2, float qmax2, float* resid, int qgroup,
                       void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, a_codes && b_codes && M > 0 && N > 0 && K > 0, "null operand or empty problem");
  MQ_REQUIRE(c, K % 16 == 0, "K must be a multiple of 16 bytes (TMA row pitch)");
  MQ_REQUIRE(c, !w4 || K % 32 == 0, "packed 4-bit weights: K must be a multiple of 32 (16-byte rows of packed codes)");
  MQ_REQUIRE(c, (reinterpret_cast<uintptr_t>(a_codes) & 15) == 0 && (reinterpret_cast<uintptr_t>(b_codes) & 15) == 0,
             "operands must be 16-byte aligned");
  MQ_REQUIRE(c, rowsum && sxw && ow && c0, "rowsum / sxw / ow / c0 are required");
  MQ_REQUIRE(c, mode >= EPI_QUANT && mode <= EPI_I32, "unknown epilogue mode");
  MQ_REQUIRE(c, mode != EPI_QUANT || (out && so && oo && (out_bits == 8 || out_bits == 16)), "EPI_QUANT needs out/so/oo, 8 or 16 bits");
  MQ_REQUIRE(c, mode != EPI_ACTMUL || (out && so && oo && lut && N % 256 == 0), "EPI_ACTMUL needs out/so/oo/lut and N % 256 == 0");
  MQ_REQUIRE(c, mode != EPI_RESID || (resid && so && oo && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(resid) & 15) == 0),
             "EPI_RESID needs so/oo and a 16-byte aligned resid with ldo % 4 == 0");
  MQ_REQUIRE(c, (mode != EPI_F32 && mode != EPI_I32) || out, "raw output needs out");
  MQ_REQUIRE(c, qmax < 4194304.f && qmax2 < 4194304.f, "qmax must be below 2^22");
  MQ_REQUIRE(c, (mode == EPI_F32 || mode == EPI_I32) || (qgroup > 0 && qgroup % 32 == 0), "qgroup must be a positive multiple of 32");
  MQ_REQUIRE(c, mode != EPI_ACTMUL || 128 % qgroup == 0, "EPI_ACTMUL needs qgroup to divide 128");
  QGemmArgs args;
  args.M = M; args.N = N; args.K = K; args.rowsum = rowsum; args.sxw = sxw; args.ow = ow; args.c0 = c0;
  args.bias = bias; args.so = so; args.oo = oo; args.qmax = qmax; args.out_bits = out_bits; args.out = out; args.ldo = ldo;
  args.dbg = 0; args.split_f = 1;
#ifdef MQ_MEASURE_KNOBS   // measurement builds only (MQB200_MEASURE=1 python -m mobilequant_b200.build --force): dbg 1 drops the epilogue
  { const char* e = getenv("MQ_QGEMM_DBG"); args.dbg = e ? atoi(e) : 0; }
#endif
  args.qgroup = qgroup > 0 ? qgroup : 32; args.rowsum_out = rowsum_out; args.lut = lut; args.s2 = s2; args.o2 = o2; args.qmax2 = qmax2;
  cudaStream_t st = (cudaStream_t)stream;
  if (w4) {          // packed weights: single-CTA variant, 8 epilogue warps + 4 nibble-expansion warps; codes are unsigned nibbles
    switch (mode) {
      case EPI_QUANT: return launch_qgemm2<EPI_QUANT, 1, 8, true>(c, a_codes, b_codes, args, resid, a_signed, 0, st);
      case EPI_ACTMUL: return launch_qgemm2<EPI_ACTMUL, 1, 8, true>(c, a_codes, b_codes, args, resid, a_signed, 0, st);
      case EPI_RESID: return launch_qgemm2<EPI_RESID, 1, 8, true>(c, a_codes, b_codes, args, resid, a_signed, 0, st);
      case EPI_F32: return launch_qgemm2<EPI_F32, 1, 8, true>(c, a_codes, b_codes, args, resid, a_signed, 0, st);
      default: return launch_qgemm2<EPI_I32, 1, 8, true>(c, a_codes, b_codes, args, resid, a_signed, 0, st);
    }
  }
  switch (mode) {
    case EPI_QUANT: return launch_qgemm<EPI_QUANT>(c, a_codes, b_codes, args, resid, a_signed, b_signed, st);
    case EPI_ACTMUL: return launch_qgemm<EPI_ACTMUL>(c, a_codes, b_codes, args, resid, a_signed, b_signed, st);
    case EPI_RESID: return launch_qgemm<EPI_RESID>(c, a_codes, b_codes, args, resid, a_signed, b_signed, st);
    case EPI_F32: return launch_qgemm<EPI_F32>(c, a_codes, b_codes, args, resid, a_signed, b_signed, st);
    default: return launch_qgemm<EPI_I32>(c, a_codes, b_codes, args, resid, a_signed, b_signed, st);
  }
}

extern "C" int mq_qgemm(void* ctx, const void* a_codes, int a_signed, const void* b_codes, int b_signed, int M, int N, int K,
                        const int32_t* rowsum, const float* sxw, const int32_t* ow, const int32_t* c0, const float* bias,
                        int mode, const float* so, const float* oo, float qmax, int out_bits, void* out, int64_t ldo,
                        int32_t* rowsum_out, const float* lut, float s2, float o2, float qmax2, float* resid, int qgroup,
                        void* stream) {
  return qgemm_entry(ctx, false, a_codes, a_signed, b_codes, b_signed, M, N, K, rowsum, sxw, ow, c0, bias, mode, so, oo, qmax, out_bits, out,
                     ldo, rowsum_out, lut, s2, o2, qmax2, resid, qgroup, stream);
}

extern "C" int mq_qgemm_w4a8(void* ctx, const void* a_codes, int a_signed, const void* b_packed, int M, int N, int K,
                             const int32_t* rowsum, const float* sxw, const int32_t* ow, const int32_t* c0, const float* bias,
                             int mode, const float* so, const float* oo, float qmax, int out_bits, void* out, int64_t ldo,
                             int32_t* rowsum_out, const float* lut, float s2, float o2, float qmax2, float* resid, int qgroup,
                             void* stream) {
  return qgemm_entry(ctx, true, a_codes, a_signed, b_packed, 0, M, N, K, rowsum, sxw, ow, c0, bias, mode, so, oo, qmax, out_bits, out,
                     ldo, rowsum_out, lut, s2, o2, qmax2, resid, qgroup, stream);
}
