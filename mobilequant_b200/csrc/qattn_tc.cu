// K6 on the 5th-generation tensor cores: exact quantised causal attention == HFAttention.forward hm:510-534 with
// QMatMul qk_bmm / pv_bmm (qm:453-466), the arithmetic of engine.cu:qattn_kernel / oracle/int_ref.py:qattn_int bit for bit:
//   I_ij = sum_d (q-oq)(k-ok)         c_ij = clamp(rne((float(I)*sq*sk)/s_s)+o_s, 0, qmax_s)       (16-bit score code)
//   E_ij = (A[k>>8]*B[k&255])>>31, k = cmax_i - c_ij        p_ij = float(E_ij)/float(sum_j E_ij)   (exact u64 row sum)
//   cp_ij = clamp(rne(p/s_p), 0, qmax_p)                    A_id = sum_j cp_ij*(v_jd-ov)           (16-bit prob code)
//   out = clamp(rne((float(A)*s_p*s_v)/s_out)+o_out, 0, 255)
//
// Structure (FA4-like; one persistent CTA per SM, work item = 128 query rows of one (batch, head)):
//   warp 16       TMA producer  : Q tile [128 x hd], K tiles [128 keys x hd], V^T tiles [hd x 128 keys]
//                                 (cp.async.bulk.tensor, 64B / 128B swizzle) + the per-key zero-point terms -oq*rsk[j]
//   warp 17       MMA issuer    : S = Q.K^T  (tcgen05.mma kind::i8, M 128, N 128, s32 in TMEM, two S buffers)
//                                 O_lo += P_lo.V, O_hi += P_hi.V (M 128, N hd; P = hi / lo bytes of the 16-bit prob codes,
//                                 written by the softmax warps into 128B-swizzled shared memory)
//   warps 0..15   softmax       : thread = one query row (TMEM lane) x 32 keys of each 128-key tile (tcgen05.ld 32x32b.x32);
//                                 row max / row sum / prob codes are thread-local, the four column groups of a row meet
//                                 through shared memory once per pass.
// The exact softmax needs the row maximum before any E and the row sum before any prob code, so the key tiles are
// walked three times (max | sum | P.V); Q.K^T is simply re-issued -- the tensor pipe is idle otherwise (a pass costs
// 128 MMA cycles per tile against ~6000 ALU cycles) -- and no score ever leaves the SM: nothing is parked, any T works.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_host.h"
#include "ctx.h"
#include "attn.cuh"
#include <string>
#include <cstdlib>
#include <cstdio>

namespace mq {
using namespace tc;

constexpr int kTM = 128;            // query rows per work item (UMMA M)
constexpr int kTN = 128;            // keys per tile (UMMA N of S, K extent of P.V)
constexpr int kSoftWarps = 16;
constexpr int kTcThreads = kSoftWarps * 32 + 128;   // 16 softmax warps (4 warpgroups) + one control warpgroup (TMA, MMA, 2 idle)
constexpr int kCkRing = 8;          // ring of per-key correction vectors (> K ring + S buffers, see the producer)
constexpr int kRepA = 8, kRepB = 16; // bank replication of the two exp tables (32 B / 64 B per entry)
constexpr int kTabBytes = 256 * (kRepA + kRepB) * 4;

template <int HD>
struct TcLayout {
  static constexpr int kQBytes = kTM * HD;
  static constexpr int kKBytes = kTN * HD;              // one K tile
  static constexpr int kVBytes = HD * kTN;              // one V^T tile
  static constexpr int kPBytes = kTM * kTN;             // one byte plane of P
  static constexpr int kKStages = HD <= 64 ? 4 : 3;
  static constexpr int kVStages = HD <= 64 ? 3 : 2;
  static constexpr int kQOff = 0;                                   // [2][kQBytes]
  static constexpr int kKOff = kQOff + 2 * kQBytes;                 // [kKStages][kKBytes]
  static constexpr int kVOff = kKOff + kKStages * kKBytes;          // [kVStages][kVBytes]
  static constexpr int kPOff = kVOff + kVStages * kVBytes;          // [2 bufs][lo, hi][kPBytes]
  static constexpr int kTabOff = kPOff + 4 * kPBytes;
  static constexpr int kCkOff = kTabOff + kTabBytes + 16384;        // [kCkRing][128] int (16 KB of slack: the tables are aligned at run time)
  static constexpr int kXiOff = kCkOff + kCkRing * kTN * 4;         // [4][128] int   (row max / prob-code sums)
  static constexpr int kXsOff = kXiOff + 4 * kTM * 4;               // [4][128] u64   (row sums)
  static constexpr int kBarOff = kXsOff + 4 * kTM * 8;
  static constexpr int kNumBars = 2 + 2 + 2 * kKStages + 2 * kVStages + 2 + 2 + 2 + 2 + 2 + kCkRing;
  static constexpr int kTotal = kBarOff + kNumBars * 8 + 16;
  static_assert(kKOff % 1024 == 0 && kVOff % 1024 == 0 && kPOff % 1024 == 0, "operand tiles must keep the swizzle phase");
  static_assert(kTotal <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
};

// K-major operand tile with rows of exactly 64 bytes, 64-byte swizzle (TMA SWIZZLE_64B box of 64-byte rows):
// 8-row groups are 512 bytes apart (cute: Swizzle<2,4,3> o ((8,n),2):((4,SBO),1) in 16-byte units)
__device__ __forceinline__ uint64_t smem_desc_k64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;              // SWIZZLE_64B
  return d;
}
template <int HD>
__device__ __forceinline__ uint64_t qk_desc(uint32_t smem_addr) { return HD == 64 ? smem_desc_k64(smem_addr) : smem_desc_k128(smem_addr); }

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int HD, bool FIVE>
__global__ void __launch_bounds__(kTcThreads, 1)
qattn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnArgs a, const int n_items, const uint32_t idesc_s,
                const uint32_t idesc_pv) {
  using L = TcLayout<HD>;
  constexpr int KS = L::kKStages, VS = L::kVStages;
  constexpr int KSTEPS = HD / 32;               // MMAs per S tile
  constexpr int OC = HD / 4;                    // output columns per softmax thread
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_q = smem + L::kQOff;
  uint8_t* s_k = smem + L::kKOff;
  uint8_t* s_v = smem + L::kVOff;
  uint8_t* s_p = smem + L::kPOff;
  // exp tables: B (256 x 64 B) on a 16 KB boundary of the shared window, A (256 x 32 B) right behind it on an 8 KB boundary,
  // so that a lookup address is (index field of k) | (table base + bank copy of the lane): one shift + one LOP3
  const uint32_t tab_addr = (smem_u32(smem + L::kTabOff) + 16383u) & ~16383u;
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem + L::kTabOff + (tab_addr - smem_u32(smem + L::kTabOff)));
  int* s_ck = reinterpret_cast<int*>(smem + L::kCkOff);
  int* s_xi = reinterpret_cast<int*>(smem + L::kXiOff);
  unsigned long long* s_xs = reinterpret_cast<unsigned long long*>(smem + L::kXsOff);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* q_full = bars;            uint64_t* q_empty = q_full + 2;
  uint64_t* k_full = q_empty + 2;     uint64_t* k_empty = k_full + KS;
  uint64_t* v_full = k_empty + KS;    uint64_t* v_empty = v_full + VS;
  uint64_t* t_full = v_empty + VS;    uint64_t* t_empty = t_full + 2;
  uint64_t* p_full = t_empty + 2;     uint64_t* p_empty = p_full + 2;
  uint64_t* o_full = p_empty + 2;     uint64_t* o_empty = o_full + 1;
  uint64_t* ck_full = o_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L::kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = (a.Tq + kTM - 1) / kTM;
  const int per_qt = a.B * a.nh;
  const int rep = a.nh / a.nkv;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    prefetch_tmap(&tmap_q); prefetch_tmap(&tmap_k); prefetch_tmap(&tmap_v);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < KS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < VS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], kSoftWarps);
      mbar_init(&p_full[i], kSoftWarps); mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_full, 1); mbar_init(o_empty, kSoftWarps);
    for (int i = 0; i < kCkRing; ++i) mbar_init(&ck_full[i], 1);
    fence_barrier_init();
  }
  if (warp == kSoftWarps + 1) tmem_alloc(tmem_slot, 512);
  if (warp < kSoftWarps) {                  // exp tables, entries replicated across banks (lane & 7 / lane & 15 picks the copy)
    for (int i = threadIdx.x; i < 256; i += kSoftWarps * 32) {
      const uint32_t va = __ldg(a.lut + i), vb = __ldg(a.lut + 256 + i);
#pragma unroll
      for (int j = 0; j < kRepB / 4; ++j) reinterpret_cast<uint4*>(s_tab)[(kRepB / 4) * i + j] = make_uint4(vb, vb, vb, vb);
#pragma unroll
      for (int j = 0; j < kRepA / 4; ++j) reinterpret_cast<uint4*>(s_tab)[256 * (kRepB / 4) + (kRepA / 4) * i + j] = make_uint4(va, va, va, va);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColS = 0, kColOlo = 256, kColOhi = 256 + HD;      // TMEM columns: S0 | S1 | O_lo | O_hi

  // work item -> (query tile, batch, head); heavy (late) query tiles first
  auto item_coords = [&](int item, int& qt, int& b, int& h) {
    qt = nqt - 1 - item / per_qt;
    const int rem = item % per_qt;
    h = rem % a.nh; b = rem / a.nh;
  };
  // key tiles an item walks: keys 0 .. q_start + (qt+1)*128 - 1, clipped to T
  auto item_tiles = [&](int qt) { return min((a.q_start + (qt + 1) * kTM + kTN - 1) / kTN, (a.T + kTN - 1) / kTN); };

  // register rebalancing: the control warpgroup hands registers to the softmax warpgroups (setmaxnreg is per warpgroup)
  if (warp >= kSoftWarps) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");

  if (warp == kSoftWarps) {
    // ================================================= TMA producer =================================================
    uint32_t gk = 0, gv = 0, it = 0;
    const int ioq = (int)a.oq;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      int qt, b, h; item_coords(item, qt, b, h);
      const int kvh = h / rep;
      const int n = item_tiles(qt);
      const uint32_t qb = it & 1;
      if (lane == 0) {
        mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[qb], L::kQBytes);
        tma_load_2d(s_q + qb * L::kQBytes, &tmap_q, &q_full[qb], 0, (b * a.nh + h) * a.Tq + qt * kTM);
      }
      const int krow0 = (b * a.nkv + kvh) * a.T;
      const int32_t* rskb = a.rsk + int64_t(krow0);
      for (int pass = 0; pass < 3; ++pass) {
        for (int t = 0; t < n; ++t, ++gk) {
          const uint32_t ks = gk % KS, c8 = gk % kCkRing;
          // the ring of correction vectors needs no empty barrier: the K ring (<= 4) plus the two S buffers bound how far
          // this warp can run ahead of the slowest softmax warp (<= 6 tiles < kCkRing)
          const int key = t * kTN + lane * 4;
          int4 rk = make_int4(0, 0, 0, 0);
          if (key < a.T) rk = __ldg(reinterpret_cast<const int4*>(rskb + key));      // T % 16 == 0: whole int4 inside the row
          if (lane == 0) {
            mbar_wait(&k_empty[ks], ((gk / KS) & 1) ^ 1);
            mbar_expect_tx(&k_full[ks], L::kKBytes);
            tma_load_2d(s_k + ks * L::kKBytes, &tmap_k, &k_full[ks], 0, krow0 + t * kTN);
          }
          __syncwarp();                  // every lane writes its slot only after K(gk - KS) has been consumed
          reinterpret_cast<int4*>(s_ck + c8 * kTN)[lane] = make_int4(-ioq * rk.x, -ioq * rk.y, -ioq * rk.z, -ioq * rk.w);
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&ck_full[c8]);
            if (pass == 2) {
              const uint32_t vs = gv % VS;
              mbar_wait(&v_empty[vs], ((gv / VS) & 1) ^ 1);
              mbar_expect_tx(&v_full[vs], L::kVBytes);
              tma_load_2d(s_v + vs * L::kVBytes, &tmap_v, &v_full[vs], t * kTN, (b * a.nkv + kvh) * HD);
              ++gv;
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kSoftWarps + 1) {
    // ================================================= MMA issuer ===================================================
    uint32_t gk = 0, gv = 0, gp = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      int qt, b, h; item_coords(item, qt, b, h);
      const int n = item_tiles(qt);
      const uint32_t qb = it & 1;
      mbar_wait(&q_full[qb], (it >> 1) & 1);
      const uint64_t qdesc = qk_desc<HD>(smem_u32(s_q + qb * L::kQBytes));
      auto issue_pv = [&](int t) {
        const uint32_t pb = gp & 1, vs = gv % VS;
        mbar_wait(&p_full[pb], (gp >> 1) & 1);
        mbar_wait(&v_full[vs], (gv / VS) & 1);
        if (t == 0) mbar_wait(o_empty, (it & 1) ^ 1);        // the previous item's output has been read out of TMEM
        tc_fence_after();
        if (lane == 0) {
          const uint64_t plo = smem_desc_k128(smem_u32(s_p + (pb * 2 + 0) * L::kPBytes));
          const uint64_t phi = smem_desc_k128(smem_u32(s_p + (pb * 2 + 1) * L::kPBytes));
          const uint64_t vdesc = smem_desc_k128(smem_u32(s_v + vs * L::kVBytes));
#pragma unroll
          for (int kk = 0; kk < kTN / 32; ++kk) {
            mma_i8(tmem_base + kColOlo, plo + uint64_t(kk * 2), vdesc + uint64_t(kk * 2), idesc_pv, (t | kk) != 0);
            mma_i8(tmem_base + kColOhi, phi + uint64_t(kk * 2), vdesc + uint64_t(kk * 2), idesc_pv, (t | kk) != 0);
          }
          tc_commit(&p_empty[pb]);
          tc_commit(&v_empty[vs]);
          if (t == n - 1) tc_commit(o_full);
        }
        __syncwarp();
        ++gp; ++gv;
      };
      for (int pass = 0; pass < 3; ++pass) {
        for (int t = 0; t < n; ++t, ++gk) {
          const uint32_t ks = gk % KS, sb = gk & 1;
          mbar_wait(&k_full[ks], (gk / KS) & 1);
          mbar_wait(&t_empty[sb], ((gk >> 1) & 1) ^ 1);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t kdesc = qk_desc<HD>(smem_u32(s_k + ks * L::kKBytes));
#pragma unroll
            for (int kk = 0; kk < KSTEPS; ++kk)
              mma_i8(tmem_base + kColS + sb * kTN, qdesc + uint64_t(kk * 2), kdesc + uint64_t(kk * 2), idesc_s, kk != 0);
            tc_commit(&k_empty[ks]);
            tc_commit(&t_full[sb]);
            if (pass == 2 && t == n - 1) tc_commit(&q_empty[qb]);      // every S tile of the item has been issued
          }
          __syncwarp();
          if (pass == 2 && t > 0) issue_pv(t - 1);          // S(t) is in flight while the softmax warps finish P(t-1)
        }
      }
      issue_pv(n - 1);
    }
  } else if (warp < kSoftWarps) {
    // ================================================= softmax warps ================================================
    const int sw = warp;
    const int quarter = warp & 3;               // TMEM lane quarter this warp may access
    const int cg = sw >> 2;                     // column group: keys [32 cg, 32 cg + 32) of every tile
    const int row = quarter * 32 + lane;        // row of the work item == TMEM lane
    const uint32_t t_lane = tmem_base + (uint32_t(quarter * 32) << 16);
    const QParam qs = make_qparam(a.s_s, a.o_s, a.qmax_s);
    const QParam qp = make_qparam(a.s_p, 0.f, a.qmax_p);
    const QParam qo = make_qparam(a.s_out, a.o_out, 255.f);
    const int iok = (int)a.ok, ioq = (int)a.oq, iov = (int)a.ov;
    static_assert(kRepA == 8 && kRepB == 16, "lookup addresses hard-code the entry sizes (32 / 64 bytes)");
    const uint32_t tabB = tab_addr + (lane & (kRepB - 1)) * 4, tabA = tab_addr + 16384u + (lane & (kRepA - 1)) * 4;
    auto lds32 = [](uint32_t addr) -> uint32_t { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; };
    auto exp_tab = [&](int k) -> uint32_t {      // E(k) = (A[k >> 8] * B[k & 255]) >> 31; indices masked: discarded lanes stay in range
      const uint32_t ea = lds32((((uint32_t)k >> 3) & 0x1FE0u) | tabA);
      const uint32_t eb = lds32((((uint32_t)k << 6) & 0x3FC0u) | tabB);
      return (uint32_t)(((unsigned long long)ea * eb) >> 31);
    };
    // hd <= 64: |I| < 2^22, so float(I) comes from the magic-number trick (IADD3 + FADD) instead of I2F
    constexpr bool MAGIC = HD <= 64;
    auto score_bits = [&](int v) -> int {        // v = I (or I + magic bits); bits of magic + (clamped code - o_s)
      const float f = MAGIC ? __fsub_rn(__int_as_float(v), kRoundMagic) : __int2float_rn(v);
      return __float_as_int(quant_magic<FIVE>(__fmul_rn(f, a.sqk), qs));
    };
    uint32_t gs = 0, gp = 0, it = 0;
    const int ldo = a.nh * HD;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      int qt, b, h; item_coords(item, qt, b, h);
      const int n = item_tiles(qt);
      const int qloc = qt * kTM + row;                       // query row inside this call's q / out
      const int qi = a.q_start + qloc;                       // absolute position
      const bool row_ok = qloc < a.Tq;
      const int rsq = row_ok ? __ldg(a.rsq + (int64_t(b) * a.nh + h) * a.Tq + qloc) : 0;
      // I = acc - iok*rsq - ioq*rsk + HD*ioq*iok = acc + ck[key] + rc
      const int rc = HD * ioq * iok - iok * rsq + (MAGIC ? kRoundMagicBits : 0);
      // tile t covers keys [128 t, 128 t + 128); only the last tile of an item can hold keys beyond a query of the item
      const int diag_t = (a.q_start + qt * kTM) / kTN;       // q_start % 128 == 0: the diagonal tile starts at the item's first row
      // classification of this thread's 32-key chunk in tile t: 0 = fully visible, 1 = triangular (key <= row inside the chunk),
      // 2 = fully masked (warp-uniform)
      auto chunk_kind = [&](int t) -> int { return t < diag_t ? 0 : (cg < quarter ? 0 : (cg == quarter ? 1 : 2)); };

      // ------------------------------------------- pass 0: row maximum of I (the code is monotone in I)
      int mx = INT_MIN;
      for (int t = 0; t < n; ++t, ++gs) {
        const uint32_t sb = gs & 1, c8 = gs % kCkRing;
        mbar_wait(&ck_full[c8], (gs / kCkRing) & 1);
        mbar_wait(&t_full[sb], (gs >> 1) & 1);
        tc_fence_after();
        const int kind = chunk_kind(t);
        uint32_t r[32];
        tmem_ld32(t_lane + kColS + sb * kTN + cg * 32, r);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[sb]);
        const int4* ck4 = reinterpret_cast<const int4*>(s_ck + c8 * kTN + cg * 32);
        if (kind == 0) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const int4 c = ck4[j4];
            mx = max(mx, max(max((int)r[4 * j4] + c.x, (int)r[4 * j4 + 1] + c.y), max((int)r[4 * j4 + 2] + c.z, (int)r[4 * j4 + 3] + c.w)));
          }
        } else if (kind == 1) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const int4 c = ck4[j4];
            const int cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 4 * j4 + e;
              if (j <= lane) mx = max(mx, (int)r[j] + cc[e]);
            }
          }
        }
      }
      s_xi[cg * kTM + row] = mx;
      named_bar_sync(1, kSoftWarps * 32);
      mx = max(max(s_xi[row], s_xi[kTM + row]), max(s_xi[2 * kTM + row], s_xi[3 * kTM + row]));
      const int cm = score_bits(mx + rc);                   // bits of magic + (code of the row maximum - o_s)

      // ------------------------------------------- pass 1: exact row sum of E
      // (16 columns at a time: the S buffer is handed back after the second load; the other buffer is already being filled)
      unsigned long long sum = 0;
      for (int t = 0; t < n; ++t, ++gs) {
        const uint32_t sb = gs & 1, c8 = gs % kCkRing;
        mbar_wait(&ck_full[c8], (gs / kCkRing) & 1);
        mbar_wait(&t_full[sb], (gs >> 1) & 1);
        tc_fence_after();
        const int kind = chunk_kind(t);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t r[16];
          tmem_ld16(t_lane + kColS + sb * kTN + cg * 32 + hf * 16, r);
          tc_wait_ld();
          if (hf == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[sb]);
          }
          const int4* ck4 = reinterpret_cast<const int4*>(s_ck + c8 * kTN + cg * 32 + hf * 16);
          auto body = [&](auto mask_tag) {
            constexpr bool MASK = decltype(mask_tag)::value;
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const int4 c = ck4[j4];
              const int cc[4] = {c.x, c.y, c.z, c.w};
              uint32_t e4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = hf * 16 + 4 * j4 + e;
                const uint32_t ev = exp_tab(cm - score_bits((int)r[4 * j4 + e] + cc[e] + rc));
                e4[e] = (!MASK || j <= lane) ? ev : 0u;
              }
              sum += ((unsigned long long)e4[0] + e4[1]) + ((unsigned long long)e4[2] + e4[3]);
            }
          };
          if (kind == 0) body(std::false_type{}); else if (kind == 1) body(std::true_type{});
        }
      }
      s_xs[cg * kTM + row] = sum;
      named_bar_sync(1, kSoftWarps * 32);
      sum = (s_xs[row] + s_xs[kTM + row]) + (s_xs[2 * kTM + row] + s_xs[3 * kTM + row]);
      const float den = __ull2float_rn(sum);               // >= 2^31: the row maximum contributes E(0)
      const float rden = __frcp_rn(den);
      // the Markstein division needs its second step only for an all-ones significand of the divisor (common.cuh)
      const bool den_five = __any_sync(0xffffffffu, mantissa_all_ones(den));

      // ------------------------------------------- pass 2: prob codes -> P (hi / lo byte planes) for the P.V MMAs
      int ps_lo = 0, ps_hi = 0;                            // sum_j cp_ij, by byte plane (zero-point correction of V)
      for (int t = 0; t < n; ++t, ++gs, ++gp) {
        const uint32_t sb = gs & 1, c8 = gs % kCkRing, pb = gp & 1;
        mbar_wait(&ck_full[c8], (gs / kCkRing) & 1);
        mbar_wait(&t_full[sb], (gs >> 1) & 1);
        tc_fence_after();
        const int kind = chunk_kind(t);
        // row `row` of a K-major 128B-swizzled tile: 16-byte chunk c lives at chunk (c ^ (row & 7))
        uint8_t* plo = s_p + (pb * 2 + 0) * L::kPBytes + row * 128;
        uint8_t* phi = plo + L::kPBytes;
        const int sw7 = row & 7;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t r[16];
          tmem_ld16(t_lane + kColS + sb * kTN + cg * 32 + hf * 16, r);
          tc_wait_ld();
          if (hf == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[sb]);
          }
          uint32_t wlo[4] = {0u, 0u, 0u, 0u}, whi[4] = {0u, 0u, 0u, 0u};
          if (kind != 2) {
            const int4* ck4 = reinterpret_cast<const int4*>(s_ck + c8 * kTN + cg * 32 + hf * 16);
            auto body = [&](auto df_tag, auto mask_tag) {
              constexpr bool DF = decltype(df_tag)::value;
              constexpr bool MASK = decltype(mask_tag)::value;
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const int4 c = ck4[j4];
                const int cc[4] = {c.x, c.y, c.z, c.w};
                uint32_t cb[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int j = hf * 16 + 4 * j4 + e;
                  uint32_t ev = exp_tab(cm - score_bits((int)r[4 * j4 + e] + cc[e] + rc));
                  if (MASK && j > lane) ev = 0u;
                  // (magic + code) keeps the 16-bit prob code in its low half-word (o_p == 0); p >= 0: the lower clamp never binds
                  const float pr = div_rn<DF>(__uint2float_rn(ev), den, rden);
                  cb[e] = (uint32_t)__float_as_int(__fadd_rn(fminf(div_rn<FIVE>(pr, qp.s, qp.rs), qp.hi), kRoundMagic));
                }
                const uint32_t p01 = __byte_perm(cb[0], cb[1], 0x5410), p23 = __byte_perm(cb[2], cb[3], 0x5410);
                wlo[j4] = __byte_perm(p01, p23, 0x6420);
                whi[j4] = __byte_perm(p01, p23, 0x7531);
                ps_lo = (int)__dp4a(wlo[j4], 0x01010101u, (unsigned)ps_lo);
                ps_hi = (int)__dp4a(whi[j4], 0x01010101u, (unsigned)ps_hi);
              }
            };
            if (den_five) { if (kind == 0) body(std::true_type{}, std::false_type{}); else body(std::true_type{}, std::true_type{}); }
            else if (kind == 0) body(std::false_type{}, std::false_type{});
            else body(std::false_type{}, std::true_type{});
          }
          if (hf == 0) mbar_wait(&p_empty[pb], ((gp >> 1) & 1) ^ 1);        // the MMAs that read this P buffer have retired
          *reinterpret_cast<uint4*>(plo + (((2 * cg + hf) ^ sw7) << 4)) = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]);
          *reinterpret_cast<uint4*>(phi + (((2 * cg + hf) ^ sw7) << 4)) = make_uint4(whi[0], whi[1], whi[2], whi[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[pb]);
      }
      s_xi[cg * kTM + row] = ps_lo + 256 * ps_hi;
      named_bar_sync(1, kSoftWarps * 32);
      const int psum = (s_xi[row] + s_xi[kTM + row]) + (s_xi[2 * kTM + row] + s_xi[3 * kTM + row]);

      // ------------------------------------------- epilogue: this thread's OC output columns of its row
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      int csum = 0;
      uint8_t* dst = a.out + (int64_t(b) * a.Tq + qloc) * ldo + h * HD + cg * OC;
#pragma unroll
      for (int v = 0; v < OC / 16; ++v) {
        uint32_t olo[16], ohi[16];
        tmem_ld16(t_lane + kColOlo + cg * OC + v * 16, olo);
        tmem_ld16(t_lane + kColOhi + cg * OC + v * 16, ohi);
        tc_wait_ld();
        if (v == OC / 16 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_empty);
        }
        uint32_t packed[4];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          uint32_t w = 0;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int A = (int)olo[4 * j4 + e] + 256 * (int)ohi[4 * j4 + e] - iov * psum;
            w |= (uint32_t)quant_int<FIVE>(__fmul_rn(__int2float_rn(A), a.spv), qo) << (8 * e);
          }
          packed[j4] = w;
          csum = (int)__dp4a(w, 0x01010101u, (unsigned)csum);
        }
        if (row_ok) reinterpret_cast<uint4*>(dst)[v] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      }
      if (row_ok && a.rowsum_out) atomicAdd(a.rowsum_out + int64_t(b) * a.Tq + qloc, csum);
      (void)qi;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kSoftWarps + 1) tmem_dealloc(tmem_base, 512);
}

bool qattn_tc_supported(const AttnArgs& a) {
  if (a.hd != 64 && a.hd != 128) return false;
  if (a.T % 16 != 0 || a.q_start % kTM != 0) return false;                  // TMA row pitch of V^T; diagonal tile alignment
  if ((reinterpret_cast<uintptr_t>(a.q) & 15) || (reinterpret_cast<uintptr_t>(a.k) & 15) || (reinterpret_cast<uintptr_t>(a.vt) & 15) ||
      (reinterpret_cast<uintptr_t>(a.rsk) & 15) || (reinterpret_cast<uintptr_t>(a.out) & 15))
    return false;
  if ((int64_t)a.B * a.nh * a.Tq >= (1ll << 31) || (int64_t)a.B * a.nkv * a.T >= (1ll << 31)) return false;
  return true;
}

template <int HD, bool FIVE>
static int launch_tc2(Ctx* c, const AttnArgs& a, cudaStream_t st) {
  using L = TcLayout<HD>;
  CUtensorMap tq, tk, tv;
  const int64_t qrows = (int64_t)a.B * a.nh * a.Tq, krows = (int64_t)a.B * a.nkv * a.T, vrows = (int64_t)a.B * a.nkv * HD;
  if (!make_tmap_2d(&tq, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, a.q, qrows, HD, HD, kTM, HD) ||
      !make_tmap_2d(&tk, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, a.k, krows, HD, HD, kTN, HD) ||
      !make_tmap_2d(&tv, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, a.vt, vrows, a.T, a.T, HD, 128))
    return fail(c, MQ_RUNTIME_ERROR, "mq_qattn: cuTensorMapEncodeTiled failed");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(qattn_tc_kernel<HD, FIVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return fail(c, MQ_RUNTIME_ERROR, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    attr_set = true;
  }
  const long long n_items = (long long)((a.Tq + kTM - 1) / kTM) * a.B * a.nh;
  if (n_items > 0x7fffffffLL) return fail(c, MQ_INVALID_ARGUMENT, "mq_qattn: too many work items");
  const int grid = (int)std::min<long long>(n_items, (long long)c->sm_count);
  const uint32_t idesc_s = make_idesc(2u, 0u, 0u, 0u, 0u, kTM, kTN), idesc_pv = make_idesc(2u, 0u, 0u, 0u, 0u, kTM, HD);
  qattn_tc_kernel<HD, FIVE><<<grid, kTcThreads, L::kTotal, st>>>(tq, tk, tv, a, (int)n_items, idesc_s, idesc_pv);
  return check_launch(c, "mq_qattn(tc)");
}

int launch_qattn_tc(Ctx* c, const AttnArgs& a, cudaStream_t st) {
  if (!qattn_tc_supported(a)) return -1;
  const bool five = mantissa_all_ones(a.s_s) || mantissa_all_ones(a.s_p) || mantissa_all_ones(a.s_out);
  if (a.hd == 64) return five ? launch_tc2<64, true>(c, a, st) : launch_tc2<64, false>(c, a, st);
  return five ? launch_tc2<128, true>(c, a, st) : launch_tc2<128, false>(c, a, st);
}

}  // namespace mq
