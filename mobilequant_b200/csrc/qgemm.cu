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
// Structure (one CTA per SM, persistent over 128 x BN output tiles):
//   warp 0      TMA producer   : cp.async.bulk.tensor 128B-swizzled A/B k-slices into a kStages-deep smem ring
//   warp 1      MMA issuer     : one elected lane issues tcgen05.mma.kind::i8 (M=128, N=BN, K=32) into TMEM; owns TMEM alloc
//   warps 2..9  epilogue       : tcgen05.ld the s32 tile (2 warps per TMEM lane quarter), requantise, store
//   TMEM holds two BN-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "tc_common.cuh"
#include "ctx.h"
#include <string>

namespace mq {
using namespace tc;

constexpr int kBM = 128;          // UMMA M
constexpr int kBK = 128;          // bytes of K per stage == one 128B swizzle row
constexpr int kUmmaK = 32;        // 8-bit operands: 32 elements per MMA
constexpr int kNumEpiWarps = 8;
constexpr int kThreads = 64 + kNumEpiWarps * 32;

enum { EPI_QUANT = 0, EPI_ACTMUL = 1, EPI_RESID = 2, EPI_F32 = 3, EPI_I32 = 4 };
enum { CP_SXW = 0, CP_OW, CP_C0, CP_BIAS, CP_SO, CP_OO, CP_COUNT };

struct QGemmArgs {
  int M, N, K;
  int mode;
  const int32_t* rowsum;   // [M]
  const float* sxw;        // [N]
  const int32_t* ow;       // [N]
  const int32_t* c0;       // [N]
  const float* bias;       // [N] or null
  const float* so;         // [N] output quantizer scale
  const float* oo;         // [N] output quantizer offset
  float qmax;              // output quantizer qmax (qmin = 0: activations are asymmetric)
  int out_bits;            // 8 or 16 (EPI_QUANT)
  void* out;               // codes / fp32 / int32
  int64_t ldo;
  int32_t* rowsum_out;     // [M] atomically accumulated sum of the emitted codes (or null)
  const float* lut;        // [256] EPI_ACTMUL: act(w1 code) as fp32 (QSiLU/QGELU folded, qm:739-753)
  float s2, o2, qmax2;     // EPI_ACTMUL: w2.input_quantizer ; EPI_RESID unused
  float* resid;            // EPI_RESID: [M, ldo] fp32 residual stream, updated in place
};

template <int BN>
struct SmemLayout {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kABytes = kBM * kBK;
  static constexpr int kBBytes = BN * kBK;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kColParamBytes = 2 * CP_COUNT * BN * 4;
  static constexpr int kBarOffset = kStages * kStageBytes + kColParamBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;   // + barriers + alignment slack
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
qgemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const QGemmArgs p,
             const uint32_t idesc) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + L::kStages * L::kABytes;
  float* colp = reinterpret_cast<float*>(smem + L::kStages * L::kStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + L::kStages;
  uint64_t* tfull_bar = empty_bar + L::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + kBM - 1) / kBM, n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_iters = (p.K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int i = 0; i < L::kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], kNumEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / n_tiles) * kBM, n0 = (t % n_tiles) * BN;
        for (int k = 0; k < k_iters; ++k) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          tma_load_2d(smem_a + stage * L::kABytes, &tmap_a, &full_bar[stage], k * kBK, m0);
          tma_load_2d(smem_b + stage * L::kBBytes, &tmap_b, &full_bar[stage], k * kBK, n0);
          if (++stage == L::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
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
            mma_i8(tmem_d, adesc + uint64_t(kk * (kUmmaK >> 4)), bdesc + uint64_t(kk * (kUmmaK >> 4)), idesc,
                   (k | kk) != 0);
          }
          tc_commit(&empty_bar[stage]);                      // smem slot free once these MMAs retire
          if (k == k_iters - 1) tc_commit(&tfull_bar[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == L::kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;                 // 0..7
    const int quarter = warp & 3;            // TMEM lane quarter this warp may read
    const int half = ew >> 2;                // column half
    const int etid = threadIdx.x - 64;       // 0..255
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t / n_tiles) * kBM, n0 = (t % n_tiles) * BN;
      // stage the per-column parameters of this tile (double buffered with the accumulator)
      float* cp = colp + acc * CP_COUNT * BN;
      for (int c = etid; c < BN; c += kNumEpiWarps * 32) {
        const int n = n0 + c;
        const bool ok = n < p.N;
        cp[CP_SXW * BN + c] = ok ? __ldg(p.sxw + n) : 0.f;
        reinterpret_cast<int*>(cp)[CP_OW * BN + c] = ok ? __ldg(p.ow + n) : 0;
        reinterpret_cast<int*>(cp)[CP_C0 * BN + c] = ok ? __ldg(p.c0 + n) : 0;
        cp[CP_BIAS * BN + c] = (ok && p.bias) ? __ldg(p.bias + n) : 0.f;
        cp[CP_SO * BN + c] = (ok && p.so) ? __ldg(p.so + n) : 1.f;
        cp[CP_OO * BN + c] = (ok && p.oo) ? __ldg(p.oo + n) : 0.f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kNumEpiWarps * 32) : "memory");
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();

      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const int rs = row_ok ? __ldg(p.rowsum + row) : 0;
      const uint32_t trow = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN;
      const int* cpi = reinterpret_cast<const int*>(cp);
      int code_sum = 0;

      if (p.mode == EPI_ACTMUL) {
        // columns [0,BN/2) of the tile are w1 rows, [BN/2,BN) the matching w3 rows
        constexpr int H = BN / 2;
        for (int cc = half * (H / 2); cc < (half + 1) * (H / 2); cc += 32) {
          uint32_t r1[32], r3[32];
          tmem_ld32(trow + cc, r1);
          tmem_ld32(trow + H + cc, r3);
          tc_wait_ld();
          uint32_t packed[8];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c1 = cc + j, c3 = H + cc + j;
            const int i1 = (int)r1[j] - cpi[CP_OW * BN + c1] * rs + cpi[CP_C0 * BN + c1];
            const int i3 = (int)r3[j] - cpi[CP_OW * BN + c3] * rs + cpi[CP_C0 * BN + c3];
            const float y1 = fadd(fmul(__int2float_rn(i1), cp[CP_SXW * BN + c1]), cp[CP_BIAS * BN + c1]);
            const float y3 = fadd(fmul(__int2float_rn(i3), cp[CP_SXW * BN + c3]), cp[CP_BIAS * BN + c3]);
            const float q1 = quant_code(y1, cp[CP_SO * BN + c1], cp[CP_OO * BN + c1], 0.f, p.qmax);
            const float q3 = quant_code(y3, cp[CP_SO * BN + c3], cp[CP_OO * BN + c3], 0.f, p.qmax);
            const float a = __ldg(p.lut + (int)q1);                                   // fq_out(act(w1x))
            const float u = dequant(q3, cp[CP_SO * BN + c3], cp[CP_OO * BN + c3]);  // fq(w3x)
            const int code = (int)quant_code(fmul(a, u), p.s2, p.o2, 0.f, p.qmax2); // w2.input_quantizer
            code_sum += code;
            if ((j & 3) == 0) packed[j >> 2] = 0;
            packed[j >> 2] |= (uint32_t)code << (8 * (j & 3));
          }
          const int ncol = n0 / 2 + cc;       // output column (N/2 wide)
          if (row_ok) {
            uint8_t* dst = reinterpret_cast<uint8_t*>(p.out) + int64_t(row) * p.ldo + ncol;
            if (ncol + 32 <= p.N / 2 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
              reinterpret_cast<uint4*>(dst)[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
              reinterpret_cast<uint4*>(dst)[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
            } else {
              for (int j = 0; j < 32; ++j)
                if (ncol + j < p.N / 2) dst[j] = (uint8_t)(packed[j >> 2] >> (8 * (j & 3)));
            }
          }
        }
      } else {
        for (int cc = half * (BN / 2); cc < (half + 1) * (BN / 2); cc += 32) {
          if (n0 + cc >= p.N) break;
          uint32_t r[32];
          tmem_ld32(trow + cc, r);
          tc_wait_ld();
          float y[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = cc + j;
            const int ii = (int)r[j] - cpi[CP_OW * BN + c] * rs + cpi[CP_C0 * BN + c];
            y[j] = fadd(fmul(__int2float_rn(ii), cp[CP_SXW * BN + c]), cp[CP_BIAS * BN + c]);
            if (p.mode == EPI_I32) y[j] = __int_as_float(ii);
          }
          const int ncol = n0 + cc;
          const bool full = (ncol + 32 <= p.N);
          if (p.mode == EPI_QUANT) {
            uint32_t q[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              q[j] = (uint32_t)quant_code(y[j], cp[CP_SO * BN + cc + j], cp[CP_OO * BN + cc + j], 0.f, p.qmax);
              if (ncol + j < p.N) code_sum += (int)q[j];
            }
            if (row_ok) {
              if (p.out_bits == 8) {
                uint8_t* dst = reinterpret_cast<uint8_t*>(p.out) + int64_t(row) * p.ldo + ncol;
                if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                  uint32_t w[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) w[j] = q[4 * j] | (q[4 * j + 1] << 8) | (q[4 * j + 2] << 16) | (q[4 * j + 3] << 24);
                  reinterpret_cast<uint4*>(dst)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                  reinterpret_cast<uint4*>(dst)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                } else {
                  for (int j = 0; j < 32; ++j) if (ncol + j < p.N) dst[j] = (uint8_t)q[j];
                }
              } else {
                uint16_t* dst = reinterpret_cast<uint16_t*>(p.out) + int64_t(row) * p.ldo + ncol;
                if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                  for (int v = 0; v < 4; ++v) {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) w[j] = q[8 * v + 2 * j] | (q[8 * v + 2 * j + 1] << 16);
                    reinterpret_cast<uint4*>(dst)[v] = make_uint4(w[0], w[1], w[2], w[3]);
                  }
                } else {
                  for (int j = 0; j < 32; ++j) if (ncol + j < p.N) dst[j] = (uint16_t)q[j];
                }
              }
            }
          } else if (p.mode == EPI_RESID) {
            if (row_ok) {
              float* dst = p.resid + int64_t(row) * p.ldo + ncol;
              if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                float4 h4[8];
#pragma unroll
                for (int v = 0; v < 8; ++v) h4[v] = reinterpret_cast<const float4*>(dst)[v];
                float* hv = reinterpret_cast<float*>(h4);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float so = cp[CP_SO * BN + cc + j], oo = cp[CP_OO * BN + cc + j];
                  hv[j] = fadd(hv[j], dequant(quant_code(y[j], so, oo, 0.f, p.qmax), so, oo));   // hm:1257,1270
                }
#pragma unroll
                for (int v = 0; v < 8; ++v) reinterpret_cast<float4*>(dst)[v] = h4[v];
              } else {
                for (int j = 0; j < 32; ++j) {
                  if (ncol + j < p.N) {
                    const float so = cp[CP_SO * BN + cc + j], oo = cp[CP_OO * BN + cc + j];
                    dst[j] = fadd(dst[j], dequant(quant_code(y[j], so, oo, 0.f, p.qmax), so, oo));
                  }
                }
              }
            }
          } else {   // EPI_F32 / EPI_I32
            if (row_ok) {
              float* dst = reinterpret_cast<float*>(p.out) + int64_t(row) * p.ldo + ncol;
              if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                for (int v = 0; v < 8; ++v)
                  reinterpret_cast<float4*>(dst)[v] = make_float4(y[4 * v], y[4 * v + 1], y[4 * v + 2], y[4 * v + 3]);
              } else {
                for (int j = 0; j < 32; ++j) if (ncol + j < p.N) dst[j] = y[j];
              }
            }
          }
        }
      }
      if (p.rowsum_out && row_ok && (p.mode == EPI_QUANT || p.mode == EPI_ACTMUL)) atomicAdd(p.rowsum_out + row, code_sum);
      // release the accumulator back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

// ---- host -----------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2D row-major byte matrix [rows, cols] (cols contiguous), box = 128 bytes x box_rows, 128B swizzle
static bool make_tmap_u8(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols};
  cuuint32_t box[2] = {128u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN>
static int launch_qgemm(Ctx* c, const void* a, const void* b, const QGemmArgs& args, int a_signed, int b_signed, cudaStream_t st) {
  CUtensorMap ta, tb;
  if (!make_tmap_u8(&ta, a, args.M, args.K, kBM) || !make_tmap_u8(&tb, b, args.N, args.K, BN))
    return fail(c, MQ_RUNTIME_ERROR, "cuTensorMapEncodeTiled failed (pointers must be 16B aligned, K a multiple of 16)");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(qgemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<BN>::kTotal);
    if (e != cudaSuccess) return fail(c, MQ_RUNTIME_ERROR, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    attr_set = true;
  }
  const int m_tiles = (args.M + kBM - 1) / kBM, n_tiles = (args.N + BN - 1) / BN;
  int grid = m_tiles * n_tiles;
  if (grid > c->sm_count) grid = c->sm_count;
  const uint32_t idesc = make_idesc(2u, a_signed ? 1u : 0u, b_signed ? 1u : 0u, 0u, 0u, kBM, BN);
  qgemm_kernel<BN><<<grid, kThreads, SmemLayout<BN>::kTotal, st>>>(ta, tb, args, idesc);
  return check_launch(c, "mq_qgemm");
}

}  // namespace mq

using namespace mq;

extern "C" int mq_qgemm(void* ctx, const void* a_codes, int a_signed, const void* b_codes, int b_signed, int M, int N, int K,
                        const int32_t* rowsum, const float* sxw, const int32_t* ow, const int32_t* c0, const float* bias,
                        int mode, const float* so, const float* oo, float qmax, int out_bits, void* out, int64_t ldo,
                        int32_t* rowsum_out, const float* lut, float s2, float o2, float qmax2, float* resid, void* stream) {
  MQ_CTX(c, ctx);
  MQ_REQUIRE(c, a_codes && b_codes && M > 0 && N > 0 && K > 0, "null operand or empty problem");
  MQ_REQUIRE(c, K % 16 == 0, "K must be a multiple of 16 bytes (TMA row pitch)");
  MQ_REQUIRE(c, (reinterpret_cast<uintptr_t>(a_codes) & 15) == 0 && (reinterpret_cast<uintptr_t>(b_codes) & 15) == 0,
             "operands must be 16-byte aligned");
  MQ_REQUIRE(c, rowsum && sxw && ow && c0, "rowsum / sxw / ow / c0 are required");
  MQ_REQUIRE(c, mode >= EPI_QUANT && mode <= EPI_I32, "unknown epilogue mode");
  MQ_REQUIRE(c, mode != EPI_QUANT || (out && so && oo && (out_bits == 8 || out_bits == 16)), "EPI_QUANT needs out/so/oo, 8 or 16 bits");
  MQ_REQUIRE(c, mode != EPI_ACTMUL || (out && so && oo && lut && N % 256 == 0), "EPI_ACTMUL needs out/so/oo/lut and N % 256 == 0");
  MQ_REQUIRE(c, mode != EPI_RESID || (resid && so && oo), "EPI_RESID needs resid/so/oo");
  MQ_REQUIRE(c, (mode != EPI_F32 && mode != EPI_I32) || out, "raw output needs out");
  QGemmArgs args;
  args.M = M; args.N = N; args.K = K; args.mode = mode; args.rowsum = rowsum; args.sxw = sxw; args.ow = ow; args.c0 = c0;
  args.bias = bias; args.so = so; args.oo = oo; args.qmax = qmax; args.out_bits = out_bits; args.out = out; args.ldo = ldo;
  args.rowsum_out = rowsum_out; args.lut = lut; args.s2 = s2; args.o2 = o2; args.qmax2 = qmax2; args.resid = resid;
  return launch_qgemm<256>(c, a_codes, b_codes, args, a_signed, b_signed, (cudaStream_t)stream);
}
