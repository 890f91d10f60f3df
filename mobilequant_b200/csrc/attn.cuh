// Arguments of the quantised causal attention kernels (engine.cu: mma.sync variants; qattn_tc.cu: tcgen05 variant).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mq {

struct Ctx;

struct AttnArgs {
  const uint8_t *q, *k, *vt;
  const int32_t *rsq, *rsk;
  int B, T, nh, nkv, hd;
  float oq, ok, ov;            // integer zero points
  float sqk;                   // sq*sk
  float s_s, o_s, qmax_s;      // score quantizer
  const uint32_t* lut;         // [512]: A[256] then B[256]
  float s_p, qmax_p;           // prob quantizer (offset 0)
  float spv;                   // s_p*s_v
  float s_out, o_out;          // output quantizer (8 bit)
  uint8_t* out;                // [B*T, nh*hd]
  int32_t* rowsum_out;         // [B*T] atomically accumulated
  int q_start;                 // first query position of this call (0 = full causal prefill); keys always start at 0
  int Tq;                      // query rows per (batch, head) in q / out (== T unless the queries are a sequence shard)
};

// tcgen05 / TMA / TMEM kernel (qattn_tc.cu).  Returns -1 when the shape is not covered (caller falls back).
int launch_qattn_tc(Ctx* c, const AttnArgs& a, cudaStream_t st);
bool qattn_tc_supported(const AttnArgs& a);

}  // namespace mq
