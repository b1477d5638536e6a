// Persistent tensor-core GRU recurrence (sm_100a): all T steps of a GRU for up to 64 sequences in ONE cooperative launch
// -- the policy rollout of ACT/models/ppo.py:67-96 / ACT/models/gfv_net.py:110 and the classifier GRU of
// ACT/models/gfv_net.py:427-435 at bench batch sizes, where the per-step form (h W_hh^T GEMM launch + gate kernel
// launch, 17-24 us per step) is bound by launch / pipeline-ramp latency and by re-streaming W_hh from L2 every step.
//
//   * CTA c owns hidden units [8c, 8c+8): its 24 rows of W_hh (gates r, z, n) stay in shared memory for the whole
//     sequence as K-major SWIZZLE_128B UMMA operand tiles (N = 32 with 8 zero rows); H / 8 CTAs (128 for H = 1024);
//   * per step the CTA streams h_{t-1} of ALL units (64 sequences x H, fp16; TMA, 8-stage ring) as the A operand and
//     accumulates D[64 sequences, 24 gate rows] in TMEM (tcgen05.mma, M = 128 with the upper 64 rows unused);
//   * two epilogue warps (thread = sequence) keep their 8 h values in fp32 registers across the steps, apply the gate
//     math (torch.nn.GRU, gates r, z, n), publish the fp16 operand rows of h_t for everybody's next step and emit h_t
//     for the consumers of the sequence; a grid-wide barrier (atomic counter, all CTAs co-resident) separates the steps;
//   * SPLIT: split-precision operands (engine.pack_conv_split): weights [W_hi | W_hi | W_lo], h as [hi | lo]; per h_hi
//     k-block the MMAs run against W_hi and W_lo, per h_lo k-block against W_hi: x_hi W_hi + x_hi W_lo + x_lo W_hi.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace af {

constexpr int kGruTcThreads = 192;   // warp 0: TMA producer, warp 1: MMA issuer, warps 4-5: gate math (TMEM lanes 0..63)
constexpr int kGruTcStages = 8;

struct GruTcParams {
  const float* xg;        // [B*T][3H] fp32, rows b*T+t: W_ih x + b_ih
  const float* b_hh;      // [3H]
  const float* h0;        // [B][H] fp32 or nullptr (zeros)
  __half* hbuf;           // scratch [2][B][(SPLIT ? 2 : 1) * H] fp16: operand rows of h, ping-pong between steps
  __half* hseq;           // h_t rows b*T+t, row stride hseq_stride; SPLIT: [hi | lo | hi] (3H wide)
  long long hseq_stride;
  float* h_out;           // [B][H] final state or nullptr
  unsigned int* counter;  // grid-barrier counter (zeroed by the launcher)
  int B, T, H;
};

struct GruTcMaps {
  CUtensorMap w;      // packed W_hh (+ split parts) [3H rows][Kw] fp16, box {64, 8}
  CUtensorMap h[2];   // hbuf[i] as {Kh, B}, box {64, 64}: sequences past B read as zeros
};

bool gru_tc_supported(int B, int H, int sm_count, int split);
size_t gru_tc_smem_bytes(int H, int split);
cudaError_t launch_gru_tc(const GruTcMaps& maps, const GruTcParams& p, int split, cudaStream_t s);

}  // namespace af
