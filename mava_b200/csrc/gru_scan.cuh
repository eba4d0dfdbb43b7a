// Host interface of the persistent GRU scan kernels (gru_scan.cu), used by rnn_f32.cu.
#pragma once
#include "common.cuh"

namespace mava {

// One network's buffers of a sequence minibatch (fp32, the layout of the per-step schedule):
// forward reads Gx [L][S][3H] and position 0 of Hin, writes Hout [L][S][H], the gate stash
// [L][S][4H] and positions 1.. of Hin; backward reads dHout (in Hout), the stash and Hin, and
// overwrites the stash with dGh = [da_r | da_z | dq | .] and Gx with dGx = [da_r | da_z | da_n].
struct GruScanNet {
  const float* Wh;    // [H][3H]
  const float* b_hn;  // [H]
  float *Gx, *gates, *Hin, *Hout;
  const int32_t* steps;    // [L][Senv]: env-step of (position, env-sequence)
  const uint8_t* done_in;  // by env-step
  int64_t S, Senv;         // rows per position, env-sequences per position (S = Senv * rpe)
  int rpe, L;
};

int launch_gru_scan_fwd(const GruScanNet* nets, int n, cudaStream_t s);
int launch_gru_scan_bwd(const GruScanNet* nets, int n, cudaStream_t s);

}  // namespace mava
