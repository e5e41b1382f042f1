// tokens.cu -- device-side coefficient-token decoder (the DCT partitions of a VP8 frame).
//
// Replaces, for throughput-oriented callers (vp8gpu_decode_ivf), the token half of the CPU front
// end: Frame::parse_tokens -> Macroblock::parse_tokens -> Block::parse_tokens (decoder/frame.cc:
// 122-137, macroblock.cc:468-502, tokens.cc:50-135) over BoolDecoder (bool_decoder.hh:82-107).
// The arithmetic code of one partition is inherently serial, and a macroblock row needs the
// "has non-zero" context of the row above, so ONE thread walks one frame in raster order; the
// parallelism is across frames: token partitions do not depend on pixels, so the host parses the
// first partitions of many frames (cheap: ~12 % of the bytes) and launches this kernel over all of
// them while the pixel kernels work on earlier frames.  One warp per frame: lane 0 decodes, the
// other lanes only help to stage the probability table.
//
// In:  vp8gpu_mb records written by the host with y_mode and VP8GPU_MB_SKIP (mb_skip_coeff) set,
//      the frame's coefficient probabilities (after the header's updates), the raw partitions.
// Out: the same token stream the CPU front end emits (csrc/parser.cc parse_block) and, per
//      record, tok_off / tok_cnt; VP8GPU_MB_SKIP is cleared, so the records end up byte-identical
//      to the ones the CPU path produces.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "engine.h"
#include "tokens_core.cuh"

namespace vp8 {
namespace {

constexpr int kMaxCols = 1024;  // 16383 px / 16
// jobs live at the start of equally spaced slots of a ring (engine.hpp TokenRing); one warp per frame,
// kTokWarps frames per CTA (measured: one-warp CTAs spread over the SMs best; VP8GPU_TOK_WARPS=8 packs them)
template <int kTokWarps>
__global__ void __launch_bounds__(32 * kTokWarps) k_tokens(const uint8_t* ring, size_t stride, int first, int count,
                                                          int nslots, Geom g) {
  __shared__ __align__(16) uint8_t probs_all[kTokWarps][tok::kProbBytes];
  __shared__ uint16_t above_all[kTokWarps][kMaxCols];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int job = blockIdx.x * kTokWarps + warp;
  if (job >= count) return;
  uint8_t* probs = probs_all[warp];
  uint16_t* above_nz = above_all[warp];
  const TokJob& J = *reinterpret_cast<const TokJob*>(ring + static_cast<size_t>((first + job) % nslots) * stride);
  for (int e = lane; e < tok::kProbEntries; e += 32) tok::expand_prob_entry(J.coef_probs, probs, e);
  for (int i = lane; i < g.mb_cols; i += 32) above_nz[i] = 0;
  __syncwarp();
  if (lane != 0) return;
  tok::decode_frame_tokens(J, g, probs, above_nz);
}

// lock-step variant: one LANE per frame, 32 frames per warp (tokens_core.cuh decode_frame_tokens_lockstep)
__global__ void __launch_bounds__(32) k_tokens_lockstep(const uint8_t* ring, size_t stride, int first, int count,
                                                         int nslots, Geom g) {
  __shared__ tok::LockstepTables T;
  const int lane = threadIdx.x;
  tok::fill_lockstep_tables(T, lane, 32);
  __syncwarp();
  const int job = blockIdx.x * 32 + lane;
  if (job >= count) return;
  const TokJob& J = *reinterpret_cast<const TokJob*>(ring + static_cast<size_t>((first + job) % nslots) * stride);
  tok::decode_frame_tokens_lockstep(J, g, T);
}

}  // namespace

int launch_tokens(const uint8_t* ring, size_t stride, int first, int count, int nslots, const Geom& g, void* stream) {
  if (g.mb_cols > kMaxCols) return (int)cudaErrorInvalidValue;
  static const int warps = [] {  // tuning knob: frames per CTA (1 or 8); 32 = one lane per frame
    const char* v = getenv("VP8GPU_TOK_WARPS");
    return v && atoi(v) == 8 ? 8 : (v && atoi(v) == 32 ? 32 : 1);
  }();
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (warps == 32) k_tokens_lockstep<<<(count + 31) / 32, 32, 0, s>>>(ring, stride, first, count, nslots, g);
  else if (warps == 8) k_tokens<8><<<(count + 7) / 8, 256, 0, s>>>(ring, stride, first, count, nslots, g);
  else k_tokens<1><<<count, 32, 0, s>>>(ring, stride, first, count, nslots, g);
  return (int)cudaGetLastError();
}

}  // namespace vp8
