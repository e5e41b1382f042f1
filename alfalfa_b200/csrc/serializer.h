// serializer.h -- host-side VP8 bitstream writer for the encoder path.
//
// Replaces Frame::serialize (encoder/serializer.cc:388-829) + BoolEncoder (encoder/bool_encoder.hh:60-152)
// for frames described by the flat records of include/vp8gpu.h: the device emits, per macroblock,
// the chosen modes / motion vector and the quantised non-zero coefficients (tokens); this module
// turns them into a standard VP8 frame (RFC 6386): frame tag, frame header, per-macroblock modes
// with the same context modelling the decoder uses (mode contexts from the MV census, B_PRED
// contexts, token contexts), one DCT partition.  It stays on the CPU.
//
// Subset written: no segmentation, no loop-filter deltas, one token partition, LAST reference
// only, default probabilities with optional per-frame token-probability updates.
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/vp8gpu.h"

namespace vp8 {

struct EncodeHeader {
  bool key_frame = true, show_frame = true;
  int width = 0, height = 0;
  int y_ac_qi = 0;            // QuantIndices::y_ac_qi, all deltas zero
  int loop_filter_level = 0;  // 0..63
  int sharpness = 0;          // 0..7
  bool optimize_token_probs = false;  // per-frame coefficient probability updates (encoder.cc:419-440)
};

// RFC 6386 section 7 arithmetic encoder (same code stream as encoder/bool_encoder.hh)
class BoolWriter {
 public:
  void put(int bit, int prob = 128);
  void literal(int value, int width);
  std::vector<uint8_t> finish();
  size_t size_estimate() const { return out_.size(); }

 private:
  void add_one();
  std::vector<uint8_t> out_;
  uint32_t range_ = 255, bottom_ = 0;
  int bit_count_ = 24;
};

// mbs: mb_cols*mb_rows records (y_mode, uv_mode, ref_frame, mv / split / b_modes, tok_off, tok_cnt);
// tokens: (block, raster position, value) of every non-zero quantised coefficient.
// Returns the compressed frame, or an empty vector if a record cannot be represented
// (e.g. a reference other than LAST, a motion vector out of range).
std::vector<uint8_t> serialize_frame(const EncodeHeader& h, const vp8gpu_mb* mbs, const vp8gpu_token* tokens,
                                     const vp8gpu_split_mvs* split);

}  // namespace vp8
