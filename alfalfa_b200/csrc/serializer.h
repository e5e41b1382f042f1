// serializer.h -- host-side VP8 bitstream writer for the encoder path.
//
// Replaces Frame::serialize (encoder/serializer.cc:388-829) + BoolEncoder (encoder/bool_encoder.hh:60-152)
// for frames described by the flat records of include/vp8gpu.h: the device emits, per macroblock,
// the chosen modes / motion vector and the quantised non-zero coefficients (tokens); this module
// turns them into a standard VP8 frame (RFC 6386): frame tag, frame header, per-macroblock modes
// with the same context modelling the decoder uses (mode contexts from the MV census, B_PRED
// contexts, token contexts), one DCT partition.  It stays on the CPU.
//
// Written by the encoder path: no segmentation, no loop-filter deltas, one token partition, LAST
// reference only, default probabilities with optional per-frame token-probability updates.  With an
// EncodeFeatures block the writer also covers the rest of the format (segmentation, loop-filter deltas,
// quantiser deltas, 1-8 token partitions, golden / altref references with sign bias, reference
// refresh / copy flags, persistent probabilities) -- used to synthesise feature-complete test streams
// (bitstream B); the truth for those is the reference decoder.
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/vp8gpu.h"
#include "parser.h"

namespace vp8 {

struct EncodeHeader {
  bool key_frame = true, show_frame = true;
  int width = 0, height = 0;
  int y_ac_qi = 0;            // QuantIndices::y_ac_qi, all deltas zero
  int loop_filter_level = 0;  // 0..63
  int sharpness = 0;          // 0..7
  bool optimize_token_probs = false;  // per-frame coefficient probability updates (encoder.cc:419-440)
};

// frame_header.hh:37-131, 213-325: everything the plain header leaves at its default
struct EncodeFeatures {
  int log2_partitions = 0;  // 0..3
  bool segmentation_enabled = false, update_mb_segmentation_map = false, update_segment_feature_data = false;
  bool segment_feature_absolute = false;
  int segment_quant[4] = {0, 0, 0, 0}, segment_lf[4] = {0, 0, 0, 0};
  int segment_tree_probs[3] = {255, 255, 255};  // 255 = not sent
  bool lf_delta_enabled = false, lf_delta_update = false;
  int ref_lf_delta[4] = {0, 0, 0, 0}, mode_lf_delta[4] = {0, 0, 0, 0};
  int y_dc_delta = 0, y2_dc_delta = 0, y2_ac_delta = 0, uv_dc_delta = 0, uv_ac_delta = 0;  // -15..15
  bool refresh_golden = false, refresh_alternate = false, refresh_last = true, refresh_entropy_probs = false;
  int copy_to_golden = 0, copy_to_alternate = 0;  // 0 none, 1 last, 2 the other one
  bool sign_bias_golden = false, sign_bias_alternate = false;
  // the stream's saved coefficient probabilities (DecoderState): read as the base of this frame's
  // updates and written back when refresh_entropy_probs is set (key frames reset it to the defaults
  // first); nullptr = stateless writer, every frame relative to the default tables
  uint8_t* saved_coef_probs = nullptr;
  // Reference-encoder writer policy (encoder/encoder.cc:419-457, encode_inter.cc:527-576): what the reference's
  // Encoder puts into a header is decided by its own rules, not by what is cheapest -- every token probability whose
  // estimate calc_prob( falses, total ) differs from the current one is sent (counted over the Y / U / V blocks of ALL
  // macroblocks, skipped ones included, Y2 blocks never: serializer.cc:456-587), flags set in an earlier frame stay set
  // in the Encoder's frame object, prob_skip / prob_inter / prob_references_last are calc_prob values that may be 0
  // or left over from the previous frame, and a block of explicit zero loop-filter deltas is written
  // (encoder.cc:464-470).  With ref_writer set the frame is written by exactly those rules, which makes the
  // device encoder's output byte-identical to the reference encoder's; the struct is the state of the
  // reference's per-type frame object and must persist between frames of one encoder.
  struct RefWriterState {
    uint8_t upd_flag[1056] = {0}, upd_val[1056] = {0};  // token_prob_update as left by earlier frames
    int prob_inter = 0, prob_last = 0, prob_golden = 0; // InterFrameHeader fields start at 0 (zero_decoder) and persist
  };
  RefWriterState* ref_writer = nullptr;
  // Frame::serialize of a PARSED frame (encoder/serializer.cc:388-405; gate: src/tests/roundtrip.cc:93-112):
  // header decisions and ambiguous labels come from the parse (parser.h Verbatim), everything else -- modes,
  // vectors, token partitions with their contexts -- is written from the flat records.  Set by serialize_parsed.
  const Verbatim* verbatim = nullptr;
  bool ref_estimate = false;  // with ref_writer: a size estimate (size_estimation.cc): no token-probability updates
  // A frame built by Encoder::update_residues (encoder/reencode.cc:248-313): a NEW InterFrame object whose header
  // takes the segmentation update, filter type / level / sharpness, loop-filter adjustments, quantiser deltas,
  // sign biases, refresh_entropy_probs and prob_references_* of the frame it started from (replayed from that
  // frame's tape, parser.h Verbatim marks) -- the reference flags too unless residue_refresh_all (last_frame) --
  // while one token partition, prob_skip_false, prob_inter, the token probability updates and the absence of mode /
  // motion-vector probability updates are what the reference Encoder decides (ref_writer must point to a FRESH
  // state whose prob_last / prob_golden were preset to the source frame's).  The SPLITMV layouts and sub-block
  // labels are the source frame's (the macroblock headers are copied, reencode.cc:141).
  const Verbatim* residue_of = nullptr;
  bool residue_refresh_all = false;
  // A frame built by Encoder::reencode_as_interframe (encoder/reencode.cc:39-129) from the key frame `from_key`: the
  // reference Encoder's inter-frame header (ref_writer, fresh) with the key frame's quantiser deltas (replayed from
  // its tape) and sharpness (EncodeHeader), all three references refreshed, and intra_16x16_prob /
  // intra_chroma_prob sent explicitly with their default values (reencode.cc:66-75)
  const Verbatim* from_key = nullptr;
  // EncodeHeader::loop_filter_level supplied late: called (once) where the header writes the level, i.e. after the
  // token partitions have been recorded and coded -- nothing before that point depends on the level, so the encoder
  // runs its loop-filter search (device) while this function works (host) and answers here (encoder.cu encode_final)
  int (*late_loop_filter_level)(void* ctx) = nullptr;
  void* late_ctx = nullptr;
  // the stream's saved mode / motion-vector probabilities (DecoderState) that macroblock headers are coded with;
  // nullptr = the default tables (an Encoder that started from a key frame never changes them)
  const uint8_t* ymode_probs = nullptr;
  const uint8_t* uvmode_probs = nullptr;
  const uint8_t (*mv_probs)[19] = nullptr;
};

// RFC 6386 section 7 arithmetic encoder (same code stream as encoder/bool_encoder.hh)
class BoolWriter {
 public:
  // RFC 6386 section 7.3.  `low_` holds the bits of the interval's lower end that may still change: 24 bits below
  // the byte that goes out next; count_ counts the shifts until that byte is complete (from -24).  One
  // renormalisation step per decision (count-leading-zeros), one branch when a byte leaves; a carry out of `low_`
  // ripples through the bytes already written (at most once per output byte).
  inline void put(int bit, int prob = 128) {
    const uint32_t split = 1 + (((range_ - 1) * static_cast<uint32_t>(prob)) >> 8);
    uint32_t range = split, low = low_;
    if (bit) {
      low += split;
      range = range_ - split;
    }
    int shift = __builtin_clz(range) - 24;
    range <<= shift;
    count_ += shift;
    if (count_ >= 0) {
      const int offset = shift - count_;
      if ((low << (offset - 1)) & 0x80000000u) add_one();
      out_.push_back(static_cast<uint8_t>(low >> (24 - offset)));
      low <<= offset;
      shift = count_;
      low &= 0xFFFFFFu;
      count_ -= 8;
    }
    low_ = low << shift;
    range_ = range;
  }
  void literal(int value, int width);
  std::vector<uint8_t> finish();
  size_t size_estimate() const { return out_.size(); }
  void reserve(size_t bytes) { out_.reserve(bytes); }

 private:
  void add_one();
  std::vector<uint8_t> out_;
  uint32_t range_ = 255, low_ = 0;
  int count_ = -24;
};

// mbs: mb_cols*mb_rows records (y_mode, uv_mode, ref_frame, mv / split / b_modes, tok_off, tok_cnt);
// tokens: (block, raster position, value) of every non-zero quantised coefficient.
// Returns the compressed frame, or an empty vector if a record cannot be represented
// (e.g. a reference other than LAST, a motion vector out of range).
std::vector<uint8_t> serialize_frame(const EncodeHeader& h, const vp8gpu_mb* mbs, const vp8gpu_token* tokens,
                                     const vp8gpu_split_mvs* split, const EncodeFeatures* features = nullptr);

// Frame::serialize( probability_tables ) of a frame that parse_frame produced with keep_verbatim
// (encoder/serializer.cc:388-405, the reference's own round-trip gate src/tests/roundtrip.cc:93-112): the
// result equals the parsed input byte for byte.  Empty vector if the frame was not parsed with keep_verbatim.
std::vector<uint8_t> serialize_parsed(const ParsedFrame& frame);

}  // namespace vp8
