// parser.h -- CPU entropy front end of the B200 VP8 pipeline.
//
// Replaces, on the host, what the reference does in DecoderState::parse_and_apply
// (decoder/decoder_state.hh:73-167): bool-decode the first partition (frame header, macroblock
// modes, motion vectors) and the DCT partitions (coefficient tokens).  Instead of the
// reference's per-frame object graph (TwoD<Macroblock>, 25 Block objects per macroblock,
// decoder/frame.hh:56-61) it emits the flat records of include/vp8gpu.h directly into
// (optionally pinned) staging memory that is copied to HBM as is: one 32-byte vp8gpu_mb per
// macroblock, one 32-bit token per non-zero coefficient, one 64-byte entry per SPLITMV
// macroblock.  Header and token parsing are fused into a single raster pass that only keeps
// one row of neighbour context.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "../../include/vp8gpu.h"

namespace vp8 {

// DecoderState (decoder/decoder.hh:190-225) as a plain copyable value.
struct State {
  int width = 0, height = 0, mb_cols = 0, mb_rows = 0;
  uint8_t coef_probs[1056];
  uint8_t ymode_probs[4];
  uint8_t uvmode_probs[3];
  uint8_t mv_probs[2][19];
  bool seg_enabled = false, seg_abs = false;  // Optional<Segmentation>
  int8_t seg_quant[4] = {0, 0, 0, 0}, seg_lf[4] = {0, 0, 0, 0};
  std::vector<uint8_t> seg_map;               // mb_cols * mb_rows
  bool lf_adj_enabled = false;                // Optional<FilterAdjustments>
  int8_t ref_adj[4] = {0, 0, 0, 0}, mode_adj[4] = {0, 0, 0, 0};

  State(int w, int h);
  void reset_probs();
  bool operator==(const State& o) const;  // DecoderState::operator==, decoder.cc:257-264
  uint64_t hash() const;
  // DecoderState::serialize / deserialize (decoder.cc:283-330) in the reference's tag-length-value format
  // (enc_state_serializer.hh): byte-compatible with the reference's EncoderStateSerializer / Deserializer
  std::vector<uint8_t> serialize() const;
  static bool deserialize(const uint8_t* data, size_t len, State& out, size_t* used = nullptr);
};

// Growable array whose storage comes from a pluggable allocator, so that the engine can
// hand the parser pinned host memory (cudaHostAlloc) and DMA it without a staging copy.
struct Allocator {
  void* (*alloc)(size_t);
  void (*free)(void*);
};
extern const Allocator kMallocAllocator;

template <class T>
class Buffer {
 public:
  explicit Buffer(const Allocator& a) : a_(a) {}
  ~Buffer() { if (p_) a_.free(p_); }
  Buffer(const Buffer&) = delete;
  Buffer& operator=(const Buffer&) = delete;
  T* data() { return p_; }
  const T* data() const { return p_; }
  size_t capacity() const { return cap_; }
  // keeps the first `keep` elements
  bool reserve(size_t n, size_t keep);
 private:
  Allocator a_;
  T* p_ = nullptr;
  size_t cap_ = 0;
};

// What is left to do when parse_frame ran with defer_tokens: the DCT partitions, decoded on the
// device by csrc/tokens.cu (Frame::parse_tokens, frame.cc:122-137).
struct TokenWork {
  bool deferred = false;
  uint32_t nparts = 0;
  uint32_t part_off[8] = {0}, part_len[8] = {0};  // relative to `bits`
  const uint8_t* bits = nullptr;                  // points into the caller's frame data
  uint32_t bits_len = 0;
  uint8_t coef_probs[1056];                       // the frame's probabilities (after header updates)
};

// What the reference's Frame object keeps beyond decoded values, and Frame::serialize
// (encoder/serializer.cc:388-405) therefore writes back unchanged: the header exactly as coded (a
// Flagged<Signed<n>> field distinguishes "absent" from a flagged +0 / -0, frame_header.hh:37-131) and every
// label whose decoded meaning is ambiguous (mb_skip_coeff of a macroblock without coefficients, the
// SPLITMV layout and LEFT4x4 / ABOVE4x4 / ZERO4x4 / NEW4x4 of each partition).  Filled by parse_frame when
// ParsedFrame::keep_verbatim is set; consumed by serialize_parsed (serializer.h), which reproduces
// the input frame byte for byte (the reference's gate: src/tests/roundtrip.cc:93-112).
struct Verbatim {
  std::vector<uint16_t> header_tape;  // every frame-header decision, in order: (probability << 1) | bit
  bool key = false, show = false;
  int width = 0, height = 0, log2_parts = 0;
  bool has_skip_prob = false, read_segment = false, sign_golden = false, sign_alt = false;
  uint8_t skip_prob = 0, prob_inter = 0, prob_last = 0, prob_golden = 0;
  uint8_t seg_tree_probs[3] = {255, 255, 255};
  uint8_t coef[1056], ymode[4], uvmode[3], mv[2][19];  // the frame's probability tables (after the header's updates)
  std::vector<uint8_t> mb_coded;     // per macroblock: bit 0 = mb_skip_coeff as coded, bits 1-2 = SPLITMV layout
  // What Encoder::update_residues / reencode_as_interframe (encoder/reencode.cc:39-64, 248-281) copy from the
  // frame they start from into the header of the frame they build: whole header sections, replayed from the
  // tape so that a Flagged<Signed> +0 / -0 survives -- decisions [mark_begin, mark_parts) are the segmentation
  // update, filter type / level / sharpness and the loop-filter adjustments, [mark_qdelta, mark_qend) the five
  // quantiser deltas -- and single fields, kept decoded.
  uint32_t mark_begin = 0, mark_parts = 0, mark_qdelta = 0, mark_qend = 0;
  bool seg_enabled = false, refresh_golden = false, refresh_alt = false, refresh_last = false, refresh_entropy = false;
  uint8_t copy_golden = 0, copy_alt = 0, y_ac_qi = 0, lf_level = 0;
  int8_t q_delta[5] = {0, 0, 0, 0, 0};  // y_dc, y2_dc, y2_ac, uv_dc, uv_ac
  std::vector<uint32_t> sub_labels;  // per SPLITMV macroblock (split_idx): 2 bits per partition, coding order
};

// KeyFrame / InterFrame (decoder/frame.hh:126-127) in flat form.
struct ParsedFrame {
  explicit ParsedFrame(const Allocator& a = kMallocAllocator) : mbs(a), tokens(a), split(a) {}
  vp8gpu_frame_desc desc{};
  Buffer<vp8gpu_mb> mbs;
  Buffer<vp8gpu_token> tokens;
  Buffer<vp8gpu_split_mvs> split;
  TokenWork tw;
  bool keep_verbatim = false;  // ask parse_frame to fill `verbatim` (not compatible with defer_tokens)
  Verbatim verbatim;
};

// Parse one compressed frame and apply it to `state`.  Returns VP8GPU_OK or VP8GPU_ERR_*;
// on INVALID / UNSUPPORTED `state` is left untouched (the header is validated before anything is committed); an
// allocation failure (VP8GPU_ERR_NOMEM) during the macroblock pass leaves it advanced.  With defer_tokens only the first partition (frame header,
// macroblock modes, motion vectors) is decoded: records carry VP8GPU_MB_SKIP instead of
// tok_off / tok_cnt, no tokens are produced, desc.n_tokens is 0 and out.tw describes the rest.
int parse_frame(State& state, const uint8_t* data, size_t len, ParsedFrame& out, bool defer_tokens = false);

// true if the frame tag says key frame (uncompressed_chunk.cc:53)
inline bool is_key_frame(const uint8_t* data, size_t len) { return len > 0 && !(data[0] & 1); }

}  // namespace vp8
