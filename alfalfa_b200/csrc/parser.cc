// parser.cc -- see parser.h.  Host-only C++, no CUDA.
//
// Behavioural contract = the reference's parse path (all paths relative to
// /root/reference/src/decoder): uncompressed_chunk.cc:34-155, frame_header.hh:37-325,
// decoder_state.hh:73-167, macroblock.cc:44-456, tokens.cc:50-135, scorer.hh, frame.cc:96-137
// and :252-269.  The implementation is organised differently: one fused raster pass, a
// 64-bit-window arithmetic decoder, one row of neighbour context, flat output records.
#include "parser.h"

#include <stdlib.h>
#include <string.h>

#include "vp8_tables.h"

namespace vp8 {

// ------------------------------------------------------------------------------------------
// allocator plumbing
// ------------------------------------------------------------------------------------------
const Allocator kMallocAllocator = {&::malloc, &::free};

template <class T>
bool Buffer<T>::reserve(size_t n, size_t keep) {
  if (n <= cap_) return true;
  size_t want = cap_ ? cap_ * 2 : 1024;
  if (want < n) want = n;
  T* q = static_cast<T*>(a_.alloc(want * sizeof(T)));
  if (!q) return false;
  if (p_) {
    if (keep) memcpy(q, p_, keep * sizeof(T));
    a_.free(p_);
  }
  p_ = q;
  cap_ = want;
  return true;
}
template class Buffer<vp8gpu_mb>;
template class Buffer<vp8gpu_token>;
template class Buffer<vp8gpu_split_mvs>;

// ------------------------------------------------------------------------------------------
// DecoderState
// ------------------------------------------------------------------------------------------
State::State(int w, int h)
    : width(w), height(h), mb_cols((w + 15) / 16), mb_rows((h + 15) / 16),
      seg_map(static_cast<size_t>((w + 15) / 16) * ((h + 15) / 16), 3) {
  reset_probs();
}

void State::reset_probs() {
  memcpy(coef_probs, k_coef_default_probs, sizeof(coef_probs));
  memcpy(ymode_probs, k_ymode_default_probs, sizeof(ymode_probs));
  memcpy(uvmode_probs, k_uvmode_default_probs, sizeof(uvmode_probs));
  memcpy(mv_probs, k_mv_default_probs, sizeof(mv_probs));
}

bool State::operator==(const State& o) const {
  if (width != o.width || height != o.height) return false;
  if (memcmp(coef_probs, o.coef_probs, sizeof(coef_probs)) || memcmp(ymode_probs, o.ymode_probs, 4) ||
      memcmp(uvmode_probs, o.uvmode_probs, 3) || memcmp(mv_probs, o.mv_probs, sizeof(mv_probs)))
    return false;
  if (seg_enabled != o.seg_enabled || lf_adj_enabled != o.lf_adj_enabled) return false;
  if (seg_enabled && (seg_abs != o.seg_abs || memcmp(seg_quant, o.seg_quant, 4) ||
                      memcmp(seg_lf, o.seg_lf, 4) || seg_map != o.seg_map))
    return false;
  if (lf_adj_enabled && (memcmp(ref_adj, o.ref_adj, 4) || memcmp(mode_adj, o.mode_adj, 4))) return false;
  return true;
}

uint64_t State::hash() const {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&h](const void* p, size_t n) {
    const uint8_t* b = static_cast<const uint8_t*>(p);
    for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
  };
  mix(coef_probs, sizeof(coef_probs));
  mix(ymode_probs, 4);
  mix(uvmode_probs, 3);
  mix(mv_probs, sizeof(mv_probs));
  const uint8_t flags[2] = {seg_enabled, lf_adj_enabled};
  mix(flags, 2);
  if (seg_enabled) {
    const uint8_t a = seg_abs;
    mix(&a, 1);
    mix(seg_quant, 4);
    mix(seg_lf, 4);
    mix(seg_map.data(), seg_map.size());
  }
  if (lf_adj_enabled) {
    mix(ref_adj, 4);
    mix(mode_adj, 4);
  }
  return h;
}

// DecoderState::serialize (decoder.cc:283-314) in the reference's own tag-length-value format
// (enc_state_serializer.hh:43-116): every record is a one-byte EncoderSerDesTag, a little-endian u32 length and
// the payload; integers are little-endian.  A blob written here is readable by the reference's
// EncoderStateDeserializer and vice versa (tests/test_state_format.py compares the bytes).
namespace {
enum SerDesTag : uint8_t {  // EncoderSerDesTag, enc_state_serializer.hh:43-56
  TAG_PROB_TABLE = 0, TAG_FILT_ADJ, TAG_SEGM_ABS, TAG_SEGM_REL, TAG_DECODER_STATE, TAG_OPT_EMPTY, TAG_OPT_FULL,
  TAG_REFERENCES, TAG_REF_LAST, TAG_REF_GOLD, TAG_REF_ALT, TAG_DECODER
};
void put_u16(std::vector<uint8_t>& b, uint32_t v) {
  b.push_back(static_cast<uint8_t>(v & 0xFF));
  b.push_back(static_cast<uint8_t>((v >> 8) & 0xFF));
}
void put_u32(std::vector<uint8_t>& b, uint32_t v) {
  put_u16(b, v & 0xFFFF);
  put_u16(b, v >> 16);
}
uint32_t get_u16(const uint8_t* p) { return p[0] | (p[1] << 8); }
uint32_t get_u32(const uint8_t* p) { return get_u16(p) | (get_u16(p + 2) << 16); }
}  // namespace

std::vector<uint8_t> State::serialize() const {
  std::vector<uint8_t> b;
  b.push_back(TAG_DECODER_STATE);
  put_u32(b, 0);  // patched below
  put_u16(b, static_cast<uint32_t>(width));
  put_u16(b, static_cast<uint32_t>(height));
  // ProbabilityTables::serialize (probability_tables.cc:126-158)
  b.push_back(TAG_PROB_TABLE);
  put_u32(b, 1056 + 4 + 3 + 38);
  b.insert(b.end(), coef_probs, coef_probs + 1056);
  b.insert(b.end(), ymode_probs, ymode_probs + 4);
  b.insert(b.end(), uvmode_probs, uvmode_probs + 3);
  b.insert(b.end(), &mv_probs[0][0], &mv_probs[0][0] + 38);
  if (seg_enabled) {
    // Segmentation::serialize (decoder.cc:396-422).  The reference sizes its SegmentationMap with the frame's
    // PIXEL dimensions (decoder_state.hh:170-176: map( width, height, 3 )) and only ever touches the top-left
    // mb_cols x mb_rows corner; the rest keeps the initial value 3.
    b.push_back(TAG_OPT_FULL);
    b.push_back(seg_abs ? TAG_SEGM_ABS : TAG_SEGM_REL);
    put_u32(b, 4 + 4 + 4 + static_cast<uint32_t>(width) * static_cast<uint32_t>(height));
    put_u16(b, static_cast<uint32_t>(width));
    put_u16(b, static_cast<uint32_t>(height));
    for (int i = 0; i < 4; i++) b.push_back(static_cast<uint8_t>(seg_quant[i]));
    for (int i = 0; i < 4; i++) b.push_back(static_cast<uint8_t>(seg_lf[i]));
    const size_t at = b.size();
    b.resize(at + static_cast<size_t>(width) * height, 3);
    for (int r = 0; r < mb_rows && r < height; r++)
      memcpy(&b[at + static_cast<size_t>(r) * width], &seg_map[static_cast<size_t>(r) * mb_cols],
             static_cast<size_t>(mb_cols < width ? mb_cols : width));
  } else {
    b.push_back(TAG_OPT_EMPTY);
  }
  if (lf_adj_enabled) {
    // FilterAdjustments::serialize (decoder.cc:342-357)
    b.push_back(TAG_OPT_FULL);
    b.push_back(TAG_FILT_ADJ);
    put_u32(b, 8);
    for (int i = 0; i < 4; i++) b.push_back(static_cast<uint8_t>(ref_adj[i]));
    for (int i = 0; i < 4; i++) b.push_back(static_cast<uint8_t>(mode_adj[i]));
  } else {
    b.push_back(TAG_OPT_EMPTY);
  }
  const uint32_t len = static_cast<uint32_t>(b.size() - 5);  // = 4 + sub-records + the two option tags
  b[1] = static_cast<uint8_t>(len & 0xFF), b[2] = static_cast<uint8_t>((len >> 8) & 0xFF);
  b[3] = static_cast<uint8_t>((len >> 16) & 0xFF), b[4] = static_cast<uint8_t>(len >> 24);
  return b;
}

// DecoderState::deserialize (decoder.cc:316-330, :242-255); `used` (optional) = bytes consumed
bool State::deserialize(const uint8_t* d, size_t len, State& out, size_t* used) {
  if (len < 9 || d[0] != TAG_DECODER_STATE) return false;
  const size_t body = get_u32(d + 1);
  if (body + 5 > len) return false;
  const uint8_t* p = d + 5;
  const uint8_t* const end = d + 5 + body;
  const int w = static_cast<int>(get_u16(p)), h = static_cast<int>(get_u16(p + 2));
  p += 4;
  if (w <= 0 || h <= 0) return false;
  State s(w, h);
  const size_t prob_len = 1056 + 4 + 3 + 38;
  if (static_cast<size_t>(end - p) < 5 + prob_len + 2 || p[0] != TAG_PROB_TABLE || get_u32(p + 1) != prob_len) return false;
  p += 5;
  memcpy(s.coef_probs, p, 1056);
  memcpy(s.ymode_probs, p + 1056, 4);
  memcpy(s.uvmode_probs, p + 1060, 3);
  memcpy(s.mv_probs, p + 1063, 38);
  p += prob_len;
  const uint8_t seg_opt = *p++;
  if (seg_opt == TAG_OPT_FULL) {
    if (end - p < 5 + 12 || (p[0] != TAG_SEGM_ABS && p[0] != TAG_SEGM_REL)) return false;
    s.seg_enabled = true;
    s.seg_abs = p[0] == TAG_SEGM_ABS;
    const size_t seg_len = get_u32(p + 1);
    const size_t mw = get_u16(p + 5), mh = get_u16(p + 7);
    if (seg_len != 12 + mw * mh || static_cast<size_t>(end - p) < 5 + seg_len) return false;
    if (mw < static_cast<size_t>(s.mb_cols) || mh < static_cast<size_t>(s.mb_rows)) return false;
    memcpy(s.seg_quant, p + 9, 4);
    memcpy(s.seg_lf, p + 13, 4);
    const uint8_t* map = p + 17;
    for (int r = 0; r < s.mb_rows; r++) memcpy(&s.seg_map[static_cast<size_t>(r) * s.mb_cols], map + static_cast<size_t>(r) * mw, s.mb_cols);
    p += 5 + seg_len;
  } else if (seg_opt != TAG_OPT_EMPTY) {
    return false;
  }
  if (p >= end) return false;
  const uint8_t filt_opt = *p++;
  if (filt_opt == TAG_OPT_FULL) {
    if (end - p < 13 || p[0] != TAG_FILT_ADJ || get_u32(p + 1) != 8) return false;
    s.lf_adj_enabled = true;
    memcpy(s.ref_adj, p + 5, 4);
    memcpy(s.mode_adj, p + 9, 4);
    p += 13;
  } else if (filt_opt != TAG_OPT_EMPTY) {
    return false;
  }
  if (p != end) return false;
  if (used) *used = 5 + body;
  out = s;
  return true;
}

// ------------------------------------------------------------------------------------------
// arithmetic decoder.  Same code as bool_decoder.hh:82-107 produces, implemented with a
// 64-bit look-ahead window so the renormalisation is one shift and bytes are fetched eight
// at a time.  Reading past the end of the partition yields zero bits, as in the reference
// (load_octet does nothing once the chunk is empty).
// ------------------------------------------------------------------------------------------
class BoolReader {
 public:
  void init(const uint8_t* p, size_t n) {
    p_ = p;
    end_ = p + n;
    value_ = 0;
    count_ = -8;
    range_ = 255;
    fill();
  }
  __attribute__((always_inline)) inline int get(uint32_t prob) {
    const uint32_t split = 1 + (((range_ - 1) * prob) >> 8);
    if (count_ < 0) fill();
    const uint64_t bigsplit = static_cast<uint64_t>(split) << 56;
    int bit;
    uint32_t range;
    if (value_ >= bigsplit) {
      range = range_ - split;
      value_ -= bigsplit;
      bit = 1;
    } else {
      range = split;
      bit = 0;
    }
    const int shift = __builtin_clz(range) - 24;
    range_ = range << shift;
    value_ <<= shift;
    count_ -= shift;
    return bit;
  }
  inline int bit() { return get(128); }
  inline int literal(int width) {  // Unsigned<width>, MSB first
    int v = 0;
    for (int i = 0; i < width; i++) v = (v << 1) | get(128);
    return v;
  }
  inline int signed_literal(int width) {  // Signed<width>: magnitude, then sign
    const int v = literal(width);
    return get(128) ? -v : v;
  }
  inline int flagged_signed(int width) { return get(128) ? signed_literal(width) : 0; }
  // tree.cc:35-57
  inline int tree(const int8_t* nodes, const uint8_t* probs) {
    int i = 0;
    while ((i = nodes[i + get(probs[i >> 1])]) > 0) {
    }
    return -i;
  }

 private:
  void fill() {
    int shift = 48 - count_;  // where the next byte goes
    const ptrdiff_t left = end_ - p_;
    if (left >= 8) {
      uint64_t big;
      memcpy(&big, p_, 8);
      big = __builtin_bswap64(big);
      const int bits = (shift & ~7) + 8;  // whole bytes that fit
      const uint64_t nv = big >> (64 - bits);
      count_ += bits;
      p_ += bits >> 3;
      value_ |= nv << (shift & 7);
      return;
    }
    while (shift >= 0 && p_ < end_) {
      count_ += 8;
      value_ |= static_cast<uint64_t>(*p_++) << shift;
      shift -= 8;
    }
    if (p_ >= end_) count_ += 0x40000000;  // partition exhausted: only zero bits from here on
  }
  const uint8_t* p_ = nullptr;
  const uint8_t* end_ = nullptr;
  uint64_t value_ = 0;
  int count_ = 0;
  uint32_t range_ = 255;
};

// ------------------------------------------------------------------------------------------
// constant trees (modemv_data.cc:186-250)
// ------------------------------------------------------------------------------------------
namespace {

const int8_t kKfYModeTree[8] = {-VP8GPU_B_PRED, 2, 4, 6, -VP8GPU_DC_PRED, -VP8GPU_V_PRED, -VP8GPU_H_PRED,
                                -VP8GPU_TM_PRED};
const int8_t kYModeTree[8] = {-VP8GPU_DC_PRED, 2, 4, 6, -VP8GPU_V_PRED, -VP8GPU_H_PRED, -VP8GPU_TM_PRED,
                              -VP8GPU_B_PRED};
const int8_t kUvModeTree[6] = {-VP8GPU_DC_PRED, 2, -VP8GPU_V_PRED, 4, -VP8GPU_H_PRED, -VP8GPU_TM_PRED};
const int8_t kBModeTree[18] = {-VP8GPU_B_DC_PRED, 2,  -VP8GPU_B_TM_PRED, 4,  -VP8GPU_B_VE_PRED, 6,
                               8,                 12, -VP8GPU_B_HE_PRED, 10, -VP8GPU_B_RD_PRED, -VP8GPU_B_VR_PRED,
                               -VP8GPU_B_LD_PRED, 14, -VP8GPU_B_VL_PRED, 16, -VP8GPU_B_HD_PRED, -VP8GPU_B_HU_PRED};
const int8_t kSmallMvTree[14] = {2, 8, 4, 6, -0, -1, -2, -3, 10, 12, -4, -5, -6, -7};
const int8_t kMvRefTree[8] = {-VP8GPU_ZEROMV, 2, -VP8GPU_NEARESTMV, 4, -VP8GPU_NEARMV, 6, -VP8GPU_NEWMV,
                              -VP8GPU_SPLITMV};
enum { kSubLeft = 0, kSubAbove = 1, kSubZero = 2, kSubNew = 3 };
const int8_t kSubMvTree[6] = {-kSubLeft, 2, -kSubAbove, 4, -kSubZero, -kSubNew};
const int8_t kSplitTree[6] = {-3, 2, -2, 4, -0, -1};
const int8_t kSegmentTree[6] = {2, 4, -0, -1, -2, -3};

// split layouts (modemv_data.cc:252-278): bit i of kSplitFill[layout][part] is set when luma
// sub-block i belongs to partition `part`; the lowest set bit is the partition's first block.
const uint16_t kSplitFill[4][16] = {
    {0x00FF, 0xFF00},
    {0x3333, 0xCCCC},
    {0x0033, 0x00CC, 0x3300, 0xCC00},
    {0x0001, 0x0002, 0x0004, 0x0008, 0x0010, 0x0020, 0x0040, 0x0080, 0x0100, 0x0200, 0x0400, 0x0800, 0x1000,
     0x2000, 0x4000, 0x8000}};
const uint8_t kSplitCount[4] = {2, 2, 4, 16};

// tokens.hh:59-60; band offsets pre-multiplied by 3 contexts * 11 nodes
const uint16_t kBandOff[16] = {0 * 33, 1 * 33, 2 * 33, 3 * 33, 6 * 33, 4 * 33, 5 * 33, 6 * 33,
                               6 * 33, 6 * 33, 6 * 33, 6 * 33, 6 * 33, 6 * 33, 6 * 33, 7 * 33};
const uint32_t kZigzagShifted[16] = {0u << 16,  1u << 16,  4u << 16,  8u << 16, 5u << 16,  2u << 16,
                                     3u << 16,  6u << 16,  9u << 16,  12u << 16, 13u << 16, 10u << 16,
                                     7u << 16,  11u << 16, 14u << 16, 15u << 16};

// what later macroblocks need to know about an already parsed neighbour
struct Neighbour {
  uint8_t inter = 0, flipped = 0, y_mode = 0, pad = 0;
  int16_t mvx = 0, mvy = 0;   // base motion vector (sub-block 15)
  uint8_t bm[4] = {0, 0, 0, 0};  // sub-block modes along the shared edge
  int16_t emv[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};  // sub-block vectors along the shared edge
};

struct Mv {
  int x, y;
};

struct FrameHeader {
  bool key = false, show = false;
  bool seg_enabled = false, seg_update_map = false, seg_update_data = false, seg_abs = false;
  int8_t seg_quant[4] = {0, 0, 0, 0}, seg_lf[4] = {0, 0, 0, 0};
  uint8_t seg_tree_probs[3] = {255, 255, 255};
  bool filter_type = false;
  int lf_level = 0, sharpness = 0;
  bool lf_adj_enabled = false, lf_delta_update = false;
  int8_t ref_upd[4] = {0, 0, 0, 0}, mode_upd[4] = {0, 0, 0, 0};
  int log2_parts = 0;
  int y_ac_qi = 0, y_dc = 0, y2_dc = 0, y2_ac = 0, uv_dc = 0, uv_ac = 0;
  bool refresh_golden = true, refresh_alt = true, refresh_last = true, refresh_entropy = false;
  int copy_golden = 0, copy_alt = 0;
  bool sign_golden = false, sign_alt = false;
  bool has_skip_prob = false;
  int skip_prob = 0, prob_inter = 0, prob_last = 0, prob_golden = 0;
};

// A view of the first partition's reader that also writes every decision to a tape (Verbatim::header_tape);
// the header readers below are templates so that the plain path pays nothing for it.
struct TapedReader {
  BoolReader& r;
  std::vector<uint16_t>& tape;
  inline int get(uint32_t prob) {
    const int b = r.get(prob);
    tape.push_back(static_cast<uint16_t>((prob << 1) | static_cast<uint32_t>(b)));
    return b;
  }
  inline int bit() { return get(128); }
  inline int literal(int width) {
    int v = 0;
    for (int i = 0; i < width; i++) v = (v << 1) | get(128);
    return v;
  }
  inline int signed_literal(int width) {
    const int v = literal(width);
    return get(128) ? -v : v;
  }
  inline int flagged_signed(int width) { return get(128) ? signed_literal(width) : 0; }
  uint32_t* marks = nullptr;  // Verbatim::mark_* (positions in the tape), in the order the header reader passes them
  inline void mark(int id) { if (marks) marks[id] = static_cast<uint32_t>(tape.size()); }
};
inline void mark_header(BoolReader&, int) {}
inline void mark_header(TapedReader& r, int id) { r.mark(id); }

// the part of the frame header shared by key and inter frames (frame_header.hh:104-131, :70-84, :37-66)
template <class Reader>
void read_common_header(Reader& br, FrameHeader& h) {
  mark_header(br, 0);
  h.seg_enabled = br.bit();
  if (h.seg_enabled) {
    h.seg_update_map = br.bit();
    h.seg_update_data = br.bit();
    if (h.seg_update_data) {
      h.seg_abs = br.bit();
      for (int i = 0; i < 4; i++) h.seg_quant[i] = static_cast<int8_t>(br.flagged_signed(7));
      for (int i = 0; i < 4; i++) h.seg_lf[i] = static_cast<int8_t>(br.flagged_signed(6));
    }
    if (h.seg_update_map)
      for (int i = 0; i < 3; i++) h.seg_tree_probs[i] = br.bit() ? static_cast<uint8_t>(br.literal(8)) : 255;
  }
  h.filter_type = br.bit();
  h.lf_level = br.literal(6);
  h.sharpness = br.literal(3);
  h.lf_adj_enabled = br.bit();
  if (h.lf_adj_enabled) {
    h.lf_delta_update = br.bit();
    if (h.lf_delta_update) {
      for (int i = 0; i < 4; i++) h.ref_upd[i] = static_cast<int8_t>(br.flagged_signed(6));
      for (int i = 0; i < 4; i++) h.mode_upd[i] = static_cast<int8_t>(br.flagged_signed(6));
    }
  }
  mark_header(br, 1);
  h.log2_parts = br.literal(2);
  h.y_ac_qi = br.literal(7);
  mark_header(br, 2);
  h.y_dc = br.flagged_signed(4);
  h.y2_dc = br.flagged_signed(4);
  h.y2_ac = br.flagged_signed(4);
  h.uv_dc = br.flagged_signed(4);
  h.uv_ac = br.flagged_signed(4);
  mark_header(br, 3);
}

template <class Reader>
void read_coef_updates(Reader& br, uint8_t* probs) {
  for (int i = 0; i < 1056; i++)
    if (br.get(k_coef_update_probs[i])) probs[i] = static_cast<uint8_t>(br.literal(8));
}

inline int clamp_q(int q) { return q < 0 ? 0 : (q > 127 ? 127 : q); }

// Quantizer::Quantizer, quantization.cc:83-93
vp8gpu_quant resolve_quant(int y_ac_qi, const FrameHeader& h) {
  vp8gpu_quant q;
  q.y_ac = k_ac_q[clamp_q(y_ac_qi)];
  q.y_dc = k_dc_q[clamp_q(y_ac_qi + h.y_dc)];
  q.y2_ac = static_cast<uint16_t>(k_ac_q[clamp_q(y_ac_qi + h.y2_ac)] * 155 / 100);
  q.y2_dc = static_cast<uint16_t>(k_dc_q[clamp_q(y_ac_qi + h.y2_dc)] * 2);
  q.uv_ac = k_ac_q[clamp_q(y_ac_qi + h.uv_ac)];
  q.uv_dc = k_dc_q[clamp_q(y_ac_qi + h.uv_dc)];
  if (q.y2_ac < 8) q.y2_ac = 8;
  if (q.uv_dc > 132) q.uv_dc = 132;
  return q;
}

// MotionVector::read_component, macroblock.cc:198-229
inline int read_mv_component(BoolReader& br, const uint8_t* p) {
  int x = 0;
  if (br.get(p[0])) {  // long form
    for (int i = 0; i < 3; i++) x += br.get(p[9 + i]) << i;
    for (int i = 9; i > 3; i--) x += br.get(p[9 + i]) << i;
    if (!(x & 0xFFF0) || br.get(p[9 + 3])) x += 8;
  } else {
    x = br.tree(kSmallMvTree, p + 2);
  }
  x <<= 1;
  if (x && br.get(p[1])) x = -x;
  return x;
}
inline Mv read_mv(BoolReader& br, const uint8_t (*probs)[19]) {
  Mv m;
  m.y = read_mv_component(br, probs[0]);  // row first (macroblock.cc:284-286)
  m.x = read_mv_component(br, probs[1]);
  return m;
}

// vectors are kept as int16 in the reference (vp8_header_structures.hh:169); sums wrap likewise
inline int wrap16(int v) { return static_cast<int16_t>(v); }

struct MvBounds {
  int left, right, top, bottom;
};
inline Mv clamp_mv(Mv m, const MvBounds& b) {  // Scorer::clamp, macroblock.cc:183-195
  m.x = m.x < b.left ? b.left : (m.x > b.right ? b.right : m.x);
  m.y = m.y < b.top ? b.top : (m.y > b.bottom ? b.bottom : m.y);
  return m;
}

// Scorer (scorer.hh:35-78 + macroblock.cc:141-171)
struct Census {
  int score[4] = {0, 0, 0, 0};
  Mv mv[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  int index = 0, split_score = 0;
  bool flipped;
  explicit Census(bool f) : flipped(f) {}
  void add(int weight, const Neighbour& nb) {
    if (!nb.inter) return;
    int x = nb.mvx, y = nb.mvy;
    if ((nb.flipped != 0) != flipped) {
      x = -x;
      y = -y;
    }
    if ((x | y) == 0) {
      score[0] += weight;
    } else {
      if (x != mv[index].x || y != mv[index].y) {
        index++;
        mv[index].x = x;
        mv[index].y = y;
      }
      score[index] += weight;
    }
    if (nb.y_mode == VP8GPU_SPLITMV) split_score += weight;
  }
  void finish() {
    if (score[3] && mv[index].x == mv[1].x && mv[index].y == mv[1].y) score[1] += score[3];
    if (score[2] > score[1]) {
      const int s = score[1];
      score[1] = score[2];
      score[2] = s;
      const Mv m = mv[1];
      mv[1] = mv[2];
      mv[2] = m;
    }
    if (score[1] >= score[0]) mv[0] = mv[1];
  }
};

inline uint8_t implied_bmode(int y_mode) {  // macroblock.hh:151-160
  static const uint8_t t[4] = {VP8GPU_B_DC_PRED, VP8GPU_B_VE_PRED, VP8GPU_B_HE_PRED, VP8GPU_B_TM_PRED};
  return t[y_mode];
}

// One 4x4 block of tokens (tokens.cc:50-135).  `tp` = probabilities of the block type,
// `i` = first coefficient index, `tag` = block number << 20.  Returns has_nonzero.
inline int parse_block(BoolReader& br, const uint8_t* tp, int ctx, int i, uint32_t tag, vp8gpu_token*& out) {
  static const uint8_t cat2[2] = {165, 145}, cat3[3] = {173, 148, 140}, cat4[4] = {176, 155, 140, 135},
                       cat5[5] = {180, 157, 141, 134, 130},
                       cat6[11] = {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129};
  const uint8_t* p = tp + kBandOff[i] + ctx * 11;
  if (!br.get(p[0])) return 0;  // immediate end of block
  int nz = 0;
  for (;;) {
    while (!br.get(p[1])) {  // run of zero tokens: no end-of-block test after a zero
      if (++i == 16) return nz;
      p = tp + kBandOff[i];
    }
    int v;
    if (!br.get(p[2])) {
      v = 1;
      ctx = 1;
    } else {
      ctx = 2;
      if (!br.get(p[3])) {
        if (!br.get(p[4])) v = 2;
        else v = 3 + br.get(p[5]);
      } else {
        const uint8_t* extra;
        int n, base;
        if (!br.get(p[6])) {
          if (!br.get(p[7])) {
            extra = nullptr, n = 0, base = 5 + br.get(159);
          } else {
            extra = cat2, n = 2, base = 7;
          }
        } else if (!br.get(p[8])) {
          if (!br.get(p[9])) extra = cat3, n = 3, base = 11;
          else extra = cat4, n = 4, base = 19;
        } else {
          if (!br.get(p[10])) extra = cat5, n = 5, base = 35;
          else extra = cat6, n = 11, base = 67;
        }
        int inc = 0;
        for (int k = 0; k < n; k++) inc = (inc << 1) + br.get(extra[k]);
        v = base + inc;
      }
    }
    if (br.get(128)) v = -v;
    *out++ = tag | kZigzagShifted[i] | static_cast<uint16_t>(v);
    nz = 1;
    if (++i == 16) return 1;
    p = tp + kBandOff[i] + ctx * 11;
    if (!br.get(p[0])) return 1;  // end of block
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// parse_frame
// ------------------------------------------------------------------------------------------
// the whole frame header (frame_header.hh:213-325) and the frame's probability tables: a key frame starts
// from the defaults (decoder_state.hh:90, decoder.cc:236-243), an inter frame from the saved ones.  Returns
// color_space | clamping_type of a key frame.
template <class Reader>
bool read_frame_header(Reader& br, FrameHeader& h, const State& state, uint8_t* frame_coef, uint8_t* frame_ymode,
                       uint8_t* frame_uvmode, uint8_t (*frame_mv)[19]) {
  bool color_or_clamp = false;
  if (h.key) {
    color_or_clamp = br.bit() | br.bit();
    read_common_header(br, h);
    h.refresh_entropy = br.bit();
    // a key frame starts from default probabilities (decoder_state.hh:90, decoder.cc:236-243)
    memcpy(frame_coef, k_coef_default_probs, 1056);
    memcpy(frame_ymode, k_ymode_default_probs, 4);
    memcpy(frame_uvmode, k_uvmode_default_probs, 3);
    memcpy(frame_mv, k_mv_default_probs, 38);
    read_coef_updates(br, frame_coef);
    h.has_skip_prob = br.bit();
    if (h.has_skip_prob) h.skip_prob = br.literal(8);
  } else {
    read_common_header(br, h);
    h.refresh_golden = br.bit();
    h.refresh_alt = br.bit();
    h.copy_golden = h.refresh_golden ? 0 : br.literal(2);
    h.copy_alt = h.refresh_alt ? 0 : br.literal(2);
    h.sign_golden = br.bit();
    h.sign_alt = br.bit();
    h.refresh_entropy = br.bit();
    h.refresh_last = br.bit();
    memcpy(frame_coef, state.coef_probs, 1056);
    memcpy(frame_ymode, state.ymode_probs, 4);
    memcpy(frame_uvmode, state.uvmode_probs, 3);
    memcpy(frame_mv, state.mv_probs, 38);
    read_coef_updates(br, frame_coef);
    h.has_skip_prob = br.bit();
    if (h.has_skip_prob) h.skip_prob = br.literal(8);
    h.prob_inter = br.literal(8);
    h.prob_last = br.literal(8);
    h.prob_golden = br.literal(8);
    if (br.bit())
      for (int i = 0; i < 4; i++) frame_ymode[i] = static_cast<uint8_t>(br.literal(8));
    if (br.bit())
      for (int i = 0; i < 3; i++) frame_uvmode[i] = static_cast<uint8_t>(br.literal(8));
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 19; j++)
        if (br.get(k_mv_update_probs[i * 19 + j])) {
          const int x = br.literal(7);
          frame_mv[i][j] = static_cast<uint8_t>(x ? x << 1 : 1);
        }
  }
  return color_or_clamp;
}

int parse_frame(State& state, const uint8_t* data, size_t len, ParsedFrame& out, bool defer_tokens) {
  // ---- frame tag and partition layout (uncompressed_chunk.cc:34-130) ----
  if (len < 3) return VP8GPU_ERR_INVALID;
  const uint32_t tag = data[0] | (data[1] << 8) | (static_cast<uint32_t>(data[2]) << 16);
  FrameHeader h;
  h.key = !(tag & 1);
  h.show = (tag >> 4) & 1;
  if (((tag >> 1) & 7) != 0) return VP8GPU_ERR_UNSUPPORTED;  // only VP8 version 0 decodes
  const size_t first_len = (tag >> 5) & 0x7FFFF;
  const size_t first_off = h.key ? 10 : 3;
  if (len <= first_off + first_len) return VP8GPU_ERR_INVALID;
  if (h.key) {
    if (data[3] != 0x9d || data[4] != 0x01 || data[5] != 0x2a) return VP8GPU_ERR_INVALID;
    const int fw = (data[6] | (data[7] << 8)) & 0x3FFF, hscale = data[7] >> 6;
    const int fh = (data[8] | (data[9] << 8)) & 0x3FFF, vscale = data[9] >> 6;
    if (fw != state.width || fh != state.height || hscale || vscale) return VP8GPU_ERR_UNSUPPORTED;
  }
  const uint8_t* rest = data + first_off + first_len;
  const size_t rest_len = len - first_off - first_len;

  BoolReader br;
  br.init(data + first_off, first_len);

  // ---- frame header; the new persistent state is staged and committed only when the whole
  //      header (including the partition table) has been validated ----
  uint8_t frame_coef[1056], frame_ymode[4], frame_uvmode[3], frame_mv[2][19];
  bool color_or_clamp = false;
  if (out.keep_verbatim) {
    if (defer_tokens) return VP8GPU_ERR_LOGIC;
    out.verbatim.header_tape.clear();
    uint32_t marks[4] = {0, 0, 0, 0};
    TapedReader taped{br, out.verbatim.header_tape, marks};
    color_or_clamp = read_frame_header(taped, h, state, frame_coef, frame_ymode, frame_uvmode, frame_mv);
    Verbatim& v = out.verbatim;
    v.mark_begin = marks[0], v.mark_parts = marks[1], v.mark_qdelta = marks[2], v.mark_qend = marks[3];
  } else {
    color_or_clamp = read_frame_header(br, h, state, frame_coef, frame_ymode, frame_uvmode, frame_mv);
  }
  if (color_or_clamp || h.filter_type) return VP8GPU_ERR_UNSUPPORTED;  // frame_header.hh:221-227,292-294

  // ---- DCT partitions (uncompressed_chunk.cc:132-155) ----
  const int nparts = 1 << h.log2_parts;
  BoolReader parts[8];
  TokenWork& tw = out.tw;
  tw.deferred = defer_tokens;
  {
    const size_t table = static_cast<size_t>(3) * (nparts - 1);
    if (rest_len < table) return VP8GPU_ERR_INVALID;
    const uint8_t* p = rest + table;
    size_t left = rest_len - table;
    tw.nparts = static_cast<uint32_t>(nparts);
    tw.bits = p;
    tw.bits_len = static_cast<uint32_t>(left);
    for (int i = 0; i < nparts; i++) {
      size_t n = left;
      if (i < nparts - 1) {
        n = rest[3 * i] | (rest[3 * i + 1] << 8) | (static_cast<size_t>(rest[3 * i + 2]) << 16);
        if (n > left) return VP8GPU_ERR_INVALID;
      }
      if (defer_tokens) {
        tw.part_off[i] = static_cast<uint32_t>(p - tw.bits);
        tw.part_len[i] = static_cast<uint32_t>(n);
      } else {
        parts[i].init(p, n);
      }
      p += n;
      left -= n;
    }
  }
  if (defer_tokens) memcpy(tw.coef_probs, frame_coef, sizeof(frame_coef));

  const int cols = state.mb_cols, rows = state.mb_rows;
  const size_t n_mbs = static_cast<size_t>(cols) * rows;
  if (!out.mbs.reserve(n_mbs, 0)) return VP8GPU_ERR_NOMEM;

  // ---- commit the persistent state (decoder_state.hh:90-96 key, :124-152 inter) ----
  if (h.key) {
    state.reset_probs();
    state.seg_enabled = h.seg_enabled;
    state.seg_abs = false;
    memset(state.seg_quant, 0, 4);
    memset(state.seg_lf, 0, 4);
    if (h.seg_enabled) {
      memset(state.seg_map.data(), 3, n_mbs);  // a fresh Segmentation: map( width, height, 3 )
      if (h.seg_update_data) {
        state.seg_abs = h.seg_abs;
        memcpy(state.seg_quant, h.seg_quant, 4);
        memcpy(state.seg_lf, h.seg_lf, 4);
      }
    }
    state.lf_adj_enabled = h.lf_adj_enabled;
    memset(state.ref_adj, 0, 4);
    memset(state.mode_adj, 0, 4);
    if (h.lf_adj_enabled && h.lf_delta_update) {
      memcpy(state.ref_adj, h.ref_upd, 4);
      memcpy(state.mode_adj, h.mode_upd, 4);
    }
    if (h.refresh_entropy) memcpy(state.coef_probs, frame_coef, sizeof(frame_coef));
  } else {
    if (h.refresh_entropy) {
      memcpy(state.coef_probs, frame_coef, sizeof(frame_coef));
      memcpy(state.ymode_probs, frame_ymode, 4);
      memcpy(state.uvmode_probs, frame_uvmode, 3);
      memcpy(state.mv_probs, frame_mv, sizeof(frame_mv));
    }
    if (h.lf_adj_enabled) {
      if (!state.lf_adj_enabled) {
        memset(state.ref_adj, 0, 4);
        memset(state.mode_adj, 0, 4);
      }
      state.lf_adj_enabled = true;
      if (h.lf_delta_update) {  // FilterAdjustments::update: unflagged entries become 0
        memcpy(state.ref_adj, h.ref_upd, 4);
        memcpy(state.mode_adj, h.mode_upd, 4);
      }
    } else {
      state.lf_adj_enabled = false;
    }
    if (h.seg_enabled) {
      if (!state.seg_enabled) {
        state.seg_abs = false;
        memset(state.seg_quant, 0, 4);
        memset(state.seg_lf, 0, 4);
        memset(state.seg_map.data(), 3, n_mbs);
      }
      state.seg_enabled = true;
      if (h.seg_update_data) {
        state.seg_abs = h.seg_abs;
        memcpy(state.seg_quant, h.seg_quant, 4);
        memcpy(state.seg_lf, h.seg_lf, 4);
      }
    } else {
      state.seg_enabled = false;
    }
  }

  // ---- frame descriptor ----
  vp8gpu_frame_desc& d = out.desc;
  memset(&d, 0, sizeof(d));
  d.width = static_cast<uint16_t>(state.width);
  d.height = static_cast<uint16_t>(state.height);
  d.mb_cols = static_cast<uint16_t>(cols);
  d.mb_rows = static_cast<uint16_t>(rows);
  d.key_frame = h.key;
  d.show_frame = h.show;
  d.loop_filter_level = static_cast<uint8_t>(h.lf_level);
  d.sharpness = static_cast<uint8_t>(h.sharpness);
  d.refresh_last = h.refresh_last;
  d.refresh_golden = h.refresh_golden;
  d.refresh_alternate = h.refresh_alt;
  d.copy_to_golden = static_cast<uint8_t>(h.copy_golden);
  d.copy_to_alternate = static_cast<uint8_t>(h.copy_alt);
  for (int s = 0; s < 4; s++) {
    int qi = h.y_ac_qi;
    // the segment index passes through Unsigned<7> = uint8_t (frame.cc:196-199): it wraps
    if (state.seg_enabled) qi = static_cast<uint8_t>(state.seg_quant[s] + (state.seg_abs ? 0 : h.y_ac_qi));
    d.quant[s] = resolve_quant(qi, h);
  }
  // per-segment base loop-filter level, still unclamped (frame.cc:150-166)
  int seg_level[4];
  for (int s = 0; s < 4; s++)
    seg_level[s] = state.seg_enabled ? state.seg_lf[s] + (state.seg_abs ? 0 : h.lf_level) : h.lf_level;

  // ---- one fused raster pass over the macroblocks ----
  std::vector<Neighbour> above(cols);       // context from the row above, per column
  std::vector<uint16_t> above_nz(cols, 0);  // has_nonzero of the blocks above: Y0-3 | U<<4 | V<<6 | Y2<<8
  vp8gpu_mb* mbs = out.mbs.data();
  uint8_t* seg_map = state.seg_map.data();
  size_t n_tok = 0;
  uint32_t n_split = 0;
  const uint8_t* const coef_y_after_y2 = frame_coef + 0 * 264;
  const uint8_t* const coef_y2 = frame_coef + 1 * 264;
  const uint8_t* const coef_uv = frame_coef + 2 * 264;
  const uint8_t* const coef_y_full = frame_coef + 3 * 264;
  const bool read_segment = h.seg_enabled && h.seg_update_map;
  const uint8_t(*mv_probs)[19] = frame_mv;
  Verbatim* const vb = out.keep_verbatim ? &out.verbatim : nullptr;
  if (vb) {
    vb->key = h.key, vb->show = h.show;
    vb->width = state.width, vb->height = state.height, vb->log2_parts = h.log2_parts;
    vb->has_skip_prob = h.has_skip_prob, vb->read_segment = read_segment;
    vb->sign_golden = h.sign_golden, vb->sign_alt = h.sign_alt;
    vb->skip_prob = static_cast<uint8_t>(h.skip_prob), vb->prob_inter = static_cast<uint8_t>(h.prob_inter);
    vb->prob_last = static_cast<uint8_t>(h.prob_last), vb->prob_golden = static_cast<uint8_t>(h.prob_golden);
    memcpy(vb->seg_tree_probs, h.seg_tree_probs, 3);
    vb->seg_enabled = h.seg_enabled;
    vb->refresh_golden = h.refresh_golden, vb->refresh_alt = h.refresh_alt, vb->refresh_last = h.refresh_last;
    vb->refresh_entropy = h.refresh_entropy;
    vb->copy_golden = static_cast<uint8_t>(h.copy_golden), vb->copy_alt = static_cast<uint8_t>(h.copy_alt);
    vb->y_ac_qi = static_cast<uint8_t>(h.y_ac_qi), vb->lf_level = static_cast<uint8_t>(h.lf_level);
    vb->q_delta[0] = static_cast<int8_t>(h.y_dc), vb->q_delta[1] = static_cast<int8_t>(h.y2_dc), vb->q_delta[2] = static_cast<int8_t>(h.y2_ac);
    vb->q_delta[3] = static_cast<int8_t>(h.uv_dc), vb->q_delta[4] = static_cast<int8_t>(h.uv_ac);
    memcpy(vb->coef, frame_coef, 1056);
    memcpy(vb->ymode, frame_ymode, 4);
    memcpy(vb->uvmode, frame_uvmode, 3);
    memcpy(vb->mv, frame_mv, 38);
    vb->mb_coded.assign(n_mbs, 0);
    vb->sub_labels.clear();
  }

  for (int row = 0; row < rows; row++) {
    Neighbour left, above_left;  // outside the frame: not inter, B_DC_PRED, zero vectors
    unsigned left_nz = 0;
    BoolReader& tr = parts[row & (nparts - 1)];  // row r -> partition r % n (frame.cc:131-136)
    MvBounds bounds;
    bounds.top = -((row * 16) << 3) - 128;
    bounds.bottom = (((rows - 1 - row) * 16) << 3) + 128;
    if (bounds.top < -32768) bounds.top = -32768;
    if (bounds.bottom > 32767) bounds.bottom = 32767;

    for (int col = 0; col < cols; col++) {
      const size_t idx = static_cast<size_t>(row) * cols + col;
      const Neighbour up = above[col];  // copy: above[col] is replaced at the end of this iteration

      // -- macroblock header (macroblock.cc:44-71) --
      if (read_segment) seg_map[idx] = static_cast<uint8_t>(br.tree(kSegmentTree, h.seg_tree_probs));
      const int segment = state.seg_enabled ? seg_map[idx] : 0;
      const int skip = h.has_skip_prob ? br.get(h.skip_prob) : 0;

      int ref = VP8GPU_REF_CURRENT;
      if (!h.key && br.get(h.prob_inter)) {
        ref = VP8GPU_REF_LAST;
        if (br.get(h.prob_last)) ref = br.get(h.prob_golden) ? VP8GPU_REF_ALTREF : VP8GPU_REF_GOLDEN;
      }

      int y_mode, uv_mode = 0;
      uint64_t b_modes = 0;
      uint32_t split_idx = 0;
      uint8_t bm[16];       // sub-block intra modes (key-frame contexts)
      int16_t mv[16][2];    // sub-block motion vectors
      bool flipped = false;

      if (ref == VP8GPU_REF_CURRENT) {
        // -- intra modes: key frames macroblock.cc:84-111, inter frames :354-374 --
        memset(mv, 0, sizeof(mv));
        if (h.key) {
          y_mode = br.tree(kKfYModeTree, k_kf_ymode_probs);
          if (y_mode == VP8GPU_B_PRED) {
            for (int i = 0; i < 16; i++) {
              const int a = i >= 4 ? bm[i - 4] : up.bm[i];
              const int l = (i & 3) ? bm[i - 1] : left.bm[i >> 2];
              bm[i] = static_cast<uint8_t>(br.tree(kBModeTree, k_kf_bmode_probs + (a * 10 + l) * 9));
              b_modes |= static_cast<uint64_t>(bm[i]) << (4 * i);
            }
          } else {
            memset(bm, implied_bmode(y_mode), 16);
          }
          uv_mode = br.tree(kUvModeTree, k_kf_uvmode_probs);
        } else {
          y_mode = br.tree(kYModeTree, frame_ymode);
          if (y_mode == VP8GPU_B_PRED)
            for (int i = 0; i < 16; i++)
              b_modes |= static_cast<uint64_t>(br.tree(kBModeTree, k_bmode_probs)) << (4 * i);
          uv_mode = br.tree(kUvModeTree, frame_uvmode);
          memset(bm, 0, 16);  // not used as context in inter frames
        }
      } else {
        // -- inter modes and vectors (macroblock.cc:376-455) --
        memset(bm, 0, 16);
        flipped = (ref == VP8GPU_REF_GOLDEN && h.sign_golden) || (ref == VP8GPU_REF_ALTREF && h.sign_alt);
        Census census(flipped);
        census.add(2, up);
        census.add(2, left);
        census.add(1, above_left);
        census.finish();
        const uint8_t ref_probs[4] = {k_mv_count_probs[census.score[0] * 4 + 0],
                                      k_mv_count_probs[census.score[1] * 4 + 1],
                                      k_mv_count_probs[census.score[2] * 4 + 2],
                                      k_mv_count_probs[census.split_score * 4 + 3]};
        y_mode = br.tree(kMvRefTree, ref_probs);
        bounds.left = -((col * 16) << 3) - 128;
        bounds.right = (((cols - 1 - col) * 16) << 3) + 128;
        if (bounds.left < -32768) bounds.left = -32768;
        if (bounds.right > 32767) bounds.right = 32767;
        Mv base = {0, 0};
        if (y_mode == VP8GPU_NEARESTMV) {
          base = clamp_mv(census.mv[1], bounds);
        } else if (y_mode == VP8GPU_NEARMV) {
          base = clamp_mv(census.mv[2], bounds);
        } else if (y_mode == VP8GPU_NEWMV) {
          const Mv delta = read_mv(br, mv_probs);
          const Mv best = clamp_mv(census.mv[0], bounds);
          base.x = wrap16(delta.x + best.x);
          base.y = wrap16(delta.y + best.y);
        } else if (y_mode == VP8GPU_SPLITMV) {
          const int layout = br.tree(kSplitTree, k_split_probs);
          const Mv best = clamp_mv(census.mv[0], bounds);
          uint32_t labels = 0;
          for (int part = 0; part < kSplitCount[layout]; part++) {
            const unsigned members = kSplitFill[layout][part];
            const int first = __builtin_ctz(members);
            const int bx = first & 3, by = first >> 2;
            // read_subblock_inter_prediction (macroblock.cc:231-280)
            const int lx = bx ? mv[first - 1][0] : left.emv[by][0], ly = bx ? mv[first - 1][1] : left.emv[by][1];
            const int ax = by ? mv[first - 4][0] : up.emv[bx][0], ay = by ? mv[first - 4][1] : up.emv[bx][1];
            const bool left_zero = (lx | ly) == 0, above_zero = (ax | ay) == 0, same = lx == ax && ly == ay;
            int ctx = 0;
            if (same && left_zero) ctx = 4;
            else if (same) ctx = 3;
            else if (above_zero) ctx = 2;
            else if (left_zero) ctx = 1;
            int sx = 0, sy = 0;
            const int label = br.tree(kSubMvTree, k_submv_ref_probs + ctx * 3);
            labels |= static_cast<uint32_t>(label) << (2 * part);
            switch (label) {
              case kSubLeft: sx = lx, sy = ly; break;
              case kSubAbove: sx = ax, sy = ay; break;
              case kSubZero: break;
              default: {
                const Mv delta = read_mv(br, mv_probs);
                sx = wrap16(delta.x + best.x);
                sy = wrap16(delta.y + best.y);
              }
            }
            for (unsigned m = members; m; m &= m - 1) {
              const int i = __builtin_ctz(m);
              mv[i][0] = static_cast<int16_t>(sx);
              mv[i][1] = static_cast<int16_t>(sy);
            }
          }
          base.x = mv[15][0];
          base.y = mv[15][1];
          if (!out.split.reserve(n_split + 1, n_split)) return VP8GPU_ERR_NOMEM;
          memcpy(out.split.data()[n_split].mv, mv, sizeof(mv));
          split_idx = n_split++;
          if (vb) {
            vb->sub_labels.push_back(labels);
            vb->mb_coded[idx] |= static_cast<uint8_t>(layout << 1);
          }
        }
        if (y_mode != VP8GPU_SPLITMV)
          for (int i = 0; i < 16; i++) {
            mv[i][0] = static_cast<int16_t>(base.x);
            mv[i][1] = static_cast<int16_t>(base.y);
          }
      }
      const bool has_y2 = y_mode != VP8GPU_B_PRED && y_mode != VP8GPU_SPLITMV;
      if (vb) vb->mb_coded[idx] |= static_cast<uint8_t>(skip);

      // -- loop-filter level of this macroblock: frame.cc:150-166, loopfilter.cc:57-79,
      //    macroblock.cc:621, and the single clamp of loopfilter.cc:85 --
      int level = 0;
      if (h.lf_level) {
        level = seg_level[segment];
        if (state.lf_adj_enabled) {
          int mode_adj;
          if (ref == VP8GPU_REF_CURRENT) mode_adj = y_mode == VP8GPU_B_PRED ? state.mode_adj[0] : 0;
          else if (y_mode == VP8GPU_ZEROMV) mode_adj = state.mode_adj[1];
          else if (y_mode == VP8GPU_SPLITMV) mode_adj = state.mode_adj[3];
          else mode_adj = state.mode_adj[2];
          level += state.ref_adj[ref] + mode_adj;
        }
        level = level <= 0 ? 0 : (level > 63 ? 63 : level);
      }

      // -- coefficient tokens (macroblock.cc:468-502, tokens.cc:50-135) --
      unsigned a_nz = above_nz[col];
      const size_t tok_off = n_tok;
      unsigned tok_cnt = 0;
      if (defer_tokens) {
        // left to csrc/tokens.cu, which needs to know mb_skip_coeff
      } else if (skip) {
        // every block of a skipped macroblock has has_nonzero_ == false; a coded Y2 becomes the
        // new (zero) Y2 context, a macroblock without Y2 leaves the previous one (frame.cc:252-269)
        const unsigned keep = has_y2 ? 0u : 0x100u;
        a_nz &= keep;
        left_nz &= keep;
      } else {
        if (!out.tokens.reserve(n_tok + 400, n_tok)) return VP8GPU_ERR_NOMEM;
        vp8gpu_token* const t0 = out.tokens.data() + n_tok;
        vp8gpu_token* t = t0;
        const uint8_t* y_probs = coef_y_full;
        int first = 0;
        if (has_y2) {
          const int ctx = ((a_nz >> 8) & 1) + ((left_nz >> 8) & 1);
          const unsigned nz = parse_block(tr, coef_y2, ctx, 0, VP8GPU_BLK_Y2 << 20, t);
          a_nz = (a_nz & ~0x100u) | (nz << 8);
          left_nz = (left_nz & ~0x100u) | (nz << 8);
          y_probs = coef_y_after_y2;
          first = 1;
        }
        for (int i = 0; i < 16; i++) {
          const int bx = i & 3, by = i >> 2;
          const int ctx = ((a_nz >> bx) & 1) + ((left_nz >> by) & 1);
          const unsigned nz = parse_block(tr, y_probs, ctx, first, static_cast<uint32_t>(i) << 20, t);
          a_nz = (a_nz & ~(1u << bx)) | (nz << bx);
          left_nz = (left_nz & ~(1u << by)) | (nz << by);
        }
        for (int plane = 0; plane < 2; plane++) {  // U then V
          const int sh = 4 + 2 * plane;
          for (int i = 0; i < 4; i++) {
            const int bx = sh + (i & 1), by = sh + (i >> 1);
            const int ctx = ((a_nz >> bx) & 1) + ((left_nz >> by) & 1);
            const unsigned nz =
                parse_block(tr, coef_uv, ctx, 0, static_cast<uint32_t>(VP8GPU_BLK_U + 4 * plane + i) << 20, t);
            a_nz = (a_nz & ~(1u << bx)) | (nz << bx);
            left_nz = (left_nz & ~(1u << by)) | (nz << by);
          }
        }
        tok_cnt = static_cast<unsigned>(t - t0);
        n_tok += tok_cnt;
      }
      above_nz[col] = static_cast<uint16_t>(a_nz);

      // -- emit the record --
      vp8gpu_mb& mb = mbs[idx];
      mb.tok_off = static_cast<uint32_t>(tok_off);
      mb.tok_cnt = static_cast<uint16_t>(tok_cnt);
      mb.y_mode = static_cast<uint8_t>(y_mode);
      mb.uv_mode = static_cast<uint8_t>(uv_mode);
      mb.ref_frame = static_cast<uint8_t>(ref);
      mb.segment_id = static_cast<uint8_t>(segment);
      mb.lf_level = static_cast<uint8_t>(level);
      mb.flags = static_cast<uint8_t>((has_y2 ? VP8GPU_MB_HAS_Y2 : 0) | (defer_tokens && skip ? VP8GPU_MB_SKIP : 0));
      mb.mv_x = mv[15][0];
      mb.mv_y = mv[15][1];
      mb.split_idx = split_idx;
      mb.reserved = 0;
      mb.b_modes = b_modes;

      // -- neighbour context for the macroblocks to the right and below --
      Neighbour common;
      common.inter = ref != VP8GPU_REF_CURRENT;
      common.flipped = flipped;
      common.y_mode = static_cast<uint8_t>(y_mode);
      common.mvx = mv[15][0];
      common.mvy = mv[15][1];
      Neighbour below_view = common, right_view = common;
      for (int k = 0; k < 4; k++) {
        below_view.bm[k] = bm[12 + k];          // bottom row
        below_view.emv[k][0] = mv[12 + k][0];
        below_view.emv[k][1] = mv[12 + k][1];
        right_view.bm[k] = bm[4 * k + 3];       // right column
        right_view.emv[k][0] = mv[4 * k + 3][0];
        right_view.emv[k][1] = mv[4 * k + 3][1];
      }
      above_left = up;
      above[col] = below_view;
      left = right_view;
    }
  }
  d.n_tokens = static_cast<uint32_t>(n_tok);
  d.n_split = n_split;
  return VP8GPU_OK;
}

}  // namespace vp8
