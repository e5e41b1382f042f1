// tokens_core.cuh -- the device-side token decoder's logic (see tokens.cu), written so that it also
// compiles as plain host C++: tests/test_tokens_host.py runs exactly this code on the CPU against
// the CPU front end (csrc/parser.cc) on every golden vector before it ever sees a GPU.
//
// What the code is shaped by: one thread walks one frame, so everything is latency.  The decoder
// state lives in registers (nothing here takes its address), and the eleven probabilities of a
// (block type, band, context) position are fetched with ONE 16-byte shared-memory load per token
// instead of one byte load per decision.
#pragma once
#include <stdint.h>

#include "engine.h"

#ifdef __CUDACC__
#define TK_DEV __device__ __forceinline__
#define TK_CONST __constant__
#define TK_LDG(p) __ldg(p)
#define TK_LDCG(p) __ldcg(p)
#define TK_CLZ(x) __clz(x)
#else
#define TK_DEV inline __attribute__((always_inline))
#define TK_CONST static const
#define TK_LDG(p) (*(p))
#define TK_LDCG(p) (*(p))
#define TK_CLZ(x) __builtin_clz(x)
#endif

namespace vp8 {
namespace tok {

// coefficient bands (tokens.hh:59-60)
TK_CONST uint8_t c_band[16] = {0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7};
TK_CONST uint8_t c_zigzag[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
// extra-bit probabilities of DCT_CAT2..6 (tokens.hh:74-78), rows padded to 11
TK_CONST uint8_t c_cat[5][11] = {{165, 145},
                                 {173, 148, 140},
                                 {176, 155, 140, 135},
                                 {180, 157, 141, 134, 130},
                                 {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129}};

// Probability table as the decoder wants it: [type 4][band 8][ctx 3] entries of 16 bytes (11 used).
constexpr int kProbEntries = 4 * 8 * 3;
constexpr int kProbBytes = kProbEntries * 16;
struct alignas(16) Probs16 {
  uint32_t w[4];
  TK_DEV uint32_t at(int k) const { return (w[k >> 2] >> ((k & 3) * 8)) & 0xFF; }
};
// entry e of the 16-byte layout from the 1056-byte table of the frame header (type, band, ctx, node)
TK_DEV void expand_prob_entry(const uint8_t* src, uint8_t* dst, int e) {
  for (int k = 0; k < 11; k++) dst[e * 16 + k] = src[e * 11 + k];
  for (int k = 11; k < 16; k++) dst[e * 16 + k] = 0;
}

// BoolDecoder with a 64-bit look-ahead window; bit-for-bit the decisions of csrc/parser.cc's
// BoolReader (and therefore bool_decoder.hh:82-107): past the end of the partition only zero
// bits arrive.  Plain struct, always handled by value / reference to a local: stays in registers.
struct BoolReader {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t value;
  int count;
  uint32_t range;
};
TK_DEV void br_fill(BoolReader& b) {
  int shift = 48 - b.count;  // where the next byte goes
  while (shift >= 0 && b.p < b.end) {
    b.count += 8;
    b.value |= static_cast<uint64_t>(TK_LDG(b.p++)) << shift;
    shift -= 8;
  }
  if (b.p >= b.end) b.count += 0x40000000;  // exhausted: zero bits from here on
}
TK_DEV void br_init(BoolReader& b, const uint8_t* data, uint32_t n) {
  b.p = data;
  b.end = data + n;
  b.value = 0;
  b.count = -8;
  b.range = 255;
  br_fill(b);
}
TK_DEV int br_get(BoolReader& b, uint32_t prob) {
  const uint32_t split = 1 + (((b.range - 1) * prob) >> 8);
  if (b.count < 0) br_fill(b);
  // the window keeps the 8 compared bits on top: only the high word takes part
  uint32_t hi = static_cast<uint32_t>(b.value >> 32);
  const uint32_t big = split << 24;
  const int bit = hi >= big;
  const uint32_t r = bit ? b.range - split : split;
  hi -= bit ? big : 0u;
  const int shift = TK_CLZ(r) - 24;
  b.range = r << shift;
  b.value = ((static_cast<uint64_t>(hi) << 32) | static_cast<uint32_t>(b.value)) << shift;
  b.count -= shift;
  return bit;
}

// One 4x4 block (tokens.cc:50-135).  tp = 16-byte probability entries of the block type,
// i = first coefficient, tag = block number << 20.  Returns has_nonzero.
TK_DEV int parse_block(BoolReader& br, const uint8_t* tp, int ctx, int i, uint32_t tag, vp8gpu_token*& out) {
  Probs16 P = *reinterpret_cast<const Probs16*>(tp + (c_band[i] * 3 + ctx) * 16);
  if (!br_get(br, P.at(0))) return 0;
  int nz = 0;
  for (;;) {
    while (!br_get(br, P.at(1))) {  // zero tokens: no end-of-block test after a zero
      if (++i == 16) return nz;
      P = *reinterpret_cast<const Probs16*>(tp + c_band[i] * 48);
    }
    int v;
    if (!br_get(br, P.at(2))) {
      v = 1;
      ctx = 1;
    } else {
      ctx = 2;
      if (!br_get(br, P.at(3))) {
        if (!br_get(br, P.at(4))) v = 2;
        else v = 3 + br_get(br, P.at(5));
      } else {
        int cat, n, base;
        if (!br_get(br, P.at(6))) {
          if (!br_get(br, P.at(7))) {
            cat = 0, n = 0, base = 5 + br_get(br, 159);
          } else {
            cat = 0, n = 2, base = 7;
          }
        } else if (!br_get(br, P.at(8))) {
          if (!br_get(br, P.at(9))) cat = 1, n = 3, base = 11;
          else cat = 2, n = 4, base = 19;
        } else {
          if (!br_get(br, P.at(10))) cat = 3, n = 5, base = 35;
          else cat = 4, n = 11, base = 67;
        }
        int inc = 0;
        for (int k = 0; k < n; k++) inc = (inc << 1) + br_get(br, c_cat[cat][k]);
        v = base + inc;
      }
    }
    if (br_get(br, 128)) v = -v;
    *out++ = tag | (static_cast<uint32_t>(c_zigzag[i]) << 16) | static_cast<uint16_t>(v);
    nz = 1;
    if (++i == 16) return 1;
    P = *reinterpret_cast<const Probs16*>(tp + (c_band[i] * 3 + ctx) * 16);
    if (!br_get(br, P.at(0))) return 1;
  }
}

// One frame, raster order (Frame::parse_tokens, frame.cc:122-137 + Macroblock::parse_tokens,
// macroblock.cc:468-502).  probs: the frame's coefficient probabilities in the 16-byte layout
// (kProbBytes, 16-byte aligned); above_nz: mb_cols zeroed words (Y0-3 | U << 4 | V << 6 | Y2 << 8
// per column).  On the device both live in shared memory.
TK_DEV void decode_frame_tokens(const TokJob& J, const Geom& g, const uint8_t* probs, uint16_t* above_nz) {
  BoolReader parts[8];
  const int nparts = static_cast<int>(J.nparts);
  for (int i = 0; i < nparts; i++) br_init(parts[i], J.bits + J.part_off[i], J.part_len[i]);

  vp8gpu_mb* const mbs = J.mbs;
  vp8gpu_token* const t_begin = J.tokens;
  vp8gpu_token* t = t_begin;
  const vp8gpu_token* const t_limit = t_begin + J.tok_cap;
  uint32_t overflow = 0;
  const uint8_t* const coef_y_after_y2 = probs + 0 * 384;
  const uint8_t* const coef_y2 = probs + 1 * 384;
  const uint8_t* const coef_uv = probs + 2 * 384;
  const uint8_t* const coef_y_full = probs + 3 * 384;

  // word 1 of a record = tok_cnt | y_mode << 16 | uv_mode << 24, word 2 = ref | segment | lf | flags << 24
  const uint32_t* rec = reinterpret_cast<const uint32_t*>(mbs);
  uint32_t w1 = TK_LDCG(rec + 1), w2 = TK_LDCG(rec + 2);
  const int n_mbs = g.mb_cols * g.mb_rows;
  int idx = 0;
  for (int row = 0; row < g.mb_rows; row++) {
    unsigned left_nz = 0;
    BoolReader tr = parts[row & (nparts - 1)];  // row r -> partition r % n (frame.cc:131-136)
    for (int col = 0; col < g.mb_cols; col++, idx++) {
      const uint32_t cur1 = w1, cur2 = w2;
      if (idx + 1 < n_mbs) {  // next record: requested now, needed after this macroblock's tokens
        w1 = TK_LDCG(rec + 8 * (idx + 1) + 1);
        w2 = TK_LDCG(rec + 8 * (idx + 1) + 2);
      }
      const int y_mode = (cur1 >> 16) & 0xFF;
      const bool skip = (cur2 >> 24) & VP8GPU_MB_SKIP;
      const bool has_y2 = y_mode != VP8GPU_B_PRED && y_mode != VP8GPU_SPLITMV;
      unsigned a_nz = above_nz[col];
      vp8gpu_token* const t0 = t;
      if (skip) {
        // frame.cc:252-269: a macroblock without Y2 leaves the previous Y2 context in place
        const unsigned keep = has_y2 ? 0u : 0x100u;
        a_nz &= keep;
        left_nz &= keep;
      } else if (t + 400 > t_limit) {
        overflow = 1;  // cannot happen with the capacity rule of the host (engine.cu); stay in bounds
        a_nz = 0;
        left_nz = 0;
      } else {
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
        // 16 luma blocks, then U and V: block b uses context bits (bx, by) of the above / left words
#pragma unroll 1
        for (int b = 0; b < 24; b++) {
          int bx, by;
          const uint8_t* tp;
          int f;
          if (b < 16) {
            bx = b & 3, by = b >> 2, tp = y_probs, f = first;
          } else {
            const int sh = 4 + 2 * ((b - 16) >> 2);
            bx = sh + (b & 1), by = sh + ((b >> 1) & 1), tp = coef_uv, f = 0;
          }
          const int ctx = ((a_nz >> bx) & 1) + ((left_nz >> by) & 1);
          const unsigned nz = parse_block(tr, tp, ctx, f, static_cast<uint32_t>(b) << 20, t);
          a_nz = (a_nz & ~(1u << bx)) | (nz << bx);
          left_nz = (left_nz & ~(1u << by)) | (nz << by);
        }
      }
      above_nz[col] = static_cast<uint16_t>(a_nz);
      uint32_t* wr = reinterpret_cast<uint32_t*>(mbs) + 8 * idx;
      wr[0] = static_cast<uint32_t>(t0 - t_begin);
      wr[1] = (cur1 & 0xFFFF0000u) | static_cast<uint32_t>(t - t0);
      wr[2] = cur2 & ~(static_cast<uint32_t>(VP8GPU_MB_SKIP) << 24);
    }
    parts[row & (nparts - 1)] = tr;
  }
  J.result[0] = static_cast<uint32_t>(t - t_begin);
  J.result[1] = overflow;
}

}  // namespace tok
}  // namespace vp8
