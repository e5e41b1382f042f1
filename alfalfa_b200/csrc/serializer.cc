// serializer.cc -- see serializer.h.  Host-only C++.
//
// The syntax written here is the exact inverse of what parser.cc reads (and therefore of the
// reference's decoder/*.cc); the reference's own writer is encoder/serializer.cc:388-829,
// encode_tree.cc and bool_encoder.hh.  Every macroblock element is written with the context the
// decoder will have when it reads it.
#include "serializer.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <thread>

#include "hostpool.h"

#include "vp8_tables.h"

namespace vp8 {

// ------------------------------------------------------------------------------------------
// BoolWriter: RFC 6386 section 7.3
// ------------------------------------------------------------------------------------------
void BoolWriter::add_one() {
  size_t i = out_.size();
  while (i > 0 && out_[i - 1] == 255) out_[--i] = 0;
  if (i > 0) ++out_[i - 1];
}
void BoolWriter::literal(int value, int width) {
  for (int i = width - 1; i >= 0; i--) put((value >> i) & 1, 128);
}
std::vector<uint8_t> BoolWriter::finish() {
  // pad like libvpx's vp8_stop_encode (and the reference, bool_encoder.hh:78-82): 32 zero bits
  for (int i = 0; i < 32; i++) put(0, 128);
  return std::move(out_);
}

namespace {

// trees: same arrays the decoder walks (modemv_data.cc:186-250)
const int8_t kKfYModeTree[8] = {-VP8GPU_B_PRED, 2, 4, 6, -VP8GPU_DC_PRED, -VP8GPU_V_PRED, -VP8GPU_H_PRED, -VP8GPU_TM_PRED};
const int8_t kYModeTree[8] = {-VP8GPU_DC_PRED, 2, 4, 6, -VP8GPU_V_PRED, -VP8GPU_H_PRED, -VP8GPU_TM_PRED, -VP8GPU_B_PRED};
const int8_t kUvModeTree[6] = {-VP8GPU_DC_PRED, 2, -VP8GPU_V_PRED, 4, -VP8GPU_H_PRED, -VP8GPU_TM_PRED};
const int8_t kBModeTree[18] = {-VP8GPU_B_DC_PRED, 2,  -VP8GPU_B_TM_PRED, 4,  -VP8GPU_B_VE_PRED, 6,
                               8,                 12, -VP8GPU_B_HE_PRED, 10, -VP8GPU_B_RD_PRED, -VP8GPU_B_VR_PRED,
                               -VP8GPU_B_LD_PRED, 14, -VP8GPU_B_VL_PRED, 16, -VP8GPU_B_HD_PRED, -VP8GPU_B_HU_PRED};
const int8_t kSmallMvTree[14] = {2, 8, 4, 6, -0, -1, -2, -3, 10, 12, -4, -5, -6, -7};
const int8_t kMvRefTree[8] = {-VP8GPU_ZEROMV, 2, -VP8GPU_NEARESTMV, 4, -VP8GPU_NEARMV, 6, -VP8GPU_NEWMV, -VP8GPU_SPLITMV};
enum { kSubLeft = 0, kSubAbove = 1, kSubZero = 2, kSubNew = 3 };
const int8_t kSubMvTree[6] = {-kSubLeft, 2, -kSubAbove, 4, -kSubZero, -kSubNew};
const int8_t kSplitTree[6] = {-3, 2, -2, 4, -0, -1};
const uint16_t kSplitFill[4][16] = {{0x00FF, 0xFF00},
                                    {0x3333, 0xCCCC},
                                    {0x0033, 0x00CC, 0x3300, 0xCC00},
                                    {0x0001, 0x0002, 0x0004, 0x0008, 0x0010, 0x0020, 0x0040, 0x0080, 0x0100, 0x0200,
                                     0x0400, 0x0800, 0x1000, 0x2000, 0x4000, 0x8000}};
const uint8_t kSplitCount[4] = {2, 2, 4, 16};
const uint8_t kBand[16] = {0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7};
const uint8_t kZigzag[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};

// write `value` with the tree the decoder walks (inverse of tree.cc:35-57).  The path to every leaf
// is looked up once per tree (the trees are the handful of constant arrays above) and cached.
struct TreePaths {
  // per leaf value: up to 8 steps of (probability index, bit)
  uint8_t len[32] = {0};
  uint8_t prob_index[32][8];
  uint8_t bit[32][8];
  void build(const int8_t* nodes, int at, uint8_t* pi, uint8_t* bi, int depth) {
    for (int b = 0; b < 2; b++) {
      pi[depth] = static_cast<uint8_t>(at >> 1);
      bi[depth] = static_cast<uint8_t>(b);
      const int n = nodes[at + b];
      if (n <= 0) {
        const int v = -n;
        len[v] = static_cast<uint8_t>(depth + 1);
        memcpy(prob_index[v], pi, depth + 1);
        memcpy(bit[v], bi, depth + 1);
      } else {
        build(nodes, n, pi, bi, depth + 1);
      }
    }
  }
  explicit TreePaths(const int8_t* nodes) {
    uint8_t pi[8], bi[8];
    build(nodes, 0, pi, bi, 0);
  }
};
inline void write_path(BoolWriter& bw, const TreePaths& t, const uint8_t* probs, int value) {
  for (int k = 0; k < t.len[value]; k++) bw.put(t.bit[value][k], probs[t.prob_index[value][k]]);
}
const TreePaths kKfYModePaths(kKfYModeTree), kYModePaths(kYModeTree), kUvModePaths(kUvModeTree), kBModePaths(kBModeTree),
    kSmallMvPaths(kSmallMvTree), kMvRefPaths(kMvRefTree), kSubMvPaths(kSubMvTree), kSplitPaths(kSplitTree);

// inverse of MotionVector::read_component (macroblock.cc:198-229); v in 1/8 pel, luma values even
void write_mv_component(BoolWriter& bw, int v, const uint8_t* p) {
  const int a = abs(v) >> 1;
  if (a < 8) {
    bw.put(0, p[0]);
    write_path(bw, kSmallMvPaths, p + 2, a);
  } else {
    bw.put(1, p[0]);
    for (int i = 0; i < 3; i++) bw.put((a >> i) & 1, p[9 + i]);
    for (int i = 9; i > 3; i--) bw.put((a >> i) & 1, p[9 + i]);
    if (a & 0xFFF0) bw.put((a >> 3) & 1, p[9 + 3]);
  }
  if (a) bw.put(v < 0, p[1]);
}

struct Mv {
  int x, y;
};
struct MbInfo {
  uint8_t inter = 0, y_mode = 0, flipped = 0;
  uint8_t bm[16] = {0};
  int16_t mv[16][2] = {{0, 0}};
};
struct Bounds {
  int left, right, top, bottom;
};
inline Mv clamp_mv(Mv m, const Bounds& b) {
  m.x = m.x < b.left ? b.left : (m.x > b.right ? b.right : m.x);
  m.y = m.y < b.top ? b.top : (m.y > b.bottom ? b.bottom : m.y);
  return m;
}
// Scorer (scorer.hh:35-78 + macroblock.cc:141-171): a neighbour whose reference has the other sign
// bias contributes its vector negated
struct Census {
  int score[4] = {0, 0, 0, 0};
  Mv mv[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  int index = 0, split_score = 0;
  bool flipped = false;
  void add(int weight, const MbInfo* nb) {
    if (!nb || !nb->inter) return;
    int x = nb->mv[15][0], y = nb->mv[15][1];
    if ((nb->flipped != 0) != flipped) {
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
    if (nb->y_mode == VP8GPU_SPLITMV) split_score += weight;
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
inline uint8_t implied_bmode(int y_mode) {
  static const uint8_t t[4] = {VP8GPU_B_DC_PRED, VP8GPU_B_VE_PRED, VP8GPU_B_HE_PRED, VP8GPU_B_TM_PRED};
  return t[y_mode];
}

// One arithmetic-coded decision of the token partition, recorded first so that branch statistics can
// choose the frame's probabilities before anything is written: (slot << 1) | bit, where slot < 1056
// indexes the coefficient probability table and 1056 + p stands for the fixed probability p.
// The statistics (how often each decision came out 0 / 1) are kept while recording: cnt[(slot << 1) | bit].
struct TokenRecorder {
  std::vector<uint16_t> bits;
  uint32_t* cnt = nullptr;
  void coef(int slot, int bit) {
    const uint16_t d = static_cast<uint16_t>((slot << 1) | bit);
    bits.push_back(d);
    cnt[d]++;
  }
  void fixed(int prob, int bit) {
    const uint16_t d = static_cast<uint16_t>(((1056 + prob) << 1) | bit);
    bits.push_back(d);
    cnt[d]++;
  }
};
// When the frame's probabilities are known before its tokens are walked (a parsed frame written back, a size
// estimate, a frame without probability optimisation) the decisions go straight into the arithmetic coder.
struct DirectWriter {
  BoolWriter& bw;
  const uint8_t* probs;  // the frame's 1056 coefficient probabilities
  void coef(int slot, int bit) { bw.put(bit, probs[slot]); }
  void fixed(int prob, int bit) { bw.put(bit, prob); }
};

template <class Sink>
inline void put_extra(Sink& t, int v, const uint8_t* probs, int n) {
  for (int i = n - 1; i >= 0; i--) t.fixed(probs[n - 1 - i], (v >> i) & 1);
}

// inverse of Block::parse_tokens (tokens.cc:50-135); returns has_nonzero
template <class Sink>
inline int record_block(Sink& t, const int16_t* coefs /* raster order */, int type, int ctx, int first) {
  static const uint8_t cat2[2] = {165, 145}, cat3[3] = {173, 148, 140}, cat4[4] = {176, 155, 140, 135},
                       cat5[5] = {180, 157, 141, 134, 130},
                       cat6[11] = {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129};
  int last = -1;
  for (int i = first; i < 16; i++)
    if (coefs[kZigzag[i]]) last = i;
  bool prev_zero = false;
  for (int i = first; i <= last; i++) {
    const int base = ((type * 8 + kBand[i]) * 3 + ctx) * 11;
    const int v = coefs[kZigzag[i]];
    const int a = abs(v);
    if (!prev_zero) t.coef(base + 0, 1);  // not end of block
    if (a == 0) {
      t.coef(base + 1, 0);
      prev_zero = true;
      ctx = 0;
      continue;
    }
    prev_zero = false;
    t.coef(base + 1, 1);
    if (a == 1) {
      t.coef(base + 2, 0);
      ctx = 1;
    } else {
      ctx = 2;
      t.coef(base + 2, 1);
      if (a <= 4) {
        t.coef(base + 3, 0);
        if (a == 2) {
          t.coef(base + 4, 0);
        } else {
          t.coef(base + 4, 1);
          t.coef(base + 5, a == 4);
        }
      } else {
        t.coef(base + 3, 1);
        if (a <= 10) {
          t.coef(base + 6, 0);
          if (a <= 6) {
            t.coef(base + 7, 0);
            t.fixed(159, a - 5);
          } else {
            t.coef(base + 7, 1);
            put_extra(t, a - 7, cat2, 2);
          }
        } else {
          t.coef(base + 6, 1);
          if (a <= 34) {
            t.coef(base + 8, 0);
            if (a <= 18) {
              t.coef(base + 9, 0);
              put_extra(t, a - 11, cat3, 3);
            } else {
              t.coef(base + 9, 1);
              put_extra(t, a - 19, cat4, 4);
            }
          } else {
            t.coef(base + 8, 1);
            if (a <= 66) {
              t.coef(base + 10, 0);
              put_extra(t, a - 35, cat5, 5);
            } else {
              t.coef(base + 10, 1);
              put_extra(t, a - 67, cat6, 11);
            }
          }
        }
      }
    }
    t.fixed(128, v < 0);
  }
  if (last < 15) {
    // end of block -- impossible to signal right after a zero token, but `last` is non-zero
    const int i = last + 1 < first ? first : last + 1;
    const int base = ((type * 8 + kBand[i]) * 3 + ctx) * 11;
    t.coef(base + 0, 0);
  }
  return last >= first;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// serialize_frame
// ------------------------------------------------------------------------------------------
std::vector<uint8_t> serialize_frame(const EncodeHeader& h, const vp8gpu_mb* mbs, const vp8gpu_token* tokens,
                                     const vp8gpu_split_mvs* split, const EncodeFeatures* features) {
  static const EncodeFeatures kPlain;
  const EncodeFeatures& x = features ? *features : kPlain;
  if (x.log2_partitions < 0 || x.log2_partitions > 3) return {};
  const int nparts = 1 << x.log2_partitions;
  const Verbatim* const vb = x.verbatim;  // re-serialisation of a parsed frame: header and labels as coded
  const int cols = (h.width + 15) / 16, rows = (h.height + 15) / 16;
  const size_t n_mbs = static_cast<size_t>(cols) * rows;
  std::vector<MbInfo> info(n_mbs);
  static const bool ser_trace = getenv("VP8GPU_SER_TRACE") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double tt[8] = {0};
  tt[0] = now();

  // ---- pass 0 (serial, cheap): which blocks of every macroblock hold a non-zero coefficient, and from
  //      that the "above" token contexts of every macroblock.  The Y2 context of a column / row is the
  //      last macroblock WITH a Y2 block (frame.cc:252-269), so it is carried through, not looked up. ----
  std::vector<uint32_t> nzmask(n_mbs, 0);
  std::vector<uint16_t> above_ctx(n_mbs, 0);  // bits 0-3 Y, 4-5 U, 6-7 V, 8 Y2
  std::vector<uint8_t> skip(n_mbs, 0);
  size_t n_skipped = 0;
  {
    std::vector<uint16_t> above(static_cast<size_t>(cols), 0);
    for (size_t idx = 0; idx < n_mbs; idx++) {
      const vp8gpu_mb& mb = mbs[idx];
      const bool has_y2 = mb.y_mode != VP8GPU_B_PRED && mb.y_mode != VP8GPU_SPLITMV;
      uint32_t m = 0;
      for (unsigned t = 0; t < mb.tok_cnt; t++) {
        const uint32_t tk = tokens[mb.tok_off + t];
        const unsigned blk = (tk >> 20) & 31, pos = (tk >> 16) & 15;
        if (blk > 24) return {};
        if ((tk & 0xFFFF) && !(has_y2 && blk < 16 && pos == 0)) m |= 1u << blk;
      }
      nzmask[idx] = m;
      uint16_t& a = above[idx % cols];
      above_ctx[idx] = a;
      // mb_skip_coeff: the encoder path skips exactly the macroblocks without coefficients; a parsed frame
      // keeps the flag it was coded with (a macroblock may spell out 25 empty blocks instead)
      const bool coded_skip = vb ? (vb->has_skip_prob && (vb->mb_coded[idx] & 1)) : mb.tok_cnt == 0;
      if (coded_skip) {
        if (mb.tok_cnt) return {};
        skip[idx] = 1;
        n_skipped++;
        a = has_y2 ? 0 : static_cast<uint16_t>(a & 0x100);
        continue;
      }
      // bottom row of Y blocks (12-15), bottom blocks of U (18, 19) and V (22, 23), Y2
      uint16_t na = static_cast<uint16_t>(((m >> 12) & 15) | (((m >> 18) & 3) << 4) | (((m >> 22) & 3) << 6));
      na |= has_y2 ? static_cast<uint16_t>(((m >> 24) & 1) << 8) : static_cast<uint16_t>(a & 0x100);
      a = na;
    }
  }

  tt[1] = now();
  // ---- the frame's coefficient probabilities before any update of this frame ----
  uint8_t coef_probs[1056];
  if (x.saved_coef_probs && h.key_frame) memcpy(x.saved_coef_probs, k_coef_default_probs, 1056);
  memcpy(coef_probs, x.saved_coef_probs ? x.saved_coef_probs : k_coef_default_probs, sizeof(coef_probs));
  if (vb) memcpy(coef_probs, vb->coef, sizeof(coef_probs));
  // they are final already when nothing will be derived from this frame's own statistics: a parsed frame written
  // back, a size estimate, a frame without probability optimisation -- then the tokens are coded in one walk
  const bool direct = vb || (x.ref_writer ? x.ref_estimate : !h.optimize_token_probs);

  // ---- pass 1: the decisions of the token partitions.  Row r belongs to partition r % n
  //      (frame.cc:131-136) and, given the above contexts, rows are independent: one recorder (and, with
  //      more than one partition, one thread) per partition ----
  struct PartWork {
    TokenRecorder rec;
    std::vector<uint32_t> cnt;
    std::vector<uint32_t> ref_skip_eob;  // reference writer policy: end-of-block counts of skipped macroblocks
    std::vector<uint8_t> bytes;
  };
  // Recording (as opposed to coding) a partition needs nothing from the row before once the above contexts are known,
  // so a partition's rows are recorded in `nsub` contiguous chunks on as many threads -- the reference's writer has
  // ONE partition, which would otherwise leave this pass serial -- and replayed into the coder chunk after chunk.
  // work[p * nsub + k]: chunk k of partition p; the partition's bytes end up in chunk 0.
  const int nsub = (!direct && nparts == 1 && n_mbs >= 256) ? 4 : 1;  // (CIF and up: the small golden-vector sizes exercise it too)
  std::vector<PartWork> work(static_cast<size_t>(nparts) * nsub);
  auto record_partition = [&](int pk) {
    const int p = pk / nsub, k = pk % nsub;
    const int part_rows = (rows - p + nparts - 1) / nparts;  // rows p, p + nparts, ... of this partition
    const int first = static_cast<int>(static_cast<int64_t>(part_rows) * k / nsub), last = static_cast<int>(static_cast<int64_t>(part_rows) * (k + 1) / nsub);
    PartWork& W = work[pk];
    const double t_begin = ser_trace ? now() : 0.0;
    W.cnt.assign(2 * (1056 + 256), 0);
    W.ref_skip_eob.assign(1056, 0);
    auto walk = [&](auto& sink) {
    int16_t c[25][16];  // coefficients of the current macroblock, raster order; all zero between macroblocks
    memset(c, 0, sizeof(c));
    for (int row = p + first * nparts; row < p + last * nparts && row < rows; row += nparts) {
      unsigned left = 0;  // same bit layout as above_ctx
      for (int col = 0; col < cols; col++) {
        const size_t idx = static_cast<size_t>(row) * cols + col;
        const vp8gpu_mb& mb = mbs[idx];
        const bool has_y2 = mb.y_mode != VP8GPU_B_PRED && mb.y_mode != VP8GPU_SPLITMV;
        if (skip[idx]) {
          if (x.ref_writer) {
            // the reference counts the (immediate) end of block of every Y / U / V block of a skipped
            // macroblock too (Macroblock::accumulate_token_branches runs for every macroblock)
            unsigned a = above_ctx[idx], l = left;
            auto eob = [&](int type, int bx, int by, int first) {
              const int ctx = ((a >> bx) & 1) + ((l >> by) & 1);
              W.ref_skip_eob[((type * 8 + kBand[first]) * 3 + ctx) * 11]++;
              a &= ~(1u << bx);
              l &= ~(1u << by);
            };
            const int ytype = has_y2 ? 0 : 3, yfirst = has_y2 ? 1 : 0;
            for (int i = 0; i < 16; i++) eob(ytype, i & 3, i >> 2, yfirst);
            for (int pl = 0; pl < 2; pl++)
              for (int i = 0; i < 4; i++) eob(2, 4 + 2 * pl + (i & 1), 4 + 2 * pl + (i >> 1), 0);
          }
          left = has_y2 ? 0 : (left & 0x100);
          continue;
        }
        const uint32_t m = nzmask[idx];
        for (unsigned t = 0; t < mb.tok_cnt; t++) {
          const uint32_t tk = tokens[mb.tok_off + t];
          c[(tk >> 20) & 31][(tk >> 16) & 15] = static_cast<int16_t>(tk & 0xFFFF);
        }
        unsigned a = above_ctx[idx];
        auto block = [&](int blk, int type, int bx, int by, int first) {
          const int ctx = ((a >> bx) & 1) + ((left >> by) & 1);
          unsigned nz = 0;
          if (m >> blk & 1) nz = static_cast<unsigned>(record_block(sink, c[blk], type, ctx, first));
          else sink.coef(((type * 8 + kBand[first]) * 3 + ctx) * 11, 0);  // immediate end of block
          a = (a & ~(1u << bx)) | (nz << bx);
          left = (left & ~(1u << by)) | (nz << by);
        };
        if (has_y2) block(24, 1, 8, 8, 0);
        const int ytype = has_y2 ? 0 : 3, yfirst = has_y2 ? 1 : 0;
        for (int i = 0; i < 16; i++) block(i, ytype, i & 3, i >> 2, yfirst);
        for (int pl = 0; pl < 2; pl++)
          for (int i = 0; i < 4; i++) block(16 + 4 * pl + i, 2, 4 + 2 * pl + (i & 1), 4 + 2 * pl + (i >> 1), 0);
        for (unsigned t = 0; t < mb.tok_cnt; t++) {
          const uint32_t tk = tokens[mb.tok_off + t];
          c[(tk >> 20) & 31][(tk >> 16) & 15] = 0;
        }
      }
    }
    };
    if (direct) {
      BoolWriter tw;
      tw.reserve(n_mbs * 24 / nparts + 1024);
      DirectWriter dw{tw, coef_probs};
      walk(dw);
      W.bytes = tw.finish();
    } else {
      W.rec.bits.reserve(n_mbs * 160 / (nparts * nsub) + 1024);
      W.rec.cnt = W.cnt.data();
      walk(W.rec);
    }
    if (ser_trace) fprintf(stderr, "  chunk %d: rows %d..%d  %.2f ms (started at +%.2f)\n", pk, first, last, now() - t_begin, t_begin - tt[1]);
  };
  const bool threaded = nsub > 1 || (nparts > 1 && n_mbs >= 1024);
  {
    HostPool::Group g;
    for (int p = 1; p < nparts * nsub && threaded; p++) g.run([&record_partition, p] { record_partition(p); });
    record_partition(0);
    for (int p = 1; p < nparts * nsub && !threaded; p++) record_partition(p);
    g.wait();
  }

  tt[2] = now();
  // ---- frame probabilities ----
  std::vector<uint8_t> updated(1056, 0);
  auto calc_prob = [](uint64_t falses, uint64_t total) -> int {  // Encoder::calc_prob (encoder.cc:48-55)
    if (falses == 0) return 0;
    const uint64_t p = 256 * falses / total;
    return static_cast<int>(p < 1 ? 1 : (p > 255 ? 255 : p));
  };
  if (vb) {
    // nothing to choose: the header is replayed
  } else if (x.ref_writer) {
    EncodeFeatures::RefWriterState& st = *x.ref_writer;
    if (!x.ref_estimate) {
      // Encoder::optimize_probability_tables (encoder.cc:419-440) over the reference's branch counts
      for (int i = 0; i < 1056; i++) {
        if (i / 264 == 1) continue;  // Y2 blocks are never counted
        uint64_t f = 0, t = 0;
        for (const PartWork& W : work) {
          f += W.cnt[2 * i];
          t += W.cnt[2 * i + 1];
          if (i % 11 == 0) f += W.ref_skip_eob[i];
        }
        const int p = calc_prob(f, f + t);
        if (p > 0 && p != coef_probs[i]) {
          st.upd_flag[i] = 1;
          st.upd_val[i] = static_cast<uint8_t>(p);
        }
      }
      for (int i = 0; i < 1056; i++)
        if (st.upd_flag[i]) {
          updated[i] = 1;
          coef_probs[i] = st.upd_val[i];
        }
    }
  } else if (h.optimize_token_probs) {
    std::vector<uint32_t> cnt(2 * 1056, 0);
    for (const PartWork& W : work)
      for (int i = 0; i < 2 * 1056; i++) cnt[i] += W.cnt[i];
    for (int i = 0; i < 1056; i++) {
      const uint32_t total = cnt[2 * i] + cnt[2 * i + 1];
      if (!total) continue;
      int p = static_cast<int>((static_cast<uint64_t>(cnt[2 * i]) * 256 + total / 2) / total);
      p = p < 1 ? 1 : (p > 255 ? 255 : p);
      // update only when the saving on the coded bits pays for the 8 + 1 bit update (rough cost model)
      const int old = coef_probs[i];
      if (abs(p - old) >= 8 && total >= 32) {
        coef_probs[i] = static_cast<uint8_t>(p);
        updated[i] = 1;
      }
    }
  }
  int skip_prob = static_cast<int>(((n_mbs - n_skipped) * 256 + n_mbs / 2) / n_mbs);  // P(not skipped)
  skip_prob = skip_prob < 1 ? 1 : (skip_prob > 255 ? 255 : skip_prob);
  if (x.ref_writer) skip_prob = calc_prob(n_mbs - n_skipped, n_mbs);  // Encoder::optimize_prob_skip (encoder.cc:442-457)
  size_t n_inter = 0;
  for (size_t i = 0; i < n_mbs; i++) n_inter += mbs[i].ref_frame != VP8GPU_REF_CURRENT;
  int prob_inter = static_cast<int>(((n_mbs - n_inter) * 256 + n_mbs / 2) / n_mbs);  // P(intra) = P(bit 0)
  prob_inter = prob_inter < 1 ? 1 : (prob_inter > 255 ? 255 : prob_inter);
  size_t n_last = 0, n_golden = 0;
  for (size_t i = 0; i < n_mbs; i++) {
    n_last += mbs[i].ref_frame == VP8GPU_REF_LAST;
    n_golden += mbs[i].ref_frame == VP8GPU_REF_GOLDEN;
  }
  auto prob_of = [](size_t zeros, size_t total) {  // P(bit 0) scaled to 1..255
    if (!total) return 128;
    const int p = static_cast<int>((zeros * 256 + total / 2) / total);
    return p < 1 ? 1 : (p > 255 ? 255 : p);
  };
  int prob_last = features ? prob_of(n_last, n_inter) : 255;
  int prob_golden = features ? prob_of(n_golden, n_inter - n_last) : 128;
  if (x.ref_writer && !h.key_frame) {
    // Encoder::optimize_interframe_probs (encode_inter.cc:527-576): a zero estimate leaves the header field alone
    EncodeFeatures::RefWriterState& st = *x.ref_writer;
    const size_t n_other = n_inter - n_last;
    int p = calc_prob(n_mbs - n_inter, n_mbs);
    if (p > 0) st.prob_inter = p;
    p = calc_prob(n_last, n_inter);
    if (p > 0) st.prob_last = p;
    p = calc_prob(n_golden, n_other);
    if (p > 0) st.prob_golden = p;
    prob_inter = st.prob_inter, prob_last = st.prob_last, prob_golden = st.prob_golden;
  }
  if (vb) skip_prob = vb->skip_prob, prob_inter = vb->prob_inter, prob_last = vb->prob_last, prob_golden = vb->prob_golden;
  auto put_flagged_signed = [](BoolWriter& w, int v, int width) {  // frame_header.hh Flagged<Signed<width>>
    w.put(v != 0);
    if (v) {
      w.literal(abs(v), width);
      w.put(v < 0);
    }
  };

  // ---- token partitions, written while this thread writes the first partition ----
  uint8_t prob_of_slot[1056 + 256];
  memcpy(prob_of_slot, coef_probs, 1056);
  for (int p = 0; p < 256; p++) prob_of_slot[1056 + p] = static_cast<uint8_t>(p);
  auto write_partition = [&](int p) {
    BoolWriter tw;
    size_t n_bits = 0;
    for (int k = 0; k < nsub; k++) n_bits += work[p * nsub + k].rec.bits.size();
    tw.reserve(n_bits / 4 + 64);
    for (int k = 0; k < nsub; k++)
      for (const uint16_t b : work[p * nsub + k].rec.bits) tw.put(b & 1, prob_of_slot[b >> 1]);
    work[p * nsub].bytes = tw.finish();
  };
  HostPool::Group writers;  // every return path below waits for the writers (the destructor does)
  for (int p = 0; p < nparts && !direct; p++) {
    if (threaded) writers.run([&write_partition, p] { write_partition(p); });
    else write_partition(p);
  }

  tt[3] = now();
  // ---- first partition: frame header ----
  BoolWriter bw;
  const Verbatim* const ro = x.residue_of;
  if (ro && (vb || !x.ref_writer || h.key_frame || ro->key || ro->read_segment || ro->mark_qend > ro->header_tape.size())) return {};
  if (vb) {
    for (const uint16_t d : vb->header_tape) bw.put(d & 1, d >> 1);
  } else if (ro) {
    auto replay = [&](uint32_t from, uint32_t to) {
      for (uint32_t i = from; i < to; i++) bw.put(ro->header_tape[i] & 1, ro->header_tape[i] >> 1);
    };
    replay(ro->mark_begin, ro->mark_parts);  // update_segmentation, filter_type, loop_filter_level, sharpness_level, mode_lf_adjustments
    bw.literal(x.log2_partitions, 2);        // not copied: the new frame object's own field
    bw.literal(h.y_ac_qi, 7);
    replay(ro->mark_qdelta, ro->mark_qend);  // quant_indices = a copy of the source's with (possibly) another y_ac_qi
    const bool all = x.residue_refresh_all;
    const bool rg = all || ro->refresh_golden, ra = all || ro->refresh_alt;
    bw.put(rg);
    bw.put(ra);
    if (!rg) bw.literal(ro->copy_golden, 2);
    if (!ra) bw.literal(ro->copy_alt, 2);
    bw.put(ro->sign_golden);
    bw.put(ro->sign_alt);
    bw.put(ro->refresh_entropy);
    bw.put(all || ro->refresh_last);
    if (x.saved_coef_probs && ro->refresh_entropy) memcpy(x.saved_coef_probs, coef_probs, 1056);
    for (int i = 0; i < 1056; i++) {
      bw.put(updated[i], k_coef_update_probs[i]);
      if (updated[i]) bw.literal(coef_probs[i], 8);
    }
    bw.put(1);  // prob_skip_false.reset( calc_prob ): present, possibly 0
    bw.literal(skip_prob, 8);
    bw.literal(prob_inter, 8);
    bw.literal(prob_last, 8);
    bw.literal(prob_golden, 8);
    bw.put(0);  // intra_16x16_prob, intra_chroma_prob, mv_prob_update: the new frame object carries none
    bw.put(0);
    for (int i = 0; i < 38; i++) bw.put(0, k_mv_update_probs[i]);
  } else {
    if (h.key_frame) {
      bw.put(0);  // color_space
      bw.put(0);  // clamping_type
    }
    bw.put(x.segmentation_enabled);
    if (x.segmentation_enabled) {  // frame_header.hh:37-66
      bw.put(x.update_mb_segmentation_map);
      bw.put(x.update_segment_feature_data);
      if (x.update_segment_feature_data) {
        bw.put(x.segment_feature_absolute);
        for (int i = 0; i < 4; i++) put_flagged_signed(bw, x.segment_quant[i], 7);
        for (int i = 0; i < 4; i++) put_flagged_signed(bw, x.segment_lf[i], 6);
      }
      if (x.update_mb_segmentation_map)
        for (int i = 0; i < 3; i++) {
          bw.put(x.segment_tree_probs[i] != 255);
          if (x.segment_tree_probs[i] != 255) bw.literal(x.segment_tree_probs[i], 8);
        }
    }
    bw.put(0);  // filter_type: normal
    bw.literal(x.late_loop_filter_level ? x.late_loop_filter_level(x.late_ctx) : h.loop_filter_level, 6);
    bw.literal(h.sharpness, 3);
    if (x.ref_writer && x.ref_estimate) {
      bw.put(0);  // the sampled frame of a size estimate never gets loop-filter settings (size_estimation.cc)
    } else if (x.ref_writer) {
      // Encoder::apply_best_loopfilter_settings (encoder.cc:464-470): mode_lf_adjustments present, updated, all
      // eight deltas flagged with the value 0
      bw.put(1);
      bw.put(1);
      for (int i = 0; i < 8; i++) {
        bw.put(1);
        bw.literal(0, 6);
        bw.put(0);
      }
    } else {
      bw.put(x.lf_delta_enabled);
      if (x.lf_delta_enabled) {  // frame_header.hh:70-84
        bw.put(x.lf_delta_update);
        if (x.lf_delta_update) {
          for (int i = 0; i < 4; i++) put_flagged_signed(bw, x.ref_lf_delta[i], 6);
          for (int i = 0; i < 4; i++) put_flagged_signed(bw, x.mode_lf_delta[i], 6);
        }
      }
    }
    bw.literal(x.log2_partitions, 2);
    bw.literal(h.y_ac_qi, 7);
    if (x.from_key) {
      if (!x.from_key->key || x.from_key->mark_qend > x.from_key->header_tape.size()) return {};
      for (uint32_t i = x.from_key->mark_qdelta; i < x.from_key->mark_qend; i++)
        bw.put(x.from_key->header_tape[i] & 1, x.from_key->header_tape[i] >> 1);
    } else {
      put_flagged_signed(bw, x.y_dc_delta, 4);
      put_flagged_signed(bw, x.y2_dc_delta, 4);
      put_flagged_signed(bw, x.y2_ac_delta, 4);
      put_flagged_signed(bw, x.uv_dc_delta, 4);
      put_flagged_signed(bw, x.uv_ac_delta, 4);
    }
    // Without a saved table (stateless writer) refresh_entropy_probs = 0 whenever probabilities are
    // updated: the updates are then valid for this frame only and every frame is coded relative to
    // the default tables.
    const int refresh_entropy =
        x.saved_coef_probs ? x.refresh_entropy_probs : (features ? 0 : (h.optimize_token_probs ? 0 : 1));
    if (h.key_frame) {
      bw.put(refresh_entropy);
    } else {
      bw.put(x.refresh_golden);
      bw.put(x.refresh_alternate);
      if (!x.refresh_golden) bw.literal(x.copy_to_golden, 2);
      if (!x.refresh_alternate) bw.literal(x.copy_to_alternate, 2);
      bw.put(x.sign_bias_golden);
      bw.put(x.sign_bias_alternate);
      bw.put(refresh_entropy);
      bw.put(features ? x.refresh_last : 1);
    }
    if (x.saved_coef_probs && refresh_entropy) memcpy(x.saved_coef_probs, coef_probs, 1056);
    for (int i = 0; i < 1056; i++) {
      bw.put(updated[i], k_coef_update_probs[i]);
      if (updated[i]) bw.literal(coef_probs[i], 8);
    }
    bw.put(1);  // mb_no_coeff_skip
    bw.literal(skip_prob, 8);
    if (!h.key_frame) {
      bw.literal(prob_inter, 8);
      bw.literal(prob_last, 8);
      bw.literal(prob_golden, 8);
      if (x.from_key) {
        bw.put(1);
        for (int i = 0; i < 4; i++) bw.literal(k_ymode_default_probs[i], 8);
        bw.put(1);
        for (int i = 0; i < 3; i++) bw.literal(k_uvmode_default_probs[i], 8);
      } else {
        bw.put(0);           // intra_16x16_prob unchanged
        bw.put(0);           // intra_chroma_prob unchanged
      }
      for (int i = 0; i < 38; i++) bw.put(0, k_mv_update_probs[i]);  // motion vector probabilities unchanged
    }
  }
  const uint8_t(*mv_probs)[19] =
      vb ? vb->mv : (x.mv_probs ? x.mv_probs : reinterpret_cast<const uint8_t(*)[19]>(k_mv_default_probs));
  const uint8_t* const ymode_probs = vb ? vb->ymode : (x.ymode_probs ? x.ymode_probs : k_ymode_default_probs);
  const uint8_t* const uvmode_probs = vb ? vb->uvmode : (x.uvmode_probs ? x.uvmode_probs : k_uvmode_default_probs);
  const Verbatim* const labels_of = vb ? vb : ro;  // SPLITMV layouts and sub-block labels as the source frame coded them
  const bool write_segment = vb ? vb->read_segment : (x.segmentation_enabled && x.update_mb_segmentation_map);

  // ---- first partition: macroblock headers (macroblock.cc:44-71, 84-111, 343-456 inverted) ----
  for (int row = 0; row < rows; row++) {
    for (int col = 0; col < cols; col++) {
      const size_t idx = static_cast<size_t>(row) * cols + col;
      const vp8gpu_mb& mb = mbs[idx];
      MbInfo& me = info[idx];
      const MbInfo* above = row > 0 ? &info[idx - cols] : nullptr;
      const MbInfo* left = col > 0 ? &info[idx - 1] : nullptr;
      const MbInfo* above_left = (row > 0 && col > 0) ? &info[idx - cols - 1] : nullptr;
      if (write_segment) {
        // segment tree {2, 4, -0, -1, -2, -3} (modemv_data.cc): the first decision picks the pair
        uint8_t sp[3] = {static_cast<uint8_t>(x.segment_tree_probs[0]), static_cast<uint8_t>(x.segment_tree_probs[1]),
                         static_cast<uint8_t>(x.segment_tree_probs[2])};
        if (vb) memcpy(sp, vb->seg_tree_probs, 3);
        if (mb.segment_id > 3) return {};
        bw.put(mb.segment_id >> 1, sp[0]);
        bw.put(mb.segment_id & 1, sp[1 + (mb.segment_id >> 1)]);
      }
      if (!vb || vb->has_skip_prob) bw.put(skip[idx], skip_prob);
      me.y_mode = mb.y_mode;
      if (mb.ref_frame == VP8GPU_REF_CURRENT) {
        if (!h.key_frame) bw.put(0, prob_inter);
        if (h.key_frame) {
          write_path(bw, kKfYModePaths, k_kf_ymode_probs, mb.y_mode);
          for (int i = 0; i < 16; i++) {
            if (mb.y_mode == VP8GPU_B_PRED) {
              const int m = static_cast<int>((mb.b_modes >> (4 * i)) & 15);
              const int am = i >= 4 ? me.bm[i - 4] : (above ? above->bm[12 + (i & 3)] : VP8GPU_B_DC_PRED);
              const int lm = (i & 3) ? me.bm[i - 1] : (left ? left->bm[((i >> 2) & 3) * 4 + 3] : VP8GPU_B_DC_PRED);
              write_path(bw, kBModePaths, k_kf_bmode_probs + (am * 10 + lm) * 9, m);
              me.bm[i] = static_cast<uint8_t>(m);
            } else {
              me.bm[i] = implied_bmode(mb.y_mode);
            }
          }
          write_path(bw, kUvModePaths, k_kf_uvmode_probs, mb.uv_mode);
        } else {
          write_path(bw, kYModePaths, ymode_probs, mb.y_mode);
          if (mb.y_mode == VP8GPU_B_PRED)
            for (int i = 0; i < 16; i++) write_path(bw, kBModePaths, k_bmode_probs, static_cast<int>((mb.b_modes >> (4 * i)) & 15));
          write_path(bw, kUvModePaths, uvmode_probs, mb.uv_mode);
        }
        continue;
      }
      if (h.key_frame || mb.ref_frame > VP8GPU_REF_ALTREF || (!features && mb.ref_frame != VP8GPU_REF_LAST)) return {};
      me.inter = 1;
      me.flipped = (mb.ref_frame == VP8GPU_REF_GOLDEN && (ro ? ro->sign_golden : x.sign_bias_golden)) ||
                   (mb.ref_frame == VP8GPU_REF_ALTREF && (ro ? ro->sign_alt : x.sign_bias_alternate));
      bw.put(1, prob_inter);
      bw.put(mb.ref_frame != VP8GPU_REF_LAST, prob_last);
      if (mb.ref_frame != VP8GPU_REF_LAST) bw.put(mb.ref_frame == VP8GPU_REF_ALTREF, prob_golden);
      Census census;
      census.flipped = me.flipped != 0;
      census.add(2, above);
      census.add(2, left);
      census.add(1, above_left);
      census.finish();
      const uint8_t ref_probs[4] = {k_mv_count_probs[census.score[0] * 4 + 0], k_mv_count_probs[census.score[1] * 4 + 1],
                                    k_mv_count_probs[census.score[2] * 4 + 2], k_mv_count_probs[census.split_score * 4 + 3]};
      Bounds b;
      b.left = -((col * 16) << 3) - 128;
      b.right = (((cols - 1 - col) * 16) << 3) + 128;
      b.top = -((row * 16) << 3) - 128;
      b.bottom = (((rows - 1 - row) * 16) << 3) + 128;
      const Mv nearest = clamp_mv(census.mv[1], b), near = clamp_mv(census.mv[2], b), best = clamp_mv(census.mv[0], b);
      if (mb.y_mode == VP8GPU_SPLITMV) {
        const int16_t(*mv)[2] = split[mb.split_idx].mv;
        memcpy(me.mv, mv, sizeof(me.mv));
        int layout = 3;
        for (int cand = 0; cand < 3; cand++) {
          bool ok = true;
          for (int part = 0; part < kSplitCount[cand] && ok; part++) {
            const unsigned members = kSplitFill[cand][part];
            const int first = __builtin_ctz(members);
            for (unsigned m = members; m; m &= m - 1) {
              const int i = __builtin_ctz(m);
              if (mv[i][0] != mv[first][0] || mv[i][1] != mv[first][1]) ok = false;
            }
          }
          if (ok) {
            layout = cand;
            break;
          }
        }
        uint32_t labels = 0;
        if (labels_of) {
          if (mb.split_idx >= labels_of->sub_labels.size() || idx >= labels_of->mb_coded.size()) return {};
          layout = (labels_of->mb_coded[idx] >> 1) & 3;
          labels = labels_of->sub_labels[mb.split_idx];
        }
        write_path(bw, kMvRefPaths, ref_probs, VP8GPU_SPLITMV);
        write_path(bw, kSplitPaths, k_split_probs, layout);
        int16_t done[16][2];  // vectors as the decoder knows them so far
        memset(done, 0, sizeof(done));
        for (int part = 0; part < kSplitCount[layout]; part++) {
          const unsigned members = kSplitFill[layout][part];
          const int first = __builtin_ctz(members);
          const int bx = first & 3, by = first >> 2;
          const int lx = bx ? done[first - 1][0] : (left ? left->mv[by * 4 + 3][0] : 0);
          const int ly = bx ? done[first - 1][1] : (left ? left->mv[by * 4 + 3][1] : 0);
          const int ax = by ? done[first - 4][0] : (above ? above->mv[12 + bx][0] : 0);
          const int ay = by ? done[first - 4][1] : (above ? above->mv[12 + bx][1] : 0);
          const bool lz = (lx | ly) == 0, az = (ax | ay) == 0, same = lx == ax && ly == ay;
          int ctx = 0;
          if (same && lz) ctx = 4;
          else if (same) ctx = 3;
          else if (az) ctx = 2;
          else if (lz) ctx = 1;
          const int vx = mv[first][0], vy = mv[first][1];
          const uint8_t* sp = k_submv_ref_probs + ctx * 3;
          int label = kSubNew;
          if (vx == lx && vy == ly) label = kSubLeft;
          else if (vx == ax && vy == ay) label = kSubAbove;
          else if ((vx | vy) == 0) label = kSubZero;
          if (labels_of) label = static_cast<int>((labels >> (2 * part)) & 3);  // as coded (any label that decodes to the vector is legal)
          if (label != kSubNew) {
            write_path(bw, kSubMvPaths, sp, label);
          } else {
            write_path(bw, kSubMvPaths, sp, kSubNew);
            const int dx = vx - best.x, dy = vy - best.y;
            if (abs(dx) > 2046 || abs(dy) > 2046) return {};
            write_mv_component(bw, dy, mv_probs[0]);
            write_mv_component(bw, dx, mv_probs[1]);
          }
          for (unsigned m = members; m; m &= m - 1) {
            const int i = __builtin_ctz(m);
            done[i][0] = static_cast<int16_t>(vx);
            done[i][1] = static_cast<int16_t>(vy);
          }
        }
      } else {
        // unsplit: pick the cheapest representation that decodes to exactly (mv_x, mv_y)
        const int vx = mb.mv_x, vy = mb.mv_y;
        int mode;
        if ((vx | vy) == 0) mode = VP8GPU_ZEROMV;
        else if (vx == nearest.x && vy == nearest.y) mode = VP8GPU_NEARESTMV;
        else if (vx == near.x && vy == near.y) mode = VP8GPU_NEARMV;
        else mode = VP8GPU_NEWMV;
        // the record's own label wins when it decodes to the same vector (an encoder may have chosen NEARMV or
        // NEWMV where a cheaper label exists; a parsed stream is re-written the way it was written)
        if (mb.y_mode == VP8GPU_NEWMV || (mb.y_mode == VP8GPU_NEARMV && vx == near.x && vy == near.y) ||
            (mb.y_mode == VP8GPU_NEARESTMV && vx == nearest.x && vy == nearest.y) || (mb.y_mode == VP8GPU_ZEROMV && (vx | vy) == 0))
          mode = mb.y_mode;
        me.y_mode = static_cast<uint8_t>(mode);
        write_path(bw, kMvRefPaths, ref_probs, mode);
        if (mode == VP8GPU_NEWMV) {
          const int dx = vx - best.x, dy = vy - best.y;
          if (abs(dx) > 2046 || abs(dy) > 2046 || (dx & 1) || (dy & 1)) return {};
          write_mv_component(bw, dy, mv_probs[0]);
          write_mv_component(bw, dx, mv_probs[1]);
        }
        for (int i = 0; i < 16; i++) {
          me.mv[i][0] = static_cast<int16_t>(vx);
          me.mv[i][1] = static_cast<int16_t>(vy);
        }
      }
    }
  }
  const std::vector<uint8_t> first = bw.finish();

  writers.wait();

  tt[4] = now();
  if (ser_trace) fprintf(stderr, "serialize_frame: pass0 %.2f record %.2f probs+write %.2f first partition %.2f ms\n", tt[1] - tt[0], tt[2] - tt[1], tt[3] - tt[2], tt[4] - tt[3]);
  // ---- frame tag (uncompressed_chunk.cc:49-77 inverted) ----
  std::vector<uint8_t> out;
  const uint32_t tag = (h.key_frame ? 0u : 1u) | (0u << 1) | (static_cast<uint32_t>(h.show_frame) << 4) |
                       (static_cast<uint32_t>(first.size()) << 5);
  out.push_back(tag & 0xFF);
  out.push_back((tag >> 8) & 0xFF);
  out.push_back((tag >> 16) & 0xFF);
  if (h.key_frame) {
    out.push_back(0x9d);
    out.push_back(0x01);
    out.push_back(0x2a);
    out.push_back(h.width & 0xFF);
    out.push_back((h.width >> 8) & 0x3F);
    out.push_back(h.height & 0xFF);
    out.push_back((h.height >> 8) & 0x3F);
  }
  out.insert(out.end(), first.begin(), first.end());
  for (int p = 0; p + 1 < nparts; p++) {  // partition sizes, all but the last (uncompressed_chunk.cc:132-155)
    out.push_back(work[p * nsub].bytes.size() & 0xFF);
    out.push_back((work[p * nsub].bytes.size() >> 8) & 0xFF);
    out.push_back((work[p * nsub].bytes.size() >> 16) & 0xFF);
  }
  for (int p = 0; p < nparts; p++) out.insert(out.end(), work[p * nsub].bytes.begin(), work[p * nsub].bytes.end());
  return out;
}

std::vector<uint8_t> serialize_parsed(const ParsedFrame& frame) {
  const Verbatim& v = frame.verbatim;
  if (v.header_tape.empty() || v.mb_coded.size() != static_cast<size_t>(frame.desc.mb_cols) * frame.desc.mb_rows) return {};
  EncodeHeader h;
  h.key_frame = v.key;
  h.show_frame = v.show;
  h.width = v.width;
  h.height = v.height;
  EncodeFeatures x;
  x.log2_partitions = v.log2_parts;
  x.sign_bias_golden = v.sign_golden;
  x.sign_bias_alternate = v.sign_alt;
  x.verbatim = &v;
  return serialize_frame(h, frame.mbs.data(), frame.tokens.data(), frame.split.data(), &x);
}

}  // namespace vp8
