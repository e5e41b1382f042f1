/* vp8_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see vp8_oracle.h).
 *
 * Plain-C restatement of excamera/alfalfa's VP8 decoder, function by function.  Every
 * routine names the reference code it follows (paths relative to /root/reference/src).
 * Written for clarity and bit-exactness, not speed: scalar, raster order, no SIMD.
 */
#define _POSIX_C_SOURCE 200809L
#include "vp8_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "vp8_tables.h"

/* ============================ small helpers ============================ */

static inline uint8_t clamp255(int x) { return x < 0 ? 0 : (x > 255 ? 255 : (uint8_t)x); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* tokens.hh:59-60 */
static const uint8_t k_band[16] = {0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7};
static const uint8_t k_zigzag[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};

/* trees, modemv_data.cc:186-250 (leaf = -value, inner = index of next pair) */
static const int8_t t_kf_ymode[8] = {-VP8GPU_B_PRED, 2, 4, 6, -VP8GPU_DC_PRED, -VP8GPU_V_PRED,
                                     -VP8GPU_H_PRED, -VP8GPU_TM_PRED};
static const int8_t t_ymode[8] = {-VP8GPU_DC_PRED, 2, 4, 6, -VP8GPU_V_PRED, -VP8GPU_H_PRED,
                                  -VP8GPU_TM_PRED, -VP8GPU_B_PRED};
static const int8_t t_uvmode[6] = {-VP8GPU_DC_PRED, 2, -VP8GPU_V_PRED, 4, -VP8GPU_H_PRED,
                                   -VP8GPU_TM_PRED};
static const int8_t t_bmode[18] = {-VP8GPU_B_DC_PRED, 2, -VP8GPU_B_TM_PRED, 4, -VP8GPU_B_VE_PRED, 6,
                                   8, 12, -VP8GPU_B_HE_PRED, 10, -VP8GPU_B_RD_PRED,
                                   -VP8GPU_B_VR_PRED, -VP8GPU_B_LD_PRED, 14, -VP8GPU_B_VL_PRED, 16,
                                   -VP8GPU_B_HD_PRED, -VP8GPU_B_HU_PRED};
static const int8_t t_small_mv[14] = {2, 8, 4, 6, -0, -1, -2, -3, 10, 12, -4, -5, -6, -7};
static const int8_t t_mv_ref[8] = {-VP8GPU_ZEROMV, 2, -VP8GPU_NEARESTMV, 4, -VP8GPU_NEARMV, 6,
                                   -VP8GPU_NEWMV, -VP8GPU_SPLITMV};
enum { SUB_LEFT = 0, SUB_ABOVE, SUB_ZERO, SUB_NEW };
static const int8_t t_submv[6] = {-SUB_LEFT, 2, -SUB_ABOVE, 4, -SUB_ZERO, -SUB_NEW};
static const int8_t t_split[6] = {-3, 2, -2, 4, -0, -1};
static const int8_t t_segment[6] = {2, 4, -0, -1, -2, -3};

/* mv_partitions, modemv_data.cc:252-278: for each of the 4 layouts, the partition number
 * of each of the 16 sub-blocks (raster order). */
static const uint8_t k_split_part[4][16] = {
    {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1}, /* top / bottom   */
    {0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1}, /* left / right   */
    {0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3}, /* quarters       */
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}};
static const uint8_t k_split_count[4] = {2, 2, 4, 16};

/* ============================ bool decoder ============================ */
/* bool_decoder.hh:45-120 */
typedef struct {
  const uint8_t* p;
  size_t n;
  uint32_t range, value;
  int bit_count;
} boolr;

static void br_load(boolr* b) {
  if (b->n) {
    b->value |= *b->p++;
    b->n--;
  }
}
static void br_init(boolr* b, const uint8_t* p, size_t n) {
  b->p = p;
  b->n = n;
  b->range = 255;
  b->value = 0;
  b->bit_count = 0;
  br_load(b);
  b->value <<= 8;
  br_load(b);
}
static int br_get(boolr* b, int prob) {
  const uint32_t split = 1 + (((b->range - 1) * (uint32_t)prob) >> 8);
  const uint32_t SPLIT = split << 8;
  int ret;
  if (b->value >= SPLIT) {
    ret = 1;
    b->range -= split;
    b->value -= SPLIT;
  } else {
    ret = 0;
    b->range = split;
  }
  while (b->range < 128) {
    b->value <<= 1;
    b->range <<= 1;
    if (++b->bit_count == 8) {
      b->bit_count = 0;
      br_load(b);
    }
  }
  return ret;
}
/* vp8_header_structures.hh:51-83: Unsigned<w> is MSB first, Signed<w> = magnitude then sign */
static int br_uint(boolr* b, int width) {
  int v = 0;
  for (int i = 0; i < width; i++) v = (v << 1) | br_get(b, 128);
  return v;
}
static int br_sint(boolr* b, int width) {
  int v = br_uint(b, width);
  return br_get(b, 128) ? -v : v;
}
static int br_flagged_sint(boolr* b, int width) { return br_get(b, 128) ? br_sint(b, width) : 0; }
/* tree.cc:35-57 */
static int br_tree(boolr* b, const int8_t* nodes, const uint8_t* probs) {
  int i = 0;
  while ((i = nodes[i + br_get(b, probs[i >> 1])]) > 0) {
  }
  return -i;
}

/* ============================ rasters ============================ */

vp8o_raster* vp8o_raster_new(int width, int height) {
  vp8o_raster* r = (vp8o_raster*)calloc(1, sizeof(*r));
  r->w16 = 16 * ((width + 15) / 16);
  r->h16 = 16 * ((height + 15) / 16);
  r->y = (uint8_t*)calloc((size_t)r->w16 * r->h16, 1);
  r->u = (uint8_t*)calloc((size_t)r->w16 * r->h16 / 4, 1);
  r->v = (uint8_t*)calloc((size_t)r->w16 * r->h16 / 4, 1);
  return r;
}
void vp8o_raster_free(vp8o_raster* r) {
  if (!r) return;
  free(r->y);
  free(r->u);
  free(r->v);
  free(r);
}
void vp8o_raster_copy(vp8o_raster* dst, const vp8o_raster* src) {
  memcpy(dst->y, src->y, (size_t)src->w16 * src->h16);
  memcpy(dst->u, src->u, (size_t)src->w16 * src->h16 / 4);
  memcpy(dst->v, src->v, (size_t)src->w16 * src->h16 / 4);
}
size_t vp8o_raster_dump_display(const vp8o_raster* r, int width, int height, uint8_t* dst) {
  /* util/raster.cc:85-104 */
  uint8_t* p = dst;
  const int cw = (width + 1) / 2, ch = (height + 1) / 2;
  for (int y = 0; y < height; y++, p += width) memcpy(p, r->y + (size_t)y * r->w16, width);
  for (int y = 0; y < ch; y++, p += cw) memcpy(p, r->u + (size_t)y * (r->w16 / 2), cw);
  for (int y = 0; y < ch; y++, p += cw) memcpy(p, r->v + (size_t)y * (r->w16 / 2), cw);
  return (size_t)(p - dst);
}

/* ============================ decoder state ============================ */
/* decoder.hh:57-225 */
struct vp8o_state {
  int width, height, mb_cols, mb_rows;
  uint8_t coef_probs[1056];
  uint8_t ymode_probs[4];
  uint8_t uvmode_probs[3];
  uint8_t mv_probs[2][19];
  int seg_enabled, seg_abs; /* Optional<Segmentation> */
  int8_t seg_quant[4], seg_lf[4];
  uint8_t* seg_map;
  int lf_adj_enabled; /* Optional<FilterAdjustments> */
  int8_t ref_adj[4], mode_adj[4];
};

static void state_default_probs(vp8o_state* s) {
  memcpy(s->coef_probs, vp8t_coef_default_probs, sizeof(s->coef_probs));
  memcpy(s->ymode_probs, vp8t_ymode_default_probs, 4);
  memcpy(s->uvmode_probs, vp8t_uvmode_default_probs, 3);
  memcpy(s->mv_probs, vp8t_mv_default_probs, 38);
}

vp8o_state* vp8o_state_new(int width, int height) {
  vp8o_state* s = (vp8o_state*)calloc(1, sizeof(*s));
  s->width = width;
  s->height = height;
  s->mb_cols = (width + 15) / 16;
  s->mb_rows = (height + 15) / 16;
  s->seg_map = (uint8_t*)malloc((size_t)s->mb_cols * s->mb_rows);
  memset(s->seg_map, 3, (size_t)s->mb_cols * s->mb_rows);
  state_default_probs(s);
  return s;
}
void vp8o_state_free(vp8o_state* s) {
  if (!s) return;
  free(s->seg_map);
  free(s);
}

/* ============================ parsed frame buffers ============================ */

vp8o_parsed* vp8o_parsed_new(void) { return (vp8o_parsed*)calloc(1, sizeof(vp8o_parsed)); }
void vp8o_parsed_free(vp8o_parsed* p) {
  if (!p) return;
  free(p->mbs);
  free(p->tokens);
  free(p->split);
  free(p);
}
static void ensure_tokens(vp8o_parsed* p, size_t need) {
  if (need > p->tokens_cap) {
    p->tokens_cap = need * 2 + 4096;
    p->tokens = (vp8gpu_token*)realloc(p->tokens, p->tokens_cap * sizeof(vp8gpu_token));
  }
}
static void ensure_split(vp8o_parsed* p, size_t need) {
  if (need > p->split_cap) {
    p->split_cap = need * 2 + 64;
    p->split = (vp8gpu_split_mvs*)realloc(p->split, p->split_cap * sizeof(vp8gpu_split_mvs));
  }
}

/* ============================ frame header ============================ */
/* frame_header.hh:37-325 */
typedef struct {
  int key, show;
  int seg_enabled, seg_update_map, seg_update_data, seg_abs;
  int8_t seg_quant[4], seg_lf[4];
  uint8_t seg_tree_probs[3];
  int filter_type, lf_level, sharpness;
  int lf_adj_enabled, lf_delta_update;
  int8_t ref_upd[4], mode_upd[4];
  int log2_parts;
  int y_ac_qi, y_dc, y2_dc, y2_ac, uv_dc, uv_ac;
  int refresh_golden, refresh_alt, copy_golden, copy_alt, sign_golden, sign_alt;
  int refresh_entropy, refresh_last;
  int has_skip_prob, skip_prob;
  int prob_inter, prob_last, prob_golden;
} frame_hdr;

static void read_segmentation_and_filter(boolr* b, frame_hdr* h) {
  /* Flagged<UpdateSegmentation>, frame_header.hh:104-131 */
  h->seg_enabled = br_get(b, 128);
  h->seg_update_map = h->seg_update_data = h->seg_abs = 0;
  memset(h->seg_quant, 0, 4);
  memset(h->seg_lf, 0, 4);
  memset(h->seg_tree_probs, 255, 3);
  if (h->seg_enabled) {
    h->seg_update_map = br_get(b, 128);
    h->seg_update_data = br_get(b, 128);
    if (h->seg_update_data) {
      h->seg_abs = br_get(b, 128);
      for (int i = 0; i < 4; i++) h->seg_quant[i] = (int8_t)br_flagged_sint(b, 7);
      for (int i = 0; i < 4; i++) h->seg_lf[i] = (int8_t)br_flagged_sint(b, 6);
    }
    if (h->seg_update_map)
      for (int i = 0; i < 3; i++) h->seg_tree_probs[i] = br_get(b, 128) ? (uint8_t)br_uint(b, 8) : 255;
  }
  h->filter_type = br_get(b, 128);
  h->lf_level = br_uint(b, 6);
  h->sharpness = br_uint(b, 3);
  /* Flagged<Flagged<ModeRefLFDeltaUpdate>>, frame_header.hh:70-84 */
  h->lf_adj_enabled = br_get(b, 128);
  h->lf_delta_update = 0;
  memset(h->ref_upd, 0, 4);
  memset(h->mode_upd, 0, 4);
  if (h->lf_adj_enabled) {
    h->lf_delta_update = br_get(b, 128);
    if (h->lf_delta_update) {
      for (int i = 0; i < 4; i++) h->ref_upd[i] = (int8_t)br_flagged_sint(b, 6);
      for (int i = 0; i < 4; i++) h->mode_upd[i] = (int8_t)br_flagged_sint(b, 6);
    }
  }
  h->log2_parts = br_uint(b, 2);
  /* QuantIndices, frame_header.hh:37-66 */
  h->y_ac_qi = br_uint(b, 7);
  h->y_dc = br_flagged_sint(b, 4);
  h->y2_dc = br_flagged_sint(b, 4);
  h->y2_ac = br_flagged_sint(b, 4);
  h->uv_dc = br_flagged_sint(b, 4);
  h->uv_ac = br_flagged_sint(b, 4);
}

/* token_prob_update (frame_header.hh:133-150) applied as ProbabilityTables::coeff_prob_update
 * does (probability_tables.cc:73-89) */
static void read_coef_prob_updates(boolr* b, uint8_t* probs) {
  for (int i = 0; i < 1056; i++)
    if (br_get(b, vp8t_coef_update_probs[i])) probs[i] = (uint8_t)br_uint(b, 8);
}

/* ============================ quantizer / loop-filter level ============================ */
/* quantization.cc:66-93 */
static int clamp_q(int q) { return q < 0 ? 0 : (q > 127 ? 127 : q); }
static vp8gpu_quant make_quant(int y_ac_qi /* already uint8-wrapped */, const frame_hdr* h) {
  vp8gpu_quant q;
  q.y_ac = vp8t_ac_q[clamp_q(y_ac_qi)];
  q.y_dc = vp8t_dc_q[clamp_q(y_ac_qi + h->y_dc)];
  q.y2_ac = (uint16_t)(vp8t_ac_q[clamp_q(y_ac_qi + h->y2_ac)] * 155 / 100);
  q.y2_dc = (uint16_t)(vp8t_dc_q[clamp_q(y_ac_qi + h->y2_dc)] * 2);
  q.uv_ac = vp8t_ac_q[clamp_q(y_ac_qi + h->uv_ac)];
  q.uv_dc = vp8t_dc_q[clamp_q(y_ac_qi + h->uv_dc)];
  if (q.y2_ac < 8) q.y2_ac = 8;
  if (q.uv_dc > 132) q.uv_dc = 132;
  return q;
}

/* frame.cc:139-182 (segment level), loopfilter.cc:57-79 (adjust), macroblock.cc:621 (skip),
 * loopfilter.cc:85 (the one clamp) */
static int mb_filter_level(const vp8o_state* st, const frame_hdr* h, int seg, int ref, int y_mode) {
  if (!h->lf_level) return 0;
  int level = h->lf_level;
  if (st->seg_enabled) level = st->seg_lf[seg] + (st->seg_abs ? 0 : h->lf_level);
  if (st->lf_adj_enabled) {
    int mode_adj;
    if (ref == VP8GPU_REF_CURRENT) mode_adj = (y_mode == VP8GPU_B_PRED) ? st->mode_adj[0] : 0;
    else if (y_mode == VP8GPU_ZEROMV) mode_adj = st->mode_adj[1];
    else if (y_mode == VP8GPU_SPLITMV) mode_adj = st->mode_adj[3];
    else mode_adj = st->mode_adj[2];
    level += st->ref_adj[ref] + mode_adj;
  }
  if (level <= 0) return 0;
  return level > 63 ? 63 : level;
}

/* ============================ macroblock headers ============================ */

typedef struct {
  uint8_t y_mode, ref, flipped, inter;
  uint8_t bmodes[16];
  int16_t mv[16][2]; /* x, y per luma sub-block */
} mbinfo;

typedef struct { int x, y; } mvec;

/* MotionVector::read_component, macroblock.cc:198-229 */
static int read_mv_component(boolr* b, const uint8_t* p) {
  enum { IS_SHORT = 0, SIGN = 1, SHORT = 2, BITS = SHORT + 8 - 1, LONG_WIDTH = 10 };
  int x = 0;
  if (br_get(b, p[IS_SHORT])) {
    for (int i = 0; i < 3; i++) x += br_get(b, p[BITS + i]) << i;
    for (int i = LONG_WIDTH - 1; i > 3; i--) x += br_get(b, p[BITS + i]) << i;
    if (!(x & 0xFFF0) || br_get(b, p[BITS + 3])) x += 8;
  } else {
    x = br_tree(b, t_small_mv, p + SHORT);
  }
  x <<= 1;
  if (x && br_get(b, p[SIGN])) x = -x;
  return x;
}
/* MotionVector(BoolDecoder&, probs), macroblock.cc:282-287: y (row) first */
static mvec read_mv(boolr* b, const uint8_t probs[2][19]) {
  mvec m;
  m.y = read_mv_component(b, probs[0]);
  m.x = read_mv_component(b, probs[1]);
  return m;
}
/* Scorer::clamp, macroblock.cc:183-195 */
static mvec clamp_mv(mvec mv, int col, int row, int cols, int rows) {
  const int to_left = imax(-((col * 16) << 3) - 128, -32768);
  const int to_right = imin((((cols - 1 - col) * 16) << 3) + 128, 32767);
  const int to_top = imax(-((row * 16) << 3) - 128, -32768);
  const int to_bottom = imin((((rows - 1 - row) * 16) << 3) + 128, 32767);
  mv.x = imin(imax(mv.x, to_left), to_right);
  mv.y = imin(imax(mv.y, to_top), to_bottom);
  return mv;
}

/* Scorer (scorer.hh:35-78, macroblock.cc:141-171) */
typedef struct {
  int flipped, index, split_score;
  int scores[4];
  mvec mvs[4];
} census;
static void census_add(census* c, int score, const mbinfo* mb) {
  if (!mb || !mb->inter) return;
  mvec mv = {mb->mv[15][0], mb->mv[15][1]};
  if (mb->flipped != c->flipped) {
    mv.x = -mv.x;
    mv.y = -mv.y;
  }
  if (mv.x == 0 && mv.y == 0) {
    c->scores[0] += score;
  } else {
    if (!(mv.x == c->mvs[c->index].x && mv.y == c->mvs[c->index].y)) {
      c->index++;
      c->mvs[c->index] = mv;
    }
    c->scores[c->index] += score;
  }
  if (mb->y_mode == VP8GPU_SPLITMV) c->split_score += score;
}
static void census_calculate(census* c) {
  if (c->scores[3]) {
    if (c->mvs[c->index].x == c->mvs[1].x && c->mvs[c->index].y == c->mvs[1].y) c->scores[1] += c->scores[3];
  }
  if (c->scores[2] > c->scores[1]) {
    int t = c->scores[1];
    c->scores[1] = c->scores[2];
    c->scores[2] = t;
    mvec m = c->mvs[1];
    c->mvs[1] = c->mvs[2];
    c->mvs[2] = m;
  }
  if (c->scores[1] >= c->scores[0]) c->mvs[0] = c->mvs[1];
}

/* MotionVector::luma_to_chroma, macroblock.cc:289-299 -- restated where it is consumed
 * (reconstruct) since the records carry luma vectors only. */
static mvec chroma_mv(const int16_t mv[16][2], int cx, int cy) {
  const int a = (cy * 2) * 4 + cx * 2;
  const int x = (int16_t)(mv[a][0] + mv[a + 1][0] + mv[a + 4][0] + mv[a + 5][0]);
  const int y = (int16_t)(mv[a][1] + mv[a + 1][1] + mv[a + 4][1] + mv[a + 5][1]);
  mvec r;
  r.x = x >= 0 ? (x + 4) >> 3 : -((-x + 4) >> 3);
  r.y = y >= 0 ? (y + 4) >> 3 : -((-y + 4) >> 3);
  return r;
}

/* implied_subblock_mode, macroblock.hh:151-160 */
static int implied_bmode(int y_mode) {
  switch (y_mode) {
    case VP8GPU_DC_PRED: return VP8GPU_B_DC_PRED;
    case VP8GPU_V_PRED: return VP8GPU_B_VE_PRED;
    case VP8GPU_H_PRED: return VP8GPU_B_HE_PRED;
    default: return VP8GPU_B_TM_PRED;
  }
}

/* ============================ token parsing ============================ */
/* tokens.cc:50-135.  Returns has_nonzero of the block; appends tokens. */
static int parse_block(boolr* b, const uint8_t* coef_probs, int type, int ctx, int first, int blk,
                       vp8gpu_token** out) {
  /* TokenDecoder tables, tokens.hh:74-78 */
  static const uint8_t cat2[2] = {165, 145}, cat3[3] = {173, 148, 140}, cat4[4] = {176, 155, 140, 135},
                       cat5[5] = {180, 157, 141, 134, 130},
                       cat6[11] = {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129};
  int last_was_zero = 0, has_nonzero = 0;
  for (int index = first; index < 16; index++) {
    const uint8_t* prob = coef_probs + ((type * 8 + k_band[index]) * 3 + ctx) * 11;
    if (!last_was_zero) {
      if (!br_get(b, prob[0])) break; /* EOB */
    }
    if (!br_get(b, prob[1])) {
      last_was_zero = 1;
      ctx = 0;
      continue;
    }
    last_was_zero = 0;
    has_nonzero = 1;
    int value;
    if (!br_get(b, prob[2])) {
      value = 1;
      ctx = 1;
    } else {
      ctx = 2;
      if (!br_get(b, prob[3])) {
        if (!br_get(b, prob[4])) value = 2;
        else value = br_get(b, prob[5]) ? 4 : 3;
      } else {
        const uint8_t* extra;
        int nbits, base;
        if (!br_get(b, prob[6])) {
          if (!br_get(b, prob[7])) {
            value = 5 + br_get(b, 159);
            goto have_value;
          }
          extra = cat2, nbits = 2, base = 7;
        } else if (!br_get(b, prob[8])) {
          if (!br_get(b, prob[9])) extra = cat3, nbits = 3, base = 11;
          else extra = cat4, nbits = 4, base = 19;
        } else {
          if (!br_get(b, prob[10])) extra = cat5, nbits = 5, base = 35;
          else extra = cat6, nbits = 11, base = 67;
        }
        int inc = 0;
        for (int i = 0; i < nbits; i++) inc = (inc << 1) + br_get(b, extra[i]);
        value = base + inc;
      }
    }
  have_value:
    if (br_get(b, 128)) value = -value;
    *(*out)++ = VP8GPU_TOKEN(blk, k_zigzag[index], value);
  }
  return has_nonzero;
}

/* ============================ frame parse ============================ */

int vp8o_parse_frame(vp8o_state* st_io, const uint8_t* data, size_t len, vp8o_parsed* out) {
  /* ---- UncompressedChunk, uncompressed_chunk.cc:34-130 ---- */
  if (len < 3) return VP8GPU_ERR_INVALID;
  const uint32_t tag = data[0] | (data[1] << 8) | ((uint32_t)data[2] << 16);
  frame_hdr h;
  memset(&h, 0, sizeof(h));
  h.key = !(tag & 1);
  const int version = (tag >> 1) & 7;
  h.show = (tag >> 4) & 1;
  const uint32_t first_len = (tag >> 5) & 0x7FFFF;
  if (version != 0) return VP8GPU_ERR_UNSUPPORTED; /* 4/6 = experimental, others rejected */
  const size_t first_off = h.key ? 10 : 3;
  if (len <= first_off + first_len) return VP8GPU_ERR_INVALID;
  if (h.key) {
    if (data[3] != 0x9d || data[4] != 0x01 || data[5] != 0x2a) return VP8GPU_ERR_INVALID;
    const uint32_t sizes = data[6] | (data[7] << 8) | (data[8] << 16) | ((uint32_t)data[9] << 24);
    const int fw = sizes & 0x3FFF, hs = (sizes >> 14) & 3, fh = (sizes >> 16) & 0x3FFF, vs = (sizes >> 30) & 3;
    if (fw != st_io->width || fh != st_io->height || hs || vs) return VP8GPU_ERR_UNSUPPORTED;
  }
  const uint8_t* rest = data + first_off + first_len;
  size_t rest_len = len - first_off - first_len;

  /* work on a copy of the state, commit on success */
  vp8o_state st = *st_io;
  const size_t n_mbs = (size_t)st.mb_cols * st.mb_rows;
  uint8_t* seg_map = (uint8_t*)malloc(n_mbs);
  memcpy(seg_map, st_io->seg_map, n_mbs);
  st.seg_map = seg_map;
  mbinfo* info = NULL;
  uint8_t* above_nz = NULL;
  int rc = VP8GPU_OK;

  boolr b;
  br_init(&b, data + first_off, first_len);

  /* ---- frame header + DecoderState update, decoder_state.hh:73-167 ---- */
  uint8_t frame_coef[1056], frame_ymode[4], frame_uvmode[3], frame_mv[2][19];
  if (h.key) {
    const int color_space = br_get(&b, 128), clamping_type = br_get(&b, 128);
    read_segmentation_and_filter(&b, &h);
    h.refresh_entropy = br_get(&b, 128);
    /* *this = DecoderState(header, w, h): decoder.cc:236-243 */
    state_default_probs(&st);
    st.seg_enabled = h.seg_enabled;
    st.seg_abs = 0;
    memset(st.seg_quant, 0, 4);
    memset(st.seg_lf, 0, 4);
    if (h.seg_enabled) {
      memset(st.seg_map, 3, n_mbs); /* Segmentation::map( width, height, 3 ), decoder_state.hh:170-176 */
      if (h.seg_update_data) {
        st.seg_abs = h.seg_abs;
        memcpy(st.seg_quant, h.seg_quant, 4);
        memcpy(st.seg_lf, h.seg_lf, 4);
      }
    }
    st.lf_adj_enabled = h.lf_adj_enabled;
    memset(st.ref_adj, 0, 4);
    memset(st.mode_adj, 0, 4);
    if (h.lf_adj_enabled && h.lf_delta_update) {
      memcpy(st.ref_adj, h.ref_upd, 4);
      memcpy(st.mode_adj, h.mode_upd, 4);
    }
    memcpy(frame_coef, st.coef_probs, 1056);
    read_coef_prob_updates(&b, frame_coef);
    h.has_skip_prob = br_get(&b, 128);
    h.skip_prob = h.has_skip_prob ? br_uint(&b, 8) : 0;
    if (color_space || clamping_type || h.filter_type) {
      rc = VP8GPU_ERR_UNSUPPORTED;
      goto done;
    }
    if (h.refresh_entropy) memcpy(st.coef_probs, frame_coef, 1056);
    memcpy(frame_ymode, st.ymode_probs, 4); /* unused for key frames */
    memcpy(frame_uvmode, st.uvmode_probs, 3);
    memcpy(frame_mv, st.mv_probs, 38);
    h.refresh_last = h.refresh_golden = h.refresh_alt = 1;
  } else {
    read_segmentation_and_filter(&b, &h);
    h.refresh_golden = br_get(&b, 128);
    h.refresh_alt = br_get(&b, 128);
    h.copy_golden = h.refresh_golden ? 0 : br_uint(&b, 2);
    h.copy_alt = h.refresh_alt ? 0 : br_uint(&b, 2);
    h.sign_golden = br_get(&b, 128);
    h.sign_alt = br_get(&b, 128);
    h.refresh_entropy = br_get(&b, 128);
    h.refresh_last = br_get(&b, 128);
    memcpy(frame_coef, st.coef_probs, 1056);
    memcpy(frame_ymode, st.ymode_probs, 4);
    memcpy(frame_uvmode, st.uvmode_probs, 3);
    memcpy(frame_mv, st.mv_probs, 38);
    read_coef_prob_updates(&b, frame_coef);
    h.has_skip_prob = br_get(&b, 128);
    h.skip_prob = h.has_skip_prob ? br_uint(&b, 8) : 0;
    h.prob_inter = br_uint(&b, 8);
    h.prob_last = br_uint(&b, 8);
    h.prob_golden = br_uint(&b, 8);
    if (br_get(&b, 128))
      for (int i = 0; i < 4; i++) frame_ymode[i] = (uint8_t)br_uint(&b, 8);
    if (br_get(&b, 128))
      for (int i = 0; i < 3; i++) frame_uvmode[i] = (uint8_t)br_uint(&b, 8);
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 19; j++)
        if (br_get(&b, vp8t_mv_update_probs[i * 19 + j])) {
          const int x = br_uint(&b, 7);
          frame_mv[i][j] = (uint8_t)(x ? x << 1 : 1); /* MVProbUpdate::read_half_prob */
        }
    if (h.filter_type) {
      rc = VP8GPU_ERR_UNSUPPORTED;
      goto done;
    }
    /* ProbabilityTables::update + refresh, decoder_state.hh:126-130 */
    if (h.refresh_entropy) {
      memcpy(st.coef_probs, frame_coef, 1056);
      memcpy(st.ymode_probs, frame_ymode, 4);
      memcpy(st.uvmode_probs, frame_uvmode, 3);
      memcpy(st.mv_probs, frame_mv, 38);
    }
    /* filter adjustments, decoder_state.hh:132-141 + FilterAdjustments::update :54-65 */
    if (h.lf_adj_enabled) {
      if (!st.lf_adj_enabled) {
        memset(st.ref_adj, 0, 4);
        memset(st.mode_adj, 0, 4);
      }
      st.lf_adj_enabled = 1;
      if (h.lf_delta_update) {
        memcpy(st.ref_adj, h.ref_upd, 4);
        memcpy(st.mode_adj, h.mode_upd, 4);
      }
    } else {
      st.lf_adj_enabled = 0;
    }
    /* segmentation, decoder_state.hh:143-152 + Segmentation::update :37-52 */
    if (h.seg_enabled) {
      if (!st.seg_enabled) {
        st.seg_abs = 0;
        memset(st.seg_quant, 0, 4);
        memset(st.seg_lf, 0, 4);
        memset(st.seg_map, 3, n_mbs);
      }
      st.seg_enabled = 1;
      if (h.seg_update_data) {
        st.seg_abs = h.seg_abs;
        memcpy(st.seg_quant, h.seg_quant, 4);
        memcpy(st.seg_lf, h.seg_lf, 4);
      }
    } else {
      st.seg_enabled = 0;
    }
  }

  /* ---- frame descriptor ---- */
  vp8gpu_frame_desc* d = &out->desc;
  memset(d, 0, sizeof(*d));
  d->width = (uint16_t)st.width;
  d->height = (uint16_t)st.height;
  d->mb_cols = (uint16_t)st.mb_cols;
  d->mb_rows = (uint16_t)st.mb_rows;
  d->key_frame = (uint8_t)h.key;
  d->show_frame = (uint8_t)h.show;
  d->loop_filter_level = (uint8_t)h.lf_level;
  d->sharpness = (uint8_t)h.sharpness;
  d->refresh_last = (uint8_t)h.refresh_last;
  d->refresh_golden = (uint8_t)h.refresh_golden;
  d->refresh_alternate = (uint8_t)h.refresh_alt;
  d->copy_to_golden = (uint8_t)h.copy_golden;
  d->copy_to_alternate = (uint8_t)h.copy_alt;
  /* calculate_segment_quantizers, frame.cc:186-206: the index goes through Unsigned<7>,
   * i.e. a uint8_t, so negative sums wrap before clamp_q */
  for (int i = 0; i < 4; i++) {
    int qi = h.y_ac_qi;
    if (st.seg_enabled) qi = (uint8_t)(st.seg_quant[i] + (st.seg_abs ? 0 : h.y_ac_qi));
    d->quant[i] = make_quant(qi, &h);
  }

  if (out->mbs_cap < n_mbs) {
    out->mbs_cap = n_mbs;
    out->mbs = (vp8gpu_mb*)realloc(out->mbs, n_mbs * sizeof(vp8gpu_mb));
  }
  memset(out->mbs, 0, n_mbs * sizeof(vp8gpu_mb));
  info = (mbinfo*)calloc(n_mbs, sizeof(mbinfo));
  uint32_t n_split = 0;

  /* ---- pass 1: macroblock headers, frame.cc:96-113 ---- */
  for (int row = 0; row < st.mb_rows; row++) {
    for (int col = 0; col < st.mb_cols; col++) {
      const size_t idx = (size_t)row * st.mb_cols + col;
      vp8gpu_mb* mb = &out->mbs[idx];
      mbinfo* me = &info[idx];
      const mbinfo* above = row > 0 ? &info[idx - st.mb_cols] : NULL;
      const mbinfo* left = col > 0 ? &info[idx - 1] : NULL;
      const mbinfo* above_left = (row > 0 && col > 0) ? &info[idx - st.mb_cols - 1] : NULL;

      /* Macroblock ctor, macroblock.cc:44-71 */
      if (h.seg_enabled && h.seg_update_map) st.seg_map[idx] = (uint8_t)br_tree(&b, t_segment, h.seg_tree_probs);
      /* update_segmentation, macroblock.cc:73-82: id is read from the persistent map */
      mb->segment_id = st.seg_enabled ? st.seg_map[idx] : 0;
      const int skip = h.has_skip_prob ? br_get(&b, h.skip_prob) : 0;
      mb->reserved = (uint32_t)skip; /* scratch until pass 2 */

      int inter = 0, ref = VP8GPU_REF_CURRENT;
      if (!h.key) {
        /* InterFrameMacroblockHeader, macroblock.cc:458-466 */
        inter = br_get(&b, h.prob_inter);
        if (inter) {
          ref = VP8GPU_REF_LAST;
          if (br_get(&b, h.prob_last)) ref = br_get(&b, h.prob_golden) ? VP8GPU_REF_ALTREF : VP8GPU_REF_GOLDEN;
        }
      }
      me->inter = (uint8_t)inter;
      me->ref = (uint8_t)ref;
      me->flipped = (uint8_t)((ref == VP8GPU_REF_GOLDEN && h.sign_golden) || (ref == VP8GPU_REF_ALTREF && h.sign_alt));

      if (!inter) {
        int y_mode;
        if (h.key) {
          /* KeyFrameMacroblock::decode_prediction_modes, macroblock.cc:84-111 */
          y_mode = br_tree(&b, t_kf_ymode, vp8t_kf_ymode_probs);
          for (int i = 0; i < 16; i++) {
            if (y_mode == VP8GPU_B_PRED) {
              const int bx = i & 3, by = i >> 2;
              int am, lm;
              if (by > 0) am = me->bmodes[i - 4];
              else am = above ? above->bmodes[12 + bx] : VP8GPU_B_DC_PRED;
              if (bx > 0) lm = me->bmodes[i - 1];
              else lm = left ? left->bmodes[by * 4 + 3] : VP8GPU_B_DC_PRED;
              me->bmodes[i] = (uint8_t)br_tree(&b, t_bmode, vp8t_kf_bmode_probs + (am * 10 + lm) * 9);
            } else {
              me->bmodes[i] = (uint8_t)implied_bmode(y_mode);
            }
          }
          mb->uv_mode = (uint8_t)br_tree(&b, t_uvmode, vp8t_kf_uvmode_probs);
        } else {
          /* intra MB of an inter frame, macroblock.cc:354-374 */
          y_mode = br_tree(&b, t_ymode, frame_ymode);
          for (int i = 0; i < 16; i++)
            me->bmodes[i] = (uint8_t)(y_mode == VP8GPU_B_PRED ? br_tree(&b, t_bmode, vp8t_bmode_probs)
                                                              : implied_bmode(y_mode));
          mb->uv_mode = (uint8_t)br_tree(&b, t_uvmode, frame_uvmode);
        }
        me->y_mode = (uint8_t)y_mode;
        if (y_mode == VP8GPU_B_PRED) {
          uint64_t packed = 0;
          for (int i = 0; i < 16; i++) packed |= (uint64_t)me->bmodes[i] << (4 * i);
          mb->b_modes = packed;
        }
      } else {
        /* macroblock.cc:376-455 */
        census c;
        memset(&c, 0, sizeof(c));
        c.flipped = me->flipped;
        census_add(&c, 2, above);
        census_add(&c, 2, left);
        census_add(&c, 1, above_left);
        census_calculate(&c);
        const int counts[4] = {c.scores[0], c.scores[1], c.scores[2], c.split_score};
        uint8_t mv_ref_p[4];
        for (int i = 0; i < 4; i++) mv_ref_p[i] = vp8t_mv_count_probs[counts[i] * 4 + i];
        const int y_mode = br_tree(&b, t_mv_ref, mv_ref_p);
        me->y_mode = (uint8_t)y_mode;
        mvec base = {0, 0};
        switch (y_mode) {
          case VP8GPU_NEARESTMV: base = clamp_mv(c.mvs[1], col, row, st.mb_cols, st.mb_rows); break;
          case VP8GPU_NEARMV: base = clamp_mv(c.mvs[2], col, row, st.mb_cols, st.mb_rows); break;
          case VP8GPU_ZEROMV: break;
          case VP8GPU_NEWMV: {
            mvec nm = read_mv(&b, (const uint8_t(*)[19])frame_mv);
            const mvec best = clamp_mv(c.mvs[0], col, row, st.mb_cols, st.mb_rows);
            base.x = (int16_t)(nm.x + best.x);
            base.y = (int16_t)(nm.y + best.y);
            break;
          }
          default: { /* SPLITMV */
            const int layout = br_tree(&b, t_split, vp8t_split_probs);
            const mvec best = clamp_mv(c.mvs[0], col, row, st.mb_cols, st.mb_rows);
            for (int part = 0; part < k_split_count[layout]; part++) {
              int first = 0;
              while (k_split_part[layout][first] != part) first++;
              const int bx = first & 3, by = first >> 2;
              /* read_subblock_inter_prediction, macroblock.cc:231-280 */
              mvec lmv = {0, 0}, amv = {0, 0};
              if (bx > 0) { lmv.x = me->mv[first - 1][0]; lmv.y = me->mv[first - 1][1]; }
              else if (left) { lmv.x = left->mv[by * 4 + 3][0]; lmv.y = left->mv[by * 4 + 3][1]; }
              if (by > 0) { amv.x = me->mv[first - 4][0]; amv.y = me->mv[first - 4][1]; }
              else if (above) { amv.x = above->mv[12 + bx][0]; amv.y = above->mv[12 + bx][1]; }
              const int lz = (lmv.x == 0 && lmv.y == 0), az = (amv.x == 0 && amv.y == 0);
              const int eq = (lmv.x == amv.x && lmv.y == amv.y);
              int ctx = 0;
              if (eq && lz) ctx = 4;
              else if (eq) ctx = 3;
              else if (az) ctx = 2;
              else if (lz) ctx = 1;
              mvec m = {0, 0};
              switch (br_tree(&b, t_submv, vp8t_submv_ref_probs + ctx * 3)) {
                case SUB_LEFT: m = lmv; break;
                case SUB_ABOVE: m = amv; break;
                case SUB_ZERO: break;
                default: {
                  mvec nm = read_mv(&b, (const uint8_t(*)[19])frame_mv);
                  m.x = (int16_t)(nm.x + best.x);
                  m.y = (int16_t)(nm.y + best.y);
                }
              }
              for (int i = 0; i < 16; i++)
                if (k_split_part[layout][i] == part) {
                  me->mv[i][0] = (int16_t)m.x;
                  me->mv[i][1] = (int16_t)m.y;
                }
            }
            base.x = me->mv[15][0];
            base.y = me->mv[15][1];
            ensure_split(out, n_split + 1);
            memcpy(out->split[n_split].mv, me->mv, sizeof(me->mv));
            mb->split_idx = n_split++;
          }
        }
        if (y_mode != VP8GPU_SPLITMV)
          for (int i = 0; i < 16; i++) {
            me->mv[i][0] = (int16_t)base.x;
            me->mv[i][1] = (int16_t)base.y;
          }
        mb->mv_x = (int16_t)base.x;
        mb->mv_y = (int16_t)base.y;
      }
      mb->y_mode = me->y_mode;
      mb->ref_frame = me->ref;
      mb->flags = (me->y_mode != VP8GPU_B_PRED && me->y_mode != VP8GPU_SPLITMV) ? VP8GPU_MB_HAS_Y2 : 0;
      mb->lf_level = (uint8_t)mb_filter_level(&st, &h, mb->segment_id, me->ref, me->y_mode);
    }
  }

  /* ---- pass 2: tokens, frame.cc:122-137 + dct_partitions uncompressed_chunk.cc:132-155 ---- */
  {
    const int nparts = 1 << h.log2_parts;
    boolr parts[8];
    if (rest_len < (size_t)3 * (nparts - 1)) {
      rc = VP8GPU_ERR_INVALID;
      goto done;
    }
    const uint8_t* pdata = rest + 3 * (nparts - 1);
    size_t pleft = rest_len - 3 * (nparts - 1);
    for (int i = 0; i < nparts; i++) {
      size_t plen = pleft;
      if (i < nparts - 1) {
        plen = rest[3 * i] | (rest[3 * i + 1] << 8) | ((size_t)rest[3 * i + 2] << 16);
        if (plen > pleft) {
          rc = VP8GPU_ERR_INVALID;
          goto done;
        }
      }
      br_init(&parts[i], pdata, plen);
      pdata += plen;
      pleft -= plen;
    }
    /* above / left has_nonzero contexts: 4 Y, 2 U, 2 V, 1 Y2 per MB column / row */
    above_nz = (uint8_t*)calloc((size_t)st.mb_cols, 9);
    size_t n_tok = 0;
    for (int row = 0; row < st.mb_rows; row++) {
      uint8_t left_nz[9] = {0};
      boolr* pb = &parts[row % nparts];
      for (int col = 0; col < st.mb_cols; col++) {
        vp8gpu_mb* mb = &out->mbs[(size_t)row * st.mb_cols + col];
        uint8_t* a = above_nz + (size_t)col * 9;
        const int skip = (int)mb->reserved;
        mb->reserved = 0;
        mb->tok_off = (uint32_t)n_tok;
        const int has_y2 = mb->flags & VP8GPU_MB_HAS_Y2;
        if (skip) {
          /* Macroblock::parse_tokens returns early (macroblock.cc:479-481): every block of the
           * MB keeps has_nonzero_ = false; a coded Y2 becomes the new (zero) context, a missing
           * Y2 leaves the previous one in place (relink_y2_blocks, frame.cc:252-269) */
          memset(a, 0, 8);
          memset(left_nz, 0, 8);
          if (has_y2) a[8] = left_nz[8] = 0;
          mb->tok_cnt = 0;
          continue;
        }
        ensure_tokens(out, n_tok + 400);
        vp8gpu_token* t = out->tokens + n_tok;
        if (has_y2) a[8] = left_nz[8] = (uint8_t)parse_block(pb, frame_coef, 1, a[8] + left_nz[8], 0, VP8GPU_BLK_Y2, &t);
        const int ytype = has_y2 ? 0 : 3, yfirst = has_y2 ? 1 : 0;
        for (int i = 0; i < 16; i++) {
          const int bx = i & 3, by = i >> 2;
          a[bx] = left_nz[by] = (uint8_t)parse_block(pb, frame_coef, ytype, a[bx] + left_nz[by], yfirst, i, &t);
        }
        for (int i = 0; i < 4; i++) {
          const int bx = i & 1, by = i >> 1;
          a[4 + bx] = left_nz[4 + by] = (uint8_t)parse_block(pb, frame_coef, 2, a[4 + bx] + left_nz[4 + by], 0, VP8GPU_BLK_U + i, &t);
        }
        for (int i = 0; i < 4; i++) {
          const int bx = i & 1, by = i >> 1;
          a[6 + bx] = left_nz[6 + by] = (uint8_t)parse_block(pb, frame_coef, 2, a[6 + bx] + left_nz[6 + by], 0, VP8GPU_BLK_V + i, &t);
        }
        mb->tok_cnt = (uint16_t)(t - (out->tokens + n_tok));
        n_tok += mb->tok_cnt;
      }
    }
    d->n_tokens = (uint32_t)n_tok;
    d->n_split = n_split;
  }

  /* the loop-filter "skip inner edges" decision needs tok_cnt, resolved by consumers as
   * (flags & HAS_Y2) && tok_cnt == 0  (macroblock.cc:608) */

  /* commit */
  memcpy(st_io->seg_map, st.seg_map, n_mbs);
  st.seg_map = st_io->seg_map;
  *st_io = st;

done:
  free(seg_map);
  free(info);
  free(above_nz);
  return rc;
}

/* ============================ prediction ============================ */
/* A plane view: pointer, stride = plane width, size. */
typedef struct {
  uint8_t* p;
  int w, h;
} plane;

/* VP8Raster::Block<size>::predictors, prediction.cc:99-167.  col/row are in units of `size`.
 * above points at above[0]; above[-1] .. above[2*size-1] are valid. */
static void predictors(const plane* pl, int size, int col, int row, uint8_t* above, uint8_t* left) {
  const uint8_t* P = pl->p;
  const int W = pl->w;
  if (col > 0)
    for (int i = 0; i < size; i++) left[i] = P[(size * row + i) * W + size * col - 1];
  else
    memset(left, 129, size);
  if (row > 0) memcpy(above, &P[(size * row - 1) * W + size * col], size);
  else memset(above, 127, size);
  if (col > 0 && row > 0) above[-1] = P[(size * row - 1) * W + size * col - 1];
  else if (row > 0) above[-1] = 129;
  else above[-1] = 127;
  if (size != 4) return;
  /* above-right, 4x4 only */
  if (row == 0) {
    memset(above + 4, 127, 4);
  } else if (4 * (col + 1) >= W) {
    if (row >= 4) memset(above + 4, P[(4 * ((row / 4) * 4) - 1) * W + 4 * (col + 1) - 1], 4);
    else memset(above + 4, 127, 4);
  } else if (col % 4 == 3 && row % 4 != 0) {
    if (row >= 4) memcpy(above + 4, &P[(4 * ((row / 4) * 4) - 1) * W + 4 * (col + 1)], 4);
    else memset(above + 4, 127, 4);
  } else {
    memcpy(above + 4, &P[(4 * row - 1) * W + 4 * (col + 1)], 4);
  }
}

/* 16x16 / 8x8 intra modes, prediction.cc:197-209, 241-248, 280-287, 385-467 */
static void intra_predict_mb(const plane* pl, int size, int col, int row, int mode) {
  uint8_t above_store[16 + 32], left[16];
  uint8_t* above = above_store + 16;
  predictors(pl, size, col, row, above, left);
  uint8_t* out = pl->p + (size * row) * pl->w + size * col;
  const int W = pl->w;
  const int log2size = size == 16 ? 4 : 3;
  switch (mode) {
    case VP8GPU_DC_PRED: {
      int value = 128;
      if (col && row) {
        int s = 0;
        for (int i = 0; i < size; i++) s += above[i] + left[i];
        value = (s + (1 << log2size)) >> (log2size + 1);
      } else if (row > 0) {
        int s = 0;
        for (int i = 0; i < size; i++) s += above[i];
        value = (s + (1 << (log2size - 1))) >> log2size;
      } else if (col > 0) {
        int s = 0;
        for (int i = 0; i < size; i++) s += left[i];
        value = (s + (1 << (log2size - 1))) >> log2size;
      }
      for (int y = 0; y < size; y++) memset(out + y * W, value, size);
      break;
    }
    case VP8GPU_V_PRED:
      for (int y = 0; y < size; y++) memcpy(out + y * W, above, size);
      break;
    case VP8GPU_H_PRED:
      for (int y = 0; y < size; y++) memset(out + y * W, left[y], size);
      break;
    default: /* TM_PRED */
      for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++) out[y * W + x] = clamp255(left[y] + above[x] - above[-1]);
  }
}

static inline uint8_t avg3(int x, int y, int z) { return (uint8_t)((x + 2 * y + z + 2) >> 2); }
static inline uint8_t avg2(int x, int y) { return (uint8_t)((x + y + 1) >> 1); }

/* 4x4 intra modes, prediction.cc:469-643.  col/row in 4-pixel units over the plane. */
static void intra_predict_4x4(const plane* pl, int col, int row, int mode) {
  uint8_t above_store[16 + 8], left[4];
  uint8_t* A = above_store + 16;
  predictors(pl, 4, col, row, A, left);
  uint8_t* out = pl->p + (4 * row) * pl->w + 4 * col;
  const int W = pl->w;
#define O(x, y) out[(y)*W + (x)]
  /* east(i): i<=3 ? left[3-i] : above[i-5]  (vp8_raster.hh:80) */
  uint8_t E[9];
  for (int i = 0; i < 9; i++) E[i] = i <= 3 ? left[3 - i] : A[i - 5];
  switch (mode) {
    case VP8GPU_B_DC_PRED: {
      int s = 0;
      for (int i = 0; i < 4; i++) s += A[i] + left[i];
      const int v = (s + 4) >> 3;
      for (int y = 0; y < 4; y++) memset(out + y * W, v, 4);
      break;
    }
    case VP8GPU_B_TM_PRED:
      for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) O(x, y) = clamp255(left[y] + A[x] - A[-1]);
      break;
    case VP8GPU_B_VE_PRED:
      for (int x = 0; x < 4; x++) {
        const uint8_t v = avg3(A[x - 1], A[x], A[x + 1]);
        for (int y = 0; y < 4; y++) O(x, y) = v;
      }
      break;
    case VP8GPU_B_HE_PRED: {
      const uint8_t r0 = avg3(A[-1], left[0], left[1]), r1 = avg3(left[0], left[1], left[2]),
                    r2 = avg3(left[1], left[2], left[3]), r3 = avg3(left[2], left[3], left[3]);
      memset(out, r0, 4);
      memset(out + W, r1, 4);
      memset(out + 2 * W, r2, 4);
      memset(out + 3 * W, r3, 4);
      break;
    }
    case VP8GPU_B_LD_PRED:
      O(0, 0) = avg3(A[0], A[1], A[2]);
      O(1, 0) = O(0, 1) = avg3(A[1], A[2], A[3]);
      O(2, 0) = O(1, 1) = O(0, 2) = avg3(A[2], A[3], A[4]);
      O(3, 0) = O(2, 1) = O(1, 2) = O(0, 3) = avg3(A[3], A[4], A[5]);
      O(3, 1) = O(2, 2) = O(1, 3) = avg3(A[4], A[5], A[6]);
      O(3, 2) = O(2, 3) = avg3(A[5], A[6], A[7]);
      O(3, 3) = avg3(A[6], A[7], A[7]);
      break;
    case VP8GPU_B_RD_PRED:
      O(0, 3) = avg3(E[0], E[1], E[2]);
      O(1, 3) = O(0, 2) = avg3(E[1], E[2], E[3]);
      O(2, 3) = O(1, 2) = O(0, 1) = avg3(E[2], E[3], E[4]);
      O(3, 3) = O(2, 2) = O(1, 1) = O(0, 0) = avg3(E[3], E[4], E[5]);
      O(3, 2) = O(2, 1) = O(1, 0) = avg3(E[4], E[5], E[6]);
      O(3, 1) = O(2, 0) = avg3(E[5], E[6], E[7]);
      O(3, 0) = avg3(E[6], E[7], E[8]);
      break;
    case VP8GPU_B_VR_PRED:
      O(0, 3) = avg3(E[1], E[2], E[3]);
      O(0, 2) = avg3(E[2], E[3], E[4]);
      O(1, 3) = O(0, 1) = avg3(E[3], E[4], E[5]);
      O(1, 2) = O(0, 0) = avg2(E[4], E[5]);
      O(2, 3) = O(1, 1) = avg3(E[4], E[5], E[6]);
      O(2, 2) = O(1, 0) = avg2(E[5], E[6]);
      O(3, 3) = O(2, 1) = avg3(E[5], E[6], E[7]);
      O(3, 2) = O(2, 0) = avg2(E[6], E[7]);
      O(3, 1) = avg3(E[6], E[7], E[8]);
      O(3, 0) = avg2(E[7], E[8]);
      break;
    case VP8GPU_B_VL_PRED:
      O(0, 0) = avg2(A[0], A[1]);
      O(0, 1) = avg3(A[0], A[1], A[2]);
      O(0, 2) = O(1, 0) = avg2(A[1], A[2]);
      O(1, 1) = O(0, 3) = avg3(A[1], A[2], A[3]);
      O(1, 2) = O(2, 0) = avg2(A[2], A[3]);
      O(1, 3) = O(2, 1) = avg3(A[2], A[3], A[4]);
      O(2, 2) = O(3, 0) = avg2(A[3], A[4]);
      O(2, 3) = O(3, 1) = avg3(A[3], A[4], A[5]);
      O(3, 2) = avg3(A[4], A[5], A[6]);
      O(3, 3) = avg3(A[5], A[6], A[7]);
      break;
    case VP8GPU_B_HD_PRED:
      O(0, 3) = avg2(E[0], E[1]);
      O(1, 3) = avg3(E[0], E[1], E[2]);
      O(0, 2) = O(2, 3) = avg2(E[1], E[2]);
      O(1, 2) = O(3, 3) = avg3(E[1], E[2], E[3]);
      O(2, 2) = O(0, 1) = avg2(E[2], E[3]);
      O(3, 2) = O(1, 1) = avg3(E[2], E[3], E[4]);
      O(2, 1) = O(0, 0) = avg2(E[3], E[4]);
      O(3, 1) = O(1, 0) = avg3(E[3], E[4], E[5]);
      O(2, 0) = avg3(E[4], E[5], E[6]);
      O(3, 0) = avg3(E[5], E[6], E[7]);
      break;
    default: /* B_HU_PRED */
      O(0, 0) = avg2(left[0], left[1]);
      O(1, 0) = avg3(left[0], left[1], left[2]);
      O(2, 0) = O(0, 1) = avg2(left[1], left[2]);
      O(3, 0) = O(1, 1) = avg3(left[1], left[2], left[3]);
      O(2, 1) = O(0, 2) = avg2(left[2], left[3]);
      O(3, 1) = O(1, 2) = avg3(left[2], left[3], left[3]);
      O(2, 2) = O(3, 2) = O(0, 3) = O(1, 3) = O(2, 3) = O(3, 3) = left[3];
  }
#undef O
}

/* sixtap_filters, prediction.cc:645-653 */
static const int16_t k_sixtap[8][6] = {{0, 0, 128, 0, 0, 0},     {0, -6, 123, 12, -1, 0},
                                       {2, -11, 108, 36, -8, 1}, {0, -9, 93, 50, -6, 0},
                                       {3, -16, 77, 77, -16, 3}, {0, -6, 50, 93, -9, 0},
                                       {1, -8, 36, 108, -11, 2}, {0, -1, 12, 123, -6, 0}};

/* EdgeExtendedRaster::at, vp8_raster.hh:327-338 */
static inline int ref_at(const plane* r, int x, int y) {
  x = x < 0 ? 0 : (x > r->w - 1 ? r->w - 1 : x);
  y = y < 0 ? 0 : (y > r->h - 1 ? r->h - 1 : y);
  return r->p[y * r->w + x];
}

/* Block<size>::inter_predict (prediction.cc:655-674) via safe_inter_predict (:919-971); the
 * in-bounds fast path (:813-917) computes the same values. col/row in units of `size`. */
static void inter_predict(const plane* out, const plane* ref, int size, int col, int row, int mvx, int mvy) {
  const int sx = col * size + (mvx >> 3), sy = row * size + (mvy >> 3);
  uint8_t* o = out->p + (size * row) * out->w + size * col;
  const int mx = mvx & 7, my = mvy & 7;
  if (mx == 0 && my == 0) {
    for (int y = 0; y < size; y++)
      for (int x = 0; x < size; x++) o[y * out->w + x] = (uint8_t)ref_at(ref, sx + x, sy + y);
    return;
  }
  uint8_t mid[21][16];
  const int16_t* hf = k_sixtap[mx];
  for (int r = 0; r < size + 5; r++)
    for (int c = 0; c < size; c++) {
      const int ry = sy + r - 2, rx = sx + c;
      int s = 64;
      for (int k = 0; k < 6; k++) s += ref_at(ref, rx - 2 + k, ry) * hf[k];
      mid[r][c] = clamp255(s >> 7);
    }
  const int16_t* vf = k_sixtap[my];
  for (int r = 0; r < size; r++)
    for (int c = 0; c < size; c++) {
      int s = 64;
      for (int k = 0; k < 6; k++) s += mid[r + k][c] * vf[k];
      o[r * out->w + c] = clamp255(s >> 7);
    }
}

/* ============================ inverse transforms ============================ */

/* DCTCoefficients::dequantize, quantization.cc:95-126 (int16 wrap) */
static void dequantize(const int16_t in[16], int dc_q, int ac_q, int16_t out[16]) {
  out[0] = (int16_t)(in[0] * dc_q);
  for (int i = 1; i < 16; i++) out[i] = (int16_t)(in[i] * ac_q);
}

/* DCTCoefficients::iwht, transform.cc:47-88: writes coefficient 0 of the 16 Y blocks */
static void iwht(const int16_t c[16], int16_t y[16][16]) {
  int16_t m[16];
  for (int i = 0; i < 4; i++) {
    const int a1 = c[i] + c[i + 12], b1 = c[i + 4] + c[i + 8];
    const int c1 = c[i + 4] - c[i + 8], d1 = c[i] - c[i + 12];
    m[i] = (int16_t)(a1 + b1);
    m[i + 4] = (int16_t)(c1 + d1);
    m[i + 8] = (int16_t)(a1 - b1);
    m[i + 12] = (int16_t)(d1 - c1);
  }
  for (int i = 0; i < 4; i++) {
    const int o = i * 4;
    const int a1 = m[o] + m[o + 3], b1 = m[o + 1] + m[o + 2];
    const int c1 = m[o + 1] - m[o + 2], d1 = m[o] - m[o + 3];
    const int a2 = a1 + b1, b2 = c1 + d1, c2 = a1 - b1, d2 = d1 - c1;
    y[i * 4 + 0][0] = (int16_t)((a2 + 3) >> 3);
    y[i * 4 + 1][0] = (int16_t)((b2 + 3) >> 3);
    y[i * 4 + 2][0] = (int16_t)((c2 + 3) >> 3);
    y[i * 4 + 3][0] = (int16_t)((d2 + 3) >> 3);
  }
}

/* DCTCoefficients::idct_add, transform.cc:100-137 */
static inline int mul_20091(int a) { return ((a * 20091) >> 16) + a; }
static inline int mul_35468(int a) { return (a * 35468) >> 16; }
static void idct_add(const int16_t c[16], uint8_t* out, int stride) {
  int16_t m[16];
  for (int i = 0; i < 4; i++) {
    const int t0 = c[i] + c[i + 8], t1 = c[i] - c[i + 8];
    const int t2 = mul_35468(c[i + 4]) - mul_20091(c[i + 12]);
    const int t3 = mul_20091(c[i + 4]) + mul_35468(c[i + 12]);
    m[i * 4 + 0] = (int16_t)(t0 + t3);
    m[i * 4 + 1] = (int16_t)(t1 + t2);
    m[i * 4 + 2] = (int16_t)(t1 - t2);
    m[i * 4 + 3] = (int16_t)(t0 - t3);
  }
  for (int i = 0; i < 4; i++) {
    const int t0 = m[i] + m[i + 8], t1 = m[i] - m[i + 8];
    const int t2 = mul_35468(m[i + 4]) - mul_20091(m[i + 12]);
    const int t3 = mul_20091(m[i + 4]) + mul_35468(m[i + 12]);
    uint8_t* t = out + i * stride;
    t[0] = clamp255(t[0] + ((t0 + t3 + 4) >> 3));
    t[1] = clamp255(t[1] + ((t1 + t2 + 4) >> 3));
    t[2] = clamp255(t[2] + ((t1 - t2 + 4) >> 3));
    t[3] = clamp255(t[3] + ((t0 - t3 + 4) >> 3));
  }
}

/* ============================ Frame::decode ============================ */
/* frame.cc:208-250, macroblock.cc:504-601 */
void vp8o_reconstruct(const vp8gpu_frame_desc* desc, const vp8gpu_mb* mbs, const vp8gpu_token* tokens,
                      const vp8gpu_split_mvs* split, const vp8o_raster* last, const vp8o_raster* golden,
                      const vp8o_raster* alt, vp8o_raster* out) {
  const int W = out->w16, H = out->h16;
  const plane oy = {out->y, W, H}, ou = {out->u, W / 2, H / 2}, ov = {out->v, W / 2, H / 2};
  for (int row = 0; row < desc->mb_rows; row++) {
    for (int col = 0; col < desc->mb_cols; col++) {
      const vp8gpu_mb* mb = &mbs[(size_t)row * desc->mb_cols + col];
      const vp8gpu_quant* q = &desc->quant[mb->segment_id];
      const int has_y2 = mb->flags & VP8GPU_MB_HAS_Y2;
      const int has_nonzero = mb->tok_cnt != 0;

      /* rebuild the MB's coefficient blocks from its tokens */
      int16_t raw[25][16];
      if (has_nonzero) {
        memset(raw, 0, sizeof(raw));
        for (unsigned t = 0; t < mb->tok_cnt; t++) {
          const uint32_t tok = tokens[mb->tok_off + t];
          raw[(tok >> 20) & 31][(tok >> 16) & 15] = (int16_t)(tok & 0xFFFF);
        }
      }
      int16_t ycoef[16][16], ucoef[4][16], vcoef[4][16];
      if (has_nonzero) {
        for (int i = 0; i < 16; i++) dequantize(raw[i], q->y_dc, q->y_ac, ycoef[i]);
        for (int i = 0; i < 4; i++) dequantize(raw[VP8GPU_BLK_U + i], q->uv_dc, q->uv_ac, ucoef[i]);
        for (int i = 0; i < 4; i++) dequantize(raw[VP8GPU_BLK_V + i], q->uv_dc, q->uv_ac, vcoef[i]);
        if (has_y2) { /* apply_walsh, macroblock.cc:504-521 */
          int16_t y2[16];
          dequantize(raw[VP8GPU_BLK_Y2], q->y2_dc, q->y2_ac, y2);
          iwht(y2, ycoef);
        }
      }

      if (mb->ref_frame == VP8GPU_REF_CURRENT) {
        /* reconstruct_intra, macroblock.cc:523-551 */
        intra_predict_mb(&ou, 8, col, row, mb->uv_mode);
        intra_predict_mb(&ov, 8, col, row, mb->uv_mode);
        if (has_nonzero)
          for (int i = 0; i < 4; i++) {
            const int bx = i & 1, by = i >> 1;
            idct_add(ucoef[i], ou.p + (8 * row + 4 * by) * ou.w + 8 * col + 4 * bx, ou.w);
            idct_add(vcoef[i], ov.p + (8 * row + 4 * by) * ov.w + 8 * col + 4 * bx, ov.w);
          }
        if (mb->y_mode == VP8GPU_B_PRED) {
          for (int i = 0; i < 16; i++) {
            const int bx = i & 3, by = i >> 2;
            intra_predict_4x4(&oy, 4 * col + bx, 4 * row + by, (int)((mb->b_modes >> (4 * i)) & 15));
            if (has_nonzero) idct_add(ycoef[i], oy.p + (16 * row + 4 * by) * W + 16 * col + 4 * bx, W);
          }
        } else {
          intra_predict_mb(&oy, 16, col, row, mb->y_mode);
          if (has_nonzero)
            for (int i = 0; i < 16; i++) {
              const int bx = i & 3, by = i >> 2;
              idct_add(ycoef[i], oy.p + (16 * row + 4 * by) * W + 16 * col + 4 * bx, W);
            }
        }
      } else {
        /* reconstruct_inter, macroblock.cc:553-601 */
        const vp8o_raster* r = mb->ref_frame == VP8GPU_REF_LAST ? last : (mb->ref_frame == VP8GPU_REF_GOLDEN ? golden : alt);
        const plane ry = {r->y, W, H}, ru = {r->u, W / 2, H / 2}, rv = {r->v, W / 2, H / 2};
        if (mb->y_mode == VP8GPU_SPLITMV) {
          const int16_t(*mv)[2] = split[mb->split_idx].mv;
          for (int i = 0; i < 16; i++) inter_predict(&oy, &ry, 4, 4 * col + (i & 3), 4 * row + (i >> 2), mv[i][0], mv[i][1]);
          for (int i = 0; i < 4; i++) {
            const mvec cm = chroma_mv(mv, i & 1, i >> 1);
            inter_predict(&ou, &ru, 4, 2 * col + (i & 1), 2 * row + (i >> 1), cm.x, cm.y);
            inter_predict(&ov, &rv, 4, 2 * col + (i & 1), 2 * row + (i >> 1), cm.x, cm.y);
          }
        } else {
          int16_t mv[16][2];
          for (int i = 0; i < 16; i++) {
            mv[i][0] = mb->mv_x;
            mv[i][1] = mb->mv_y;
          }
          const mvec cm = chroma_mv(mv, 0, 0);
          inter_predict(&oy, &ry, 16, col, row, mb->mv_x, mb->mv_y);
          inter_predict(&ou, &ru, 8, col, row, cm.x, cm.y);
          inter_predict(&ov, &rv, 8, col, row, cm.x, cm.y);
        }
        if (has_nonzero) {
          for (int i = 0; i < 16; i++) {
            const int bx = i & 3, by = i >> 2;
            idct_add(ycoef[i], oy.p + (16 * row + 4 * by) * W + 16 * col + 4 * bx, W);
          }
          for (int i = 0; i < 4; i++) {
            const int bx = i & 1, by = i >> 1;
            idct_add(ucoef[i], ou.p + (8 * row + 4 * by) * ou.w + 8 * col + 4 * bx, ou.w);
            idct_add(vcoef[i], ov.p + (8 * row + 4 * by) * ov.w + 8 * col + 4 * bx, ov.w);
          }
        }
      }
    }
  }
}

/* ============================ loop filter ============================ */
/* loopfilter_filters.hh:50-183 (libvpx normal filter) */
static inline int8_t sclamp(int t) { return (int8_t)(t < -128 ? -128 : (t > 127 ? 127 : t)); }
static inline int iabs(int a) { return a < 0 ? -a : a; }

static inline int8_t filter_mask(int limit, int blimit, int p3, int p2, int p1, int p0, int q0, int q1, int q2, int q3) {
  int mask = 0;
  mask |= (iabs(p3 - p2) > limit);
  mask |= (iabs(p2 - p1) > limit);
  mask |= (iabs(p1 - p0) > limit);
  mask |= (iabs(q1 - q0) > limit);
  mask |= (iabs(q2 - q1) > limit);
  mask |= (iabs(q3 - q2) > limit);
  mask |= (iabs(p0 - q0) * 2 + iabs(p1 - q1) / 2 > blimit);
  return (int8_t)(mask - 1);
}
static inline int8_t hev_mask(int thresh, int p1, int p0, int q0, int q1) {
  int hev = 0;
  hev |= (iabs(p1 - p0) > thresh) * -1;
  hev |= (iabs(q1 - q0) > thresh) * -1;
  return (int8_t)hev;
}
/* vp8_filter: inner-edge filter */
static void lf_inner(int8_t mask, int8_t hev, uint8_t* op1, uint8_t* op0, uint8_t* oq0, uint8_t* oq1) {
  const int8_t ps1 = (int8_t)(*op1 ^ 0x80), ps0 = (int8_t)(*op0 ^ 0x80);
  const int8_t qs0 = (int8_t)(*oq0 ^ 0x80), qs1 = (int8_t)(*oq1 ^ 0x80);
  int8_t f = sclamp(ps1 - qs1);
  f &= hev;
  f = sclamp(f + 3 * (qs0 - ps0));
  f &= mask;
  int8_t f1 = sclamp(f + 4), f2 = sclamp(f + 3);
  f1 >>= 3;
  f2 >>= 3;
  *oq0 = (uint8_t)(sclamp(qs0 - f1) ^ 0x80);
  *op0 = (uint8_t)(sclamp(ps0 + f2) ^ 0x80);
  f = f1;
  f += 1;
  f >>= 1;
  f &= ~hev;
  *oq1 = (uint8_t)(sclamp(qs1 - f) ^ 0x80);
  *op1 = (uint8_t)(sclamp(ps1 + f) ^ 0x80);
}
/* vp8_mbfilter: macroblock-edge filter */
static void lf_mbedge(int8_t mask, int8_t hev, uint8_t* op2, uint8_t* op1, uint8_t* op0, uint8_t* oq0, uint8_t* oq1, uint8_t* oq2) {
  const int8_t ps2 = (int8_t)(*op2 ^ 0x80), ps1 = (int8_t)(*op1 ^ 0x80);
  int8_t ps0 = (int8_t)(*op0 ^ 0x80), qs0 = (int8_t)(*oq0 ^ 0x80);
  const int8_t qs1 = (int8_t)(*oq1 ^ 0x80), qs2 = (int8_t)(*oq2 ^ 0x80);
  int8_t f = sclamp(ps1 - qs1);
  f = sclamp(f + 3 * (qs0 - ps0));
  f &= mask;
  int8_t f2 = f;
  f2 &= hev;
  int8_t f1 = sclamp(f2 + 4);
  f2 = sclamp(f2 + 3);
  f1 >>= 3;
  f2 >>= 3;
  qs0 = sclamp(qs0 - f1);
  ps0 = sclamp(ps0 + f2);
  f &= ~hev;
  f2 = f;
  int8_t u = sclamp((63 + f2 * 27) >> 7);
  *oq0 = (uint8_t)(sclamp(qs0 - u) ^ 0x80);
  *op0 = (uint8_t)(sclamp(ps0 + u) ^ 0x80);
  u = sclamp((63 + f2 * 18) >> 7);
  *oq1 = (uint8_t)(sclamp(qs1 - u) ^ 0x80);
  *op1 = (uint8_t)(sclamp(ps1 + u) ^ 0x80);
  u = sclamp((63 + f2 * 9) >> 7);
  *oq2 = (uint8_t)(sclamp(qs2 - u) ^ 0x80);
  *op2 = (uint8_t)(sclamp(ps2 + u) ^ 0x80);
}

typedef struct { int interior, mb_edge, sub_edge, hev; } lf_params;

/* filter one edge of `n` positions: p points at q0 of position 0, `step` moves across the
 * edge (towards q1), `next` moves to the next position along the edge */
static void edge(uint8_t* p, int step, int next, int n, const lf_params* lp, int mb_edge) {
  for (int i = 0; i < n; i++, p += next) {
    const int8_t mask = filter_mask(lp->interior, mb_edge ? lp->mb_edge : lp->sub_edge, p[-4 * step], p[-3 * step],
                                    p[-2 * step], p[-step], p[0], p[step], p[2 * step], p[3 * step]);
    const int8_t hev = hev_mask(lp->hev, p[-2 * step], p[-step], p[0], p[step]);
    if (mb_edge) lf_mbedge(mask, hev, p - 3 * step, p - 2 * step, p - step, p, p + step, p + 2 * step);
    else lf_inner(mask, hev, p - 2 * step, p - step, p, p + step);
  }
}

/* Frame::loopfilter (frame.cc:139-182) -> Macroblock::loopfilter (macroblock.cc:603-641) ->
 * NormalLoopFilter (loopfilter.cc:81-154) */
void vp8o_loopfilter(const vp8gpu_frame_desc* desc, const vp8gpu_mb* mbs, vp8o_raster* out) {
  if (!desc->loop_filter_level) return;
  const int W = out->w16, CW = W / 2;
  for (int row = 0; row < desc->mb_rows; row++) {
    for (int col = 0; col < desc->mb_cols; col++) {
      const vp8gpu_mb* mb = &mbs[(size_t)row * desc->mb_cols + col];
      const int level = mb->lf_level;
      if (!level) continue;
      const int skip_inner = (mb->flags & VP8GPU_MB_HAS_Y2) && mb->tok_cnt == 0;
      lf_params lp;
      int interior = level; /* SimpleLoopFilter ctor, loopfilter.cc:81-104 */
      if (desc->sharpness) {
        interior >>= desc->sharpness > 4 ? 2 : 1;
        if (interior > 9 - desc->sharpness) interior = 9 - desc->sharpness;
      }
      if (interior < 1) interior = 1;
      lp.interior = interior;
      lp.mb_edge = ((level + 2) * 2) + interior;
      lp.sub_edge = (level * 2) + interior;
      lp.hev = (level >= 15) + (level >= 40) + (level >= 20 && !desc->key_frame); /* loopfilter.cc:113-122 */
      uint8_t* y = out->y + (size_t)16 * row * W + 16 * col;
      uint8_t* u = out->u + (size_t)8 * row * CW + 8 * col;
      uint8_t* v = out->v + (size_t)8 * row * CW + 8 * col;
      if (col > 0) { /* 1: left macroblock edge */
        edge(y, 1, W, 16, &lp, 1);
        edge(u, 1, CW, 8, &lp, 1);
        edge(v, 1, CW, 8, &lp, 1);
      }
      if (!skip_inner) { /* 2: vertical sub-block edges */
        for (int x = 4; x < 16; x += 4) edge(y + x, 1, W, 16, &lp, 0);
        edge(u + 4, 1, CW, 8, &lp, 0);
        edge(v + 4, 1, CW, 8, &lp, 0);
      }
      if (row > 0) { /* 3: top macroblock edge */
        edge(y, W, 1, 16, &lp, 1);
        edge(u, CW, 1, 8, &lp, 1);
        edge(v, CW, 1, 8, &lp, 1);
      }
      if (!skip_inner) { /* 4: horizontal sub-block edges */
        for (int r = 4; r < 16; r += 4) edge(y + r * W, W, 1, 16, &lp, 0);
        edge(u + 4 * CW, CW, 1, 8, &lp, 0);
        edge(v + 4 * CW, CW, 1, 8, &lp, 0);
      }
    }
  }
}

/* ============================ Decoder ============================ */
struct vp8o_decoder {
  int width, height;
  vp8o_state* state;
  vp8o_parsed* parsed;
  vp8o_raster* pool[4];
  int refcnt[4];
  int last, golden, alt; /* indices into pool: References, decoder.hh:123-149 */
};

vp8o_decoder* vp8o_decoder_new(int width, int height) {
  vp8o_decoder* d = (vp8o_decoder*)calloc(1, sizeof(*d));
  d->width = width;
  d->height = height;
  d->state = vp8o_state_new(width, height);
  d->parsed = vp8o_parsed_new();
  for (int i = 0; i < 4; i++) d->pool[i] = vp8o_raster_new(width, height);
  d->last = d->golden = d->alt = 0; /* References( MutableRasterHandle ): all three share one raster */
  d->refcnt[0] = 3;
  return d;
}
void vp8o_decoder_free(vp8o_decoder* d) {
  if (!d) return;
  vp8o_state_free(d->state);
  vp8o_parsed_free(d->parsed);
  for (int i = 0; i < 4; i++) vp8o_raster_free(d->pool[i]);
  free(d);
}
static void set_ref(vp8o_decoder* d, int* slot, int idx) {
  d->refcnt[idx]++;
  d->refcnt[*slot]--;
  *slot = idx;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static int decode_timed(vp8o_decoder* d, const uint8_t* data, size_t len, int* shown, const vp8o_raster** out,
                        vp8o_raster* pre_lf, double phase[3]) {
  const double t0 = now_s();
  const int rc = vp8o_parse_frame(d->state, data, len, d->parsed);
  if (rc != VP8GPU_OK) return rc;
  const double t1 = now_s();
  const vp8gpu_frame_desc* desc = &d->parsed->desc;
  int cur = 0;
  while (d->refcnt[cur]) cur++;
  vp8o_raster* r = d->pool[cur];
  /* Decoder::decode_frame, decoder.cc:101-118 */
  vp8o_reconstruct(desc, d->parsed->mbs, d->parsed->tokens, d->parsed->split, d->pool[d->last], d->pool[d->golden],
                   d->pool[d->alt], r);
  const double t2 = now_s();
  if (pre_lf) vp8o_raster_copy(pre_lf, r);
  const double t3 = now_s();
  vp8o_loopfilter(desc, d->parsed->mbs, r);
  const double t4 = now_s();
  /* Frame::copy_to, frame.cc:272-307 -- note the order: alternate, golden, then refreshes */
  d->refcnt[cur]++; /* the output handle */
  if (desc->key_frame) {
    set_ref(d, &d->last, cur);
    set_ref(d, &d->golden, cur);
    set_ref(d, &d->alt, cur);
  } else {
    if (desc->copy_to_alternate == 1) set_ref(d, &d->alt, d->last);
    else if (desc->copy_to_alternate == 2) set_ref(d, &d->alt, d->golden);
    if (desc->copy_to_golden == 1) set_ref(d, &d->golden, d->last);
    else if (desc->copy_to_golden == 2) set_ref(d, &d->golden, d->alt);
    if (desc->refresh_golden) set_ref(d, &d->golden, cur);
    if (desc->refresh_alternate) set_ref(d, &d->alt, cur);
    if (desc->refresh_last) set_ref(d, &d->last, cur);
  }
  d->refcnt[cur]--; /* caller only borrows it until the next call */
  if (shown) *shown = desc->show_frame;
  if (out) *out = r;
  if (phase) {
    phase[0] += t1 - t0;
    phase[1] += t2 - t1;
    phase[2] += t4 - t3;
  }
  return VP8GPU_OK;
}

int vp8o_decoder_decode(vp8o_decoder* d, const uint8_t* data, size_t len, int* shown, const vp8o_raster** out,
                        vp8o_raster* pre_lf) {
  return decode_timed(d, data, len, shown, out, pre_lf, NULL);
}
const vp8o_parsed* vp8o_decoder_last_parsed(const vp8o_decoder* d) { return d->parsed; }
const vp8o_raster* vp8o_decoder_ref(const vp8o_decoder* d, int which) {
  return d->pool[which == 0 ? d->last : (which == 1 ? d->golden : d->alt)];
}

/* IVF walk (util/ivf.cc:36-82) + FilePlayer start rule (player.cc:101-109) */
int vp8o_time_ivf(const uint8_t* ivf, size_t len, int reps, uint32_t max_frames, double phase_s[3], uint32_t* frames) {
  if (len < 32 || memcmp(ivf, "DKIF", 4)) return VP8GPU_ERR_INVALID;
  const int w = ivf[12] | (ivf[13] << 8), h = ivf[14] | (ivf[15] << 8);
  const uint32_t count = ivf[24] | (ivf[25] << 8) | (ivf[26] << 16) | ((uint32_t)ivf[27] << 24);
  double best[3] = {0, 0, 0}, best_total = 1e30;
  for (int rep = 0; rep < reps; rep++) {
    vp8o_decoder* d = vp8o_decoder_new(w, h);
    double ph[3] = {0, 0, 0};
    size_t pos = 32;
    uint32_t done = 0;
    int started = 0;
    for (uint32_t i = 0; i < count && done < max_frames; i++) {
      if (pos + 12 > len) break;
      const uint32_t flen = ivf[pos] | (ivf[pos + 1] << 8) | (ivf[pos + 2] << 16) | ((uint32_t)ivf[pos + 3] << 24);
      const uint8_t* f = ivf + pos + 12;
      pos += 12 + flen;
      if (pos > len) break;
      if (!started && (flen < 1 || (f[0] & 1))) continue;
      started = 1;
      const int rc = decode_timed(d, f, flen, NULL, NULL, NULL, ph);
      if (rc != VP8GPU_OK) {
        vp8o_decoder_free(d);
        return rc;
      }
      done++;
    }
    vp8o_decoder_free(d);
    const double total = ph[0] + ph[1] + ph[2];
    if (total < best_total) {
      best_total = total;
      memcpy(best, ph, sizeof(best));
    }
    *frames = done;
  }
  memcpy(phase_s, best, sizeof(best));
  return VP8GPU_OK;
}

/* ============================ unit-level test hooks ============================ */
/* (used by tests/test_math_host.py to check the arithmetic of csrc/vp8_math.cuh on the CPU) */
void vp8o_test_idct_add(const int16_t c[16], uint8_t px[16]) { idct_add(c, px, 4); }
void vp8o_test_iwht(const int16_t c[16], int16_t dc[16]) {
  int16_t y[16][16];
  memset(y, 0, sizeof(y));
  iwht(c, y);
  for (int i = 0; i < 16; i++) dc[i] = y[i][0];
}
void vp8o_test_lf_edge(uint8_t px[8], int level, int sharpness, int key_frame, int mb_edge) {
  lf_params lp;
  int interior = level;
  if (sharpness) {
    interior >>= sharpness > 4 ? 2 : 1;
    if (interior > 9 - sharpness) interior = 9 - sharpness;
  }
  if (interior < 1) interior = 1;
  lp.interior = interior;
  lp.mb_edge = ((level + 2) * 2) + interior;
  lp.sub_edge = (level * 2) + interior;
  lp.hev = (level >= 15) + (level >= 40) + (level >= 20 && !key_frame);
  edge(px + 4, 1, 8, 1, &lp, mb_edge);
}
/* s = edge vector: s[0..3] = left[3..0], s[4] = above[-1], s[5..12] = above[0..7] */
void vp8o_test_bpred(int mode, const uint8_t s[13], uint8_t out[16]) {
  uint8_t buf[16 * 16];
  memset(buf, 0, sizeof(buf));
  plane pl = {buf, 16, 16};
  /* place the 4x4 block at sub-block (1,1) of a 16x16 plane so every neighbour is a real pixel */
  for (int k = 0; k < 4; k++) buf[(4 + k) * 16 + 3] = s[3 - k];
  for (int k = -1; k < 8; k++) buf[3 * 16 + 4 + k] = s[5 + k];
  intra_predict_4x4(&pl, 1, 1, mode);
  for (int y = 0; y < 4; y++)
    for (int x = 0; x < 4; x++) out[y * 4 + x] = buf[(4 + y) * 16 + 4 + x];
}
/* N x N six-tap prediction from a (N+5)^2 window (window[2][2] = block origin) */
void vp8o_test_sixtap(const uint8_t* window, int n, int mx, int my, uint8_t* out) {
  const int ws = n + 5;
  plane ref = {(uint8_t*)window, ws, ws};
  uint8_t* tmp = (uint8_t*)calloc((size_t)ws * ws, 1);
  plane o = {tmp, ws, ws};
  /* block (col,row) = (0,0) of size n shifted by mv = (2*8 + mx, 2*8 + my) reads window[2+..] */
  inter_predict(&o, &ref, n, 0, 0, 16 + mx, 16 + my);
  for (int y = 0; y < n; y++) memcpy(out + y * n, tmp + y * ws, n);
  free(tmp);
}

/* ---- forward transforms of the encoder (decoder/dct.cc:45-164), restated for unit tests ---- */
void vp8o_test_fdct(const uint8_t src[16], const uint8_t pred[16], int16_t out[16]) {
  int16_t in[16], c[16];
  for (int i = 0; i < 16; i++) in[i] = (int16_t)(src[i] - pred[i]);
  for (int i = 0; i < 4; i++) { /* rows */
    const int a1 = (in[4 * i + 0] + in[4 * i + 3]) * 8, b1 = (in[4 * i + 1] + in[4 * i + 2]) * 8;
    const int c1 = (in[4 * i + 1] - in[4 * i + 2]) * 8, d1 = (in[4 * i + 0] - in[4 * i + 3]) * 8;
    c[4 * i + 0] = (int16_t)(a1 + b1);
    c[4 * i + 2] = (int16_t)(a1 - b1);
    c[4 * i + 1] = (int16_t)((c1 * 2217 + d1 * 5352 + 14500) >> 12);
    c[4 * i + 3] = (int16_t)((d1 * 2217 - c1 * 5352 + 7500) >> 12);
  }
  for (int i = 0; i < 4; i++) { /* columns */
    const int a1 = c[i + 0] + c[i + 12], b1 = c[i + 4] + c[i + 8];
    const int c1 = c[i + 4] - c[i + 8], d1 = c[i + 0] - c[i + 12];
    out[i + 0] = (int16_t)((a1 + b1 + 7) >> 4);
    out[i + 8] = (int16_t)((a1 - b1 + 7) >> 4);
    out[i + 4] = (int16_t)(((c1 * 2217 + d1 * 5352 + 12000) >> 16) + (d1 != 0));
    out[i + 12] = (int16_t)((d1 * 2217 - c1 * 5352 + 51000) >> 16);
  }
}
void vp8o_test_fwht(const int16_t in[16], int16_t out[16]) {
  int16_t c[16];
  for (int i = 0; i < 4; i++) {
    const int a1 = (in[4 * i + 0] + in[4 * i + 2]) * 4, d1 = (in[4 * i + 1] + in[4 * i + 3]) * 4;
    const int c1 = (in[4 * i + 1] - in[4 * i + 3]) * 4, b1 = (in[4 * i + 0] - in[4 * i + 2]) * 4;
    c[4 * i + 0] = (int16_t)(a1 + d1 + (a1 != 0));
    c[4 * i + 1] = (int16_t)(b1 + c1);
    c[4 * i + 2] = (int16_t)(b1 - c1);
    c[4 * i + 3] = (int16_t)(a1 - d1);
  }
  for (int i = 0; i < 4; i++) {
    const int a1 = c[i + 0] + c[i + 8], d1 = c[i + 4] + c[i + 12];
    const int c1 = c[i + 4] - c[i + 12], b1 = c[i + 0] - c[i + 8];
    int a2 = a1 + d1, b2 = b1 + c1, c2 = b1 - c1, d2 = a1 - d1;
    a2 += a2 < 0;
    b2 += b2 < 0;
    c2 += c2 < 0;
    d2 += d2 < 0;
    out[i + 0] = (int16_t)((a2 + 3) >> 3);
    out[i + 4] = (int16_t)((b2 + 3) >> 3);
    out[i + 8] = (int16_t)((c2 + 3) >> 3);
    out[i + 12] = (int16_t)((d2 + 3) >> 3);
  }
}
