// enc_costs.cc -- see enc_costs.h.  Restates Costs::fill_mode_costs / fill_mv_ref_costs /
// fill_mv_component_costs / fill_mv_sad_costs (encoder/costs.cc:64-221) over flat tables.
#include "enc_costs.h"

#include <math.h>
#include <string.h>

#include "vp8_enc_tables.h"
#include "vp8_tables.h"

namespace vp8 {
namespace {

inline uint16_t cost_zero(uint8_t p) { return k_prob_cost[p]; }
inline uint16_t cost_one(uint8_t p) { return k_prob_cost[255 - p]; }  // complement() = 255 - p (costs.cc:53)
inline uint16_t cost_bit(uint8_t p, bool b) { return b ? cost_one(p) : cost_zero(p); }

// Costs::compute_cost (costs.cc:145-168): cost of every leaf of a token tree; leaves are <= 0 (value = -entry)
void tree_costs(uint16_t* out, const uint8_t* probs, const int8_t* tree, int index, uint16_t cost) {
  const uint8_t p = probs[index / 2];
  for (int i = 0; i < 2; i++) {
    const int entry = tree[index + i];
    const uint16_t c = (uint16_t)(cost + cost_bit(p, i != 0));
    if (entry <= 0) out[-entry] = c;
    else tree_costs(out, probs, tree, entry, c);
  }
}

// modemv_data.cc:162-215
const int8_t kKfYModeTree[8] = {-4, 2, 4, 6, 0, -1, -2, -3};
const int8_t kYModeTree[8] = {0, 2, 4, 6, -1, -2, -3, -4};
const int8_t kBModeTree[18] = {0, 2, -1, 4, -2, 6, 8, 12, -3, 10, -5, -6, -4, 14, -7, 16, -8, -9};
const int8_t kSmallMvTree[14] = {2, 8, 4, 6, 0, -1, -2, -3, 10, 12, -4, -5, -6, -7};

enum { IS_SHORT = 0, SIGN = 1, SHORT = 2, BITS = SHORT + 8 - 1, LONG_MV_WIDTH = 10 };

// Costs::mv_component_cost (costs.cc:64-100) without the sign
uint32_t mv_magnitude_cost(int num, const uint8_t* probs) {
  const int x = num >> 1;
  uint32_t cost;
  if (x < 8) {
    cost = cost_zero(probs[IS_SHORT]);
    int idx = 0;
    for (int n = 2; n >= 0; n--) {  // tree_cost( x, 3, small_mv_tree, probs + SHORT )
      const int bit = (x >> n) & 1;
      cost += cost_bit(probs[SHORT + idx / 2], bit);
      idx = kSmallMvTree[idx + bit];
    }
  } else {
    cost = cost_one(probs[IS_SHORT]);
    for (int i = 0; i < 3; i++) cost += cost_bit(probs[BITS + i], (x >> i) & 1);
    for (int i = LONG_MV_WIDTH - 1; i > 3; i--) cost += cost_bit(probs[BITS + i], (x >> i) & 1);
    if (x & 0xfff0) cost += cost_bit(probs[BITS + 3], (x >> 3) & 1);
  }
  return cost;
}

}  // namespace

void build_enc_tables(EncTables& t, const uint8_t* mv_probs) {
  if (!mv_probs) mv_probs = k_mv_default_probs;
  memset(&t, 0, sizeof(t));
  for (int a = 0; a < 10; a++)
    for (int l = 0; l < 10; l++) tree_costs(t.bmode_cost[a][l], k_kf_bmode_probs + (a * 10 + l) * 9, kBModeTree, 0, 0);
  tree_costs(t.ymode_cost[0], k_kf_ymode_probs, kKfYModeTree, 0, 0);
  tree_costs(t.ymode_cost[1], k_ymode_default_probs, kYModeTree, 0, 0);
  for (int k = 0; k < 4; k++)
    for (int c = 0; c < 6; c++) {
      t.mvref_zero[k][c] = cost_zero(k_mv_count_probs[c * 4 + k]);
      t.mvref_one[k][c] = cost_one(k_mv_count_probs[c * 4 + k]);
    }
  // Costs::fill_mv_component_costs( decoder_state_.probability_tables.motion_vector_probs ) (encode_inter.cc:601,
  // reencode.cc:85).  The encoder never updates these probabilities itself (optimize_mv_probs is not called from
  // encode_raster, encode_inter.cc:578-653): they are the defaults unless the Encoder was built from a Decoder that
  // had seen updates.
  for (int comp = 0; comp < 2; comp++) {
    const uint8_t* probs = mv_probs + comp * 19;
    for (int i = 0; i < 1024; i++) t.mv_mag_cost[comp][i] = (uint16_t)mv_magnitude_cost(i, probs);
    t.mv_sign_cost[comp][0] = cost_zero(probs[SIGN]);
    t.mv_sign_cost[comp][1] = cost_one(probs[SIGN]);
  }
  t.mv_sad_cost[0] = 300;
  for (size_t i = 1; i <= 255; i++) {
    const size_t cost = 256 * (2 * log2f(8 * i) + 0.6);  // the reference's expression, float then double (costs.cc:136)
    t.mv_sad_cost[i] = (uint16_t)cost;
  }
}

// ---- two-pass key frames: token costs and the cost of a coefficient value beyond its token ----
namespace {
// tokens.hh:36-49: ZERO ONE TWO THREE FOUR CAT1..CAT6 EOB = 0..11; costs.cc:38-52
const int8_t kCoefTree[22] = {-11, 2, 0, 4, -1, 6, 8, 12, -2, 10, -3, -4, 14, 16, -5, -6, 18, 20, -7, -8, -9, -10};
// extra bits of DCT_VAL_CATEGORY1..6: first value, number of bits, their probabilities (tokens.hh:62-78)
const int kCatBase[6] = {5, 7, 11, 19, 35, 67};
const int kCatBits[6] = {1, 2, 3, 4, 5, 11};
const uint8_t kCatProbs[6][11] = {{159}, {165, 145}, {173, 148, 140}, {176, 155, 140, 135}, {180, 157, 141, 134, 130},
                                  {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129}};
}  // namespace

void build_trellis_tables(TrellisTables& t) {
  memset(&t, 0, sizeof(t));
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++)
      for (int k = 0; k < 3; k++) {
        const uint8_t* probs = k_coef_default_probs + ((i * 8 + j) * 3 + k) * 11;
        // after a zero (context 0, beyond the first position of the block type) a block cannot end: the tree is
        // entered below its end-of-block decision (costs.cc:180-185)
        tree_costs(t.token_cost[i][j][k], probs, kCoefTree, (k == 0 && j > (i == 0 ? 1 : 0)) ? 2 : 0, 0);
      }
  // dct_value_cost (libvpx tokenize.c fill_value_tokens): the extra bits of the value's category, most significant
  // first, each with its own probability, plus the sign at probability one half; nothing for 0
  for (int v = -2048; v < 2048; v++) {
    const int a = v < 0 ? -v : v;
    if (a == 0) continue;
    uint32_t cost = cost_bit(128, v < 0);
    if (a > 4) {
      int cat = 0;
      while (cat + 1 < 6 && kCatBase[cat + 1] <= a) cat++;
      const int extra = a - kCatBase[cat];
      for (int b = 0; b < kCatBits[cat]; b++) cost += cost_bit(kCatProbs[cat][b], (extra >> (kCatBits[cat] - 1 - b)) & 1);
    }
    t.value_cost[v + 2048] = (uint16_t)cost;
  }
}

}  // namespace vp8
