// enc_costs.h -- rate tables of the encoder's mode decision (encoder/costs.cc), built once on the host
// and kept in HBM for k_enc_rd.  All entries are the reference's integers (costs in 1/256 bit).
#pragma once
#include <stdint.h>

namespace vp8 {

struct EncTables {
  uint16_t bmode_cost[10][10][10];  // [above mode][left mode][mode]      Costs::bmode_costs (costs.cc:196-203)
  uint16_t ymode_cost[2][5];        // [0 key / 1 inter frame][DC V H TM B] Costs::mbmode_costs (costs.cc:205-207)
  uint16_t mvref_zero[4][6];        // cost of a 0 at node k of mv_ref_tree given census count c: mv_counts_to_probs[c][k]
  uint16_t mvref_one[4][6];         //         of a 1                       (Costs::fill_mv_ref_costs, costs.cc:218-221)
  uint16_t mv_mag_cost[2][1024];    // [0 row (y) / 1 column (x)][|component|] without the sign (costs.cc:64-124)
  uint16_t mv_sign_cost[2][2];      // [component][negative]
  uint16_t mv_sad_cost[256];        // Costs::fill_mv_sad_costs (costs.cc:127-142)
  uint16_t pad[2];
};

// Tables of the second (trellis) pass of a two-pass key frame (encoder/encoder.cc:220-408)
struct TrellisTables {
  uint16_t token_cost[4][8][3][12];  // [block type][band][context][token]  Costs::fill_token_costs of the DEFAULT
                                     // coefficient probabilities (encode_intra.cc:413-414: decoder_state_ is a fresh one)
  uint16_t value_cost[4096];         // Costs::coeff_base_cost( v ) at [v + 2048]: extra bits + sign of the token of v
};
void build_trellis_tables(TrellisTables& t);

// mv_probs: the stream's saved motion-vector probabilities ([2][19]); nullptr = the default table
void build_enc_tables(EncTables& t, const uint8_t* mv_probs = nullptr);

// Encoder::update_rd_multipliers (encoder.cc:179-194)
inline void rd_multipliers(int y_ac, uint32_t* rate_mult, uint32_t* dist_mult) {
  const double q_ac = y_ac < 160 ? y_ac : 160.0;
  uint32_t rm = (uint32_t)(q_ac * q_ac * 2.80);
  if (rm > 1000) {
    *dist_mult = 1;
    rm /= 100;
  } else {
    *dist_mult = 100;
  }
  *rate_mult = rm;
}

}  // namespace vp8
