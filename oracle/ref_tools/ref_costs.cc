/* ref_costs -- TEST INFRASTRUCTURE.  Prints the rate tables of the UNMODIFIED reference encoder
 * (encoder/costs.cc: Costs::fill_mode_costs, fill_mv_ref_costs, fill_mv_component_costs, fill_mv_sad_costs)
 * as flat integers, in the layout of alfalfa_b200/csrc/enc_costs.h, for tests/test_enc_costs_host.py. */
#include <cstdio>

#include "costs.hh"
#include "modemv_data.hh"
#include "decoder.hh"

int main() {
  Costs c;
  c.fill_mode_costs();
  ProbabilityTables pt;
  c.fill_mv_component_costs(pt.motion_vector_probs);
  c.fill_mv_sad_costs();
  printf("bmode");
  for (unsigned a = 0; a < 10; a++)
    for (unsigned l = 0; l < 10; l++)
      for (unsigned m = 0; m < 10; m++) printf(" %u", (unsigned)c.bmode_costs.at(a).at(l).at(m));
  printf("\nymode");
  for (unsigned k = 0; k < 2; k++)
    for (unsigned m = 0; m < 5; m++) printf(" %u", (unsigned)c.mbmode_costs.at(k).at(m));
  /* mv_ref costs for every census count vector with counts 0..5 and split count 0: ZERO NEAREST NEAR NEW */
  printf("\nmvref");
  for (unsigned c0 = 0; c0 < 6; c0++)
    for (unsigned c1 = 0; c1 < 6; c1++)
      for (unsigned c2 = 0; c2 < 6; c2++) {
        const ProbabilityArray<num_mv_refs> probs = {{mv_counts_to_probs.at(c0).at(0), mv_counts_to_probs.at(c1).at(1),
                                                      mv_counts_to_probs.at(c2).at(2), mv_counts_to_probs.at(0).at(3)}};
        c.fill_mv_ref_costs(probs);
        printf(" %u %u %u %u", (unsigned)c.mbmode_costs.at(1).at(ZEROMV), (unsigned)c.mbmode_costs.at(1).at(NEARESTMV),
               (unsigned)c.mbmode_costs.at(1).at(NEARMV), (unsigned)c.mbmode_costs.at(1).at(NEWMV));
      }
  printf("\nmvcomp");
  for (unsigned comp = 0; comp < 2; comp++)
    for (unsigned sign = 0; sign < 2; sign++)
      for (unsigned i = 0; i < 1024; i++) printf(" %u", (unsigned)c.mv_component_costs.at(comp).at(sign).at(i));
  printf("\nmvsad");
  for (unsigned i = 0; i < 256; i++) printf(" %u", (unsigned)c.mv_sad_costs.at(0).at(0).at(i));
  /* two-pass encoding (trellis quantisation, encoder.cc:220-408): token costs of the default probability tables
   * (Costs::fill_token_costs, costs.cc:168-186) and the extra-bit + sign cost of every coefficient value */
  c.fill_token_costs(ProbabilityTables());
  printf("\ntokcost");
  for (unsigned i = 0; i < BLOCK_TYPES; i++)
    for (unsigned j = 0; j < COEF_BANDS; j++)
      for (unsigned k = 0; k < PREV_COEF_CONTEXTS; k++)
        for (unsigned t = 0; t < MAX_ENTROPY_TOKENS; t++) printf(" %u", (unsigned)c.token_costs.at(i).at(j).at(k).at(t));
  printf("\nvalcost");
  for (int v = -2048; v < 2048; v++) printf(" %u", (unsigned)Costs::coeff_base_cost((int16_t)v));
  printf("\n");
  return 0;
}
