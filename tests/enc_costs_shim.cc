// Host build of the encoder's rate tables (alfalfa_b200/csrc/enc_costs.cc) for tests/test_enc_costs_host.py.
#include <string.h>

#include "../alfalfa_b200/csrc/enc_costs.h"

extern "C" {
int ec_size() { return (int)sizeof(vp8::EncTables); }
void ec_build(vp8::EncTables* t) { vp8::build_enc_tables(*t); }
void ec_rd(int y_ac, unsigned* rm, unsigned* dm) { vp8::rd_multipliers(y_ac, rm, dm); }
int ec_trellis_size() { return (int)sizeof(vp8::TrellisTables); }
void ec_trellis(vp8::TrellisTables* t) { vp8::build_trellis_tables(*t); }
}
