// Host build of the device-side token decoder (alfalfa_b200/csrc/tokens_core.cuh) for
// tests/test_tokens_host.py: every frame is parsed twice from the same state -- once by the CPU
// front end (parser.cc), once with defer_tokens + decode_frame_tokens -- and the records compared.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../alfalfa_b200/csrc/parser.h"
#include "../alfalfa_b200/csrc/tokens_core.cuh"

struct Harness {
  vp8::State a, b;
  vp8::ParsedFrame pa, pb;
  std::vector<vp8gpu_token> tokens;
  std::vector<uint16_t> above;
  bool lockstep = false;
  Harness(int w, int h) : a(w, h), b(w, h) {}
};

extern "C" {
void* th_new(int w, int h) { return new Harness(w, h); }
void th_free(void* p) { delete static_cast<Harness*>(p); }
// which form of the decoder th_frame runs: 0 = one thread per frame, 1 = the lock-step state machine
void th_variant(void* p, int lockstep) { static_cast<Harness*>(p)->lockstep = lockstep != 0; }

// 0 = identical; 1 = parse error mismatch; 2 = descriptor; 3 = records; 4 = tokens; 5 = overflow; 6 = skip flag left
// negative = both parsers rejected the frame with that code
int th_frame(void* hp, const uint8_t* data, size_t len, uint32_t tok_cap_override, uint32_t* n_tokens) {
  Harness& H = *static_cast<Harness*>(hp);
  const int ra = vp8::parse_frame(H.a, data, len, H.pa);
  const int rb = vp8::parse_frame(H.b, data, len, H.pb, true);
  if (ra != rb) return 1;
  if (ra != VP8GPU_OK) return ra < 0 ? ra : -ra;
  const vp8gpu_frame_desc& d = H.pa.desc;
  const size_t n_mbs = (size_t)d.mb_cols * d.mb_rows;
  vp8::Geom g{};
  g.mb_cols = d.mb_cols;
  g.mb_rows = d.mb_rows;
  // capacity rule of Engine::token_ring_create
  size_t cap = (size_t)H.pb.tw.bits_len * 9 + 1024;
  if (cap > n_mbs * 400) cap = n_mbs * 400;
  if (tok_cap_override) cap = tok_cap_override;
  H.tokens.assign(cap + 1, 0xDEADBEEFu);
  H.above.assign(g.mb_cols, 0);
  uint32_t result[2] = {0, 0};
  vp8::TokJob J{};
  J.mbs = H.pb.mbs.data();
  J.tokens = H.tokens.data();
  J.bits = H.pb.tw.bits;
  J.coef_probs = H.pb.tw.coef_probs;
  J.result = result;
  memcpy(J.part_off, H.pb.tw.part_off, sizeof(J.part_off));
  memcpy(J.part_len, H.pb.tw.part_len, sizeof(J.part_len));
  J.nparts = H.pb.tw.nparts;
  J.tok_cap = (uint32_t)cap;
  if (H.lockstep) {
    // what Engine::token_ring_stage prepares: 2 bits per macroblock (flags & 3), 16 macroblocks per word
    std::vector<uint32_t> mbinfo((n_mbs + 15) / 16 + 1, 0);
    for (size_t i = 0; i < n_mbs; i++) mbinfo[i >> 4] |= (uint32_t)(H.pb.mbs.data()[i].flags & 3u) << (2 * (i & 15));
    J.mbinfo = mbinfo.data();
    vp8::tok::LockstepTables T;
    vp8::tok::fill_lockstep_tables(T, 0, 1);
    // one lane: the "transposed" tables are the plain ones
    vp8::tok::decode_frame_tokens_lockstep<1>(J, g, T, H.pb.tw.coef_probs, H.above.data());
  } else {
    alignas(16) uint8_t probs16[vp8::tok::kProbBytes];
    for (int e = 0; e < vp8::tok::kProbEntries; e++) vp8::tok::expand_prob_entry(H.pb.tw.coef_probs, probs16, e);
    vp8::tok::decode_frame_tokens(J, g, probs16, H.above.data());
  }
  if (n_tokens) *n_tokens = result[0];
  if (H.tokens[cap] != 0xDEADBEEFu) return 5;
  if (result[1]) return 5;
  vp8gpu_frame_desc db = H.pb.desc;
  db.n_tokens = result[0];
  if (memcmp(&db, &d, sizeof(d)) != 0) return 2;
  if (memcmp(H.pa.mbs.data(), H.pb.mbs.data(), n_mbs * sizeof(vp8gpu_mb)) != 0) return 3;
  if (d.n_tokens && memcmp(H.pa.tokens.data(), H.tokens.data(), (size_t)d.n_tokens * 4) != 0) return 4;
  if (d.n_split && memcmp(H.pa.split.data(), H.pb.split.data(), (size_t)d.n_split * 64) != 0) return 3;
  if (!(H.a == H.b)) return 1;
  return 0;
}
}
