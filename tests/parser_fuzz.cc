// Mutated frames (truncations, bit flips, random bytes) through the host front end -- whole parse, first partition
// only, and with labels kept -- and back through the re-serialiser, built with AddressSanitizer + UBSan by
// tests/test_parser_fuzz.py: the front end reads untrusted bytes and must answer with a status, never with a
// wild access (the reference throws Invalid / out_of_range, uncompressed_chunk.cc:34-130, chunk.hh:54-59).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <random>
#include <vector>
#include "../alfalfa_b200/csrc/parser.h"
#include "../alfalfa_b200/csrc/serializer.h"
int main(int argc, char** argv) {
  std::mt19937 rng(argc > 2 ? atoi(argv[2]) : 1);
  const int rounds = argc > 3 ? atoi(argv[3]) : 100;
  std::ifstream in(argv[1], std::ios::binary);
  std::vector<uint8_t> f((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  int w = f[12] | (f[13] << 8), h = f[14] | (f[15] << 8);
  uint32_t n = f[24] | (f[25] << 8) | (f[26] << 16) | (f[27] << 24);
  std::vector<std::vector<uint8_t>> frames;
  size_t pos = 32;
  for (uint32_t i = 0; i < n && pos + 12 <= f.size(); i++) { uint32_t len = f[pos] | (f[pos+1]<<8) | (f[pos+2]<<16) | (f[pos+3]<<24); frames.emplace_back(f.begin()+pos+12, f.begin()+pos+12+len); pos += 12+len; }
  long ok = 0, bad = 0, ser = 0;
  for (int round = 0; round < rounds; round++) {
    vp8::State st(w, h);
    vp8::ParsedFrame pf;
    pf.keep_verbatim = (round & 1);
    for (size_t i = 0; i < frames.size() && i < 12; i++) {
      std::vector<uint8_t> m = frames[i];
      const int kind = rng() % 4;
      if (kind == 0 && m.size() > 4) m.resize(rng() % m.size());                         // truncate
      else if (kind == 1) for (int k = 0; k < 1 + (int)(rng() % 8); k++) m[rng() % m.size()] ^= 1u << (rng() % 8);  // bit flips
      else if (kind == 2) for (int k = 0; k < 16 && !m.empty(); k++) m[rng() % m.size()] = rng();  // random bytes
      // kind 3: intact
      const bool defer = !pf.keep_verbatim && (rng() & 1);
      int rc = vp8::parse_frame(st, m.data(), m.size(), pf, defer);
      if (rc == 0) { ok++; if (pf.keep_verbatim) { auto b = vp8::serialize_parsed(pf); ser += !b.empty(); } } else bad++;
    }
  }
  printf("%s: parsed %ld rejected %ld reserialised %ld\n", argv[1], ok, bad, ser);
}
