// C++ caller in the shape of salsify/salsify-sender.cc:492-518: one Encoder, copied twice per frame, the two
// copies encode the same source at different quantisers on two threads (std::async), one of the results is
// "sent", and the sender continues from the copy that produced it.  Checks, through the host mirror only:
//   - the two concurrent encodes equal the same encodes done one after the other on fresh copies,
//   - a Decoder fed with the sent frames equals the chosen encoder's export_decoder() after every frame
//     (Decoder::operator== = state + the three rasters, decoder.cc:153) and the minihashes agree,
//   - an Encoder constructed from that Decoder continues the stream with an inter frame that the Decoder decodes.
// Prints "ok <frames>" and exits 0.
#include <cmath>
#include <cstdio>
#include <exception>
#include <future>
#include <vector>

#include "../../alfalfa_b200/host/alfalfa_gpu.hh"

using namespace alfalfa_gpu;

struct Picture {
  std::vector<uint8_t> y, u, v;
  SourceFrame view(int w) const { return SourceFrame{y.data(), u.data(), v.data(), size_t(w), size_t((w + 1) / 2)}; }
};

static Picture synth(int w, int h, int t) {
  Picture p;
  p.y.resize(size_t(w) * h);
  p.u.resize(size_t((w + 1) / 2) * ((h + 1) / 2));
  p.v.resize(p.u.size());
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      p.y[size_t(y) * w + x] = uint8_t(128 + 60 * std::sin(0.05 * (x + 3 * t)) * std::cos(0.04 * (y + 2 * t)) + ((x * 7 + y * 13 + t) % 5));
  for (int y = 0; y < (h + 1) / 2; y++)
    for (int x = 0; x < (w + 1) / 2; x++) {
      p.u[size_t(y) * ((w + 1) / 2) + x] = uint8_t(128 + 30 * std::sin(0.03 * (x + t)));
      p.v[size_t(y) * ((w + 1) / 2) + x] = uint8_t(128 + 30 * std::cos(0.02 * (y - t)));
    }
  return p;
}

#define REQUIRE(c)                                                    \
  do {                                                                \
    if (!(c)) {                                                       \
      std::fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #c);     \
      return 1;                                                       \
    }                                                                 \
  } while (0)

int main() {
  try {
    const int w = 320, h = 240, frames = 5;
    Context ctx(0, w, h, 64);
    Encoder encoder(ctx, w, h);
    Decoder receiver(ctx, w, h);
    for (int t = 0; t < frames; t++) {
      const Picture pic = synth(w, h, t);
      const SourceFrame src = pic.view(w);
      Encoder good(encoder), bad(encoder);  // two copies per frame
      auto fut_good = std::async(std::launch::async, [&] { return good.encode_with_quantizer(src, 40); });
      auto fut_bad = std::async(std::launch::async, [&] { return bad.encode_with_quantizer(src, 90); });
      const std::vector<uint8_t> frame_good = fut_good.get(), frame_bad = fut_bad.get();
      REQUIRE(!frame_good.empty() && !frame_bad.empty());
      // the same encodes, sequentially, on fresh copies: the concurrent ones did not disturb each other
      REQUIRE(Encoder(encoder).encode_with_quantizer(src, 40) == frame_good);
      REQUIRE(Encoder(encoder).encode_with_quantizer(src, 90) == frame_bad);
      REQUIRE(frame_good.size() > frame_bad.size());
      // the original is untouched by its copies
      REQUIRE(encoder.export_decoder() == receiver);
      // "send" one of them (alternate) and continue from the copy that made it
      const bool pick_good = (t & 1) == 0;
      receiver.get_frame_output(Chunk(pick_good ? frame_good : frame_bad));
      encoder = pick_good ? good : bad;
      REQUIRE(encoder.export_decoder() == receiver);
      REQUIRE(encoder.minihash() == receiver.minihash());
    }
    // the same pattern through the target-size search, which is what salsify-sender runs on its two copies: two
    // bisections at once, each coding its probes in one launch on its own lane, both writers on the shared host pool
    for (int t = frames; t < frames + 3; t++) {
      const Picture pic2 = synth(w, h, t);
      const SourceFrame src = pic2.view(w);
      Encoder big(encoder), small(encoder);
      auto fut_big = std::async(std::launch::async, [&] { return big.encode_with_target_size(src, 3000); });
      auto fut_small = std::async(std::launch::async, [&] { return small.encode_with_target_size(src, 1000); });
      const std::vector<uint8_t> frame_big = fut_big.get(), frame_small = fut_small.get();
      REQUIRE(!frame_big.empty() && !frame_small.empty());
      REQUIRE(Encoder(encoder).encode_with_target_size(src, 3000) == frame_big);
      REQUIRE(Encoder(encoder).encode_with_target_size(src, 1000) == frame_small);
      receiver.get_frame_output(Chunk(frame_big));
      encoder = big;
      REQUIRE(encoder.export_decoder() == receiver);
    }
    // Encoder( const Decoder & ): continue the receiver's stream
    Encoder continued(receiver);
    const Picture pic = synth(w, h, frames + 3);
    const std::vector<uint8_t> next = continued.encode_with_quantizer(pic.view(w), 60);
    REQUIRE(!next.empty() && (next[0] & 1) == 1);  // an inter frame
    receiver.get_frame_output(Chunk(next));
    REQUIRE(continued.export_decoder() == receiver);
    std::printf("ok %d\n", frames + 4);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "exception: %s\n", e.what());
    return 1;
  }
  return 0;
}
