// C++ caller in the shape of "xc-enc --reencode" (frontend/xc-enc.cc:262-327) written against the host mirror: the
// prediction stream is parsed by its own Decoder into frame objects (labels kept, like the reference's KeyFrame /
// InterFrame), the Encoder is built from a deserialized Decoder (xc-enc -I), Encoder::reencode emits the chunk again.
// usage: reencode_chunk OUT.ivf WIDTH HEIGHT TARGETS.yuv PRED.ivf STATE.bin KF_Q_WEIGHT EXTRA_FRAME_CHUNK
// OUT.ivf gets the emitted frames (32-byte DKIF header, 12-byte frame headers).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <vector>

#include "../../alfalfa_b200/host/alfalfa_gpu.hh"

using namespace alfalfa_gpu;

static std::vector<uint8_t> slurp(const char* path) {
  std::vector<uint8_t> b;
  FILE* f = fopen(path, "rb");
  if (!f) throw std::runtime_error(std::string("cannot open ") + path);
  uint8_t buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) b.insert(b.end(), buf, buf + n);
  fclose(f);
  return b;
}
static uint32_t le32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | (uint32_t(p[3]) << 24); }

int main(int argc, char** argv) {
  try {
    if (argc < 9) {
      fprintf(stderr, "usage: reencode_chunk OUT.ivf W H TARGETS.yuv PRED.ivf STATE.bin KF_Q_WEIGHT EXTRA_FRAME_CHUNK\n");
      return 2;
    }
    const int w = atoi(argv[2]), h = atoi(argv[3]);
    const double kf_q_weight = atof(argv[7]);
    const bool extra_frame_chunk = atoi(argv[8]) != 0;
    const std::vector<uint8_t> ivf = slurp(argv[5]), targets = slurp(argv[4]), state = slurp(argv[6]);
    if (ivf.size() < 32 || memcmp(ivf.data(), "DKIF", 4) != 0) throw std::runtime_error("not an IVF file");

    Context ctx(0, w, h);
    Decoder pred_decoder(ctx, w, h);
    std::vector<ParsedFrame> prediction_frames;
    for (size_t off = 32; off + 12 <= ivf.size();) {
      const uint32_t n = le32(&ivf[off]);
      off += 12;
      if (off + n > ivf.size()) throw std::runtime_error("truncated IVF");
      ParsedFrame frame = pred_decoder.parse_frame(Chunk{&ivf[off], n}, true);
      pred_decoder.decode_frame(frame);
      prediction_frames.push_back(frame);
      off += n;
    }
    const size_t cw = (w + 1) / 2, ch = (h + 1) / 2, frame_bytes = size_t(w) * h + 2 * cw * ch;
    if (targets.size() < frame_bytes * prediction_frames.size()) throw std::runtime_error("target input too short");
    std::vector<SourceFrame> original_rasters;
    for (size_t i = 0; i < prediction_frames.size(); i++) {
      const uint8_t* p = targets.data() + i * frame_bytes;
      original_rasters.push_back(SourceFrame{p, p + size_t(w) * h, p + size_t(w) * h + cw * ch, size_t(w), cw});
    }

    Encoder encoder(Decoder::deserialize(ctx, state, w, h));
    const std::vector<std::vector<uint8_t>> out = encoder.reencode(original_rasters, prediction_frames, kf_q_weight, extra_frame_chunk);

    FILE* f = fopen(argv[1], "wb");
    if (!f) throw std::runtime_error("cannot write output");
    uint8_t hdr[32] = {'D', 'K', 'I', 'F', 0, 0, 32, 0, 'V', 'P', '8', '0'};
    hdr[12] = w & 255, hdr[13] = w >> 8, hdr[14] = h & 255, hdr[15] = h >> 8;
    hdr[16] = 1, hdr[20] = 1;
    hdr[24] = out.size() & 255, hdr[25] = (out.size() >> 8) & 255;
    fwrite(hdr, 1, 32, f);
    for (size_t i = 0; i < out.size(); i++) {
      uint8_t fh[12] = {0};
      const uint32_t n = (uint32_t)out[i].size();
      fh[0] = n & 255, fh[1] = (n >> 8) & 255, fh[2] = (n >> 16) & 255, fh[3] = n >> 24;
      fh[4] = i & 255, fh[5] = (i >> 8) & 255;
      fwrite(fh, 1, 12, f);
      fwrite(out[i].data(), 1, out[i].size(), f);
    }
    fclose(f);
    printf("ok %zu\n", out.size());
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 1;
  }
}
