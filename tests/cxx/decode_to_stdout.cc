// C++ caller of the host mirror (alfalfa_b200/host/alfalfa_gpu.hh), in the shape of the reference's
// src/tests/decode-to-stdout.cc:43-49: open an IVF, feed every frame to Decoder::parse_and_decode_frame,
// write the display rectangle of every shown frame to stdout.  tests/test_gpu_cxx_host.py runs it over the
// golden vectors on the GPU box and compares the SHA-1 of the output with the vector's name
// (tests/decoding.test:14-15).
// usage: decode_to_stdout FILE.ivf [device_tokens(0|1)]
//        decode_to_stdout --out DIR device_tokens FILE.ivf...   (same, several files of ONE frame size in one process:
//                                                             one Context, a fresh Decoder per file, DIR/<name>.yuv)
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "../../alfalfa_b200/host/alfalfa_gpu.hh"

using namespace alfalfa_gpu;

static uint32_t le32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | (uint32_t(p[3]) << 24); }

static void decode_file(const Context& ctx, const std::vector<uint8_t>& file, bool device_tokens, FILE* out) {
  const uint16_t width = file[12] | (file[13] << 8), height = file[14] | (file[15] << 8);
  const uint32_t frame_count = le32(&file[24]);
  Decoder decoder(ctx, width, height);
  if (device_tokens) decoder.set_device_tokens(true);
  size_t pos = 32;
  bool started = false;
  for (uint32_t i = 0; i < frame_count && pos + 12 <= file.size(); i++) {
    const uint32_t n = le32(&file[pos]);
    const Chunk frame(&file[pos + 12], n);
    pos += 12 + n;
    if (!started && (n == 0 || (frame.buffer[0] & 1))) continue;  // FilePlayer starts at the first key frame
    started = true;
    const RasterHandle raster = decoder.parse_and_decode_frame(frame);
    if (raster.initialized()) {
      const std::vector<uint8_t> pixels = raster.dump(width, height);
      std::fwrite(pixels.data(), 1, pixels.size(), out);
    }
  }
}

static std::vector<uint8_t> slurp(const char* path) {
  std::ifstream in(path, std::ios::binary);
  std::vector<uint8_t> file((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  if (file.size() < 32) throw Invalid("not an IVF file");
  return file;
}

int main(int argc, char** argv) {
  try {
    if (argc < 2) {
      std::fprintf(stderr, "usage: %s FILE.ivf [device_tokens] | --out DIR device_tokens FILE.ivf...\n", argv[0]);
      return 2;
    }
    if (std::string(argv[1]) == "--out") {
      if (argc < 5) return 2;
      const bool device_tokens = std::atoi(argv[3]) != 0;
      const std::vector<uint8_t> first = slurp(argv[4]);
      Context ctx(0, first[12] | (first[13] << 8), first[14] | (first[15] << 8), 16);
      for (int k = 4; k < argc; k++) {
        const std::string path = argv[k];
        const std::string name = path.substr(path.find_last_of('/') + 1);
        FILE* out = std::fopen((std::string(argv[2]) + "/" + name + ".yuv").c_str(), "wb");
        if (!out) throw Invalid("cannot write the output");
        decode_file(ctx, slurp(argv[k]), device_tokens, out);
        std::fclose(out);
      }
      return 0;
    }
    const std::vector<uint8_t> file = slurp(argv[1]);
    Context ctx(0, file[12] | (file[13] << 8), file[14] | (file[15] << 8), 16);
    decode_file(ctx, file, argc > 2 && std::atoi(argv[2]), stdout);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s: %s\n", argv[0], e.what());
    return 1;
  }
  return 0;
}
