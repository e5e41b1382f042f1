/* ref_encode -- TEST INFRASTRUCTURE. Synthesises a seeded YUV420 clip and encodes it to VP8/IVF
 * with the UNMODIFIED reference encoder (Encoder::encode_with_quantizer, encoder/encoder.cc:559),
 * because no >=1080p VP8 material and no external encoder exist in this image.
 * A new Encoder per GOP makes every GOP start with a key frame.
 *
 * usage: ref_encode OUT.ivf WIDTH HEIGHT FRAMES GOP QINDEX [SEED] [KIND]
 *   environment: REF_RAW=file.yuv  read planar YUV420 frames (display size) instead of synthesising;
 *                REF_TARGET=bytes  use Encoder::encode_with_target_size instead of a fixed quantiser;
 *                REF_MIN_SSIM=x    use Encoder::encode_with_minimum_ssim;
 *                REF_TWO_PASS=1    Encoder( ..., two_pass = true, ... ): key frames get the trellis pass (encoder.cc:220-408);
 *   prints one JSON line with encode seconds (source generation excluded), bytes and luma PSNR of
 *   the encoder's own reconstruction (Encoder::export_decoder, encoder.hh:378).
 *   KIND 0: smooth moving sinusoid + noise ("easy");  1: translating random-texture tiles ("hard");
 *        2: translating softly textured tiles, light noise ("medium", broadcast-like bitrate)
 */
/* diagnostics below look at the sampled frame of Encoder::estimate_size (a private member): this is test
 * infrastructure compiled against the unmodified headers, so the access specifiers are lifted for this file */
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#define private public
#define protected public
#include <chrono>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <random>
#include <vector>

#include "encoder.hh"
#include "ivf_writer.hh"

using namespace std;

static void synth(MutableRasterHandle& h, int w, int hgt, int t, int seed, int kind) {
  VP8Raster& r = h.get();
  const int W = r.width(), H = r.height();
  mt19937 rng(seed + t);
  uniform_int_distribution<int> noise(kind == 2 ? -1 : -3, kind == 2 ? 1 : 3);
  static vector<uint8_t> tex;
  if (kind >= 1 && tex.empty()) {
    const int rad = kind == 2 ? 9 : 3;
    mt19937 trng(seed * 7919 + 1);
    tex.resize(512 * 512);
    /* low-pass random texture so that sub-pel motion is meaningful */
    vector<int> raw(512 * 512);
    for (auto& v : raw) v = trng() & 255;
    for (int y = 0; y < 512; y++)
      for (int x = 0; x < 512; x++) {
        int s = 0;
        for (int dy = 0; dy < rad; dy++)
          for (int dx = 0; dx < rad; dx++) s += raw[((y + dy) & 511) * 512 + ((x + dx) & 511)];
        s /= rad * rad;
        if (kind == 2) s = 128 + (s - 128) * 4; /* restore contrast lost to the blur */
        tex[y * 512 + x] = s < 0 ? 0 : (s > 255 ? 255 : s);
      }
  }
  for (int y = 0; y < H; y++) {
    const int yy = y < hgt ? y : hgt - 1;
    for (int x = 0; x < W; x++) {
      const int xx = x < w ? x : w - 1; /* edge-extend like yuv4mpeg.cc:231-271 */
      int v;
      if (kind == 0) {
        v = 128 + (int)lround(60.0 * sin(0.02 * (xx + 3 * t)) * cos(0.015 * (yy + 2 * t)));
      } else {
        const int tile = ((xx >> 6) + (yy >> 6)) & 3;
        const double vx = (tile & 1 ? 3.25 : -3.25), vy = (tile & 2 ? 1.75 : -1.75);
        const int sx = (int)floor(xx + vx * t), sy = (int)floor(yy + vy * t);
        v = tex[(sy & 511) * 512 + (sx & 511)];
      }
      if (x < w && y < hgt) v += noise(rng);
      r.Y().at(x, y) = v < 0 ? 0 : (v > 255 ? 255 : v);
    }
  }
  for (int y = 0; y < H / 2; y++)
    for (int x = 0; x < W / 2; x++) {
      const int xx = min(x, (w + 1) / 2 - 1), yy = min(y, (hgt + 1) / 2 - 1);
      r.U().at(x, y) = 128 + (int)lround(30.0 * sin(0.01 * (xx + t)));
      r.V().at(x, y) = 128 + (int)lround(30.0 * cos(0.012 * (yy - t)));
    }
}

int main(int argc, char** argv) {
  try {
    if (argc < 7) { cerr << "usage: ref_encode OUT.ivf W H FRAMES GOP QINDEX [SEED] [KIND]\n"; return 2; }
    const int w = atoi(argv[2]), h = atoi(argv[3]), frames = atoi(argv[4]), gop = atoi(argv[5]), qi = atoi(argv[6]);
    const int seed = argc > 7 ? atoi(argv[7]) : 1234, kind = argc > 8 ? atoi(argv[8]) : 0;
    IVFWriter out(argv[1], "VP80", w, h, 30, 1);
    Optional<Encoder> enc;
    size_t total = 0;
    const char* raw_path = getenv("REF_RAW");
    const char* target_env = getenv("REF_TARGET");
    FILE* raw = raw_path ? fopen(raw_path, "rb") : nullptr;
    double enc_seconds = 0, sse = 0;
    for (int t = 0; t < frames; t++) {
      if (t % gop == 0) { enc.clear(); enc.initialize(w, h, getenv("REF_TWO_PASS") != nullptr, REALTIME_QUALITY); }
      MutableRasterHandle raster(w, h);
      if (raw) {
        VP8Raster& r = raster.get();
        vector<uint8_t> line(w);
        for (int y = 0; y < (int)r.height(); y++) {
          if (y < h && fread(line.data(), 1, w, raw) != (size_t)w) throw runtime_error("raw input too short");
          for (int x = 0; x < (int)r.width(); x++) r.Y().at(x, y) = line[x < w ? x : w - 1];
        }
        const int cw = (w + 1) / 2, ch = (h + 1) / 2;
        for (int pl = 0; pl < 2; pl++)
          for (int y = 0; y < (int)r.height() / 2; y++) {
            if (y < ch && fread(line.data(), 1, cw, raw) != (size_t)cw) throw runtime_error("raw input too short");
            for (int x = 0; x < (int)r.width() / 2; x++) (pl ? r.V() : r.U()).at(x, y) = line[x < cw ? x : cw - 1];
          }
      } else {
        synth(raster, w, h, t, seed, kind);
      }
      if (const char* ef = getenv("REF_EST_FRAME")) {  /* diagnostic: the size estimates the target-size search sees */
        if (atoi(ef) == t) {
          const char* lo = getenv("REF_EST_LO"); const char* hi = getenv("REF_EST_HI");
          for (int q = lo ? atoi(lo) : 4; q <= (hi ? atoi(hi) : 127); q++)
          {
            cerr << "estimate frame " << t << " qi " << q << " " << enc.get().estimate_frame_size(raster.get(), q) << "\n";
            const char* dump = getenv("REF_EST_DUMP");
            if (dump && t > 0) {  /* the sampled inter frame of that estimate, serialised the way estimate_size does */
              const vector<uint8_t> b = enc.get().subsampled_inter_frame_.get().serialize(enc.get().decoder_state_.probability_tables);
              ofstream(string(dump) + "." + to_string(q), ios::binary).write(reinterpret_cast<const char*>(b.data()), b.size());
            }
          }
        }
      }
      const auto t0 = chrono::steady_clock::now();
      const char* ssim_env = getenv("REF_MIN_SSIM");
      const vector<uint8_t> f = ssim_env ? enc.get().encode_with_minimum_ssim(raster.get(), atof(ssim_env))
                                : target_env ? enc.get().encode_with_target_size(raster.get(), atoi(target_env))
                                             : enc.get().encode_with_quantizer(raster.get(), qi);
      enc_seconds += chrono::duration<double>(chrono::steady_clock::now() - t0).count();
      out.append_frame(Chunk(&f.at(0), f.size()));
      total += f.size();
      const VP8Raster& rec = enc.get().export_decoder().get_references().last;
      for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
          const double d = (double)rec.Y().at(x, y) - (double)raster.get().Y().at(x, y);
          sse += d * d;
        }
      cerr << "frame " << t << " " << f.size() << " bytes\n";
    }
    if (raw) fclose(raw);
    cerr << "total " << total << " bytes\n";
    const double mse = sse / ((double)w * h * frames);
    printf("{\"frames\": %d, \"encode_s\": %.4f, \"fps\": %.3f, \"bytes\": %zu, \"psnr_y\": %.3f}\n", frames,
           enc_seconds, frames / enc_seconds, total, mse > 0 ? 10 * log10(255.0 * 255.0 / mse) : 99.0);
  } catch (const exception& e) {
    cerr << "ref_encode: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
