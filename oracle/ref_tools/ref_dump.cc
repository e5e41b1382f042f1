/* ref_dump -- TEST INFRASTRUCTURE. A driver (ours) around the UNMODIFIED reference
 * decoder classes, linked against oracle/_ref/libalfalfa_ref.a which the Makefile
 * compiles in place from /root/reference/src.  It calls the reference at the seam the
 * product replaces: DecoderState::parse_and_apply (decoder_state.hh:73),
 * Frame::decode (frame.cc:208/227), Frame::loopfilter (frame.cc:139), Frame::copy_to
 * (frame.cc:272), so that per-frame rasters before and after the loop filter can be
 * diffed against our oracle and CUDA path, and so the reference's CPU time per phase
 * can be measured (bench.py --impl reference, cpu_baseline.kind = "reference").
 *
 * usage:
 *   ref_dump shown FILE.ivf            display rectangle of shown frames -> stdout
 *                                      (same bytes as the reference's decode-to-stdout)
 *   ref_dump full  FILE.ivf OUT.bin    every frame, MB-aligned planes, pre+post loop filter
 *   ref_dump time  FILE.ivf [REPS] [FIRST] [COUNT]
 *                                      one JSON line with seconds per phase (best of REPS)
 *   ref_dump state FILE.ivf N          decode the first N frames, then Decoder::serialize (decoder.cc:54-69,
 *                                      the EncoderStateSerializer format of xc-enc -O) -> stdout
 *   ref_dump resume STATE FILE.ivf N   Decoder::deserialize( STATE ) (decoder.cc:71-81), then decode frames
 *                                      N.. of the file; display rectangle of shown frames -> stdout
 */
#include <chrono>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "decoder.hh"
#include "enc_state_serializer.hh"
#include "decoder_state.hh"
#include "frame.hh"
#include "ivf.hh"
#include "uncompressed_chunk.hh"

using namespace std;
using clk = chrono::steady_clock;

static void write_plane(FILE* f, const TwoD<uint8_t>& p) {
  for (unsigned r = 0; r < p.height(); r++) fwrite(&p.at(0, r), 1, p.width(), f);
}

static void write_raster(FILE* f, const VP8Raster& r) {
  write_plane(f, r.Y()); write_plane(f, r.U()); write_plane(f, r.V());
}

struct Phase { double parse = 0, recon = 0, lf = 0; };

template <class FrameType>
static void run_frame(DecoderState& state, References& refs, const UncompressedChunk& uc,
                      FILE* full, bool shown_mode, Phase& ph, uint32_t frame_no) {
  auto t0 = clk::now();
  FrameType frame = state.parse_and_apply<FrameType>(uc);
  auto t1 = clk::now();
  MutableRasterHandle raster(state.width, state.height);
  frame.decode(state.segmentation, refs, raster);
  auto t2 = clk::now();
  if (full) {
    uint32_t hdr[6] = {0x46525031u /* 'FRP1' */, frame_no, (uint32_t)uc.key_frame(),
                       (uint32_t)frame.show_frame(), raster.get().width(), raster.get().height()};
    fwrite(hdr, sizeof(hdr), 1, full);
    write_raster(full, raster.get());
  }
  auto t3 = clk::now();
  frame.loopfilter(state.segmentation, state.filter_adjustments, raster);
  auto t4 = clk::now();
  if (full) write_raster(full, raster.get());
  RasterHandle frozen(move(raster));
  frame.copy_to(frozen, refs);
  if (shown_mode && frame.show_frame()) frozen.get().dump(stdout);
  ph.parse += chrono::duration<double>(t1 - t0).count();
  ph.recon += chrono::duration<double>(t2 - t1).count();
  ph.lf += chrono::duration<double>(t4 - t3).count();
}

int main(int argc, char** argv) {
  try {
    if (argc < 3) { cerr << "usage: ref_dump shown|full|time FILE.ivf [...]\n"; return 2; }
    const string mode = argv[1];
    if (mode == "resume") {
      if (argc < 5) { cerr << "resume needs STATE FILE.ivf N\n"; return 2; }
      IVF file(argv[3]);
      Decoder decoder = EncoderStateDeserializer::build<Decoder>(string(argv[2]));
      for (uint32_t i = atoi(argv[4]); i < file.frame_count(); i++) {
        const Optional<RasterHandle> raster = decoder.parse_and_decode_frame(file.frame(i));
        if (raster.initialized()) raster.get().get().dump(stdout);
      }
      return 0;
    }
    IVF ivf(argv[2]);
    const uint16_t w = ivf.width(), h = ivf.height();
    FILE* full = nullptr;
    if (mode == "full") {
      if (argc < 4) { cerr << "full needs OUT\n"; return 2; }
      full = fopen(argv[3], "wb");
      if (!full) { perror("fopen"); return 1; }
    }
    int reps = 1;
    uint32_t first = 0, count = ivf.frame_count();
    if (mode == "state") {
      if (argc < 4) { cerr << "state needs N\n"; return 2; }
      count = atoi(argv[3]);
      if (count > ivf.frame_count()) count = ivf.frame_count();
    }
    if (mode == "time") {
      if (argc > 3) reps = atoi(argv[3]);
      if (argc > 4) first = atoi(argv[4]);
      if (argc > 5) count = atoi(argv[5]);
      if (first + count > ivf.frame_count()) count = ivf.frame_count() - first;
    }
    Phase best; double best_total = 1e30; uint32_t decoded = 0;
    for (int rep = 0; rep < reps; rep++) {
      DecoderState state(w, h);
      References refs(w, h);
      Phase ph; decoded = 0;
      bool started = false;
      for (uint32_t i = first; i < first + count; i++) {
        UncompressedChunk uc(ivf.frame(i), w, h, false);
        if (!started && !uc.key_frame()) continue; /* like FilePlayer: start at first key frame */
        started = true;
        if (uc.key_frame()) run_frame<KeyFrame>(state, refs, uc, full, mode == "shown", ph, i);
        else run_frame<InterFrame>(state, refs, uc, full, mode == "shown", ph, i);
        decoded++;
      }
      if (mode == "state") {
        EncoderStateSerializer odata;
        Decoder(state, refs).serialize(odata);
        odata.write(stdout);
      }
      double total = ph.parse + ph.recon + ph.lf;
      if (total < best_total) { best_total = total; best = ph; }
    }
    if (full) fclose(full);
    if (mode == "time") {
      printf("{\"frames\": %u, \"width\": %u, \"height\": %u, \"parse_s\": %.6f, \"recon_s\": %.6f, "
             "\"loopfilter_s\": %.6f, \"total_s\": %.6f, \"mpix_per_s\": %.3f}\n",
             decoded, w, h, best.parse, best.recon, best.lf, best_total,
             (double)w * h * decoded / 1e6 / best_total);
    }
  } catch (const exception& e) {
    cerr << "ref_dump: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
