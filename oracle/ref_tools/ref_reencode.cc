/* ref_reencode -- TEST INFRASTRUCTURE.  The UNMODIFIED reference's Encoder::reencode (encoder/reencode.cc:315-381),
 * driven the way frontend/xc-enc.cc:262-327 drives it ("xc-enc --reencode"): an Encoder built from a serialized
 * Decoder (the state a receiver is in when the chunk starts), the chunk's target rasters, and the chunk as it was
 * coded independently ("prediction" frames, parsed by their own decoder).  Truth for tests/test_reencode*.py.
 *
 * usage: ref_reencode OUT.ivf WIDTH HEIGHT TARGETS.yuv PRED.ivf STATE.bin KF_Q_WEIGHT EXTRA_FRAME_CHUNK [STATE_OUT.bin]
 *   TARGETS.yuv  one planar YUV420 frame (display size) per frame of PRED.ivf
 *   STATE.bin    Decoder::serialize output (EncoderStateSerializer, enc_state_serializer.hh) the Encoder starts from
 *   the prediction decoder starts fresh (PRED.ivf begins with a key frame)
 */
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "decoder.hh"
#include "enc_state_serializer.hh"
#include "encoder.hh"
#include "ivf.hh"
#include "ivf_writer.hh"
#include "uncompressed_chunk.hh"

using namespace std;

static void read_plane(FILE* f, TwoD<uint8_t>& plane, int w, int h) {
  vector<uint8_t> line(w);
  for (int y = 0; y < (int)plane.height(); y++) {
    if (y < h && fread(line.data(), 1, w, f) != (size_t)w) throw runtime_error("target input too short");
    for (int x = 0; x < (int)plane.width(); x++) plane.at(x, y) = line[x < w ? x : w - 1];
  }
}

int main(int argc, char** argv) {
  try {
    if (argc < 9) {
      cerr << "usage: ref_reencode OUT.ivf W H TARGETS.yuv PRED.ivf STATE.bin KF_Q_WEIGHT EXTRA_FRAME_CHUNK [STATE_OUT.bin]\n";
      return 2;
    }
    const string out_path = argv[1];
    const int w = atoi(argv[2]), h = atoi(argv[3]);
    const double kf_q_weight = atof(argv[7]);
    const bool extra_frame_chunk = atoi(argv[8]) != 0;

    IVF pred_ivf{argv[5]};
    if (pred_ivf.width() != w || pred_ivf.height() != h) throw runtime_error("prediction ivf size mismatch");

    vector<RasterHandle> originals;
    FILE* f = fopen(argv[4], "rb");
    if (!f) throw runtime_error("cannot open targets");
    for (unsigned i = 0; i < pred_ivf.frame_count(); i++) {
      MutableRasterHandle r(w, h);
      read_plane(f, r.get().Y(), w, h);
      read_plane(f, r.get().U(), (w + 1) / 2, (h + 1) / 2);
      read_plane(f, r.get().V(), (w + 1) / 2, (h + 1) / 2);
      originals.emplace_back(move(r));
    }
    fclose(f);

    // the chunk as its own decoder sees it: every frame parsed into the reference's frame object and decoded, so that
    // the next one parses against the right state (what xc-enc does before it calls Encoder::reencode)
    typedef pair<Optional<KeyFrame>, Optional<InterFrame>> PredictionFrame;
    vector<PredictionFrame> prediction_frames;
    Decoder chunk_decoder(w, h);
    const unsigned n_frames = pred_ivf.frame_count();
    prediction_frames.reserve(n_frames);
    for (unsigned idx = 0; idx < n_frames; idx++) {
      const UncompressedChunk chunk(pred_ivf.frame(idx), w, h, false);
      PredictionFrame slot;
      if (chunk.key_frame()) {
        slot.first.initialize(chunk_decoder.parse_frame<KeyFrame>(chunk));
        chunk_decoder.decode_frame(slot.first.get());
      } else {
        slot.second.initialize(chunk_decoder.parse_frame<InterFrame>(chunk));
        chunk_decoder.decode_frame(slot.second.get());
      }
      prediction_frames.push_back(move(slot));
    }

    Encoder encoder(EncoderStateDeserializer::build<Decoder>(argv[6]), false, REALTIME_QUALITY);
    {
      IVFWriter output{out_path, "VP80", (uint16_t)w, (uint16_t)h, 1, 1};
      output.set_expected_decoder_entry_hash(encoder.export_decoder().get_hash().hash());
      encoder.reencode(originals, prediction_frames, kf_q_weight, extra_frame_chunk, output);
    }
    if (argc > 9) {
      EncoderStateSerializer odata = {};
      encoder.export_decoder().serialize(odata);
      odata.write(argv[9]);
    }
    return 0;
  } catch (const exception& e) {
    cerr << "ref_reencode: " << e.what() << "\n";
    return 1;
  }
}
