/* ref_flatten -- TEST INFRASTRUCTURE, and the worked example of INTEGRATION.md section B.
 * The UNMODIFIED reference front end (UncompressedChunk + DecoderState::parse_and_apply, linked from
 * oracle/_ref/libalfalfa_ref.a) parses every frame; flatten() below turns the reference's parsed Frame object
 * (TwoD<Macroblock>, 25 Block objects per macroblock, frame.hh:56-61) into the flat records of include/vp8gpu.h;
 * the records then cross the narrow seam of SURVEY.md 8b -- vp8gpu_decode_parsed in place of the two calls
 * frame.decode( ... ); frame.loopfilter( ... ) at decoder/decoder.cc:109-111 -- and Frame::copy_to
 * (frame.cc:272-307) is applied to device raster handles.  This is what a maintainer of the reference adds to
 * Decoder::decode_frame to put the product behind the reference's own parser.
 *
 * usage:
 *   ref_flatten records FILE.ivf             flat records of every frame -> stdout, for a byte comparison
 *                                            with the product's own front end (tests/test_ref_flatten.py):
 *                                            per frame  desc (80 B) | mbs | tokens | split vectors
 *   ref_flatten decode LIBVP8GPU.so FILE.ivf the library is dlopen()ed (the binary is built in the container,
 *                                            the GPU is on another box); display rectangle of every shown
 *                                            frame -> stdout, same bytes as the reference's decode-to-stdout
 *   ref_flatten decode LIBVP8GPU.so --out DIR FILE.ivf...
 *                                            several files of ONE frame size in one process (the reference's frame
 *                                            pool allows a single size per process, frame_pool.cc:54-57): DIR/<name>.yuv
 */
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "decoder.hh"
#include "decoder_state.hh"
#include "frame.hh"
#include "ivf.hh"
#include "uncompressed_chunk.hh"

#include "../../include/vp8gpu.h"

using namespace std;

struct Flat {
  vp8gpu_frame_desc desc;
  vector<vp8gpu_mb> mbs;
  vector<vp8gpu_token> tokens;
  vector<vp8gpu_split_mvs> split;
};

static vp8gpu_quant to_quant(const Quantizer& q) {
  vp8gpu_quant o;
  o.y_dc = q.y_dc, o.y_ac = q.y_ac, o.y2_dc = q.y2_dc, o.y2_ac = q.y2_ac, o.uv_dc = q.uv_dc, o.uv_ac = q.uv_ac;
  return o;
}

/* tokens of one block: the non-zero coefficients in scan order (tokens.cc:50-135 stores them de-zigzagged) */
template <class BlockT>
static void flatten_block(const BlockT& b, unsigned block_no, unsigned first, vector<vp8gpu_token>& out) {
  for (unsigned i = first; i < 16; i++) {
    const unsigned pos = zigzag.at(i);
    const int16_t v = b.coefficients().at(pos);
    if (v) out.push_back(VP8GPU_TOKEN(block_no, pos, v));
  }
}

static void reference_fields(const KeyFrameHeader&, vp8gpu_frame_desc& d) {
  d.refresh_last = d.refresh_golden = d.refresh_alternate = 1;  /* KeyFrame::copy_to */
}
static void reference_fields(const InterFrameHeader& h, vp8gpu_frame_desc& d) {
  d.refresh_last = h.refresh_last;
  d.refresh_golden = h.refresh_golden_frame;
  d.refresh_alternate = h.refresh_alternate_frame;
  d.copy_to_golden = h.copy_buffer_to_golden.initialized() ? uint8_t(h.copy_buffer_to_golden.get()) : uint8_t(0);
  d.copy_to_alternate = h.copy_buffer_to_alternate.initialized() ? uint8_t(h.copy_buffer_to_alternate.get()) : uint8_t(0);
}

template <class FrameType>
static void flatten(const FrameType& frame, const DecoderState& state, Flat& out) {
  const auto& hdr = frame.header();
  const auto& mbs = frame.macroblocks();
  const unsigned cols = mbs.width(), rows = mbs.height();
  vp8gpu_frame_desc& d = out.desc;
  memset(&d, 0, sizeof(d));
  d.width = state.width, d.height = state.height, d.mb_cols = cols, d.mb_rows = rows;
  d.key_frame = remove_reference<decltype(hdr)>::type::key_frame();
  d.show_frame = frame.show_frame();
  d.loop_filter_level = hdr.loop_filter_level;
  d.sharpness = hdr.sharpness_level;
  reference_fields(hdr, d);
  /* Frame::calculate_segment_quantizers (frame.cc:186-206) with the reference's own Quantizer */
  for (unsigned s = 0; s < 4; s++) {
    QuantIndices qi(hdr.quant_indices);
    if (state.segmentation.initialized()) {
      const Segmentation& seg = state.segmentation.get();
      qi.y_ac_qi = seg.segment_quantizer_adjustments.at(s) +
                   (seg.absolute_segment_adjustments ? static_cast<Unsigned<7>>(0) : qi.y_ac_qi);
    }
    d.quant[s] = to_quant(Quantizer(qi));
  }
  out.mbs.assign(size_t(cols) * rows, vp8gpu_mb());
  out.tokens.clear();
  out.split.clear();
  mbs.forall_ij([&](const typename remove_reference<decltype(mbs.at(0, 0))>::type& mb, const unsigned col, const unsigned row) {
    vp8gpu_mb& r = out.mbs[size_t(row) * cols + col];
    memset(&r, 0, sizeof(r));
    const mbmode y_mode = mb.y_prediction_mode();
    const reference_frame ref = mb.header().reference();
    const bool has_y2 = mb.Y2().coded();
    r.y_mode = y_mode;
    r.uv_mode = ref == CURRENT_FRAME ? mb.uv_prediction_mode() : 0;
    r.ref_frame = ref;
    r.segment_id = state.segmentation.initialized() ? mb.segment_id() : 0;
    r.flags = has_y2 ? VP8GPU_MB_HAS_Y2 : 0;
    /* loop-filter level: Frame::loopfilter (frame.cc:139-182) + Macroblock::loopfilter (macroblock.cc:603-641)
       + the single clamp in the filter's constructor (loopfilter.cc:85) */
    int level = 0;
    if (hdr.loop_filter_level) {
      FilterParameters fp(hdr.filter_type, hdr.loop_filter_level, hdr.sharpness_level);
      if (state.segmentation.initialized()) {
        const Segmentation& seg = state.segmentation.get();
        fp.filter_level = seg.segment_filter_adjustments.at(r.segment_id) + (seg.absolute_segment_adjustments ? 0 : fp.filter_level);
      }
      if (state.filter_adjustments.initialized())
        fp.adjust(state.filter_adjustments.get().loopfilter_ref_adjustments,
                  state.filter_adjustments.get().loopfilter_mode_adjustments, ref, y_mode);
      level = fp.filter_level <= 0 ? 0 : (fp.filter_level > 63 ? 63 : fp.filter_level);
    }
    r.lf_level = level;
    /* modes and vectors */
    if (ref == CURRENT_FRAME) {
      if (y_mode == B_PRED)
        for (unsigned i = 0; i < 16; i++) r.b_modes |= uint64_t(mb.Y().at(i & 3, i >> 2).prediction_mode()) << (4 * i);
    } else {
      const MotionVector& base = mb.base_motion_vector();
      r.mv_x = base.x(), r.mv_y = base.y();
      if (y_mode == SPLITMV) {
        vp8gpu_split_mvs s;
        for (unsigned i = 0; i < 16; i++) {
          const MotionVector& mv = mb.Y().at(i & 3, i >> 2).motion_vector();
          s.mv[i][0] = mv.x(), s.mv[i][1] = mv.y();
        }
        r.split_idx = out.split.size();
        out.split.push_back(s);
      }
    }
    /* coefficients: Y2, Y 0..15, U, V (macroblock.cc:468-502) */
    r.tok_off = out.tokens.size();
    if (mb.has_nonzero()) {
      if (has_y2) flatten_block(mb.Y2(), VP8GPU_BLK_Y2, 0, out.tokens);
      for (unsigned i = 0; i < 16; i++) flatten_block(mb.Y().at(i & 3, i >> 2), i, has_y2 ? 1 : 0, out.tokens);
      for (unsigned i = 0; i < 4; i++) flatten_block(mb.U().at(i & 1, i >> 1), VP8GPU_BLK_U + i, 0, out.tokens);
      for (unsigned i = 0; i < 4; i++) flatten_block(mb.V().at(i & 1, i >> 1), VP8GPU_BLK_V + i, 0, out.tokens);
    }
    r.tok_cnt = out.tokens.size() - r.tok_off;
  });
  d.n_tokens = out.tokens.size();
  d.n_split = out.split.size();
}

/* the product's C ABI, resolved at run time */
struct Gpu {
  void* so = nullptr;
  decltype(&vp8gpu_ctx_create) ctx_create;
  decltype(&vp8gpu_ctx_destroy) ctx_destroy;
  decltype(&vp8gpu_last_error) last_error;
  decltype(&vp8gpu_frame_alloc) frame_alloc;
  decltype(&vp8gpu_frame_retain) frame_retain;
  decltype(&vp8gpu_frame_release) frame_release;
  decltype(&vp8gpu_decode_parsed) decode_parsed;
  decltype(&vp8gpu_frame_download_display) download_display;
  template <class F>
  void bind(F& f, const char* name) {
    f = reinterpret_cast<F>(dlsym(so, name));
    if (!f) throw runtime_error(string("missing symbol ") + name);
  }
  explicit Gpu(const char* path) {
    so = dlopen(path, RTLD_NOW);
    if (!so) throw runtime_error(string("dlopen: ") + dlerror());
    bind(ctx_create, "vp8gpu_ctx_create");
    bind(ctx_destroy, "vp8gpu_ctx_destroy");
    bind(last_error, "vp8gpu_last_error");
    bind(frame_alloc, "vp8gpu_frame_alloc");
    bind(frame_retain, "vp8gpu_frame_retain");
    bind(frame_release, "vp8gpu_frame_release");
    bind(decode_parsed, "vp8gpu_decode_parsed");
    bind(download_display, "vp8gpu_frame_download_display");
  }
};

/* References (decoder.hh:123-141) of device rasters; every slot owns one reference count */
struct GpuReferences {
  Gpu& g;
  vp8gpu_ctx* ctx;
  vp8gpu_frame_id last = -1, golden = -1, alternative = -1;
  void assign(vp8gpu_frame_id& slot, vp8gpu_frame_id v) {
    if (v >= 0) g.frame_retain(ctx, v);
    if (slot >= 0) g.frame_release(ctx, slot);
    slot = v;
  }
  /* Frame::copy_to (frame.cc:272-307), same order */
  void update(const vp8gpu_frame_desc& d, vp8gpu_frame_id raster) {
    if (d.key_frame) {
      assign(last, raster), assign(golden, raster), assign(alternative, raster);
      return;
    }
    if (d.copy_to_alternate == 1) assign(alternative, last);
    else if (d.copy_to_alternate == 2) assign(alternative, golden);
    if (d.copy_to_golden == 1) assign(golden, last);
    else if (d.copy_to_golden == 2) assign(golden, alternative);
    if (d.refresh_golden) assign(golden, raster);
    if (d.refresh_alternate) assign(alternative, raster);
    if (d.refresh_last) assign(last, raster);
  }
};

static void run_file(const string& mode, const char* path, Gpu* gpu, vp8gpu_ctx* ctx, FILE* sink) {
  const bool decode = mode == "decode";
  IVF ivf(path);
  const uint16_t w = ivf.width(), h = ivf.height();
  DecoderState state(w, h);
  Flat flat;
  unique_ptr<GpuReferences> refs;
  if (decode) refs.reset(new GpuReferences{*gpu, ctx});
  vector<uint8_t> display(size_t(w) * h + 2 * size_t((w + 1) / 2) * ((h + 1) / 2));
  bool started = false;
  for (uint32_t i = 0; i < ivf.frame_count(); i++) {
    UncompressedChunk uc(ivf.frame(i), w, h, false);
    if (!started && !uc.key_frame()) continue;
    started = true;
    if (uc.key_frame()) flatten(state.parse_and_apply<KeyFrame>(uc), state, flat);
    else flatten(state.parse_and_apply<InterFrame>(uc), state, flat);
    if (!decode) {
      fwrite(&flat.desc, sizeof(flat.desc), 1, sink);
      fwrite(flat.mbs.data(), sizeof(vp8gpu_mb), flat.mbs.size(), sink);
      fwrite(flat.tokens.data(), sizeof(vp8gpu_token), flat.tokens.size(), sink);
      fwrite(flat.split.data(), sizeof(vp8gpu_split_mvs), flat.split.size(), sink);
      continue;
    }
    /* the seam: frame.decode( segmentation, references, raster ); frame.loopfilter( ... ) */
    vp8gpu_frame_id raster = -1;
    if (gpu->frame_alloc(ctx, &raster) != VP8GPU_OK) throw runtime_error(gpu->last_error(ctx));
    const vp8gpu_frame_id three[3] = {refs->last, refs->golden, refs->alternative};
    if (gpu->decode_parsed(ctx, 0, &flat.desc, flat.mbs.data(), flat.tokens.data(), flat.split.data(), three, raster) != VP8GPU_OK)
      throw runtime_error(gpu->last_error(ctx));
    refs->update(flat.desc, raster);
    if (flat.desc.show_frame) {
      if (gpu->download_display(ctx, raster, display.data(), display.size()) != VP8GPU_OK) throw runtime_error(gpu->last_error(ctx));
      fwrite(display.data(), 1, display.size(), sink);
    }
    gpu->frame_release(ctx, raster);
  }
  if (decode) refs->assign(refs->last, -1), refs->assign(refs->golden, -1), refs->assign(refs->alternative, -1);
}

int main(int argc, char** argv) {
  try {
    if (argc < 3) {
      cerr << "usage: ref_flatten records FILE.ivf | ref_flatten decode LIBVP8GPU.so [--out DIR] FILE.ivf...\n";
      return 2;
    }
    const string mode = argv[1];
    if (mode != "decode") {
      run_file(mode, argv[2], nullptr, nullptr, stdout);
      return 0;
    }
    if (argc < 4) return 2;
    Gpu gpu(argv[2]);
    const bool many = string(argv[3]) == "--out";
    if (many && argc < 6) return 2;
    const int first = many ? 5 : 3;
    vp8gpu_ctx* ctx = nullptr;
    {
      IVF probe(argv[first]);
      if (gpu.ctx_create(0, probe.width(), probe.height(), 16, &ctx) != VP8GPU_OK)
        throw runtime_error("vp8gpu_ctx_create failed (no CUDA device?)");
    }
    for (int k = first; k < (many ? argc : first + 1); k++) {
      FILE* sink = stdout;
      if (many) {
        const string path = argv[k];
        sink = fopen((string(argv[4]) + "/" + path.substr(path.find_last_of('/') + 1) + ".yuv").c_str(), "wb");
        if (!sink) throw runtime_error("cannot write the output");
      }
      run_file(mode, argv[k], &gpu, ctx, sink);
      if (many) fclose(sink);
    }
    gpu.ctx_destroy(ctx);
  } catch (const exception& e) {
    cerr << "ref_flatten: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
