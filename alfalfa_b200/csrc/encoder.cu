// encoder.cu -- host side of the encoder path: Encoder (encoder/encoder.hh:345-382) on the device.
//
// First slice of SURVEY.md 8a row a16: encode_with_quantizer / encode_with_target_size for a
// key frame followed by inter frames predicted from LAST, 16x16 intra modes, SAD-driven decisions
// (the reference's RD search, B_PRED / SPLITMV, trellis and SSIM-driven loop-filter search are not
// reproduced yet).  What is exact: transform / quantiser arithmetic (dct.cc, quantization.cc) and
// the closed loop -- the emitted frame decodes (reference decoder, oracle, this library) to exactly
// the reconstruction the encoder keeps as its LAST reference, which is what Encoder::export_decoder
// (encoder.hh:378) promises.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/vp8gpu.h"
#include "engine.hpp"
#include "serializer.h"
#include "vp8_tables.h"

using vp8::Engine;

// shared with capi.cc
struct vp8gpu_ctx_view {
  Engine* engine;
};
extern "C" Engine* vp8gpu_ctx_engine(vp8gpu_ctx* ctx);
extern "C" int vp8gpu_ctx_next_lane(vp8gpu_ctx* ctx);

struct vp8gpu_encoder {
  vp8gpu_ctx* ctx = nullptr;
  Engine* e = nullptr;
  int lane = 0;
  bool has_state = false;
  int last = -1;      // LAST reference = previous reconstruction
  int src = -1;       // device raster holding the (edge-extended) source frame
  int last_qi = -1;   // last_y_ac_qi_
  int last_lf = -1;   // loop_filter_level_ (encoder.hh:144): -1 = not initialised
  double last_ssim = -1.0;  // encode_stats_.ssim of the last frame
  // device scratch: EncJob | DevJob | sync ints | mbs | mv | sad | tokens
  uint8_t* dev = nullptr;
  size_t off_encjob = 0, off_devjob = 0, off_sync = 0, off_mbs = 0, off_mv = 0, off_sad = 0, off_tokens = 0, dev_bytes = 0;
  uint32_t tok_cap = 0;
  // pinned host buffers
  uint8_t* h_hdr = nullptr;      // EncJob + DevJob
  vp8gpu_mb* h_mbs = nullptr;
  vp8gpu_token* h_tokens = nullptr;
  uint8_t* h_src = nullptr;      // padded planes
  uint32_t* h_count = nullptr;
  uint64_t stat_frames = 0;
  // candidates of a quantiser search run as concurrent device passes: two more sets of buffers on their
  // own lanes (created on first use); they read this encoder's source, LAST and motion vectors
  vp8gpu_encoder* helper[2] = {nullptr, nullptr};
  vp8gpu_encoder* owner = nullptr;   // set in a helper
  cudaEvent_t motion_done = nullptr; // the motion search of the current source frame (owner's lane)
  int pending_out = -1;              // output raster of a pass that has been launched but not collected
  bool pending_key = false;
};

namespace {
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
int clamp_q(int q) { return q < 0 ? 0 : (q > 127 ? 127 : q); }
vp8gpu_quant make_quant(int qi) {  // Quantizer::Quantizer, quantization.cc:83-93 (all deltas zero)
  vp8gpu_quant q;
  q.y_ac = k_ac_q[clamp_q(qi)];
  q.y_dc = k_dc_q[clamp_q(qi)];
  q.y2_ac = static_cast<uint16_t>(k_ac_q[clamp_q(qi)] * 155 / 100);
  q.y2_dc = static_cast<uint16_t>(k_dc_q[clamp_q(qi)] * 2);
  q.uv_ac = k_ac_q[clamp_q(qi)];
  q.uv_dc = k_dc_q[clamp_q(qi)];
  if (q.y2_ac < 8) q.y2_ac = 8;
  if (q.uv_dc > 132) q.uv_dc = 132;
  return q;
}
#define CUE(call)                                                       \
  do {                                                                  \
    cudaError_t e__ = (call);                                           \
    if (e__ != cudaSuccess) return enc->e->cuda_fail(e__, #call);       \
  } while (0)

// launch half of a pass: everything up to the asynchronous download of the token count and the records
int encode_launch(vp8gpu_encoder* enc, bool key, int qi, bool search_motion) {
  Engine* e = enc->e;
  const vp8::Geom& g = e->geom();
  const size_t n_mbs = (size_t)g.mb_cols * g.mb_rows;
  vp8gpu_encoder* own = enc->owner ? enc->owner : enc;  // whose source / LAST / vectors are used
  if (int rc = e->ensure_lane(enc->lane)) return rc;
  cudaStream_t s = e->stream(enc->lane);
  int out = -1;
  int rc = e->frame_alloc(&out);
  if (rc != VP8GPU_OK) return rc;
  int ids[3] = {own->src, out, own->last};
  rc = e->acquire_frames(enc->lane, ids, key ? 2 : 3, 2u);  // only `out` is written
  if (rc != VP8GPU_OK) {
    e->frame_release(out);
    return rc;
  }
  vp8::EncJob* ej = reinterpret_cast<vp8::EncJob*>(enc->h_hdr);
  memset(enc->h_hdr, 0, 512);
  int* d_sync = reinterpret_cast<int*>(enc->dev + enc->off_sync);
  ej->src = e->frame_dev(own->src);
  ej->ref = key ? nullptr : e->frame_dev(own->last);
  ej->out = e->frame_dev(out);
  ej->mbs = reinterpret_cast<vp8gpu_mb*>(enc->dev + enc->off_mbs);
  ej->tokens = reinterpret_cast<vp8gpu_token*>(enc->dev + enc->off_tokens);
  ej->tok_counter = reinterpret_cast<uint32_t*>(d_sync + 96);
  ej->tok_cap = enc->tok_cap;
  ej->mv = reinterpret_cast<int*>(own->dev + own->off_mv);
  ej->sad = reinterpret_cast<uint32_t*>(own->dev + own->off_sad);
  ej->progress = d_sync + 128;
  ej->q = make_quant(qi);
  ej->key_frame = key;
  ej->lf_level = 1;  // records carry "filtered"; the level itself is chosen afterwards (choose_loop_filter)
  auto fail = [&](int code) {
    e->frame_release(out);
    return code;
  };
#define CUF(call)                                                        \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) return fail(e->cuda_fail(e__, #call));       \
  } while (0)
  CUF(cudaMemcpyAsync(enc->dev + enc->off_encjob, enc->h_hdr, 512, cudaMemcpyHostToDevice, s));
  CUF(cudaMemsetAsync(d_sync, 0, sizeof(int) * (128 + 2 * (size_t)g.mb_rows), s));
  const vp8::EncJob* d_ej = reinterpret_cast<const vp8::EncJob*>(enc->dev + enc->off_encjob);
  int launches = 0;
  if (!key) {
    if (search_motion) {  // vectors do not depend on the quantiser: searched once per source frame, by the owner
      if (enc->owner) return fail(e->fail(VP8GPU_ERR_LOGIC, "motion search belongs to the owning encoder"));
      if (int ce = vp8::launch_enc_motion(d_ej, g, s)) return fail(e->cuda_fail((cudaError_t)ce, "k_enc_motion"));
      launches++;
      if (!enc->motion_done) CUF(cudaEventCreateWithFlags(&enc->motion_done, cudaEventDisableTiming));
      CUF(cudaEventRecord(enc->motion_done, s));
    } else if (enc->owner && own->motion_done) {
      CUF(cudaStreamWaitEvent(s, own->motion_done, 0));  // the vectors come from the owner's lane
    }
  }
  if (int ce = vp8::launch_enc_mb(d_ej, g, d_sync + 0, s)) return fail(e->cuda_fail((cudaError_t)ce, "k_enc_mb"));
  launches++;
  e->count_launches(launches);
  e->mark_frames(enc->lane, ids, key ? 2 : 3, 2u);
  // results back: token count first, then the records
  CUF(cudaMemcpyAsync(enc->h_count, ej->tok_counter, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUF(cudaMemcpyAsync(enc->h_mbs, ej->mbs, n_mbs * sizeof(vp8gpu_mb), cudaMemcpyDeviceToHost, s));
  enc->pending_out = out;
  enc->pending_key = key;
  return VP8GPU_OK;
#undef CUF
}

// collect half: wait for the pass, fetch the tokens.  *out_frame = the reconstruction BEFORE the loop
// filter (the caller releases it or keeps it as LAST).
int encode_collect(vp8gpu_encoder* enc, int* out_frame) {
  Engine* e = enc->e;
  cudaStream_t s = e->stream(enc->lane);
  const int out = enc->pending_out;
  enc->pending_out = -1;
  if (out < 0) return e->fail(VP8GPU_ERR_LOGIC, "encode_collect without a launched pass");
  auto fail = [&](int code) {
    e->frame_release(out);
    return code;
  };
  cudaError_t ce = cudaStreamSynchronize(s);
  if (ce != cudaSuccess) return fail(e->cuda_fail(ce, "encoder pass"));
  const uint32_t n_tok = *enc->h_count;
  if (n_tok > enc->tok_cap) return fail(e->fail(VP8GPU_ERR_NOMEM, "encoder token pool overflow"));
  if (n_tok) {
    ce = cudaMemcpyAsync(enc->h_tokens, enc->dev + enc->off_tokens, (size_t)n_tok * sizeof(vp8gpu_token), cudaMemcpyDeviceToHost, s);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    if (ce != cudaSuccess) return fail(e->cuda_fail(ce, "encoder token download"));
  }
  *out_frame = out;
  return VP8GPU_OK;
}

// One encoding pass at quantiser index qi: motion search (once per source frame), the mode-decision /
// transform / reconstruction wavefront, records and tokens back on the host.
int encode_core(vp8gpu_encoder* enc, bool key, int qi, bool search_motion, int* out_frame) {
  const int rc = encode_launch(enc, key, qi, search_motion);
  return rc == VP8GPU_OK ? encode_collect(enc, out_frame) : rc;
}

// the compressed frame of the last encode_core pass
int encode_bytes(vp8gpu_encoder* enc, bool key, int qi, int lf_level, std::vector<uint8_t>& bytes) {
  Engine* e = enc->e;
  vp8::EncodeHeader h;
  h.key_frame = key;
  h.show_frame = true;
  h.width = e->width();
  h.height = e->height();
  h.y_ac_qi = qi;
  h.loop_filter_level = lf_level;
  h.sharpness = 0;
  h.optimize_token_probs = true;
  // eight DCT partitions (row r -> partition r % 8, frame.cc:131-136): the writer records and codes them
  // on eight host threads, which is most of the host time of an encoding pass
  vp8::EncodeFeatures ft;
  ft.log2_partitions = (h.height + 15) / 16 >= 16 ? 3 : 0;
  if (const char* v = getenv("VP8GPU_ENC_LOG2_PARTS")) ft.log2_partitions = atoi(v) & 3;  // tuning knob
  bytes = vp8::serialize_frame(h, enc->h_mbs, enc->h_tokens, nullptr, &ft);
  if (bytes.empty()) return e->fail(VP8GPU_ERR_LOGIC, "serializer rejected the device records");
  return VP8GPU_OK;
}

// one loop-filter pass over `frame` with every macroblock at `level` (in place)
int filter_frame(vp8gpu_encoder* enc, int frame, bool key, int level) {
  Engine* e = enc->e;
  const vp8::Geom& g = e->geom();
  cudaStream_t s = e->stream(enc->lane);
  if (level <= 0) return VP8GPU_OK;  // a frame-level 0 disables the filter (frame.cc:144)
  vp8::DevJob* dj = reinterpret_cast<vp8::DevJob*>(enc->h_hdr + 512);
  memset(dj, 0, 512);
  int* d_sync = reinterpret_cast<int*>(enc->dev + enc->off_sync);
  dj->mbs = reinterpret_cast<const vp8gpu_mb*>(enc->dev + enc->off_mbs);
  dj->tokens = reinterpret_cast<const vp8gpu_token*>(enc->dev + enc->off_tokens);
  dj->out = e->frame_dev(frame);
  dj->lf_progress = d_sync + 128 + g.mb_rows;
  dj->intra_progress = d_sync + 128;
  dj->key_frame = key;
  dj->sharpness = 0;
  dj->lf_enabled = 1;
  dj->lf_force = (uint8_t)level;
  int ids[1] = {frame};
  int rc = e->acquire_frames(enc->lane, ids, 1);
  if (rc != VP8GPU_OK) return rc;
  CUE(cudaMemcpyAsync(enc->dev + enc->off_encjob + 512, dj, 512, cudaMemcpyHostToDevice, s));
  CUE(cudaMemsetAsync(d_sync + 32, 0, sizeof(int), s));                                     // ticket
  CUE(cudaMemsetAsync(d_sync + 128 + g.mb_rows, 0, sizeof(int) * (size_t)g.mb_rows, s));    // row progress
  const vp8::DevJob* d_dj = reinterpret_cast<const vp8::DevJob*>(enc->dev + enc->off_encjob + 512);
  if (int ce = vp8::launch_loopfilter(d_dj, 1, g, d_sync + 32, s)) return e->cuda_fail((cudaError_t)ce, "k_loopfilter");
  e->count_launches(1);
  e->mark_frames(enc->lane, ids, 1);
  // the pinned descriptor is rewritten by the next call: wait until it has been read
  CUE(cudaStreamSynchronize(s));
  return VP8GPU_OK;
}

// Encoder::apply_best_loopfilter_settings (encoder.cc:460-508): try loop-filter levels in ascending
// order -- all of 0..63 for the first frame, the previous level +-1 afterwards -- on a copy of the
// reconstruction, keep going while the luma SSIM against the source improves, then filter the
// reconstruction itself at the best level.
int choose_loop_filter(vp8gpu_encoder* enc, int recon, bool key, int* level_out, double* ssim_out) {
  Engine* e = enc->e;
  int lo = 0, hi = 63;
  if (enc->last_lf >= 0) {
    lo = enc->last_lf > 0 ? enc->last_lf - 1 : 0;
    hi = enc->last_lf + 1 > 63 ? 63 : enc->last_lf + 1;
  }
  int temp = -1;
  int rc = e->frame_alloc(&temp);
  if (rc != VP8GPU_OK) return rc;
  int best = 0;
  double best_ssim = -1.0;
  for (int level = lo; level <= hi; level++) {
    rc = e->frame_copy(temp, recon, enc->lane);
    if (rc == VP8GPU_OK) rc = filter_frame(enc, temp, key, level);
    double q = 0;
    if (rc == VP8GPU_OK) rc = e->frames_ssim(temp, enc->src, enc->lane, &q);
    if (rc != VP8GPU_OK) break;
    if (q > best_ssim) {
      best_ssim = q;
      best = level;
    } else {
      break;
    }
  }
  e->frame_release(temp);
  if (rc != VP8GPU_OK) return rc;
  rc = filter_frame(enc, recon, key, best);
  *level_out = best;
  *ssim_out = best_ssim;
  return rc;
}

// source planes (display size) -> MB-aligned raster on the device, edges replicated like the
// reference's input reader (input/yuv4mpeg.cc:231-271)
int upload_source(vp8gpu_encoder* enc, const uint8_t* y, size_t ys, const uint8_t* u, const uint8_t* v, size_t cs) {
  Engine* e = enc->e;
  const vp8::Geom& g = e->geom();
  const int w = e->width(), h = e->height(), cw = (w + 1) / 2, ch = (h + 1) / 2;
  uint8_t* py = enc->h_src;
  uint8_t* pu = py + (size_t)g.W * g.H;
  uint8_t* pv = pu + (size_t)(g.W / 2) * (g.H / 2);
  for (int r = 0; r < g.H; r++) {
    const uint8_t* srow = y + (size_t)(r < h ? r : h - 1) * ys;
    uint8_t* drow = py + (size_t)r * g.W;
    memcpy(drow, srow, w);
    if (g.W > w) memset(drow + w, srow[w - 1], g.W - w);
  }
  for (int pl = 0; pl < 2; pl++) {
    const uint8_t* sp = pl ? v : u;
    uint8_t* dp = pl ? pv : pu;
    for (int r = 0; r < g.H / 2; r++) {
      const uint8_t* srow = sp + (size_t)(r < ch ? r : ch - 1) * cs;
      uint8_t* drow = dp + (size_t)r * (g.W / 2);
      memcpy(drow, srow, cw);
      if (g.W / 2 > cw) memset(drow + cw, srow[cw - 1], g.W / 2 - cw);
    }
  }
  return e->frame_upload(enc->src, py, g.W, pu, pv, g.W / 2);
}

}  // namespace

extern "C" {

int vp8gpu_encoder_create(vp8gpu_ctx* ctx, vp8gpu_encoder** out) {
  if (!ctx || !out) return VP8GPU_ERR_LOGIC;
  vp8gpu_encoder* enc = new vp8gpu_encoder();
  enc->ctx = ctx;
  enc->e = vp8gpu_ctx_engine(ctx);
  enc->lane = vp8gpu_ctx_next_lane(ctx);
  Engine* e = enc->e;
  const vp8::Geom& g = e->geom();
  const size_t n_mbs = (size_t)g.mb_cols * g.mb_rows;
  enc->tok_cap = (uint32_t)(n_mbs * 400);
  size_t off = 0;
  enc->off_encjob = off;
  off = align_up(off + 1024, 256);
  enc->off_sync = off;
  off = align_up(off + sizeof(int) * (128 + 2 * (size_t)g.mb_rows), 256);
  enc->off_mbs = off;
  off = align_up(off + n_mbs * sizeof(vp8gpu_mb), 256);
  enc->off_mv = off;
  off = align_up(off + n_mbs * 2 * sizeof(int), 256);
  enc->off_sad = off;
  off = align_up(off + n_mbs * sizeof(uint32_t), 256);
  enc->off_tokens = off;
  off = align_up(off + (size_t)enc->tok_cap * sizeof(vp8gpu_token), 256);
  enc->dev_bytes = off;
  cudaSetDevice(e->device());
  if (cudaMalloc(&enc->dev, enc->dev_bytes) != cudaSuccess || cudaHostAlloc(&enc->h_hdr, 1024, cudaHostAllocDefault) != cudaSuccess ||
      cudaHostAlloc(&enc->h_mbs, n_mbs * sizeof(vp8gpu_mb), cudaHostAllocDefault) != cudaSuccess ||
      cudaHostAlloc(&enc->h_tokens, (size_t)enc->tok_cap * sizeof(vp8gpu_token), cudaHostAllocDefault) != cudaSuccess ||
      cudaHostAlloc(&enc->h_src, (size_t)g.W * g.H * 3 / 2, cudaHostAllocDefault) != cudaSuccess ||
      cudaHostAlloc(&enc->h_count, 64, cudaHostAllocDefault) != cudaSuccess || e->frame_alloc(&enc->src) != VP8GPU_OK) {
    vp8gpu_encoder_destroy(enc);
    return e->fail(VP8GPU_ERR_NOMEM, "encoder allocation failed");
  }
  *out = enc;
  return VP8GPU_OK;
}

void vp8gpu_encoder_destroy(vp8gpu_encoder* enc) {
  if (!enc) return;
  for (vp8gpu_encoder*& hlp : enc->helper) {
    vp8gpu_encoder_destroy(hlp);
    hlp = nullptr;
  }
  if (enc->motion_done) cudaEventDestroy(enc->motion_done);
  cudaSetDevice(enc->e->device());
  cudaStreamSynchronize(enc->e->stream(enc->lane));
  if (enc->last >= 0) enc->e->frame_release(enc->last);
  if (enc->src >= 0) enc->e->frame_release(enc->src);
  if (enc->dev) cudaFree(enc->dev);
  if (enc->h_hdr) cudaFreeHost(enc->h_hdr);
  if (enc->h_mbs) cudaFreeHost(enc->h_mbs);
  if (enc->h_tokens) cudaFreeHost(enc->h_tokens);
  if (enc->h_src) cudaFreeHost(enc->h_src);
  if (enc->h_count) cudaFreeHost(enc->h_count);
  delete enc;
}

static int finish_frame(vp8gpu_encoder* enc, const std::vector<uint8_t>& bytes, int out_frame, int qi, int lf, double ssim,
                        uint8_t* out, size_t cap, size_t* size) {
  *size = bytes.size();
  if (!out || cap < bytes.size()) {
    enc->e->frame_release(out_frame);
    return enc->e->fail(VP8GPU_ERR_NOMEM, "output buffer too small");
  }
  memcpy(out, bytes.data(), bytes.size());
  if (enc->last >= 0) enc->e->frame_release(enc->last);
  enc->last = out_frame;  // Frame::copy_to: key frames and refresh_last inter frames replace LAST
  enc->has_state = true;
  enc->last_qi = qi;
  enc->last_lf = lf;      // encoder.cc:165
  enc->last_ssim = ssim;
  enc->stat_frames++;
  return VP8GPU_OK;
}

// encode at qi, choose the loop filter, serialize: Encoder::encode_raster + write_frame (encoder.cc:140-178)
static int encode_final(vp8gpu_encoder* enc, bool key, int qi, bool search_motion, uint8_t* out, size_t cap, size_t* size) {
  int frame = -1, lf = 0;
  double ssim = -1.0;
  int rc = encode_core(enc, key, qi, search_motion, &frame);
  if (rc != VP8GPU_OK) return rc;
  rc = choose_loop_filter(enc, frame, key, &lf, &ssim);
  std::vector<uint8_t> bytes;
  if (rc == VP8GPU_OK) rc = encode_bytes(enc, key, qi, lf, bytes);
  if (rc != VP8GPU_OK) {
    enc->e->frame_release(frame);
    return rc;
  }
  return finish_frame(enc, bytes, frame, qi, lf, ssim, out, cap, size);
}

int vp8gpu_encoder_encode_with_quantizer(vp8gpu_encoder* enc, const uint8_t* y, size_t y_stride, const uint8_t* u,
                                         const uint8_t* v, size_t uv_stride, int y_ac_qi, uint8_t* out, size_t cap,
                                         size_t* size) {
  if (!enc || !y || !u || !v || !size || y_ac_qi < 0 || y_ac_qi > 127) return VP8GPU_ERR_LOGIC;
  cudaSetDevice(enc->e->device());
  int rc = upload_source(enc, y, y_stride, u, v, uv_stride);
  if (rc != VP8GPU_OK) return rc;
  return encode_final(enc, !enc->has_state, y_ac_qi, true, out, cap, size);
}

int vp8gpu_encoder_encode_with_target_size(vp8gpu_encoder* enc, const uint8_t* y, size_t y_stride, const uint8_t* u,
                                           const uint8_t* v, size_t uv_stride, size_t target_size, uint8_t* out,
                                           size_t cap, size_t* size, int* chosen_qi) {
  if (!enc || !y || !u || !v || !size) return VP8GPU_ERR_LOGIC;
  cudaSetDevice(enc->e->device());
  int rc = upload_source(enc, y, y_stride, u, v, uv_stride);
  if (rc != VP8GPU_OK) return rc;
  // bisection over y_ac_qi exactly as Encoder::encode_with_target_size (encoder.cc:597-626); the size
  // of a candidate is its real size (a device pass is cheap), not the 1/16-sampled estimate of
  // estimate_frame_size; the loop filter does not change the size, so candidates skip it
  int lo = 4, hi = 127;
  if (enc->last_qi >= 0) {
    if (enc->last_qi - 16 >= lo) lo = enc->last_qi - 16;
    if (enc->last_qi + 16 < hi) hi = enc->last_qi + 16;
  }
  // Candidates run three at a time as concurrent device passes (this encoder's buffers and two helpers' on
  // their own lanes; the wavefront kernel is latency bound, so three cost about as much as one): the range
  // shrinks to a quarter per round instead of a half.  Sizes fall as the index rises, which is also what
  // the reference's bisection relies on, so the answer is the same: the smallest index that fits, or the
  // top of the range if none does.
  int best = -1;
  const int top = hi;
  std::vector<uint8_t> bytes;
  const bool key = !enc->has_state;
  bool first_round = true;
  for (int k = 0; k < 2; k++)
    if (!enc->helper[k]) {
      rc = vp8gpu_encoder_create(enc->ctx, &enc->helper[k]);
      if (rc != VP8GPU_OK) return rc;
      enc->helper[k]->owner = enc;
    }
  vp8gpu_encoder* const slot[3] = {enc, enc->helper[0], enc->helper[1]};
  while (lo <= hi) {
    const int n = hi - lo + 1;
    int q[3], nq;
    if (n <= 3) {
      nq = n;
      for (int k = 0; k < n; k++) q[k] = lo + k;
    } else {
      nq = 3;
      q[0] = lo + n / 4;
      q[1] = lo + n / 2;
      q[2] = lo + (3 * n) / 4;
    }
    int launched = 0;
    for (int k = 0; k < nq && rc == VP8GPU_OK; k++) {
      rc = encode_launch(slot[k], key, q[k], first_round && k == 0);
      if (rc == VP8GPU_OK) launched++;
    }
    first_round = false;
    size_t sz[3] = {0, 0, 0};
    for (int k = 0; k < launched; k++) {  // every launched pass is collected, also after an error
      int frame = -1;
      const int r2 = encode_collect(slot[k], &frame);
      if (r2 != VP8GPU_OK) {
        if (rc == VP8GPU_OK) rc = r2;
        continue;
      }
      enc->e->frame_release(frame);
      if (rc != VP8GPU_OK) continue;
      const int r3 = encode_bytes(slot[k], key, q[k], 0, bytes);
      if (r3 != VP8GPU_OK) rc = r3;
      sz[k] = bytes.size();
    }
    if (rc != VP8GPU_OK) return rc;
    int fit = -1;
    for (int k = 0; k < nq && fit < 0; k++)
      if (sz[k] <= target_size) fit = k;
    if (fit >= 0) {
      best = q[fit];
      hi = q[fit] - 1;
      if (fit > 0) lo = q[fit - 1] + 1;
    } else {
      lo = q[nq - 1] + 1;
    }
  }
  if (best < 0) best = top;  // encoder.cc:618: nothing fits -> the last index of the range is taken
  if (best < 0) return enc->e->fail(VP8GPU_ERR_LOGIC, "target size search failed");
  if (chosen_qi) *chosen_qi = best;
  return encode_final(enc, key, best, false, out, cap, size);  // encoder.cc:628: encode again at the chosen index
}

// Encoder::encode_with_minimum_ssim -> encode_with_quantizer_search (encoder.cc:510-557, 577-590): the
// coarsest quantiser whose reconstruction (after the loop-filter choice) still reaches minimum_ssim
int vp8gpu_encoder_encode_with_minimum_ssim(vp8gpu_encoder* enc, const uint8_t* y, size_t y_stride, const uint8_t* u,
                                            const uint8_t* v, size_t uv_stride, double minimum_ssim, uint8_t* out,
                                            size_t cap, size_t* size, int* chosen_qi) {
  if (!enc || !y || !u || !v || !size) return VP8GPU_ERR_LOGIC;
  cudaSetDevice(enc->e->device());
  int rc = upload_source(enc, y, y_stride, u, v, uv_stride);
  if (rc != VP8GPU_OK) return rc;
  const bool key = !enc->has_state;
  int lo = 0, hi = 127, best = 0;
  bool found = false, first_probe = true;
  while (lo <= hi) {
    const int qi = (lo + hi) / 2;
    int frame = -1, lf = 0;
    double ssim = -1.0;
    rc = encode_core(enc, key, qi, first_probe, &frame);
    first_probe = false;
    if (rc != VP8GPU_OK) return rc;
    rc = choose_loop_filter(enc, frame, key, &lf, &ssim);
    enc->e->frame_release(frame);
    if (rc != VP8GPU_OK) return rc;
    if (ssim >= minimum_ssim || (lo == hi && !found)) {
      found = true;
      best = qi;
    }
    if (lo == hi) break;
    if (ssim < minimum_ssim) hi = qi - 1;
    else lo = qi + 1;
  }
  if (chosen_qi) *chosen_qi = best;
  return encode_final(enc, key, best, false, out, cap, size);
}

// Encoder::estimate_frame_size (size_estimation.cc:36-99): exact instead of sampled
int vp8gpu_encoder_estimate_frame_size(vp8gpu_encoder* enc, const uint8_t* y, size_t y_stride, const uint8_t* u,
                                       const uint8_t* v, size_t uv_stride, int y_ac_qi, size_t* size) {
  if (!enc || !y || !u || !v || !size || y_ac_qi < 0 || y_ac_qi > 127) return VP8GPU_ERR_LOGIC;
  cudaSetDevice(enc->e->device());
  int rc = upload_source(enc, y, y_stride, u, v, uv_stride);
  if (rc != VP8GPU_OK) return rc;
  const bool key = !enc->has_state;
  int frame = -1;
  rc = encode_core(enc, key, y_ac_qi, true, &frame);
  if (rc != VP8GPU_OK) return rc;
  enc->e->frame_release(frame);
  std::vector<uint8_t> bytes;
  rc = encode_bytes(enc, key, y_ac_qi, 0, bytes);
  if (rc == VP8GPU_OK) *size = bytes.size();
  return rc;
}

// EncoderStats (encoder.hh:118-127) of the last frame: luma SSIM after the loop filter, the chosen
// loop-filter level and quantiser index
int vp8gpu_encoder_stats(const vp8gpu_encoder* enc, double* ssim, int* loop_filter_level, int* y_ac_qi) {
  if (!enc) return VP8GPU_ERR_LOGIC;
  if (ssim) *ssim = enc->last_ssim;
  if (loop_filter_level) *loop_filter_level = enc->last_lf;
  if (y_ac_qi) *y_ac_qi = enc->last_qi;
  return VP8GPU_OK;
}

// Encoder::export_decoder (encoder.hh:378): the reconstruction kept as LAST (one new reference for the caller)
int vp8gpu_encoder_reconstruction(vp8gpu_encoder* enc, vp8gpu_frame_id* out) {
  if (!enc || !out || enc->last < 0) return VP8GPU_ERR_LOGIC;
  const int rc = enc->e->frame_retain(enc->last);
  if (rc == VP8GPU_OK) *out = enc->last;
  return rc;
}

}  // extern "C"
