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
// loop-filter strength from the quantiser (libvpx's initial guess; the reference searches by SSIM)
int default_filter_level(int qi) {
  int l = qi * 3 / 8;
  return l > 63 ? 63 : l;
}
#define CUE(call)                                                       \
  do {                                                                  \
    cudaError_t e__ = (call);                                           \
    if (e__ != cudaSuccess) return enc->e->cuda_fail(e__, #call);       \
  } while (0)

// run the device pipeline for one candidate quantiser; returns the compressed frame in `bytes` and
// the new reconstruction in *out_frame (caller releases it or keeps it as LAST)
int run_encode(vp8gpu_encoder* enc, bool key, int qi, bool search_motion, std::vector<uint8_t>& bytes, int* out_frame) {
  Engine* e = enc->e;
  const vp8::Geom& g = e->geom();
  const size_t n_mbs = (size_t)g.mb_cols * g.mb_rows;
  cudaStream_t s = e->stream(enc->lane);
  int out = -1;
  int rc = e->frame_alloc(&out);
  if (rc != VP8GPU_OK) return rc;
  int ids[3] = {enc->src, out, enc->last};
  rc = e->acquire_frames(enc->lane, ids, key ? 2 : 3);
  if (rc != VP8GPU_OK) {
    e->frame_release(out);
    return rc;
  }
  const int lf_level = default_filter_level(qi);
  vp8::EncJob* ej = reinterpret_cast<vp8::EncJob*>(enc->h_hdr);
  vp8::DevJob* dj = reinterpret_cast<vp8::DevJob*>(enc->h_hdr + 512);
  memset(enc->h_hdr, 0, 1024);
  int* d_sync = reinterpret_cast<int*>(enc->dev + enc->off_sync);
  ej->src = e->frame_dev(enc->src);
  ej->ref = key ? nullptr : e->frame_dev(enc->last);
  ej->out = e->frame_dev(out);
  ej->mbs = reinterpret_cast<vp8gpu_mb*>(enc->dev + enc->off_mbs);
  ej->tokens = reinterpret_cast<vp8gpu_token*>(enc->dev + enc->off_tokens);
  ej->tok_counter = reinterpret_cast<uint32_t*>(d_sync + 96);
  ej->tok_cap = enc->tok_cap;
  ej->mv = reinterpret_cast<int*>(enc->dev + enc->off_mv);
  ej->sad = reinterpret_cast<uint32_t*>(enc->dev + enc->off_sad);
  ej->progress = d_sync + 128;
  ej->q = make_quant(qi);
  ej->key_frame = key;
  ej->lf_level = (uint8_t)lf_level;
  dj->mbs = ej->mbs;
  dj->tokens = ej->tokens;
  dj->split = nullptr;
  dj->out = ej->out;
  dj->intra_progress = d_sync + 128;
  dj->lf_progress = d_sync + 128 + g.mb_rows;
  for (int i = 0; i < 4; i++) dj->quant[i] = ej->q;
  dj->key_frame = key;
  dj->sharpness = 0;
  dj->lf_enabled = lf_level > 0;
  auto fail = [&](int code) {
    e->frame_release(out);
    return code;
  };
#define CUF(call)                                                        \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) return fail(e->cuda_fail(e__, #call));       \
  } while (0)
  CUF(cudaMemcpyAsync(enc->dev + enc->off_encjob, enc->h_hdr, 1024, cudaMemcpyHostToDevice, s));
  CUF(cudaMemsetAsync(d_sync, 0, sizeof(int) * (128 + 2 * (size_t)g.mb_rows), s));
  const vp8::EncJob* d_ej = reinterpret_cast<const vp8::EncJob*>(enc->dev + enc->off_encjob);
  const vp8::DevJob* d_dj = reinterpret_cast<const vp8::DevJob*>(enc->dev + enc->off_encjob + 512);
  int launches = 0;
  if (!key && search_motion) {  // vectors do not depend on the quantiser: searched once per source frame
    if (int ce = vp8::launch_enc_motion(d_ej, g, s)) return fail(e->cuda_fail((cudaError_t)ce, "k_enc_motion"));
    launches++;
  }
  if (int ce = vp8::launch_enc_mb(d_ej, g, d_sync + 0, s)) return fail(e->cuda_fail((cudaError_t)ce, "k_enc_mb"));
  launches++;
  if (lf_level > 0) {
    if (int ce = vp8::launch_loopfilter(d_dj, 1, g, d_sync + 32, s)) return fail(e->cuda_fail((cudaError_t)ce, "k_loopfilter"));
    launches++;
  }
  e->count_launches(launches);
  e->mark_frames(enc->lane, ids, key ? 2 : 3);
  // results back: token count first, then the records
  CUF(cudaMemcpyAsync(enc->h_count, ej->tok_counter, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUF(cudaMemcpyAsync(enc->h_mbs, ej->mbs, n_mbs * sizeof(vp8gpu_mb), cudaMemcpyDeviceToHost, s));
  CUF(cudaStreamSynchronize(s));
  const uint32_t n_tok = *enc->h_count;
  if (n_tok > enc->tok_cap) return fail(e->fail(VP8GPU_ERR_NOMEM, "encoder token pool overflow"));
  if (n_tok) {
    CUF(cudaMemcpyAsync(enc->h_tokens, ej->tokens, (size_t)n_tok * sizeof(vp8gpu_token), cudaMemcpyDeviceToHost, s));
    CUF(cudaStreamSynchronize(s));
  }
  vp8::EncodeHeader h;
  h.key_frame = key;
  h.show_frame = true;
  h.width = e->width();
  h.height = e->height();
  h.y_ac_qi = qi;
  h.loop_filter_level = lf_level;
  h.sharpness = 0;
  h.optimize_token_probs = true;
  bytes = vp8::serialize_frame(h, enc->h_mbs, enc->h_tokens, nullptr);
  if (bytes.empty()) return fail(e->fail(VP8GPU_ERR_LOGIC, "serializer rejected the device records"));
  *out_frame = out;
  return VP8GPU_OK;
#undef CUF
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

static int finish_frame(vp8gpu_encoder* enc, const std::vector<uint8_t>& bytes, int out_frame, int qi, uint8_t* out,
                        size_t cap, size_t* size) {
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
  enc->stat_frames++;
  return VP8GPU_OK;
}

int vp8gpu_encoder_encode_with_quantizer(vp8gpu_encoder* enc, const uint8_t* y, size_t y_stride, const uint8_t* u,
                                         const uint8_t* v, size_t uv_stride, int y_ac_qi, uint8_t* out, size_t cap,
                                         size_t* size) {
  if (!enc || !y || !u || !v || !size || y_ac_qi < 0 || y_ac_qi > 127) return VP8GPU_ERR_LOGIC;
  cudaSetDevice(enc->e->device());
  int rc = upload_source(enc, y, y_stride, u, v, uv_stride);
  if (rc != VP8GPU_OK) return rc;
  std::vector<uint8_t> bytes;
  int frame = -1;
  rc = run_encode(enc, !enc->has_state, y_ac_qi, true, bytes, &frame);
  if (rc != VP8GPU_OK) return rc;
  return finish_frame(enc, bytes, frame, y_ac_qi, out, cap, size);
}

int vp8gpu_encoder_encode_with_target_size(vp8gpu_encoder* enc, const uint8_t* y, size_t y_stride, const uint8_t* u,
                                           const uint8_t* v, size_t uv_stride, size_t target_size, uint8_t* out,
                                           size_t cap, size_t* size, int* chosen_qi) {
  if (!enc || !y || !u || !v || !size) return VP8GPU_ERR_LOGIC;
  cudaSetDevice(enc->e->device());
  int rc = upload_source(enc, y, y_stride, u, v, uv_stride);
  if (rc != VP8GPU_OK) return rc;
  // bisection over y_ac_qi exactly as Encoder::encode_with_target_size (encoder.cc:597-626); the size
  // of a candidate is its real size (the device encode is cheap), not the 1/16-sampled estimate
  int lo = 4, hi = 127;
  if (enc->last_qi >= 0) {
    if (enc->last_qi - 16 >= lo) lo = enc->last_qi - 16;
    if (enc->last_qi + 16 < hi) hi = enc->last_qi + 16;
  }
  int best = -1, best_frame = -1;
  std::vector<uint8_t> best_bytes, bytes;
  const bool key = !enc->has_state;
  bool first_probe = true;
  while (lo <= hi) {
    const int qi = (lo + hi) / 2;
    int frame = -1;
    rc = run_encode(enc, key, qi, first_probe, bytes, &frame);
    first_probe = false;
    if (rc != VP8GPU_OK) {
      if (best_frame >= 0) enc->e->frame_release(best_frame);
      return rc;
    }
    if (bytes.size() <= target_size || (lo == hi && best < 0)) {
      if (best_frame >= 0) enc->e->frame_release(best_frame);
      best = qi;
      best_frame = frame;
      best_bytes.swap(bytes);
      hi = qi - 1;
    } else {
      enc->e->frame_release(frame);
      lo = qi + 1;
    }
  }
  if (best < 0) return enc->e->fail(VP8GPU_ERR_LOGIC, "target size search failed");
  if (chosen_qi) *chosen_qi = best;
  return finish_frame(enc, best_bytes, best_frame, best, out, cap, size);
}

// Encoder::export_decoder (encoder.hh:378): the reconstruction kept as LAST (one new reference for the caller)
int vp8gpu_encoder_reconstruction(vp8gpu_encoder* enc, vp8gpu_frame_id* out) {
  if (!enc || !out || enc->last < 0) return VP8GPU_ERR_LOGIC;
  const int rc = enc->e->frame_retain(enc->last);
  if (rc == VP8GPU_OK) *out = enc->last;
  return rc;
}

}  // extern "C"
