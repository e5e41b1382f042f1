/* vp8_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the VP8 decode algorithm of
 * excamera/alfalfa (src/decoder), used only as the parity checker for the CUDA path
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).  Nothing under
 * alfalfa_b200/ may include, link or call it.
 *
 * Parity status: PINNED.  The restatement reproduces all 53 golden SHA-1 vectors of the
 * reference's tests/decoding.test (tests/test_oracle_golden.py) and matches the unmodified
 * reference compiled in oracle/_ref frame by frame, before and after the loop filter.
 *
 * It is split at the same seam as the product (include/vp8gpu.h):
 *   vp8o_parse_frame    restates uncompressed_chunk.cc + DecoderState::parse_and_apply and
 *                       emits the flat records of vp8gpu.h
 *   vp8o_reconstruct    restates Frame::decode  (prediction, dequant, IWHT, IDCT)
 *   vp8o_loopfilter     restates Frame::loopfilter
 * so each half of the product can be checked against its own half of the oracle.
 */
#ifndef VP8_ORACLE_H
#define VP8_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#include "../include/vp8gpu.h" /* record layouts only (types, no functions) */

#ifdef __cplusplus
extern "C" {
#endif

/* MB-aligned planar raster, tight strides (VP8Raster, vp8_raster.hh:53). */
typedef struct vp8o_raster {
  int w16, h16; /* luma plane size = 16*mb_cols x 16*mb_rows */
  uint8_t *y, *u, *v;
} vp8o_raster;

vp8o_raster* vp8o_raster_new(int width, int height);
void vp8o_raster_free(vp8o_raster* r);
void vp8o_raster_copy(vp8o_raster* dst, const vp8o_raster* src);

/* DecoderState (decoder.hh:190-225) */
typedef struct vp8o_state vp8o_state;
vp8o_state* vp8o_state_new(int width, int height);
void vp8o_state_free(vp8o_state* s);

/* growable parsed-frame buffers */
typedef struct vp8o_parsed {
  vp8gpu_frame_desc desc;
  vp8gpu_mb* mbs;
  vp8gpu_token* tokens;
  vp8gpu_split_mvs* split;
  size_t mbs_cap, tokens_cap, split_cap;
} vp8o_parsed;

vp8o_parsed* vp8o_parsed_new(void);
void vp8o_parsed_free(vp8o_parsed* p);

/* returns VP8GPU_OK / VP8GPU_ERR_* */
int vp8o_parse_frame(vp8o_state* st, const uint8_t* data, size_t len, vp8o_parsed* out);

void vp8o_reconstruct(const vp8gpu_frame_desc* desc, const vp8gpu_mb* mbs,
                      const vp8gpu_token* tokens, const vp8gpu_split_mvs* split,
                      const vp8o_raster* last, const vp8o_raster* golden, const vp8o_raster* alt,
                      vp8o_raster* out);
void vp8o_loopfilter(const vp8gpu_frame_desc* desc, const vp8gpu_mb* mbs, vp8o_raster* out);

/* Decoder (decoder.hh:244-300) */
typedef struct vp8o_decoder vp8o_decoder;
vp8o_decoder* vp8o_decoder_new(int width, int height);
void vp8o_decoder_free(vp8o_decoder* d);
/* Decodes one compressed frame.  *shown = show_frame.  The returned raster (owned by the
 * decoder, valid until the next call) is the loop-filtered output; if pre_lf is not NULL
 * it receives a copy of the raster before the loop filter. */
int vp8o_decoder_decode(vp8o_decoder* d, const uint8_t* data, size_t len, int* shown,
                        const vp8o_raster** out, vp8o_raster* pre_lf);
const vp8o_parsed* vp8o_decoder_last_parsed(const vp8o_decoder* d);
/* reference rasters: 0 last, 1 golden, 2 alternative */
const vp8o_raster* vp8o_decoder_ref(const vp8o_decoder* d, int which);

/* BaseRaster::dump (util/raster.cc:85-114): display rectangle, packed planar. Returns
 * bytes written. */
size_t vp8o_raster_dump_display(const vp8o_raster* r, int width, int height, uint8_t* dst);

/* Time `reps` full decodes of an in-memory IVF (for bench.py's cpu_baseline "port").
 * phase_s[3] = seconds in parse / reconstruct / loopfilter of the best repetition. */
int vp8o_time_ivf(const uint8_t* ivf, size_t len, int reps, uint32_t max_frames, double phase_s[3],
                  uint32_t* frames);

#ifdef __cplusplus
}
#endif
#endif
