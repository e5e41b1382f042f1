// kernels.cu -- sm_100a kernels of the VP8 pixel pipeline.
//
// Three kernels per frame batch, all "one warp per macroblock (row)":
//   k_inter        every inter-coded macroblock is independent (the reference frame is frozen):
//                  one warp per macroblock, six-tap motion compensation + token expansion +
//                  dequant + IWHT + IDCT + add, fully parallel.         (macroblock.cc:553-601)
//   k_intra        intra macroblocks read unfiltered pixels of their left / above / above-right
//                  neighbours in the same frame: one warp per macroblock ROW sweeping left to
//                  right, rows chained by progress counters in HBM (2-macroblock lag).
//                                                                       (macroblock.cc:523-551)
//   k_loopfilter   same wavefront shape for the in-loop deblocking filter: lanes 0-15 filter
//                  the 16 luma positions of an edge, lanes 16-23 / 24-31 the 8 U / 8 V ones.
//                                                    (loopfilter.cc:133-154, frame.cc:139-182)
// Arithmetic lives in vp8_math.cuh (shared with the CPU unit tests); this file is data movement.
//
// Memory-model notes for the wavefront kernels: pixels written by another warp are read with
// ld.global.cg (L2, never a stale L1 line); a finished macroblock is published with
// __syncwarp() (orders every lane's stores before lane 0) and one st.release.gpu by lane 0; the
// consumer polls with ld.acquire.gpu on lane 0 and __syncwarp()s before the other lanes read.
#include <cuda_runtime.h>
#include <stdint.h>

#include "engine.h"
#include "vp8_math.cuh"

namespace vp8 {
namespace {

#define VP8_LUT_QUALIFIER __device__ const
#include "bpred_lut.inc"

// sixtap_filters (prediction.cc:645-653)
__constant__ int16_t c_sixtap[8][6] = {{0, 0, 128, 0, 0, 0},     {0, -6, 123, 12, -1, 0}, {2, -11, 108, 36, -8, 1},
                                      {0, -9, 93, 50, -6, 0},   {3, -16, 77, 77, -16, 3}, {0, -6, 50, 93, -9, 0},
                                      {1, -8, 36, 108, -11, 2}, {0, -1, 12, 123, -6, 0}};

// Optional phase profiling of the wavefront kernels (build with -DVP8_PROFILE, tools/phase_profile.py):
// lane 0 of every warp accumulates clock64() deltas per phase and adds them to g_prof at exit.
#ifdef VP8_PROFILE
__device__ unsigned long long g_prof[32];
#define PROF_DECL unsigned long long prof_t = clock64(), prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; unsigned prof_n = 0
#define PROF(i)                                  \
  do {                                           \
    const unsigned long long now_ = clock64();   \
    prof_acc[i] += now_ - prof_t;                \
    prof_t = now_;                               \
  } while (0)
#define PROF_COUNT() (prof_n++)
#define PROF_FLUSH(base)                                                           \
  do {                                                                             \
    if (lane == 0) {                                                               \
      for (int i_ = 0; i_ < 8; i_++) atomicAdd(&g_prof[(base) + i_], prof_acc[i_]); \
      atomicAdd(&g_prof[(base) + 8], (unsigned long long)prof_n);                  \
    }                                                                              \
  } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define PROF_COUNT()
#define PROF_FLUSH(base)
#endif

constexpr int CS = 18;          // int16 stride of one 4x4 coefficient block in shared memory (bank spread)
constexpr int COEF_WORDS = 25 * CS / 2;  // 225 32-bit words

struct MbFields {
  uint32_t tok_off, tok_cnt;
  int y_mode, uv_mode, ref, segment, lf_level, flags;
  int mv_x, mv_y;
  uint32_t split_idx;
  uint32_t bm_lo, bm_hi;
};

__device__ __forceinline__ MbFields load_mb(const vp8gpu_mb* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  const uint4 a = __ldg(q), b = __ldg(q + 1);
  MbFields f;
  f.tok_off = a.x;
  f.tok_cnt = a.y & 0xFFFF;
  f.y_mode = (a.y >> 16) & 0xFF;
  f.uv_mode = a.y >> 24;
  f.ref = a.z & 0xFF;
  f.segment = (a.z >> 8) & 0xFF;
  f.lf_level = (a.z >> 16) & 0xFF;
  f.flags = a.z >> 24;
  f.mv_x = (int16_t)(a.w & 0xFFFF);
  f.mv_y = (int16_t)(a.w >> 16);
  f.split_idx = b.x;
  f.bm_lo = b.z;
  f.bm_hi = b.w;
  return f;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// MotionVector::luma_to_chroma (macroblock.cc:289-299) on the int16 sum of four luma components
__device__ __forceinline__ int chroma_component(int sum) {
  const int s = (int16_t)sum;
  return s >= 0 ? (s + 4) >> 3 : -((-s + 4) >> 3);
}

// ------------------------------------------------------------------------------------------------
// Motion compensation of one N x N block by the whole warp (prediction.cc:655-674, 919-971).
// (x0, y0) = block origin in the plane, mv in 1/8 pel.
//  * the source window ((N+5)^2, or N^2 for whole-pel vectors) is staged in shared memory; when it
//    lies inside the plane it is fetched as aligned 32-bit words, otherwise pixel by pixel with
//    clamped coordinates (EdgeExtendedRaster::at, vp8_raster.hh:327-338) -- no padded planes needed;
//  * each lane filters strips of 4 outputs from a 9-pixel run (9 loads, 24 MACs);
//  * a pass whose fraction is 0 is skipped: its taps are {0,0,128,0,0,0}, (128 p + 64) >> 7 = p.
// ------------------------------------------------------------------------------------------------
template <int N>
struct Mc {
  static constexpr int NW = (N + 11) / 4;  // words per staged row: covers 3 + N + 5 bytes
  static constexpr int TS = 4 * NW;        // staged row stride in bytes (24 / 16 / 12)
};

// returns the byte offset of window column 0 inside a staged row
template <int N>
__device__ __forceinline__ int load_window(const uint8_t* __restrict__ ref, int pitch, int PW, int PH, int wx, int wy,
                                           int wcols, int wrows, uint8_t* tile, int lane) {
  constexpr int NW = Mc<N>::NW, TS = Mc<N>::TS;
  if (wx >= 0 && wy >= 0 && wx + wcols <= PW && wy + wrows <= PH) {
    const int o = wx & 3;
    const uint8_t* base = ref + (size_t)wy * pitch + (wx - o);
    uint32_t* tw = reinterpret_cast<uint32_t*>(tile);
    for (int i = lane; i < wrows * NW; i += 32) {
      const int r = i / NW, w = i - r * NW;
      tw[i] = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)r * pitch) + w);
    }
    return o;
  }
  for (int i = lane; i < wrows * wcols; i += 32) {
    const int r = i / wcols, c = i - r * wcols;
    tile[r * TS + c] = __ldg(ref + (size_t)clampi(wy + r, 0, PH - 1) * pitch + clampi(wx + c, 0, PW - 1));
  }
  return 0;
}

// horizontal 6-tap over `nrows` staged rows: out[r][c] from win[r][c .. c+5]; 4 outputs per item
template <int N>
__device__ __forceinline__ void hpass(const uint8_t* win, int nrows, const int16_t* hf, uint8_t* out, int ostride,
                                      int lane) {
  constexpr int G = N / 4, TS = Mc<N>::TS;
  for (int i = lane; i < nrows * G; i += 32) {
    const int r = i / G, g = i - r * G;
    const uint8_t* t = win + r * TS + 4 * g;
    int p[9];
#pragma unroll
    for (int k = 0; k < 9; k++) p[k] = t[k];
    uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
      o |= (uint32_t)vp8m::sixtap(p[j], p[j + 1], p[j + 2], p[j + 3], p[j + 4], p[j + 5], hf) << (8 * j);
    *reinterpret_cast<uint32_t*>(out + r * ostride + 4 * g) = o;
  }
}
// vertical 6-tap: dst[r][c] from src[r .. r+5][c]; each item = one column, 4 rows
template <int N>
__device__ __forceinline__ void vpass(const uint8_t* src, int sstride, const int16_t* vf, uint8_t* dst, int dstride,
                                      int lane) {
  constexpr int G = N / 4;
  for (int i = lane; i < N * G; i += 32) {
    const int g = i / N, c = i - g * N;
    const uint8_t* m = src + (4 * g) * sstride + c;
    int p[9];
#pragma unroll
    for (int k = 0; k < 9; k++) p[k] = m[k * sstride];
#pragma unroll
    for (int j = 0; j < 4; j++)
      dst[(4 * g + j) * dstride + c] = (uint8_t)vp8m::sixtap(p[j], p[j + 1], p[j + 2], p[j + 3], p[j + 4], p[j + 5], vf);
  }
}

template <int N>
__device__ __forceinline__ void mc_block(const uint8_t* __restrict__ ref, int pitch, int PW, int PH, int x0, int y0,
                                         int mvx, int mvy, uint8_t* dst, int dstride, uint8_t* tile, uint8_t* mid,
                                         int lane) {
  constexpr int TS = Mc<N>::TS, G = N / 4;
  const int sx = x0 + (mvx >> 3), sy = y0 + (mvy >> 3);
  const int mx = mvx & 7, my = mvy & 7;
  if ((mx | my) == 0) {
    const int o = load_window<N>(ref, pitch, PW, PH, sx, sy, N, N, tile, lane);
    __syncwarp();
    for (int i = lane; i < N * G; i += 32) {
      const int r = i / G, g = i - r * G;
      const uint8_t* t = tile + r * TS + o + 4 * g;
      *reinterpret_cast<uint32_t*>(dst + r * dstride + 4 * g) =
          (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
    }
    __syncwarp();
    return;
  }
  const int o = load_window<N>(ref, pitch, PW, PH, sx - 2, sy - 2, N + 5, N + 5, tile, lane);
  __syncwarp();
  const uint8_t* win = tile + o;
  if (mx && my) {
    hpass<N>(win, N + 5, c_sixtap[mx], mid, N, lane);
    __syncwarp();
    vpass<N>(mid, N, c_sixtap[my], dst, dstride, lane);
  } else if (mx) {
    hpass<N>(win + 2 * TS, N, c_sixtap[mx], dst, dstride, lane);
  } else {
    vpass<N>(win + 2, TS, c_sixtap[my], dst, dstride, lane);
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// GPU back end of the entropy decoder: expand the macroblock's token list into dequantised
// coefficient blocks (quantization.cc:95-126, int16 wrap), run the inverse WHT (transform.cc:47-88)
// and the inverse DCT (transform.cc:100-137).  On return coef[blk*CS + y*4 + x] holds the
// RESIDUAL of pixel (x, y) of block blk (0-15 Y, 16-19 U, 20-23 V).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void build_residuals(const DevJob& J, const MbFields& f, int16_t* coef, int lane) {
  uint32_t* w = reinterpret_cast<uint32_t*>(coef);
  for (int i = lane; i < COEF_WORDS; i += 32) w[i] = 0;
  __syncwarp();
  const vp8gpu_quant q = J.quant[f.segment];
  const vp8gpu_token* tok = J.tokens + f.tok_off;
  for (uint32_t t = lane; t < f.tok_cnt; t += 32) {
    const uint32_t v = __ldg(tok + t);
    const int blk = (v >> 20) & 31, pos = (v >> 16) & 15;
    const int val = (int16_t)(v & 0xFFFF);
    int factor;
    if (blk < 16) factor = pos ? q.y_ac : q.y_dc;
    else if (blk < 24) factor = pos ? q.uv_ac : q.uv_dc;
    else factor = pos ? q.y2_ac : q.y2_dc;
    coef[blk * CS + pos] = (int16_t)(val * factor);
  }
  __syncwarp();
  if (f.flags & VP8GPU_MB_HAS_Y2) {
    if (lane == 0) {
      int16_t dc[16];
      vp8m::iwht16(coef + 24 * CS, dc);
#pragma unroll
      for (int k = 0; k < 16; k++) coef[k * CS] = dc[k];
    }
    __syncwarp();
  }
  if (lane < 24) {
    int16_t* c = coef + lane * CS;
    const uint32_t* cw = reinterpret_cast<const uint32_t*>(c);
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) any |= cw[k];
    if (any) {
      int16_t in[16], r[16];
#pragma unroll
      for (int k = 0; k < 16; k++) in[k] = c[k];
      vp8m::idct16(in, r);
#pragma unroll
      for (int k = 0; k < 16; k++) c[k] = r[k];
    }
  }
  __syncwarp();
}

// pixel = clamp255(prediction + residual) over the whole macroblock buffer
// pix layout: Y 16x16 at 0, U 8x8 at 256, V 8x8 at 320.
__device__ __forceinline__ void add_residuals(uint8_t* pix, const int16_t* coef, int lane) {
  for (int g4 = lane; g4 < 96; g4 += 32) {
    int blk, ry, off;
    if (g4 < 64) {
      const int y = g4 >> 2, x4 = (g4 & 3) * 4;
      blk = (y >> 2) * 4 + (x4 >> 2);
      ry = y & 3;
      off = y * 16 + x4;
    } else {
      const int c = g4 - 64, plane = c >> 4, cc = c & 15;
      const int y = cc >> 1, x4 = (cc & 1) * 4;
      blk = 16 + plane * 4 + (y >> 2) * 2 + (x4 >> 2);
      ry = y & 3;
      off = 256 + plane * 64 + y * 8 + x4;
    }
    const int16_t* r = coef + blk * CS + ry * 4;
    uint32_t p = *reinterpret_cast<uint32_t*>(pix + off);
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) o |= (uint32_t)vp8m::clamp255((int)((p >> (8 * k)) & 0xFF) + r[k]) << (8 * k);
    *reinterpret_cast<uint32_t*>(pix + off) = o;
  }
  __syncwarp();
}

// macroblock buffer -> frame: 16-byte luma rows by lanes 0-15, 8-byte chroma rows by 16-31
__device__ __forceinline__ void store_mb(const uint8_t* pix, uint8_t* frame, const Geom& g, int col, int row,
                                         int lane) {
  if (lane < 16) {
    const uint4 v = *reinterpret_cast<const uint4*>(pix + lane * 16);
    *reinterpret_cast<uint4*>(frame + (size_t)(16 * row + lane) * g.y_pitch + 16 * col) = v;
  } else {
    const int plane = (lane - 16) >> 3, y = (lane - 16) & 7;
    const uint2 v = *reinterpret_cast<const uint2*>(pix + 256 + plane * 64 + y * 8);
    uint8_t* base = frame + (plane ? g.v_off : g.u_off);
    *reinterpret_cast<uint2*>(base + (size_t)(8 * row + y) * g.c_pitch + 8 * col) = v;
  }
}

// ================================================================================================
// k_inter
// ================================================================================================
constexpr int INTER_WARPS = 4;

__global__ void __launch_bounds__(INTER_WARPS * 32) k_inter(const DevJob* __restrict__ jobs, Geom g) {
  __shared__ __align__(16) uint8_t s_pix[INTER_WARPS][384];
  __shared__ __align__(16) int16_t s_coef[INTER_WARPS][25 * CS];
  __shared__ __align__(16) uint8_t s_tile[INTER_WARPS][21 * 24];
  __shared__ __align__(16) uint8_t s_mid[INTER_WARPS][21 * 16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const DevJob& J = jobs[blockIdx.y];
  const int mbi = blockIdx.x * INTER_WARPS + warp;
  if (mbi >= g.mb_cols * g.mb_rows) return;
  const MbFields f = load_mb(J.mbs + mbi);
  if (f.ref == VP8GPU_REF_CURRENT) return;  // intra macroblocks belong to k_intra
  const int row = mbi / g.mb_cols, col = mbi - row * g.mb_cols;
  uint8_t* pix = s_pix[warp];
  uint8_t* tile = s_tile[warp];
  uint8_t* mid = s_mid[warp];
  const uint8_t* ref = J.ref[f.ref - 1];
  const uint8_t* refU = ref + g.u_off;
  const uint8_t* refV = ref + g.v_off;
  const int CW = g.W >> 1, CH = g.H >> 1;

  if (f.y_mode == VP8GPU_SPLITMV) {
    // lane i < 16 holds the vector of luma sub-block i
    int my_x = 0, my_y = 0;
    if (lane < 16) {
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(J.split + f.split_idx) + lane);
      my_x = (int16_t)(v & 0xFFFF);
      my_y = (int16_t)(v >> 16);
    }
    for (int b = 0; b < 16; b++) {
      const int bx = b & 3, by = b >> 2;
      const int mvx = __shfl_sync(0xffffffffu, my_x, b), mvy = __shfl_sync(0xffffffffu, my_y, b);
      mc_block<4>(ref, g.y_pitch, g.W, g.H, 16 * col + 4 * bx, 16 * row + 4 * by, mvx, mvy, pix + by * 64 + bx * 4, 16,
                  tile, mid, lane);
    }
    for (int b = 0; b < 4; b++) {
      const int cx = b & 1, cy = b >> 1, a = cy * 8 + cx * 2;
      const int sx = __shfl_sync(0xffffffffu, my_x, a) + __shfl_sync(0xffffffffu, my_x, a + 1) +
                     __shfl_sync(0xffffffffu, my_x, a + 4) + __shfl_sync(0xffffffffu, my_x, a + 5);
      const int sy = __shfl_sync(0xffffffffu, my_y, a) + __shfl_sync(0xffffffffu, my_y, a + 1) +
                     __shfl_sync(0xffffffffu, my_y, a + 4) + __shfl_sync(0xffffffffu, my_y, a + 5);
      const int cmx = chroma_component(sx), cmy = chroma_component(sy);
      mc_block<4>(refU, g.c_pitch, CW, CH, 8 * col + 4 * cx, 8 * row + 4 * cy, cmx, cmy, pix + 256 + cy * 32 + cx * 4, 8,
                  tile, mid, lane);
      mc_block<4>(refV, g.c_pitch, CW, CH, 8 * col + 4 * cx, 8 * row + 4 * cy, cmx, cmy, pix + 320 + cy * 32 + cx * 4, 8,
                  tile, mid, lane);
    }
  } else {
    mc_block<16>(ref, g.y_pitch, g.W, g.H, 16 * col, 16 * row, f.mv_x, f.mv_y, pix, 16, tile, mid, lane);
    const int cmx = chroma_component(4 * f.mv_x), cmy = chroma_component(4 * f.mv_y);
    mc_block<8>(refU, g.c_pitch, CW, CH, 8 * col, 8 * row, cmx, cmy, pix + 256, 8, tile, mid, lane);
    mc_block<8>(refV, g.c_pitch, CW, CH, 8 * col, 8 * row, cmx, cmy, pix + 320, 8, tile, mid, lane);
  }

  if (f.tok_cnt) {  // Macroblock::has_nonzero_ (macroblock.cc:579,593)
    build_residuals(J, f, s_coef[warp], lane);
    add_residuals(pix, s_coef[warp], lane);
  }
  store_mb(pix, J.out, g, col, row, lane);
}

// ================================================================================================
// wavefront plumbing
// ================================================================================================
__device__ __forceinline__ int ld_progress(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_progress(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until the row above has finished every macroblock left of `need`
__device__ __forceinline__ void wait_row(const int* progress_above, int need, int lane) {
  if (lane == 0) {
    while (ld_progress(progress_above) < need) __nanosleep(20);
  }
  __syncwarp();  // the other lanes' later loads are ordered after lane 0's acquire through this barrier
}
__device__ __forceinline__ void publish_row(int* progress, int value, int lane) {
  __syncwarp();  // every lane's stores happen-before lane 0's release store (cumulative at gpu scope)
  if (lane == 0) st_progress(progress, value);
}
__device__ __forceinline__ uint8_t ldcg_u8(const uint8_t* p) { return __ldcg(p); }

// next set bit at position >= from in a bitmask spread one 32-bit word per lane; -1 if none
__device__ __forceinline__ int next_marked(uint32_t my_word, int from, int nwords) {
  for (int w = from >> 5; w < nwords; w++) {
    uint32_t bits = __shfl_sync(0xffffffffu, my_word, w);
    if (w == (from >> 5)) bits &= 0xffffffffu << (from & 31);
    if (bits) return w * 32 + __ffs(bits) - 1;
  }
  return -1;
}

// ================================================================================================
// k_intra
// ================================================================================================
__global__ void __launch_bounds__(32) k_intra(const DevJob* __restrict__ jobs, int njobs, Geom g, int* ticket) {
  __shared__ __align__(16) uint8_t pix[384];
  __shared__ __align__(16) int16_t coef[25 * CS];
  __shared__ uint8_t aboveY[24];  // [0] = above-left, [1..16] = above, [17..20] = above-right
  __shared__ uint8_t leftY[16];
  __shared__ uint8_t aboveC[2][12];  // [0] = above-left, [1..8] = above
  __shared__ uint8_t leftC[2][8];
  __shared__ uint8_t edge[16];  // 13-entry edge vector of the current 4x4 sub-block
  const int lane = threadIdx.x;
  int t = 0;
  if (lane == 0) t = atomicAdd(ticket, 1);
  t = __shfl_sync(0xffffffffu, t, 0);
  const int row = t / njobs, job = t - row * njobs;
  if (row >= g.mb_rows) return;
  const DevJob& J = jobs[job];
  if (J.n_intra == 0) return;
  const int cols = g.mb_cols;
  const vp8gpu_mb* row_mbs = J.mbs + (size_t)row * cols;

  // which macroblocks of this row are intra-coded: bit c of a mask spread one word per lane
  const int nwords = (cols + 31) >> 5;
  uint32_t my_word = 0;
  for (int w = 0; w < nwords; w++) {
    const int c = w * 32 + lane;
    const bool intra = c < cols && (__ldg(reinterpret_cast<const uint32_t*>(row_mbs + c) + 2) & 0xFF) == VP8GPU_REF_CURRENT;
    const uint32_t bits = __ballot_sync(0xffffffffu, intra);
    if (lane == w) my_word = bits;
  }
  int col = next_marked(my_word, 0, nwords);
  int* progress = J.intra_progress + row;
  // progress = P means: every macroblock of this row with column < P is reconstructed
  publish_row(progress, col < 0 ? cols : col, lane);

  uint8_t* const Y = J.out;
  uint8_t* const U = J.out + g.u_off;
  uint8_t* const V = J.out + g.v_off;

  PROF_DECL;
  while (col >= 0) {
    PROF(7);
    const MbFields f = load_mb(row_mbs + col);
    // the residual only depends on this macroblock's tokens: build it before waiting on the row above
    const bool has_res = f.tok_cnt != 0;
    PROF(0);
    if (has_res) build_residuals(J, f, coef, lane);
    PROF(1);
    if (row > 0) wait_row(progress - 1, min(col + 2, cols), lane);
    PROF(2);

    // ---- edges (prediction.cc:99-167), read through L2 ----
    {
      // luma above row incl. corner and above-right: 21 entries
      if (lane < 21) {
        int v;
        if (row == 0) v = 127;
        else if (lane == 0) v = col > 0 ? ldcg_u8(Y + (size_t)(16 * row - 1) * g.y_pitch + 16 * col - 1) : 129;
        else if (lane <= 16) v = ldcg_u8(Y + (size_t)(16 * row - 1) * g.y_pitch + 16 * col + lane - 1);
        else if (col == cols - 1) v = ldcg_u8(Y + (size_t)(16 * row - 1) * g.y_pitch + 16 * col + 15);
        else v = ldcg_u8(Y + (size_t)(16 * row - 1) * g.y_pitch + 16 * col + lane - 1);
        aboveY[lane] = (uint8_t)v;
      }
      if (lane < 16) leftY[lane] = col > 0 ? ldcg_u8(Y + (size_t)(16 * row + lane) * g.y_pitch + 16 * col - 1) : 129;
      if (lane < 18) {
        const int plane = lane / 9, k = lane % 9;
        const uint8_t* P = plane ? V : U;
        int v;
        if (row == 0) v = 127;
        else if (k == 0) v = col > 0 ? ldcg_u8(P + (size_t)(8 * row - 1) * g.c_pitch + 8 * col - 1) : 129;
        else v = ldcg_u8(P + (size_t)(8 * row - 1) * g.c_pitch + 8 * col + k - 1);
        aboveC[plane][k] = (uint8_t)v;
      }
      if (lane >= 16) {
        const int plane = (lane - 16) >> 3, k = (lane - 16) & 7;
        const uint8_t* P = plane ? V : U;
        leftC[plane][k] = col > 0 ? ldcg_u8(P + (size_t)(8 * row + k) * g.c_pitch + 8 * col - 1) : 129;
      }
    }
    __syncwarp();
    PROF(3);

    // ---- chroma 8x8 prediction (prediction.cc:435-449): 128 pixels, 4 per lane ----
    int cdc[2] = {128, 128};
    if (f.uv_mode == VP8GPU_DC_PRED) {
#pragma unroll
      for (int plane = 0; plane < 2; plane++) {
        int s = 0, n = 0;
        if (row > 0) { for (int k = 0; k < 8; k++) s += aboveC[plane][1 + k]; n += 8; }
        if (col > 0) { for (int k = 0; k < 8; k++) s += leftC[plane][k]; n += 8; }
        cdc[plane] = n == 16 ? (s + 8) >> 4 : (n == 8 ? (s + 4) >> 3 : 128);
      }
    }
    for (int i = lane; i < 128; i += 32) {
      const int plane = i >> 6, y = (i >> 3) & 7, x = i & 7;
      const uint8_t* A = aboveC[plane] + 1;
      const uint8_t* L = leftC[plane];
      int v;
      switch (f.uv_mode) {
        case VP8GPU_DC_PRED: v = plane ? cdc[1] : cdc[0]; break;
        case VP8GPU_V_PRED: v = A[x]; break;
        case VP8GPU_H_PRED: v = L[y]; break;
        default: v = vp8m::clamp255(L[y] + A[x] - A[-1]);
      }
      pix[256 + i] = (uint8_t)v;
    }

    if (f.y_mode != VP8GPU_B_PRED) {
      // ---- luma 16x16 prediction (prediction.cc:451-467) ----
      const uint8_t* A = aboveY + 1;
      int dc = 128;
      if (f.y_mode == VP8GPU_DC_PRED) {
        int s = 0, n = 0;
        if (row > 0) { for (int k = 0; k < 16; k++) s += A[k]; n += 16; }
        if (col > 0) { for (int k = 0; k < 16; k++) s += leftY[k]; n += 16; }
        dc = n == 32 ? (s + 16) >> 5 : (n == 16 ? (s + 8) >> 4 : 128);
      }
      for (int i = lane; i < 256; i += 32) {
        const int y = i >> 4, x = i & 15;
        int v;
        switch (f.y_mode) {
          case VP8GPU_DC_PRED: v = dc; break;
          case VP8GPU_V_PRED: v = A[x]; break;
          case VP8GPU_H_PRED: v = leftY[y]; break;
          default: v = vp8m::clamp255(leftY[y] + A[x] - A[-1]);
        }
        pix[i] = (uint8_t)v;
      }
      __syncwarp();
      if (has_res) add_residuals(pix, coef, lane);
    } else {
      // ---- B_PRED: 16 sub-blocks in raster order, each predicted from reconstructed
      //      neighbours, residual added before the next one starts (macroblock.cc:540-545) ----
      __syncwarp();
      if (has_res) {  // chroma residual first (it does not interact with luma)
        for (int g4 = 64 + lane; g4 < 96; g4 += 32) {
          const int c = g4 - 64, plane = c >> 4, cc = c & 15;
          const int y = cc >> 1, x4 = (cc & 1) * 4;
          const int blk = 16 + plane * 4 + (y >> 2) * 2 + (x4 >> 2);
          const int16_t* r = coef + blk * CS + (y & 3) * 4;
          uint8_t* p = pix + 256 + plane * 64 + y * 8 + x4;
          for (int k = 0; k < 4; k++) p[k] = (uint8_t)vp8m::clamp255(p[k] + r[k]);
        }
      }
      const uint64_t modes = ((uint64_t)f.bm_hi << 32) | f.bm_lo;
      for (int b = 0; b < 16; b++) {
        const int bx = b & 3, by = b >> 2;
        const int mode = (int)((modes >> (4 * b)) & 15);
        // edge vector: s[0..3] = left[3..0], s[4] = above[-1], s[5..12] = above[0..7]
        if (lane < 13) {
          int v;
          if (lane < 4) {
            const int k = 3 - lane;  // left[k]
            v = bx ? pix[(4 * by + k) * 16 + 4 * bx - 1] : leftY[4 * by + k];
          } else {
            const int k = lane - 5;  // above[k], k = -1..7
            if (by == 0) v = aboveY[1 + 4 * bx + k];
            else if (k < 0) v = bx ? pix[(4 * by - 1) * 16 + 4 * bx - 1] : leftY[4 * by - 1];
            else if (k < 4 || bx < 3) v = pix[(4 * by - 1) * 16 + 4 * bx + k];
            else v = aboveY[17 + (k - 4)];  // right column, rows 1-3: the row above the macroblock
          }
          edge[lane] = (uint8_t)v;
        }
        __syncwarp();
        if (lane < 16) {
          const int x = lane & 3, y = lane >> 2;
          int v;
          if (mode == VP8GPU_B_DC_PRED) {
            int s = 4;
            for (int k = 0; k < 4; k++) s += edge[k] + edge[5 + k];
            v = s >> 3;
          } else if (mode == VP8GPU_B_TM_PRED) {
            v = vp8m::clamp255(edge[3 - y] + edge[5 + x] - edge[4]);
          } else {
            v = vp8m::bpred_eval(k_bpred_lut[(mode - 2) * 16 + lane], edge);
          }
          if (has_res) v = vp8m::clamp255(v + coef[b * CS + lane]);
          pix[(4 * by + y) * 16 + 4 * bx + x] = (uint8_t)v;
        }
        __syncwarp();
      }
    }
    __syncwarp();
    PROF(4);
    store_mb(pix, J.out, g, col, row, lane);
    const int next = next_marked(my_word, col + 1, nwords);
    PROF(5);
    publish_row(progress, next < 0 ? cols : next, lane);
    PROF(6);
    PROF_COUNT();
    col = next;
  }
  PROF_FLUSH(0);
}

// ================================================================================================
// k_loopfilter
// ================================================================================================
// One line of pixels across edges (a row for vertical edges, a column for horizontal ones), held in
// registers: px[0..3] = the 4 pixels before the macroblock, px[4..] = the macroblock's own.
// Edge order along the line = the reference's order for this direction (loopfilter.cc:133-154):
// macroblock edge (position 4) first, then the sub-block edges at 8, 12, 16 (chroma: 8 only).
__device__ __forceinline__ void filter_edge_at(int* px, int q, const vp8m::LfParams& lp, bool mb_edge) {
  const int mask = vp8m::lf_mask(lp.interior, mb_edge ? lp.mb_edge : lp.sub_edge, px[q - 4], px[q - 3], px[q - 2],
                                 px[q - 1], px[q], px[q + 1], px[q + 2], px[q + 3]);
  if (!mask) return;  // both filters are the identity when the mask is 0
  const int hev = vp8m::lf_hev(lp.hev, px[q - 2], px[q - 1], px[q], px[q + 1]);
  if (mb_edge) vp8m::lf_mbedge(mask, hev, px[q - 3], px[q - 2], px[q - 1], px[q], px[q + 1], px[q + 2]);
  else vp8m::lf_inner(mask, hev, px[q - 2], px[q - 1], px[q], px[q + 1]);
}
__device__ __forceinline__ void filter_line(int* px, bool luma, bool do_mb_edge, bool do_inner,
                                            const vp8m::LfParams& lp) {
  if (do_mb_edge) filter_edge_at(px, 4, lp, true);
  if (do_inner) {
    filter_edge_at(px, 8, lp, false);
    if (luma) {
      filter_edge_at(px, 12, lp, false);
      filter_edge_at(px, 16, lp, false);
    }
  }
}

__global__ void __launch_bounds__(32) k_loopfilter(const DevJob* __restrict__ jobs, int njobs, Geom g, int* ticket) {
  // region = the macroblock plus 4 pixels above and to the left: luma 20x20, chroma 12x12
  constexpr int YS = 20, CSZ = 12;
  __shared__ __align__(16) uint8_t ry[20 * YS];
  __shared__ __align__(16) uint8_t rc[2][12 * CSZ];
  const int lane = threadIdx.x;
  int t = 0;
  if (lane == 0) t = atomicAdd(ticket, 1);
  t = __shfl_sync(0xffffffffu, t, 0);
  const int row = t / njobs, job = t - row * njobs;
  if (row >= g.mb_rows) return;
  const DevJob& J = jobs[job];
  if (!J.lf_enabled) return;
  const int cols = g.mb_cols;
  const vp8gpu_mb* row_mbs = J.mbs + (size_t)row * cols;

  const int nwords = (cols + 31) >> 5;
  uint32_t my_word = 0;
  for (int w = 0; w < nwords; w++) {
    const int c = w * 32 + lane;
    const bool filtered = c < cols && ((__ldg(reinterpret_cast<const uint32_t*>(row_mbs + c) + 2) >> 16) & 0xFF) != 0;
    const uint32_t bits = __ballot_sync(0xffffffffu, filtered);
    if (lane == w) my_word = bits;
  }
  int col = next_marked(my_word, 0, nwords);
  int* progress = J.lf_progress + row;
  publish_row(progress, col < 0 ? cols : col, lane);

  uint8_t* const Y = J.out;
  uint8_t* const U = J.out + g.u_off;
  uint8_t* const V = J.out + g.v_off;
  const int y_lo = row > 0 ? 0 : 4;  // first region row that exists in the frame

  // per-lane addressing of the three transfer patterns (all in 32-bit words)
  //  own block (3 words per lane): words 0-63 luma 16 rows x 4, 64-95 chroma 2 planes x 8 rows x 2
  //  top rows  (1 word per lane):  lanes 0-15 luma 4 rows x 4, lanes 16-31 chroma 2 x 4 rows x 2
  auto own_ptr = [&](int k, int c, const uint8_t*& gp, uint8_t*& sp) {
    const int w = lane + 32 * k;
    if (w < 64) {
      const int r = w >> 2, wx = w & 3;
      gp = Y + (size_t)(16 * row + r) * g.y_pitch + 16 * c + 4 * wx;
      sp = ry + (4 + r) * YS + 4 + 4 * wx;
    } else {
      const int cw = w - 64, plane = cw >> 4, k2 = cw & 15, r = k2 >> 1, wx = k2 & 1;
      gp = (plane ? V : U) + (size_t)(8 * row + r) * g.c_pitch + 8 * c + 4 * wx;
      sp = rc[plane] + (4 + r) * CSZ + 4 + 4 * wx;
    }
  };
  uint32_t own[3];
  auto prefetch_own = [&](int c) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const uint8_t* gp;
      uint8_t* sp;
      own_ptr(k, c, gp, sp);
      own[k] = __ldcg(reinterpret_cast<const uint32_t*>(gp));
    }
  };
  if (col >= 0) prefetch_own(col);
  int prev = -2;  // last column this warp filtered (its right 4 columns are still in shared memory)

  PROF_DECL;
  while (col >= 0) {
    PROF(7);
    const MbFields f = load_mb(row_mbs + col);
    const bool have_left = prev == col - 1;
    PROF(0);
    if (row > 0) wait_row(progress - 1, min(col + 2, cols), lane);
    PROF(1);

    // ---- top 4 rows (final output of the row above), through L2: one word per lane ----
    uint32_t top = 0;
    if (row > 0) {
      if (lane < 16) {
        const int r = lane >> 2, wx = lane & 3;
        top = __ldcg(reinterpret_cast<const uint32_t*>(Y + (size_t)(16 * row - 4 + r) * g.y_pitch + 16 * col + 4 * wx));
      } else {
        const int cw = lane - 16, plane = cw >> 3, k2 = cw & 7, r = k2 >> 1, wx = k2 & 1;
        top = __ldcg(reinterpret_cast<const uint32_t*>((plane ? V : U) + (size_t)(8 * row - 4 + r) * g.c_pitch + 8 * col + 4 * wx));
      }
    }
    // ---- left 4 columns: slide them over from the previous macroblock, or fetch them ----
    uint32_t left0 = 0, left1 = 0;
    if (col > 0) {
      if (have_left) {
        if (lane < 20) left0 = *reinterpret_cast<const uint32_t*>(ry + lane * YS + 16);
        if (lane < 24) left1 = *reinterpret_cast<const uint32_t*>(rc[lane / 12] + (lane % 12) * CSZ + 8);
      } else {
        if (lane < 20 && lane >= y_lo)
          left0 = __ldcg(reinterpret_cast<const uint32_t*>(Y + (size_t)(16 * row - 4 + lane) * g.y_pitch + 16 * col - 4));
        if (lane < 24 && (lane % 12) >= y_lo)
          left1 = __ldcg(reinterpret_cast<const uint32_t*>((lane / 12 ? V : U) + (size_t)(8 * row - 4 + lane % 12) * g.c_pitch + 8 * col - 4));
      }
    }
    __syncwarp();  // everybody has read the old region before it is overwritten
    if (col > 0) {
      if (lane < 20) *reinterpret_cast<uint32_t*>(ry + lane * YS) = left0;
      if (lane < 24) *reinterpret_cast<uint32_t*>(rc[lane / 12] + (lane % 12) * CSZ) = left1;
    }
    if (row > 0) {
      if (lane < 16) *reinterpret_cast<uint32_t*>(ry + (lane >> 2) * YS + 4 + 4 * (lane & 3)) = top;
      else {
        const int cw = lane - 16, plane = cw >> 3, k2 = cw & 7;
        *reinterpret_cast<uint32_t*>(rc[plane] + (k2 >> 1) * CSZ + 4 + 4 * (k2 & 1)) = top;
      }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const uint8_t* gp;
      uint8_t* sp;
      own_ptr(k, col, gp, sp);
      *reinterpret_cast<uint32_t*>(sp) = own[k];
    }
    __syncwarp();
    PROF(2);
    const int next = next_marked(my_word, col + 1, nwords);
    if (next >= 0) prefetch_own(next);  // in flight while this macroblock is filtered

    const vp8m::LfParams lp = vp8m::lf_params(f.lf_level, J.sharpness, J.key_frame);
    const bool do_inner = !((f.flags & VP8GPU_MB_HAS_Y2) && f.tok_cnt == 0);  // macroblock.cc:608
    // lane roles on an edge: 0-15 luma positions, 16-23 U, 24-31 V
    const bool luma = lane < 16;
    uint8_t* const plane_base = luma ? ry : rc[(lane - 16) >> 3];
    const int stride = luma ? YS : CSZ, idx = luma ? lane : (lane & 7), len = luma ? 20 : 12;
    int px[20];

    // ---- vertical edges: one region row (4 + idx) per lane, in registers ----
    {
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(plane_base + (4 + idx) * stride);
#pragma unroll
      for (int k = 0; k < 5; k++) {
        const uint32_t v = (k < 3 || luma) ? rw[k] : 0u;
        px[4 * k] = v & 0xFF, px[4 * k + 1] = (v >> 8) & 0xFF, px[4 * k + 2] = (v >> 16) & 0xFF, px[4 * k + 3] = v >> 24;
      }
      filter_line(px, luma, col > 0, do_inner, lp);
      uint32_t* ww = reinterpret_cast<uint32_t*>(plane_base + (4 + idx) * stride);
#pragma unroll
      for (int k = 0; k < 5; k++)
        if (k < 3 || luma) ww[k] = (uint32_t)px[4 * k] | ((uint32_t)px[4 * k + 1] << 8) | ((uint32_t)px[4 * k + 2] << 16) | ((uint32_t)px[4 * k + 3] << 24);
    }
    __syncwarp();
    // ---- horizontal edges: one region column (4 + idx) per lane ----
    {
      uint8_t* cp = plane_base + 4 + idx;
#pragma unroll
      for (int k = 0; k < 20; k++) px[k] = k < len ? cp[k * stride] : 0;
      filter_line(px, luma, row > 0, do_inner, lp);
#pragma unroll
      for (int k = 1; k < 19; k++)
        if (k < len - 1) cp[k * stride] = (uint8_t)px[k];
    }
    __syncwarp();

    PROF(3);
    // ---- write back: region columns 0..15 (x -4..11); the last 4 columns travel with the next
    //      macroblock unless this warp will not filter it ----
    const int x_lo = col > 0 ? 0 : 1;
    const bool flush_right = next != col + 1;
    const int lw = flush_right ? 5 : 4, cw_n = flush_right ? 3 : 2;
    for (int i = lane; i < 20 * lw; i += 32) {
      const int r = i / lw, wx = i - r * lw;
      if (r >= y_lo && wx >= x_lo)
        *reinterpret_cast<uint32_t*>(Y + (size_t)(16 * row - 4 + r) * g.y_pitch + 16 * col - 4 + 4 * wx) =
            *reinterpret_cast<const uint32_t*>(ry + r * YS + 4 * wx);
    }
    for (int i = lane; i < 24 * cw_n; i += 32) {
      const int plane = i / (12 * cw_n), k = i - plane * 12 * cw_n, r = k / cw_n, wx = k - r * cw_n;
      if (r >= y_lo && wx >= x_lo)
        *reinterpret_cast<uint32_t*>((plane ? V : U) + (size_t)(8 * row - 4 + r) * g.c_pitch + 8 * col - 4 + 4 * wx) =
            *reinterpret_cast<const uint32_t*>(rc[plane] + r * CSZ + 4 * wx);
    }
    PROF(4);
    publish_row(progress, next < 0 ? cols : next, lane);
    PROF(5);
    PROF_COUNT();
    prev = col;
    col = next;
  }
  PROF_FLUSH(16);
}

// ================================================================================================
// k_compare: References::operator== (decoder.cc:249-254) on the device; flag != 0 when any visible
// pixel of the MB-aligned planes differs (pitch padding is ignored).
// ================================================================================================
__global__ void k_compare(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, Geom g, int* flag) {
  const int words_y = g.W / 4, words_c = g.W / 8;
  const int rows_total = g.H + g.H;  // H luma rows + H/2 U rows + H/2 V rows
  int diff = 0;
  for (int r = blockIdx.x; r < rows_total; r += gridDim.x) {
    size_t off;
    int nw;
    if (r < g.H) off = (size_t)r * g.y_pitch, nw = words_y;
    else if (r < g.H + g.H / 2) off = g.u_off + (size_t)(r - g.H) * g.c_pitch, nw = words_c;
    else off = g.v_off + (size_t)(r - g.H - g.H / 2) * g.c_pitch, nw = words_c;
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(a + off);
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(b + off);
    for (int i = threadIdx.x; i < nw; i += blockDim.x) diff |= (pa[i] != pb[i]);
  }
  if (diff) atomicOr(flag, 1);
}

}  // namespace

// ================================================================================================
// launchers
// ================================================================================================
int launch_inter(const DevJob* jobs, int njobs, const Geom& g, void* stream) {
  const int n_mbs = g.mb_cols * g.mb_rows;
  dim3 grid((n_mbs + INTER_WARPS - 1) / INTER_WARPS, njobs);
  k_inter<<<grid, INTER_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(jobs, g);
  return (int)cudaGetLastError();
}
int launch_intra(const DevJob* jobs, int njobs, const Geom& g, int* ticket, void* stream) {
  k_intra<<<g.mb_rows * njobs, 32, 0, static_cast<cudaStream_t>(stream)>>>(jobs, njobs, g, ticket);
  return (int)cudaGetLastError();
}
int launch_loopfilter(const DevJob* jobs, int njobs, const Geom& g, int* ticket, void* stream) {
  k_loopfilter<<<g.mb_rows * njobs, 32, 0, static_cast<cudaStream_t>(stream)>>>(jobs, njobs, g, ticket);
  return (int)cudaGetLastError();
}

#ifdef VP8_PROFILE
extern "C" void vp8gpu_debug_profile(unsigned long long out[32], int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_prof, sizeof(unsigned long long) * 32);
  if (reset) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(g_prof, z, sizeof(z));
  }
}
#endif

int launch_compare(const uint8_t* a, const uint8_t* b, const Geom& g, int* d_flag, void* stream) {
  k_compare<<<296, 128, 0, static_cast<cudaStream_t>(stream)>>>(a, b, g, d_flag);
  return (int)cudaGetLastError();
}

}  // namespace vp8
