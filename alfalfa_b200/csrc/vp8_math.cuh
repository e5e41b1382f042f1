// vp8_math.cuh -- the integer arithmetic of the VP8 pixel pipeline, written once and usable
// from both the sm_100a kernels and (for CPU unit tests of the arithmetic, tests/test_math_host.py)
// a plain g++ build.  Everything here is per-value math; the warp-level data movement lives in
// kernels.cu.  Each function names the reference code whose results it must reproduce
// (paths relative to /root/reference/src).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define VP8_HD __host__ __device__ __forceinline__
#else
#define VP8_HD inline
#endif

namespace vp8m {

VP8_HD int clamp255(int x) { return x < 0 ? 0 : (x > 255 ? 255 : x); }
VP8_HD int sclamp(int t) { return t < -128 ? -128 : (t > 127 ? 127 : t); }
VP8_HD int iabs(int a) { return a < 0 ? -a : a; }
// value of (int16_t)x computed in int registers
VP8_HD int wrap16(int x) { return (int)(int16_t)x; }

// ---- inverse Walsh-Hadamard: decoder/transform.cc:47-88 -------------------------------------
// in: 16 dequantised Y2 coefficients; out[k] = DC of luma sub-block k (raster order).
VP8_HD void iwht16(const int16_t* in, int16_t* out) {
  int m[16];
  for (int i = 0; i < 4; i++) {
    const int a1 = in[i] + in[i + 12], b1 = in[i + 4] + in[i + 8];
    const int c1 = in[i + 4] - in[i + 8], d1 = in[i] - in[i + 12];
    m[i] = wrap16(a1 + b1);
    m[i + 4] = wrap16(c1 + d1);
    m[i + 8] = wrap16(a1 - b1);
    m[i + 12] = wrap16(d1 - c1);
  }
  for (int i = 0; i < 4; i++) {
    const int o = 4 * i;
    const int a1 = m[o] + m[o + 3], b1 = m[o + 1] + m[o + 2];
    const int c1 = m[o + 1] - m[o + 2], d1 = m[o] - m[o + 3];
    out[o + 0] = (int16_t)((a1 + b1 + 3) >> 3);
    out[o + 1] = (int16_t)((c1 + d1 + 3) >> 3);
    out[o + 2] = (int16_t)((a1 - b1 + 3) >> 3);
    out[o + 3] = (int16_t)((d1 - c1 + 3) >> 3);
  }
}

// ---- inverse DCT: decoder/transform.cc:100-137 ------------------------------------------------
// Produces the residual r[y*4+x] that idct_add adds to the prediction before clamping:
// pixel = clamp255(pred + r).  (|r| <= ~16k, so it fits an int16.)
VP8_HD int mul_20091(int a) { return ((a * 20091) >> 16) + a; }
VP8_HD int mul_35468(int a) { return (a * 35468) >> 16; }
VP8_HD void idct16(const int16_t* c, int16_t* r) {
  int m[16];
  for (int i = 0; i < 4; i++) {
    const int t0 = c[i] + c[i + 8], t1 = c[i] - c[i + 8];
    const int t2 = mul_35468(c[i + 4]) - mul_20091(c[i + 12]);
    const int t3 = mul_20091(c[i + 4]) + mul_35468(c[i + 12]);
    m[i * 4 + 0] = wrap16(t0 + t3);
    m[i * 4 + 1] = wrap16(t1 + t2);
    m[i * 4 + 2] = wrap16(t1 - t2);
    m[i * 4 + 3] = wrap16(t0 - t3);
  }
  for (int i = 0; i < 4; i++) {
    const int t0 = m[i] + m[i + 8], t1 = m[i] - m[i + 8];
    const int t2 = mul_35468(m[i + 4]) - mul_20091(m[i + 12]);
    const int t3 = mul_20091(m[i + 4]) + mul_35468(m[i + 12]);
    r[i * 4 + 0] = (int16_t)((t0 + t3 + 4) >> 3);
    r[i * 4 + 1] = (int16_t)((t1 + t2 + 4) >> 3);
    r[i * 4 + 2] = (int16_t)((t1 - t2 + 4) >> 3);
    r[i * 4 + 3] = (int16_t)((t0 - t3 + 4) >> 3);
  }
}

// ---- forward DCT / WHT and quantiser (encoder): decoder/dct.cc:45-164, quantization.cc:148-178 ----
// d = residual (source - prediction), raster order; out[4 * vertical_freq + horizontal_freq]
VP8_HD void fdct16(const int16_t* d, int16_t* out) {
  int t[16];
  for (int i = 0; i < 4; i++) {
    const int a1 = (d[4 * i + 0] + d[4 * i + 3]) * 8, b1 = (d[4 * i + 1] + d[4 * i + 2]) * 8;
    const int c1 = (d[4 * i + 1] - d[4 * i + 2]) * 8, d1 = (d[4 * i + 0] - d[4 * i + 3]) * 8;
    t[4 * i + 0] = wrap16(a1 + b1);
    t[4 * i + 2] = wrap16(a1 - b1);
    t[4 * i + 1] = wrap16((c1 * 2217 + d1 * 5352 + 14500) >> 12);
    t[4 * i + 3] = wrap16((d1 * 2217 - c1 * 5352 + 7500) >> 12);
  }
  for (int i = 0; i < 4; i++) {
    const int a1 = t[i] + t[i + 12], b1 = t[i + 4] + t[i + 8];
    const int c1 = t[i + 4] - t[i + 8], d1 = t[i] - t[i + 12];
    out[i] = (int16_t)((a1 + b1 + 7) >> 4);
    out[i + 8] = (int16_t)((a1 - b1 + 7) >> 4);
    out[i + 4] = (int16_t)(((c1 * 2217 + d1 * 5352 + 12000) >> 16) + (d1 != 0));
    out[i + 12] = (int16_t)((d1 * 2217 - c1 * 5352 + 51000) >> 16);
  }
}
// in = the 16 luma DC coefficients (raster order of sub-blocks); out = Y2 coefficients
VP8_HD void fwht16(const int16_t* in, int16_t* out) {
  int t[16];
  for (int i = 0; i < 4; i++) {
    const int a1 = (in[4 * i + 0] + in[4 * i + 2]) * 4, d1 = (in[4 * i + 1] + in[4 * i + 3]) * 4;
    const int c1 = (in[4 * i + 1] - in[4 * i + 3]) * 4, b1 = (in[4 * i + 0] - in[4 * i + 2]) * 4;
    t[4 * i + 0] = wrap16(a1 + d1 + (a1 != 0));
    t[4 * i + 1] = wrap16(b1 + c1);
    t[4 * i + 2] = wrap16(b1 - c1);
    t[4 * i + 3] = wrap16(a1 - d1);
  }
  for (int i = 0; i < 4; i++) {
    const int a1 = t[i] + t[i + 8], d1 = t[i + 4] + t[i + 12];
    const int c1 = t[i + 4] - t[i + 12], b1 = t[i] - t[i + 8];
    int a2 = a1 + d1, b2 = b1 + c1, c2 = b1 - c1, d2 = a1 - d1;
    a2 += a2 < 0;
    b2 += b2 < 0;
    c2 += c2 < 0;
    d2 += d2 < 0;
    out[i] = (int16_t)((a2 + 3) >> 3);
    out[i + 4] = (int16_t)((b2 + 3) >> 3);
    out[i + 8] = (int16_t)((c2 + 3) >> 3);
    out[i + 12] = (int16_t)((d2 + 3) >> 3);
  }
}
// DCTCoefficients::quantize: C++ integer division, i.e. truncation toward zero
VP8_HD int quantize_trunc(int coef, int factor) { return coef / factor; }

// ---- six-tap sub-pixel filter: decoder/prediction.cc:645-653, 919-971 ------------------------
VP8_HD int sixtap(int p0, int p1, int p2, int p3, int p4, int p5, const int16_t* t) {
  return clamp255((p0 * t[0] + p1 * t[1] + p2 * t[2] + p3 * t[3] + p4 * t[4] + p5 * t[5] + 64) >> 7);
}

// ---- the same filter on packed pixels (what k_inter executes) ---------------------------------
// Every non-identity tap set fits a signed byte (|tap| <= 123), so a row of six taps is two 4-byte
// dot products: t03 = taps 0..3, t45 = taps 4..5 (upper two bytes zero).
VP8_HD uint32_t pack_taps03(const int16_t* t) {
  return (uint32_t)(uint8_t)t[0] | ((uint32_t)(uint8_t)t[1] << 8) | ((uint32_t)(uint8_t)t[2] << 16) | ((uint32_t)(uint8_t)t[3] << 24);
}
VP8_HD uint32_t pack_taps45(const int16_t* t) { return (uint32_t)(uint8_t)t[4] | ((uint32_t)(uint8_t)t[5] << 8); }
// c + sum over the four bytes of (unsigned byte of a) * (signed byte of b)
VP8_HD int dot4_us(uint32_t a, uint32_t b, int c) {
#ifdef __CUDA_ARCH__
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  for (int k = 0; k < 4; k++) c += (int)((a >> (8 * k)) & 0xFF) * (int)(int8_t)((b >> (8 * k)) & 0xFF);
  return c;
#endif
}
// bytes [s/8 .. s/8+3] of the 8-byte sequence lo | hi << 32 (s = 0, 8, 16, 24)
VP8_HD uint32_t bytes_at(uint32_t lo, uint32_t hi, int s) {
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, s);
#else
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}
// sat_u8(v3) << 24 | sat_u8(v2) << 16 | sat_u8(v1) << 8 | sat_u8(v0)
VP8_HD uint32_t pack_sat4(int v0, int v1, int v2, int v3) {
#ifdef __CUDA_ARCH__
  uint32_t t, d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(v3), "r"(v2), "r"(0));
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v1), "r"(v0), "r"(t));
  return d;
#else
  return (uint32_t)clamp255(v0) | ((uint32_t)clamp255(v1) << 8) | ((uint32_t)clamp255(v2) << 16) | ((uint32_t)clamp255(v3) << 24);
#endif
}
VP8_HD uint32_t pack_sat2(int v0, int v1) {
#ifdef __CUDA_ARCH__
  uint32_t d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v1), "r"(v0), "r"(0));
  return d;
#else
  return (uint32_t)clamp255(v0) | ((uint32_t)clamp255(v1) << 8);
#endif
}
// horizontal pass: four outputs from twelve consecutive pixels w0 | w1 | w2 (output j uses pixels j .. j+5)
VP8_HD uint32_t sixtap_h4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t t03, uint32_t t45) {
  int v[4];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int j = 0; j < 4; j++)
    v[j] = dot4_us(bytes_at(w1, w2, 8 * j), t45, dot4_us(bytes_at(w0, w1, 8 * j), t03, 64)) >> 7;
  return pack_sat4(v[0], v[1], v[2], v[3]);
}
// vertical pass on pixel PAIRS: each p[k] holds two pixels of row k in its 16-bit halves
// (pixel | pixel' << 16).  With a bias of 8192 + 64 per half neither half can borrow from or carry into
// the other (a filtered sum lies in [-8160, 40800]), so six 32-bit multiply-adds filter two columns.
// Returns sat_u8(out) | sat_u8(out') << 8.
VP8_HD uint32_t pair_of(uint32_t two_bytes) { return (two_bytes & 0xFF) | ((two_bytes & 0xFF00) << 8); }
VP8_HD uint32_t sixtap_v2(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t p4, uint32_t p5, const int16_t* t) {
  const uint32_t acc = 0x20402040u + p0 * (uint32_t)(int)t[0] + p1 * (uint32_t)(int)t[1] + p2 * (uint32_t)(int)t[2] +
                       p3 * (uint32_t)(int)t[3] + p4 * (uint32_t)(int)t[4] + p5 * (uint32_t)(int)t[5];
  return pack_sat2((int)((acc >> 7) & 0x1FF) - 64, (int)(acc >> 23) - 64);
}
// pixel + residual with saturation on four pixels: r01 / r23 hold the int16 residuals of pixels 0,1 / 2,3
VP8_HD uint32_t add_residual4(uint32_t pix, uint32_t r01, uint32_t r23) {
  return pack_sat4((int)(pix & 0xFF) + (int)(int16_t)(r01 & 0xFFFF), (int)((pix >> 8) & 0xFF) + ((int)r01 >> 16),
                   (int)((pix >> 16) & 0xFF) + (int)(int16_t)(r23 & 0xFFFF), (int)(pix >> 24) + ((int)r23 >> 16));
}

// ---- 4x4 directional intra prediction through the generated table (tools/gen_bpred_lut.py) ----
// s = 13-entry edge vector: s[0..3] = left[3..0], s[4] = above[-1], s[5..12] = above[0..7]
VP8_HD int bpred_eval(unsigned entry, const uint8_t* s) {
  const int a = s[entry & 15], b = s[(entry >> 4) & 15], c = s[(entry >> 8) & 15];
  return (entry & 0x1000) ? ((a + 2 * b + c + 2) >> 2) : ((a + b + 1) >> 1);
}

// ---- normal loop filter: decoder/loopfilter_filters.hh:50-183, loopfilter.cc:81-125 -----------
struct LfParams {
  int interior, mb_edge, sub_edge, hev;
};
VP8_HD LfParams lf_params(int level /* 1..63 */, int sharpness, int key_frame) {
  LfParams p;
  int interior = level;
  if (sharpness) {
    interior >>= sharpness > 4 ? 2 : 1;
    if (interior > 9 - sharpness) interior = 9 - sharpness;
  }
  if (interior < 1) interior = 1;
  p.interior = interior;
  p.mb_edge = ((level + 2) * 2) + interior;
  p.sub_edge = (level * 2) + interior;
  p.hev = (level >= 15) + (level >= 40) + ((level >= 20) && !key_frame);
  return p;
}
// returns -1 (filter) or 0 (leave alone)
VP8_HD int lf_mask(int limit, int blimit, int p3, int p2, int p1, int p0, int q0, int q1, int q2, int q3) {
  int m = 0;
  m |= (iabs(p3 - p2) > limit);
  m |= (iabs(p2 - p1) > limit);
  m |= (iabs(p1 - p0) > limit);
  m |= (iabs(q1 - q0) > limit);
  m |= (iabs(q2 - q1) > limit);
  m |= (iabs(q3 - q2) > limit);
  m |= (iabs(p0 - q0) * 2 + iabs(p1 - q1) / 2 > blimit);
  return m - 1;
}
VP8_HD int lf_hev(int thresh, int p1, int p0, int q0, int q1) {
  return ((iabs(p1 - p0) > thresh) || (iabs(q1 - q0) > thresh)) ? -1 : 0;
}
// vp8_filter (sub-block edges): modifies p1 p0 q0 q1 (pixel values 0..255)
VP8_HD void lf_inner(int mask, int hev, int& p1, int& p0, int& q0, int& q1) {
  const int ps1 = p1 - 128, ps0 = p0 - 128, qs0 = q0 - 128, qs1 = q1 - 128;
  int f = sclamp(ps1 - qs1) & hev;
  f = sclamp(f + 3 * (qs0 - ps0)) & mask;
  const int f1 = sclamp(f + 4) >> 3, f2 = sclamp(f + 3) >> 3;
  q0 = sclamp(qs0 - f1) + 128;
  p0 = sclamp(ps0 + f2) + 128;
  f = ((f1 + 1) >> 1) & ~hev;
  q1 = sclamp(qs1 - f) + 128;
  p1 = sclamp(ps1 + f) + 128;
}
// vp8_mbfilter (macroblock edges): modifies p2 p1 p0 q0 q1 q2
VP8_HD void lf_mbedge(int mask, int hev, int& p2, int& p1, int& p0, int& q0, int& q1, int& q2) {
  const int ps2 = p2 - 128, ps1 = p1 - 128, qs1 = q1 - 128, qs2 = q2 - 128;
  int ps0 = p0 - 128, qs0 = q0 - 128;
  int f = sclamp(ps1 - qs1);
  f = sclamp(f + 3 * (qs0 - ps0)) & mask;
  int f2 = f & hev;
  const int f1 = sclamp(f2 + 4) >> 3;
  f2 = sclamp(f2 + 3) >> 3;
  qs0 = sclamp(qs0 - f1);
  ps0 = sclamp(ps0 + f2);
  f &= ~hev;
  int u = sclamp((63 + f * 27) >> 7);
  q0 = sclamp(qs0 - u) + 128;
  p0 = sclamp(ps0 + u) + 128;
  u = sclamp((63 + f * 18) >> 7);
  q1 = sclamp(qs1 - u) + 128;
  p1 = sclamp(ps1 + u) + 128;
  u = sclamp((63 + f * 9) >> 7);
  q2 = sclamp(qs2 - u) + 128;
  p2 = sclamp(ps2 + u) + 128;
}

}  // namespace vp8m
