// CPU build of the kernels' arithmetic header (alfalfa_b200/csrc/vp8_math.cuh + the generated
// B_PRED table) for tests/test_math_host.py.  Test scaffolding; not part of the product library.
#include <stdint.h>
#include <string.h>

#include "../alfalfa_b200/csrc/vp8_math.cuh"
#define VP8_LUT_QUALIFIER static const
#include "../alfalfa_b200/csrc/bpred_lut.inc"

static const int16_t kSixtap[8][6] = {{0, 0, 128, 0, 0, 0},     {0, -6, 123, 12, -1, 0}, {2, -11, 108, 36, -8, 1},
                                      {0, -9, 93, 50, -6, 0},   {3, -16, 77, 77, -16, 3}, {0, -6, 50, 93, -9, 0},
                                      {1, -8, 36, 108, -11, 2}, {0, -1, 12, 123, -6, 0}};
extern "C" {
void m_idct_add(const int16_t* c, uint8_t* px) {
  int16_t r[16];
  vp8m::idct16(c, r);
  for (int i = 0; i < 16; i++) px[i] = (uint8_t)vp8m::clamp255(px[i] + r[i]);
}
void m_fdct(const uint8_t* src, const uint8_t* pred, int16_t* out) {
  int16_t d[16];
  for (int i = 0; i < 16; i++) d[i] = (int16_t)(src[i] - pred[i]);
  vp8m::fdct16(d, out);
}
void m_fwht(const int16_t* in, int16_t* out) { vp8m::fwht16(in, out); }
void m_iwht(const int16_t* c, int16_t* dc) { vp8m::iwht16(c, dc); }
void m_lf_edge(uint8_t* px, int level, int sharpness, int key_frame, int mb_edge) {
  const vp8m::LfParams lp = vp8m::lf_params(level, sharpness, key_frame);
  int p[8];
  for (int i = 0; i < 8; i++) p[i] = px[i];
  const int mask = vp8m::lf_mask(lp.interior, mb_edge ? lp.mb_edge : lp.sub_edge, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]);
  const int hev = vp8m::lf_hev(lp.hev, p[2], p[3], p[4], p[5]);
  if (mb_edge) vp8m::lf_mbedge(mask, hev, p[1], p[2], p[3], p[4], p[5], p[6]);
  else vp8m::lf_inner(mask, hev, p[2], p[3], p[4], p[5]);
  for (int i = 0; i < 8; i++) px[i] = (uint8_t)p[i];
}
void m_bpred(int mode, const uint8_t* s, uint8_t* out) {
  for (int i = 0; i < 16; i++) {
    const int x = i & 3, y = i >> 2;
    int v;
    if (mode == 0) {
      int t = 4;
      for (int k = 0; k < 4; k++) t += s[k] + s[5 + k];
      v = t >> 3;
    } else if (mode == 1) {
      v = vp8m::clamp255(s[3 - y] + s[5 + x] - s[4]);
    } else {
      v = vp8m::bpred_eval(k_bpred_lut[(mode - 2) * 16 + i], s);
    }
    out[i] = (uint8_t)v;
  }
}
void m_sixtap(const uint8_t* window, int n, int mx, int my, uint8_t* out) {
  const int ws = n + 5;
  uint8_t mid[21 * 16];
  if (mx == 0 && my == 0) {
    for (int y = 0; y < n; y++) memcpy(out + y * n, window + (y + 2) * ws + 2, n);
    return;
  }
  for (int r = 0; r < n + 5; r++)
    for (int c = 0; c < n; c++) {
      const uint8_t* t = window + r * ws + c;
      mid[r * n + c] = mx ? (uint8_t)vp8m::sixtap(t[0], t[1], t[2], t[3], t[4], t[5], kSixtap[mx]) : t[2];
    }
  for (int r = 0; r < n; r++)
    for (int c = 0; c < n; c++) {
      const uint8_t* m = mid + r * n + c;
      out[r * n + c] = my ? (uint8_t)vp8m::sixtap(m[0], m[n], m[2 * n], m[3 * n], m[4 * n], m[5 * n], kSixtap[my]) : m[2 * n];
    }
}
// the packed filter arithmetic of k_inter (sixtap_h4 / sixtap_v2 / add_residual4) on the same window
void m_sixtap_packed(const uint8_t* window, int n, int mx, int my, uint8_t* out) {
  const int ws = n + 5;
  uint8_t mid[21 * 16];
  if (mx == 0 && my == 0) {
    for (int y = 0; y < n; y++) memcpy(out + y * n, window + (y + 2) * ws + 2, n);
    return;
  }
  for (int r = 0; r < n + 5; r++)
    for (int g = 0; g < n / 4; g++) {
      uint8_t b[12];
      for (int k = 0; k < 12; k++) b[k] = (4 * g + k < ws) ? window[r * ws + 4 * g + k] : 0xEE;  // bytes past the window never matter
      uint32_t w[3];
      memcpy(w, b, 12);
      uint32_t o;
      if (mx) o = vp8m::sixtap_h4(w[0], w[1], w[2], vp8m::pack_taps03(kSixtap[mx]), vp8m::pack_taps45(kSixtap[mx]));
      else memcpy(&o, b + 2, 4);
      memcpy(mid + r * n + 4 * g, &o, 4);
    }
  for (int r = 0; r < n; r++)
    for (int c = 0; c < n; c += 2) {
      if (!my) {
        out[r * n + c] = mid[(r + 2) * n + c];
        out[r * n + c + 1] = mid[(r + 2) * n + c + 1];
        continue;
      }
      uint32_t p[6];
      for (int k = 0; k < 6; k++) p[k] = vp8m::pair_of((uint32_t)mid[(r + k) * n + c] | ((uint32_t)mid[(r + k) * n + c + 1] << 8));
      const uint32_t o = vp8m::sixtap_v2(p[0], p[1], p[2], p[3], p[4], p[5], kSixtap[my]);
      out[r * n + c] = (uint8_t)(o & 0xFF);
      out[r * n + c + 1] = (uint8_t)(o >> 8);
    }
}
void m_add_residual4(const uint8_t* px, const int16_t* r, uint8_t* out) {
  uint32_t p, r01, r23;
  memcpy(&p, px, 4);
  memcpy(&r01, r, 4);
  memcpy(&r23, r + 2, 4);
  const uint32_t o = vp8m::add_residual4(p, r01, r23);
  memcpy(out, &o, 4);
}
}
