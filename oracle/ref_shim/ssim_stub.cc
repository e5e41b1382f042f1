/* Link-time stand-in for util/ssim.cc, which binds libx264-internal symbols
 * (x264_8_pixel_ssim_wxh, util/ssim.cc:36-45) that are not in this image and not vendored by the
 * reference (configure.ac:99 only asks pkg-config for some x264; version unpinned).
 * The decode path never calls ssim().  The reference ENCODER uses it to pick the loop-filter
 * level (encoder.cc:489-508), so for synthesising benchmark streams we restate x264's published
 * algorithm (common/pixel.c: ssim_4x4x2_core / ssim_end1 / ssim_end4 / pixel_ssim_wxh: integer sums
 * over 4x4 windows stepped by 4, combined over overlapping 8x8 neighbourhoods, float ratio) from
 * memory.  PARITY UNPINNED: no x264 source or golden SSIM value exists here to check it against;
 * it only steers which (valid) bitstream the reference encoder emits.  Test scaffolding only. */
#include <vector>

#include "2d.hh"
#include "ssim.hh"

static void core_4x4x2(const uint8_t* p1, int s1, const uint8_t* p2, int s2, int sums[2][4]) {
  for (int z = 0; z < 2; z++) {
    uint32_t a1 = 0, a2 = 0, ss = 0, s12 = 0;
    for (int y = 0; y < 4; y++)
      for (int x = 0; x < 4; x++) {
        const int a = p1[x + y * s1], b = p2[x + y * s2];
        a1 += a; a2 += b; ss += a * a; ss += b * b; s12 += a * b;
      }
    sums[z][0] = a1; sums[z][1] = a2; sums[z][2] = ss; sums[z][3] = s12;
    p1 += 4; p2 += 4;
  }
}
static float end1(int s1, int s2, int ss, int s12) {
  static const int c1 = (int)(.01 * .01 * 255 * 255 * 64 + .5);
  static const int c2 = (int)(.03 * .03 * 255 * 255 * 64 * 63 + .5);
  const int vars = ss * 64 - s1 * s1 - s2 * s2, covar = s12 * 64 - s1 * s2;
  return (float)(2 * s1 * s2 + c1) * (float)(2 * covar + c2) / ((float)(s1 * s1 + s2 * s2 + c1) * (float)(vars + c2));
}

double ssim(const TwoD<uint8_t>& image, const TwoD<uint8_t>& other) {
  const uint8_t *pix1 = &image.at(0, 0), *pix2 = &other.at(0, 0);
  const int stride1 = image.width(), stride2 = other.width();
  int width = image.width() >> 2, height = image.height() >> 2;
  std::vector<int> buf(8 * (width + 3));
  int(*sum0)[4] = reinterpret_cast<int(*)[4]>(buf.data());
  int(*sum1)[4] = sum0 + width + 3;
  float total = 0.0f;
  int z = 0;
  for (int y = 1; y < height; y++) {
    for (; z <= y; z++) {
      int(*t)[4] = sum0; sum0 = sum1; sum1 = t;
      for (int x = 0; x < width; x += 2)
        core_4x4x2(&pix1[4 * (x + z * stride1)], stride1, &pix2[4 * (x + z * stride2)], stride2, &sum0[x]);
    }
    for (int x = 0; x < width - 1; x++)
      total += end1(sum0[x][0] + sum0[x + 1][0] + sum1[x][0] + sum1[x + 1][0], sum0[x][1] + sum0[x + 1][1] + sum1[x][1] + sum1[x + 1][1],
                    sum0[x][2] + sum0[x + 1][2] + sum1[x][2] + sum1[x + 1][2], sum0[x][3] + sum0[x + 1][3] + sum1[x][3] + sum1[x + 1][3]);
  }
  const int count = (height - 1) * (width - 1);
  return count > 0 ? total / count : 0.0;
}
