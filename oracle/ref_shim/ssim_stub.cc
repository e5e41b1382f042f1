/* Link-time stand-in for util/ssim.cc (binds libx264 internals that are not in
 * this image). The decode path never calls it. Test scaffolding only. */
#include "2d.hh"
#include "ssim.hh"
double ssim(const TwoD<uint8_t>&, const TwoD<uint8_t>&) { return 0.0; }
