/* Link-time glue (ours) that lets the UNMODIFIED reference encoder build without an assembler:
 * two things in the reference exist only under HAVE_SSE2:
 *   - the explicit Encoder::sad/sse/variance specialisations (encoder/variance.cc:84-164); the
 *     scalar templates at :34-82 are fine but never instantiated -> instantiate them here;
 *   - VP8Raster::Block<16>::inter_predict(mv, const SafeRaster&, out) (decoder/prediction.cc:680-734)
 *     -> scalar two-pass six-tap over SafeRaster::at, same arithmetic as safe_inter_predict (:919-971).
 * Test scaffolding only (used to synthesise >=1080p VP8 streams for the benchmarks). */
#include "encoder.hh"
#include "variance.cc"

template uint32_t Encoder::sad<16>( const VP8Raster::Block<16> &, const TwoDSubRange<uint8_t, 16, 16> & );
template uint32_t Encoder::sse<4>( const VP8Raster::Block<4> &, const TwoDSubRange<uint8_t, 4, 4> & );
template uint32_t Encoder::sse<8>( const VP8Raster::Block<8> &, const TwoDSubRange<uint8_t, 8, 8> & );
template uint32_t Encoder::sse<16>( const VP8Raster::Block<16> &, const TwoDSubRange<uint8_t, 16, 16> & );
template uint32_t Encoder::variance<16>( const VP8Raster::Block<16> &, const TwoDSubRange<uint8_t, 16, 16> & );

static const int kTaps[8][6] = { { 0, 0, 128, 0, 0, 0 },     { 0, -6, 123, 12, -1, 0 }, { 2, -11, 108, 36, -8, 1 },
                                 { 0, -9, 93, 50, -6, 0 },   { 3, -16, 77, 77, -16, 3 }, { 0, -6, 50, 93, -9, 0 },
                                 { 1, -8, 36, 108, -11, 2 }, { 0, -1, 12, 123, -6, 0 } };
static inline uint8_t c255( int x ) { return x < 0 ? 0 : ( x > 255 ? 255 : x ); }

template <>
void VP8Raster::Block<16>::inter_predict( const MotionVector & mv, const SafeRaster & reference,
                                          TwoDSubRange<uint8_t, 16, 16> & output ) const
{
  const int sc = column_ * 16 + ( mv.x() >> 3 ), sr = row_ * 16 + ( mv.y() >> 3 );
  const int mx = mv.x() & 7, my = mv.y() & 7;
  if ( mx == 0 and my == 0 ) {
    for ( int r = 0; r < 16; r++ )
      for ( int c = 0; c < 16; c++ ) output.at( c, r ) = reference.at( sc + c, sr + r );
    return;
  }
  uint8_t mid[ 21 ][ 16 ];
  for ( int r = 0; r < 21; r++ )
    for ( int c = 0; c < 16; c++ ) {
      int s = 64;
      for ( int k = 0; k < 6; k++ ) s += reference.at( sc + c + k - 2, sr + r - 2 ) * kTaps[ mx ][ k ];
      mid[ r ][ c ] = c255( s >> 7 );
    }
  for ( int r = 0; r < 16; r++ )
    for ( int c = 0; c < 16; c++ ) {
      int s = 64;
      for ( int k = 0; k < 6; k++ ) s += mid[ r + k ][ c ] * kTaps[ my ][ k ];
      output.at( c, r ) = c255( s >> 7 );
    }
}
