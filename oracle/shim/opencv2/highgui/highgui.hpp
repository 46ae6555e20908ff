// TEST INFRASTRUCTURE: see opencv2/imgproc/imgproc.hpp in this directory.  cv::imread for the one
// uncompressed format OpenCV reads that needs no codec: binary PGM ("P5"), 8 or 16 bit (big endian),
// so that the reference app's own main() can load its frames (FrameAlignment.cpp:74-81).
#ifndef PHOVO_SHIM_OPENCV_HIGHGUI_HPP_
#define PHOVO_SHIM_OPENCV_HIGHGUI_HPP_
#include <cstdio>
#include <string>
#include <vector>
#include "../imgproc/imgproc.hpp"
namespace cv
{
// what imread returns: converts to whichever Mat_<T> it is assigned to (like cv::Mat -> cv::Mat_<T>)
struct Mat
{
  Mat() : rows( 0 ), cols( 0 ) {}
  int rows, cols;
  std::vector< unsigned short > values;
  template< class T > operator Mat_< T >() const
  {
    Mat_< T > m( rows, cols );
    for( int r = 0; r < rows; r++ ) for( int c = 0; c < cols; c++ ) m( r, c ) = T( values[ size_t( r ) * size_t( cols ) + size_t( c ) ] );
    return m;
  }
};
inline Mat imread( const std::string & path, int /*flags*/ )
{
  Mat m;
  FILE * f = std::fopen( path.c_str(), "rb" );
  if( !f ) return m;
  int w = 0, h = 0, maxval = 0;
  if( std::fscanf( f, "P5 %d %d %d", &w, &h, &maxval ) == 3 && std::fgetc( f ) != EOF )
  {
    m.rows = h; m.cols = w; m.values.resize( size_t( w ) * size_t( h ) );
    for( size_t k = 0; k < m.values.size(); k++ )
    {
      int v = std::fgetc( f );
      if( maxval > 255 ) v = ( v << 8 ) | std::fgetc( f );
      m.values[k] = static_cast< unsigned short >( v );
    }
  }
  std::fclose( f );
  return m;
}
} // namespace cv
#endif
