// TEST INFRASTRUCTURE.  Minimal stand-in for the parts of OpenCV that the reference's
// CPhotoconsistencyOdometryAnalytic.h / CPhotoconsistencyOdometry.h touch, so that those headers can
// be compiled UNMODIFIED from /root/reference into oracle/_ref (see oracle/Makefile).  Third-party
// arithmetic (resize, GaussianBlur, Scharr, convertTo) is delegated to the oracle's restatements in
// phovo_oracle.c, which tests/ pin against the real OpenCV (python cv2) -- only first-party
// reference code runs "for real" in oracle/_ref.
#ifndef PHOVO_SHIM_OPENCV_IMGPROC_HPP_
#define PHOVO_SHIM_OPENCV_IMGPROC_HPP_

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <vector>

extern "C" {
void pho_level_size(int rows, int cols, int level, int* out_rows, int* out_cols);
void pho_resize_level(const double* src, int rows, int cols, int level, double* dst);
void pho_gaussian_blur(double* img, int rows, int cols, int ksize, double sigma);
void pho_scharr(const double* src, int rows, int cols, int dx, int dy, double scale, double* dst);
}

#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_64F 6

namespace cv
{
enum { BORDER_DEFAULT = 4 };

struct Size
{
  Size() : width( 0 ), height( 0 ) {}
  Size( int w, int h ) : width( w ), height( h ) {}
  int width, height;
};

template< class T > struct DepthCode;
template<> struct DepthCode< unsigned char > { enum { value = CV_8U }; };
template<> struct DepthCode< unsigned short > { enum { value = CV_16U }; };
template<> struct DepthCode< float > { enum { value = CV_32F }; };
template<> struct DepthCode< double > { enum { value = CV_64F }; };

// Reference-counted dense row-major image: copies are shallow, like cv::Mat.  `data` / `step` are
// cv::Mat's raw view (first byte, row stride in BYTES); create() can leave padding behind every row
// so that strided inputs can be exercised (images made by the stand-in's own functions are continuous).
template< class T >
class Mat_
{
public:
  typedef T value_type;
  Mat_() : rows( 0 ), cols( 0 ), data( 0 ), step( 0 ) {}
  Mat_( int r, int c ) { create( r, c ); }
  void create( int r, int c, size_t rowBytes = 0 )
  {
    rows = r; cols = c;
    step = rowBytes ? rowBytes : size_t( c ) * sizeof( T );
    m_Buf.reset( new std::vector< T >( ( step * size_t( r ) + sizeof( T ) - 1 ) / sizeof( T ), T( 0 ) ) );
    data = reinterpret_cast< unsigned char * >( m_Buf->data() );
  }
  static Mat_ zeros( int r, int c ) { return Mat_( r, c ); }
  bool empty() const { return !m_Buf || m_Buf->empty(); }
  bool isContinuous() const { return step == size_t( cols ) * sizeof( T ); }
  int type() const { return DepthCode< T >::value; }
  T * ptr() { return m_Buf ? m_Buf->data() : 0; }
  const T * ptr() const { return m_Buf ? m_Buf->data() : 0; }
  T & operator()( int r, int c ) { return *reinterpret_cast< T * >( data + step * size_t( r ) + sizeof( T ) * size_t( c ) ); }
  const T & operator()( int r, int c ) const { return *reinterpret_cast< const T * >( data + step * size_t( r ) + sizeof( T ) * size_t( c ) ); }
  T & operator()( int i ) { return ( *m_Buf )[ size_t( i ) ]; }                 // continuous images only
  const T & operator()( int i ) const { return ( *m_Buf )[ size_t( i ) ]; }
  // Mat::convertTo(dst, rtype, alpha): dst = saturate_cast<U>( src * alpha ); floating destinations only here
  template< class U >
  void convertTo( Mat_< U > & dst, int /*rtype*/, double alpha = 1. ) const
  {
    Mat_< U > out( rows, cols );
    for( int r = 0; r < rows; r++ )
      for( int c = 0; c < cols; c++ ) out( r, c ) = U( double( ( *this )( r, c ) ) * alpha );
    dst = out;
  }
  int rows, cols;
  unsigned char * data;
  size_t step;
private:
  std::shared_ptr< std::vector< T > > m_Buf;
};

// MatExpr of the apps (FrameAlignment.cpp:77,81: `img = img * 1. / 1000.`): OpenCV folds `e * a` into a
// scale factor and `e / b` into `e * (1 / b)`, then evaluates once; the stand-in evaluates eagerly with
// the same roundings for that expression (x * 1. is exact).
template< class T > inline Mat_< T > operator*( const Mat_< T > & a, double s )
{
  Mat_< T > out( a.rows, a.cols );
  for( int r = 0; r < a.rows; r++ ) for( int c = 0; c < a.cols; c++ ) out( r, c ) = T( double( a( r, c ) ) * s );
  return out;
}
template< class T > inline Mat_< T > operator/( const Mat_< T > & a, double s ) { return a * ( 1. / s ); }

inline void resize( const Mat_< double > & src, Mat_< double > & dst, Size /*dsize*/, double fx, double /*fy*/ )
{
  const int level = int( std::lround( -std::log2( fx ) ) );
  int r, c;
  pho_level_size( src.rows, src.cols, level, &r, &c );
  Mat_< double > out( r, c );
  pho_resize_level( src.ptr(), src.rows, src.cols, level, out.ptr() );
  dst = out;
}

inline void GaussianBlur( const Mat_< double > & src, Mat_< double > & dst, Size ksize, double sigma )
{
  Mat_< double > out( src.rows, src.cols );
  for( int k = 0; k < src.rows * src.cols; k++ ) out( k ) = src( k );
  pho_gaussian_blur( out.ptr(), out.rows, out.cols, ksize.width, sigma );
  dst = out;
}

inline void blur( const Mat_< double > &, Mat_< double > &, Size ) {}   // ENABLE_BOX_FILTER_BLUR is 0 in the reference

inline void Scharr( const Mat_< double > & src, Mat_< double > & dst, int /*ddepth*/, int dx, int dy,
                    double scale, double /*delta*/, int /*borderType*/ )
{
  Mat_< double > out( src.rows, src.cols );
  pho_scharr( src.ptr(), src.rows, src.cols, dx, dy, scale, out.ptr() );
  dst = out;
}

// cv::mean of a single-channel image: only .val[0] is used by the reference (BiObjective.h:299)
struct Scalar { double val[4]; };
inline Scalar mean( const Mat_< double > & src )
{
  Scalar s; s.val[0] = s.val[1] = s.val[2] = s.val[3] = 0.;
  const size_t n = size_t( src.rows ) * size_t( src.cols );
  double acc = 0.;
  for( size_t k = 0; k < n; k++ ) acc += src( int( k ) );
  s.val[0] = n ? acc / double( n ) : 0.;
  return s;
}

template< class T >
inline void absdiff( const Mat_< T > &, const Mat_< T > &, Mat_< T > & ) {}
template< class T >
inline void imshow( const char *, const Mat_< T > & ) {}
inline int waitKey( int ) { return 0; }

// cv::FileStorage reader for the YAML-1.0 subset of the reference's config files:
// `key: scalar` and `key: [a, b, ...]`, keys may contain spaces and parentheses.
class FileNode
{
public:
  FileNode() {}
  explicit FileNode( const std::vector< double > & v ) : m_Values( v ) {}
  const std::vector< double > & values() const { return m_Values; }
private:
  std::vector< double > m_Values;
};
inline void operator>>( const FileNode & n, int & v ) { if( !n.values().empty() ) v = int( n.values()[0] ); }
inline void operator>>( const FileNode & n, bool & v ) { if( !n.values().empty() ) v = n.values()[0] != 0; }
inline void operator>>( const FileNode & n, double & v ) { if( !n.values().empty() ) v = n.values()[0]; }
inline void operator>>( const FileNode & n, std::vector< int > & v )
{
  v.clear();
  for( size_t i = 0; i < n.values().size(); i++ ) v.push_back( int( n.values()[i] ) );
}
inline void operator>>( const FileNode & n, std::vector< double > & v ) { v = n.values(); }

class FileStorage
{
public:
  enum { READ = 0 };
  FileStorage( const std::string & fileName, int /*flags*/ )
  {
    FILE * f = std::fopen( fileName.c_str(), "r" );
    if( !f ) return;
    char line[4096];
    while( std::fgets( line, sizeof( line ), f ) )
    {
      std::string s( line );
      if( s.empty() || s[0] == '%' || s[0] == '#' ) continue;
      const size_t colon = s.rfind( ": " ) != std::string::npos ? s.find( ": " ) : s.find( ':' );
      if( colon == std::string::npos ) continue;
      std::string key = s.substr( 0, colon ), rest = s.substr( colon + 1 );
      while( !key.empty() && ( key[ key.size() - 1 ] == ' ' ) ) key.erase( key.size() - 1 );
      std::vector< double > vals;
      const char * p = rest.c_str();
      while( *p )
      {
        char * end = 0;
        const double v = std::strtod( p, &end );
        if( end != p ) { vals.push_back( v ); p = end; }
        else p++;
      }
      m_Nodes[ key ] = FileNode( vals );
    }
    std::fclose( f );
  }
  FileNode operator[]( const char * key ) const
  {
    std::map< std::string, FileNode >::const_iterator it = m_Nodes.find( key );
    return it == m_Nodes.end() ? FileNode() : it->second;
  }
private:
  std::map< std::string, FileNode > m_Nodes;
};

} // namespace cv
#endif
