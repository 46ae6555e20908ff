// TEST-ONLY stand-in for the reference's phovo/include/CPhotoconsistencyOdometry.h.
//
// Neither OpenCV nor Eigen is installed in the build container, so the drop-in adapter
// (include/CPhotoconsistencyOdometryCuda.h) cannot be compiled against the real header here.
// This file declares the same names with the same access patterns the adapter relies on --
// cv::Mat_<T>::{rows, cols, data, step}, operator()(i,j) / operator()(i) on the matrix types, and the
// seven pure virtuals of the abstract solver (reference :153-178) -- and nothing else.  Inside the
// reference tree the adapter picks up the real header instead (same file name on the include path).
#ifndef PHOVO_TEST_SHIM_CPHOTOCONSISTENCY_ODOMETRY_H_
#define PHOVO_TEST_SHIM_CPHOTOCONSISTENCY_ODOMETRY_H_

#include <cstddef>
#include <cstring>
#include <vector>

namespace cv
{
// Dense 2-D array with a byte row stride, like cv::Mat_<T> (owning, for the tests).
template< class T >
class Mat_
{
public:
  Mat_() : rows( 0 ), cols( 0 ), data( 0 ), step( 0 ) {}
  Mat_( int r, int c, std::size_t rowBytes = 0 ) { create( r, c, rowBytes ); }
  Mat_( const Mat_ & o ) { copyFrom( o ); }
  Mat_ & operator=( const Mat_ & o ) { if( this != &o ) copyFrom( o ); return *this; }
  void create( int r, int c, std::size_t rowBytes = 0 )
  {
    rows = r; cols = c;
    step = rowBytes ? rowBytes : std::size_t( c ) * sizeof( T );
    m_Storage.assign( step * std::size_t( r ), 0 );
    data = m_Storage.empty() ? 0 : &m_Storage[0];
  }
  T & operator()( int i, int j ) { return *reinterpret_cast< T * >( data + step * std::size_t( i ) + sizeof( T ) * std::size_t( j ) ); }
  const T & operator()( int i, int j ) const { return *reinterpret_cast< const T * >( data + step * std::size_t( i ) + sizeof( T ) * std::size_t( j ) ); }
  int rows, cols;
  unsigned char * data;
  std::size_t step;
private:
  void copyFrom( const Mat_ & o )
  {
    rows = o.rows; cols = o.cols; step = o.step; m_Storage = o.m_Storage;
    data = m_Storage.empty() ? 0 : &m_Storage[0];
  }
  std::vector< unsigned char > m_Storage;
};
} // namespace cv

namespace phovo
{
namespace Numeric
{
template< class T, int R, int C >
class FixedRowMajor
{
public:
  FixedRowMajor() { for( int k = 0; k < R * C; k++ ) m_V[k] = T( 0 ); }
  T & operator()( int i, int j ) { return m_V[ i * C + j ]; }
  const T & operator()( int i, int j ) const { return m_V[ i * C + j ]; }
  T & operator()( int i ) { return m_V[i]; }
  const T & operator()( int i ) const { return m_V[i]; }
private:
  T m_V[ R * C ];
};
template< class T > class Matrix33RowMajor : public FixedRowMajor< T, 3, 3 > {};
template< class T > class Matrix44RowMajor : public FixedRowMajor< T, 4, 4 > {};
template< class T > class VectorCol6 : public FixedRowMajor< T, 6, 1 > {};
template< class T > class VectorCol4 : public FixedRowMajor< T, 4, 1 > {};
} // namespace Numeric

template< class TPixel, class TCoordinate >
class CPhotoconsistencyOdometry
{
public:
  typedef TPixel                PixelType;
  typedef cv::Mat_< PixelType > IntensityImageType;
  typedef TCoordinate                CoordinateType;
  typedef cv::Mat_< CoordinateType > DepthImageType;
  typedef Numeric::Matrix33RowMajor< CoordinateType > Matrix33Type;
  typedef Numeric::Matrix44RowMajor< CoordinateType > Matrix44Type;
  typedef Numeric::VectorCol6< CoordinateType >       Vector6Type;
  typedef Numeric::VectorCol4< CoordinateType >       Vector4Type;

  virtual ~CPhotoconsistencyOdometry() {}
  virtual void SetIntrinsicMatrix( const Matrix33Type & intrinsicMatrix ) = 0;
  virtual void SetSourceFrame( const IntensityImageType & intensityImage, const DepthImageType & depthImage ) = 0;
  virtual void SetTargetFrame( const IntensityImageType & intensityImage, const DepthImageType & depthImage ) = 0;
  virtual void SetInitialStateVector( const Vector6Type & initialStateVector ) = 0;
  virtual void Optimize() = 0;
  virtual Vector6Type GetOptimalStateVector() const = 0;
  virtual Matrix44Type GetOptimalRigidTransformationMatrix() const = 0;
};
} // namespace phovo

#endif
