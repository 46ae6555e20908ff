// Headless mirror of apps/PhotoconsistencyFrameAlignment/PhotoconsistencyFrameAlignment.cpp:54-105
// and of the loop of apps/PhotoconsistencyVisualOdometry/PhotoconsistencyVisualOdometry.cpp:212-259,
// written against the drop-in adapter exactly as the apps are written against the analytic solver
// (same calls, same order).  cv::imread / imshow are replaced by raw binary files because OpenCV is
// not installed here; tests/test_gpu_cpp_adapter.py writes the inputs and checks the outputs
// against the CPU oracle.
//
//   frame_alignment_app align  <config.yml> <rows> <cols> <fx> <fy> <ox> <oy> <gray0.u8> <depth0.f64> <gray1.u8> <depth1.f64> [row_padding_bytes]
//   frame_alignment_app vo     <config.yml> <rows> <cols> <fx> <fy> <ox> <oy> <num_frames> <gray_%d.u8 pattern> <depth_%d.f64 pattern> <trajectory_out>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "CPhotoconsistencyOdometryCuda.h"

typedef double CoordinateType;
typedef unsigned char PixelType;
typedef phovo::Cuda::CPhotoconsistencyOdometryCuda< PixelType, CoordinateType > OdometryType;
typedef OdometryType::Matrix33Type Matrix33Type;
typedef OdometryType::Matrix44Type Matrix44Type;
typedef OdometryType::Vector6Type Vector6Type;
typedef OdometryType::IntensityImageType IntensityImageType;
typedef OdometryType::DepthImageType DepthImageType;

template< class ImageType, class T >
static bool readRaw( const std::string & path, int rows, int cols, size_t padding, ImageType & img )
{
  img.create( rows, cols, size_t( cols ) * sizeof( T ) + padding );
  FILE * f = std::fopen( path.c_str(), "rb" );
  if( !f ) { std::cerr << "cannot open " << path << std::endl; return false; }
  bool ok = true;
  for( int r = 0; r < rows && ok; r++ )
    ok = std::fread( img.data + img.step * size_t( r ), sizeof( T ), size_t( cols ), f ) == size_t( cols );
  std::fclose( f );
  return ok;
}

static void mul44( const double * A, const double * B, double * C )
{
  for( int i = 0; i < 4; i++ )
    for( int j = 0; j < 4; j++ )
    {
      double s = 0;
      for( int k = 0; k < 4; k++ ) s += A[ 4 * i + k ] * B[ 4 * k + j ];
      C[ 4 * i + j ] = s;
    }
}

// inverse of a rigid transformation [R t; 0 1]
static void rigidInverse( const Matrix44Type & Rt, double * inv )
{
  for( int i = 0; i < 3; i++ )
  {
    for( int j = 0; j < 3; j++ ) inv[ 4 * i + j ] = Rt( j, i );
    inv[ 4 * i + 3 ] = -( Rt( 0, i ) * Rt( 0, 3 ) + Rt( 1, i ) * Rt( 1, 3 ) + Rt( 2, i ) * Rt( 2, 3 ) );
  }
  inv[12] = 0; inv[13] = 0; inv[14] = 0; inv[15] = 1;
}

// unit quaternion (x y z w) of the rotation block, the convention of Eigen::Quaterniond(R)
static void quaternionOf( const double * P, double * q )
{
  const double m00 = P[0], m01 = P[1], m02 = P[2], m10 = P[4], m11 = P[5], m12 = P[6], m20 = P[8], m21 = P[9], m22 = P[10];
  double t = m00 + m11 + m22;
  if( t > 0 )
  {
    t = std::sqrt( t + 1.0 );
    q[3] = 0.5 * t; t = 0.5 / t;
    q[0] = ( m21 - m12 ) * t; q[1] = ( m02 - m20 ) * t; q[2] = ( m10 - m01 ) * t;
  }
  else
  {
    int i = 0;
    if( m11 > m00 ) i = 1;
    if( m22 > ( i == 0 ? m00 : m11 ) ) i = 2;
    const int j = ( i + 1 ) % 3, k = ( j + 1 ) % 3;
    const double * M = P;
    t = std::sqrt( M[ 4 * i + i ] - M[ 4 * j + j ] - M[ 4 * k + k ] + 1.0 );
    q[i] = 0.5 * t; t = 0.5 / t;
    q[3] = ( M[ 4 * k + j ] - M[ 4 * j + k ] ) * t;
    q[j] = ( M[ 4 * j + i ] + M[ 4 * i + j ] ) * t;
    q[k] = ( M[ 4 * k + i ] + M[ 4 * i + k ] ) * t;
  }
}

int main( int argc, char ** argv )
{
  if( argc < 12 ) { std::cerr << "usage: see the header of this file" << std::endl; return -1; }
  const std::string mode = argv[1];
  const int rows = std::atoi( argv[3] ), cols = std::atoi( argv[4] );
  Matrix33Type intrinsicMatrix;
  intrinsicMatrix( 0, 0 ) = std::atof( argv[5] ); intrinsicMatrix( 1, 1 ) = std::atof( argv[6] );
  intrinsicMatrix( 0, 2 ) = std::atof( argv[7] ); intrinsicMatrix( 1, 2 ) = std::atof( argv[8] );
  intrinsicMatrix( 2, 2 ) = 1.;
  try
  {
    OdometryType photoconsistencyOdometry;
    Vector6Type stateVector;   // x,y,z,yaw,pitch,roll = 0
    photoconsistencyOdometry.ReadConfigurationFile( std::string( argv[2] ) );
    photoconsistencyOdometry.SetIntrinsicMatrix( intrinsicMatrix );
    std::cout.precision( 17 );
    if( mode == "align" )
    {
      if( argc < 13 ) return -1;
      const size_t padding = argc > 13 ? size_t( std::atoi( argv[13] ) ) : 0;
      IntensityImageType imgGray0, imgGray1;
      DepthImageType imgDepth0, imgDepth1;
      if( !readRaw< IntensityImageType, PixelType >( argv[9], rows, cols, padding, imgGray0 ) ) return 2;
      if( !readRaw< DepthImageType, CoordinateType >( argv[10], rows, cols, padding, imgDepth0 ) ) return 2;
      if( !readRaw< IntensityImageType, PixelType >( argv[11], rows, cols, padding, imgGray1 ) ) return 2;
      if( !readRaw< DepthImageType, CoordinateType >( argv[12], rows, cols, padding, imgDepth1 ) ) return 2;
      photoconsistencyOdometry.SetSourceFrame( imgGray0, imgDepth0 );
      photoconsistencyOdometry.SetTargetFrame( imgGray1, imgDepth1 );
      photoconsistencyOdometry.SetInitialStateVector( stateVector );
      photoconsistencyOdometry.Optimize();
      float setupMs = 0, optimizeMs = 0;
      photoconsistencyOdometry.GetTimings( setupMs, optimizeMs );
      std::cout << "Time = " << optimizeMs * 1e-3 << " sec." << std::endl;
      const Vector6Type s = photoconsistencyOdometry.GetOptimalStateVector();
      std::cout << "state";
      for( int i = 0; i < 6; i++ ) std::cout << " " << s( i );
      std::cout << std::endl;
      const Matrix44Type Rt = photoconsistencyOdometry.GetOptimalRigidTransformationMatrix();
      std::cout << "Rt";
      for( int i = 0; i < 4; i++ ) for( int j = 0; j < 4; j++ ) std::cout << " " << Rt( i, j );
      std::cout << std::endl;
      std::cout << "iterations " << photoconsistencyOdometry.GetIterationStats().size() << std::endl;
      // FrameAlignment.cpp:106-110: warp the source with the result and compare with the target
      IntensityImageType warpedImage;
      photoconsistencyOdometry.WarpImage( imgGray0, imgDepth0, warpedImage, Rt, intrinsicMatrix );
      unsigned long long sumAbsDiff = 0, written = 0;
      for( int r = 0; r < rows; r++ )
        for( int c = 0; c < cols; c++ )
        {
          const int w = warpedImage( r, c ), t = imgGray1( r, c );
          if( w ) { written++; sumAbsDiff += (unsigned long long)( w > t ? w - t : t - w ); }
        }
      std::cout << "warped " << written << " " << sumAbsDiff << std::endl;
      return 0;
    }
    if( mode == "vo" )
    {
      const int numFrames = std::atoi( argv[9] );
      const std::string grayPattern = argv[10], depthPattern = argv[11];
      FILE * trajectory = std::fopen( argv[12], "w" );
      if( !trajectory ) return 2;
      double pose[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
      char name[1024];
      IntensityImageType previousIntensity, currentIntensity;
      DepthImageType previousDepth, currentDepth;
      std::snprintf( name, sizeof( name ), grayPattern.c_str(), 0 );
      if( !readRaw< IntensityImageType, PixelType >( name, rows, cols, 0, previousIntensity ) ) return 2;
      std::snprintf( name, sizeof( name ), depthPattern.c_str(), 0 );
      if( !readRaw< DepthImageType, CoordinateType >( name, rows, cols, 0, previousDepth ) ) return 2;
      for( int k = 1; k < numFrames; k++ )
      {
        std::snprintf( name, sizeof( name ), grayPattern.c_str(), k );
        if( !readRaw< IntensityImageType, PixelType >( name, rows, cols, 0, currentIntensity ) ) return 2;
        std::snprintf( name, sizeof( name ), depthPattern.c_str(), k );
        if( !readRaw< DepthImageType, CoordinateType >( name, rows, cols, 0, currentDepth ) ) return 2;
        // VisualOdometry.cpp:222-224: the state vector handed in is always zero (:175)
        if( k == 1 ) photoconsistencyOdometry.SetSourceFrame( previousIntensity, previousDepth );
        else photoconsistencyOdometry.PromoteTargetToSource( previousDepth );   // frame k-1's pyramid is already on the device
        photoconsistencyOdometry.SetTargetFrame( currentIntensity, currentDepth );
        photoconsistencyOdometry.SetInitialStateVector( stateVector );
        photoconsistencyOdometry.Optimize();
        const Matrix44Type Rt = photoconsistencyOdometry.GetOptimalRigidTransformationMatrix();
        double inv[16], next[16], q[4];
        rigidInverse( Rt, inv );                       // VisualOdometry.cpp:234 pose = pose * Rt.inverse()
        mul44( pose, inv, next );
        for( int i = 0; i < 16; i++ ) pose[i] = next[i];
        quaternionOf( pose, q );                       // :235-237
        std::fprintf( trajectory, "%d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", k, pose[3], pose[7], pose[11], q[0], q[1], q[2], q[3] );   // :240-243
        previousIntensity = currentIntensity;          // :256-257
        previousDepth = currentDepth;
      }
      std::fclose( trajectory );
      return 0;
    }
  }
  catch( const std::exception & e )
  {
    std::cerr << "error: " << e.what() << std::endl;
    return 3;
  }
  return -1;
}
