// TEST INFRASTRUCTURE.  C entry points around the reference's OWN analytic solver, compiled
// unmodified from /root/reference/phovo/include (see oracle/Makefile -> oracle/_ref/libphovo_ref.so).
// Only OpenCV / Eigen are replaced by the stand-ins in this directory.  Used by
// tests/test_oracle_vs_reference_source.py and tests/golden/make_reference_golden.py to pin the
// oracle's restatement against the reference source itself.
#include <cstdint>
#include <cstring>
#include <vector>

#include "CPhotoconsistencyOdometryAnalytic.h"

namespace
{
typedef phovo::Analytic::CPhotoconsistencyOdometryAnalytic< unsigned char, double > Solver;

struct IterRecord { int n; double H[36]; double g[6]; bool haveH, haveG; };
std::vector< IterRecord > * g_Log = 0;

// Optimize() forms J^T r first (AN:538), then J^T J (AN:540): one record per iteration.
void tap( int rowsA, int colsA, int colsB, const double * r )
{
  if( !g_Log || rowsA != 6 || colsA <= 6 ) return;
  if( colsB == 1 )
  {
    IterRecord rec; std::memset( &rec, 0, sizeof( rec ) );
    rec.n = colsA; rec.haveG = true;
    for( int k = 0; k < 6; k++ ) rec.g[k] = r[k];
    g_Log->push_back( rec );
  }
  else if( colsB == 6 && !g_Log->empty() )
  {
    IterRecord & rec = g_Log->back();
    for( int k = 0; k < 36; k++ ) rec.H[k] = r[k];
    rec.haveH = true;
  }
}

struct Ref
{
  Solver solver;
  std::vector< IterRecord > log;
};

cv::Mat_< unsigned char > wrapGray( const uint8_t * p, int rows, int cols )
{
  cv::Mat_< unsigned char > m( rows, cols );
  std::memcpy( m.ptr(), p, size_t( rows ) * size_t( cols ) );
  return m;
}
cv::Mat_< double > wrapDepth( const double * p, int rows, int cols )
{
  cv::Mat_< double > m( rows, cols );
  if( p ) std::memcpy( m.ptr(), p, sizeof( double ) * size_t( rows ) * size_t( cols ) );
  return m;
}
} // namespace

extern "C" {

void * ref_create() { return new Ref(); }
void ref_destroy( void * h ) { delete static_cast< Ref * >( h ); }
void ref_read_config( void * h, const char * path ) { static_cast< Ref * >( h )->solver.ReadConfigurationFile( path ); }
void ref_set_depth_range( void * h, double lo, double hi )
{
  static_cast< Ref * >( h )->solver.SetMinDepth( lo );
  static_cast< Ref * >( h )->solver.SetMaxDepth( hi );
}
void ref_set_intrinsics( void * h, const double K[9] )
{
  Solver::Matrix33Type M;
  for( int i = 0; i < 3; i++ ) for( int j = 0; j < 3; j++ ) M( i, j ) = K[ 3 * i + j ];
  static_cast< Ref * >( h )->solver.SetIntrinsicMatrix( M );
}
void ref_set_source( void * h, const uint8_t * gray, const double * depth, int rows, int cols )
{
  static_cast< Ref * >( h )->solver.SetSourceFrame( wrapGray( gray, rows, cols ), wrapDepth( depth, rows, cols ) );
}
void ref_set_target( void * h, const uint8_t * gray, int rows, int cols )
{
  static_cast< Ref * >( h )->solver.SetTargetFrame( wrapGray( gray, rows, cols ), wrapDepth( 0, rows, cols ) );
}
void ref_set_initial_state( void * h, const double s[6] )
{
  Solver::Vector6Type v;
  for( int i = 0; i < 6; i++ ) v( i ) = s[i];
  static_cast< Ref * >( h )->solver.SetInitialStateVector( v );
}
void ref_optimize( void * h )
{
  Ref * r = static_cast< Ref * >( h );
  r->log.clear();
  g_Log = &r->log;
  Eigen::productTap() = tap;
  r->solver.Optimize();
  Eigen::productTap() = 0;
  g_Log = 0;
}
void ref_get_state( void * h, double s[6] )
{
  const Solver::Vector6Type v = static_cast< Ref * >( h )->solver.GetOptimalStateVector();
  for( int i = 0; i < 6; i++ ) s[i] = v( i );
}
void ref_get_rt( void * h, double rt[16] )
{
  const Solver::Matrix44Type M = static_cast< Ref * >( h )->solver.GetOptimalRigidTransformationMatrix();
  for( int i = 0; i < 4; i++ ) for( int j = 0; j < 4; j++ ) rt[ 4 * i + j ] = M( i, j );
}
int ref_num_iterations( void * h ) { return int( static_cast< Ref * >( h )->log.size() ); }
// n = pixels of the level the iteration ran on; H row-major 6x6; g = J^T r
void ref_get_iteration( void * h, int index, int * n, double H[36], double g[6] )
{
  const IterRecord & rec = static_cast< Ref * >( h )->log[ size_t( index ) ];
  *n = rec.n;
  std::memcpy( H, rec.H, sizeof( rec.H ) );
  std::memcpy( g, rec.g, sizeof( rec.g ) );
}
// BASE:73-134 warpImage (the apps' post-hoc visualisation), level 0
void ref_warp_image( const uint8_t * gray, const double * depth, int rows, int cols, const double rt[16], const double K[9], uint8_t * out )
{
  phovo::Numeric::Matrix44RowMajor< double > Rt;
  phovo::Numeric::Matrix33RowMajor< double > Km;
  for( int i = 0; i < 4; i++ ) for( int j = 0; j < 4; j++ ) Rt( i, j ) = rt[ 4 * i + j ];
  for( int i = 0; i < 3; i++ ) for( int j = 0; j < 3; j++ ) Km( i, j ) = K[ 3 * i + j ];
  cv::Mat_< unsigned char > warped;
  phovo::warpImage< unsigned char, double >( wrapGray( gray, rows, cols ), wrapDepth( depth, rows, cols ), warped, Rt, Km );
  std::memcpy( out, warped.ptr(), size_t( rows ) * size_t( cols ) );
}

} // extern "C"
