// TEST INFRASTRUCTURE.  C entry points around the reference's OWN photometric + depth solver
// (CPhotoconsistencyOdometryBiObjective.h), compiled unmodified from /root/reference/phovo/include
// into oracle/_ref/libphovo_ref.so together with ref_driver.cpp.  Separate translation unit because
// both reference headers define the same configuration macros.
#include <cstdint>
#include <cstring>
#include <vector>

#include "CPhotoconsistencyOdometryBiObjective.h"

namespace
{
typedef phovo::Analytic::CPhotoconsistencyOdometryBiObjective< unsigned char, double > Solver;

struct IterRecord { int n; double H[36]; double g[6]; };
std::vector< IterRecord > * g_Log = 0;

// Optimize() forms J^T r first (BiObjective.h:622), then J^T J (:623-624): one record per iteration.
void tap( int rowsA, int colsA, int colsB, const double * r )
{
  if( !g_Log || rowsA != 6 || colsA <= 6 ) return;
  if( colsB == 1 )
  {
    IterRecord rec; std::memset( &rec, 0, sizeof( rec ) );
    rec.n = colsA;
    for( int k = 0; k < 6; k++ ) rec.g[k] = r[k];
    g_Log->push_back( rec );
  }
  else if( colsB == 6 && !g_Log->empty() )
    for( int k = 0; k < 36; k++ ) g_Log->back().H[k] = r[k];
}

struct Ref { Solver solver; std::vector< IterRecord > log; };

cv::Mat_< unsigned char > wrapGray( const uint8_t * p, int rows, int cols )
{
  cv::Mat_< unsigned char > m( rows, cols );
  std::memcpy( m.ptr(), p, size_t( rows ) * size_t( cols ) );
  return m;
}
cv::Mat_< double > wrapDepth( const double * p, int rows, int cols )
{
  cv::Mat_< double > m( rows, cols );
  std::memcpy( m.ptr(), p, sizeof( double ) * size_t( rows ) * size_t( cols ) );
  return m;
}
} // namespace

extern "C" {

void * refbi_create() { return new Ref(); }
void refbi_destroy( void * h ) { delete static_cast< Ref * >( h ); }
void refbi_read_config( void * h, const char * path ) { static_cast< Ref * >( h )->solver.ReadConfigurationFile( path ); }
void refbi_set_intrinsics( void * h, const double K[9] )
{
  Solver::Matrix33Type M;
  for( int i = 0; i < 3; i++ ) for( int j = 0; j < 3; j++ ) M( i, j ) = K[ 3 * i + j ];
  static_cast< Ref * >( h )->solver.SetIntrinsicMatrix( M );
}
void refbi_set_source( void * h, const uint8_t * gray, const double * depth, int rows, int cols )
{
  static_cast< Ref * >( h )->solver.SetSourceFrame( wrapGray( gray, rows, cols ), wrapDepth( depth, rows, cols ) );
}
void refbi_set_target( void * h, const uint8_t * gray, const double * depth, int rows, int cols )
{
  static_cast< Ref * >( h )->solver.SetTargetFrame( wrapGray( gray, rows, cols ), wrapDepth( depth, rows, cols ) );
}
void refbi_set_initial_state( void * h, const double s[6] )
{
  Solver::Vector6Type v;
  for( int i = 0; i < 6; i++ ) v( i ) = s[i];
  static_cast< Ref * >( h )->solver.SetInitialStateVector( v );
}
void refbi_optimize( void * h )
{
  Ref * r = static_cast< Ref * >( h );
  r->log.clear();
  g_Log = &r->log;
  Eigen::productTap() = tap;
  r->solver.Optimize();
  Eigen::productTap() = 0;
  g_Log = 0;
}
void refbi_get_state( void * h, double s[6] )
{
  const Solver::Vector6Type v = static_cast< Ref * >( h )->solver.GetOptimalStateVector();
  for( int i = 0; i < 6; i++ ) s[i] = v( i );
}
int refbi_num_iterations( void * h ) { return int( static_cast< Ref * >( h )->log.size() ); }
// n = rows of the stacked system (2 x pixels of the level); H row-major 6x6; g = J^T r
void refbi_get_iteration( void * h, int index, int * n, double H[36], double g[6] )
{
  const IterRecord & rec = static_cast< Ref * >( h )->log[ size_t( index ) ];
  *n = rec.n;
  std::memcpy( H, rec.H, sizeof( rec.H ) );
  std::memcpy( g, rec.g, sizeof( rec.g ) );
}

} // extern "C"
