// TEST INFRASTRUCTURE.  C entry points around the reference's OWN Ceres-based solver class
// (CPhotoconsistencyOdometryCeres.h) with its residual functor (CE:156-269) and sampler
// (third_party/sample.h, third_party/jet_extras.h), all compiled UNMODIFIED from /root/reference
// into oracle/_ref/libphovo_ref.so.  Ceres itself is absent from this image: oracle/shim/ceres/ is a
// stand-in whose AutoDiffCostFunction evaluates the functor on double and on Jet<double,6>, and
// whose ceres::Solve() calls the hook below.  The hook
//   * CAPTURE: evaluates the problem of the requested level at a caller-supplied state and keeps
//     residuals + N x 6 Jacobian (what pins the oracle's and the CUDA path's functor), or
//   * SOLVE: runs the oracle's restated Levenberg-Marquardt loop (pho_lm_minimize) with the
//     reference functor as the evaluator -- "reference functor + restated LM".
// Ceres' own trust-region code never runs here: the LM trajectory stays unpinned against Ceres.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <vector>

#include "CPhotoconsistencyOdometryCeres.h"
#include "../phovo_oracle.h"

namespace
{
typedef phovo::Ceres::CPhotoconsistencyOdometryCeres< unsigned char, double > Solver;

struct Ref
{
  Solver solver;
  std::vector< phovo_iter_stats > log;
  std::vector< int > itersPerCall;
};

enum HookMode { HOOK_NONE, HOOK_CAPTURE, HOOK_SOLVE };
struct HookState
{
  HookMode mode;
  Ref * ref;
  // capture
  int wantResiduals;            // the level is identified by its pixel count
  const double * state;
  double * residuals; double * jacobian;
  int captured;
  // solve
  int level;
} g_Hook = { HOOK_NONE, 0, 0, 0, 0, 0, 0, 0 };

struct EvalCtx { ceres::Problem * problem; std::vector< double > res, jac; };

void evalForLm( void * user, const double x[6], int wantJac, double H[21], double g[6], double * cost, int * count )
{
  EvalCtx * c = static_cast< EvalCtx * >( user );
  const int n = c->problem->cost()->num_residuals();
  const double * params[1] = { x };
  double * jacs[1] = { c->jac.data() };
  c->problem->cost()->Evaluate( params, c->res.data(), wantJac ? jacs : 0 );
  *count = -1;                  // the functor does not count its valid pixels
  if( wantJac ) pho_normal_equations_rowmajor( c->jac.data(), c->res.data(), size_t( n ), H, g, cost );
  else { double s = 0; for( int i = 0; i < n; i++ ) s += c->res[ size_t( i ) ] * c->res[ size_t( i ) ]; *cost = 0.5 * s; }
}

phovo_iter_stats * newEntry( void * user )
{
  Ref * r = static_cast< Ref * >( user );
  phovo_iter_stats e; std::memset( &e, 0, sizeof( e ) );
  r->log.push_back( e );
  return &r->log.back();
}

void hook( const ceres::Solver::Options & options, ceres::Problem * problem, ceres::Solver::Summary * )
{
  const int n = problem->cost()->num_residuals();
  if( g_Hook.mode == HOOK_CAPTURE )
  {
    if( n != g_Hook.wantResiduals || g_Hook.captured ) return;
    const double * params[1] = { g_Hook.state };
    double * jacs[1] = { g_Hook.jacobian };
    problem->cost()->Evaluate( params, g_Hook.residuals, g_Hook.jacobian ? jacs : 0 );
    g_Hook.captured = 1;
  }
  else if( g_Hook.mode == HOOK_SOLVE )
  {
    EvalCtx ctx; ctx.problem = problem; ctx.res.resize( size_t( n ) ); ctx.jac.resize( size_t( n ) * 6 );
    pho_lm_options opt;                                                  // CE:464-477
    opt.max_num_iterations = options.max_num_iterations;
    opt.function_tolerance = options.function_tolerance;
    opt.gradient_tolerance = options.gradient_tolerance;
    opt.parameter_tolerance = options.parameter_tolerance;
    opt.initial_trust_region_radius = options.initial_trust_region_radius;
    opt.max_trust_region_radius = options.max_trust_region_radius;
    opt.min_trust_region_radius = options.min_trust_region_radius;
    opt.min_relative_decrease = options.min_relative_decrease;
    // levels are solved coarse to fine; the hook numbers them by call order (the caller maps back)
    const int it = pho_lm_minimize( &opt, evalForLm, &ctx, newEntry, g_Hook.ref, g_Hook.level, problem->parameters() );
    g_Hook.ref->itersPerCall.push_back( it );
    g_Hook.level += 1;
  }
}

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

// the reference prints summary.BriefReport() per level (CE:495): keep the test output clean
struct QuietCout
{
  QuietCout() : m_Old( std::cout.rdbuf( m_Sink.rdbuf() ) ) {}
  ~QuietCout() { std::cout.rdbuf( m_Old ); }
  std::ostringstream m_Sink; std::streambuf * m_Old;
};

void runOptimize( Ref * r )
{
  QuietCout quiet;
  ceres::solveHook() = hook;
  r->solver.Optimize();
  ceres::solveHook() = 0;
  g_Hook.mode = HOOK_NONE;
}
} // namespace

extern "C" {

void * refce_create() { return new Ref(); }
void refce_destroy( void * h ) { delete static_cast< Ref * >( h ); }
void refce_read_config( void * h, const char * path ) { static_cast< Ref * >( h )->solver.ReadConfigurationFile( path ); }
void refce_set_intrinsics( void * h, const double K[9] )
{
  Solver::Matrix33Type M;
  for( int i = 0; i < 3; i++ ) for( int j = 0; j < 3; j++ ) M( i, j ) = K[ 3 * i + j ];
  static_cast< Ref * >( h )->solver.SetIntrinsicMatrix( M );
}
void refce_set_source( void * h, const uint8_t * gray, const double * depth, int rows, int cols )
{
  static_cast< Ref * >( h )->solver.SetSourceFrame( wrapGray( gray, rows, cols ), wrapDepth( depth, rows, cols ) );
}
void refce_set_target( void * h, const uint8_t * gray, int rows, int cols )
{
  static_cast< Ref * >( h )->solver.SetTargetFrame( wrapGray( gray, rows, cols ), wrapDepth( 0, rows, cols ) );
}
void refce_set_initial_state( void * h, const double s[6] )
{
  Solver::Vector6Type v;
  for( int i = 0; i < 6; i++ ) v( i ) = s[i];
  static_cast< Ref * >( h )->solver.SetInitialStateVector( v );
}
void refce_get_state( void * h, double s[6] )
{
  const Solver::Vector6Type v = static_cast< Ref * >( h )->solver.GetOptimalStateVector();
  for( int i = 0; i < 6; i++ ) s[i] = v( i );
}
// Evaluates the reference functor of the level with `num_pixels` residuals at `state` through the
// reference's own Optimize() (which builds the ceres::Problem from ITS pyramids, CE:441-460).
// residuals: num_pixels; jacobian: num_pixels x 6 row-major or NULL (then T = double is used).
// Returns 1 if a level of that size with max_num_iterations > 0 exists.
int refce_evaluate( void * h, int num_pixels, const double state[6], double * residuals, double * jacobian )
{
  Ref * r = static_cast< Ref * >( h );
  g_Hook.mode = HOOK_CAPTURE; g_Hook.ref = r; g_Hook.wantResiduals = num_pixels; g_Hook.state = state;
  g_Hook.residuals = residuals; g_Hook.jacobian = jacobian; g_Hook.captured = 0;
  runOptimize( r );
  return g_Hook.captured;
}
// Optimize() with the restated LM behind ceres::Solve.  Log entries carry the call order in `level`
// (0 = coarsest solved level).
void refce_optimize( void * h )
{
  Ref * r = static_cast< Ref * >( h );
  r->log.clear(); r->itersPerCall.clear();
  g_Hook.mode = HOOK_SOLVE; g_Hook.ref = r; g_Hook.level = 0;
  runOptimize( r );
}
int refce_num_iter_stats( void * h ) { return int( static_cast< Ref * >( h )->log.size() ); }
void refce_get_iter_stats( void * h, int index, phovo_iter_stats * out ) { *out = static_cast< Ref * >( h )->log[ size_t( index ) ]; }

} // extern "C"
