// TEST INFRASTRUCTURE.  Stand-in for Ceres' dual number ceres::Jet<T,N> (ceres/jet.h of Ceres
// Solver >= 1.6, which the reference's CMakeLists.txt:50 requires and this image does not have), so
// that the reference's residual functor (CPhotoconsistencyOdometryCeres.h:156-269) and its sampler
// (third_party/sample.h, third_party/jet_extras.h) compile UNMODIFIED into oracle/_ref.
//
// A Jet is a + sum_k v[k] e_k with e_i e_j = 0.  Only the operations the functor uses exist; each
// follows the dual-number algebra as Ceres states it in jet.h (product rule, quotient through the
// reciprocal of the denominator's scalar part, sin/cos, pow(double, Jet)).  Rounding-level
// differences against a particular Ceres release are possible (e.g. whether h.a of a quotient is
// f.a * (1 / g.a) or f.a / g.a changed between releases); the parity tests allow 1e-11.
#ifndef PHOVO_SHIM_CERES_JET_H_
#define PHOVO_SHIM_CERES_JET_H_

#include <cmath>
#include "Eigen/Core"

namespace ceres
{
template< typename T, int N >
struct Jet
{
  enum { DIMENSION = N };
  Jet() : a() {}
  // not explicit: the functor writes T( 1. ), T( c ), T( image( r, c ) )
  Jet( const T & value ) : a( value ) {}
  Jet( const T & value, int k ) : a( value ) { v( k ) = T( 1. ); }

  T a;                           // scalar part
  Eigen::Matrix< T, N, 1 > v;    // infinitesimal part (zero-initialised by the stand-in matrix)
};

#define PHOVO_JET_FOR for( int k = 0; k < N; k++ )

template< typename T, int N > inline Jet< T, N > operator+( const Jet< T, N > & f, const Jet< T, N > & g )
{ Jet< T, N > h( f.a + g.a ); PHOVO_JET_FOR h.v( k ) = f.v( k ) + g.v( k ); return h; }
template< typename T, int N > inline Jet< T, N > operator-( const Jet< T, N > & f, const Jet< T, N > & g )
{ Jet< T, N > h( f.a - g.a ); PHOVO_JET_FOR h.v( k ) = f.v( k ) - g.v( k ); return h; }
template< typename T, int N > inline Jet< T, N > operator-( const Jet< T, N > & f )
{ Jet< T, N > h( -f.a ); PHOVO_JET_FOR h.v( k ) = -f.v( k ); return h; }
// (a + u)(b + w) = ab + (a w + b u)
template< typename T, int N > inline Jet< T, N > operator*( const Jet< T, N > & f, const Jet< T, N > & g )
{ Jet< T, N > h( f.a * g.a ); PHOVO_JET_FOR h.v( k ) = f.a * g.v( k ) + f.v( k ) * g.a; return h; }
// (a + u)/(b + w) = a/b + (u - (a/b) w)/b, with 1/b formed once
template< typename T, int N > inline Jet< T, N > operator/( const Jet< T, N > & f, const Jet< T, N > & g )
{
  const T g_a_inverse = T( 1.0 ) / g.a;
  const T f_a_by_g_a = f.a * g_a_inverse;
  Jet< T, N > h( f_a_by_g_a );
  PHOVO_JET_FOR h.v( k ) = ( f.v( k ) - f_a_by_g_a * g.v( k ) ) * g_a_inverse;
  return h;
}
// mixed with scalars
template< typename T, int N > inline Jet< T, N > operator+( const Jet< T, N > & f, T s ) { Jet< T, N > h( f ); h.a = f.a + s; return h; }
template< typename T, int N > inline Jet< T, N > operator+( T s, const Jet< T, N > & f ) { Jet< T, N > h( f ); h.a = s + f.a; return h; }
template< typename T, int N > inline Jet< T, N > operator-( const Jet< T, N > & f, T s ) { Jet< T, N > h( f ); h.a = f.a - s; return h; }
template< typename T, int N > inline Jet< T, N > operator-( T s, const Jet< T, N > & f )
{ Jet< T, N > h( s - f.a ); PHOVO_JET_FOR h.v( k ) = -f.v( k ); return h; }
template< typename T, int N > inline Jet< T, N > operator*( const Jet< T, N > & f, T s )
{ Jet< T, N > h( f.a * s ); PHOVO_JET_FOR h.v( k ) = f.v( k ) * s; return h; }
template< typename T, int N > inline Jet< T, N > operator*( T s, const Jet< T, N > & f )
{ Jet< T, N > h( f.a * s ); PHOVO_JET_FOR h.v( k ) = f.v( k ) * s; return h; }
template< typename T, int N > inline Jet< T, N > operator/( const Jet< T, N > & f, T s )
{ const T inv = T( 1.0 ) / s; Jet< T, N > h( f.a * inv ); PHOVO_JET_FOR h.v( k ) = f.v( k ) * inv; return h; }
template< typename T, int N > inline Jet< T, N > operator/( T s, const Jet< T, N > & g )
{ const T minus_s_g_a_inverse2 = -s / ( g.a * g.a ); Jet< T, N > h( s / g.a ); PHOVO_JET_FOR h.v( k ) = g.v( k ) * minus_s_g_a_inverse2; return h; }

// comparisons look at the scalar part only (CE:246-247)
#define PHOVO_JET_CMP( op )                                                                                              \
  template< typename T, int N > inline bool operator op( const Jet< T, N > & f, const Jet< T, N > & g ) { return f.a op g.a; } \
  template< typename T, int N > inline bool operator op( const Jet< T, N > & f, const T & s ) { return f.a op s; }             \
  template< typename T, int N > inline bool operator op( const T & s, const Jet< T, N > & g ) { return s op g.a; }
PHOVO_JET_CMP( < )
PHOVO_JET_CMP( <= )
PHOVO_JET_CMP( > )
PHOVO_JET_CMP( >= )
PHOVO_JET_CMP( == )
PHOVO_JET_CMP( != )
#undef PHOVO_JET_CMP

template< typename T, int N > inline Jet< T, N > sin( const Jet< T, N > & f )
{ const T c = std::cos( f.a ); Jet< T, N > h( std::sin( f.a ) ); PHOVO_JET_FOR h.v( k ) = c * f.v( k ); return h; }
template< typename T, int N > inline Jet< T, N > cos( const Jet< T, N > & f )
{ const T s = -std::sin( f.a ); Jet< T, N > h( std::cos( f.a ) ); PHOVO_JET_FOR h.v( k ) = s * f.v( k ); return h; }
// pow(f, a + u) = f^a + log(f) f^a u     (CE:163-168: pow( 2, T( level ) ))
template< typename T, int N > inline Jet< T, N > pow( double f, const Jet< T, N > & g )
{
  const T tmp = std::pow( f, g.a );
  const T d = std::log( f ) * tmp;
  Jet< T, N > h( tmp ); PHOVO_JET_FOR h.v( k ) = d * g.v( k ); return h;
}
#undef PHOVO_JET_FOR

} // namespace ceres
#endif
