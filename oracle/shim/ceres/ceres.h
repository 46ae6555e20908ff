// TEST INFRASTRUCTURE.  Stand-in for the part of the Ceres Solver API that the reference's
// CPhotoconsistencyOdometryCeres.h:433-500 touches (Problem, AutoDiffCostFunction<.., DYNAMIC, 6>,
// Solver::Options / Summary, Solve, IterationCallback), so that header compiles UNMODIFIED into
// oracle/_ref.  What is real here: AutoDiffCostFunction evaluates the reference's own functor on
// T = double and on T = Jet<double,6> exactly as Ceres' autodiff does (one Jet per parameter with a
// unit infinitesimal, residual = scalar part, Jacobian row = infinitesimal part, row-major).
// What is NOT here: Ceres' trust-region minimiser.  ceres::Solve() hands the problem to a hook the
// driver installs (oracle/shim/ref_driver_ceres.cpp): it either only evaluates (pinning residuals and
// Jacobians) or runs the oracle's restated Levenberg-Marquardt loop on the reference functor.
#ifndef PHOVO_SHIM_CERES_CERES_H_
#define PHOVO_SHIM_CERES_CERES_H_

#include <string>
#include <vector>
#include "ceres/jet.h"

namespace ceres
{
enum { DYNAMIC = -1 };
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum CallbackReturnType { SOLVER_CONTINUE, SOLVER_ABORT, SOLVER_TERMINATE_SUCCESSFULLY };

struct IterationSummary { int iteration; double cost; double cost_change; double gradient_max_norm; double step_norm; double trust_region_radius; };

class IterationCallback
{
public:
  virtual ~IterationCallback() {}
  virtual CallbackReturnType operator()( const IterationSummary & summary ) = 0;
};

class LossFunction;

class CostFunction
{
public:
  virtual ~CostFunction() {}
  // parameters[0]: the single 6-vector block; jacobians may be NULL, jacobians[0] is num_residuals x 6 row-major
  virtual bool Evaluate( double const * const * parameters, double * residuals, double ** jacobians ) const = 0;
  int num_residuals() const { return m_NumResiduals; }
protected:
  int m_NumResiduals;
};

template< class Functor, int M, int N0 >
class AutoDiffCostFunction : public CostFunction
{
public:
  AutoDiffCostFunction( Functor * functor, int numResiduals ) : m_Functor( functor ) { m_NumResiduals = numResiduals; }
  ~AutoDiffCostFunction() { delete m_Functor; }   // takes ownership, like Ceres
  virtual bool Evaluate( double const * const * parameters, double * residuals, double ** jacobians ) const
  {
    if( !jacobians || !jacobians[0] ) return ( *m_Functor )( parameters[0], residuals );
    typedef Jet< double, N0 > JetT;
    JetT x[ N0 ];
    for( int k = 0; k < N0; k++ ) x[k] = JetT( parameters[0][k], k );
    std::vector< JetT > out;
    out.resize( size_t( m_NumResiduals ) );
    if( !( *m_Functor )( x, out.data() ) ) return false;
    for( int i = 0; i < m_NumResiduals; i++ )
    {
      residuals[i] = out[ size_t( i ) ].a;
      for( int k = 0; k < N0; k++ ) jacobians[0][ size_t( i ) * N0 + k ] = out[ size_t( i ) ].v( k );
    }
    return true;
  }
private:
  Functor * m_Functor;
};

class Problem
{
public:
  Problem() : m_Cost( 0 ), m_Parameters( 0 ) {}
  ~Problem() { delete m_Cost; }                   // owns its cost functions, like Ceres
  void AddResidualBlock( CostFunction * cost, LossFunction *, double * parameters ) { delete m_Cost; m_Cost = cost; m_Parameters = parameters; }
  CostFunction * cost() const { return m_Cost; }
  double * parameters() const { return m_Parameters; }
private:
  Problem( const Problem & );
  Problem & operator=( const Problem & );
  CostFunction * m_Cost;
  double * m_Parameters;
};

class Solver
{
public:
  struct Options
  {
    Options() : max_num_iterations( 50 ), linear_solver_type( SPARSE_NORMAL_CHOLESKY ), minimizer_progress_to_stdout( false ),
      function_tolerance( 1e-6 ), gradient_tolerance( 1e-10 ), parameter_tolerance( 1e-8 ),
      initial_trust_region_radius( 1e4 ), max_trust_region_radius( 1e16 ), min_trust_region_radius( 1e-32 ),
      min_relative_decrease( 1e-3 ), num_linear_solver_threads( 1 ), num_threads( 1 ),
      max_num_consecutive_invalid_steps( 5 ), update_state_every_iteration( false ) {}
    int max_num_iterations;
    LinearSolverType linear_solver_type;
    bool minimizer_progress_to_stdout;
    double function_tolerance, gradient_tolerance, parameter_tolerance;
    double initial_trust_region_radius, max_trust_region_radius, min_trust_region_radius;
    double min_relative_decrease;
    int num_linear_solver_threads, num_threads;
    int max_num_consecutive_invalid_steps;
    bool update_state_every_iteration;
    std::vector< IterationCallback * > callbacks;
  };
  struct Summary
  {
    Summary() : num_iterations( 0 ), initial_cost( 0. ), final_cost( 0. ) {}
    int num_iterations; double initial_cost, final_cost;
    std::string BriefReport() const { return std::string(); }
  };
};

typedef void ( *SolveHook )( const Solver::Options & options, Problem * problem, Solver::Summary * summary );
inline SolveHook & solveHook() { static SolveHook hook = 0; return hook; }
inline void Solve( const Solver::Options & options, Problem * problem, Solver::Summary * summary )
{
  if( solveHook() ) solveHook()( options, problem, summary );
}

} // namespace ceres
#endif
