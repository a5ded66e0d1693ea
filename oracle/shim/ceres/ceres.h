// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for the part of the Ceres API the reference uses: the `Create` factories
// of src/CeresResidues.h compile against AutoDiffCostFunction, and src/PoseGraphSLAM.cpp builds its problem on a
// ceres::Problem that RECORDS what it is given (parameter blocks, parameterisations, residual blocks with their
// parameter pointers, constant blocks, removals).  ceres::Solve does not minimise anything: it calls a hook the test
// driver installs (default: nothing), so what the tests compare is the problem the reference's own front-end code
// constructs and the initial guesses it writes — not a solve.
#pragma once
#include <functional>
#include <map>
#include <string>
#include <typeinfo>
#include <vector>
namespace ceres {
class CostFunction { public: virtual ~CostFunction() {} virtual const char* functor_name() const { return "?"; } virtual int num_residuals() const { return 0; } virtual const void* functor() const { return nullptr; } };
template <class F, int kNumResiduals, int... Ns> class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(F* f) : f_(f) {}
  ~AutoDiffCostFunction() { delete f_; }
  const char* functor_name() const override { return typeid(F).name(); }
  int num_residuals() const override { return kNumResiduals; }
  const void* functor() const override { return f_; }
 private:
  F* f_;
};
class LossFunction { public: virtual ~LossFunction() {} };
class CauchyLoss : public LossFunction { public: explicit CauchyLoss(double) {} };
class HuberLoss : public LossFunction { public: explicit HuberLoss(double) {} };
class LocalParameterization { public: virtual ~LocalParameterization() {} };
class EigenQuaternionParameterization : public LocalParameterization {};
class QuaternionParameterization : public LocalParameterization {};
template <class F, int kGlobal, int kLocal> class AutoDiffLocalParameterization : public LocalParameterization {};
struct ResidualBlock { CostFunction* cost; LossFunction* loss; std::vector<double*> params; bool removed; };
typedef ResidualBlock* ResidualBlockId;
class Problem {
 public:
  ~Problem() { for (ResidualBlock* b : blocks) delete b; }
  void AddParameterBlock(double* p, int size) { param_size[p] = size; }
  void AddParameterBlock(double* p, int size, LocalParameterization* lp) { param_size[p] = size; param_lp[p] = lp; }
  void SetParameterization(double* p, LocalParameterization* lp) { param_lp[p] = lp; }
  void SetParameterBlockConstant(double* p) { param_const[p] = true; }
  void SetParameterBlockVariable(double* p) { param_const[p] = false; }
  void SetParameterLowerBound(double*, int, double) {}
  void SetParameterUpperBound(double*, int, double) {}
  template <class... P> ResidualBlockId AddResidualBlock(CostFunction* c, LossFunction* l, P... ps) { ResidualBlock* b = new ResidualBlock{c, l, {ps...}, false}; blocks.push_back(b); return b; }
  void RemoveResidualBlock(ResidualBlockId b) { b->removed = true; }
  int NumResidualBlocks() const { int n = 0; for (ResidualBlock* b : blocks) n += !b->removed; return n; }
  int NumParameterBlocks() const { return (int)param_size.size(); }
  std::vector<ResidualBlock*> blocks;
  std::map<double*, int> param_size; std::map<double*, LocalParameterization*> param_lp; std::map<double*, bool> param_const;
};
enum LinearSolverType { DENSE_QR, DENSE_SCHUR, SPARSE_SCHUR, SPARSE_NORMAL_CHOLESKY, ITERATIVE_SCHUR, CGNR, DENSE_NORMAL_CHOLESKY };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };
enum TrustRegionStrategyType { LEVENBERG_MARQUARDT, DOGLEG };
enum MinimizerType { LINE_SEARCH, TRUST_REGION };
class Solver {
 public:
  struct Options {
    bool minimizer_progress_to_stdout = false; LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY; int max_num_iterations = 50; int num_threads = 1;
    int num_linear_solver_threads = 1; double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8, max_solver_time_in_seconds = 1e9;
    TrustRegionStrategyType trust_region_strategy_type = LEVENBERG_MARQUARDT; MinimizerType minimizer_type = TRUST_REGION; bool update_state_every_iteration = false;
  };
  struct Summary {
    TerminationType termination_type = NO_CONVERGENCE; double initial_cost = 0, final_cost = 0, total_time_in_seconds = 0; int num_successful_steps = 0, num_unsuccessful_steps = 0;
    std::vector<int> iterations;
    std::string BriefReport() const { return "shim: no solve"; }
    std::string FullReport() const { return "shim: no solve"; }
    bool IsSolutionUsable() const { return true; }
  };
};
inline std::function<void(const Solver::Options&, Problem*, Solver::Summary*)>& solve_hook() { static std::function<void(const Solver::Options&, Problem*, Solver::Summary*)> h; return h; }
inline void Solve(const Solver::Options& o, Problem* p, Solver::Summary* s) { if (solve_hook()) solve_hook()(o, p, s); }
}  // namespace ceres
