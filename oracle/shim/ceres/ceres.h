// ORACLE — TEST INFRASTRUCTURE ONLY.  Just enough of the Ceres API for the `Create` factories of the reference's
// src/CeresResidues.h to compile (they are never called through this shim; the functors' operator() is called directly).
#pragma once
namespace ceres {
class CostFunction { public: virtual ~CostFunction() {} };
template <class F, int kNumResiduals, int... Ns> class AutoDiffCostFunction : public CostFunction { public: explicit AutoDiffCostFunction(F* f) : f_(f) {} ~AutoDiffCostFunction() { delete f_; } private: F* f_; };
class LocalParameterization { public: virtual ~LocalParameterization() {} };
template <class F, int kGlobal, int kLocal> class AutoDiffLocalParameterization : public LocalParameterization {};
}  // namespace ceres
