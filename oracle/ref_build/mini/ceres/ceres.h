// TEST INFRASTRUCTURE (oracle/): the three Ceres interface classes the reference's residual source derives from (ceres/cost_function.h,
// sized_cost_function.h, local_parameterization.h of Ceres 2.0), declarations only - enough to compile
// /root/reference/src/ceres_parametrization.cpp here and call its Evaluate() / Plus() directly.  No solver.
#pragma once

namespace ceres {

class CostFunction {
public:
    virtual ~CostFunction() {}
    virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
};

template <int kNumResiduals, int... Ns> class SizedCostFunction : public CostFunction {};

class LocalParameterization {
public:
    virtual ~LocalParameterization() {}
    virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
    virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;
    virtual int GlobalSize() const = 0;
    virtual int LocalSize() const = 0;
};

}  // namespace ceres
