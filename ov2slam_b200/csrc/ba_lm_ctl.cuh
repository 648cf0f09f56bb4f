// L (controller): Ceres 2.0's TrustRegionMinimizer + LevenbergMarquardtStrategy decision logic as a small
// state machine that runs ON THE DEVICE (every CTA of the persistent solve kernel replays it from the same
// reduced scalars, so all CTAs take the same branch without a broadcast) and compiles for the host as well
// (tests/test_host_logic.py drives it with the oracle's per-iteration scalars).
//
// Reference behaviour (file:line in /root/reference/Thirdparty/ceres-solver/internal/ceres):
//   trust_region_minimizer.cc:67-134   Minimize(): iteration loop, FinalizeIterationAndCheckIfMinimizerCanContinue
//   trust_region_minimizer.cc:377-448  ComputeTrustRegionStep: invalid steps (model_cost_change <= 0, failed
//                                      factorisation), max_consecutive_invalid_steps = 5
//   trust_region_minimizer.cc:706-748  ParameterToleranceReached (1e-8), FunctionToleranceReached
//   trust_region_minimizer.cc:786-826  HandleSuccessfulStep / HandleUnsuccessfulStep
//   levenberg_marquardt_strategy.cc:147-160  StepAccepted: radius /= max(1/3, 1 - (2 rho - 1)^3), capped 1e16;
//                                      StepRejected: radius /= decrease_factor, decrease_factor *= 2
// Options fixed by Optimizer::localBA (src/optimizer.cpp:436-470): initial radius 1e4, min radius 1e-32,
// min_relative_decrease 1e-3, gradient_tolerance 1e-10, parameter_tolerance 1e-8, monotonic steps.
#pragma once
#include <float.h>
#include <math.h>

#if defined(__CUDACC__)
#define LMC_HD __host__ __device__ __forceinline__
#else
#define LMC_HD inline
#endif

namespace lmctl {

enum Action { ACT_CONTINUE_REJECTED = 0, ACT_CONTINUE_ACCEPTED = 1, ACT_STOP = 2 };

struct State {
    double x_cost, minimum_cost, xnorm, gmax, radius, decrease_factor;
    double initial_cost;
    int step_successful, x_is_new, cost_known, iteration, num_invalid, first_iter, termination;   // termination: 0 convergence, 1 max iterations, 2 failure
};

LMC_HD void init(State& s) {
    s.x_cost = 0.0; s.minimum_cost = DBL_MAX; s.xnorm = -1.0; s.gmax = DBL_MAX; s.radius = 1e4; s.decrease_factor = 2.0;
    s.initial_cost = 0.0;
    s.step_successful = 1; s.x_is_new = 1; s.cost_known = 0; s.iteration = 0; s.num_invalid = 0; s.first_iter = 1; s.termination = 0;
}

// Top of the loop: FinalizeIterationAndCheckIfMinimizerCanContinue.  Returns false when the minimizer stops
// (termination set); otherwise the iteration counter has been advanced and a trust-region step must be computed.
LMC_HD bool begin_iteration(State& s, int max_iters) {
    if (s.step_successful && s.cost_known && s.x_cost < s.minimum_cost) s.minimum_cost = s.x_cost;
    if (s.iteration >= max_iters) { s.termination = 1; return false; }
    if (s.radius <= 1e-32) { s.termination = 0; return false; }
    s.iteration++;
    return true;
}

// After the step: cost at x (valid when the Jacobians were re-evaluated this iteration), candidate cost, model cost
// change, squared step norm, squared candidate norm, gradient max norm at x, factorisation failure flag.
LMC_HD Action end_iteration(State& s, double cost_at_x, double cand_cost, double model_cost_change, double step2, double candx2,
                            double gmax_at_x, bool chol_fail, double function_tolerance) {
    if (s.x_is_new) {
        s.x_cost = cost_at_x;            // cost of the (re-)evaluation at x, as Ceres uses it
        s.cost_known = 1;
        if (s.iteration == 1) s.initial_cost = s.x_cost;
        if (s.x_cost < s.minimum_cost) s.minimum_cost = s.x_cost;
        // GradientToleranceReached() for the point this system was assembled at: Ceres would have stopped before
        // this iteration, so the iteration does not count
        s.gmax = gmax_at_x;
        if (s.gmax <= 1e-10) { s.termination = 0; s.iteration--; return ACT_STOP; }
    }
    s.x_is_new = 0;
    s.first_iter = 0;
    const bool step_valid = !chol_fail && isfinite(model_cost_change) && model_cost_change > 0.0;
    if (!step_valid) {
        if (++s.num_invalid >= 5) { s.termination = 2; return ACT_STOP; }
        s.radius /= s.decrease_factor;
        s.decrease_factor *= 2.0;
        s.step_successful = 0;
        return ACT_CONTINUE_REJECTED;
    }
    s.num_invalid = 0;
    if (!isfinite(cand_cost)) cand_cost = DBL_MAX;
    const double step_norm = sqrt(step2);
    if (step_norm <= 1e-8 * (s.xnorm + 1e-8)) { s.termination = 0; return ACT_STOP; }          // candidate NOT adopted
    const double cost_change = s.x_cost - cand_cost;
    if (fabs(cost_change) <= function_tolerance * s.x_cost) { s.termination = 0; return ACT_STOP; }
    const double rel = cand_cost >= DBL_MAX ? -DBL_MAX : cost_change / model_cost_change;
    if (rel > 1e-3) {
        s.xnorm = sqrt(candx2);
        s.x_cost = cand_cost;            // replaced by the re-evaluated value at the next end_iteration
        s.x_is_new = 1;
        s.step_successful = 1;
        const double t = 2.0 * rel - 1.0;
        s.radius = s.radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
        s.radius = fmin(1e16, s.radius);
        s.decrease_factor = 2.0;
        return ACT_CONTINUE_ACCEPTED;
    }
    s.step_successful = 0;
    s.radius /= s.decrease_factor;
    s.decrease_factor *= 2.0;
    return ACT_CONTINUE_REJECTED;
}

LMC_HD double final_cost(const State& s) { return s.minimum_cost < DBL_MAX ? s.minimum_cost : s.x_cost; }

}  // namespace lmctl
