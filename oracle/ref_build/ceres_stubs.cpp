// TEST INFRASTRUCTURE (oracle/): entry points of three Ceres sources that are NOT compiled here (dogleg_strategy.cc, polynomial.cc,
// line_search_direction.cc need decompositions the stand-in linear-algebra header does not provide) but are referenced by factories
// in files that are.  None of them lies on the path the checks run (Levenberg-Marquardt trust region, DENSE_SCHUR, no line search):
// reaching one aborts loudly.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ceres/dogleg_strategy.h"
#include "ceres/line_search_direction.h"
#include "ceres/polynomial.h"

namespace ceres {
namespace internal {

static void off_path(const char* what) {
    fprintf(stderr, "oracle/_ref: %s is not part of this build (see oracle/ref_build/ceres_stubs.cpp)\n", what);
    abort();
}

DoglegStrategy::DoglegStrategy(const TrustRegionStrategy::Options&)
    : linear_solver_(nullptr), radius_(0), max_radius_(0), min_diagonal_(0), max_diagonal_(0), min_mu_(0), max_mu_(0), mu_increase_factor_(0),
      increase_threshold_(0), decrease_threshold_(0) {
    off_path("DoglegStrategy");
}
TrustRegionStrategy::Summary DoglegStrategy::ComputeStep(const TrustRegionStrategy::PerSolveOptions&, SparseMatrix*, const double*, double*) {
    off_path("DoglegStrategy::ComputeStep");
    return TrustRegionStrategy::Summary();
}
void DoglegStrategy::StepAccepted(double) { off_path("DoglegStrategy::StepAccepted"); }
void DoglegStrategy::StepRejected(double) { off_path("DoglegStrategy::StepRejected"); }
void DoglegStrategy::StepIsInvalid() { off_path("DoglegStrategy::StepIsInvalid"); }
double DoglegStrategy::Radius() const { off_path("DoglegStrategy::Radius"); return 0.0; }

LineSearchDirection* LineSearchDirection::Create(const LineSearchDirection::Options&) { off_path("LineSearchDirection::Create"); return nullptr; }

void MinimizeInterpolatingPolynomial(const std::vector<FunctionSample>&, double, double, double*, double*) { off_path("MinimizeInterpolatingPolynomial"); }

}  // namespace internal
}  // namespace ceres
