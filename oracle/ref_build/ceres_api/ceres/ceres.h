// TEST INFRASTRUCTURE (oracle/): what the reference's residual headers and the check driver need of <ceres/ceres.h> - the REAL Ceres 2.0
// public headers of the vendored tree (Thirdparty/ceres-solver/include), minus the automatic-differentiation / Jet / numeric-diff
// headers, which need parts of Eigen the stand-in linear algebra header does not provide and which this path never uses.
#pragma once
#include "ceres/cost_function.h"
#include "ceres/iteration_callback.h"
#include "ceres/local_parameterization.h"
#include "ceres/loss_function.h"
#include "ceres/ordered_groups.h"
#include "ceres/problem.h"
#include "ceres/sized_cost_function.h"
#include "ceres/solver.h"
#include "ceres/types.h"
