// TEST INFRASTRUCTURE (oracle/): stands in for the config.h that Ceres' CMake generates from cmake/config.h.in - the options of a
// dependency-free build (no SuiteSparse / CXSparse / LAPACK / threads), for compiling a few Ceres 2.0 sources in place.
#pragma once
#define CERES_NO_SUITESPARSE
#define CERES_NO_CXSPARSE
#define CERES_NO_ACCELERATE_SPARSE
#define CERES_NO_LAPACK
#define CERES_NO_THREADS
#define CERES_RESTRICT_SCHUR_SPECIALIZATION
