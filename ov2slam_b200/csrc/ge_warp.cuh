// Warp-synchronous Gauss-Jordan solve of a small SPD system held in shared memory (reduced camera system of
// localBA windows with <= 96 unknowns; Optimizer::localBA's DENSE_SCHUR back end,
// /root/reference/src/optimizer.cpp:439-443).  ONE warp, no block barrier: lane l owns rows l, l+32, l+64 of
// the augmented matrix [S | b]; pivot step j eliminates column j from every other row, so no substitution
// passes follow.  The only synchronisation is one __syncwarp per pivot (row j+1 must be complete before it is
// read as the next pivot row).  Candidate replacement of the 12-warp register version (mode 2/3 of
// ba_reduced_solve_kernel), whose 48 block barriers + publish/read round trips cost ~0.9 us per pivot (ncu).
// Written per lane so that tests/test_host_logic.py can run the same code on the host, lane by lane.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define GE_HD __host__ __device__ inline
#else
#define GE_HD inline
#endif

namespace gewarp {

GE_HD int pitch(int n) { return (n + 1) | 1; }   // odd: rows of consecutive lanes fall into different banks

// reciprocal of a positive pivot without the IEEE division routine on the critical path:
// float reciprocal + two Newton steps in double (relative error ~2^-92 before the final rounding)
GE_HD double recip(double p) {
#if defined(__CUDA_ARCH__)
    double r = (double)__frcp_rn((float)p);
#else
    double r = (double)(1.0f / (float)p);
#endif
    r = r * (2.0 - p * r);
    r = r * (2.0 - p * r);
    return r;
}

// pivot step j for this lane's rows.  Returns false when the pivot is not a usable positive number
// (the same test Ceres' LLT failure path amounts to; uniform across lanes: every lane reads the same pivot).
GE_HD bool step(int lane, double* A, int n, int P, int j) {
    const double p = A[j * P + j];
    if (!(p > 0.0) || !isfinite(p)) return false;
    const double ip = recip(p);
    const double* rj = A + j * P;
    for (int i = lane; i < n; i += 32) {
        if (i == j) continue;
        double* ri = A + i * P;
        const double m = ri[j] * ip;
        for (int c = j + 1; c <= n; ++c) ri[c] -= m * rj[c];
    }
    return true;
}

// after the n steps: x_i = A[i][n] / A[i][i]
GE_HD void finish(int lane, const double* A, int n, int P, double* x) {
    for (int i = lane; i < n; i += 32) x[i] = A[i * P + n] * recip(A[i * P + i]);
}

}  // namespace gewarp
