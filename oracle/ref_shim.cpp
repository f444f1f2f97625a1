// ref_shim.cpp -- C-ABI door onto the reference's own, unmodified C++ headers.
//
// TEST INFRASTRUCTURE ONLY (see oracle/bsls_oracle.c for the rules).  The headers
// are compiled from where they lie under /root/reference/python/c_extensions
// (oracle/Makefile passes -I); no reference source is copied into this repo.
// The output, oracle/_ref/libbsls_ref.so, is git-ignored but travels to the GPU
// box, where it serves as the strongest checker and as the "reference" CPU baseline.
//
// Each wrapper forwards to the routine the reference's Cython layer binds
// (python/c_extensions/c_extensions.pyx:15-19,52-61).
#include "proj_simplex.h"
#include "isotonic_regression.h"

extern "C" {

void ref_proj_simplex(double *y, int start, int end) { proj_simplex(y, start, end); }
void ref_proj_multi_simplex(double *y, int *blocks, int numblocks, int n) { proj_multi_simplex(y, blocks, numblocks, n); }
void ref_proj_multi_ball(double *y, int *blocks, int numblocks, int n) { proj_multi_ball(y, blocks, numblocks, n); }

void ref_isotonic_regression(double *y, int start, int end, int *weight, int update) { isotonic_regression(y, start, end, weight, update); }
void ref_isotonic_regression_multi(double *y, int *blocks, int numblocks, int n, int *weight, int update) { isotonic_regression_multi(y, blocks, numblocks, n, weight, update); }
void ref_isotonic_regression_2(double *y, int start, int end) { isotonic_regression_2(y, start, end); }
void ref_isotonic_regression_multi_2(double *y, int *blocks, int numblocks, int n) { isotonic_regression_multi_2(y, blocks, numblocks, n); }
void ref_isotonic_regression_3(double *y, int start, int end, int *weight, int update) { isotonic_regression_3(y, start, end, weight, update); }
void ref_isotonic_regression_multi_3(double *y, int *blocks, int numblocks, int n, int *weight, int update) { isotonic_regression_multi_3(y, blocks, numblocks, n, weight, update); }

}  // extern "C"
