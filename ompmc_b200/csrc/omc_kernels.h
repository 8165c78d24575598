// omc_kernels.h -- host-callable launchers of the CUDA kernels (internal to libompmc_b200.so).
#pragma once
#include "omc_types.cuh"

namespace omc {

// omc_lockstep.cu
void launch_lockstep(const DevProblem &P, Part *stack, int depth, int blocks, int threads, long long first, long long nhist,
                     cudaStream_t stream);
int lockstep_blocks_per_sm(int threads);
void launch_test_geometry(const DevProblem &P, int n, const double *xyzuvw, const int *ir, const double *ustep_in, int *idisc,
                          int *irnew, double *ustep_out, double *tperp, cudaStream_t stream);
void launch_test_rng(uint32_t s0, uint32_t s1, unsigned long long hist, int n, double *out, cudaStream_t stream);
void launch_accum(double *endep, double *accum, double *accum2, long long n, cudaStream_t stream);

}  // namespace omc
