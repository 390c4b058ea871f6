// Compile-only harness for the fast-path kernels (ptxas -v / SASS checks in seconds instead of the minutes
// sampler.cu takes):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v -c tools/dev_fast_kernel.cu
#include "../magi_v2_b200/csrc/sampler_fast.cuh"
template __global__ void logpost_grad_fast_kernel<Seir4, 168>(magi_problem_t, const double*, const double*, const double*,
                                                              const double*, double*, double*, double*, double*);
#ifdef DEV_HMC
template __global__ void hmc_fast_kernel<Seir4, 168>(magi_problem_t, magi_hmc_config_t, double*, double*, double*, double*,
                                                     double*, HmcOut, double*);
#endif
