// Internal (C++) entry points shared between the stage files and api.cu.
#pragma once
#include "common.cuh"

namespace paig {

// rollout.cu
int rollout_forward(int cell, int n, int B, int steps, const float* dt, const double* p0, const double* p1, float* seq,
                    cudaStream_t st);
int rollout_backward(int cell, int n, int B, int steps, const float* dt, const double* p0, const double* p1,
                     const float* seq, const float* dpos, const float* dvel, long batch_stride, long row_stride,
                     int with_row0, float* d_state0, double* d_phys, cudaStream_t st);

}  // namespace paig
