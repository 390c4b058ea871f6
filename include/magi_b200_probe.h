/* Measurement probes of libmagi_b200.so: synthetic kernels that give bench.py the FP64 roofs of the GPU it runs on
 * (MEASURED_PEAKS.json carries HBM and bf16 numbers only).  Not part of the reference-facing path: nothing in the
 * reference (magi_v2.py) corresponds to them. */
#ifndef MAGI_B200_PROBE_H
#define MAGI_B200_PROBE_H
#include "magi_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* kind 0: dependent DFMA chains on the FP64 vector pipe (8 per thread); kind 1: mma.sync.m8n8k4.f64 (DMMA, 4
 * accumulator pairs per warp).  Launches blocks x threads, `iters` inner iterations; `out` needs blocks * threads
 * doubles.  *flops receives the floating-point operations the launch executes (host value).  Time it with CUDA
 * events on `stream`. */
MAGI_API int magi_b200_probe_fp64(int kind, int iters, int blocks, int threads, double* out, double* flops,
                                  magi_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
