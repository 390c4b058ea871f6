// Shared helpers for the sm_100a MAGI kernels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/magi_b200.h"

#define MAGI_CHAINS_PER_CTA 8   // chains of one dataset processed together (DMMA n = 8)
#define MAGI_FULL_MASK 0xffffffffu

static inline int magi_cuda_status(cudaError_t e) { return e == cudaSuccess ? MAGI_OK : MAGI_ERR_CUDA + (int)e; }

// rows/cols of the packed matrices are padded to a multiple of 8 doubles (64 B)
__host__ __device__ static inline int magi_pad8(int n) { return (n + 7) & ~7; }

// Position of element (r, col) inside an 8 x 8 tile of the packed matrices (512 B).  The tile is stored as 32 pairs of
// doubles (16 B each: (r, 2 cp), (r, 2 cp + 1)), pair (r, cp) in slot
//     sigma(r, cp) = 4 (r ^ ((r >> 1) & 1)) + ((cp + 2 ((r >> 2) & 1)) & 3)
// -- a bijection of the 32 slots chosen so that BOTH fragment shapes of the FP64 MMA are conflict-free reads out of
// shared memory once a tile has been bulk-copied there verbatim: the forward fragment of lane 4g+c (one 16-byte load of
// pair (g, c): the eight lanes of a quarter warp hit eight different slots mod 8) and the transposed fragment (two
// 8-byte loads, elements (2c, g) and (2c+1, g): the sixteen lanes of a half warp hit sixteen different bank pairs;
// with the plain row-major tile they were a 4-way conflict, which is why round 2 first transposed in registers with
// six shuffles per tile).
__host__ __device__ static inline int magi_tile_slot(int r, int cp) {
  return 4 * (r ^ ((r >> 1) & 1)) + ((cp + 2 * ((r >> 2) & 1)) & 3);
}
__host__ __device__ static inline int magi_tile_pos(int r, int col) { return 2 * magi_tile_slot(r, col >> 1) + (col & 1); }

// packed layout: [B][D][3][np/8][np/8][8][8] (8x8 tiles, row-major in the tile), slot 0 = sym(C^-1),
// 1 = m, 2 = sym(K^-1); padding is zero.
__host__ __device__ static inline size_t magi_packed_mat_elems(int n) {
  return (size_t)magi_pad8(n) * (size_t)magi_pad8(n);
}

// softplus / sigmoid in the overflow-safe form; equals log(1+exp(z)) (magi_v2.py:318-319) to
// rounding wherever the naive form is finite.
__device__ __forceinline__ double magi_softplus(double z) { return fmax(z, 0.0) + log1p(exp(-fabs(z))); }
__device__ __forceinline__ double magi_sigmoid(double z) {
  double e = exp(-fabs(z));
  return z >= 0.0 ? 1.0 / (1.0 + e) : e / (1.0 + e);
}

__device__ __forceinline__ double magi_shfl_xor(double v, int m) { return __shfl_xor_sync(MAGI_FULL_MASK, v, m); }
__device__ __forceinline__ double magi_shfl_down(double v, int d) { return __shfl_down_sync(MAGI_FULL_MASK, v, d); }

__device__ __forceinline__ double magi_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += magi_shfl_xor(v, o);
  return v;
}
