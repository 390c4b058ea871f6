// Counter-based RNG of the sampler: Philox4x32-10 (Salmon et al., SC'11) + Box-Muller.
// Bit-for-bit the same stream as oracle/magi_oracle.py::rng_normals / rng_uniform so that a whole
// HMC chain (momenta AND accept decisions) can be checked against the CPU oracle.
//   counter = (pair index, global chain id, global iteration, purpose), key = 64-bit seed.
#pragma once
#include "common.cuh"

#define MAGI_RNG_MOMENTUM 0u
#define MAGI_RNG_ACCEPT 1u

__device__ __forceinline__ uint4 magi_philox(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// (0,1) double from two 32-bit words: 53 random bits + 1/2 ulp.
__device__ __forceinline__ double magi_u53(uint32_t hi, uint32_t lo) {
  const uint64_t k = ((uint64_t)hi << 21) ^ ((uint64_t)lo >> 11);
  return ((double)k + 0.5) * 1.1102230246251565e-16;  // 2^-53
}

// two standard normals for (pair, chain, iteration)
__device__ __forceinline__ void magi_normal_pair(uint64_t seed, uint32_t pair, uint32_t chain, uint32_t iter,
                                                 double& z0, double& z1) {
  const uint4 r = magi_philox(make_uint4(pair, chain, iter, MAGI_RNG_MOMENTUM),
                              make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const double u1 = magi_u53(r.x, r.y), u2 = magi_u53(r.z, r.w);
  const double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  z0 = rad * cs;
  z1 = rad * sn;
}

__device__ __forceinline__ double magi_uniform(uint64_t seed, uint32_t chain, uint32_t iter) {
  const uint4 r = magi_philox(make_uint4(0u, chain, iter, MAGI_RNG_ACCEPT),
                              make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  return magi_u53(r.x, r.y);
}
