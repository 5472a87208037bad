// cdk_rng.cuh -- the counter-based random stream shared by the EnKF kernel, the path sampler and the CPU oracle.
#pragma once
#include <stdint.h>

namespace cdk {

enum { RNG_INIT = 0, RNG_OBS = 1, RNG_DYN = 2 };

// ---- Philox4x32-10 + Box-Muller: bit-identical to oracle/cd_oracle.py philox4x32 / philox_normal_pair ----------------
__device__ __forceinline__ static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&r)[4]) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  r[0] = c0;
  r[1] = c1;
  r[2] = c2;
  r[3] = c3;
}

// FOUR standard normals per Philox call, for (member, trajectory, step, stream | substep | quad): the four 32-bit words
// become four uniforms (r + 0.5) 2^-32 in (0, 1) -- exact in fp64 -- and two Box-Muller pairs (z0, z1) from (u0, u1),
// (z2, z3) from (u2, u3); |z| <= sqrt(2 * 33 ln 2) = 6.76.  (Round 1 drew two normals per call from 53-bit uniforms: the
// ten Philox rounds, ~190 integer instructions, were 58 % of the EnKF's instruction stream.)
__device__ __noinline__ static void normal_quad(uint32_t member, uint32_t traj, uint32_t step, uint32_t c3, uint64_t seed,
                                                double (&z)[4]) {
  uint32_t r[4];
  philox4x32_10(member, traj, step, c3, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), r);
  const double sc = 1.0 / 4294967296.0;
  const double u0 = ((double)r[0] + 0.5) * sc, u1 = ((double)r[1] + 0.5) * sc;
  const double u2 = ((double)r[2] + 0.5) * sc, u3 = ((double)r[3] + 0.5) * sc;
  const double rad0 = sqrt(-2.0 * log(u0)), rad1 = sqrt(-2.0 * log(u2));
  double s0, c0, s1, c1;
  sincospi(2.0 * u1, &s0, &c0);
  sincospi(2.0 * u3, &s1, &c1);
  z[0] = rad0 * c0;
  z[1] = rad0 * s0;
  z[2] = rad1 * c1;
  z[3] = rad1 * s1;
}

__device__ __forceinline__ static uint32_t rng_c3(int stream, int substep, int quad) {
  return ((uint32_t)stream << 28) | (((uint32_t)substep & 0xFFFFFu) << 8) | (uint32_t)quad;
}


}  // namespace cdk
