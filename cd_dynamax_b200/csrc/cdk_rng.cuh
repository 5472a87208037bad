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

// two standard normals for (member, trajectory, step, stream | substep | pair)
__device__ __noinline__ static void normal_pair(uint32_t member, uint32_t traj, uint32_t step, uint32_t c3, uint64_t seed,
                                            double& z0, double& z1) {
  uint32_t r[4];
  philox4x32_10(member, traj, step, c3, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), r);
  const double u1 = ((double)(r[0] >> 5) * 67108864.0 + (double)(r[1] >> 6) + 0.5) * (1.0 / 9007199254740992.0);
  const double u2 = ((double)(r[2] >> 5) * 67108864.0 + (double)(r[3] >> 6) + 0.5) * (1.0 / 9007199254740992.0);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincos(6.283185307179586 * u2, &s, &c);
  z0 = rad * c;
  z1 = rad * s;
}

__device__ __forceinline__ static uint32_t rng_c3(int stream, int substep, int pair) {
  return ((uint32_t)stream << 28) | (((uint32_t)substep & 0xFFFFFu) << 8) | (uint32_t)pair;
}


}  // namespace cdk
