// cdk_rng.cuh -- the counter-based random stream shared by the EnKF kernel, the path sampler and the CPU oracle.
#pragma once
#include <stdint.h>

namespace cdk {

enum { RNG_INIT = 0, RNG_OBS = 1, RNG_DYN = 2 };

// ---- Philox4x32-10 + Box-Muller: bit-identical to oracle/cd_oracle.py philox4x32 / box_muller_f32 / philox_normal_quad ----
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

// One Box-Muller pair from two 32-bit words, in fp32 arithmetic built ONLY from correctly rounded IEEE operations (add, mul,
// div, sqrt; never contracted into FMAs), so that oracle/cd_oracle.py:box_muller_f32 reproduces every deviate BIT FOR BIT
// with NumPy float32 arrays -- libm and CUDA's log / sincospi differ in the last ulp, which the chaotic test problems then
// amplify.  (History: fp64 log / sqrt / sincospi, ~130 FP64-pipe instructions + ~70 constant moves per quad, more than the
// ten Philox rounds themselves; 53 % of the EnKF's instruction stream.)
//   radius: u = (float(ra) + 0.5) 2^-32 in (0, 1];  ln u = ln m + (e - 32) ln 2 with u 2^32 = m 2^e, m in [sqrt(.5), sqrt 2),
//           ln m = 2 s (1 + s^2/3 + s^4/5 + s^6/7 + s^8/9), s = (m - 1) / (m + 1)  (|s| <= 0.172: truncation 2e-9);
//           rad = sqrt(-2 ln u) <= 6.76
//   angle:  2 pi v = q pi/2 + pi/4 + y, q = rb >> 30, y = g pi/2, g = (bits 29..7 of rb + 0.5) 2^-23 - 0.5 in (-0.5, 0.5);
//           sin y, cos y by their Taylor polynomials to y^9 / y^8 (|y| <= pi/4: truncation 2e-9 / 2.5e-8), the rotation by
//           pi/4 as (cos y -+ sin y) sqrt(.5) with sqrt(.5) folded into the radius, the quadrant by swap / negate.
// The deviates carry fp32 precision (relative 6e-8) and are exactly representable in fp32; the filters widen them to T.
__device__ __forceinline__ static void box_muller_f32(uint32_t ra, uint32_t rb, float& z0, float& z1) {
  const float uf = __fadd_rn(__uint2float_rn(ra), 0.5f);
  const int bits = __float_as_int(uf);
  int e = (bits >> 23) - 127;
  float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
  if (m > 1.41421354f) {
    m = __fmul_rn(m, 0.5f);
    e += 1;
  }
  const float s = __fdiv_rn(__fadd_rn(m, -1.0f), __fadd_rn(m, 1.0f));
  const float s2 = __fmul_rn(s, s);
  float p = 0.111111112f;
  p = __fadd_rn(__fmul_rn(p, s2), 0.142857149f);
  p = __fadd_rn(__fmul_rn(p, s2), 0.2f);
  p = __fadd_rn(__fmul_rn(p, s2), 0.333333343f);
  p = __fadd_rn(__fmul_rn(p, s2), 1.0f);
  const float lnm = __fmul_rn(__fmul_rn(2.0f, s), p);
  const float lnu = __fadd_rn(lnm, __fmul_rn((float)(e - 32), 0.693147182f));
  const float radh = __fmul_rn(__fsqrt_rn(__fmul_rn(-2.0f, lnu)), 0.707106769f);
  const uint32_t q = rb >> 30;
  const float g = __fadd_rn(__fmul_rn(__fadd_rn(__uint2float_rn((rb >> 7) & 0x7fffffu), 0.5f), 1.1920929e-07f), -0.5f);
  const float y = __fmul_rn(g, 1.57079637f);
  const float y2 = __fmul_rn(y, y);
  float ps = 2.75573188e-06f;  // 1/9!
  ps = __fadd_rn(__fmul_rn(ps, y2), -1.98412701e-04f);
  ps = __fadd_rn(__fmul_rn(ps, y2), 8.33333377e-03f);
  ps = __fadd_rn(__fmul_rn(ps, y2), -0.166666672f);
  ps = __fadd_rn(__fmul_rn(ps, y2), 1.0f);
  const float sy = __fmul_rn(y, ps);
  float pc = 2.48015876e-05f;  // 1/8!
  pc = __fadd_rn(__fmul_rn(pc, y2), -1.38888892e-03f);
  pc = __fadd_rn(__fmul_rn(pc, y2), 4.16666679e-02f);
  pc = __fadd_rn(__fmul_rn(pc, y2), -0.5f);
  const float cy = __fadd_rn(__fmul_rn(pc, y2), 1.0f);
  const float c45 = __fadd_rn(cy, -sy), s45 = __fadd_rn(cy, sy);  // sqrt(2) cos / sin (pi/4 + y)
  float cq = (q & 1u) ? -s45 : c45;
  float sq = (q & 1u) ? c45 : s45;
  if (q & 2u) {
    cq = -cq;
    sq = -sq;
  }
  z0 = __fmul_rn(radh, cq);
  z1 = __fmul_rn(radh, sq);
}

// FOUR standard normals per Philox call, for (member, trajectory, step, stream | substep | quad): two Box-Muller pairs,
// (z0, z1) from words (r0, r1), (z2, z3) from (r2, r3).  (Round 1 drew two normals per call from 53-bit uniforms.)
__device__ __noinline__ static void normal_quad(uint32_t member, uint32_t traj, uint32_t step, uint32_t c3, uint64_t seed,
                                                double (&z)[4]) {
  uint32_t r[4];
  philox4x32_10(member, traj, step, c3, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), r);
  float f[4];
  box_muller_f32(r[0], r[1], f[0], f[1]);
  box_muller_f32(r[2], r[3], f[2], f[3]);
#pragma unroll
  for (int u = 0; u < 4; ++u) z[u] = (double)f[u];
}

// Two quads (counters c3a, c3b) in one call: the same eight values normal_quad would return, with the two 10-round Philox
// chains and the four Box-Muller pairs interleaved -- one chain alone leaves a scheduler with two resident warps waiting
// on its IMAD.WIDE -> LOP3 dependency most of the time.
__device__ __noinline__ static void normal_oct(uint32_t member, uint32_t traj, uint32_t step, uint32_t c3a, uint32_t c3b,
                                               uint64_t seed, double (&z)[8]) {
  uint32_t ra[4], rb[4];
  const uint32_t k0 = (uint32_t)(seed & 0xffffffffu), k1 = (uint32_t)(seed >> 32);
  philox4x32_10(member, traj, step, c3a, k0, k1, ra);
  philox4x32_10(member, traj, step, c3b, k0, k1, rb);
  float f[8];
  box_muller_f32(ra[0], ra[1], f[0], f[1]);
  box_muller_f32(ra[2], ra[3], f[2], f[3]);
  box_muller_f32(rb[0], rb[1], f[4], f[5]);
  box_muller_f32(rb[2], rb[3], f[6], f[7]);
#pragma unroll
  for (int u = 0; u < 8; ++u) z[u] = (double)f[u];
}

__device__ __forceinline__ static uint32_t rng_c3(int stream, int substep, int quad) {
  return ((uint32_t)stream << 28) | (((uint32_t)substep & 0xFFFFFu) << 8) | (uint32_t)quad;
}


}  // namespace cdk
