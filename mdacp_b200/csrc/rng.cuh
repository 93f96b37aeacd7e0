// Per-particle counter RNG and the Langevin momentum update, shared by integrate.cu (k_langevin) and the fused
// force kernel (force.cu, FORCE_KICK_LANGEVIN).
#pragma once
#include <cuda_runtime.h>

namespace mdb {

// Counter-based per-particle RNG (Philox4x32-10): the reference draws 3 normals per particle in
// index order from one sequential std::mt19937 per unit (fcalculator.cc:772-777), which a
// data-parallel kernel cannot reproduce bit-wise (SURVEY.md finding 5 / row a10).  Parity is
// exact for AimedTemperature = 0 (noise amplitude 0) and statistical otherwise.
__device__ __forceinline__ void philox_round(unsigned &c0, unsigned &c1, unsigned &c2, unsigned &c3, unsigned k0,
                                             unsigned k1) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  unsigned hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
  unsigned hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
  unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__device__ __forceinline__ void philox4x32(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0,
                                           unsigned k1, unsigned out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u01(unsigned hi, unsigned lo) {
  // (0,1]: 53 random bits, never 0 so that log() is finite
  unsigned long long b = (((unsigned long long)hi << 32) | lo) >> 11;
  return ((double)b + 1.0) * (1.0 / 9007199254740992.0);
}


// ForceCalculator::Langevin (fcalculator.cc:765-778) for one particle: p += (-gamma p + xi) dt with three normals
// xi of amplitude amp, evaluated in the reference's operation order without FMA.  id = the particle's reference
// index, (step_lo, step_hi) = the step counter: the stream of a particle does not depend on where it is stored.
__device__ __forceinline__ void langevin_update(double4 &v, unsigned id, double gamma, double amp, double dt, unsigned key0,
                                                unsigned key1, unsigned step_lo, unsigned step_hi) {
  double n0 = 0.0, n1 = 0.0, n2 = 0.0;
  if (amp != 0.0) {
    unsigned r[4], s[4];
    philox4x32(id, step_lo, step_hi, 0u, key0, key1, r);
    philox4x32(id, step_lo, step_hi, 1u, key0, key1, s);
    double ua = u01(r[0], r[1]), ub = u01(r[2], r[3]);
    double uc = u01(s[0], s[1]), ud = u01(s[2], s[3]);
    double ra = sqrt(-2.0 * log(ua)), rc = sqrt(-2.0 * log(uc));
    double sa, ca, sc, cc;
    sincospi(2.0 * ub, &sa, &ca);
    sincospi(2.0 * ud, &sc, &cc);
    n0 = ra * ca; n1 = ra * sa; n2 = rc * cc;
    (void)sc;
  }
  v.x = __dadd_rn(v.x, __dmul_rn(__dadd_rn(__dmul_rn(-gamma, v.x), __dmul_rn(amp, n0)), dt));
  v.y = __dadd_rn(v.y, __dmul_rn(__dadd_rn(__dmul_rn(-gamma, v.y), __dmul_rn(amp, n1)), dt));
  v.z = __dadd_rn(v.z, __dmul_rn(__dadd_rn(__dmul_rn(-gamma, v.z), __dmul_rn(amp, n2)), dt));
}

}  // namespace mdb
