// Counter-based Philox4x32-10 + the repo's normal transform (see oracle/philox_ref.py for the spec).
// Replaces torch.randn_like / torch.randn at ps_vae/model.py:57 and ps_vae/inference.py:23,73,95.
#pragma once
#include "common.cuh"

namespace psvae {

struct PhiloxKey { uint32_t k0, k1; };

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// (ra, rb) -> two normals: s*sin(pi t), s*cos(pi t);  u = fma(ra,2^-32,2^-33), t = fma(rb,2^-31,-1)
__device__ __forceinline__ void box_muller(uint32_t ra, uint32_t rb, float& n_even, float& n_odd) {
  const float u = fmaf(__uint2float_rn(ra), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float t = fmaf(__uint2float_rn(rb), 4.6566128730773926e-10f, -1.0f);
#ifndef PSVAE_ACCURATE_BOX_MULLER
  // radius with the hardware log2 / sqrt approximations: 3 instructions instead of ~36 -- logf + IEEE sqrtf were 31 % of the Langevin kernel's
  // instructions (profiles/r02_ncu_sampling_kernels.txt).  ln u carries an absolute error of ~1e-7, the radius s one of ~1e-7 / s: measured over
  // 2^20 samples against the fp64 spec (oracle/philox_ref.py) mean 1.6e-7 (accurate form: 1.2e-7), 99.9th percentile 9.5e-7 (8.3e-7), maximum
  // 7e-5 at the rare small-radius sample (1.3e-6).  eps is noise: its value is pinned by the counter, not by the last bits of the transform.
  // Conditional sampling +35 %, bare generator +57 %, unconditional sampling +6 %, train step -5 us.  -DPSVAE_ACCURATE_BOX_MULLER restores logf / sqrtf.
  float s;
  {
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u));
    const float m2ln = -1.3862943611198906f * lg;          // -2 ln 2 * log2(u)
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(m2ln));
  }
#else
  const float s = sqrtf(-2.0f * logf(u));          // accurate log: keeps |error| ~1e-7 even for u -> 1
#endif
  float sn, cs;
  __sincosf(3.14159265358979323846f * t, &sn, &cs);   // argument in [-pi, pi]: fast intrinsic is at its best
  n_even = s * sn;
  n_odd = s * cs;
}

// The four normals of block q (elements 4q .. 4q+3 of the flat [rows, cols] tensor).
__device__ __forceinline__ float4 philox_normal4(uint64_t q, uint64_t seed, uint64_t offset) {
  const uint4 r = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)offset, (uint32_t)(offset >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
  float4 o;
  box_muller(r.x, r.y, o.x, o.y);
  box_muller(r.z, r.w, o.z, o.w);
  return o;
}

}  // namespace psvae
