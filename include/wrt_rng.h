/* wrt_rng.h — counter-based RNG for soft-shadow area-light samples.
 *
 * The reference draws from a random_device-seeded mt19937 (global.hpp:139-150),
 * so its soft shadows are not reproducible.  This build replaces the generator
 * by Philox4x32-10 keyed on (seed) with counter (pixel, path id, light, sample):
 * the image is then independent of GPU count, tile order and scheduling, and
 * the CPU oracle can reproduce the GPU's samples bit for bit.
 *
 * The sample distribution is the reference's *effective* one: because the
 * uniform_real_distribution in getRandomFloat is a function-local static, the
 * second call's (0, 1-u) range is ignored and u, v are both U[0,1)
 * (Triangle.hpp:139-145).
 */
#ifndef WRT_RNG_H
#define WRT_RNG_H

#include <stdint.h>

#ifdef __CUDACC__
#define WRT_HD __host__ __device__ __forceinline__
#else
#define WRT_HD static inline
#endif

#ifndef WRT_DEFAULT_SEED
#define WRT_DEFAULT_SEED 0x5EEDu
#endif

typedef struct WrtRand4 { uint32_t x, y, z, w; } WrtRand4;

WRT_HD WrtRand4 wrt_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                  uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    WrtRand4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

/* 24 random bits -> float in [0,1) */
WRT_HD float wrt_u01(uint32_t bits)
{
    return (float)(bits >> 8) * (1.0f / 16777216.0f);
}

/* u,v for soft-shadow sample `sample` of light `light` at ray-tree node
 * `path` (root = 1, reflection child = 2p, transmission child = 2p+1) of
 * global pixel `pixel` (y*width + x).  One Philox block serves two samples:
 * counter (pixel, path, light, sample >> 1); even samples take words x,y, odd
 * samples words z,w. */
WRT_HD void wrt_light_sample_uv(uint32_t seed, uint32_t pixel, uint32_t path,
                                uint32_t light, uint32_t sample, float* u, float* v)
{
    WrtRand4 r = wrt_philox4x32_10(pixel, path, light, sample >> 1, seed, 0x57525421u);
    if (sample & 1u) { *u = wrt_u01(r.z); *v = wrt_u01(r.w); }
    else             { *u = wrt_u01(r.x); *v = wrt_u01(r.y); }
}

/* Both samples of a pair at once (what the CUDA kernel uses). */
WRT_HD void wrt_light_sample_uv_pair(uint32_t seed, uint32_t pixel, uint32_t path, uint32_t light,
                                     uint32_t pair, float* u0, float* v0, float* u1, float* v1)
{
    WrtRand4 r = wrt_philox4x32_10(pixel, path, light, pair, seed, 0x57525421u);
    *u0 = wrt_u01(r.x); *v0 = wrt_u01(r.y);
    *u1 = wrt_u01(r.z); *v1 = wrt_u01(r.w);
}

#endif /* WRT_RNG_H */
