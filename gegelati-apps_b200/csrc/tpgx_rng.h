/*
 * tpgx_rng.h — the environments' random draws, restated for host and device.
 *
 * [UP] Mutator::RNG wraps std::mt19937_64 and draws through std::uniform_int_distribution<uint64_t>
 * and std::uniform_real_distribution<double>.  The episode kernels do not
 * run the Mersenne twister on the device: because [UP] LearningAgent::evaluateJob seeds every root's
 * episode identically, the host generates the raw 64-bit engine outputs once per iteration
 * (tpgx_rng_fill_raw) and each episode walks that shared table with its own cursor.  The two
 * distribution transforms below restate libstdc++ (GCC >= 11):
 *   uniform_int_distribution  — Lemire's nearly-divisionless bounded draw on a 64x64->128 product
 *                               (bits/uniform_int_dist.h, _S_nd), range == 2^64-1 returns the raw draw;
 *   uniform_real_distribution — generate_canonical<double,53> = (double)raw / 2^64 clipped below 1,
 *                               then canonical * (b - a) + a.
 */
#ifndef TPGX_RNG_H
#define TPGX_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define TPGX_RNG_HD __host__ __device__ __forceinline__
#else
#define TPGX_RNG_HD static inline
#endif

#define TPGX_RNG_TABLE 256 /* raw draws available to one episode */

struct tpgx_rng_cursor {
    const uint64_t* raw;
    int pos;
    int exhausted;
};

TPGX_RNG_HD uint64_t tpgx_rng_next(tpgx_rng_cursor& c) {
    if (c.pos >= TPGX_RNG_TABLE) { c.exhausted = 1; return 0; }
    return c.raw[c.pos++];
}

TPGX_RNG_HD void tpgx_mul64(uint64_t a, uint64_t b, uint64_t* hi, uint64_t* lo) {
#if defined(__CUDA_ARCH__)
    *lo = a * b;
    *hi = __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    *lo = (uint64_t)p;
    *hi = (uint64_t)(p >> 64);
#endif
}

/* [UP] Mutator::RNG::getUnsignedInt64(lo, hi), inclusive bounds */
TPGX_RNG_HD uint64_t tpgx_rng_u64(tpgx_rng_cursor& c, uint64_t lo, uint64_t hi) {
    const uint64_t urange = hi - lo;
    if (urange == 0xffffffffffffffffULL) return tpgx_rng_next(c);
    const uint64_t range = urange + 1;
    uint64_t ph, pl;
    tpgx_mul64(tpgx_rng_next(c), range, &ph, &pl);
    if (pl < range) {
        const uint64_t threshold = (0 - range) % range;
        while (pl < threshold && !c.exhausted) tpgx_mul64(tpgx_rng_next(c), range, &ph, &pl);
    }
    return ph + lo;
}

/* [UP] Mutator::RNG::getDouble(lo, hi) */
TPGX_RNG_HD double tpgx_rng_f64(tpgx_rng_cursor& c, double lo, double hi) {
    const uint64_t r = tpgx_rng_next(c);
#if defined(__CUDA_ARCH__)
    double canon = __ull2double_rn(r) / 18446744073709551616.0;
#else
    double canon = (double)r / 18446744073709551616.0;
#endif
    if (canon >= 1.0) canon = 0.99999999999999988898; /* nextafter(1.0, 0.0) */
    return canon * (hi - lo) + lo;
}

/* The environments seed their RNG with Data::Hash(seed) ^ Data::Hash(mode) ([REF] pendulum/src/Learn/pendulum.cpp:45).
 * A caller that links GEGELATI passes that value itself (tpgx_episode_request.rng_seeds: the C++ adapter does); this is
 * the default for callers that do not, assuming [UP] Data::Hash<T> = std::hash<T>, the identity on libstdc++ — an
 * assumption about upstream, "parity unpinned". */
TPGX_RNG_HD uint64_t tpgx_env_seed(uint64_t seed, int mode) { return seed ^ (uint64_t)mode; }

#endif
