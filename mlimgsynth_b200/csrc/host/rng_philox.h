/* rng_philox.h -- Philox4x32-10 + Box-Muller normal stream, bit-compatible with the reference's
 * ccommon/rng_philox.c (which imitates torch's CUDA randn as in A1111's rng_philox.py).
 * Host-side on purpose: the noise stream must be bit-exact (BASELINE.json north_star), and the
 * double-precision log/sin of Box-Muller are not guaranteed bit-identical between glibc and CUDA. */
#pragma once
#include <stdint.h>

typedef struct RngPhilox { uint64_t seed; uint32_t offset; } RngPhilox;
void rng_philox_randn(RngPhilox* S, unsigned n, float* out);
