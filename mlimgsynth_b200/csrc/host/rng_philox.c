#include "rng_philox.h"
#include <math.h>

/* One Philox4x32 block per output: counter = (offset, 0, index, 0), key = seed (lo, hi),
 * ten rounds; words 0 and 1 feed Box-Muller evaluated in double, result rounded to float
 * (rng_philox.c:16-51). The offset advances by one per call, not per number. */
static inline void philox_round(uint32_t c[4], const uint32_t k[2])
{
	const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
	const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
	const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
	const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
	c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
}

void rng_philox_randn(RngPhilox* S, unsigned n, float* out)
{
	const double inv32 = 2.3283064365386963e-10;      /* 2^-32 */
	const double tau32 = 1.4629180792671596e-09;      /* 2 pi 2^-32 */
	for (unsigned i = 0; i < n; ++i) {
		uint32_t c[4] = { S->offset, 0, i, 0 };
		uint32_t k[2] = { (uint32_t)S->seed, (uint32_t)(S->seed >> 32) };
		for (int r = 0; r < 10; ++r) {
			philox_round(c, k);
			k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
		}
		double u = ((double)c[0] + 0.5) * inv32, v = ((double)c[1] + 0.5) * tau32;
		out[i] = (float)(sqrt(-2.0 * log(u)) * sin(v));
	}
	S->offset++;
}
