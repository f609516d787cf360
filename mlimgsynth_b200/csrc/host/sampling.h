/* sampling.h -- denoising sampler: sigma schedule, initial / churn / ancestral noise, inpainting
 * mask, solver stepping (reference: sampling.c). State lives on the device; noise is drawn on the
 * host with the bit-exact Philox stream (one RNG per image of the batch: image i = seed + i,
 * fresh offset, the reference's batch semantics via generate.sh:55-61) and uploaded. */
#pragma once
#include "solvers.h"
#include "unet.h"
#include "rng_philox.h"

enum { DNSAMP_SCHED_UNIFORM = 1, DNSAMP_SCHED_KARRAS = 2 };

typedef struct DenoiseSampler {
	Solver solver;
	float* sigmas; int n_sigmas;
	int i_step, n_step, nfe_per_step;
	const UnetParams* unet_p;    /* fill before use */
	int nfe_per_dxdt;            /* fill before use */
	RngPhilox* rng; int n_rng;   /* fill before use: one stream per image */
	int64_t n_per_image;         /* fill before use: latent elements per image */
	float *noise_dev, *x0_dev, *noise_host; int64_t n_alloc;
	struct {
		int n_step, method, sched;
		float f_t_ini, f_t_end, s_noise, s_ancestral;
		const float* lmask_dev;  /* device [w*h] inpainting mask or NULL */
		int64_t mask_pix;
	} c;
} DenoiseSampler;

void dnsamp_free(DenoiseSampler* S);
int  dnsamp_init(DenoiseSampler* S);
int  dnsamp_step(DenoiseSampler* S, float* x_dev);   /* > 0 while steps remain, 0 when done, < 0 on error */
