#include "sampling.h"
#include "ggml-b200.h"
#include <math.h>

void dnsamp_free(DenoiseSampler* S)
{
	ggml_b200_free(S->noise_dev); ggml_b200_free(S->x0_dev); free(S->noise_host);
	S->noise_dev = S->x0_dev = NULL; S->noise_host = NULL; S->n_alloc = 0;
	solver_free(&S->solver);
	free(S->sigmas); S->sigmas = NULL;
}

/* Step count and noise levels (sampling.c:28-96): multi-NFE solvers divide the step count so the
 * number of UNet evaluations stays the same; img2img scales it by the time span; sigmas are either
 * uniform in t or Karras (rho = 7) between the model's sigma(t_end) and sigma(t_ini). */
int dnsamp_init(DenoiseSampler* S)
{
	if (S->c.method <= 0) S->c.method = SOLVER_METHOD_EULER;
	S->solver.C = solver_class_get(S->c.method);
	if (!S->solver.C) FAIL(-1, "invalid sampling method %d", S->c.method);
	S->n_step = S->c.n_step < 1 ? 20 : S->c.n_step;
	S->nfe_per_step = S->solver.C->n_fe;
	if (S->nfe_per_step > 1) S->n_step = (S->n_step + S->nfe_per_step - 1) / S->nfe_per_step;
	S->nfe_per_step *= S->nfe_per_dxdt;
	if (!(S->c.f_t_ini > 0)) S->c.f_t_ini = 1;
	S->n_step = S->n_step * (S->c.f_t_ini - S->c.f_t_end) + 0.5;
	if (S->n_step < 1) S->n_step = 1;

	S->sigmas = xrealloc(S->sigmas, (S->n_step + 1) * sizeof(float));
	S->n_sigmas = S->n_step + 1;
	S->sigmas[S->n_step] = 0;
	float t_ini = (S->unet_p->n_step_train - 1) * S->c.f_t_ini;
	float t_end = (S->unet_p->n_step_train - 1) * S->c.f_t_end;
	if (!S->c.sched) S->c.sched = DNSAMP_SCHED_UNIFORM;
	if (S->c.sched == DNSAMP_SCHED_UNIFORM) {
		float b = t_ini, f = S->n_step > 1 ? (t_end - t_ini) / (S->n_step - 1) : 0;
		for (int i = 0; i < S->n_step; ++i) S->sigmas[i] = unet_t_to_sigma(S->unet_p, b + i * f);
	} else if (S->c.sched == DNSAMP_SCHED_KARRAS) {
		float smin = unet_t_to_sigma(S->unet_p, t_end), smax = unet_t_to_sigma(S->unet_p, t_ini), p = 7,
		      sminp = pow(smin, 1 / p), smaxp = pow(smax, 1 / p),
		      b = smaxp, f = S->n_step > 1 ? (sminp - smaxp) / (S->n_step - 1) : 0;
		for (int i = 0; i < S->n_step; ++i) S->sigmas[i] = pow(b + i * f, p);
	} else FAIL(-1, "invalid sampling scheduler %d", S->c.sched);

	int64_t n = S->n_per_image * S->n_rng;
	if (n > S->n_alloc) {
		ggml_b200_free(S->noise_dev); ggml_b200_free(S->x0_dev);
		S->noise_dev = ggml_b200_malloc(n * sizeof(float));
		S->x0_dev = ggml_b200_malloc(n * sizeof(float));
		S->noise_host = xrealloc(S->noise_host, n * sizeof(float));
		S->n_alloc = n;
	}
	CHECK(solver_reset(&S->solver, n));
	S->solver.t = S->sigmas[0];
	S->i_step = 0;
	return 1;
}

/* x += randn * sigma (sampling.c:112-117); image i draws from its own stream */
static void noise_add(DenoiseSampler* S, float* x, float sigma)
{
	int64_t n = S->n_per_image * S->n_rng;
	for (int i = 0; i < S->n_rng; ++i) rng_philox_randn(&S->rng[i], (unsigned)S->n_per_image, S->noise_host + i * S->n_per_image);
	ggml_b200_upload(S->noise_dev, S->noise_host, n * sizeof(float));
	float* outs[1] = { x };
	const float* ins[2] = { x, S->noise_dev };
	float c[2] = { 1, sigma };
	ggml_b200_lincomb(1, outs, 2, ins, c, n);
}

static void mask_apply(DenoiseSampler* S, float* x)
{
	int64_t n = S->n_per_image * S->n_rng;
	ggml_b200_mask_blend(x, S->x0_dev, S->c.lmask_dev, S->c.mask_pix, n / S->c.mask_pix);
}

int dnsamp_step(DenoiseSampler* S, float* x)
{
	int s = S->i_step;
	if (!(s < S->n_step)) return 0;
	int64_t n = S->n_per_image * S->n_rng;
	float s_up = 0, s_down = S->sigmas[s + 1];

	if (s == 0) {   /* sampling.c:128-136 */
		if (S->c.lmask_dev) ggml_b200_copy(S->x0_dev, x, n * sizeof(float));
		noise_add(S, x, S->sigmas[0]);
		if (S->c.lmask_dev) mask_apply(S, x);
	}
	if (S->c.s_noise > 0 && s > 0) {   /* stochastic churn (sampling.c:138-151) */
		float s_curr = S->sigmas[s], s_hat = s_curr * sqrt(2) * S->c.s_noise,
		      s_noise = sqrt(s_hat * s_hat - s_curr * s_curr);
		noise_add(S, x, s_noise);
		if (S->c.lmask_dev) mask_apply(S, x);
		S->solver.t = s_hat;
	}
	if (S->c.s_ancestral > 0) {   /* ancestral split of the step (sampling.c:153-166) */
		float s1 = S->sigmas[s], s2 = S->sigmas[s + 1];
		s_up = sqrt((s2 * s2) * (s1 * s1 - s2 * s2) / (s1 * s1));
		s_up *= S->c.s_ancestral;
		if (s_up > s2) s_up = s2;
		s_down = sqrt(s2 * s2 - s_up * s_up);
	}
	CHECK(solver_step(&S->solver, s_down, x));
	if (s_up > 0 && s + 1 != S->n_step) {
		noise_add(S, x, s_up);
		S->solver.t = S->sigmas[s + 1];
	}
	if (S->c.lmask_dev) mask_apply(S, x);
	S->i_step++;
	return 1;
}
