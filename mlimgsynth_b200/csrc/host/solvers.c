#include "solvers.h"
#include "ggml-b200.h"
#include <math.h>

static const SolverClass* const g_solvers[] = { NULL, &g_solver_euler, &g_solver_heun, &g_solver_taylor3, &g_solver_dpmpp2m, &g_solver_dpmpp2s };

const SolverClass* solver_class_get(int idx) { return idx >= 1 && idx <= 5 ? g_solvers[idx] : NULL; }
const SolverClass* solver_class_find(const char* name)
{
	for (int i = 1; i <= 5; ++i) if (!strcmp(name, g_solvers[i]->name)) return g_solvers[i];
	return NULL;
}

int solver_reset(Solver* S, int64_t n)
{
	if (n > S->n_alloc) {
		solver_free(S);
		S->dx = ggml_b200_malloc(n * sizeof(float));
		for (int i = 0; i < 4; ++i) S->tmp[i] = ggml_b200_malloc(n * sizeof(float));
		S->n_alloc = n;
	}
	S->n = n;
	ggml_b200_memset(S->dx, 0, n * sizeof(float));
	for (int i = 0; i < 4; ++i) ggml_b200_memset(S->tmp[i], 0, n * sizeof(float));
	memset(S->var, 0, sizeof(S->var));
	S->i_step = 0;
	return 1;
}

void solver_free(Solver* S)
{
	ggml_b200_free(S->dx); S->dx = NULL;
	for (int i = 0; i < 4; ++i) { ggml_b200_free(S->tmp[i]); S->tmp[i] = NULL; }
	S->n_alloc = 0;
}

int solver_step(Solver* S, float t, float* x)
{
	int r = S->C->step(S, t, x);
	if (r < 0) return r;
	S->t = t;
	S->i_step++;
	return r;
}

/* out_j = sum_i coef[j][i] * in_i over the n state elements */
static void lin(Solver* S, int n_out, float* o0, float* o1, float* o2, int n_in, const float* i0, const float* i1, const float* i2,
	const float* i3, const float* coef)
{
	float* outs[3] = { o0, o1, o2 };
	const float* ins[4] = { i0, i1, i2, i3 };
	ggml_b200_lincomb(n_out, outs, n_in, ins, coef, S->n);
}

/* Euler: x += dx * dt  (solvers.c:82-88) */
static int euler_step(Solver* S, float t, float* x)
{
	float dt = t - S->t;
	CHECK(S->dxdt(S, S->t, x, S->dx));
	float c[2] = { 1, dt };
	lin(S, 1, x, 0, 0, 2, x, S->dx, 0, 0, c);
	return 1;
}
const SolverClass g_solver_euler = { euler_step, 1, "euler" };

/* Heun (solvers.c:100-122): predictor x1 = x + dx dt; last step (t == 0) keeps the Euler result,
 * otherwise x += (dx + d1) * 0.5 * dt */
static int heun_step(Solver* S, float t, float* x)
{
	float dt = t - S->t;
	float *x1 = S->tmp[0], *d1 = S->tmp[1];
	CHECK(S->dxdt(S, S->t, x, S->dx));
	float c1[2] = { 1, dt };
	if (!(t > 0)) { lin(S, 1, x, 0, 0, 2, x, S->dx, 0, 0, c1); return 1; }
	lin(S, 1, x1, 0, 0, 2, x, S->dx, 0, 0, c1);
	CHECK(S->dxdt(S, t, x1, d1));
	float h = (float)(0.5 * dt);
	float c2[3] = { 1, h, h };
	lin(S, 1, x, 0, 0, 3, x, S->dx, d1, 0, c2);
	return 1;
}
const SolverClass g_solver_heun = { heun_step, 2, "heun" };

/* Third-order Taylor extension of Euler (solvers.c:124-170), finite-difference derivatives from the
 * two previous slopes: d2 = (dx - dp1)/dt_prev, d3 = (d2 - dp2)/dt_prev,
 * x += dx dt + d2 dt^2/2 + d3 dt^3/6; history dp1 <- dx, dp2 <- d2. */
static int taylor3_step(Solver* S, float t, float* x)
{
	float dt = t - S->t;
	float *dp1 = S->tmp[0], *dp2 = S->tmp[1];
	CHECK(S->dxdt(S, S->t, x, S->dx));
	float idtp = S->i_step >= 1 ? 1 / S->var[0] : 0,
	      f2 = S->i_step >= 1 ? dt * dt / 2 : 0,
	      f3 = S->i_step >= 2 ? dt * dt * dt / 6 : 0;
	float k = f2 * idtp + f3 * idtp * idtp;
	/* inputs: x, dx, dp1, dp2 -> outputs: x, dp1, dp2 */
	float c[12] = {
		1, dt + k, -k, -f3 * idtp,
		0, 1, 0, 0,
		0, idtp, -idtp, 0 };
	lin(S, 3, x, dp1, dp2, 4, x, S->dx, dp1, dp2, c);
	S->var[0] = dt;
	return 1;
}
const SolverClass g_solver_taylor3 = { taylor3_step, 1, "taylor3" };

/* DPM-Solver++(2M) (solvers.c:172-236): a = s_next/s, h = -log a, c = h / (2 h_last) (0 on the first
 * and on the last step); d0 = x - s dx; D = (1+c) d0 - c d_prev; x = a x + (1-a) D; d_prev <- d0. */
static int dpmpp2m_step(Solver* S, float t, float* x)
{
	float* dprev = S->tmp[0];
	float a = t / S->t, h = -(float)log(a), h_last = S->var[0], c = h / (2 * h_last);
	if (S->i_step == 0 || !(t > 0)) c = 0;
	CHECK(S->dxdt(S, S->t, x, S->dx));
	float w = (1 - a) * (1 + c);
	/* inputs: x, dx, dprev -> outputs: x, dprev */
	float m[6] = {
		a + w, -w * S->t, -(1 - a) * c,
		1, -S->t, 0 };
	lin(S, 2, x, dprev, 0, 3, x, S->dx, dprev, 0, m);
	S->var[0] = h;
	return 1;
}
const SolverClass g_solver_dpmpp2m = { dpmpp2m_step, 1, "dpmpp2m" };

/* DPM-Solver++(2S) (solvers.c:238-296): midpoint s1 = sqrt(s_next s); x1 = x + dx (s1 - s);
 * d = x1 - s1 dx1; x = a x + (1-a) d; Euler on the last step. */
static int dpmpp2s_step(Solver* S, float t, float* x)
{
	float *x1 = S->tmp[0], *dx1 = S->tmp[1];
	CHECK(S->dxdt(S, S->t, x, S->dx));
	if (!(t > 0)) {
		float c[2] = { 1, t - S->t };
		lin(S, 1, x, 0, 0, 2, x, S->dx, 0, 0, c);
		return 1;
	}
	float t1 = (float)sqrt(t * S->t), dt1 = t1 - S->t, a = t / S->t;
	float c1[2] = { 1, dt1 };
	lin(S, 1, x1, 0, 0, 2, x, S->dx, 0, 0, c1);
	CHECK(S->dxdt(S, t1, x1, dx1));
	float c2[3] = { a, 1 - a, -(1 - a) * t1 };
	lin(S, 1, x, 0, 0, 3, x, x1, dx1, 0, c2);
	return 1;
}
const SolverClass g_solver_dpmpp2s = { dpmpp2s_step, 2, "dpmpp2s" };
