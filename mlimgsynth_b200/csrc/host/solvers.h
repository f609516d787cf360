/* solvers.h -- initial-value-problem solvers used as diffusion sampling methods, operating on a
 * DEVICE-resident state. Same classes, NFE counts, history handling and last-step fallbacks as
 * the reference's solvers.c (euler :82, heun :100, taylor3 :137, dpmpp2m :207, dpmpp2s :264);
 * every element-wise loop there is a linear combination with scalar coefficients, so each solver
 * stage is ONE fused launch of ggml_b200_lincomb with coefficients computed on the host in the
 * reference's float arithmetic. */
#pragma once
#include "base.h"

struct Solver;
typedef struct SolverClass {
	int (*step)(struct Solver*, float t, float* x);
	int n_fe;               /* calls to dxdt per step */
	const char* name;
} SolverClass;

extern const SolverClass g_solver_euler, g_solver_heun, g_solver_taylor3, g_solver_dpmpp2m, g_solver_dpmpp2s;

enum { SOLVER_METHOD_EULER = 1, SOLVER_METHOD_HEUN = 2, SOLVER_METHOD_TAYLOR3 = 3, SOLVER_METHOD_DPMPP2M = 4, SOLVER_METHOD_DPMPP2S = 5 };
const SolverClass* solver_class_get(int idx);
const SolverClass* solver_class_find(const char* name);

typedef struct Solver {
	const SolverClass* C;     /* fill before use */
	float* dx;                /* device, n elements */
	float* tmp[4];            /* device history / scratch tensors */
	float  var[4];            /* host scalars carried across steps */
	int64_t n, n_alloc;       /* elements per state tensor */
	float t;
	unsigned i_step;
	/* dx/dt at time t for the device state x -> device dx; return 0 to skip (t < 0) */
	int (*dxdt)(struct Solver*, float t, const float* x, float* dx);
	void* user;
} Solver;

int  solver_reset(Solver* S, int64_t n);      /* (re)allocate and zero the state tensors */
void solver_free(Solver* S);
int  solver_step(Solver* S, float t, float* x);
