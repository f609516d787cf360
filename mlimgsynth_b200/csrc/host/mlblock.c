#include "mlblock.h"

void ht_resize(HTensor* T, int n0, int n1, int n2, int n3)
{
	if (!(T->flags & HT_OWNMEM)) T->d = NULL;
	T->n[0] = n0; T->n[1] = n1; T->n[2] = n2; T->n[3] = n3;
	T->d = xrealloc(T->d, ht_count(T) * sizeof(float));
	T->flags |= HT_OWNMEM;
}
void ht_free(HTensor* T) { if (T->flags & HT_OWNMEM) free(T->d); memset(T, 0, sizeof(*T)); }
size_t ht_count(const HTensor* T) { return (size_t)T->n[0] * T->n[1] * T->n[2] * T->n[3]; }
void ht_copy(HTensor* dst, const HTensor* src)
{
	ht_resize(dst, src->n[0], src->n[1], src->n[2], src->n[3]);
	memcpy(dst->d, src->d, ht_count(src) * sizeof(float));
}
int ht_finite_check(const HTensor* T)
{
	size_t n = ht_count(T);
	for (size_t i = 0; i < n; ++i) if (!(T->d[i] - T->d[i] == 0)) return -1;
	return 1;
}

static void release(MLCtx* C)
{
	if (C->allocr) { ggml_gallocr_free(C->allocr); C->allocr = NULL; }   /* before ggml_free: owns the plan */
	for (int i = 0; i < C->n_ent; ++i) { free(C->ent[i].name); free(C->ent[i].key); }
	free(C->ent); C->ent = NULL; C->n_ent = C->cap_ent = 0;
	free(C->inputs); C->inputs = NULL; C->n_inputs = C->cap_inputs = 0;
	C->result = NULL; C->graph = NULL; C->prepared = false;
	if (C->cc) { ggml_free(C->cc); ggml_free(C->cp); C->cc = C->cp = NULL; }
}

void mlctx_end(MLCtx* C) { release(C); }

void mlctx_begin(MLCtx* C, const char* name)
{
	release(C);
	if (!C->c.n_tensor_max) C->c.n_tensor_max = 16384;
	struct ggml_init_params ip = { ggml_tensor_overhead() * C->c.n_tensor_max + ggml_graph_overhead(), NULL, true };
	C->cc = ggml_init(ip);
	C->cp = ggml_init(ip);
	C->c.name = name ? name : "";
	memset(&C->info, 0, sizeof(C->info));
}

void mlctx_block_begin(MLCtx* C)
{
	MLCtxEntry e = { NULL, NULL, 1, NULL };
	ARR_PUSH(C->ent, C->n_ent, C->cap_ent, e);
}

MLTensor* mlctx_tensor_add(MLCtx* C, const char* name, MLTensor* t)
{
	MLCtxEntry e = { t, xstrdup(name), 0, NULL };
	ARR_PUSH(C->ent, C->n_ent, C->cap_ent, e);
	if (!t->name[0]) ggml_set_name(t, name);
	return t;
}

MLTensor* mlctx_input_new(MLCtx* C, const char* name, enum ggml_type dtype, int n0, int n1, int n2, int n3)
{
	MLTensor* t = ggml_new_tensor_4d(C->cp, dtype, n0, n1, n2, n3);
	ggml_set_name(t, name);
	ggml_set_input(t);
	ARR_PUSH(C->inputs, C->n_inputs, C->cap_inputs, t);
	return t;
}

/* Parameter paths. Entries are pushed in build order: a block pushes BEGIN, its children, and the
 * CALLER then pushes the block's own (name, result). Walking backwards, a named op therefore opens
 * a scope that its BEGIN marker closes; a named leaf (op == NONE) is a parameter whose key is the
 * dotted path of the open scopes plus its own name. */
static int resolve_names(MLCtx* C)
{
	char path[512]; size_t len = 0;
	size_t stack[64]; int sp = 0;
	path[0] = 0;
	for (int i = C->n_ent - 1; i >= 0; --i) {
		MLCtxEntry* e = &C->ent[i];
		if (e->kind == 1) {
			if (!sp) FAIL(-1, "%s: unbalanced module graph", C->c.name);
			len = stack[--sp]; path[len] = 0;
			continue;
		}
		size_t nl = strlen(e->name);
		if (len + nl + 2 > sizeof(path)) FAIL(-1, "parameter path too long");
		size_t old = len;
		if (len) path[len++] = '.';
		memcpy(path + len, e->name, nl + 1); len += nl;
		if (e->t->op == GGML_OP_NONE) { e->key = xstrdup(path); len = old; path[len] = 0; }
		else { if (sp == 64) FAIL(-1, "module graph too deep"); stack[sp++] = old; }
	}
	return 1;
}

static int upload_param(MLCtx* C, MLCtxEntry* e)
{
	TSEntry* s = tstore_find(C->tstore, e->key);
	if (!s) FAIL(-1, "tensor '%s' not found", e->key);
	MLTensor* t = e->t;
	if (ggml_nelements(t) != tsentry_count(s))   /* element count only, like mlblock.c:243 */
		FAIL(-1, "tensor '%s': wrong element count %lld (graph wants %lld)", e->key, (long long)tsentry_count(s), (long long)ggml_nelements(t));
	int want = t->type == GGML_TYPE_F16 ? TS_F16 : t->type == GGML_TYPE_F32 ? TS_F32 : -1;
	if (want < 0) FAIL(-1, "tensor '%s': unsupported graph type %s", e->key, ggml_type_name(t->type));
	const size_t src_es = s->dtype == TS_F32 ? 4 : 2;
	if (s->nbytes < (size_t)tsentry_count(s) * src_es)      /* never read past the entry's bytes in the mapping */
		FAIL(-1, "tensor '%s': %zu bytes stored, %zu needed", e->key, s->nbytes, (size_t)tsentry_count(s) * src_es);
	void* tmp = NULL;
	const void* src = tsentry_as(s, want, &tmp);
	if (!src) FAIL(-1, "tensor '%s': unsupported file dtype", e->key);
	ggml_backend_tensor_set(t, src, 0, ggml_nbytes(t));
	if (tmp) { ggml_b200_synchronize(); free(tmp); C->info.n_conv++; }
	C->info.mem_params += ggml_nbytes(t);
	return 1;
}

int mlctx_reload_params(MLCtx* C)
{
	double t0 = time_now();
	C->info.mem_params = 0; C->info.n_conv = 0;
	for (int i = 0; i < C->n_ent; ++i)
		if (C->ent[i].kind == 0 && C->ent[i].t->op == GGML_OP_NONE && C->ent[i].key) CHECK(upload_param(C, &C->ent[i]));
	ggml_b200_synchronize();
	C->info.t_load = time_now() - t0;
	return 1;
}

int mlctx_prep(MLCtx* C)
{
	if (!C->n_ent) FAIL(-1, "%s: empty module graph", C->c.name);
	MLTensor* result = C->ent[C->n_ent - 1].t;
	if (C->c.tprefix) mlctx_tensor_add(C, C->c.tprefix, result);
	CHECK(resolve_names(C));
	ggml_set_output(result);
	C->result = result;
	C->graph = ggml_new_graph_custom(C->cc, C->c.n_tensor_max, false);
	ggml_build_forward_expand(C->graph, result);
	C->allocr = ggml_gallocr_new(ggml_backend_get_default_buffer_type(C->backend));
	if (!ggml_gallocr_reserve(C->allocr, C->graph) || !ggml_gallocr_alloc_graph(C->allocr, C->graph))
		FAIL(-1, "%s: could not allocate device memory", C->c.name);
	C->info.mem_total = ggml_gallocr_get_buffer_size(C->allocr, 0);
	CHECK(mlctx_reload_params(C));
	if (!(C->c.flags & MLB_F_QUIET))
		log_info("%s: %d nodes, params %.1f MiB loaded in %.3fs (converted %u)", C->c.name, ggml_graph_n_nodes(C->graph),
			C->info.mem_params / 1048576.0, C->info.t_load, C->info.n_conv);
	C->prepared = true;
	return 1;
}

int mlctx_compute(MLCtx* C)
{
	int r = ggml_backend_graph_compute(C->backend, C->graph);
	C->info.n_compute++;
	if (r) FAIL(-1, "%s: graph compute failed (%d)", C->c.name, r);
	return 1;
}
