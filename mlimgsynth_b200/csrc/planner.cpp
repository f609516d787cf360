// planner.cpp -- turns a recorded ggml-shaped graph into a list of fused sm_100a kernel launches.
//
// The reference executes its graphs node by node on ggml (mlblock.c:294-307). Here the node list
// is planned once per graph (first ggml_backend_graph_compute) and replayed as a CUDA graph:
//   * every value is a strided view (PT) over an SSA buffer; reshape / permute / transpose / view
//     and `cont` of a dense permutation are pure stride arithmetic, so the reference's
//     NCHW <-> token-major `permute+cont` pairs (unet.c:126-127,135-137, mlblock_nn.c:204-226)
//     and the head split/merge copies vanish;
//   * activations are stored channels-last in f16 ([N*H*W, C] rows): the A operand of every
//     contraction is K-contiguous, which is what TMA + tcgen05 want;
//   * peephole fusion: conv/linear + bias (+ time-embedding add, + activation, + residual) into
//     the GEMM epilogue; group_norm*w+b(+silu) and norm*w+b into single kernels; the
//     mul_mat/scale/mask/softmax/mul_mat chain of ggml_nn_attention into one attention kernel;
//     the GEGLU view/cont/gelu/mul chain into one gate kernel;
//   * weights are re-laid-out once (conv kernels to [Cout][kh][kw][Cin]) and cached per version;
//   * buffers are assigned by liveness into one arena so the working set stays L2-resident.
#include "engine.h"
#include "kernels.h"
#include <algorithm>
#include <cmath>
#include <functional>
#include <map>
#include <unordered_map>

namespace b200 {

// ------------------------------------------------------------------ plan data
enum BufKind { BUF_ARENA, BUF_FIXED, BUF_PERSIST };
struct Buf { BufKind kind; size_t bytes; void* fixed; int first, last; size_t off; };

struct PT {                 // planned tensor: strided view over a buffer
	int buf = -1;
	int64_t off = 0;        // elements
	DT dt = DT_F32;
	int64_t ne[4] = {1, 1, 1, 1};
	int64_t st[4] = {0, 0, 0, 0};
	int64_t numel() const { return ne[0] * ne[1] * ne[2] * ne[3]; }
};

enum StepKind {
	S_COPY, S_BINARY, S_UNARY, S_UPSCALE, S_SOFTMAX, S_GET_ROWS, S_TSEMB, S_GEMM_SIMT,
	S_GROUPNORM, S_LAYERNORM, S_GEGLU, S_IM2COL, S_WPREP_CONV, S_ATTENTION, S_GEMM_TC, S_CONV_TC, S_ZERO, S_SOFTMAX_F16, S_WPREP_GEGLU,
};

struct Step {
	StepKind kind;
	PT out, in[4];
	int n_in = 0;
	int iop = 0; float fparam = 0; int iparam[8] = {0};
	// GEMM epilogue sources
	PT bias, rowvec, residual; bool has_bias = false, has_rowvec = false, has_residual = false;
	int64_t rows_per_image = 0;
	UnaryOp act = U_NONE;
	int64_t M = 0, N = 0, K = 0, lda = 0, ldb = 0, ldc = 0;
	int64_t conv_n = 0, conv_h = 0, conv_w = 0, conv_c = 0;
	GemmTC* tc = nullptr;
	AttnTC* atc = nullptr; bool atc_checked = false;
	size_t stats_off = 0;               // groupnorm statistics slot in the zero region
	// GroupNorm statistics accumulated by the producing GEMM / conv epilogue (mlblock_nn.c:78): the producer step carries
	// gn_groups > 0 and the slot, the consumer the producer's index; whether the launch really does it is known once the
	// producer is prepared (gemm_tc_gn_fused)
	int gn_groups = 0; int64_t gn_rows_per_image = 0; int gn_producer = -1;
	const ggml_tensor* leaf = nullptr;  // weight prep: source leaf (+ version it was prepared at)
	uint64_t leaf_version = ~0ull;
	const char* name = "";
};

struct Plan {
	Backend* be;
	ggml_cgraph* graph;
	std::vector<Buf> bufs;
	std::vector<Step> prep;   // weight preparation, re-run when the leaf version changes
	std::vector<Step> steps;  // the per-compute schedule
	void* arena = nullptr; size_t arena_bytes = 0;
	void* persist = nullptr; size_t persist_bytes = 0;
	size_t zero_off = 0, zero_bytes = 0;   // region of the arena cleared at the start of each run
	int zero_buf = -1;

	cudaGraphExec_t exec = nullptr;
	bool use_graph = true;
	uint64_t launches_per_run = 0;
};

bool g_dryrun();
// ---- per-kernel-family profile (ggml_b200_profile_*): eager run with CUDA events around each step
struct ProfAcc { double ms = 0, flops = 0, bytes = 0; uint64_t launches = 0; };
static bool g_profile = false;
static ProfAcc g_prof[32];
void profile_enable(bool on) { g_profile = on; if (on) for (auto& a : g_prof) a = ProfAcc(); }
bool profile_get(int kind, double* ms, double* flops, double* bytes, uint64_t* launches)
{
	if (kind < 0 || kind >= 32) return false;
	*ms = g_prof[kind].ms; *flops = g_prof[kind].flops; *bytes = g_prof[kind].bytes; *launches = g_prof[kind].launches;
	return true;
}

static bool env_flag(const char* n) { const char* e = getenv(n); return e && *e && *e != '0'; }

// ------------------------------------------------------------------ PT helpers
static DT dt_of(ggml_type t)
{
	switch (t) {
	case GGML_TYPE_F32: return DT_F32;
	case GGML_TYPE_F16: return DT_F16;
	case GGML_TYPE_I32: return DT_I32;
	default: B200_FATAL("tensor type %s cannot be used in a compute graph of the B200 engine", ggml_type_name(t));
	}
}

static void contiguous_strides(PT& p)
{
	p.st[0] = 1;
	for (int i = 1; i < 4; ++i) p.st[i] = p.st[i-1] * p.ne[i-1];
}
static bool is_ggml_contig(const PT& p)
{
	int64_t s = 1;
	for (int i = 0; i < 4; ++i) { if (p.ne[i] != 1 && p.st[i] != s) return false; s *= p.ne[i]; }
	return true;
}
// dense = a permutation of a compact buffer (every element addressed exactly once, no gaps)
static bool is_dense(const PT& p)
{
	int ord[4] = {0, 1, 2, 3};
	std::sort(ord, ord + 4, [&](int a, int b) { return p.st[a] < p.st[b]; });
	int64_t s = 1;
	for (int k = 0; k < 4; ++k) { int i = ord[k]; if (p.ne[i] == 1) continue; if (p.st[i] != s) return false; s *= p.ne[i]; }
	return true;
}
static bool same_shape(const PT& a, const PT& b) { for (int i = 0; i < 4; ++i) if (a.ne[i] != b.ne[i]) return false; return true; }

// Try to view `a` with a new shape without moving data (torch-style stride computation, dim 0 fastest).
static bool try_reshape(const PT& a, const int64_t ne[4], PT& out)
{
	out = a;
	for (int i = 0; i < 4; ++i) out.ne[i] = ne[i];
	int64_t ost[4];
	int vd = 0;                       // next new dim to assign
	int64_t chunk_base = 0, chunk_numel = 1, view_numel = 1;
	bool in_chunk = false;
	int last_old = -1;
	for (int i = 0; i < 4; ++i) if (a.ne[i] != 1) last_old = i;
	if (last_old < 0) {               // scalar-like: any strides work
		int64_t s = 1; for (int i = 0; i < 4; ++i) { out.st[i] = s; s *= ne[i]; }
		return true;
	}
	int prev = -1;
	for (int i = 0; i <= last_old; ++i) {
		if (a.ne[i] == 1) continue;
		if (!in_chunk) { chunk_base = a.st[i]; chunk_numel = 1; in_chunk = true; }
		chunk_numel *= a.ne[i];
		// is the next non-unit old dim mergeable with this one?
		int nx = -1;
		for (int j = i + 1; j <= last_old; ++j) if (a.ne[j] != 1) { nx = j; break; }
		bool mergeable = nx >= 0 && a.st[nx] == a.st[i] * a.ne[i];
		(void)prev; prev = i;
		if (mergeable) continue;
		// close the chunk: consume new dims until their product matches
		while (vd < 4 && (view_numel < chunk_numel || ne[vd] == 1)) {
			ost[vd] = chunk_base * view_numel;
			view_numel *= ne[vd];
			vd++;
		}
		if (view_numel != chunk_numel) return false;
		in_chunk = false; view_numel = 1;
	}
	for (; vd < 4; ++vd) { if (ne[vd] != 1) return false; ost[vd] = vd ? ost[vd-1] * ne[vd-1] : 1; }
	for (int i = 0; i < 4; ++i) out.st[i] = ost[i];
	return true;
}

// ------------------------------------------------------------------ builder
struct Builder {
	Plan* P;
	ggml_cgraph* g;
	std::unordered_map<const ggml_tensor*, PT> val;
	std::unordered_map<const ggml_tensor*, std::vector<ggml_tensor*>> users;
	std::unordered_map<const ggml_tensor*, bool> done;
	std::unordered_map<const ggml_tensor*, PT> prepared;   // leaf -> prepared weight
	std::map<std::string, PT> memo;                        // CSE of conversions / unary ops on SSA values
	bool force_simt;

	static std::string pt_key(const char* tag, const PT& p, int iop = 0, float f = 0)
	{
		char b[256];
		snprintf(b, sizeof(b), "%s|%d|%lld|%d|%lld,%lld,%lld,%lld|%lld,%lld,%lld,%lld|%d|%a", tag, p.buf, (long long)p.off, (int)p.dt,
			(long long)p.ne[0], (long long)p.ne[1], (long long)p.ne[2], (long long)p.ne[3],
			(long long)p.st[0], (long long)p.st[1], (long long)p.st[2], (long long)p.st[3], iop, (double)f);
		return b;
	}

	int new_buf(BufKind k, size_t bytes, void* fixed = nullptr)
	{
		P->bufs.push_back(Buf{k, (bytes + 255) / 256 * 256, fixed, -1, -1, 0});
		return (int)P->bufs.size() - 1;
	}
	PT new_pt(DT dt, const int64_t ne[4], BufKind k = BUF_ARENA)
	{
		PT p; p.dt = dt;
		for (int i = 0; i < 4; ++i) p.ne[i] = ne[i];
		contiguous_strides(p);
		p.buf = new_buf(k, (size_t)p.numel() * dt_size(dt));
		return p;
	}
	// new tensor with the same dim ordering (stride order) as `like`, densely packed
	PT new_pt_like(DT dt, const int64_t ne[4], const PT& like)
	{
		PT p; p.dt = dt;
		for (int i = 0; i < 4; ++i) p.ne[i] = ne[i];
		int ord[4] = {0, 1, 2, 3};
		std::stable_sort(ord, ord + 4, [&](int a, int b) {
			int64_t sa = like.ne[a] == 1 ? INT64_MAX : like.st[a], sb = like.ne[b] == 1 ? INT64_MAX : like.st[b];
			return sa < sb; });
		int64_t s = 1;
		for (int k = 0; k < 4; ++k) { p.st[ord[k]] = s; s *= ne[ord[k]]; }
		p.buf = new_buf(BUF_ARENA, (size_t)p.numel() * dt_size(dt));
		return p;
	}
	// channels-last image tensor [W,H,C,N]: memory order C, W, H, N
	PT new_pt_nhwc(DT dt, int64_t W, int64_t H, int64_t C, int64_t N)
	{
		PT p; p.dt = dt; p.ne[0] = W; p.ne[1] = H; p.ne[2] = C; p.ne[3] = N;
		p.st[2] = 1; p.st[0] = C; p.st[1] = C * W; p.st[3] = C * W * H;
		p.buf = new_buf(BUF_ARENA, (size_t)p.numel() * dt_size(dt));
		return p;
	}

	PT leaf_pt(const ggml_tensor* t)
	{
		PT p; p.dt = dt_of(t->type);
		size_t es = ggml_type_size(t->type);
		for (int i = 0; i < 4; ++i) { p.ne[i] = t->ne[i]; p.st[i] = (int64_t)(t->nb[i] / es); }
		if (!t->data) B200_FATAL("leaf tensor '%s' has no storage: call ggml_gallocr_alloc_graph before compute", t->name);
		p.buf = new_buf(BUF_FIXED, ggml_nbytes(t), t->data);
		return p;
	}
	PT get(const ggml_tensor* t)
	{
		auto it = val.find(t);
		if (it != val.end()) return it->second;
		if (t->op == GGML_OP_NONE) {
			if (t->view_src) B200_FATAL("views of leaves are not expected here");
			PT p = leaf_pt(t); val[t] = p; return p;
		}
		B200_FATAL("planner: value of node '%s' (%s) requested before it was planned", t->name, ggml_op_name(t->op));
	}

	Step& emit(StepKind k, const char* name)
	{
		P->steps.emplace_back();
		Step& s = P->steps.back();
		s.kind = k; s.name = name;
		return s;
	}
	void copy(const PT& dst, const PT& src, const char* name = "copy")
	{
		Step& s = emit(S_COPY, name); s.out = dst; s.in[0] = src; s.n_in = 1;
	}
	// materialise `p` as ggml-contiguous (optionally converting dtype)
	PT to_contig(const PT& p, DT dt)
	{
		if (is_ggml_contig(p) && p.dt == dt) return p;
		PT d = new_pt(dt, p.ne);
		copy(d, p, "to_contig");
		return d;
	}
	// rows [K, R...] with K contiguous and a uniform row pitch, f16 (GEMM A operand)
	bool rows_uniform(const PT& p, int64_t& rows, int64_t& pitch)
	{
		if (p.st[0] != 1 && p.ne[0] != 1) return false;
		rows = p.ne[1] * p.ne[2] * p.ne[3];
		pitch = 0;
		int64_t expect = -1;
		for (int i = 1; i < 4; ++i) {
			if (p.ne[i] == 1) continue;
			if (expect < 0) pitch = p.st[i];
			else if (p.st[i] != expect) return false;
			expect = p.st[i] * p.ne[i];
		}
		if (pitch == 0) pitch = p.ne[0];
		return true;
	}
	PT as_gemm_rows(const PT& p)
	{
		int64_t rows, pitch;
		if (p.dt == DT_F16 && rows_uniform(p, rows, pitch) && pitch % 8 == 0 && (p.off % 8) == 0) return p;
		std::string key = pt_key("rows", p);
		auto it = memo.find(key);
		if (it != memo.end()) return it->second;
		PT d = new_pt(DT_F16, p.ne);
		if (d.ne[0] % 8) {   // pad the row pitch so TMA strides stay 16-byte multiples
			int64_t kp = (d.ne[0] + 7) / 8 * 8;
			d.st[1] = kp; d.st[2] = kp * d.ne[1]; d.st[3] = d.st[2] * d.ne[2];
			P->bufs[d.buf].bytes = ((size_t)kp * d.ne[1] * d.ne[2] * d.ne[3] * 2 + 255) / 256 * 256;
		}
		copy(d, p, "to_f16_rows");
		memo[key] = d;
		return d;
	}
	// image tensor as channels-last f16
	PT as_nhwc_f16(const PT& p)
	{
		bool ok = p.dt == DT_F16 && (p.st[2] == 1 || p.ne[2] == 1) && p.st[0] == p.ne[2] && p.st[1] == p.ne[2] * p.ne[0] &&
			(p.ne[3] == 1 || p.st[3] == p.ne[2] * p.ne[0] * p.ne[1]) && (p.off % 8) == 0;
		if (ok) return p;
		std::string key = pt_key("nhwc", p);
		auto it = memo.find(key);
		if (it != memo.end()) return it->second;
		PT d = new_pt_nhwc(DT_F16, p.ne[0], p.ne[1], p.ne[2], p.ne[3]);
		copy(d, p, "to_nhwc_f16");
		memo[key] = d;
		return d;
	}

	bool single_user(const ggml_tensor* t, ggml_tensor** u)
	{
		auto it = users.find(t);
		if (it == users.end() || it->second.size() != 1) return false;
		if (t->flags & GGML_TENSOR_FLAG_OUTPUT) return false;
		*u = it->second[0];
		return true;
	}
	// strip reshape wrappers (bias reshaped to [1,1,C,1] etc.)
	static const ggml_tensor* strip_reshape(const ggml_tensor* t)
	{
		while (t->op == GGML_OP_RESHAPE) t = t->src[0];
		return t;
	}
	// b is a per-channel vector for a (channel = a's dim `cdim`)
	static bool is_channel_vec(const ggml_tensor* b, const ggml_tensor* a, int cdim)
	{
		const ggml_tensor* r = strip_reshape(b);
		if (r->op != GGML_OP_NONE || r->type != GGML_TYPE_F32) return false;
		for (int i = 0; i < 4; ++i) if (b->ne[i] != (i == cdim ? a->ne[cdim] : 1)) return false;
		return true;
	}

	void plan();
	void plan_one(ggml_tensor* t);
	bool try_geglu(ggml_tensor* t);
	bool match_geglu(const ggml_tensor* h, ggml_tensor** xv, ggml_tensor** gv, ggml_tensor** c, ggml_tensor** ge, ggml_tensor** mu);
	void ensure_planned(const ggml_tensor* t);
	void plan_node(ggml_tensor* t);
	void plan_mul_mat(ggml_tensor* t);
	void plan_conv(ggml_tensor* t);
	bool try_attention(ggml_tensor* t);
	void plan_gemm_common(ggml_tensor* last, Step& s, int cdim);
	void finish(ggml_tensor* t, const PT& p);
	ggml_tensor* absorb_epilogue(ggml_tensor* t, Step& s, int cdim, const PT& outshape_like);
};

// After a node's value is known: materialise it into its logical ggml storage if the caller can
// read it (OUTPUT flag), or if it is an in-place update of a leaf.
void Builder::finish(ggml_tensor* t, const PT& p)
{
	val[t] = p;
	done[t] = true;
	bool need = (t->flags & GGML_TENSOR_FLAG_OUTPUT) != 0;
	if (t->view_src && t->view_src->op == GGML_OP_NONE &&
		(t->op == GGML_OP_ADD || t->op == GGML_OP_SCALE || t->op == GGML_OP_UNARY || t->op == GGML_OP_SOFT_MAX || t->op == GGML_OP_DIAG_MASK_INF))
		need = true;   // in-place op on a leaf: ggml semantics update the leaf's memory
	if (!need) return;
	if (t->op == GGML_OP_RESHAPE || t->op == GGML_OP_VIEW || t->op == GGML_OP_PERMUTE || t->op == GGML_OP_TRANSPOSE) return;
	if (!t->data) B200_FATAL("output tensor '%s' has no logical storage", t->name);
	PT d; d.dt = dt_of(t->type);
	size_t es = ggml_type_size(t->type);
	for (int i = 0; i < 4; ++i) { d.ne[i] = t->ne[i]; d.st[i] = (int64_t)(t->nb[i] / es); }
	d.buf = new_buf(BUF_FIXED, ggml_nbytes(t), t->data);
	copy(d, p, "to_output");
}

// Walk forward from GEMM/conv node t absorbing: +bias, +per-image vector, activation, +residual.
// Returns the last absorbed node. cdim = channel dim of the logical result (0 for mul_mat, 2 for conv).
ggml_tensor* Builder::absorb_epilogue(ggml_tensor* t, Step& s, int cdim, const PT& out_like)
{
	ggml_tensor* cur = t;
	for (;;) {
		ggml_tensor* u;
		if (!single_user(cur, &u)) break;
		if (u->op == GGML_OP_ADD && u->src[0] == cur && !s.has_residual) {
			const ggml_tensor* b = u->src[1];
			if (!s.has_bias && !s.has_rowvec && s.act == U_NONE && is_channel_vec(b, cur, cdim)) {
				s.bias = get(strip_reshape(b)); s.has_bias = true;
				cur = u; done[u] = true; continue;
			}
			// per-image vector: [1,1,C,N] over [W,H,C,N] (resnet emb add, mlblock_nn.c:139-144)
			if (cdim == 2 && !s.has_rowvec && s.act == U_NONE && b->ne[0] == 1 && b->ne[1] == 1 && b->ne[2] == cur->ne[2] &&
				b->ne[3] == cur->ne[3] && b->op != GGML_OP_NONE) {
				ensure_planned(b);
				PT r = get(b);
				if (r.st[2] == 1 || r.ne[2] == 1) {
					s.rowvec = r; s.has_rowvec = true; s.rows_per_image = cur->ne[0] * cur->ne[1];
					cur = u; done[u] = true; continue;
				}
			}
		}
		if (u->op == GGML_OP_UNARY && s.act == U_NONE && !s.has_residual) {
			int op = u->op_params[0];
			s.act = op == GGML_UNARY_OP_SILU ? U_SILU : op == GGML_UNARY_OP_GELU ? U_GELU :
				op == GGML_UNARY_OP_GELU_QUICK ? U_GELU_QUICK : op == GGML_UNARY_OP_RELU ? U_RELU : U_TANH;
			cur = u; done[u] = true; continue;
		}
		// residual: add(cur, x) or add(x, cur) with x already planned and laid out like the result
		if (u->op == GGML_OP_ADD && !s.has_residual) {
			const ggml_tensor* other = u->src[0] == cur ? u->src[1] : u->src[0];
			bool same = true;
			for (int i = 0; i < 4; ++i) same = same && other->ne[i] == cur->ne[i];
			if (same && other != cur && (other->op == GGML_OP_NONE || !done.count(other) || val.count(other))) {
				ensure_planned(other);
				PT r = get(other);
				bool layout_ok = r.dt != DT_I32;
				for (int i = 0; i < 4; ++i) if (r.ne[i] != 1 && r.st[i] != out_like.st[i]) layout_ok = false;
				if (layout_ok) {
					s.residual = r; s.has_residual = true;
					cur = u; done[u] = true; continue;
				}
			}
		}
		break;
	}
	return cur;
}

// attention pattern (ggml_extend.c:200-221): kq = mul_mat(k,q); scale_inplace; [diag_mask_inf_inplace];
// soft_max_inplace; out = mul_mat(v, kq)
bool Builder::try_attention(ggml_tensor* t)
{
	ggml_tensor *sc, *nx, *sm, *pv;
	if (!single_user(t, &sc) || sc->op != GGML_OP_SCALE) return false;
	if (!single_user(sc, &nx)) return false;
	bool causal = false;
	if (nx->op == GGML_OP_DIAG_MASK_INF) {
		if (nx->op_params[0] != 0) return false;
		causal = true;
		if (!single_user(nx, &sm)) return false;
	} else sm = nx;
	if (sm->op != GGML_OP_SOFT_MAX) return false;
	if (!single_user(sm, &pv) || pv->op != GGML_OP_MUL_MAT || pv->src[1] != sm) return false;
	const ggml_tensor *k = t->src[0], *q = t->src[1], *v = pv->src[0];
	if (!val.count(v)) return false;    // V must already be planned (it is: built before the first mul_mat)
	float scale; memcpy(&scale, sc->op_params, 4);
	PT pq = get(q), pk = get(k), pvv = get(v);
	// broadcast of k/v over q's batch dims is not used by the reference
	if (k->ne[2] != q->ne[2] || k->ne[3] != q->ne[3]) return false;
	// Wide heads (the VAE's single 512-wide head, vae.c:46-74) do not fit the fused kernel's tensor-memory budget:
	// run them as two tensor-core GEMMs around a register-resident row softmax, per (head, image):
	//   S[nq,nk] (f32) = Q K^T ;  P (f16) = softmax(scale * S) ;  O[nq,d] = P (V^T)^T  with V^T materialised once.
	const int64_t d = q->ne[0], nq = q->ne[1], nk = k->ne[1];
	if (!force_simt && !causal && d > 160 && d % 8 == 0 && nq >= 128 && nk % 8 == 0 && k_softmax_f32_f16_supported(nk)) {
		PT A = as_gemm_rows(pq), Bk = as_gemm_rows(pk);
		PT o; o.dt = DT_F16;
		o.ne[0] = pv->ne[0]; o.ne[1] = pv->ne[1]; o.ne[2] = pv->ne[2]; o.ne[3] = pv->ne[3];
		o.st[0] = 1; o.st[2] = o.ne[0]; o.st[1] = o.ne[0] * o.ne[2]; o.st[3] = o.st[1] * o.ne[1];
		o.buf = new_buf(BUF_ARENA, (size_t)o.numel() * 2);
		// The score block is bounded: queries are processed in chunks of <= 2^26 / nk rows (S chunk <= 256 MB f32 + 128 MB
		// f16 probabilities, reused by every chunk, head and image), so the footprint no longer grows with nq * nk
		// (16384 tokens at 1024x1024: 1.6 GB before, 0.4 GB now; the untiled 2048x2048 decode: 25 GB -> 0.4 GB).
		int64_t qc = std::max<int64_t>(128, ((int64_t)1 << 26) / nk / 128 * 128);
		if (qc > nq) qc = nq;
		int64_t ne_s[4] = { nk, qc, 1, 1 };
		PT S = new_pt(DT_F32, ne_s), Pm = new_pt(DT_F16, ne_s);
		int64_t ne_vt[4] = { nk, d, 1, 1 };
		PT Vt = new_pt(DT_F16, ne_vt);
		for (int64_t b = 0; b < q->ne[3]; ++b) for (int64_t h = 0; h < q->ne[2]; ++h) {
			PT qa = A, ka = Bk, va = pvv, oa = o;
			qa.off += h * A.st[2] + b * A.st[3]; ka.off += h * Bk.st[2] + b * Bk.st[3];
			va.off += h * pvv.st[2] + b * pvv.st[3]; oa.off += h * o.st[2] + b * o.st[3];
			for (PT* x : { &qa, &ka, &va, &oa }) { x->ne[2] = x->ne[3] = 1; }
			copy(Vt, va, "attn_v_transpose");          // [nk, d] view of token-major V -> rows of nk keys per channel
			for (int64_t r0 = 0; r0 < nq; r0 += qc) {
				const int64_t rows = std::min(qc, nq - r0);
				PT qr = qa, orr = oa, Sr = S, Pr = Pm;
				qr.off += r0 * A.st[1]; qr.ne[1] = rows;
				orr.off += r0 * o.st[1]; orr.ne[1] = rows;
				Sr.ne[1] = rows; Pr.ne[1] = rows;
				Step g1; g1.kind = S_GEMM_TC; g1.name = "attn_qk";
				g1.in[0] = qr; g1.in[1] = ka; g1.n_in = 2; g1.out = Sr;
				g1.M = rows; g1.N = nk; g1.K = d; g1.lda = A.st[1]; g1.ldb = Bk.st[1]; g1.ldc = nk;
				P->steps.push_back(g1);
				Step& sm2 = emit(S_SOFTMAX_F16, "attn_softmax");
				sm2.out = Pr; sm2.in[0] = Sr; sm2.n_in = 1; sm2.fparam = scale;
				Step g2; g2.kind = S_GEMM_TC; g2.name = "attn_pv";
				g2.in[0] = Pr; g2.in[1] = Vt; g2.n_in = 2; g2.out = orr;
				g2.M = rows; g2.N = d; g2.K = nk; g2.lda = nk; g2.ldb = nk; g2.ldc = o.st[1];
				P->steps.push_back(g2);
			}
		}
		done[t] = done[sc] = done[sm] = true;
		finish(pv, o);
		return true;
	}
	// output [d, nq, H, B]; choose token-major, heads interleaved: (d:1, H:d, nq:d*H, B:d*H*nq) so that
	// the reference's head-merge permute+cont+reshape becomes a free view
	PT o; o.dt = DT_F16;
	o.ne[0] = pv->ne[0]; o.ne[1] = pv->ne[1]; o.ne[2] = pv->ne[2]; o.ne[3] = pv->ne[3];
	o.st[0] = 1; o.st[2] = o.ne[0]; o.st[1] = o.ne[0] * o.ne[2]; o.st[3] = o.st[1] * o.ne[1];
	o.buf = new_buf(BUF_ARENA, (size_t)o.numel() * 2);
	Step& s = emit(S_ATTENTION, "attention");
	s.out = o; s.in[0] = pq; s.in[1] = pk; s.in[2] = pvv; s.n_in = 3;
	s.fparam = scale; s.iparam[0] = causal ? 1 : 0;
	done[t] = done[sc] = done[sm] = true;
	if (causal) done[nx] = true;
	finish(pv, o);
	return true;
}

void Builder::plan_mul_mat(ggml_tensor* t)
{
	if (try_attention(t)) return;
	const ggml_tensor *a = t->src[0], *b = t->src[1];
	PT pa = get(a), pb = get(b);
	const int64_t K = a->ne[0], Mw = a->ne[1];
	bool weight2d = a->ne[2] == 1 && a->ne[3] == 1;
	bool tc_ok = !force_simt && weight2d && pa.dt == DT_F16 && pa.st[0] == 1 && gemm_tc_supported(b->ne[1] * b->ne[2] * b->ne[3], Mw, K) &&
		pa.st[1] % 8 == 0 && K % 8 == 0;
	// Sibling projections of the same activation without bias / activation / residual (q, k, v of a self-attention,
	// mlblock_nn.c:199-203; the k and v projections of the text context in every cross-attention, unet.c:110-145) run
	// as ONE GEMM over the concatenated weight rows: the activation is streamed once instead of once per projection.
	// Biased sibling projections of the same (activated) vector: the time-embedding projection of every resnet,
	// emb_layers = linear(silu(emb)) (mlblock_nn.c:139-144), 22 launches with M = images in SD1.x. Each resnet builds its own
	// silu node, so siblings are matched through the unary's source. One GEMM over the concatenated weight rows and biases;
	// the members read column slices of its output (the conv epilogue's per-image vector takes any row pitch).
	if (tc_ok && !env_flag("GGML_B200_NO_PROJ_FUSION") && !env_flag("GGML_B200_NO_EMB_FUSION") && b->ne[1] * b->ne[2] * b->ne[3] <= 256) {
		const bool b_unary = b->op == GGML_OP_UNARY;
		const ggml_tensor* bsrc = b_unary ? b->src[0] : b;
		auto same_input = [&](const ggml_tensor* x) {
			if (x == b) return true;
			if (!b_unary || x->op != GGML_OP_UNARY || x->src[0] != bsrc || x->op_params[0] != b->op_params[0]) return false;
			for (int i = 0; i < 4; ++i) if (x->ne[i] != b->ne[i]) return false;
			return true;
		};
		auto biased = [&](ggml_tensor* u, ggml_tensor** addp) {
			const ggml_tensor* w = u->src[0];
			if (u->op != GGML_OP_MUL_MAT || !same_input(u->src[1]) || done.count(u) || (u->flags & GGML_TENSOR_FLAG_OUTPUT)) return false;
			if (w->op != GGML_OP_NONE || w->type != GGML_TYPE_F16 || w->ne[0] != K || w->ne[2] != 1 || w->ne[3] != 1 || (w->nb[1] / 2) % 8 || w->ne[1] % 8) return false;
			ggml_tensor* nx;
			if (!single_user(u, &nx) || nx->op != GGML_OP_ADD || nx->src[0] != u || (nx->flags & GGML_TENSOR_FLAG_OUTPUT)) return false;
			if (!is_channel_vec(nx->src[1], u, 0)) return false;
			auto it = users.find(nx);
			if (it == users.end() || it->second.empty()) return false;
			for (ggml_tensor* q : it->second) if (q->op != GGML_OP_RESHAPE) return false;      // no activation / residual to absorb: the sum is used as a vector
			*addp = nx;
			return true;
		};
		std::vector<ggml_tensor*> grp, adds;
		auto scan = [&](const ggml_tensor* in) {
			auto it = users.find(in);
			if (it == users.end()) return;
			for (ggml_tensor* u : it->second) { ggml_tensor* ad; if (std::find(grp.begin(), grp.end(), u) == grp.end() && biased(u, &ad)) { grp.push_back(u); adds.push_back(ad); } }
		};
		if (b_unary) { auto it = users.find(bsrc); if (it != users.end()) for (ggml_tensor* q : it->second) if (same_input(q)) scan(q); }
		else scan(b);
		if (grp.size() >= 2 && std::find(grp.begin(), grp.end(), t) != grp.end()) {
			int64_t n_total = 0;
			for (ggml_tensor* u : grp) n_total += u->src[0]->ne[1];
			int64_t ne_w[4] = { K, n_total, 1, 1 }, ne_b[4] = { n_total, 1, 1, 1 };
			PT wcat = new_pt(DT_F16, ne_w, BUF_PERSIST), bcat = new_pt(DT_F32, ne_b, BUF_PERSIST);
			PT A = as_gemm_rows(pb);
			int64_t rows, pitch; rows_uniform(A, rows, pitch);
			PT out; out.dt = DT_F16;
			out.ne[0] = n_total; out.ne[1] = b->ne[1]; out.ne[2] = b->ne[2]; out.ne[3] = b->ne[3];
			contiguous_strides(out);
			out.buf = new_buf(BUF_ARENA, (size_t)out.numel() * 2);
			int64_t col = 0;
			for (size_t i = 0; i < grp.size(); ++i) {
				const ggml_tensor* w = grp[i]->src[0];
				const ggml_tensor* bl = strip_reshape(adds[i]->src[1]);
				PT sub = wcat; sub.ne[1] = w->ne[1]; sub.off = col * K;
				Step ps; ps.kind = S_COPY; ps.name = "proj_weight_concat"; ps.out = sub; ps.in[0] = get(w); ps.n_in = 1; ps.leaf = w;
				P->prep.push_back(ps);
				PT bsub = bcat; bsub.ne[0] = w->ne[1]; bsub.off = col;
				PT bsrc_pt = get(bl); bsrc_pt.ne[0] = w->ne[1]; bsrc_pt.ne[1] = bsrc_pt.ne[2] = bsrc_pt.ne[3] = 1; bsrc_pt.st[0] = 1;
				Step pb2; pb2.kind = S_COPY; pb2.name = "proj_bias_concat"; pb2.out = bsub; pb2.in[0] = bsrc_pt; pb2.n_in = 1; pb2.leaf = bl;
				P->prep.push_back(pb2);
				col += w->ne[1];
			}
			Step s; s.kind = S_GEMM_TC; s.name = "linear_fused";
			s.in[0] = A; s.in[1] = wcat; s.n_in = 2; s.out = out;
			s.M = rows; s.N = n_total; s.K = K; s.lda = pitch; s.ldb = K; s.ldc = n_total;
			s.bias = bcat; s.has_bias = true;
			P->steps.push_back(s);
			col = 0;
			for (size_t i = 0; i < grp.size(); ++i) {
				PT v = out; v.ne[0] = grp[i]->src[0]->ne[1]; v.off = col;
				col += v.ne[0];
				done[grp[i]] = true;
				finish(adds[i], v);
			}
			return;
		}
	}
	if (tc_ok && !env_flag("GGML_B200_NO_PROJ_FUSION")) {
		auto plain = [&](const ggml_tensor* u) {
			const ggml_tensor* w = u->src[0];
			if (u->op != GGML_OP_MUL_MAT || u->src[1] != b || done.count(u) || (u->flags & GGML_TENSOR_FLAG_OUTPUT)) return false;
			if (w->op != GGML_OP_NONE || w->type != GGML_TYPE_F16 || w->ne[0] != K || w->ne[2] != 1 || w->ne[3] != 1 || (w->nb[1] / 2) % 8) return false;
			auto it = users.find(u);
			if (it == users.end() || it->second.empty()) return false;
			for (ggml_tensor* nx : it->second)
				if (nx->op == GGML_OP_ADD || nx->op == GGML_OP_UNARY || nx->op == GGML_OP_MUL_MAT || nx->op == GGML_OP_SCALE) return false;   // has an epilogue / is an attention operand as-is
			return true;
		};
		std::vector<ggml_tensor*> grp;
		for (ggml_tensor* u : users[b]) if (plain(u) && std::find(grp.begin(), grp.end(), u) == grp.end()) grp.push_back(u);
		if (grp.size() >= 2 && std::find(grp.begin(), grp.end(), t) != grp.end()) {
			int64_t n_total = 0;
			for (ggml_tensor* u : grp) n_total += u->src[0]->ne[1];
			int64_t ne_w[4] = { K, n_total, 1, 1 };
			PT wcat = new_pt(DT_F16, ne_w, BUF_PERSIST);
			PT A = as_gemm_rows(pb);
			int64_t rows, pitch; rows_uniform(A, rows, pitch);
			PT out; out.dt = DT_F16;
			out.ne[0] = n_total; out.ne[1] = b->ne[1]; out.ne[2] = b->ne[2]; out.ne[3] = b->ne[3];
			contiguous_strides(out);
			out.buf = new_buf(BUF_ARENA, (size_t)out.numel() * 2);
			int64_t col = 0;
			for (ggml_tensor* u : grp) {
				const ggml_tensor* w = u->src[0];
				PT sub = wcat; sub.ne[1] = w->ne[1]; sub.off = col * K;
				Step ps; ps.kind = S_COPY; ps.name = "proj_weight_concat"; ps.out = sub; ps.in[0] = get(w); ps.n_in = 1; ps.leaf = w;
				P->prep.push_back(ps);
				col += w->ne[1];
			}
			Step s; s.kind = S_GEMM_TC; s.name = "linear_fused";
			s.in[0] = A; s.in[1] = wcat; s.n_in = 2; s.out = out;
			s.M = rows; s.N = n_total; s.K = K; s.lda = pitch; s.ldb = K; s.ldc = n_total;
			P->steps.push_back(s);
			col = 0;
			for (ggml_tensor* u : grp) {
				PT v = out; v.ne[0] = u->src[0]->ne[1]; v.off = col;
				col += v.ne[0];
				done[u] = true;
				finish(u, v);
			}
			return;
		}
	}
	if (tc_ok) {
		PT A = as_gemm_rows(pb);
		int64_t rows, pitch; rows_uniform(A, rows, pitch);
		PT out; out.dt = DT_F16;
		out.ne[0] = Mw; out.ne[1] = b->ne[1]; out.ne[2] = b->ne[2]; out.ne[3] = b->ne[3];
		contiguous_strides(out);
		out.buf = new_buf(BUF_ARENA, (size_t)out.numel() * 2);
		Step s; s.kind = S_GEMM_TC; s.name = "linear";
		s.in[0] = A; s.in[1] = pa; s.n_in = 2;
		s.M = rows; s.N = Mw; s.K = K; s.lda = pitch; s.ldb = pa.st[1]; s.ldc = Mw;
		ggml_tensor* last = absorb_epilogue(t, s, 0, out);
		// GEGLU projection (mlblock_nn.c:159-172): gate in the GEMM epilogue, half the columns ever reach memory
		ggml_tensor *xv, *gv, *cg, *ge, *mu;
		if (!env_flag("GGML_B200_NO_GEGLU_FUSION") && s.act == U_NONE && !s.has_residual && !s.has_rowvec && Mw % 64 == 0 &&
			!(last->flags & GGML_TENSOR_FLAG_OUTPUT) && match_geglu(last, &xv, &gv, &cg, &ge, &mu)) {
			const int64_t D = Mw / 2;
			auto prep = [&](const ggml_tensor* leaf, const PT& src, DT dt, int64_t row_elems) {
				auto it = prepared.find(leaf);
				if (it != prepared.end()) return it->second;
				int64_t ne[4] = { row_elems, Mw, 1, 1 };
				PT wp = new_pt(dt, ne, BUF_PERSIST);
				Step ps; ps.kind = S_WPREP_GEGLU; ps.name = "geglu_weight_prep"; ps.out = wp; ps.in[0] = src; ps.n_in = 1;
				ps.iparam[0] = (int)D; ps.leaf = leaf;
				P->prep.push_back(ps);
				prepared[leaf] = wp;
				return wp;
			};
			s.in[1] = prep(a, pa, DT_F16, K); s.ldb = K;
			if (s.has_bias) {
				const ggml_tensor* bl = nullptr;
				for (ggml_tensor* u = t; u != last; ) { ggml_tensor* nx; single_user(u, &nx); if (nx->op == GGML_OP_ADD) bl = strip_reshape(nx->src[1]); u = nx; }
				if (bl && bl->op == GGML_OP_NONE && s.bias.dt == DT_F32) s.bias = prep(bl, s.bias, DT_F32, 1);
				else bl = nullptr;
				if (!bl) B200_FATAL("GEGLU fusion: bias of '%s' is not a plain f32 parameter", t->name);
			}
			PT og; og.dt = DT_F16;
			og.ne[0] = D; og.ne[1] = b->ne[1]; og.ne[2] = b->ne[2]; og.ne[3] = b->ne[3];
			contiguous_strides(og);
			og.buf = new_buf(BUF_ARENA, (size_t)og.numel() * 2);
			s.out = og; s.ldc = D; s.iparam[0] = 1; s.name = "linear_geglu";
			P->steps.push_back(s);
			done[t] = done[last] = done[xv] = done[gv] = done[cg] = done[ge] = true;
			finish(mu, og);
			return;
		}
		// the [d, tokens] result of a linear may carry the residual in token-major layout too
		s.out = out;
		P->steps.push_back(s);
		done[t] = true;
		finish(last, out);
		return;
	}
	// generic path: any strides / dtypes, batched, exact f32 accumulation
	// exact path keeps f32 results when the reference would (f32 x f32, or a graph output)
	DT odt = ((t->flags & GGML_TENSOR_FLAG_OUTPUT) || (pa.dt == DT_F32 && pb.dt == DT_F32)) ? DT_F32 : DT_F16;
	PT out = new_pt(odt, t->ne);
	Step& s = emit(S_GEMM_SIMT, "mul_mat_simt");
	s.out = out; s.in[0] = pa; s.in[1] = pb; s.n_in = 2;
	s.iparam[0] = (pa.dt == DT_F16 && pb.dt == DT_F32) ? 1 : 0;   // round activations like the reference
	finish(t, out);
}

void Builder::plan_conv(ggml_tensor* t)
{
	const ggml_tensor *w = t->src[0], *x = t->src[1];
	const int32_t* op = t->op_params;
	int s0 = op[0], s1 = op[1], p0 = op[2], p1 = op[3], d0 = op[4], d1 = op[5];
	int64_t KW = w->ne[0], KH = w->ne[1], Cin = w->ne[2], Cout = w->ne[3];
	int64_t W = x->ne[0], H = x->ne[1], N = x->ne[3], OW = t->ne[0], OH = t->ne[1];
	if (w->op != GGML_OP_NONE) B200_FATAL("conv_2d: kernel must be a parameter leaf");
	PT pw = get(w);
	PT px = as_nhwc_f16(get(x));
	PT out = new_pt_nhwc(DT_F16, OW, OH, Cout, N);
	Step s; s.name = "conv";
	const bool k1 = KW == 1 && KH == 1 && s0 == 1 && s1 == 1 && p0 == 0 && p1 == 0;
	const bool k3 = KW == 3 && KH == 3 && s0 == 1 && s1 == 1 && p0 == 1 && p1 == 1 && d0 == 1 && d1 == 1;
	if (!force_simt && k1 && Cin % 8 == 0 && pw.dt == DT_F16) {
		s.kind = S_GEMM_TC;
		s.in[0] = px; s.in[1] = pw; s.n_in = 2;
		s.M = W * H * N; s.N = Cout; s.K = Cin; s.lda = Cin; s.ldb = Cin; s.ldc = Cout;
	} else if (!force_simt && k3 && Cin % 64 == 0) {
		// weights re-laid-out once to [Cout][kh][kw][Cin]
		PT wp;
		auto it = prepared.find(w);
		if (it != prepared.end()) wp = it->second;
		else {
			int64_t ne[4] = { 9 * Cin, Cout, 1, 1 };
			wp = new_pt(DT_F16, ne, BUF_PERSIST);
			Step ps; ps.kind = S_WPREP_CONV; ps.name = "conv_weight_prep"; ps.out = wp; ps.in[0] = pw; ps.n_in = 1;
			ps.iparam[0] = (int)(9 * Cin); ps.leaf = w;
			P->prep.push_back(ps);
			prepared[w] = wp;
		}
		s.kind = S_CONV_TC;
		s.in[0] = px; s.in[1] = wp; s.n_in = 2;
		s.conv_n = N; s.conv_h = H; s.conv_w = W; s.conv_c = Cin;
		s.M = W * H * N; s.N = Cout; s.K = 9 * Cin; s.ldc = Cout;
	} else {
		// strided / narrow / odd convolutions: explicit im2col (K padded to 64) + GEMM
		int64_t Kc = KW * KH * Cin, kpad = (Kc + 63) / 64 * 64;
		PT wp;
		auto it = prepared.find(w);
		if (it != prepared.end()) wp = it->second;
		else {
			int64_t ne[4] = { kpad, Cout, 1, 1 };
			wp = new_pt(DT_F16, ne, BUF_PERSIST);
			Step ps; ps.kind = S_WPREP_CONV; ps.name = "conv_weight_prep"; ps.out = wp; ps.in[0] = pw; ps.n_in = 1;
			ps.iparam[0] = (int)kpad; ps.leaf = w;
			P->prep.push_back(ps);
			prepared[w] = wp;
		}
		int64_t ne[4] = { kpad, OW * OH * N, 1, 1 };
		PT col = new_pt(DT_F16, ne);
		Step& c = emit(S_IM2COL, "im2col");
		c.out = col; c.in[0] = px; c.n_in = 1;
		c.iparam[0] = (int)KW; c.iparam[1] = (int)KH; c.iparam[2] = s0; c.iparam[3] = s1; c.iparam[4] = p0; c.iparam[5] = p1;
		c.iparam[6] = d0; c.iparam[7] = d1;
		c.M = OW; c.N = OH; c.K = kpad;
		if (force_simt) {
			PT o2 = out;
			Step& gs = emit(S_GEMM_SIMT, "conv_simt");
			// view the NHWC output as [Cout, rows]
			PT ov; ov.buf = o2.buf; ov.dt = o2.dt; ov.ne[0] = Cout; ov.ne[1] = OW * OH * N; ov.st[0] = 1; ov.st[1] = Cout; ov.st[2] = ov.st[3] = Cout * ov.ne[1];
			gs.out = ov; gs.in[0] = wp; gs.in[1] = col; gs.n_in = 2;
			finish(t, out);
			return;
		}
		s.kind = S_GEMM_TC;
		s.in[0] = col; s.in[1] = wp; s.n_in = 2;
		s.M = OW * OH * N; s.N = Cout; s.K = kpad; s.lda = kpad; s.ldb = kpad; s.ldc = Cout;
	}
	ggml_tensor* last = absorb_epilogue(t, s, 2, out);
	s.out = out;
	P->steps.push_back(s);
	done[t] = true;
	finish(last, out);
}

void Builder::plan_node(ggml_tensor* t)
{
	switch (t->op) {
	case GGML_OP_RESHAPE: {
		PT a = get(t->src[0]), o;
		if (!try_reshape(a, t->ne, o)) {
			PT c = to_contig(a, a.dt);
			if (!try_reshape(c, t->ne, o)) B200_FATAL("reshape of a contiguous tensor failed");
		}
		finish(t, o);
	} break;
	case GGML_OP_PERMUTE: {
		PT a = get(t->src[0]), o = a;
		for (int i = 0; i < 4; ++i) { int ax = t->op_params[i]; o.ne[ax] = a.ne[i]; o.st[ax] = a.st[i]; }
		finish(t, o);
	} break;
	case GGML_OP_TRANSPOSE: {
		PT a = get(t->src[0]), o = a;
		std::swap(o.ne[0], o.ne[1]); std::swap(o.st[0], o.st[1]);
		finish(t, o);
	} break;
	case GGML_OP_VIEW: {
		// byte strides/offset are expressed in the source's logical ggml layout
		const ggml_tensor* src = t->src[0];
		PT a = get(src);
		bool logical = true;
		size_t es = ggml_type_size(src->type);
		for (int i = 0; i < 4; ++i) if (a.ne[i] != 1 && a.st[i] != (int64_t)(src->nb[i] / es)) logical = false;
		if (!logical) {
			// bring the source into its logical layout first
			PT c; c.dt = a.dt;
			for (int i = 0; i < 4; ++i) { c.ne[i] = a.ne[i]; c.st[i] = (int64_t)(src->nb[i] / es); }
			int64_t span = 1; for (int i = 0; i < 4; ++i) span += (c.ne[i] - 1) * c.st[i];
			c.buf = new_buf(BUF_ARENA, (size_t)span * dt_size(c.dt));
			copy(c, a, "to_logical");
			a = c;
		}
		size_t offset; memcpy(&offset, t->op_params, sizeof(offset));
		PT o = a;
		for (int i = 0; i < 4; ++i) { o.ne[i] = t->ne[i]; o.st[i] = (int64_t)(t->nb[i] / es); }
		o.off = a.off + (int64_t)(offset / es);
		finish(t, o);
	} break;
	case GGML_OP_CONT: {
		PT a = get(t->src[0]);
		// Every consumer works on strided views (and materialises one itself when it really needs contiguity), so an
		// f16 view is passed through as it is: a permutation of a compact buffer, or a column slice of a fused
		// projection (head split of q/k/v, mlblock_nn.c:204-222).
		if (is_dense(a) || a.dt == DT_F16) finish(t, a);
		else finish(t, to_contig(a, a.dt == DT_F32 ? DT_F16 : a.dt));
	} break;
	case GGML_OP_ADD: case GGML_OP_MUL: {
		const ggml_tensor *a = t->src[0], *b = t->src[1];
		PT pa = get(a), pb = get(b);
		DT odt = (pa.dt == DT_F32 && (t->flags & GGML_TENSOR_FLAG_OUTPUT) == 0) ? DT_F16 : pa.dt;
		if (t->type == GGML_TYPE_F16) odt = DT_F16;
		PT o = new_pt_like(odt, t->ne, pa);
		Step& s = emit(S_BINARY, t->op == GGML_OP_ADD ? "add" : "mul");
		s.out = o; s.in[0] = pa; s.in[1] = pb; s.n_in = 2; s.iop = t->op == GGML_OP_ADD ? B_ADD : B_MUL;
		finish(t, o);
	} break;
	case GGML_OP_SCALE: case GGML_OP_UNARY: {
		PT a = get(t->src[0]);
		int iop; float fp = 0;
		if (t->op == GGML_OP_SCALE) { iop = U_SCALE; memcpy(&fp, t->op_params, 4); }
		else {
			int op = t->op_params[0];
			iop = op == GGML_UNARY_OP_SILU ? U_SILU : op == GGML_UNARY_OP_GELU ? U_GELU :
				op == GGML_UNARY_OP_GELU_QUICK ? U_GELU_QUICK : op == GGML_UNARY_OP_RELU ? U_RELU : U_TANH;
		}
		// the same activation of the same value (silu(emb) in every resnet, mlblock_nn.c:140) is computed once
		std::string key = pt_key("unary", a, iop, fp);
		auto it = memo.find(key);
		if (it != memo.end()) { finish(t, it->second); break; }
		PT o = new_pt_like(a.dt == DT_F32 ? DT_F16 : a.dt, t->ne, a);
		Step& s = emit(S_UNARY, t->op == GGML_OP_SCALE ? "scale" : "unary");
		s.out = o; s.in[0] = a; s.n_in = 1; s.iop = iop; s.fparam = fp;
		memo[key] = o;
		finish(t, o);
	} break;
	case GGML_OP_NORM: {
		// norm -> mul(w) -> add(b)  (mlblock_nn.c:58-75)
		PT a = get(t->src[0]);
		float eps; memcpy(&eps, t->op_params, 4);
		ggml_tensor *m = nullptr, *ad = nullptr, *last = t;
		PT gw, gb; bool has_w = false, has_b = false;
		if (single_user(t, &m) && m->op == GGML_OP_MUL && m->src[0] == t && is_channel_vec(m->src[1], t, 0)) {
			gw = get(strip_reshape(m->src[1])); has_w = true; last = m; done[m] = true;
			if (single_user(m, &ad) && ad->op == GGML_OP_ADD && ad->src[0] == m && is_channel_vec(ad->src[1], m, 0)) {
				gb = get(strip_reshape(ad->src[1])); has_b = true; last = ad; done[ad] = true;
			}
		}
		int64_t rows, pitch;
		PT src = a;
		if (!rows_uniform(src, rows, pitch)) src = to_contig(a, a.dt);
		PT o = new_pt(DT_F16, t->ne);
		Step& s = emit(S_LAYERNORM, "layernorm");
		s.out = o; s.in[0] = src; s.n_in = 1; s.fparam = eps;
		if (has_w) { s.bias = gw; s.has_bias = true; }
		if (has_b) { s.rowvec = gb; s.has_rowvec = true; }
		done[t] = true;
		finish(last, o);
	} break;
	case GGML_OP_GROUP_NORM: {
		// group_norm -> mul(w) -> add(b) -> [silu]  (mlblock_nn.c:78-103, :136,147)
		PT a = as_nhwc_f16(get(t->src[0]));
		int groups = t->op_params[0];
		float eps; memcpy(&eps, &t->op_params[1], 4);
		ggml_tensor *m = nullptr, *ad = nullptr, *un = nullptr, *last = t;
		PT gw, gb; bool has_w = false, has_b = false, silu = false;
		if (single_user(t, &m) && m->op == GGML_OP_MUL && m->src[0] == t && is_channel_vec(m->src[1], t, 2)) {
			gw = get(strip_reshape(m->src[1])); has_w = true; last = m; done[m] = true;
			if (single_user(m, &ad) && ad->op == GGML_OP_ADD && ad->src[0] == m && is_channel_vec(ad->src[1], m, 2)) {
				gb = get(strip_reshape(ad->src[1])); has_b = true; last = ad; done[ad] = true;
				if (single_user(ad, &un) && un->op == GGML_OP_UNARY && un->op_params[0] == GGML_UNARY_OP_SILU) {
					silu = true; last = un; done[un] = true;
				}
			}
		}
		if (t->ne[2] % 8) B200_FATAL("group_norm: channel count %lld must be a multiple of 8", (long long)t->ne[2]);
		PT o = new_pt_nhwc(DT_F16, t->ne[0], t->ne[1], t->ne[2], t->ne[3]);
		// the statistics can come from the epilogue of the GEMM / conv that wrote exactly this tensor (last writer of its buffer)
		int producer = -1;
		if (!env_flag("GGML_B200_NO_GN_EPILOGUE")) {
			for (int i = (int)P->steps.size() - 1; i >= 0; --i) {
				const Step& ps = P->steps[i];
				if (ps.out.buf != a.buf) continue;
				// (a launch serves ONE group_norm: a second consumer of the same tensor keeps its own statistics pass and slot -- if
				// the producer turns out not to fuse at run time, two passes adding into one slot would count every element twice)
				bool same = (ps.kind == S_GEMM_TC || ps.kind == S_CONV_TC) && ps.out.dt == DT_F16 && ps.out.off == a.off && ps.out.numel() == a.numel() &&
					!(ps.kind == S_GEMM_TC && (ps.iparam[0] == 1 || ps.ldc != ps.N)) && ps.N == a.ne[2] && ps.gn_groups == 0;
				if (same) producer = i;
				break;
			}
		}
		Step& s = emit(S_GROUPNORM, "groupnorm");
		s.out = o; s.in[0] = a; s.n_in = 1; s.fparam = eps; s.iparam[0] = groups; s.iparam[1] = silu ? 1 : 0;
		if (has_w) { s.bias = gw; s.has_bias = true; }
		if (has_b) { s.rowvec = gb; s.has_rowvec = true; }
		s.stats_off = P->zero_bytes;
		P->zero_bytes += (size_t)4 * groups * t->ne[3] * sizeof(unsigned long long);     // fixed-point (hi, lo) sum and sum of squares
		if (producer >= 0) { Step& ps = P->steps[producer]; ps.gn_groups = groups; ps.stats_off = s.stats_off; ps.gn_rows_per_image = a.ne[0] * a.ne[1]; }
		s.gn_producer = producer;
		done[t] = true;
		finish(last, o);
	} break;
	case GGML_OP_MUL_MAT: plan_mul_mat(t); break;
	case GGML_OP_CONV_2D: plan_conv(t); break;
	case GGML_OP_SOFT_MAX: case GGML_OP_DIAG_MASK_INF: {
		PT a = to_contig(get(t->src[0]), DT_F32);
		PT o = new_pt(DT_F32, t->ne);
		if (t->op == GGML_OP_DIAG_MASK_INF) {
			// standalone mask: fold into a causal softmax if that is the only user, else unsupported
			ggml_tensor* u;
			if (single_user(t, &u) && u->op == GGML_OP_SOFT_MAX) {
				Step& s = emit(S_SOFTMAX, "softmax_causal");
				s.out = o; s.in[0] = a; s.n_in = 1; s.iparam[0] = 1; s.iparam[1] = t->op_params[0];
				done[t] = true;
				finish(u, o);
				break;
			}
			B200_FATAL("diag_mask_inf without a following soft_max is not supported");
		}
		Step& s = emit(S_SOFTMAX, "softmax");
		s.out = o; s.in[0] = a; s.n_in = 1;
		finish(t, o);
	} break;
	case GGML_OP_CONCAT: {
		PT a = get(t->src[0]), b = get(t->src[1]);
		int dim = t->op_params[0];
		PT o = new_pt_like(a.dt == DT_F32 ? DT_F16 : a.dt, t->ne, a);
		PT oa = o, ob = o;
		for (int i = 0; i < 4; ++i) { oa.ne[i] = a.ne[i]; ob.ne[i] = b.ne[i]; }
		ob.off = o.off + a.ne[dim] * o.st[dim];
		copy(oa, a, "concat_a");
		copy(ob, b, "concat_b");
		finish(t, o);
	} break;
	case GGML_OP_PAD: {
		PT a = get(t->src[0]);
		PT o = new_pt_like(a.dt == DT_F32 ? DT_F16 : a.dt, t->ne, a);
		Step& z = emit(S_ZERO, "pad_zero"); z.out = o;
		PT oa = o;
		for (int i = 0; i < 4; ++i) oa.ne[i] = a.ne[i];
		copy(oa, a, "pad_copy");
		finish(t, o);
	} break;
	case GGML_OP_UPSCALE: {
		PT a = get(t->src[0]);
		PT o = new_pt_like(a.dt == DT_F32 ? DT_F16 : a.dt, t->ne, a);
		Step& s = emit(S_UPSCALE, "upscale");
		s.out = o; s.in[0] = a; s.n_in = 1;
		finish(t, o);
	} break;
	case GGML_OP_GET_ROWS: {
		PT tab = get(t->src[0]), ids = get(t->src[1]);
		PT o = new_pt(DT_F16, t->ne);
		Step& s = emit(S_GET_ROWS, "get_rows");
		s.out = o; s.in[0] = tab; s.in[1] = ids; s.n_in = 2;
		finish(t, o);
	} break;
	case GGML_OP_TIMESTEP_EMBEDDING: {
		PT ts = get(t->src[0]);
		PT o = new_pt(DT_F16, t->ne);
		Step& s = emit(S_TSEMB, "timestep_embedding");
		s.out = o; s.in[0] = ts; s.n_in = 1; s.iparam[0] = t->op_params[0]; s.iparam[1] = t->op_params[1];
		finish(t, o);
	} break;
	default:
		B200_FATAL("op %s is not supported by the B200 engine", ggml_op_name(t->op));
	}
}

// GEGLU gate (mlblock_nn.c:164-169): x,g = chunk(h); mul(x, gelu(cont(g))).
bool Builder::match_geglu(const ggml_tensor* h, ggml_tensor** pxv, ggml_tensor** pgv, ggml_tensor** pc, ggml_tensor** pge, ggml_tensor** pmu)
{
	auto it = users.find(h);
	if (it == users.end()) return false;
	auto& us = it->second;
	if (!(us.size() == 2 && us[0]->op == GGML_OP_VIEW && us[1]->op == GGML_OP_VIEW &&
		!(h->flags & GGML_TENSOR_FLAG_OUTPUT) && h->ne[0] % 16 == 0)) return false;
	ggml_tensor *xv = us[0], *gv = us[1], *c, *ge, *mu;
	size_t off0, off1; memcpy(&off0, xv->op_params, sizeof(off0)); memcpy(&off1, gv->op_params, sizeof(off1));
	int64_t d = h->ne[0] / 2;
	if (!(off0 == 0 && off1 == (size_t)d * ggml_type_size(h->type) && xv->ne[0] == d && gv->ne[0] == d &&
		single_user(gv, &c) && c->op == GGML_OP_CONT && single_user(c, &ge) && ge->op == GGML_OP_UNARY &&
		ge->op_params[0] == GGML_UNARY_OP_GELU && single_user(ge, &mu) && mu->op == GGML_OP_MUL &&
		mu->src[0] == xv && mu->src[1] == ge && users[xv].size() == 1)) return false;
	*pxv = xv; *pgv = gv; *pc = c; *pge = ge; *pmu = mu;
	return true;
}

// stand-alone gate kernel (when the projection did not absorb it)
bool Builder::try_geglu(ggml_tensor* t)
{
	if (t->op != GGML_OP_VIEW) return false;
	const ggml_tensor* h = t->src[0];
	ggml_tensor *xv, *gv, *c, *ge, *mu;
	if (!match_geglu(h, &xv, &gv, &c, &ge, &mu) || t != xv) return false;
	PT ph = get(h);
	if (!(ph.dt == DT_F16 && is_ggml_contig(ph))) return false;
	PT o = new_pt(DT_F16, mu->ne);
	Step& s = emit(S_GEGLU, "geglu");
	s.out = o; s.in[0] = ph; s.n_in = 1;
	done[xv] = done[gv] = done[c] = done[ge] = true;
	finish(mu, o);
	return true;
}

void Builder::plan_one(ggml_tensor* t)
{
	if (done.count(t)) return;
	if (try_geglu(t)) return;
	plan_node(t);
}

// Plan the not-yet-planned ancestors of t (and t). Used when a fusion wants an operand that the
// node order schedules later (e.g. the time-embedding projection or the skip convolution of a
// resnet, which the reference builds after the convolution they are added to).
void Builder::ensure_planned(const ggml_tensor* t)
{
	std::function<void(const ggml_tensor*)> rec = [&](const ggml_tensor* x) {
		if (x->op == GGML_OP_NONE || val.count(x) || done.count(x)) return;
		for (int i = 0; i < GGML_MAX_SRC; ++i) if (x->src[i]) rec(x->src[i]);
		plan_one(const_cast<ggml_tensor*>(x));
	};
	rec(t);
	if (t->op != GGML_OP_NONE && !val.count(t))
		B200_FATAL("planner: operand '%s' (%s) was fused away and cannot be used as an operand", t->name, ggml_op_name(t->op));
}

void Builder::plan()
{
	for (ggml_tensor* t : g->nodes)
		for (int i = 0; i < GGML_MAX_SRC; ++i)
			if (t->src[i]) users[t->src[i]].push_back(t);
	P->zero_bytes = 0;
	for (ggml_tensor* t : g->nodes) plan_one(t);
}

// ------------------------------------------------------------------ memory assignment
static void touch(Plan* P, const PT& p, int step)
{
	if (p.buf < 0) return;
	Buf& b = P->bufs[p.buf];
	if (b.first < 0) b.first = step;
	b.last = std::max(b.last, step);
}
static void step_touch(Plan* P, Step& s, int idx)
{
	touch(P, s.out, idx);
	for (int i = 0; i < s.n_in; ++i) touch(P, s.in[i], idx);
	if (s.has_bias) touch(P, s.bias, idx);
	if (s.has_rowvec) touch(P, s.rowvec, idx);
	if (s.has_residual) touch(P, s.residual, idx);
}

static void assign_memory(Plan* P)
{
	int n = (int)P->steps.size();
	for (int i = 0; i < n; ++i) step_touch(P, P->steps[i], i);
	// the zero region (groupnorm statistics) lives for the whole run
	if (P->zero_bytes) {
		P->bufs.push_back(Buf{BUF_ARENA, (P->zero_bytes + 255) / 256 * 256, nullptr, 0, n, 0});
		P->zero_buf = (int)P->bufs.size() - 1;
	}

	// persistent buffers
	size_t poff = 0;
	for (Buf& b : P->bufs) if (b.kind == BUF_PERSIST) { b.off = poff; poff += b.bytes; }
	P->persist_bytes = poff;
	// arena: first-fit over a free list, buffers ordered by first use
	std::vector<int> order;
	for (int i = 0; i < (int)P->bufs.size(); ++i) if (P->bufs[i].kind == BUF_ARENA && P->bufs[i].first >= 0) order.push_back(i);
	std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return P->bufs[a].first < P->bufs[b].first; });
	struct Live { size_t off, bytes; int last; };
	std::vector<Live> live;
	size_t top = 0;
	bool no_reuse = env_flag("GGML_B200_NO_REUSE");
	for (int bi : order) {
		Buf& b = P->bufs[bi];
		// drop buffers whose last use is strictly before this one's first use
		if (!no_reuse)
			live.erase(std::remove_if(live.begin(), live.end(), [&](const Live& l) { return l.last < b.first; }), live.end());
		std::sort(live.begin(), live.end(), [](const Live& x, const Live& y) { return x.off < y.off; });
		size_t pos = 0; bool placed = false;
		for (const Live& l : live) {
			if (l.off >= pos + b.bytes) { placed = true; break; }
			pos = std::max(pos, l.off + l.bytes);
		}
		(void)placed;
		b.off = pos;
		live.push_back({pos, b.bytes, b.last});
		top = std::max(top, pos + b.bytes);
	}
	P->arena_bytes = top;
}

static void* buf_ptr(Plan* P, const PT& p)
{
	const Buf& b = P->bufs[p.buf];
	char* base = b.kind == BUF_FIXED ? (char*)b.fixed : b.kind == BUF_PERSIST ? (char*)P->persist + b.off : (char*)P->arena + b.off;
	return base + p.off * (int64_t)dt_size(p.dt);
}
static View view_of(Plan* P, const PT& p)
{
	View v; v.ptr = buf_ptr(P, p); v.dt = p.dt;
	for (int i = 0; i < 4; ++i) { v.ne[i] = p.ne[i]; v.st[i] = p.st[i]; }
	return v;
}

// ------------------------------------------------------------------ execution
static void run_step(Plan* P, Step& s, cudaStream_t st)
{
	switch (s.kind) {
	case S_COPY: k_copy(st, view_of(P, s.out), view_of(P, s.in[0])); break;
	case S_ZERO: CUDA_CHECK(cudaMemsetAsync(buf_ptr(P, s.out), 0, (size_t)s.out.numel() * dt_size(s.out.dt), st)); break;
	case S_BINARY: k_binary(st, (BinOp)s.iop, view_of(P, s.out), view_of(P, s.in[0]), view_of(P, s.in[1])); break;
	case S_UNARY: k_unary(st, (UnaryOp)s.iop, s.fparam, view_of(P, s.out), view_of(P, s.in[0])); break;
	case S_UPSCALE: k_upscale(st, view_of(P, s.out), view_of(P, s.in[0])); break;
	case S_SOFTMAX_F16:
		k_softmax_f32_f16(st, (const float*)buf_ptr(P, s.in[0]), (__half*)buf_ptr(P, s.out), s.in[0].ne[1], s.in[0].ne[0], s.in[0].ne[0], s.out.ne[0], s.fparam);
		break;
	case S_SOFTMAX: k_softmax_rows(st, view_of(P, s.out), view_of(P, s.in[0]), s.iparam[0] != 0, s.iparam[1]); break;
	case S_GET_ROWS: k_get_rows(st, view_of(P, s.out), view_of(P, s.in[0]), view_of(P, s.in[1])); break;
	case S_TSEMB: k_timestep_embedding(st, view_of(P, s.out), view_of(P, s.in[0]), s.iparam[0], s.iparam[1]); break;
	case S_GEMM_SIMT: k_gemm_simt(st, view_of(P, s.out), view_of(P, s.in[0]), view_of(P, s.in[1]), s.iparam[0] != 0); break;
	case S_GROUPNORM:
		k_groupnorm(st, view_of(P, s.out), view_of(P, s.in[0]),
			s.has_bias ? (const float*)buf_ptr(P, s.bias) : nullptr, s.has_rowvec ? (const float*)buf_ptr(P, s.rowvec) : nullptr,
			s.iparam[0], s.fparam, s.iparam[1] != 0,
			(unsigned long long*)((char*)P->arena + P->bufs[P->zero_buf].off + s.stats_off),
			s.gn_producer >= 0 && gemm_tc_gn_fused(P->steps[s.gn_producer].tc));
		break;
	case S_LAYERNORM:
		k_layernorm(st, view_of(P, s.out), view_of(P, s.in[0]),
			s.has_bias ? (const float*)buf_ptr(P, s.bias) : nullptr, s.has_rowvec ? (const float*)buf_ptr(P, s.rowvec) : nullptr, s.fparam);
		break;
	case S_GEGLU: k_geglu(st, view_of(P, s.out), view_of(P, s.in[0])); break;
	case S_IM2COL:
		k_im2col(st, (__half*)buf_ptr(P, s.out), s.K, view_of(P, s.in[0]), s.iparam[0], s.iparam[1], s.iparam[2], s.iparam[3],
			s.iparam[4], s.iparam[5], s.iparam[6], s.iparam[7], s.M, s.N);
		break;
	case S_WPREP_GEGLU: k_geglu_rows_prep(st, buf_ptr(P, s.out), view_of(P, s.in[0]), s.iparam[0]); break;
	case S_WPREP_CONV: k_conv_weight_prep(st, (__half*)buf_ptr(P, s.out), s.iparam[0], view_of(P, s.in[0])); break;
	case S_ATTENTION: {
		View o = view_of(P, s.out), q = view_of(P, s.in[0]), k = view_of(P, s.in[1]), v = view_of(P, s.in[2]);
		if (!s.atc_checked) {
			s.atc_checked = true;
			const char* e = getenv("GGML_B200_ATTN");
			bool simt = e && !strcmp(e, "simt");
			if (!simt && attn_tc_supported(o, q, k, v, s.iparam[0] != 0)) s.atc = attn_tc_prepare(o, q, k, v, s.fparam);
		}
		if (s.atc) attn_tc_launch(st, s.atc);
		else k_attention(st, o, q, k, v, s.fparam, s.iparam[0] != 0);
	} break;
	case S_GEMM_TC: case S_CONV_TC: {
		if (s.kind == S_CONV_TC && !s.tc && !s.has_rowvec && !s.has_residual && s.act == U_NONE && s.out.dt == DT_F16 && s.gn_groups == 0 &&
			k_conv3x3_small_supported(s.conv_c, s.N)) {          // 3 / 4 output channels: direct kernel (conv_out, last VAE convolution)
			k_conv3x3_small(st, (__half*)buf_ptr(P, s.out), (const __half*)buf_ptr(P, s.in[0]), (const __half*)buf_ptr(P, s.in[1]),
				s.has_bias ? (const float*)buf_ptr(P, s.bias) : nullptr, s.conv_n, s.conv_h, s.conv_w, s.conv_c, s.N);
			break;
		}
		if (!s.tc) {
			GemmEpilogue ep;
			if (s.has_bias) ep.bias = (const float*)buf_ptr(P, s.bias);
			if (s.has_rowvec) { ep.rowvec = buf_ptr(P, s.rowvec); ep.rowvec_dt = s.rowvec.dt; ep.rowvec_stride = s.rowvec.ne[3] > 1 ? s.rowvec.st[3] : 0; ep.rows_per_image = s.rows_per_image; }
			if (s.has_residual) { ep.residual = buf_ptr(P, s.residual); ep.residual_dt = s.residual.dt; ep.ldr = s.ldc; }
			ep.act = s.act;
			ep.geglu = s.kind == S_GEMM_TC && s.iparam[0] == 1;
			if (s.gn_groups > 0) {
				ep.gn_stats = (unsigned long long*)((char*)P->arena + P->bufs[P->zero_buf].off + s.stats_off);
				ep.gn_groups = s.gn_groups; ep.gn_rows_per_image = s.gn_rows_per_image;
			}
			if (s.kind == S_GEMM_TC)
				s.tc = gemm_tc_prepare((const __half*)buf_ptr(P, s.in[0]), s.lda, (const __half*)buf_ptr(P, s.in[1]), s.ldb,
					buf_ptr(P, s.out), s.out.dt, s.ldc, s.M, s.N, s.K, ep, P->be->sm_count);
			else
				s.tc = conv3x3_tc_prepare((const __half*)buf_ptr(P, s.in[0]), s.conv_n, s.conv_h, s.conv_w, s.conv_c,
					(const __half*)buf_ptr(P, s.in[1]), buf_ptr(P, s.out), s.out.dt, s.N, ep, P->be->sm_count);
		}
		gemm_tc_launch(st, s.tc);
	} break;
	}
}

// step histogram of the most recently built plan (ggml_b200_last_plan_count): lets the CPU-only tests check what the
// planner fused, in dry-run mode
static std::map<std::string, int> g_last_plan_hist;
int last_plan_count(const char* key) { auto it = g_last_plan_hist.find(key ? key : "steps"); return it == g_last_plan_hist.end() ? 0 : it->second; }

static void dump_plan(Plan* P)
{
	static const char* kn[] = { "COPY", "BINARY", "UNARY", "UPSCALE", "SOFTMAX", "GET_ROWS", "TSEMB", "GEMM_SIMT",
		"GROUPNORM", "LAYERNORM", "GEGLU", "IM2COL", "WPREP_CONV", "ATTENTION", "GEMM_TC", "CONV_TC", "ZERO", "SOFTMAX_F16", "WPREP_GEGLU" };
	std::map<std::string, int> hist;
	for (Step& s : P->steps) hist[kn[s.kind]]++;
	g_last_plan_hist = hist; g_last_plan_hist["steps"] = (int)P->steps.size(); g_last_plan_hist["preps"] = (int)P->prep.size();
	for (Step& s : P->steps) g_last_plan_hist[std::string("name:") + s.name]++;
	for (Step& s : P->steps) if (s.kind == S_GROUPNORM && s.gn_producer >= 0) g_last_plan_hist["gn_from_epilogue"]++;    // planned; the producer's tile mode decides
	std::string line;
	for (auto& kv : hist) line += kv.first + ":" + std::to_string(kv.second) + " ";
	if (env_flag("GGML_B200_QUIET")) return;      // (the histogram above is always recorded)
	B200_LOG("plan: %zu nodes -> %zu steps (+%zu weight preps), arena %.1f MiB, prepared weights %.1f MiB | %s",
		P->graph->nodes.size(), P->steps.size(), P->prep.size(), P->arena_bytes / 1048576.0, P->persist_bytes / 1048576.0, line.c_str());
	if (env_flag("GGML_B200_DUMP_PLAN")) {
		int i = 0;
		for (Step& s : P->steps) {
			fprintf(stderr, "  %4d %-10s %-18s out[%lld,%lld,%lld,%lld] st[%lld,%lld,%lld,%lld] dt%d", i++, kn[s.kind], s.name,
				(long long)s.out.ne[0], (long long)s.out.ne[1], (long long)s.out.ne[2], (long long)s.out.ne[3],
				(long long)s.out.st[0], (long long)s.out.st[1], (long long)s.out.st[2], (long long)s.out.st[3], (int)s.out.dt);
			if (s.kind == S_GEMM_TC || s.kind == S_CONV_TC)
				fprintf(stderr, " M%lld N%lld K%lld%s%s%s act%d", (long long)s.M, (long long)s.N, (long long)s.K,
					s.has_bias ? " +bias" : "", s.has_rowvec ? " +rowvec" : "", s.has_residual ? " +res" : "", (int)s.act);
			fputc('\n', stderr);
		}
	}
}

Plan* plan_build(Backend* be, ggml_cgraph* g)
{
	Plan* P = new Plan();
	P->be = be; P->graph = g;
	P->use_graph = !env_flag("GGML_B200_NO_CUDA_GRAPH");
	Builder B; B.P = P; B.g = g;
	const char* gm = getenv("GGML_B200_GEMM");
	B.force_simt = gm && !strcmp(gm, "simt");
	B.plan();
	assign_memory(P);
	if (!g_dryrun()) {
		CUDA_CHECK(cudaSetDevice(be->device));
		if (P->arena_bytes) CUDA_CHECK(cudaMalloc(&P->arena, P->arena_bytes));
		if (P->persist_bytes) CUDA_CHECK(cudaMalloc(&P->persist, P->persist_bytes));
	}
	g_stats.plans_built++;
	dump_plan(P);
	return P;
}

static void step_cost(const Step& s, double* flops, double* bytes)
{
	auto nb = [](const PT& p) { return (double)p.numel() * dt_size(p.dt); };
	*flops = 0; *bytes = nb(s.out);
	for (int i = 0; i < s.n_in; ++i) *bytes += nb(s.in[i]);
	if (s.has_residual) *bytes += nb(s.residual);
	if (s.kind == S_GEMM_TC || s.kind == S_CONV_TC) *flops = 2.0 * s.M * s.N * s.K;
	else if (s.kind == S_GEMM_SIMT) *flops = 2.0 * s.in[0].ne[0] * (double)s.out.numel();
	else if (s.kind == S_ATTENTION) *flops = 4.0 * s.in[0].ne[0] * s.in[0].ne[1] * s.in[1].ne[1] * s.in[0].ne[2] * s.in[0].ne[3];
}

static void plan_run_profiled(Plan* P)
{
	cudaStream_t st = P->be->stream;
	cudaEvent_t e0, e1;
	CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
	if (P->zero_bytes) CUDA_CHECK(cudaMemsetAsync((char*)P->arena + P->bufs[P->zero_buf].off, 0, P->zero_bytes, st));
	for (Step& s : P->steps) {
		uint64_t l0 = g_stats.kernel_launches;
		k_spin(st, 25);                        // the step's launches queue up behind it: host launch latency stays out of the timing
		CUDA_CHECK(cudaEventRecord(e0, st));
		run_step(P, s, st);
		CUDA_CHECK(cudaEventRecord(e1, st));
		CUDA_CHECK(cudaEventSynchronize(e1));
		float ms = 0; CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
		double fl, by; step_cost(s, &fl, &by);
		if (env_flag("GGML_B200_PROFILE_STEPS"))
			fprintf(stderr, "step %-12s kind %2d  %8.1f us  out[%lld,%lld,%lld,%lld] in0[%lld,%lld,%lld,%lld] M%lld N%lld K%lld %s\n", s.name, (int)s.kind, ms * 1e3,
				(long long)s.out.ne[0], (long long)s.out.ne[1], (long long)s.out.ne[2], (long long)s.out.ne[3],
				(long long)s.in[0].ne[0], (long long)s.in[0].ne[1], (long long)s.in[0].ne[2], (long long)s.in[0].ne[3],
				(long long)s.M, (long long)s.N, (long long)s.K, s.atc ? "tc" : "");
		ProfAcc& a = g_prof[s.kind];
		a.ms += ms; a.flops += fl; a.bytes += by; a.launches += g_stats.kernel_launches - l0;
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
}

void plan_run(Plan* P)
{
	cudaStream_t st = P->be->stream;
	// weight preparation: only for leaves whose contents changed since the last run
	for (Step& s : P->prep) {
		uint64_t v = trec(s.leaf)->version;
		if (v != s.leaf_version) { run_step(P, s, st); s.leaf_version = v; }
	}
	if (g_profile) { plan_run_profiled(P); return; }
	if (P->use_graph && !P->exec) {
		// first run: execute eagerly once (creates tensor maps, sets function attributes), then capture
		uint64_t l0 = g_stats.kernel_launches;
		if (P->zero_bytes) CUDA_CHECK(cudaMemsetAsync((char*)P->arena + P->bufs[P->zero_buf].off, 0, P->zero_bytes, st));
		for (Step& s : P->steps) run_step(P, s, st);
		P->launches_per_run = g_stats.kernel_launches - l0;
		CUDA_CHECK(cudaStreamSynchronize(st));
		cudaGraph_t graph;
		CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
		if (P->zero_bytes) CUDA_CHECK(cudaMemsetAsync((char*)P->arena + P->bufs[P->zero_buf].off, 0, P->zero_bytes, st));
		for (Step& s : P->steps) run_step(P, s, st);
		CUDA_CHECK(cudaStreamEndCapture(st, &graph));
		g_stats.kernel_launches -= P->launches_per_run;   // capture pass launched nothing
		CUDA_CHECK(cudaGraphInstantiate(&P->exec, graph, 0));
		CUDA_CHECK(cudaGraphDestroy(graph));
		return;
	}
	if (P->use_graph) {
		CUDA_CHECK(cudaGraphLaunch(P->exec, st));
		g_stats.graph_launches++;
		g_stats.kernel_launches += P->launches_per_run;
		return;
	}
	if (P->zero_bytes) CUDA_CHECK(cudaMemsetAsync((char*)P->arena + P->bufs[P->zero_buf].off, 0, P->zero_bytes, st));
	for (Step& s : P->steps) {
		run_step(P, s, st);
		if (env_flag("GGML_B200_SYNC_STEPS")) {
			cudaError_t e = cudaStreamSynchronize(st);
			if (e != cudaSuccess) B200_FATAL("step '%s' failed: %s", s.name, cudaGetErrorString(e));
		}
	}
}

void plan_free(Plan* P)
{
	if (!P) return;
	if (g_dryrun()) { delete P; return; }
	cudaStreamSynchronize(P->be->stream);
	for (Step& s : P->steps) { if (s.tc) gemm_tc_free(s.tc); if (s.atc) attn_tc_free(s.atc); }
	if (P->exec) cudaGraphExecDestroy(P->exec);
	if (P->arena) cudaFree(P->arena);
	if (P->persist) cudaFree(P->persist);
	delete P;
}

}  // namespace b200
