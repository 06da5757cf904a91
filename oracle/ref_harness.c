/*
 * ref_harness.c -- thin C harness over the UNMODIFIED reference (libphyc compiled from
 * /root/reference by oracle/Makefile into oracle/_ref/).
 *
 * TEST INFRASTRUCTURE ONLY.  Used to (a) validate oracle/phb_oracle.c, (b) generate the golden
 * fixtures under tests/golden/ (oracle/make_golden.py) and (c) time the reference's CPU path for
 * bench.py's cpu_baseline / --impl reference legs.  Never loaded by the product path.
 *
 * It builds a "treelikelihood" Model through the reference's own JSON constructor
 * (treelikelihood.c:819) from a JSON text passed by the caller (inline "sequences" and "newick",
 * so no files are needed at run time) and exposes plain-array getters for everything the
 * B200 path takes as input and everything it must reproduce.
 *
 * Compiled against the reference headers where they lie (-I/root/reference/src); this file
 * contains no reference code.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "phyc/branchmodel.h"
#include "phyc/datatype.h"
#include "phyc/gy94.h"
#include "phyc/sequence.h"
#include "phyc/simplex.h"
#include "phyc/hashtable.h"
#include "phyc/matrix.h"
#include "phyc/mjson.h"
#include "phyc/parameters.h"
#include "phyc/sitemodel.h"
#include "phyc/sitepattern.h"
#include "phyc/substmodel.h"
#include "phyc/tree.h"
#include "phyc/treelikelihood.h"

typedef struct RefH {
	Hashtable *hash;
	json_node *json;
	Model *model;
	SingleTreeLikelihood *tlk;
} RefH;

void *refh_create(const char *json_text) {
	RefH *h = (RefH *)calloc(1, sizeof(RefH));
	h->hash = new_Hashtable_string(100);
	hashtable_set_key_ownership(h->hash, false);
	hashtable_set_value_ownership(h->hash, false);
	h->json = create_json_tree(json_text);
	json_node *child = h->json->children[0];
	h->model = new_TreeLikelihoodModel_from_json(child, h->hash);
	h->tlk = (SingleTreeLikelihood *)h->model->obj;
	if (Tree_is_time_mode(h->tlk->tree)) Tree_update_heights(h->tlk->tree);
	return h;
}

/*
 * GY94 codon model built through the C API: the JSON factory has empty GY94/MG94 branches
 * (substmodel.c:1526-1537) and the shipped >= 60-state dispatcher is stale (SURVEY 8c caveat 1),
 * so the generic function pointers are installed by hand and tip partials are used.
 */
void *refh_create_codon(const char *newick, int n, const char **names, const char **seqs, double kappa, double omega) {
	RefH *h = (RefH *)calloc(1, sizeof(RefH));
	DataType *dt = new_CodonDataType(0);
	Sequences *aln = new_Sequences(n);
	for (int i = 0; i < n; i++) Sequences_add(aln, new_Sequence(names[i], seqs[i]));
	aln->datatype = dt;
	SitePattern *sp = new_SitePattern(aln);
	free_Sequences(aln);
	Tree *tree = new_Tree(newick, true);
	Model *mtree = new_TreeModel("tree", tree);
	int S = dt->state_count(dt);
	double *f = (double *)malloc(sizeof(double) * S);
	for (int i = 0; i < S; i++) f[i] = 1.0 / S;
	Simplex *fs = new_Simplex_with_values("freqs", f, S);
	free(f);
	Model *mfs = new_SimplexModel("freqs", fs);
	SubstitutionModel *m = new_GY94_with_values(fs, omega, kappa, 0);
	Model *mm = new_SubstitutionModel2("sm", m, mfs, NULL);
	SiteModel *sm = new_SiteModel_with_parameters(NULL, NULL, 1, DISTRIBUTION_UNIFORM, false, QUADRATURE_QUANTILE_MEDIAN);
	Model *msm = new_SiteModel2("sitemodel", sm, NULL);
	SingleTreeLikelihood *tlk = new_SingleTreeLikelihood(tree, m, sm, sp, NULL, false);
	h->model = new_TreeLikelihoodModel("treelikelihood", tlk, mtree, mm, msm, NULL);
	h->tlk = tlk;
	tlk->include_jacobian = false; /* only the JSON factory initialises this field (treelikelihood.c:935): model->logP would add the tree model's Jacobian */
	extern void refh_use_generic_kernels(void *);
	refh_use_generic_kernels(h);
	return h;
}

/* Model.clone (what gradascent.c:166-170 does per worker): a second handle on an independent copy of the model graph */
void *refh_clone(void *vh) {
	RefH *src = (RefH *)vh;
	RefH *h = (RefH *)calloc(1, sizeof(RefH));
	h->hash = new_Hashtable_string(10);
	hashtable_set_key_ownership(h->hash, false);
	hashtable_set_value_ownership(h->hash, false);
	h->model = src->model->clone(src->model, h->hash);
	h->tlk = (SingleTreeLikelihood *)h->model->obj;
	return h;
}

void refh_free(void *vh) {
	RefH *h = (RefH *)vh;
	if (h->hash != NULL) { /* models built by hand are left to the process exit */
		h->model->free(h->model);
		free_Hashtable(h->hash);
		if (h->json) json_free_tree(h->json);
	}
	free(h);
}

/* out: T, N, S, C, P, root, time_mode, scale, use_tip_states, right child of root */
void refh_dims(void *vh, int *out) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	out[0] = Tree_tip_count(tlk->tree);
	out[1] = Tree_node_count(tlk->tree);
	out[2] = tlk->m->nstate;
	out[3] = tlk->cat_count;
	out[4] = tlk->pattern_count;
	out[5] = Node_id(Tree_root(tlk->tree));
	out[6] = Tree_is_time_mode(tlk->tree);
	out[7] = tlk->scale;
	out[8] = tlk->use_tip_states;
	out[9] = Node_id(Tree_root(tlk->tree)->right);
}

void refh_topology(void *vh, int *left, int *right, int *parent) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	int N = Tree_node_count(tlk->tree);
	for (int i = 0; i < N; i++) {
		Node *n = Tree_node(tlk->tree, i);
		int id = Node_id(n);
		left[id] = n->left ? Node_id(n->left) : -1;
		right[id] = n->right ? Node_id(n->right) : -1;
		parent[id] = n->parent ? Node_id(n->parent) : -1;
	}
}

/* effective branch lengths as _calculate_partials reads them (treelikelihood.c:1652-1663) */
void refh_branch_lengths(void *vh, double *bl, double *dt) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	int N = Tree_node_count(tlk->tree);
	int time_mode = Tree_is_time_mode(tlk->tree);
	if (time_mode) Tree_update_heights(tlk->tree);
	for (int i = 0; i < N; i++) {
		Node *n = Tree_node(tlk->tree, i);
		int id = Node_id(n);
		if (Node_isroot(n)) {
			bl[id] = 0;
			if (dt) dt[id] = 0;
			continue;
		}
		if (tlk->bm == NULL || !time_mode) {
			bl[id] = Node_distance(n);
			if (dt) dt[id] = 0;
		} else {
			bl[id] = tlk->bm->get(tlk->bm, n) * Node_time_elapsed(n);
			if (dt) dt[id] = Node_time_elapsed(n);
		}
	}
}

void refh_set_branch_lengths(void *vh, const double *bl) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	int N = Tree_node_count(tlk->tree);
	for (int i = 0; i < N; i++) {
		Node *n = Tree_node(tlk->tree, i);
		if (!Node_isroot(n)) Node_set_distance(n, bl[Node_id(n)]);
	}
	SingleTreeLikelihood_update_all_nodes(tlk);
}

/* uint8 states per tip node id (rows follow node ids through tlk->mapping) */
void refh_tip_states(void *vh, uint8_t *out) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	int N = Tree_node_count(tlk->tree);
	int P = tlk->pattern_count;
	for (int i = 0; i < N; i++) {
		Node *n = Tree_node(tlk->tree, i);
		if (!Node_isleaf(n)) continue;
		int id = Node_id(n);
		memcpy(out + (size_t)id * P, tlk->sp->patterns[tlk->mapping[id]], P);
	}
}

/* tip partials [T][P][S] through sp->get_partials (treelikelihood.c:1109) */
void refh_tip_partials(void *vh, double *out) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	int N = Tree_node_count(tlk->tree);
	int P = tlk->pattern_count, S = tlk->m->nstate;
	for (int i = 0; i < N; i++) {
		Node *n = Tree_node(tlk->tree, i);
		if (!Node_isleaf(n)) continue;
		int id = Node_id(n);
		tlk->sp->get_partials(tlk->sp, tlk->mapping[id], out + (size_t)id * P * S);
	}
}

void refh_weights(void *vh, double *out) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	memcpy(out, tlk->sp->weights, sizeof(double) * tlk->pattern_count);
}

/* eigen system + frequencies; returns 1 when the model keeps an eigen decomposition */
int refh_model(void *vh, double *evec, double *eval, double *ivec, double *freqs) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	SubstitutionModel *m = tlk->m;
	int S = m->nstate;
	const double *f = tlk->get_root_frequencies(tlk);
	memcpy(freqs, f, sizeof(double) * S);
	if (m->need_update) {
		m->update_Q(m);
		if (m->eigendcmp != NULL && m->modeltype != JC69) update_eigen_system(m);
	}
	if (m->eigendcmp == NULL || m->modeltype == JC69) return 0;
	for (int i = 0; i < S; i++) {
		eval[i] = m->eigendcmp->eval[i];
		for (int j = 0; j < S; j++) {
			evec[i * S + j] = m->eigendcmp->evec[i][j];
			ivec[i * S + j] = m->eigendcmp->Invevec[i][j];
		}
	}
	return 1;
}

void refh_sitemodel(void *vh, double *rates, double *props) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	tlk->sm->update(tlk->sm);
	double *p = tlk->sm->get_proportions(tlk->sm);
	for (int c = 0; c < tlk->cat_count; c++) {
		rates[c] = tlk->sm->get_rate(tlk->sm, c);
		props[c] = p[c];
	}
}

/* row-major P(t) and dP/dt per (node, category) straight from m->p_t / m->dp_dt */
void refh_matrices(void *vh, double *Pm, double *dPm) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	int N = Tree_node_count(tlk->tree);
	int S = tlk->m->nstate, C = tlk->cat_count;
	double *bl = (double *)malloc(sizeof(double) * N);
	refh_branch_lengths(vh, bl, NULL);
	tlk->sm->update(tlk->sm);
	for (int id = 0; id < N; id++) {
		for (int c = 0; c < C; c++) {
			double t = bl[id] * tlk->sm->get_rate(tlk->sm, c);
			size_t off = ((size_t)id * C + c) * S * S;
			if (id == Node_id(Tree_root(tlk->tree))) {
				memset(Pm + off, 0, sizeof(double) * S * S);
				memset(dPm + off, 0, sizeof(double) * S * S);
				continue;
			}
			tlk->m->p_t(tlk->m, t, Pm + off);
			tlk->m->dp_dt(tlk->m, t, dPm + off);
		}
	}
	free(bl);
}

void refh_use_rescaling(void *vh, int use) { SingleTreeLikelihood_use_rescaling(((RefH *)vh)->tlk, use != 0); }
int refh_rescaling(void *vh) { return ((RefH *)vh)->tlk->scale; }
void refh_enable_sse(void *vh, int v) { SingleTreeLikelihood_enable_SSE(((RefH *)vh)->tlk, v != 0); }
void refh_set_include_jacobian(void *vh, int v) { ((RefH *)vh)->tlk->include_jacobian = v != 0; }

/* codon and other >= 60-state models: the shipped dispatcher is stale (SURVEY 8c caveat 1) */
void refh_use_generic_kernels(void *vh) {
	extern void update_partials_general(SingleTreeLikelihood *, int, int, int, int, int);
	extern void integrate_partials_general(const SingleTreeLikelihood *, const double *, const double *, double *);
	extern void node_log_likelihoods_general(const SingleTreeLikelihood *, const double *, const double *, double *);
	extern void calculate_branch_partials(SingleTreeLikelihood *, double *, int, int, int);
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	tlk->update_partials = update_partials_general;
	tlk->integrate_partials = integrate_partials_general;
	tlk->node_log_likelihoods = node_log_likelihoods_general;
	tlk->calculate_per_cat_partials = calculate_branch_partials;
}

/* full recomputation, protocol of examples/benchmarking.c:466-471 */
double refh_logP(void *vh) {
	RefH *h = (RefH *)vh;
	SingleTreeLikelihood_update_all_nodes(h->tlk);
	h->tlk->m->need_update = true;
	return h->model->logP(h->model);
}

void refh_pattern_lnl(void *vh, double *out) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	memcpy(out, tlk->pattern_lk, sizeof(double) * tlk->pattern_count);
}

/*
 * TreeLikelihood_initialize_gradient(flags) then TreeLikelihood_gradient (treelikelihood.c:237,320).
 * include_root_freqs: -1 keep what initialize_gradient set, else override (public field, :123).
 * Returns the gradient length; out receives min(length, cap) values.
 */
int refh_gradient(void *vh, int flags, int include_root_freqs, double *out, int cap) {
	RefH *h = (RefH *)vh;
	size_t len = TreeLikelihood_initialize_gradient(h->model, flags);
	if (include_root_freqs >= 0) h->tlk->include_root_freqs = include_root_freqs != 0;
	SingleTreeLikelihood_update_all_nodes(h->tlk);
	h->tlk->m->need_update = true;
	double *g = TreeLikelihood_gradient(h->model);
	for (size_t i = 0; i < len && (int)i < cap; i++) out[i] = g[i];
	return (int)len;
}

/* copy of one partials buffer: idx < N lower (NULL for state tips -> returns 0), idx >= N upper */
int refh_partials(void *vh, int idx, double *out) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	double *p = tlk->partials[tlk->current_partials_indexes[idx]][idx];
	if (p == NULL) return 0;
	memcpy(out, p, sizeof(double) * tlk->partials_size);
	return 1;
}

int refh_scaling_factors(void *vh, int idx, double *out) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	if (!tlk->scale || tlk->scaling_factors == NULL) return 0;
	double *p = tlk->scaling_factors[tlk->current_partials_indexes[idx]][idx];
	if (p == NULL) return 0;
	memcpy(out, p, sizeof(double) * tlk->pattern_count);
	return 1;
}

static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC_RAW, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* seconds per lnL evaluation, examples/benchmarking.c:466-471 */
double refh_time_logP(void *vh, int iters, double *last) {
	RefH *h = (RefH *)vh;
	double v = 0;
	double t0 = now_s();
	for (int i = 0; i < iters; i++) {
		SingleTreeLikelihood_update_all_nodes(h->tlk);
		h->tlk->m->need_update = true;
		v = h->model->logP(h->model);
	}
	double t1 = now_s();
	if (last) *last = v;
	return (t1 - t0) / iters;
}

/* seconds per lnL + gradient evaluation, examples/benchmarking.c:498-503 */
double refh_time_gradient(void *vh, int flags, int include_root_freqs, int iters) {
	RefH *h = (RefH *)vh;
	TreeLikelihood_initialize_gradient(h->model, flags);
	if (include_root_freqs >= 0) h->tlk->include_root_freqs = include_root_freqs != 0;
	double t0 = now_s();
	for (int i = 0; i < iters; i++) {
		SingleTreeLikelihood_update_all_nodes(h->tlk);
		h->tlk->m->need_update = true;
		TreeLikelihood_gradient(h->model);
	}
	double t1 = now_s();
	return (t1 - t0) / iters;
}

/* ---- hooks for the drop-in test of integration/physher_glue.c (tests/test_glue_dropin.py) ---- */

void *refh_model_handle(void *vh) { return ((RefH *)vh)->model; }

size_t refh_initialize_gradient(void *vh, int flags, int include_root_freqs) {
	RefH *h = (RefH *)vh;
	size_t len = TreeLikelihood_initialize_gradient(h->model, flags);
	if (include_root_freqs >= 0) h->tlk->include_root_freqs = include_root_freqs != 0;
	return len;
}

void refh_mark_dirty(void *vh) {
	RefH *h = (RefH *)vh;
	SingleTreeLikelihood_update_all_nodes(h->tlk);
	h->tlk->m->need_update = true;
}

/*
 * The request sequence of the reference's own known-answer test (tests/test_tree_likelihood.c:33-77), through the
 * Model vtable: prepare_gradient over {clock rate(s), ratios}, then dlogP per parameter.
 * out[0] = d logP / d clock rate, out[1..] = d logP / d ratio_i (the last one is the root height).  Returns the count.
 */
int refh_kat_dlogP(void *vh, double *out, int cap) {
	RefH *h = (RefH *)vh;
	Model *model = h->model;
	Model **models = (Model **)model->data;
	Tree *tree = (Tree *)models[0]->obj;
	BranchModel *bm = (BranchModel *)models[3]->obj;
	Parameters *ps = new_Parameters(10);
	for (size_t i = 0; i < Parameters_count(bm->rates); i++) Parameters_add(ps, Parameters_at(bm->rates, i));
	Parameters *ratios = get_reparams(tree);
	for (size_t i = 0; i < Parameters_count(ratios); i++) Parameters_add(ps, Parameters_at(ratios, i));
	model->prepare_gradient(model, ps);
	SingleTreeLikelihood_update_all_nodes(h->tlk);
	int n = 0;
	if (n < cap) out[n++] = model->dlogP(model, Parameters_at(ps, 0));
	for (size_t i = 0; i < Parameters_count(ratios) && n < cap; i++) out[n++] = model->dlogP(model, Parameters_at(ratios, i));
	free_Parameters(ps);
	return n;
}

/* inputs of the time-tree chain as the reference holds them: tip heights [T] by node id, reparameterisation values [T-1] by
 * class id (root entry = root height), clock rates (returns their count: 1 strict, else one per node by node id) */
int refh_time_tree(void *vh, double *tip_heights, double *ratios, double *rates) {
	RefH *h = (RefH *)vh;
	Model **models = (Model **)h->model->data;
	Tree *tree = (Tree *)models[0]->obj;
	BranchModel *bm = (BranchModel *)models[3]->obj;
	Tree_update_heights(tree);
	int N = Tree_node_count(tree);
	for (int i = 0; i < N; i++) {
		Node *n = Tree_node(tree, i);
		if (Node_isleaf(n)) tip_heights[Node_id(n)] = Node_height(n);
	}
	Parameters *rp = get_reparams(tree);
	for (size_t i = 0; i < Parameters_count(rp); i++) ratios[i] = Parameters_value(rp, i);
	int nr = (int)Parameters_count(bm->rates);
	if (nr == 1) rates[0] = Parameters_value(bm->rates, 0);
	else
		for (int i = 0; i < N; i++) {
			Node *n = Tree_node(tree, i);
			if (!Node_isroot(n)) rates[Node_id(n)] = bm->get(bm, n);
		}
	return nr;
}

void refh_set_ratios(void *vh, const double *ratios) {
	RefH *h = (RefH *)vh;
	Model **models = (Model **)h->model->data;
	Tree *tree = (Tree *)models[0]->obj;
	Parameters *rp = get_reparams(tree);
	for (size_t i = 0; i < Parameters_count(rp); i++) Parameters_set_value(rp, i, ratios[i]);
	Tree_update_heights(tree);
	SingleTreeLikelihood_update_all_nodes(h->tlk);
}

void refh_set_clock_rate(void *vh, double rate) {
	RefH *h = (RefH *)vh;
	Model **models = (Model **)h->model->data;
	BranchModel *bm = (BranchModel *)models[3]->obj;
	Parameters_set_value(bm->rates, 0, rate);
	SingleTreeLikelihood_update_all_nodes(h->tlk);
}

/* sp->patterns [size][count] in alignment order (NOT re-mapped to node ids), sp->weights; returns sp->size */
int refh_patterns_raw(void *vh, uint8_t *patterns, double *weights) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	SitePattern *sp = tlk->sp;
	for (int i = 0; i < sp->size; i++) memcpy(patterns + (size_t)i * sp->count, sp->patterns[i], sp->count);
	memcpy(weights, sp->weights, sizeof(double) * sp->count);
	return sp->size;
}
const char *refh_pattern_name(void *vh, int i) { return ((RefH *)vh)->tlk->sp->names[i]; }

/* ---- substitution-model parameter gradients (calculate_dlnl_dQ, treelikelihood.c:2337-2583) ---- */

/* dP/d theta_index per (node, category) straight from m->dPdp at t = bl * rate_c (what :2421 builds), [N][C][S][S] */
void refh_dPdp(void *vh, int index, double *out) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	int N = Tree_node_count(tlk->tree), S = tlk->m->nstate, C = tlk->cat_count;
	double *bl = (double *)malloc(sizeof(double) * N);
	refh_branch_lengths(vh, bl, NULL);
	tlk->m->dQ_need_update = true;
	for (int id = 0; id < N; id++)
		for (int c = 0; c < C; c++) {
			size_t off = ((size_t)id * C + c) * S * S;
			if (id == Node_id(Tree_root(tlk->tree))) memset(out + off, 0, sizeof(double) * S * S);
			else tlk->m->dPdp(tlk->m, index, out + off, bl[id] * tlk->sm->get_rate(tlk->sm, c));
		}
	free(bl);
}

/* the reference's own value for parameter `index`: lnL + upper partials through TreeLikelihood_gradient (tree flags,
 * include_root_freqs as given), then calculate_dlnl_dQ on its pattern likelihoods */
double refh_dlnl_dQ(void *vh, int index, int include_root_freqs) {
	RefH *h = (RefH *)vh;
	double tmp[8];
	refh_gradient(vh, TREELIKELIHOOD_FLAG_TREE_MODEL, include_root_freqs, tmp, 8);
	return calculate_dlnl_dQ(h->tlk, index, h->tlk->pattern_lk + h->tlk->sp->count);
}

/* ---- single-branch "upper likelihood" functions (treelikelihood.c:2195-2335, 2592-2686) ---- */

/* The reference's own lnL, d lnL/dt and d2 lnL/dt2 for the branch above node `id` at length `bl`, exactly as its Brent / Newton
 * drivers obtain them: tlk->calculate_upper (= _calculate_uppper; node_upper == NULL => _calculate_simple + update_upper_partials),
 * then calculate_dldt_uppper and d2lnldt2_uppper on exp(pattern_lk) (the protocol of _singleTreeLikelihood_d2logP, :470-527).
 * The branch length is put back afterwards.  Not for the root or the root's right child (:2202-2205).  Returns 0 on success. */
int refh_branch_derivatives(void *vh, int id, double bl, double *out) {
	RefH *h = (RefH *)vh;
	SingleTreeLikelihood *tlk = h->tlk;
	Node *node = Tree_node(tlk->tree, id);
	if (Node_isroot(node) || Tree_root(tlk->tree)->right == node) return 1;
	const int P = tlk->sp->count;
	const double old = Node_distance(node);
	Node_set_distance(node, bl);
	SingleTreeLikelihood_update_all_nodes(tlk); /* node_upper = NULL */
	tlk->m->need_update = true;
	out[0] = tlk->calculate_upper(tlk, node);
	double *lk = (double *)malloc(sizeof(double) * P), *dlk = (double *)malloc(sizeof(double) * P);
	for (int k = 0; k < P; k++) lk[k] = exp(tlk->pattern_lk[k]);
	calculate_dldt_uppper(tlk, node, dlk);
	double d1 = 0;
	for (int k = 0; k < P; k++) d1 += dlk[k] / lk[k] * tlk->sp->weights[k];
	out[1] = d1;
	out[2] = d2lnldt2_uppper(tlk, node, lk, dlk);
	free(lk);
	free(dlk);
	Node_set_distance(node, old);
	SingleTreeLikelihood_update_all_nodes(tlk);
	tlk->m->need_update = true;
	tlk->use_upper = false;
	return 0;
}

/*
 * The access pattern of serial_brent_optimize_tree (optimizer.c:111-152) with fixed candidate lengths instead of Brent's: open
 * with update_uppers (the reference's own, or the one passed in -- a drop-in backend's), then visit the branches in post-order
 * (not the root, not the root's right child) and evaluate tlk->calculate at d0 * factors[k] for each; the last factor stays.
 * out receives one lnL per (branch, factor); ids (optional) the node id of each visit.  Returns the number of values.
 */
int refh_upper_walk(void *vh, void (*update_uppers)(Model *), const double *factors, int nf, double *out, int *ids, int cap) {
	RefH *h = (RefH *)vh;
	SingleTreeLikelihood *tlk = h->tlk;
	Tree *tree = tlk->tree;
	Node **nodes = Tree_get_nodes(tree, POSTORDER);
	tlk->node_upper = NULL;
	tlk->use_upper = true;
	tlk->update_upper = true;
	if (update_uppers) update_uppers(h->model);
	else SingleTreeLikelihood_update_uppers(tlk);
	int k = 0;
	for (int i = 0; i < Tree_node_count(tree); i++) {
		Node *node = nodes[i];
		if (Node_isroot(node) || (Node_isroot(Node_parent(node)) && Node_right(Node_parent(node)) == node)) continue;
		if (tlk->node_upper == NULL) tlk->node_upper = node;
		const double d0 = Node_distance(node);
		for (int f = 0; f < nf && k < cap; f++) {
			Node_set_distance(node, d0 * factors[f]);
			SingleTreeLikelihood_update_one_node(tlk, node);
			if (ids) ids[k] = Node_id(node);
			out[k++] = tlk->calculate(tlk);
		}
	}
	tlk->use_upper = false;
	SingleTreeLikelihood_update_all_nodes(tlk);
	return k;
}

/* model->d2logP for the branch-length parameter of node `id` (Model vtable; _singleTreeLikelihood_d2logP, treelikelihood.c:470-527) */
double refh_d2logP_branch(void *vh, int id) {
	RefH *h = (RefH *)vh;
	Node *node = Tree_node(h->tlk->tree, id);
	SingleTreeLikelihood_update_all_nodes(h->tlk);
	return h->model->d2logP(h->model, node->distance);
}

/* ---- hooks for the boundary rows of tests/test_glue_dropin.py: JSON factory, direct slots, Model.store / restore ---- */

/* like refh_create, through a caller-supplied factory with the signature of new_TreeLikelihoodModel_from_json (the glue's
 * phb_physher_new_TreeLikelihoodModel_from_json, which understands "backend" / "device") */
void *refh_create_with(const char *json_text, Model *(*factory)(json_node *, Hashtable *)) {
	RefH *h = (RefH *)calloc(1, sizeof(RefH));
	h->hash = new_Hashtable_string(100);
	hashtable_set_key_ownership(h->hash, false);
	hashtable_set_value_ownership(h->hash, false);
	h->json = create_json_tree(json_text);
	h->model = factory(h->json->children[0], h->hash);
	h->tlk = (SingleTreeLikelihood *)h->model->obj;
	if (Tree_is_time_mode(h->tlk->tree)) Tree_update_heights(h->tlk->tree);
	return h;
}

/* the NON-VIRTUAL SingleTreeLikelihood_update_uppers (treelikelihood.c:1530-1538): _calculate_simple + update_upper_partials, both of
 * which drive the struct slots tlk->update_partials / integrate_partials / node_log_likelihoods themselves */
double refh_update_uppers(void *vh) {
	RefH *h = (RefH *)vh;
	SingleTreeLikelihood_update_all_nodes(h->tlk);
	h->tlk->m->need_update = true;
	SingleTreeLikelihood_update_uppers(h->tlk);
	h->tlk->use_upper = false;
	return h->tlk->lk;
}

/* the inner step of asr_marginal (asr.c:60-69): per-pattern log likelihood with node `id` pinned to `state` -- the node's lower
 * partials are masked IN tlk->partials, pushed through the calculate_per_cat_partials / integrate_partials / node_log_likelihoods
 * slots, and restored.  Upper partials must be current (refh_update_uppers, or a backend's sync).  out [P]. */
void refh_pinned_state_pattern_lnl(void *vh, int id, int state, double *out) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	const int S = tlk->m->nstate, P = tlk->sp->count;
	const size_t n = tlk->partials_size;
	double *partials = tlk->partials[tlk->current_partials_indexes[id]][id];
	double *backup = (double *)malloc(sizeof(double) * n);
	double *spare = (double *)malloc(sizeof(double) * n);
	double *root = (double *)malloc(sizeof(double) * (size_t)P * S);
	memcpy(backup, partials, sizeof(double) * n);
	for (size_t i = 0; i < n; i++)
		if ((int)(i % S) != state) partials[i] = 0;
	if (id == Node_id(Tree_root(tlk->tree))) memcpy(spare, partials, sizeof(double) * n);
	else tlk->calculate_per_cat_partials(tlk, spare, tlk->upper_partial_indexes[id], id, id);
	if (tlk->sm->integrate) tlk->integrate_partials(tlk, spare, tlk->sm->get_proportions(tlk->sm), root);
	else memcpy(root, spare, sizeof(double) * (size_t)P * S);
	tlk->node_log_likelihoods(tlk, root, tlk->get_root_frequencies(tlk), out);
	memcpy(partials, backup, sizeof(double) * n);
	free(backup), free(spare), free(root);
}

/* one call of the update_partials slot, as the tripod optimisation of SPR issues it (spropt.c:1578-1608): the upper partial of
 * `id` from its parent's upper partial and its sibling's lower partial, after the sibling's length changed */
void refh_slot_update_upper(void *vh, int id) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	Node *node = Tree_node(tlk->tree, id);
	Node *parent = Node_parent(node), *sib = Node_sibling(node);
	const int N = Tree_node_count(tlk->tree);
	SingleTreeLikelihood_update_Q(tlk, sib);
	if (Node_isroot(parent)) tlk->update_partials(tlk, id + N, Node_id(sib), Node_id(sib), -1, -1);
	else tlk->update_partials(tlk, id + N, Node_id(parent) + N, Node_id(parent), Node_id(sib), Node_id(sib));
}

void refh_set_distance(void *vh, int id, double d) {
	SingleTreeLikelihood *tlk = ((RefH *)vh)->tlk;
	Node_set_distance(Tree_node(tlk->tree, id), d);
}

void refh_store(void *vh) { ((RefH *)vh)->model->store(((RefH *)vh)->model); }
void refh_restore(void *vh) { ((RefH *)vh)->model->restore(((RefH *)vh)->model); }
double refh_plain_logP(void *vh) { return ((RefH *)vh)->model->logP(((RefH *)vh)->model); } /* no dirty marking: what a driver calls */
