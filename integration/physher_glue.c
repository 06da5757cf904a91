/*
 * physher_glue.c -- the reference-side binding a physher maintainer adds to run the tree likelihood on
 * libphysher_b200.so.  This is the ONLY translation unit that includes the reference's
 * headers; it contains no arithmetic on partials, only the plumbing between physher's objects and the C ABI
 * of include/physher_b200.h.
 *
 *   phb_physher_attach(model, device)   after new_TreeLikelihoodModel[_from_json] (treelikelihood.c:126-128):
 *                                       builds a phb_tlk from Tree / SitePattern / SubstitutionModel / SiteModel /
 *                                       BranchModel, re-points tlk->calculate (the slot every caller goes through,
 *                                       _singleTreeLikelihood_logP treelikelihood.c:163) and wraps model->dlogP / model->free
 *   phb_physher_gradient(model)         contract of TreeLikelihood_gradient (treelikelihood.c:320-340)
 *   phb_physher_detach(model)           restores the function pointers and frees the device object
 *
 * Everything else -- JSON parsing, listeners and dirty flags, eigendecomposition, gamma quantiles, node heights and the
 * chain rule to ratios / clock rate (gradient_ratios, gradient_clock, treelikelihood.c:3054-3171) -- is the reference's
 * own code, called unchanged.
 *
 * Build (oracle/Makefile `glue`): gcc -I/root/reference/src -Iinclude integration/physher_glue.c
 *                                 -Loracle/_ref -lphyc_ref -Lphysher_b200 -lphysher_b200
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "phyc/branchmodel.h"
#include "phyc/mjson.h"
#include "phyc/parameters.h"
#include "phyc/sitemodel.h"
#include "phyc/sitepattern.h"
#include "phyc/substmodel.h"
#include "phyc/tree.h"
#include "phyc/treelikelihood.h"

#include "physher_b200.h"

/* non-static helpers of treelikelihood.c that its header does not list */
extern void gradient_ratios(SingleTreeLikelihood *tlk, const double *branch_gradient, double *gradient);
extern void gradient_clock(SingleTreeLikelihood *tlk, const double *branch_gradient, double *gradient);
extern void gradient_shape_W_sitemodel(SingleTreeLikelihood *tlk, const double *branch_gradient, const double *branch_lengths, double *gradient);
extern void central_finite_differences_simplex(Model *model, Simplex *simplex, double epsilon, double *gradient);
extern void central_finite_differences_parameters(Model *model, Parameters *parameters, double epsilon, double *gradient);
extern void update_eigen_system(SubstitutionModel *m);

typedef struct Backend {
	struct Backend *next;
	Model *model;
	SingleTreeLikelihood *tlk;
	phb_tlk *h;
	double (*ref_calculate)(SingleTreeLikelihood *);
	double (*ref_dlogP)(Model *, const Parameter *);
	void (*ref_free)(Model *);
	int N, S, C;
	double *bl, *rates, *props, *freqs, *evec, *eval, *ivec, *P, *dP, *branch_gradient, *cat_gradient, *site_bl;
	int have_model; /* the substitution model has been pushed at least once */
	int incremental; /* the device object keeps its partials resident (single-branch fast path in use) */
	double (*ref_d2logP)(Model *, const Parameter *);
	Model *(*ref_clone)(Model *, Hashtable *);
	int device;
	long long evaluations;
	/* the struct slots callers drive themselves (treelikelihood.h:90-94,111) */
	void (*ref_update_partials)(SingleTreeLikelihood *, int, int, int, int, int);
	void (*ref_integrate_partials)(const SingleTreeLikelihood *, const double *, const double *, double *);
	void (*ref_node_log_likelihoods)(const SingleTreeLikelihood *, const double *, const double *, double *);
	void (*ref_calculate_per_cat_partials)(SingleTreeLikelihood *, double *, int, int, int);
	int host_partials_current; /* tlk->partials / tlk->matrices on the host mirror the device (phb_physher_sync_partials_to_host) */
	void (*ref_store)(Model *);
	void (*ref_restore)(Model *);
	int *left, *right, root; /* topology as pushed last */
	/* inputs as they were at model->store (MCMC reject returns to them) */
	int has_store, st_root;
	double *st_bl, *st_rates, *st_props, *st_freqs, *st_evec, *st_eval, *st_ivec;
	int *st_left, *st_right;
	double st_lk;
	int st_clean;
} Backend;

static long long g_total_evaluations = 0; /* device evaluations through every attached model of this process (introspection) */
static Backend *g_backends = NULL; /* one host thread per SingleTreeLikelihood, like the reference (SURVEY.md 8b threading) */

static Backend *backend_of_tlk(const SingleTreeLikelihood *tlk) {
	for (Backend *b = g_backends; b; b = b->next)
		if (b->tlk == tlk) return b;
	return NULL;
}

static void die(const char *what) { /* the reference's error convention: message + exit (treelikelihood.c:1099-1100) */
	fprintf(stderr, "physher_b200: %s: %s\n", what, phb_last_error());
	exit(1);
}

static int same(const double *a, const double *b, size_t n) { return memcmp(a, b, n * sizeof(double)) == 0; }

/* Push whatever changed since the last evaluation.  The reference reads all of this through raw pointers on every
 * evaluation (_calculate_partials, treelikelihood.c:1645-1734); here each piece is compared with the last pushed copy so
 * that an unchanged substitution model costs nothing and does not dirty the device object. */
static void sync_inputs(Backend *b) {
	SingleTreeLikelihood *tlk = b->tlk;
	const int N = b->N, S = b->S, C = b->C;
	const int time_mode = Tree_is_time_mode(tlk->tree);
	double *tmp = (double *)malloc(sizeof(double) * (size_t)(2 * S * S + 2 * S + 2 * C + N));
	double *evec = tmp, *ivec = evec + S * S, *eval = ivec + S * S, *freqs = eval + S, *rates = freqs + S, *props = rates + C, *bl = props + C;
	/* topology: the reference re-reads its Tree on every traversal, so an NNI / SPR move (nniopt.c:301-334, spropt.c:1548-1615) needs no
	 * call of its own there; here the traversal is compiled into schedules and a changed tree is handed over */
	{
		int changed = 0, root = -1;
		for (int i = 0; i < N; i++) {
			Node *n = Tree_node(tlk->tree, i);
			const int id = Node_id(n);
			const int l = n->left ? Node_id(n->left) : -1, r = n->right ? Node_id(n->right) : -1;
			if (Node_isroot(n)) root = id;
			if (l != b->left[id] || r != b->right[id]) changed = 1;
			b->left[id] = l, b->right[id] = r;
		}
		if (changed || root != b->root) {
			b->root = root;
			if (phb_tlk_set_topology(b->h, b->left, b->right, root)) die("set_topology");
		}
	}
	/* branch lengths as _calculate_partials reads them (treelikelihood.c:1652-1663) */
	for (int i = 0; i < N; i++) {
		Node *n = Tree_node(tlk->tree, i);
		const int id = Node_id(n);
		if (Node_isroot(n)) bl[id] = 0.0;
		else if (tlk->bm == NULL || !time_mode) bl[id] = Node_distance(n);
		else bl[id] = tlk->bm->get(tlk->bm, n) * Node_time_elapsed(n);
	}
	if (!same(bl, b->bl, N) || b->evaluations == 0) {
		if (phb_tlk_set_branch_lengths(b->h, bl)) { /* negative length: the reference prints and exits (:1659-1662) */
			fprintf(stderr, "%s\n", phb_last_error());
			exit(1);
		}
		memcpy(b->bl, bl, sizeof(double) * N);
	}
	/* site model: sm->get_rate already includes mu (sitemodel.c:544-549) */
	double *p = tlk->sm->get_proportions(tlk->sm);
	for (int c = 0; c < C; c++) {
		rates[c] = tlk->sm->get_rate(tlk->sm, c);
		props[c] = p[c];
	}
	if (!same(rates, b->rates, C) || !same(props, b->props, C) || b->evaluations == 0) {
		if (phb_tlk_set_site_model(b->h, rates, props)) die("set_site_model");
		memcpy(b->rates, rates, sizeof(double) * C);
		memcpy(b->props, props, sizeof(double) * C);
	}
	/* root frequencies (tlk->get_root_frequencies, treelikelihood.c:1946-1953) */
	memcpy(freqs, tlk->get_root_frequencies(tlk), sizeof(double) * S);
	if (!same(freqs, b->freqs, S) || b->evaluations == 0) {
		if (phb_tlk_set_frequencies(b->h, freqs)) die("set_frequencies");
		memcpy(b->freqs, freqs, sizeof(double) * S);
	}
	/* substitution model: the host-side eigen system (substmodel.c:518-557), or closed-form matrices (jc69.c:73, hky.c:230) */
	SubstitutionModel *m = tlk->m;
	/* JC69 and F81 (both carry modeltype JC69: jc69.c:31, f81.c:33) have closed-form p_t / dp_dt and never fill m->eigendcmp.  Their
	 * rate matrix is Q = beta (1 pi^T - I), beta = 1 / (1 - sum pi^2) (f81.c:45-61, 77; pi = 1/4 gives jc69.c:73-94), whose eigen system
	 * is known in closed form: eigenvalues (0, -beta, -beta, -beta), V = [1 | e_k - pi_k 1], V^-1 = [pi ; e_k - e_S].  Handing that
	 * over keeps these models on the eigen path: matrices are built on the device (no 2 N C host p_t calls + upload per evaluation),
	 * the fused 4-state walk serves them, and the single-branch path can rebuild P, P', P" at a candidate length. */
	const int f81_like = m->modeltype == JC69 && S == 4;
	const int has_eigen = f81_like || m->eigendcmp != NULL;
	if (has_eigen) {
		if (f81_like) {
			const double *pi = m->get_frequencies(m);
			double ss = 0.0;
			for (int i = 0; i < S; i++) ss += pi[i] * pi[i];
			const double beta = 1.0 / (1.0 - ss);
			memset(evec, 0, sizeof(double) * S * S);
			memset(ivec, 0, sizeof(double) * S * S);
			eval[0] = 0.0;
			for (int i = 0; i < S; i++) {
				evec[i * S] = 1.0;
				ivec[i] = pi[i];
			}
			for (int k = 1; k < S; k++) {
				eval[k] = -beta;
				for (int i = 0; i < S; i++) evec[i * S + k] = (i == k - 1 ? 1.0 : 0.0) - pi[k - 1];
				ivec[k * S + (k - 1)] = 1.0;
				ivec[k * S + (S - 1)] = -1.0;
			}
		} else {
			if (m->need_update) {
				m->update_Q(m);
				update_eigen_system(m);
			}
			for (int i = 0; i < S; i++) {
				eval[i] = m->eigendcmp->eval[i];
				for (int j = 0; j < S; j++) {
					evec[i * S + j] = m->eigendcmp->evec[i][j];
					ivec[i * S + j] = m->eigendcmp->Invevec[i][j];
				}
			}
		}
		if (!b->have_model || !same(evec, b->evec, (size_t)S * S) || !same(eval, b->eval, S) || !same(ivec, b->ivec, (size_t)S * S)) {
			if (phb_tlk_set_eigen(b->h, evec, eval, ivec)) die("set_eigen");
			memcpy(b->evec, evec, sizeof(double) * S * S);
			memcpy(b->eval, eval, sizeof(double) * S);
			memcpy(b->ivec, ivec, sizeof(double) * S * S);
			b->have_model = 1;
		}
	} else {
		/* O(N C S^2) host work, S = 4 here; the matrices depend on the branch lengths, so they are rebuilt whenever anything moved */
		const size_t msz = (size_t)S * S;
		if (!b->P) {
			b->P = (double *)calloc((size_t)N * C * msz, sizeof(double));
			b->dP = (double *)calloc((size_t)N * C * msz, sizeof(double));
		}
		for (int i = 0; i < N; i++) {
			Node *n = Tree_node(tlk->tree, i);
			const int id = Node_id(n);
			if (Node_isroot(n)) continue;
			for (int c = 0; c < C; c++) {
				m->p_t(m, bl[id] * rates[c], b->P + ((size_t)id * C + c) * msz);
				m->dp_dt(m, bl[id] * rates[c], b->dP + ((size_t)id * C + c) * msz);
			}
		}
		if (phb_tlk_set_matrices(b->h, b->P, b->dP)) die("set_matrices");
		b->have_model = 1;
	}
	free(tmp);
}

/* what _calculate_simple does before it touches partials (treelikelihood.c:1462-1468), then the input sync */
static int prepare(Backend *b) {
	SingleTreeLikelihood *tlk = b->tlk;
	if (!tlk->sm->update(tlk->sm)) return 0; /* :1462-1466 */
	if (Tree_is_time_mode(tlk->tree)) Tree_update_heights(tlk->tree); /* :1468 */
	sync_inputs(b);
	if (phb_tlk_rescaling(b->h) != (int)tlk->scale) phb_tlk_use_rescaling(b->h, tlk->scale);
	if (!b->incremental) phb_tlk_update_all_nodes(b->h); /* resident partials: the setters above marked exactly what moved */
	return 1;
}

/* flag handling after an evaluation (treelikelihood.c:1489-1495, 1521-1523) */
static double finish(Backend *b, double lnl) {
	SingleTreeLikelihood *tlk = b->tlk;
	b->evaluations++;
	g_total_evaluations++;
	tlk->lk = lnl;
	if (phb_tlk_rescaling(b->h) && !tlk->scale) printf("_calculate: rescaling %f\n", -INFINITY); /* the reference's own message (:1497) */
	tlk->scale = phb_tlk_rescaling(b->h) != 0; /* -inf => the device path switched rescaling on and recomputed (:1496-1519) */
	const bool bad = isnan(lnl);
	for (int i = 0; i < b->N; i++) tlk->update_nodes[i] = bad;
	tlk->update = bad;
	tlk->update_upper = true;
	return lnl;
}

/*
 * tlk->calculate while tlk->use_upper is set: the control flow of _calculate (treelikelihood.c:1552-1607).  The optimiser has
 * changed the length of ONE branch (Brent on that branch) or has just moved on to the next branch (two nodes flagged: the
 * previous one, whose final length is now pushed to the device object, and the new one).  The branch under optimisation is
 * evaluated from its upper and lower partials at its candidate length (phb_tlk_calculate_branch == _calculate_uppper, :2592-2686);
 * the device object recomputes only the partials the pushed change reaches.
 */
static double calculate_upper_mode(Backend *b) {
	SingleTreeLikelihood *tlk = b->tlk;
	if (!tlk->update) return tlk->lk; /* "no update", :1582-1587 */
	b->host_partials_current = 0;
	if (Tree_is_time_mode(tlk->tree)) {
		/* a changed node height moves three branch lengths (rate x dt of the node and of its children, :1652-1663), so "the one
		 * flagged branch" of :1564-1580 does not exist on a time tree (the reference's upper functions read Node_distance there,
		 * :2640).  All lengths are pushed and lnL comes from the resident partials: only the ancestors of what moved are recomputed. */
		if (!prepare(b)) return tlk->lk = NAN;
		double lnl = NAN;
		if (phb_tlk_calculate(b->h, &lnl)) die("calculate");
		b->evaluations++;
		for (int i = 0; i < b->N; i++) tlk->update_nodes[i] = false;
		tlk->update_upper = true;
		return tlk->lk = lnl;
	}
	const int prev = tlk->node_upper ? Node_id(tlk->node_upper) : -1;
	int count = 0, cur = -1;
	for (int i = 0; i < b->N; i++)
		if (tlk->update_nodes[i]) {
			count++;
			if (count == 1 || i != prev) cur = i; /* one flagged node: that one; more: the one that is not node_upper (:1591-1599) */
		}
	if (cur < 0) return tlk->lk;
	for (int i = 0; i < b->N; i++) { /* every other flagged branch takes its current length for good */
		if (!tlk->update_nodes[i] || i == cur) continue;
		Node *n = Tree_node(tlk->tree, i);
		const int id = Node_id(n);
		if (!Node_isroot(n) && Node_distance(n) != b->bl[id]) {
			if (phb_tlk_set_branch_length(b->h, id, Node_distance(n))) die("set_branch_length");
			b->bl[id] = Node_distance(n);
		}
		tlk->update_nodes[i] = false;
	}
	Node *node = Tree_node(tlk->tree, cur);
	const double t = Node_distance(node);
	double lnl = NAN;
	if (phb_tlk_calculate_branch(b->h, Node_id(node), 1, &t, &lnl, NULL, NULL)) die("calculate_branch");
	b->evaluations++;
	tlk->node_upper = node;
	return tlk->lk = lnl;
}

/* == tlk->calculate: control flow of _calculate_simple (treelikelihood.c:1454-1526), numerics on the device */
static double phb_physher_calculate(SingleTreeLikelihood *tlk) {
	Backend *b = backend_of_tlk(tlk);
	if (tlk->use_upper && tlk->tripod && b->incremental && tlk->node_upper != NULL && !Tree_is_time_mode(tlk->tree)) {
		/* :1553-1556, the three-branch optimisation of SPR (spropt.c:1560-1612): lnL at the current length of node_upper from its
		 * upper and lower partials, which the caller keeps current through tlk->update_partials */
		const double t = Node_distance(tlk->node_upper);
		double lnl = NAN;
		if (phb_tlk_calculate_branch(b->h, Node_id(tlk->node_upper), 1, &t, &lnl, NULL, NULL)) die("calculate_branch");
		b->evaluations++;
		return tlk->lk = lnl;
	}
	if (tlk->use_upper && b->incremental) return calculate_upper_mode(b);
	if (!tlk->update) return tlk->lk; /* :1458 */
	if (!prepare(b)) return tlk->lk = NAN;
	double lnl = NAN;
	b->host_partials_current = 0;
	if (phb_tlk_calculate(b->h, &lnl)) die("calculate");
	return finish(b, lnl);
}

/*
 * calculate_dlnl_dQ (treelikelihood.c:2337-2583) for the parameter indices [first, first + count): the per-node matrices dP/d theta
 * come from the reference's own m->dPdp (:2421, :2484), the node sweep over upper partials, lower partials and pattern likelihoods
 * runs on the device for all indices in one call (phb_tlk_matrix_gradient), and a frequency parameter adds its root term
 * sum_i (d pi_i / d theta) * d lnL / d pi_i (:2371-2404) from phb_tlk_root_frequency_gradient.
 */
static void device_dlnl_dQ(Backend *b, int first, int count, double *out) {
	SingleTreeLikelihood *tlk = b->tlk;
	SubstitutionModel *m = tlk->m;
	const int N = b->N, S = b->S, C = b->C;
	if (count <= 0) return;
	const size_t msz = (size_t)S * S, set = (size_t)N * C * msz;
	double *M = (double *)calloc((size_t)count * set, sizeof(double));
	if (!M) die("out of memory");
	for (int k = 0; k < count; k++) {
		m->dQ_need_update = true; /* :2360 */
		for (int i = 0; i < N; i++) {
			Node *n = Tree_node(tlk->tree, i);
			const int id = Node_id(n);
			if (Node_isroot(n)) continue;
			for (int c = 0; c < C; c++) m->dPdp(m, first + k, M + (size_t)k * set + ((size_t)id * C + c) * msz, b->bl[id] * b->rates[c]);
		}
	}
	if (phb_tlk_matrix_gradient(b->h, count, M, out)) die("matrix_gradient");
	free(M);
	/* frequency parameters: indices >= rateCount (:2362-2369) */
	size_t rate_count = m->rates_simplex == NULL ? Parameters_count(m->rates) : m->rates_simplex->K - 1;
	if (m->rates_simplex != NULL && !m->grad_wrt_reparam) rate_count++;
	double *G = NULL, *dphi = NULL;
	for (int k = 0; k < count; k++) {
		if ((size_t)(first + k) < rate_count) continue;
		if (tlk->get_root_frequencies(tlk) == tlk->root_frequencies) continue; /* fixed root frequencies: no root term (:2372) */
		if (!G) {
			G = (double *)calloc(S, sizeof(double));
			dphi = (double *)calloc(S, sizeof(double));
			if (phb_tlk_root_frequency_gradient(b->h, G)) die("root_frequency_gradient");
		}
		const size_t fi = (size_t)(first + k) - rate_count;
		memset(dphi, 0, sizeof(double) * S);
		if (m->grad_wrt_reparam) m->simplex->gradient(m->simplex, fi, dphi);
		else dphi[fi] = 1.0;
		double root_term = 0.0;
		for (int i = 0; i < S; i++) root_term += dphi[i] * G[i];
		out[k] += root_term;
	}
	free(G);
	free(dphi);
}

/* == TreeLikelihood_gradient (treelikelihood.c:320-340).  Supported requests: TREE_MODEL and BRANCH_MODEL (what the
 * hot path produces: branch-length gradients and everything the reference derives from them on the host).  When the
 * lower pass is stale too, lnL and the gradient come out of ONE device evaluation (the reference runs calculate() and
 * then the upper pass; the fused kernels do both in a single launch). */
double *phb_physher_gradient(Model *self) {
	SingleTreeLikelihood *tlk = (SingleTreeLikelihood *)self->obj;
	Backend *b = backend_of_tlk(tlk);
	if (!b) return TreeLikelihood_gradient(self);
	const int flags = tlk->prepared_gradient;
	const int subst_all = (flags & (TREELIKELIHOOD_FLAG_SUBSTITUTION_MODEL)) || (flags & (TREELIKELIHOOD_FLAG_SUBSTITUTION_MODEL_UNCONSTRAINED));
	const int subst_rates = subst_all || (flags & (TREELIKELIHOOD_FLAG_SUBSTITUTION_MODEL_RATES));
	const int subst_freqs = subst_all || (flags & (TREELIKELIHOOD_FLAG_SUBSTITUTION_MODEL_FREQUENCIES));
	Model **models = (Model **)self->data;
	/* which route the reference takes for the substitution-model block (TreeLikelihood_calculate_gradient, treelikelihood.c:3309-3355):
	 * analytic through m->dPdp (calculate_dlnl_dQ; here one device sweep for all indices), or differences of logP -- which run on
	 * the device too, through tlk->calculate, by the reference's own central_finite_differences_* / Model_first_derivative */
	const int no_dPdp = tlk->m->dPdp == NULL || tlk->m->modeltype == NONREVERSIBLE;
	const double fd_eps = models[1]->epsilon;
	const int rates_fd = subst_rates && (no_dPdp || (!subst_all && fd_eps > 0.0));
	const int freqs_fd = subst_freqs && ((subst_all && no_dPdp) || (!subst_all && fd_eps > 0.0));
	if (tlk->update_upper) {
		const int time_mode = Tree_is_time_mode(tlk->tree);
		double lnl = NAN;
		const double *g = NULL;
		int ok = 1;
		if (tlk->update) ok = prepare(b);
		b->host_partials_current = 0;
		if (ok) {
			phb_tlk_set_option(b->h, PHB_OPT_INCLUDE_ROOT_FREQS, tlk->include_root_freqs);
			phb_tlk_set_option(b->h, PHB_OPT_COMPAT_SCALED_GRADIENT, 1); /* a drop-in reproduces the reference's scaled form too */
			phb_tlk_set_option(b->h, PHB_OPT_UNROOTED, !time_mode);      /* :3249-3255 */
			if (phb_tlk_gradient(b->h, &g)) die("gradient");
			if (phb_tlk_calculate(b->h, &lnl)) die("calculate"); /* cached by the gradient call unless lnL was NaN */
		}
		if (tlk->update) finish(b, lnl);
		if (isnan(lnl) || isinf(lnl)) { /* :328-332 */
			for (size_t i = 0; i < tlk->gradient_length; i++) tlk->gradient[i] = NAN;
			return tlk->gradient;
		}
		memcpy(b->branch_gradient, g, sizeof(double) * b->N);
		if (tlk->sm->mu != NULL) {
			/* the reference collapses the categories with sm->cat_rates, which do NOT carry mu (gradient_branch_length_from_cat_inplace,
			 * :3129-3143, against sm->get_rate, sitemodel.c:544-549): what it calls the branch gradient is d lnL / d (mu bl) -- the tree,
			 * clock and mu blocks below are all built from that array, so a drop-in hands over the same thing */
			const double mu = Parameter_value(tlk->sm->mu);
			for (int i = 0; i < b->N; i++) b->branch_gradient[i] /= mu;
		}
		/* from here on the reference's own host code (TreeLikelihood_calculate_gradient, :3268-3309) */
		size_t offset = 0;
		if (flags & (TREELIKELIHOOD_FLAG_TREE_MODEL)) {
			if (time_mode) {
				gradient_ratios(tlk, b->branch_gradient, tlk->gradient);
				offset += Tree_tip_count(tlk->tree) - 1;
			} else {
				memcpy(tlk->gradient, b->branch_gradient, sizeof(double) * b->N);
				offset += b->N;
			}
		}
		if (flags & (TREELIKELIHOOD_FLAG_SITE_MODEL)) {
			/* the device hands over cat_branch_gradient [N][C] (gradient_cat_branch_lengths, :2793); the chain through the rate
			 * quantiles is the reference's own gradient_discrete_sitemodel / sm->derivative (:3034-3052, sitemodel.c:258-434) */
			const int N = b->N, C = b->C;
			if (!b->cat_gradient) {
				b->cat_gradient = (double *)calloc((size_t)N * C, sizeof(double));
				b->site_bl = (double *)calloc(N, sizeof(double));
			}
			if (phb_tlk_cat_branch_gradient(b->h, b->cat_gradient)) die("cat_branch_gradient");
			for (int i = 0; i < N; i++) { /* branch lengths as :3228-3243 builds them; b->bl already carries rate * dt or the distance */
				Node *n = Tree_node(tlk->tree, i);
				b->site_bl[Node_id(n)] = Node_isroot(n) ? 0.0 : b->bl[Node_id(n)];
			}
			if (!time_mode) b->site_bl[Node_id(Tree_root(tlk->tree)->right)] = 0.0; /* :3249-3255 */
			if (tlk->sm->mu != NULL) { /* :3245-3249 */
				const double mu = Parameter_value(tlk->sm->mu);
				for (int i = 0; i < N; i++) b->site_bl[i] *= mu;
			}
			if (C > 1) {
				/* gradient_discrete_sitemodel (:3034-3052).  The shape term needs only cat_branch_gradient (the reference's own
				 * gradient_shape_W_sitemodel); the invariant-site term also needs sum_k w_k sum_i pi_i (R_0 - R_j)[k,i] / L_k over the
				 * root partials (:2958-2972, :2988-3000), which the device hands over per category (phb_tlk_category_gradient) */
				if (Parameters_count(tlk->sm->rates) == 1) gradient_shape_W_sitemodel(tlk, b->cat_gradient, b->site_bl, &tlk->gradient[offset++]);
				if (tlk->sm->proportions != NULL) {
					double *A = (double *)calloc(C, sizeof(double)), *dg = (double *)calloc(C, sizeof(double));
					if (phb_tlk_category_gradient(b->h, A)) die("category_gradient");
					if (Parameters_count(tlk->sm->rates) == 0) { /* gradient_pinv_sitemodel: invariant + one variable category */
						for (int i = 0; i < N; i++) dg[1] += b->cat_gradient[(size_t)i * 2 + 1] * b->site_bl[i];
						dg[0] = A[0] - A[1];
					} else { /* gradient_pinv_W_sitemodel */
						double mean = 0.0;
						for (int j = 1; j < C; j++) mean += A[j];
						dg[0] = A[0] - mean / (C - 1);
						for (int i = 0; i < N; i++)
							for (int j = 1; j < C; j++) dg[j] += b->cat_gradient[(size_t)i * C + j] * b->site_bl[i];
					}
					tlk->gradient[offset++] = tlk->sm->derivative(tlk->sm, dg, Parameters_at(tlk->sm->proportions->parameters, 0));
					free(A);
					free(dg);
				}
			}
			if (tlk->sm->mu != NULL) { /* :3283-3296: sum of branch gradient x branch length (the reference's convention) */
				double grad = 0.0;
				for (int i = 0; i < N; i++) grad += b->branch_gradient[i] * b->bl[i];
				tlk->gradient[offset++] = grad;
			}
		}
		if ((flags & (TREELIKELIHOOD_FLAG_BRANCH_MODEL)) && time_mode) {
			gradient_clock(tlk, b->branch_gradient, tlk->gradient + offset);
			offset += Parameters_count(tlk->bm->rates);
		}
		if (subst_rates || subst_freqs) {
			/* index ranges of gradient_PMatrix / _rates / _frequencies (treelikelihood.c:3077-3113) */
			SubstitutionModel *m = tlk->m;
			size_t nrate = m->rates_simplex == NULL ? Parameters_count(m->rates) : m->rates_simplex->K;
			if (m->rates_simplex != NULL && m->grad_wrt_reparam) nrate--;
			size_t nfreq = 0;
			if (m->simplex != NULL) {
				nfreq = m->simplex->K;
				if (m->grad_wrt_reparam) nfreq--;
			}
			/* analytic indices in one device sweep (calculate_dlnl_dQ numbering: rates first, then frequencies) */
			const int a_rates = subst_rates && !rates_fd, a_freqs = subst_freqs && !freqs_fd;
			const size_t first = a_rates ? 0 : nrate;
			const size_t count = (a_rates ? nrate : 0) + (a_freqs ? nfreq : 0);
			double *analytic = (double *)calloc(count > 0 ? count : 1, sizeof(double));
			if (count > 0) device_dlnl_dQ(b, (int)first, (int)count, analytic);
			if (subst_rates) {
				if (!rates_fd) memcpy(tlk->gradient + offset, analytic, sizeof(double) * nrate);
				else if (no_dPdp) { /* :3318-3323 */
					Parameters *params = m->rates_simplex == NULL ? m->rates : m->rates_simplex->parameters;
					for (size_t i = 0; i < Parameters_count(params); i++) tlk->gradient[offset + i] = Model_first_derivative(self, Parameters_at(params, i), 0.000001);
				} else if (!m->grad_wrt_reparam && m->rates_simplex != NULL) central_finite_differences_simplex(self, m->rates_simplex, fd_eps, tlk->gradient + offset); /* :3326-3328 */
				else central_finite_differences_parameters(self, m->rates_simplex == NULL ? m->rates : m->rates_simplex->parameters, fd_eps, tlk->gradient + offset);
				offset += nrate;
			}
			if (subst_freqs) {
				if (!freqs_fd) memcpy(tlk->gradient + offset, analytic + (a_rates ? nrate : 0), sizeof(double) * nfreq);
				else if (!m->grad_wrt_reparam) central_finite_differences_simplex(self, m->simplex, fd_eps > 0.0 ? fd_eps : 0.000001, tlk->gradient + offset); /* :3344-3346 */
				else central_finite_differences_parameters(self, m->simplex->parameters, fd_eps > 0.0 ? fd_eps : 0.000001, tlk->gradient + offset);
				offset += nfreq;
			}
			free(analytic);
			if (rates_fd || freqs_fd) SingleTreeLikelihood_update_all_nodes(tlk); /* the differences went through tlk->calculate: lnL and flags are the perturbed ones */
		}
		tlk->update_upper = false;
	}
	return tlk->gradient;
}

/* == model->dlogP: fill tlk->gradient on the device when stale, then let the reference's own index lookup
 * (_singleTreeLikelihood_dlogP_prepared, treelikelihood.c:342-447) pick the entry -- it skips its CPU recomputation
 * because update_upper is false by then. */
static double phb_physher_dlogP(Model *self, const Parameter *p) {
	SingleTreeLikelihood *tlk = (SingleTreeLikelihood *)self->obj;
	Backend *b = backend_of_tlk(tlk);
	if (tlk->gradient_length != 0 && tlk->update_upper) {
		phb_physher_gradient(self);
		if (tlk->update_upper) return tlk->lk; /* NaN / inf lnL: the reference returns it (:354-356) */
	}
	return b->ref_dlogP(self, p);
}

/* == SingleTreeLikelihood_update_uppers (treelikelihood.c:1530-1538), the call that opens serial_brent_optimize_tree
 * (optimizer.c:125), NNI and SPR: lnL, every upper partial, use_upper on.  A non-virtual function in the reference, so a drop-in
 * build forwards it here when a backend is attached. */
void phb_physher_update_uppers(Model *model) {
	SingleTreeLikelihood *tlk = (SingleTreeLikelihood *)model->obj;
	Backend *b = backend_of_tlk(tlk);
	if (!b) {
		SingleTreeLikelihood_update_uppers(tlk);
		return;
	}
	if (!b->incremental) {
		b->incremental = 1;
		if (phb_tlk_set_option(b->h, PHB_OPT_INCREMENTAL, 1)) die("set_option");
	}
	double lnl = NAN;
	if (!prepare(b)) die("update_uppers: site model");
	if (phb_tlk_update_uppers(b->h)) die("update_uppers");
	if (phb_tlk_calculate(b->h, &lnl)) die("calculate"); /* cached */
	finish(b, lnl);
	b->host_partials_current = 0;
	tlk->update_upper = false;
	tlk->use_upper = true;
	tlk->node_upper = NULL;
}

/* == model->d2logP for a branch length (_singleTreeLikelihood_d2logP, treelikelihood.c:470-527): calculate_dldt_uppper +
 * d2lnldt2_uppper at the current length, from the device's upper / lower partials of that branch */
static double phb_physher_d2logP(Model *self, const Parameter *p) {
	SingleTreeLikelihood *tlk = (SingleTreeLikelihood *)self->obj;
	Backend *b = backend_of_tlk(tlk);
	Node *node = NULL;
	for (int i = 0; i < b->N; i++) {
		Node *n = Tree_node(tlk->tree, i);
		if (n->distance && strcmp(n->distance->name, Parameter_name(p)) == 0) {
			node = n;
			break;
		}
	}
	if (!node || Node_isroot(node) || Tree_is_time_mode(tlk->tree)) return b->ref_d2logP(self, p); /* finite differences (:485-487) */
	if (!b->incremental) {
		b->incremental = 1;
		if (phb_tlk_set_option(b->h, PHB_OPT_INCREMENTAL, 1)) die("set_option");
	}
	if (tlk->update) {
		double lnl = NAN;
		if (!prepare(b)) return NAN;
		if (phb_tlk_calculate(b->h, &lnl)) die("calculate");
		finish(b, lnl);
		if (isnan(lnl) || isinf(lnl)) return lnl; /* :501-503 */
	}
	const double t = Node_distance(node);
	double lnl, d1, d2;
	if (phb_tlk_calculate_branch(b->h, Node_id(node), 1, &t, &lnl, &d1, &d2)) die("calculate_branch");
	if (isnan(d2)) SingleTreeLikelihood_update_all_nodes(tlk); /* :520-522 */
	return d2;
}


/* ------------------------------------------------------------------------------------------------------------------------------ */
/* the struct slots callers drive themselves (treelikelihood.h:90-94, 111)                                                        */
/* ------------------------------------------------------------------------------------------------------------------------------ */

/*
 * tlk->update_partials(tlk, out, p1, m1, p2, m2): the slot behind update_upper_partials[2] (treelikelihood.c:2129-2190, i.e. the
 * non-virtual SingleTreeLikelihood_update_uppers[2]), the tripod optimisation of SPR (spropt.c:1578-1608) and _calculate_partials
 * itself.  One partial update by index on the device's resident buffers with the CURRENT branch lengths, and a write-through of the
 * result into the reference's own host buffer: code that reads tlk->partials directly afterwards (asr.c:60-69, _calculate_uppper
 * :2664-2672) sees what the device computed, and the remaining three slots -- pure functions of host pointers -- stay valid.
 */
static void phb_physher_update_partials(SingleTreeLikelihood *tlk, int out, int p1, int m1, int p2, int m2) {
	Backend *b = backend_of_tlk(tlk);
	if (tlk->scale) {
		fprintf(stderr, "physher_b200: tlk->update_partials driven by the caller under rescaling is not on the device path "
		                "(SingleTreeLikelihood_scalePartials rescales the host copy only)\n");
		exit(2);
	}
	if (!b->incremental) {
		b->incremental = 1;
		if (phb_tlk_set_option(b->h, PHB_OPT_INCREMENTAL, 1)) die("set_option");
	}
	sync_inputs(b); /* the caller has just changed a length and refreshed ITS matrices (SingleTreeLikelihood_update_Q, :1614-1643) */
	double *mirror = tlk->partials[tlk->current_partials_indexes[out]][out];
	if (phb_tlk_update_partials(b->h, out, p1, m1, p2, m2, mirror)) die("update_partials");
	b->evaluations++;
}

/*
 * phb_physher_sync_partials_to_host: every lower and upper partial, the transition matrices and the pattern likelihoods from the
 * device into the reference's own buffers (tlk->partials, tlk->matrices, tlk->pattern_lk).  For reference code that reads those
 * buffers directly after SingleTreeLikelihood_update_uppers -- asr_marginal (asr.c:40-110) zeroes states in tlk->partials and pushes
 * the result through calculate_per_cat_partials / integrate_partials / node_log_likelihoods.  Switches resident partials on.
 */
void phb_physher_sync_partials_to_host(Model *model) {
	SingleTreeLikelihood *tlk = (SingleTreeLikelihood *)model->obj;
	Backend *b = backend_of_tlk(tlk);
	if (!b) return;
	const int N = b->N, S = b->S, C = b->C;
	if (!b->incremental) {
		b->incremental = 1;
		if (phb_tlk_set_option(b->h, PHB_OPT_INCREMENTAL, 1)) die("set_option");
	}
	if (!prepare(b)) die("sync_partials_to_host: site model");
	if (phb_tlk_update_uppers(b->h)) die("update_uppers");
	double lnl = NAN;
	if (phb_tlk_calculate(b->h, &lnl)) die("calculate");
	finish(b, lnl);
	for (int i = 0; i < N; i++) {
		Node *n = Tree_node(tlk->tree, i);
		const int id = Node_id(n);
		double *lower = tlk->partials[tlk->current_partials_indexes[id]][id];
		if (lower && !Node_isleaf(n) && phb_tlk_get_partials(b->h, id, lower)) die("get_partials");
		if (!Node_isroot(n)) {
			tlk->upper_partial_indexes[id] = id + N; /* :2131 */
			double *upper = tlk->partials[tlk->current_partials_indexes[id + N]][id + N];
			if (phb_tlk_get_partials(b->h, id + N, upper)) die("get_partials");
		}
	}
	{ /* matrices as the device built them, in the reference's layout [category][S*S]; SSE mode keeps a tip's matrix transposed (:1676-1690) */
		double *P = (double *)malloc(sizeof(double) * (size_t)N * C * S * S);
		if (phb_tlk_get_matrices(b->h, P, NULL)) die("get_matrices");
		for (int i = 0; i < N; i++) {
			Node *n = Tree_node(tlk->tree, i);
			const int id = Node_id(n);
			if (Node_isroot(n)) continue;
			double *dst = tlk->matrices[tlk->current_matrices_indexes[id]][id];
			bool transposed = false;
#if defined(SSE3_ENABLED) || defined(AVX_ENABLED)
			transposed = tlk->use_SIMD && tlk->partials[0][id] == NULL;
#endif
			for (int c = 0; c < C; c++)
				for (int x = 0; x < S; x++)
					for (int y = 0; y < S; y++)
						dst[(size_t)c * tlk->matrix_size + (transposed ? y * S + x : x * S + y)] = P[(((size_t)id * C + c) * S + x) * S + y];
		}
		free(P);
	}
	if (phb_tlk_pattern_log_likelihoods(b->h, tlk->pattern_lk)) die("pattern_log_likelihoods");
	tlk->update_upper = false;
	tlk->use_upper = true;
	tlk->node_upper = NULL;
	b->host_partials_current = 1;
}

static void require_host_mirror(Backend *b, const char *slot) {
	if (b->host_partials_current) return;
	/* a fused device evaluation keeps no partials on the host: bring them over instead of computing on stale buffers */
	(void)slot;
	phb_physher_sync_partials_to_host(b->model);
}

static void phb_physher_integrate_partials(const SingleTreeLikelihood *tlk, const double *in, const double *props, double *out) {
	Backend *b = backend_of_tlk(tlk);
	require_host_mirror(b, "integrate_partials");
	b->ref_integrate_partials(tlk, in, props, out);
}

static void phb_physher_node_log_likelihoods(const SingleTreeLikelihood *tlk, const double *partials, const double *freqs, double *out) {
	Backend *b = backend_of_tlk(tlk);
	require_host_mirror(b, "node_log_likelihoods");
	b->ref_node_log_likelihoods(tlk, partials, freqs, out);
}

static void phb_physher_calculate_per_cat_partials(SingleTreeLikelihood *tlk, double *root_partials, int upper_index, int partial_index, int matrix_index) {
	Backend *b = backend_of_tlk(tlk);
	require_host_mirror(b, "calculate_per_cat_partials");
	b->ref_calculate_per_cat_partials(tlk, root_partials, upper_index, partial_index, matrix_index);
}

/* ------------------------------------------------------------------------------------------------------------------------------ */
/* Model.store / Model.restore (MCMC accept / reject, _singleTreeLikelihood_store / _restore treelikelihood.c:126-161)            */
/* ------------------------------------------------------------------------------------------------------------------------------ */

static void phb_physher_store(Model *self) {
	SingleTreeLikelihood *tlk = (SingleTreeLikelihood *)self->obj;
	Backend *b = backend_of_tlk(tlk);
	b->ref_store(self); /* the sub-models' own stores, tlk->stored_lk */
	const size_t N = b->N, S = b->S, C = b->C;
	if (!b->st_bl) {
		b->st_bl = (double *)calloc(N, sizeof(double));
		b->st_rates = (double *)calloc(C, sizeof(double));
		b->st_props = (double *)calloc(C, sizeof(double));
		b->st_freqs = (double *)calloc(S, sizeof(double));
		b->st_evec = (double *)calloc(S * S, sizeof(double));
		b->st_ivec = (double *)calloc(S * S, sizeof(double));
		b->st_eval = (double *)calloc(S, sizeof(double));
		b->st_left = (int *)calloc(N, sizeof(int));
		b->st_right = (int *)calloc(N, sizeof(int));
	}
	memcpy(b->st_bl, b->bl, sizeof(double) * N);
	memcpy(b->st_rates, b->rates, sizeof(double) * C);
	memcpy(b->st_props, b->props, sizeof(double) * C);
	memcpy(b->st_freqs, b->freqs, sizeof(double) * S);
	memcpy(b->st_evec, b->evec, sizeof(double) * S * S);
	memcpy(b->st_ivec, b->ivec, sizeof(double) * S * S);
	memcpy(b->st_eval, b->eval, sizeof(double) * S);
	memcpy(b->st_left, b->left, sizeof(int) * N);
	memcpy(b->st_right, b->right, sizeof(int) * N);
	b->st_root = b->root;
	b->st_lk = tlk->lk;
	b->st_clean = !tlk->update && b->evaluations > 0 && b->P == NULL; /* the device object holds exactly these inputs and their lnL */
	if (phb_tlk_store(b->h)) die("store");
	b->has_store = 1;
}

/* After the sub-models have restored their parameters the tree likelihood is a function of the STORED inputs again.  The reference
 * leaves tlk->update set and recomputes the dirty nodes on the next logP; here the device object returns to the stored inputs and
 * their lnL (phb_tlk_restore, no recomputation) whenever what the models now hold is bit for bit what was stored. */
static void phb_physher_restore(Model *self) {
	SingleTreeLikelihood *tlk = (SingleTreeLikelihood *)self->obj;
	Backend *b = backend_of_tlk(tlk);
	b->ref_restore(self);
	if (!b->has_store || !b->st_clean || b->incremental) return; /* nothing to gain: the next logP syncs and evaluates */
	const int N = b->N, S = b->S, C = b->C;
	/* what the models hold now, gathered exactly as sync_inputs would push it */
	int same_inputs = b->root == b->st_root;
	const int time_mode = Tree_is_time_mode(tlk->tree);
	if (time_mode) Tree_update_heights(tlk->tree);
	for (int i = 0; i < N && same_inputs; i++) {
		Node *n = Tree_node(tlk->tree, i);
		const int id = Node_id(n);
		const double bl = Node_isroot(n) ? 0.0 : ((tlk->bm == NULL || !time_mode) ? Node_distance(n) : tlk->bm->get(tlk->bm, n) * Node_time_elapsed(n));
		const int l = n->left ? Node_id(n->left) : -1, r = n->right ? Node_id(n->right) : -1;
		if (bl != b->st_bl[id] || l != b->st_left[id] || r != b->st_right[id]) same_inputs = 0;
	}
	if (same_inputs && tlk->sm->update(tlk->sm)) {
		const double *p = tlk->sm->get_proportions(tlk->sm);
		for (int c = 0; c < C; c++)
			if (tlk->sm->get_rate(tlk->sm, c) != b->st_rates[c] || p[c] != b->st_props[c]) same_inputs = 0;
	} else same_inputs = 0;
	if (same_inputs && !same(tlk->get_root_frequencies(tlk), b->st_freqs, S)) same_inputs = 0;
	if (same_inputs && tlk->m->need_update) same_inputs = 0; /* the eigen system would have to be rebuilt to be compared */
	if (!same_inputs) return;
	if (phb_tlk_restore(b->h)) die("restore");
	memcpy(b->bl, b->st_bl, sizeof(double) * N);
	memcpy(b->rates, b->st_rates, sizeof(double) * C);
	memcpy(b->props, b->st_props, sizeof(double) * C);
	memcpy(b->freqs, b->st_freqs, sizeof(double) * S);
	memcpy(b->evec, b->st_evec, sizeof(double) * S * S);
	memcpy(b->ivec, b->st_ivec, sizeof(double) * S * S);
	memcpy(b->eval, b->st_eval, sizeof(double) * S);
	tlk->lk = b->st_lk;
	for (int i = 0; i < N; i++) tlk->update_nodes[i] = false;
	tlk->update = false;
	tlk->update_upper = true;
}

int phb_physher_detach(Model *model);

static void phb_physher_free(Model *self) {
	Backend *b = backend_of_tlk((SingleTreeLikelihood *)self->obj);
	void (*ref_free)(Model *) = b->ref_free;
	if (self->ref_count == 1) phb_physher_detach(self);
	ref_free(self);
}

static int attach_with(Model *model, int device, phb_tlk *h);

/*
 * Model.clone of the tree likelihood (_treeLikelihood_model_clone, treelikelihood.c:715-790): the reference clones its model graph
 * and copies the function pointers of the source object (clone_SingleTreeLikelihood_with, :1310), so a clone of an attached model
 * must get a device object of its own -- phb_tlk_clone copies the data device to device -- before anybody evaluates it.
 */
static Model *phb_physher_clone(Model *self, Hashtable *hash) {
	Backend *b = backend_of_tlk((SingleTreeLikelihood *)self->obj);
	Model *c = b->ref_clone(self, hash);
	SingleTreeLikelihood *ctlk = (SingleTreeLikelihood *)c->obj;
	ctlk->calculate = b->ref_calculate; /* what attach_with records as the reference's own function */
	ctlk->update_partials = b->ref_update_partials;
	ctlk->integrate_partials = b->ref_integrate_partials;
	ctlk->node_log_likelihoods = b->ref_node_log_likelihoods;
	ctlk->calculate_per_cat_partials = b->ref_calculate_per_cat_partials;
	c->store = b->ref_store;
	c->restore = b->ref_restore;
	c->dlogP = b->ref_dlogP;
	c->d2logP = b->ref_d2logP;
	c->free = b->ref_free;
	c->clone = b->ref_clone;
	ctlk->use_upper = false;
	phb_tlk *h = phb_tlk_clone(b->h, b->device);
	if (!h) die("phb_tlk_clone");
	if (attach_with(c, b->device, h)) die("attach clone");
	return c;
}

int phb_physher_attach(Model *model, int device) {
	SingleTreeLikelihood *tlk = (SingleTreeLikelihood *)model->obj;
	if (backend_of_tlk(tlk)) return 0;
	const int N = Tree_node_count(tlk->tree), T = Tree_tip_count(tlk->tree);
	const int S = tlk->m->nstate, C = tlk->cat_count, P = tlk->pattern_count;
	int *left = (int *)malloc(sizeof(int) * N), *right = (int *)malloc(sizeof(int) * N);
	int root = -1;
	for (int i = 0; i < N; i++) {
		Node *n = Tree_node(tlk->tree, i);
		const int id = Node_id(n);
		left[id] = n->left ? Node_id(n->left) : -1;
		right[id] = n->right ? Node_id(n->right) : -1;
		if (Node_isroot(n)) root = id;
	}
	phb_tlk *h = phb_tlk_create(T, S, C, P, left, right, root, tlk->use_tip_states, device);
	free(left);
	free(right);
	if (!h) {
		fprintf(stderr, "physher_b200: %s\n", phb_last_error());
		return -1;
	}
	/* SitePattern rows follow node ids through tlk->mapping (treelikelihood.c:1085-1117) */
	if (tlk->use_tip_states) {
		uint8_t *states = (uint8_t *)malloc((size_t)T * P);
		for (int i = 0; i < N; i++) {
			Node *n = Tree_node(tlk->tree, i);
			if (Node_isleaf(n)) memcpy(states + (size_t)Node_id(n) * P, tlk->sp->patterns[tlk->mapping[Node_id(n)]], P);
		}
		if (phb_tlk_set_tip_states(h, states)) die("set_tip_states");
		free(states);
	} else {
		double *partials = (double *)malloc(sizeof(double) * (size_t)T * P * S);
		for (int i = 0; i < N; i++) {
			Node *n = Tree_node(tlk->tree, i);
			if (Node_isleaf(n)) tlk->sp->get_partials(tlk->sp, tlk->mapping[Node_id(n)], partials + (size_t)Node_id(n) * P * S);
		}
		if (phb_tlk_set_tip_partials(h, partials)) die("set_tip_partials");
		free(partials);
	}
	if (phb_tlk_set_pattern_weights(h, tlk->sp->weights)) die("set_pattern_weights");
	return attach_with(model, device, h);
}

/* common part of attach and clone: the backend record and the re-pointed slots */
static int attach_with(Model *model, int device, phb_tlk *h) {
	SingleTreeLikelihood *tlk = (SingleTreeLikelihood *)model->obj;
	const int N = Tree_node_count(tlk->tree);
	const int S = tlk->m->nstate, C = tlk->cat_count;

	Backend *b = (Backend *)calloc(1, sizeof(Backend));
	b->model = model;
	b->tlk = tlk;
	b->h = h;
	b->N = N, b->S = S, b->C = C;
	b->bl = (double *)calloc(N, sizeof(double));
	b->branch_gradient = (double *)calloc(N, sizeof(double));
	b->rates = (double *)calloc(C, sizeof(double));
	b->props = (double *)calloc(C, sizeof(double));
	b->freqs = (double *)calloc(S, sizeof(double));
	b->evec = (double *)calloc((size_t)S * S, sizeof(double));
	b->ivec = (double *)calloc((size_t)S * S, sizeof(double));
	b->eval = (double *)calloc(S, sizeof(double));
	b->ref_calculate = tlk->calculate;
	b->ref_dlogP = model->dlogP;
	b->ref_d2logP = model->d2logP;
	b->ref_free = model->free;
	b->ref_clone = model->clone;
	b->ref_store = model->store;
	b->ref_restore = model->restore;
	b->ref_update_partials = tlk->update_partials;
	b->ref_integrate_partials = tlk->integrate_partials;
	b->ref_node_log_likelihoods = tlk->node_log_likelihoods;
	b->ref_calculate_per_cat_partials = tlk->calculate_per_cat_partials;
	b->device = device;
	b->left = (int *)malloc(sizeof(int) * N);
	b->right = (int *)malloc(sizeof(int) * N);
	b->root = -1;
	for (int i = 0; i < N; i++) {
		Node *n = Tree_node(tlk->tree, i);
		b->left[Node_id(n)] = n->left ? Node_id(n->left) : -1;
		b->right[Node_id(n)] = n->right ? Node_id(n->right) : -1;
		if (Node_isroot(n)) b->root = Node_id(n);
	}
	model->clone = phb_physher_clone;
	model->store = phb_physher_store;
	model->restore = phb_physher_restore;
	tlk->update_partials = phb_physher_update_partials;
	tlk->integrate_partials = phb_physher_integrate_partials;
	tlk->node_log_likelihoods = phb_physher_node_log_likelihoods;
	tlk->calculate_per_cat_partials = phb_physher_calculate_per_cat_partials;
	tlk->calculate = phb_physher_calculate;
	model->dlogP = phb_physher_dlogP;
	model->d2logP = phb_physher_d2logP;
	model->free = phb_physher_free;
	b->next = g_backends;
	g_backends = b;
	SingleTreeLikelihood_update_all_nodes(tlk);
	return 0;
}

int phb_physher_detach(Model *model) {
	SingleTreeLikelihood *tlk = (SingleTreeLikelihood *)model->obj;
	Backend **pp = &g_backends;
	while (*pp && (*pp)->tlk != tlk) pp = &(*pp)->next;
	Backend *b = *pp;
	if (!b) return -1;
	*pp = b->next;
	tlk->calculate = b->ref_calculate;
	model->dlogP = b->ref_dlogP;
	model->d2logP = b->ref_d2logP;
	model->free = b->ref_free;
	model->clone = b->ref_clone;
	model->store = b->ref_store;
	model->restore = b->ref_restore;
	tlk->update_partials = b->ref_update_partials;
	tlk->integrate_partials = b->ref_integrate_partials;
	tlk->node_log_likelihoods = b->ref_node_log_likelihoods;
	tlk->calculate_per_cat_partials = b->ref_calculate_per_cat_partials;
	tlk->use_upper = false;
	SingleTreeLikelihood_update_all_nodes(tlk);
	phb_tlk_free(b->h);
	free(b->bl), free(b->branch_gradient), free(b->rates), free(b->props), free(b->freqs);
	free(b->evec), free(b->ivec), free(b->eval), free(b->P), free(b->dP), free(b->cat_gradient), free(b->site_bl);
	free(b->left), free(b->right), free(b->st_left), free(b->st_right);
	free(b->st_bl), free(b->st_rates), free(b->st_props), free(b->st_freqs), free(b->st_evec), free(b->st_ivec), free(b->st_eval);
	free(b);
	return 0;
}

long long phb_physher_total_evaluations(void) { return g_total_evaluations; }

long long phb_physher_evaluations(Model *model) {
	Backend *b = backend_of_tlk((SingleTreeLikelihood *)model->obj);
	return b ? b->evaluations : -1;
}

/* the device object behind an attached model (introspection: which kernels ran, launch counts) */
phb_tlk *phb_physher_handle(Model *model) {
	Backend *b = backend_of_tlk((SingleTreeLikelihood *)model->obj);
	return b ? b->h : NULL;
}

/*
 * The JSON plugin surface: new_TreeLikelihoodModel_from_json (treelikelihood.c:819-943) with two more keys,
 *     "backend": "b200"      run the tree likelihood on libphysher_b200 (anything else, or no key: the reference's CPU path)
 *     "device":  <int>       CUDA ordinal, default 0
 * The reference rejects unknown keys (json_check_allowed, :820-832), so the two keys are taken out of the node while its own
 * constructor runs and put back afterwards; every other key -- and every "&id" reference through `hash` -- is the reference's.
 * A maintainer points the "treelikelihood" entry of the model factory (physher.c:189, compoundmodel.c:391) at this function.
 */
Model *phb_physher_new_TreeLikelihoodModel_from_json(json_node *node, Hashtable *hash) {
	const char *backend = get_json_node_value_string(node, "backend");
	const int want_device = backend != NULL && (strcmp(backend, "b200") == 0 || strcmp(backend, "cuda") == 0);
	const int device = get_json_node_value_int(node, "device", 0);
	json_node **saved = (json_node **)malloc(sizeof(json_node *) * (node->child_count > 0 ? node->child_count : 1));
	const size_t saved_count = node->child_count;
	memcpy(saved, node->children, sizeof(json_node *) * saved_count);
	size_t kept = 0;
	for (size_t i = 0; i < saved_count; i++) {
		const char *key = saved[i]->key;
		if (key != NULL && (strcmp(key, "backend") == 0 || strcmp(key, "device") == 0)) continue;
		node->children[kept++] = saved[i];
	}
	node->child_count = kept;
	Model *model = new_TreeLikelihoodModel_from_json(node, hash);
	memcpy(node->children, saved, sizeof(json_node *) * saved_count);
	node->child_count = saved_count;
	free(saved);
	if (model != NULL && want_device && phb_physher_attach(model, device) != 0) {
		fprintf(stderr, "physher_b200: \"backend\": \"%s\" requested but the device path is not available: %s\n", backend, phb_last_error());
		exit(1); /* the reference's error convention (:1099-1100); there is no silent CPU fallback */
	}
	return model;
}
