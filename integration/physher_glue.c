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
extern void gradient_discrete_sitemodel(SingleTreeLikelihood *tlk, const double *branch_gradient, const double *branch_lengths, double *gradient);
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
} Backend;

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
	if (Tree_is_time_mode(tlk->tree)) {
		fprintf(stderr, "physher_b200: use_upper on a time tree is not on the device path\n");
		exit(2);
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
	if (tlk->use_upper && b->incremental) return calculate_upper_mode(b);
	if (!tlk->update) return tlk->lk; /* :1458 */
	if (!prepare(b)) return tlk->lk = NAN;
	double lnl = NAN;
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
	if (subst_rates || subst_freqs) {
		/* the analytic route of the reference (calculate_dlnl_dQ on m->dPdp); its finite-difference routes (:3318-3334, :3344-3351)
		 * re-evaluate logP and would run on the device through tlk->calculate, but are not wired here */
		Model **models = (Model **)self->data;
		if (tlk->m->dPdp == NULL || tlk->m->modeltype == NONREVERSIBLE || (!subst_all && models[1]->epsilon > 0.0)) {
			fprintf(stderr, "physher_b200: substitution-model gradient by finite differences is not on the device path\n");
			exit(2);
		}
	}
	if ((flags & (TREELIKELIHOOD_FLAG_SITE_MODEL)) && (tlk->sm->proportions != NULL || tlk->sm->mu != NULL)) {
		/* gradient_pinv_sitemodel reads the CPU root partials (treelikelihood.c:2943-2975); mu rescales rates and lengths (:3245-3249) */
		fprintf(stderr, "physher_b200: invariant-site proportion and mu gradients are not on the device path\n");
		exit(2);
	}
	if (tlk->update_upper) {
		const int time_mode = Tree_is_time_mode(tlk->tree);
		double lnl = NAN;
		const double *g = NULL;
		int ok = 1;
		if (tlk->update) ok = prepare(b);
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
			double grad_sitemodel[2] = {0.0, 0.0};
			if (C > 1) gradient_discrete_sitemodel(tlk, b->cat_gradient, b->site_bl, grad_sitemodel);
			if (Parameters_count(tlk->sm->rates) == 1) tlk->gradient[offset++] = grad_sitemodel[0];
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
			const size_t first = subst_rates ? 0 : nrate;
			const size_t count = (subst_rates ? nrate : 0) + (subst_freqs ? nfreq : 0);
			device_dlnl_dQ(b, (int)first, (int)count, tlk->gradient + offset);
			offset += count;
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
	b->device = device;
	model->clone = phb_physher_clone;
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
	tlk->use_upper = false;
	SingleTreeLikelihood_update_all_nodes(tlk);
	phb_tlk_free(b->h);
	free(b->bl), free(b->branch_gradient), free(b->rates), free(b->props), free(b->freqs);
	free(b->evec), free(b->ivec), free(b->eval), free(b->P), free(b->dP), free(b->cat_gradient), free(b->site_bl);
	free(b);
	return 0;
}

long long phb_physher_evaluations(Model *model) {
	Backend *b = backend_of_tlk((SingleTreeLikelihood *)model->obj);
	return b ? b->evaluations : -1;
}
