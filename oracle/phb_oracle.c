/*
 * phb_oracle.c -- plain-C restatement of physher's tree-likelihood hot path (see phb_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: never linked into or called from the product path.
 * Parity status: PINNED against tests/test_tree_likelihood.c known answers and against the
 * compiled reference (oracle/_ref), see tests/test_oracle.py.
 *
 * Layouts follow the reference: partials [cat][pattern][state] (treelikelihood.c:1028),
 * matrices [cat][i = parent state][j = child state] row-major (substmodel.c:547-555).
 */
#include "phb_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* substmodel.c:518-557 (_p_t): P = V diag(exp(lambda t)) V^-1, fabs() on every entry (:552) */
void oracle_p_t(int S, const double *evec, const double *eval, const double *ivec, double t, double *P) {
	double *pp = (double *)malloc(sizeof(double) * S * S);
	for (int i = 0; i < S; i++) {
		double e = exp(eval[i] * t);
		for (int j = 0; j < S; j++) pp[i * S + j] = ivec[i * S + j] * e;
	}
	for (int i = 0; i < S; i++) {
		for (int j = 0; j < S; j++) {
			double acc = 0.0;
			for (int k = 0; k < S; k++) acc += pp[k * S + j] * evec[i * S + k];
			P[i * S + j] = fabs(acc);
		}
	}
	free(pp);
}

/* substmodel.c:695-723 (_dp_dt): dP/dt = V diag(lambda exp(lambda t)) V^-1, no fabs */
void oracle_dp_dt(int S, const double *evec, const double *eval, const double *ivec, double t, double *dP) {
	double *pp = (double *)malloc(sizeof(double) * S * S);
	for (int i = 0; i < S; i++) {
		double e = eval[i] * exp(eval[i] * t);
		for (int j = 0; j < S; j++) pp[i * S + j] = ivec[i * S + j] * e;
	}
	for (int i = 0; i < S; i++) {
		for (int j = 0; j < S; j++) {
			double acc = 0.0;
			for (int k = 0; k < S; k++) acc += pp[k * S + j] * evec[i * S + k];
			dP[i * S + j] = acc;
		}
	}
	free(pp);
}

typedef struct Work {
	const OracleProblem *pb;
	double *mat;    /* [N][C][S][S] */
	double *dmat;   /* [N][C][S][S] */
	double *lower;  /* [N][C][P][S] */
	double *upper;  /* [N][C][P][S] */
	double *sf;     /* [2N][P] scaling factors: lower at n, upper at N+n */
	size_t psize;   /* C*P*S */
	size_t msize;   /* C*S*S */
} Work;

static int is_tip(const OracleProblem *pb, int n) { return pb->left[n] < 0; }

/* does this node own a partials buffer?  (treelikelihood.c:963-968: tips have none in state mode) */
static int has_partials(const OracleProblem *pb, int n) {
	return !is_tip(pb, n) || pb->tip_mode == ORACLE_TIP_PARTIALS;
}

/*
 * message[i] = sum_j M[c][i][j] * X[c,k,j] for one child operand, all (c,k).
 *  - partial operand: treelikelihoodX.c:502-577 (partials_undefined_and_undefined inner sums)
 *  - state operand, P matrices: treelikelihoodX.c:166-289 / treelikelihood4.c:922-992:
 *    known state -> column of M; unknown state (>= S) -> factor 1
 *  - state operand, derivative matrices (prob_matrix == 0): unknown state -> real row sums,
 *    treelikelihoodX.c:918-929 / treelikelihood4.c:1734-1746
 */
static void message(const Work *w, const double *M /*[C][S][S]*/, const double *X /*[C][P][S] or NULL*/,
                    const uint8_t *states /*[P] or NULL*/, int prob_matrix, double *out /*[C][P][S]*/) {
	const OracleProblem *pb = w->pb;
	const int S = pb->S, C = pb->C, P = pb->P;
	for (int c = 0; c < C; c++) {
		const double *Mc = M + (size_t)c * S * S;
		for (int k = 0; k < P; k++) {
			double *o = out + ((size_t)c * P + k) * S;
			if (X != NULL) {
				const double *x = X + ((size_t)c * P + k) * S;
				for (int i = 0; i < S; i++) {
					double acc = 0.0;
					for (int j = 0; j < S; j++) acc += Mc[i * S + j] * x[j];
					o[i] = acc;
				}
			} else {
				int s = states[k];
				if (s < S) {
					for (int i = 0; i < S; i++) o[i] = Mc[i * S + s];
				} else if (prob_matrix) {
					for (int i = 0; i < S; i++) o[i] = 1.0;
				} else {
					for (int i = 0; i < S; i++) {
						double acc = 0.0;
						for (int j = 0; j < S; j++) acc += Mc[i * S + j];
						o[i] = acc;
					}
				}
			}
		}
	}
}

static const double *node_partials(const Work *w, int idx /* < N lower, >= N upper */) {
	const OracleProblem *pb = w->pb;
	if (idx >= pb->N) return w->upper + (size_t)(idx - pb->N) * w->psize;
	if (!has_partials(pb, idx)) return NULL;
	return w->lower + (size_t)idx * w->psize;
}

static const uint8_t *node_states(const Work *w, int idx) {
	const OracleProblem *pb = w->pb;
	if (idx >= pb->N || has_partials(pb, idx)) return NULL;
	return pb->tip_states + (size_t)idx * pb->P;
}

/* SingleTreeLikelihood_scalePartials, treelikelihood.c:1790-1836 */
static void scale_partials(Work *w, int out_idx, int in1, int in2) {
	const OracleProblem *pb = w->pb;
	const int S = pb->S, C = pb->C, P = pb->P;
	double *p = (double *)node_partials(w, out_idx);
	double *sf = w->sf + (size_t)out_idx * P;
	const double *sf1 = (in1 >= 0 && node_partials(w, in1) != NULL) ? w->sf + (size_t)in1 * P : NULL;
	const double *sf2 = (in2 >= 0 && node_partials(w, in2) != NULL) ? w->sf + (size_t)in2 * P : NULL;
	for (int k = 0; k < P; k++) {
		double m = 0.0;
		for (int c = 0; c < C; c++)
			for (int i = 0; i < S; i++) {
				double v = p[((size_t)c * P + k) * S + i];
				if (v > m) m = v;
			}
		if (m < pb->scaling_threshold) {
			for (int c = 0; c < C; c++)
				for (int i = 0; i < S; i++) p[((size_t)c * P + k) * S + i] /= m;
			sf[k] = log(m);
		} else {
			sf[k] = 0.0;
		}
		if (sf1) sf[k] += sf1[k];
		if (sf2) sf[k] += sf2[k];
	}
}

/*
 * update_partials(out, p1, m1, p2, m2): out = (M1 x1) o (M2 x2); p2 < 0 => single child.
 * treelikelihoodX.c:43-102 (dispatch), treelikelihood4.c:1409-1467.
 */
static void update_partials(Work *w, int out_idx, int p1, int m1, int p2, int m2, double *tmp) {
	const OracleProblem *pb = w->pb;
	double *out = (double *)node_partials(w, out_idx);
	message(w, w->mat + (size_t)m1 * w->msize, node_partials(w, p1), node_states(w, p1), 1, out);
	if (p2 >= 0) {
		message(w, w->mat + (size_t)m2 * w->msize, node_partials(w, p2), node_states(w, p2), 1, tmp);
		for (size_t e = 0; e < w->psize; e++) out[e] *= tmp[e];
	}
	if (pb->scale) scale_partials(w, out_idx, p1, p2);
}

/* _calculate_partials, treelikelihood.c:1645-1734 (every node dirty) */
static void lower_pass(Work *w, int n, double *tmp) {
	const OracleProblem *pb = w->pb;
	if (is_tip(pb, n)) return;
	lower_pass(w, pb->left[n], tmp);
	lower_pass(w, pb->right[n], tmp);
	update_partials(w, n, pb->left[n], pb->left[n], pb->right[n], pb->right[n], tmp);
}

/* update_upper_partials, treelikelihood.c:2129-2161; upper index = id + N */
static void upper_pass(Work *w, int n, double *tmp) {
	const OracleProblem *pb = w->pb;
	const int N = pb->N;
	if (n != pb->root) {
		int parent = pb->parent[n];
		int sib = pb->left[parent] == n ? pb->right[parent] : pb->left[parent];
		if (parent != pb->root) {
			/* u_n = (P_p u_p) o (P_s L_s) */
			update_partials(w, N + n, N + parent, parent, sib, sib, tmp);
		} else {
			/* u_n = P_s L_s, optionally times the root frequencies (:2148-2153) */
			update_partials(w, N + n, sib, sib, -1, -1, tmp);
			if (pb->include_root_freqs) {
				double *u = w->upper + (size_t)n * w->psize;
				for (size_t e = 0; e < w->psize; e++) u[e] *= pb->freqs[e % pb->S];
			}
		}
	}
	if (!is_tip(pb, n)) {
		upper_pass(w, pb->left[n], tmp);
		upper_pass(w, pb->right[n], tmp);
	}
}

int oracle_evaluate(const OracleProblem *pb, OracleResult *res) {
	const int S = pb->S, C = pb->C, P = pb->P, N = pb->N;
	Work w;
	w.pb = pb;
	w.psize = (size_t)C * P * S;
	w.msize = (size_t)C * S * S;
	const int want_grad = res->grad != NULL || res->cat_grad != NULL || res->upper != NULL;
	w.mat = (double *)calloc((size_t)N * w.msize, sizeof(double));
	w.dmat = (double *)calloc((size_t)N * w.msize, sizeof(double));
	w.lower = (double *)calloc((size_t)N * w.psize, sizeof(double));
	w.upper = want_grad ? (double *)calloc((size_t)N * w.psize, sizeof(double)) : NULL;
	w.sf = (double *)calloc((size_t)2 * N * P, sizeof(double));
	double *tmp = (double *)malloc(sizeof(double) * w.psize);
	double *tmp2 = (double *)malloc(sizeof(double) * w.psize);
	if (!w.mat || !w.dmat || !w.lower || !w.sf || !tmp || !tmp2 || (want_grad && !w.upper)) return 1;

	/* transition matrices, treelikelihood.c:1672-1696 (t = bl * rate_c) */
	for (int n = 0; n < N; n++) {
		if (n == pb->root) continue;
		for (int c = 0; c < C; c++) {
			double *M = w.mat + (size_t)n * w.msize + (size_t)c * S * S;
			double *dM = w.dmat + (size_t)n * w.msize + (size_t)c * S * S;
			if (pb->P_override) {
				memcpy(M, pb->P_override + (size_t)n * w.msize + (size_t)c * S * S, sizeof(double) * S * S);
			} else {
				oracle_p_t(S, pb->evec, pb->eval, pb->ivec, pb->bl[n] * pb->rates[c], M);
			}
			if (pb->dP_override) {
				memcpy(dM, pb->dP_override + (size_t)n * w.msize + (size_t)c * S * S, sizeof(double) * S * S);
			} else {
				oracle_dp_dt(S, pb->evec, pb->eval, pb->ivec, pb->bl[n] * pb->rates[c], dM);
			}
		}
	}

	/* tip partials replicated over categories, treelikelihood.c:1106-1117 */
	if (pb->tip_mode == ORACLE_TIP_PARTIALS) {
		for (int t = 0; t < pb->T; t++)
			for (int c = 0; c < C; c++)
				memcpy(w.lower + (size_t)t * w.psize + (size_t)c * P * S, pb->tip_partials + (size_t)t * P * S,
				       sizeof(double) * P * S);
	}

	lower_pass(&w, pb->root, tmp);

	/* integrate_partials_general (treelikelihoodX.c:128-164), node_log_likelihoods_general (:104-126),
	 * weighted sum (treelikelihood.c:1482-1487) */
	double *plk = (double *)malloc(sizeof(double) * P);
	const double *Lr = w.lower + (size_t)pb->root * w.psize;
	double lnl = 0.0;
	for (int k = 0; k < P; k++) {
		double site = 0.0;
		for (int i = 0; i < S; i++) {
			double r = 0.0;
			for (int c = 0; c < C; c++) {
				double v = Lr[((size_t)c * P + k) * S + i];
				r += (C == 1) ? v : v * pb->props[c];
			}
			site += pb->freqs[i] * r;
		}
		plk[k] = log(site);
		if (pb->scale) plk[k] += w.sf[(size_t)pb->root * P + k]; /* getLogScalingFactor, :1838-1845 */
		lnl += plk[k] * pb->weights[k];
	}
	res->lnl = lnl;
	if (res->pattern_lnl) memcpy(res->pattern_lnl, plk, sizeof(double) * P);

	if (want_grad) {
		upper_pass(&w, pb->root, tmp);
		double *cg = (double *)calloc((size_t)N * C, sizeof(double));
		/* gradient_cat_branch_lengths, treelikelihood.c:2793-2941 and aux :2715-2789 */
		for (int n = 0; n < N; n++) {
			if (n == pb->root) continue;
			const double *U = w.upper + (size_t)n * w.psize;
			/* spare = (dP L_n) o U_n : calculate_branch_partials, treelikelihoodX.c:878-1001 */
			message(&w, w.dmat + (size_t)n * w.msize, node_partials(&w, n), node_states(&w, n), 0, tmp);
			for (size_t e = 0; e < w.psize; e++) tmp[e] *= U[e];
			if (pb->scale) {
				/* spare2 = (P L_n) o U_n  (:2722-2723); same helper => unknown states use row sums */
				message(&w, w.mat + (size_t)n * w.msize, node_partials(&w, n), node_states(&w, n), 0, tmp2);
				for (size_t e = 0; e < w.psize; e++) tmp2[e] *= U[e];
			}
			if (pb->scale && !pb->compat_scaled_gradient) {
				/* exact form under rescaling: sum_k w_k * (sum_c prop_c rate_c num[c,k]) / (sum_c prop_c den[c,k]);
				 * stored so that sum_c cg[n,c] prop_c rate_c reproduces it (cg[n,c] shares one denominator). */
				for (int k = 0; k < P; k++) {
					double den = 0.0;
					for (int c = 0; c < C; c++) {
						double d = 0.0;
						for (int i = 0; i < S; i++)
							d += (pb->include_root_freqs ? 1.0 : pb->freqs[i]) * tmp2[((size_t)c * P + k) * S + i];
						den += (C == 1 ? 1.0 : pb->props[c]) * d;
					}
					for (int c = 0; c < C; c++) {
						double num = 0.0;
						for (int i = 0; i < S; i++)
							num += (pb->include_root_freqs ? 1.0 : pb->freqs[i]) * tmp[((size_t)c * P + k) * S + i];
						cg[(size_t)n * C + c] += num / den * pb->weights[k];
					}
				}
				continue;
			}
			for (int c = 0; c < C; c++) {
				double acc = 0.0;
				for (int k = 0; k < P; k++) {
					double num = 0.0, den = 0.0;
					for (int i = 0; i < S; i++) {
						double f = pb->include_root_freqs ? 1.0 : pb->freqs[i];
						num += f * tmp[((size_t)c * P + k) * S + i];
						if (pb->scale) den += f * tmp2[((size_t)c * P + k) * S + i];
					}
					if (!pb->scale) den = exp(plk[k]); /* pattern_likelihoods, :3207-3210 */
					acc += num / den * pb->weights[k];
				}
				cg[(size_t)n * C + c] = acc;
			}
		}
		/* unrooted: the root's right child carries no gradient, treelikelihood.c:3249-3255 */
		if (pb->unrooted) {
			int r = pb->right[pb->root];
			for (int c = 0; c < C; c++) cg[(size_t)r * C + c] = 0.0;
		}
		if (res->cat_grad) memcpy(res->cat_grad, cg, sizeof(double) * N * C);
		if (res->grad) {
			/* gradient_branch_length_from_cat_inplace, :3129-3143; only applied when C > 1 (:3258-3266) */
			for (int n = 0; n < N; n++) {
				if (C == 1) {
					res->grad[n] = cg[n];
				} else {
					double g = 0.0;
					for (int c = 0; c < C; c++) g += cg[(size_t)n * C + c] * pb->props[c] * pb->rates[c];
					res->grad[n] = g;
				}
			}
		}
		free(cg);
	}

	if (res->lower) memcpy(res->lower, w.lower, sizeof(double) * N * w.psize);
	if (res->upper) memcpy(res->upper, w.upper, sizeof(double) * N * w.psize);
	if (res->matrices) memcpy(res->matrices, w.mat, sizeof(double) * N * w.msize);
	if (res->dmatrices) memcpy(res->dmatrices, w.dmat, sizeof(double) * N * w.msize);
	if (res->scaling) memcpy(res->scaling, w.sf, sizeof(double) * N * P);

	free(plk);
	free(tmp);
	free(tmp2);
	free(w.mat);
	free(w.dmat);
	free(w.lower);
	free(w.upper);
	free(w.sf);
	return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Time-tree chain.  Deliberately the reference's NAIVE forms (recursive products of ratios),
 * so that the device implementation (an adjoint sweep) is checked against an independent formulation.
 * ---------------------------------------------------------------------------------------------- */

/* tree_transform_collect_lowers (treetransform.c:239-252) */
static double tt_collect_lowers(const int *left, const int *right, const double *tip_heights, int n, double *lowers) {
	if (left[n] < 0) return lowers[n] = tip_heights[n];
	const double a = tt_collect_lowers(left, right, tip_heights, left[n], lowers);
	const double b = tt_collect_lowers(left, right, tip_heights, right[n], lowers);
	return lowers[n] = a > b ? a : b;
}

/* tree_transform_update_heights (treetransform.c:224-237) */
static void tt_update_heights(const int *left, const int *right, const int *parent, int T, int n, const double *ratios, const double *lowers,
                              double *heights) {
	if (left[n] < 0) {
		heights[n] = lowers[n];
		return;
	}
	const double s = ratios[n - T];
	if (parent[n] < 0) heights[n] = s;
	else heights[n] = lowers[n] + (heights[parent[n]] - lowers[n]) * s;
	tt_update_heights(left, right, parent, T, left[n], ratios, lowers, heights);
	tt_update_heights(left, right, parent, T, right[n], ratios, lowers, heights);
}

/* heights, branch lengths rate * (h_parent - h) (treelikelihood.c:1652-1663), log Jacobian (_node_transform_log_jacobian,
 * treetransform.c:215-222).  ratios by class id = node id - T, root entry = root height; nrates = 1 or N. */
void oracle_time_forward(int T, int N, int root, const int *left, const int *right, const int *parent, const double *tip_heights,
                         const double *ratios, const double *rates, int nrates, double *lowers, double *heights, double *bl, double *logjac) {
	tt_collect_lowers(left, right, tip_heights, root, lowers);
	tt_update_heights(left, right, parent, T, root, ratios, lowers, heights);
	double lj = 0.0;
	for (int n = 0; n < N; n++) {
		if (n == root) {
			bl[n] = 0.0;
			continue;
		}
		bl[n] = (nrates == 1 ? rates[0] : rates[n]) * (heights[parent[n]] - heights[n]);
		if (n >= T) lj += log(heights[parent[n]] - lowers[n]);
	}
	*logjac = lj;
}

/* product_of_ratios (treetransform.c:311-318) */
static void tt_product_of_ratios(const int *left, const int *right, int T, int n, const double *grad, const double *ratios, double prod,
                                 double *out) {
	if (left[n] < 0) return;
	const double p = ratios[n - T] * prod;
	*out += grad[n - T] * p;
	tt_product_of_ratios(left, right, T, left[n], grad, ratios, p, out);
	tt_product_of_ratios(left, right, T, right[n], grad, ratios, p, out);
}

/* _node_transform_dlog_jacobian_aux (treetransform.c:268-285) */
static void tt_dlog_jacobian_aux(const int *left, const int *right, const int *parent, int T, int ref, int n, const double *ratios,
                                 const double *lowers, const double *heights, double *dlogP, double *descendant) {
	if (left[n] < 0) return;
	if (parent[n] >= 0 && n != ref) descendant[n] = descendant[parent[n]] * ratios[n - T];
	else if (parent[n] >= 0) descendant[n] = heights[parent[n]] - lowers[n];
	else descendant[n] = 1;
	tt_dlog_jacobian_aux(left, right, parent, T, ref, left[n], ratios, lowers, heights, dlogP, descendant);
	tt_dlog_jacobian_aux(left, right, parent, T, ref, right[n], ratios, lowers, heights, dlogP, descendant);
	if (parent[n] >= 0 && n != ref) *dlogP += descendant[parent[n]] / (heights[parent[n]] - lowers[n]);
}

/* branch gradient -> gradient_heights (treelikelihood.c:3145-3156) -> node_transform_jvp (treetransform.c:320-337)
 * [+ _node_transform_log_jacobian_gradient (treetransform.c:297-309)], and gradient_clock (treelikelihood.c:3054-3075) */
void oracle_time_backward(int T, int N, int root, const int *left, const int *right, const int *parent, const double *ratios,
                          const double *rates, int nrates, const double *lowers, const double *heights, const double *branch_grad,
                          int include_jacobian, double *grad_ratios, double *grad_rates) {
	double *hg = (double *)calloc(T - 1, sizeof(double));
	double *descendant = (double *)calloc(N, sizeof(double));
	for (int n = 0; n < N; n++) {
		if (n == root) continue;
		const double g = branch_grad[n] * (nrates == 1 ? rates[0] : rates[n]);
		if (n >= T) hg[n - T] += -g;
		hg[parent[n] - T] += g;
	}
	for (int n = T; n < N; n++) {
		const double dh = n == root ? 1.0 : heights[parent[n]] - lowers[n];
		double acc = hg[n - T];
		tt_product_of_ratios(left, right, T, left[n], hg, ratios, 1.0, &acc);
		tt_product_of_ratios(left, right, T, right[n], hg, ratios, 1.0, &acc);
		grad_ratios[n - T] = acc * dh;
		if (include_jacobian) {
			double adj = 0.0;
			tt_dlog_jacobian_aux(left, right, parent, T, n, n, ratios, lowers, heights, &adj, descendant);
			grad_ratios[n - T] += adj;
		}
	}
	if (nrates == 1) {
		grad_rates[0] = 0.0;
		for (int n = 0; n < N; n++)
			if (n != root) grad_rates[0] += branch_grad[n] * (heights[parent[n]] - heights[n]);
	} else {
		for (int n = 0; n < N; n++) grad_rates[n] = n == root ? 0.0 : branch_grad[n] * (heights[parent[n]] - heights[n]);
	}
	free(hg);
	free(descendant);
}

/* ------------------------------------------------------------------------------------------------
 * Site-pattern compression (row A3): new_SitePattern2 / _make_patterns (sitepattern.c:186-251, 731-754) on top of the
 * reference's chained hash table (hashtable.c), restated with whole-column keys: insertion at the bucket head (:262-312),
 * growth through the prime list at load factor 0.65 with the list-reversing transfer (:199-249), iteration over buckets
 * in index order (:414-451).  alignment [T][nsites] encoded states; returns the pattern count, fills patterns [T][P] (the
 * caller provides room for P = nsites), weights [P] and site_to_pattern [nsites].
 * ---------------------------------------------------------------------------------------------- */
typedef struct PatEntry {
	unsigned hash;
	int first_site; /* key: the column at this site */
	int count;
	int next;
} PatEntry;

static unsigned pat_hash(const uint8_t *aln, int T, size_t nsites, size_t s) {
	unsigned hash = aln[s]; /* hashtable_hash_uint8_t, sitepattern.c:71-79 */
	for (int i = 1; i < T; i++) hash ^= aln[(size_t)i * nsites + s] + 0x9e3779b9 + (hash << 6) + (hash >> 2);
	unsigned i = hash; /* hashfn, hashtable.c:188-197 */
	i += ~(i << 9);
	i ^= ((i >> 14) | (i << 18));
	i += (i << 4);
	i ^= ((i >> 10) | (i << 22));
	return i;
}

static int pat_same(const uint8_t *aln, int T, size_t nsites, size_t a, size_t b) {
	for (int i = 0; i < T; i++)
		if (aln[(size_t)i * nsites + a] != aln[(size_t)i * nsites + b]) return 0;
	return 1;
}

long oracle_compress_patterns(int T, size_t nsites, const uint8_t *aln, unsigned initial_size, uint8_t *patterns, double *weights,
                              int *site_to_pattern) {
	static const unsigned primes[] = {5,        53,       97,       193,      389,       769,       1543,      3079,      6151,
	                                  12289,    24593,    49157,    98317,    196613,    393241,    786433,    1572869,   3145739,
	                                  6291469,  12582917, 25165843, 50331653, 100663319, 201326611, 402653189, 805306457, 1610612741};
	int pindex = 0;
	for (pindex = 0; primes[pindex] < initial_size; pindex++) {
	}
	unsigned size = primes[pindex];
	unsigned loadlimit = (unsigned)ceil(size * 0.65);
	int *table = (int *)malloc(sizeof(int) * size);
	PatEntry *entries = (PatEntry *)malloc(sizeof(PatEntry) * nsites);
	int *entry_of_site = (int *)malloc(sizeof(int) * nsites);
	if (!table || !entries || !entry_of_site) return -1;
	for (unsigned i = 0; i < size; i++) table[i] = -1;
	int length = 0;
	for (size_t site = 0; site < nsites; site++) { /* _make_patterns */
		const unsigned hv = pat_hash(aln, T, nsites, site);
		int e;
		for (e = table[hv % size]; e >= 0; e = entries[e].next) /* Hashtable_get_entry */
			if (entries[e].hash == hv && pat_same(aln, T, nsites, site, (size_t)entries[e].first_site)) break;
		if (e >= 0) {
			entries[e].count++;
			entry_of_site[site] = e;
			continue;
		}
		if ((unsigned)length == loadlimit) { /* Hashtable_add -> Hashtable_expand */
			const unsigned newsize = primes[++pindex];
			int *nt = (int *)malloc(sizeof(int) * newsize);
			for (unsigned i = 0; i < newsize; i++) nt[i] = -1;
			for (unsigned i = 0; i < size; i++) {
				int x;
				while ((x = table[i]) >= 0) {
					table[i] = entries[x].next;
					const unsigned idx = entries[x].hash % newsize;
					entries[x].next = nt[idx];
					nt[idx] = x;
				}
			}
			free(table);
			table = nt;
			size = newsize;
			loadlimit = (unsigned)ceil(size * 0.65);
		}
		e = length++;
		entries[e].hash = hv;
		entries[e].first_site = (int)site;
		entries[e].count = 1;
		entries[e].next = table[hv % size];
		table[hv % size] = e;
		entry_of_site[site] = e;
	}
	int *pos = (int *)malloc(sizeof(int) * (length ? length : 1));
	long index = 0;
	for (unsigned i = 0; i < size; i++) /* Hashtable_init_iterator / Hashtable_next, then sitepattern.c:229-240 */
		for (int e = table[i]; e >= 0; e = entries[e].next) {
			weights[index] = entries[e].count;
			for (int t = 0; t < T; t++) patterns[(size_t)t * length + index] = aln[(size_t)t * nsites + entries[e].first_site];
			pos[e] = (int)index;
			index++;
		}
	if (site_to_pattern)
		for (size_t site = 0; site < nsites; site++) site_to_pattern[site] = pos[entry_of_site[site]];
	free(pos), free(table), free(entries), free(entry_of_site);
	return index;
}
