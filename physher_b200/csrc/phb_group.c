/*
 * phb_group.c -- one tree likelihood sharded over several GPUs from ONE host process and ONE host thread.
 *
 * physher is a single-process C program (SURVEY.md 8b threading, 8e); the pattern-sharded scheme of DESIGN.md 5 therefore also
 * exists below the torch.distributed layer, as plain C on top of the C ABI: shard g owns the contiguous pattern range
 * [P g / G, P (g + 1) / G) on its own device, every model input is broadcast to all shards, an evaluation is LAUNCHED on
 * every shard before any result is collected (the shards' streams run concurrently).  Reduction of the G raw result vectors:
 *   NCCL (default when the shards sit on distinct devices): one communicator per device from ncclCommInitAll, every shard's
 *        [lnL, grad[N], inf flag] all-reduced in place on the shard's OWN stream -- ordered behind its kernels by the stream, no host
 *        wait, no host sum -- all G calls inside one ncclGroupStart / ncclGroupEnd; only shard 0's copy travels to the host;
 *   host (a device listed twice, no NCCL, or PHB_GROUP_REDUCE_HOST: the tests' cross-check): N + 1 doubles per shard summed on the
 *        host in shard order.
 * What needs the REDUCED lnL is applied after the sum, exactly as in csrc/phb_treelikelihood.c for one device: +-inf switches rescaling on on every shard and recomputes
 * (treelikelihood.c:1496-1519), NaN / inf fills the gradient with NaN (:328-332), the unrooted convention zeroes the root's right
 * child (:3249-3255).
 *
 * Only public entry points of include/physher_b200.h are used here.
 */
#include "../../include/physher_b200.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct phb_group {
	int G, T, N, S, C, P, root, root_right, use_tip_states;
	phb_tlk **shard;
	int *begin; /* [G + 1] pattern range edges */
	int unrooted, scale;
	double lk;
	double *gradient, *scratch; /* [N] / [N + 2], owned */
	int reduce;                 /* PHB_GROUP_REDUCE_* */
	phb_comm **comms;           /* [G] when NCCL serves this device list */
	int *devices;
	double **bufs;              /* [G] scratch: per-shard device operands of one all-reduce */
	void **streams;
};

int phb_internal_fail(int code, const char *msg); /* phb_treelikelihood.c: sets phb_last_error() */
int phb_internal_launch_packed(phb_tlk *t, int want_gradient, double **dev_buf, void **stream);
int phb_internal_collect_packed(phb_tlk *t, double *host);
int phb_internal_comms_init_all(int n, const int *devices, phb_comm **comms); /* phb_nccl.c */
int phb_internal_allreduce_all(int n, phb_comm **comms, double **bufs, size_t count, void **streams);

static int group_fail(const char *what) { return phb_internal_fail(PHB_EINVAL, what); }

phb_group *phb_group_create(int nshards, const int *devices, int ntips, int nstate, int ncat, int npatterns, const int *left, const int *right,
                            int root, int use_tip_states) {
	if (nshards < 1 || !devices || npatterns < nshards || !left || !right) {
		group_fail("phb_group_create: need nshards >= 1, a device list, a topology and at least one pattern per shard");
		return NULL;
	}
	phb_group *g = (phb_group *)calloc(1, sizeof(phb_group));
	if (!g) return NULL;
	g->G = nshards, g->T = ntips, g->N = 2 * ntips - 1, g->S = nstate, g->C = ncat, g->P = npatterns, g->root = root;
	g->use_tip_states = use_tip_states != 0;
	g->unrooted = 1;
	g->lk = NAN;
	g->shard = (phb_tlk **)calloc(nshards, sizeof(phb_tlk *));
	g->begin = (int *)malloc(sizeof(int) * (nshards + 1));
	g->gradient = (double *)calloc(g->N, sizeof(double));
	g->scratch = (double *)calloc((size_t)g->N + 2, sizeof(double));
	g->devices = (int *)malloc(sizeof(int) * nshards);
	g->bufs = (double **)calloc(nshards, sizeof(double *));
	g->streams = (void **)calloc(nshards, sizeof(void *));
	if (!g->shard || !g->begin || !g->gradient || !g->scratch || !g->devices || !g->bufs || !g->streams || root < 0 || root >= g->N) {
		phb_group_free(g);
		phb_internal_fail(root < 0 || root >= 2 * ntips - 1 ? PHB_EINVAL : PHB_ENOMEM, "phb_group_create: bad root or out of memory");
		return NULL;
	}
	g->root_right = right[root];
	for (int s = 0; s <= nshards; s++) g->begin[s] = (int)(((long long)npatterns * s) / nshards);
	memcpy(g->devices, devices, sizeof(int) * nshards);
	for (int s = 0; s < nshards; s++) {
		g->shard[s] = phb_tlk_create(ntips, nstate, ncat, g->begin[s + 1] - g->begin[s], left, right, root, use_tip_states, devices[s]);
		if (!g->shard[s]) { /* phb_last_error() holds the reason */
			phb_group_free(g);
			return NULL;
		}
	}
	{ /* NCCL when the shards sit on distinct devices and the library is loadable (bound only then); else the host sum */
		int distinct = nshards > 1;
		for (int a = 0; a < nshards && distinct; a++)
			for (int b = a + 1; b < nshards; b++)
				if (devices[a] == devices[b]) distinct = 0;
		if (distinct && phb_nccl_version() > 0) phb_group_set_reduction(g, PHB_GROUP_REDUCE_NCCL);
	}
	return g;
}

int phb_group_set_reduction(phb_group *g, int how) {
	if (how == PHB_GROUP_REDUCE_HOST) {
		g->reduce = PHB_GROUP_REDUCE_HOST;
		return PHB_OK;
	}
	if (how != PHB_GROUP_REDUCE_NCCL) return group_fail("phb_group_set_reduction: unknown reduction");
	if (!g->comms) {
		for (int a = 0; a < g->G; a++)
			for (int b = a + 1; b < g->G; b++)
				if (g->devices[a] == g->devices[b]) return phb_internal_fail(PHB_ESTATE, "phb_group: NCCL needs one distinct device per shard");
		phb_comm **comms = (phb_comm **)calloc(g->G, sizeof(phb_comm *));
		if (!comms) return phb_internal_fail(PHB_ENOMEM, "out of memory");
		const int rc = phb_internal_comms_init_all(g->G, g->devices, comms);
		if (rc) {
			for (int s = 0; s < g->G; s++) phb_comm_free(comms[s]);
			free(comms);
			return rc;
		}
		g->comms = comms;
	}
	g->reduce = PHB_GROUP_REDUCE_NCCL;
	return PHB_OK;
}

int phb_group_reduction(const phb_group *g) { return g->reduce; }

void phb_group_free(phb_group *g) {
	if (!g) return;
	if (g->shard)
		for (int s = 0; s < g->G; s++)
			if (g->shard[s]) phb_tlk_free(g->shard[s]);
	if (g->comms)
		for (int s = 0; s < g->G; s++) phb_comm_free(g->comms[s]);
	free(g->comms), free(g->devices), free(g->bufs), free(g->streams);
	free(g->shard), free(g->begin), free(g->gradient), free(g->scratch);
	free(g);
}

int phb_group_size(const phb_group *g) { return g->G; }

phb_tlk *phb_group_shard(phb_group *g, int shard) { return (shard >= 0 && shard < g->G) ? g->shard[shard] : NULL; }

int phb_group_shard_range(const phb_group *g, int shard, int *begin, int *end) {
	if (shard < 0 || shard >= g->G) return group_fail("phb_group_shard_range: shard out of range");
	if (begin) *begin = g->begin[shard];
	if (end) *end = g->begin[shard + 1];
	return PHB_OK;
}

/* ---- pattern-indexed inputs are sliced, everything else is broadcast ------------------------------------------------------- */

int phb_group_set_tip_states(phb_group *g, const uint8_t *states) {
	for (int s = 0; s < g->G; s++) {
		const int b = g->begin[s], n = g->begin[s + 1] - b;
		uint8_t *slice = (uint8_t *)malloc((size_t)g->T * n);
		if (!slice) return PHB_ENOMEM;
		for (int t = 0; t < g->T; t++) memcpy(slice + (size_t)t * n, states + (size_t)t * g->P + b, n);
		const int rc = phb_tlk_set_tip_states(g->shard[s], slice);
		free(slice);
		if (rc) return rc;
	}
	return PHB_OK;
}

int phb_group_set_tip_partials(phb_group *g, const double *partials) {
	const size_t S = g->S;
	for (int s = 0; s < g->G; s++) {
		const int b = g->begin[s], n = g->begin[s + 1] - b;
		double *slice = (double *)malloc(sizeof(double) * (size_t)g->T * n * S);
		if (!slice) return PHB_ENOMEM;
		for (int t = 0; t < g->T; t++) memcpy(slice + (size_t)t * n * S, partials + ((size_t)t * g->P + b) * S, sizeof(double) * n * S);
		const int rc = phb_tlk_set_tip_partials(g->shard[s], slice);
		free(slice);
		if (rc) return rc;
	}
	return PHB_OK;
}

int phb_group_set_pattern_weights(phb_group *g, const double *weights) {
	for (int s = 0; s < g->G; s++) {
		const int rc = phb_tlk_set_pattern_weights(g->shard[s], weights + g->begin[s]);
		if (rc) return rc;
	}
	return PHB_OK;
}

#define BROADCAST(call)                         \
	do {                                        \
		for (int s_ = 0; s_ < g->G; s_++) {     \
			phb_tlk *t = g->shard[s_];          \
			const int rc_ = (call);             \
			if (rc_) return rc_;                \
		}                                       \
		return PHB_OK;                          \
	} while (0)

int phb_group_set_eigen(phb_group *g, const double *evec, const double *eval, const double *ivec) { BROADCAST(phb_tlk_set_eigen(t, evec, eval, ivec)); }
int phb_group_set_frequencies(phb_group *g, const double *freqs) { BROADCAST(phb_tlk_set_frequencies(t, freqs)); }
int phb_group_set_site_model(phb_group *g, const double *rates, const double *props) { BROADCAST(phb_tlk_set_site_model(t, rates, props)); }
int phb_group_set_branch_lengths(phb_group *g, const double *bl) { BROADCAST(phb_tlk_set_branch_lengths(t, bl)); }

int phb_group_set_option(phb_group *g, int option, int value) {
	if (option == PHB_OPT_UNROOTED) { /* applied after the reduction: the shards return raw sums */
		g->unrooted = value != 0;
		return PHB_OK;
	}
	if (option == PHB_OPT_INCREMENTAL && value) return group_fail("phb_group_set_option: resident partials (PHB_OPT_INCREMENTAL) are a single-device mode");
	BROADCAST(phb_tlk_set_option(t, option, value));
}

int phb_group_use_rescaling(phb_group *g, int use) {
	g->scale = use != 0;
	BROADCAST(phb_tlk_use_rescaling(t, use));
}

int phb_group_rescaling(const phb_group *g) { return g->scale; }

/* ---- evaluation -------------------------------------------------------------------------------------------------------------- */

/* launch on every shard, then collect and sum in shard order */
static int group_evaluate(phb_group *g, int want_gradient, double *lnl) {
	for (int attempt = 0; attempt < 2; attempt++) {
		int rc;
		for (int s = 0; s < g->G; s++)
			if ((rc = phb_tlk_evaluate_launch(g->shard[s], want_gradient))) return rc;
		double total = 0.0;
		if (want_gradient) memset(g->gradient, 0, sizeof(double) * g->N);
		for (int s = 0; s < g->G; s++) {
			double part = 0.0;
			if ((rc = phb_tlk_evaluate_collect(g->shard[s], &part, want_gradient ? g->scratch : NULL))) return rc;
			total += part;
			if (want_gradient)
				for (int n = 0; n < g->N; n++) g->gradient[n] += g->scratch[n];
		}
		*lnl = total;
		if (isinf(total) && !g->scale) { /* a -inf shard makes the reduced lnL -inf: every shard switches (treelikelihood.c:1496-1519) */
			fprintf(stdout, "_calculate: rescaling %f\n", total);
			if ((rc = phb_group_use_rescaling(g, 1))) return rc;
			continue;
		}
		break;
	}
	return PHB_OK;
}

int phb_group_calculate(phb_group *g, double *lnl) {
	if (!lnl) return group_fail("phb_group_calculate: lnl is required");
	const int rc = group_evaluate(g, 0, &g->lk);
	*lnl = g->lk;
	return rc;
}

int phb_group_gradient(phb_group *g, double *lnl, const double **grad) {
	if (!grad) return group_fail("phb_group_gradient: grad is required");
	const int rc = group_evaluate(g, 1, &g->lk);
	if (rc) return rc;
	if (isnan(g->lk) || isinf(g->lk)) {
		for (int n = 0; n < g->N; n++) g->gradient[n] = NAN; /* treelikelihood.c:328-332 */
	} else {
		g->gradient[g->root] = 0.0;
		if (g->unrooted) g->gradient[g->root_right] = 0.0; /* treelikelihood.c:3249-3255 */
	}
	if (lnl) *lnl = g->lk;
	*grad = g->gradient;
	return PHB_OK;
}
