// phb_timetree.cu -- the time-tree chain around the likelihood, batched over samples, on the device.
//
// Replaces, for B parameter samples sharing one topology, the O(N) host recursions the reference runs around every
// evaluation of a time tree:
//   forward   ratios / root height -> node heights   (tree_transform_update_heights, treetransform.c:224-237)
//             heights, clock rates -> branch lengths (bl = rate * (h_parent - h_node), treelikelihood.c:1652-1663)
//             log |Jacobian| of the ratio transform  (_node_transform_log_jacobian, treetransform.c:215-222)
//   backward  branch gradient -> height gradient     (gradient_heights, treelikelihood.c:3145-3156)
//             height gradient -> ratio gradient      (Tree_node_transform_jvp, tree.c:2788; the adjoint sweep of
//                                                     node_transform_jvp_backprop, treetransform.c:75-93)
//             + gradient of log |Jacobian|           (_node_transform_log_jacobian_gradient_backprop, treetransform.c:95-120)
//             branch gradient -> clock-rate gradient (gradient_clock, treelikelihood.c:3054-3075)
// so that a batch (BASELINE config 3: 128 variational samples) needs one H2D of [B][T-1] ratios + rates and one D2H of
// the results instead of B host round trips.  One thread owns one sample and walks the tree in the host-provided pre-order /
// post-order; per-sample state lives in [node][sample] scratch so that neighbouring threads touch neighbouring addresses.
// Sequential per-sample sums in a fixed order: results are reproducible run to run.
#include "phb_ctx.cuh"

#include <stdlib.h>
#include <string.h>

// heights[n][b], bl[b][n] (the layout the matrix kernels read), logjac[b]
__global__ void k_time_forward(int T, int N, int B, int root, int nrates, const int *__restrict__ preorder, const int *__restrict__ parent,
                               const double *__restrict__ lowers, const double *__restrict__ ratios, const double *__restrict__ rates,
                               double *__restrict__ heights, double *__restrict__ bl, double *__restrict__ logjac, int *__restrict__ bad) {
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= B) return;
	const double *s = ratios + (size_t)b * (T - 1);  // by class id = node id - T; the root's entry is the root height
	const double *rt = rates + (size_t)b * nrates;
	double lj = 0.0;
	for (int k = 0; k < N; k++) {
		const int n = preorder[k];  // parents before children
		double h;
		if (n == root) {
			h = s[n - T];
			bl[(size_t)b * N + n] = 0.0;
		} else {
			const double hp = heights[(size_t)parent[n] * B + b];
			const double lo = lowers[n];
			if (n >= T) {
				// separately rounded like the host code of the reference (treetransform.c:224-237, compiled without FMA contraction): a height
				// that differs by an ulp gives branch lengths, and with them exponentials, whose ROUNDING differs -- and the t^2, t^3 entries
				// of a codon P(t) carry that as 1e-9 of their value
				h = __dadd_rn(lo, __dmul_rn(__dsub_rn(hp, lo), s[n - T]));
				lj += log(hp - lo);
			} else {
				h = lo;  // a tip's height is its sampling date
			}
			const double len = __dmul_rn(nrates == 1 ? rt[0] : rt[n], __dsub_rn(hp, h));  // treelikelihood.c:1652-1663
			if (len < 0.0) *bad = 1;  // the reference exits on a negative branch length (treelikelihood.c:1659-1662)
			bl[(size_t)b * N + n] = len;
		}
		heights[(size_t)n * B + b] = h;
	}
	logjac[b] = lj;
}

// result: [B][1+N] raw (lnL, d lnL / d bl by node id) of the batched evaluation
__global__ void k_time_backward(int T, int N, int B, int root, int nrates, int include_jacobian, const int *__restrict__ postorder,
                                const int *__restrict__ parent, const double *__restrict__ lowers, const double *__restrict__ ratios,
                                const double *__restrict__ rates, const double *__restrict__ heights, const double *__restrict__ result,
                                double *__restrict__ adj, double *__restrict__ grad_ratios, double *__restrict__ grad_rates) {
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= B) return;
	const double *s = ratios + (size_t)b * (T - 1);
	const double *rt = rates + (size_t)b * nrates;
	const double *g = result + (size_t)b * (1 + N) + 1;
	double *gr = grad_ratios + (size_t)b * (T - 1);
	double *gc = grad_rates + (size_t)b * nrates;
	// height gradient (gradient_heights): every non-root node pushes +g * rate to its parent and, when internal, -g * rate to itself;
	// clock gradient (gradient_clock): g * elapsed time
	for (int i = 0; i < T - 1; i++) adj[(size_t)i * B + b] = 0.0;
	double clock = 0.0;
	for (int n = 0; n < N; n++) {
		if (n == root) {
			if (nrates > 1) gc[n] = 0.0;
			continue;
		}
		const int p = parent[n];
		const double dt = heights[(size_t)p * B + b] - heights[(size_t)n * B + b];
		const double gn = g[n] * (nrates == 1 ? rt[0] : rt[n]);
		if (n >= T) adj[(size_t)(n - T) * B + b] -= gn;
		adj[(size_t)(p - T) * B + b] += gn;
		if (nrates == 1) clock += g[n] * dt;
		else gc[n] = g[n] * dt;
	}
	if (nrates == 1) gc[0] = clock;
	if (include_jacobian) {  // d log|J| / d h_parent = 1 / (h_parent - lower_n) for every internal non-root child n
		for (int k = 0; k < N; k++) {
			const int n = postorder[k];
			if (n < T || n == root) continue;
			const int p = parent[n];
			adj[(size_t)(p - T) * B + b] += 1.0 / (heights[(size_t)p * B + b] - lowers[n]);
		}
	}
	// adjoint sweep, children before parents: d h_n / d s_n = h_parent - lower_n, d h_n / d h_parent = s_n
	for (int k = 0; k < N; k++) {
		const int n = postorder[k];
		if (n < T || n == root) continue;
		const int p = parent[n];
		const double a = adj[(size_t)(n - T) * B + b];
		gr[n - T] = a * (heights[(size_t)p * B + b] - lowers[n]);
		adj[(size_t)(p - T) * B + b] += a * s[n - T];
	}
	gr[root - T] = adj[(size_t)(root - T) * B + b];
}

static int tt_grow(phbc_ctx *ctx, int B) {
	if (B <= ctx->tt_cap) return 0;
	const size_t N = ctx->N, T = ctx->T;
	PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
	double **bufs[] = {&ctx->d_tt_ratios, &ctx->d_tt_rates, &ctx->d_tt_heights, &ctx->d_tt_adj, &ctx->d_tt_out};
	const size_t sizes[] = {(size_t)B * (T - 1), (size_t)B * N, (size_t)B * N, (size_t)B * (T - 1), (size_t)B * (1 + (T - 1) + N)};
	for (int i = 0; i < 5; i++) {
		if (*bufs[i]) cudaFree(*bufs[i]);
		*bufs[i] = NULL;
		PHBC_CHECK(cudaMalloc((void **)bufs[i], sizes[i] * sizeof(double)));
	}
	if (ctx->h_tt) cudaFreeHost(ctx->h_tt);
	ctx->h_tt = NULL;
	PHBC_CHECK(cudaMallocHost((void **)&ctx->h_tt, (size_t)B * (1 + (T - 1) + 2 * N) * sizeof(double)));
	ctx->tt_cap = B;
	return 0;
}

extern "C" int phbc_set_time_tree(phbc_ctx *ctx, const double *lowers, const int *parent, const int *preorder, const int *postorder) {
	PHBC_CHECK(cudaSetDevice(ctx->device));
	const size_t N = ctx->N;
	if (!ctx->d_tt_lowers) {
		PHBC_CHECK(cudaMalloc((void **)&ctx->d_tt_lowers, N * sizeof(double)));
		PHBC_CHECK(cudaMalloc((void **)&ctx->d_tt_topo, 3 * N * sizeof(int)));
		PHBC_CHECK(cudaMalloc((void **)&ctx->d_tt_bad, sizeof(int)));
	}
	PHBC_CHECK(cudaMemcpyAsync(ctx->d_tt_lowers, lowers, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	PHBC_CHECK(cudaMemcpyAsync(ctx->d_tt_topo, parent, N * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
	PHBC_CHECK(cudaMemcpyAsync(ctx->d_tt_topo + N, preorder, N * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
	PHBC_CHECK(cudaMemcpyAsync(ctx->d_tt_topo + 2 * N, postorder, N * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
	PHBC_CHECK(cudaStreamSynchronize(ctx->stream));  // the sources are caller memory
	return 0;
}

// ratios [B][T-1], rates [B][nrates] (host) -> branch lengths of samples 0..B-1 in the batch slots, log|J| kept for the download.
// Returns 1 (and leaves the batch untouched) when some sample has a negative branch length.
extern "C" int phbc_time_forward(phbc_ctx *ctx, int B, const double *ratios, const double *rates, int nrates) {
	PHBC_CHECK(cudaSetDevice(ctx->device));
	if (!ctx->d_tt_lowers) {
		snprintf(phbc_errbuf, sizeof(phbc_errbuf), "time tree not set");
		return -4;
	}
	const size_t N = ctx->N, T = ctx->T;
	int rc;
	ctx->ex_count = 0;  // branch lengths are made on the device from here on
	if ((rc = tt_grow(ctx, B))) return rc;
	if (B > ctx->bl_cap || B > ctx->result_cap) {  // batch slots of the likelihood (same growth rule as phbc_upload_branch_lengths)
		PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_bl);
		cudaFreeHost(ctx->h_bl);
		cudaFree(ctx->d_result);
		ctx->d_bl = NULL, ctx->h_bl = NULL, ctx->d_result = NULL;
		PHBC_CHECK(cudaMalloc((void **)&ctx->d_bl, (size_t)B * N * sizeof(double)));
		PHBC_CHECK(cudaMallocHost((void **)&ctx->h_bl, (size_t)B * N * sizeof(double)));
		PHBC_CHECK(cudaMalloc((void **)&ctx->d_result, (size_t)B * (1 + N) * sizeof(double)));
		ctx->bl_cap = B;
		ctx->result_cap = B;
	}
	PHBC_CHECK(cudaStreamSynchronize(ctx->stream));  // pinned staging is free again
	memcpy(ctx->h_tt, ratios, (size_t)B * (T - 1) * sizeof(double));
	memcpy(ctx->h_tt + (size_t)B * (T - 1), rates, (size_t)B * nrates * sizeof(double));
	PHBC_CHECK(cudaMemcpyAsync(ctx->d_tt_ratios, ctx->h_tt, (size_t)B * (T - 1) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	PHBC_CHECK(cudaMemcpyAsync(ctx->d_tt_rates, ctx->h_tt + (size_t)B * (T - 1), (size_t)B * nrates * sizeof(double), cudaMemcpyHostToDevice,
	                           ctx->stream));
	PHBC_CHECK(cudaMemsetAsync(ctx->d_tt_bad, 0, sizeof(int), ctx->stream));
	const int *parent = ctx->d_tt_topo, *pre = ctx->d_tt_topo + N;
	k_time_forward<<<(B + 63) / 64, 64, 0, ctx->stream>>>((int)T, (int)N, B, ctx->root, nrates, pre, parent, ctx->d_tt_lowers, ctx->d_tt_ratios,
	                                                    ctx->d_tt_rates, ctx->d_tt_heights, ctx->d_bl, ctx->d_tt_out, ctx->d_tt_bad);
	ctx->launches++;
	int bad = 0;
	PHBC_CHECK(cudaMemcpyAsync(&bad, ctx->d_tt_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
	PHBC_CHECK(cudaGetLastError());
	return bad ? 1 : 0;
}

// after the batched evaluation: chain rule on the device, one D2H of [lnl | logjac | grad_ratios | grad_rates]
extern "C" int phbc_time_backward(phbc_ctx *ctx, int B, int nrates, int include_jacobian, int want_gradient, double *lnl, double *logjac,
                                  double *grad_ratios, double *grad_rates) {
	PHBC_CHECK(cudaSetDevice(ctx->device));
	const size_t N = ctx->N, T = ctx->T;
	double *d_logjac = ctx->d_tt_out, *d_gr = ctx->d_tt_out + B, *d_gc = d_gr + (size_t)B * (T - 1);
	if (want_gradient) {
		const int *parent = ctx->d_tt_topo, *post = ctx->d_tt_topo + 2 * N;
		k_time_backward<<<(B + 63) / 64, 64, 0, ctx->stream>>>((int)T, (int)N, B, ctx->root, nrates, include_jacobian, post, parent, ctx->d_tt_lowers,
		                                                     ctx->d_tt_ratios, ctx->d_tt_rates, ctx->d_tt_heights, ctx->d_result, ctx->d_tt_adj, d_gr,
		                                                     d_gc);
		ctx->launches++;
	}
	// lnL sits at stride 1 + N in the result slots: gather it with a strided copy
	double *h = ctx->h_tt;
	PHBC_CHECK(cudaMemcpy2DAsync(h, sizeof(double), ctx->d_result, (1 + N) * sizeof(double), sizeof(double), B, cudaMemcpyDeviceToHost, ctx->stream));
	const size_t tail = (size_t)B + (want_gradient ? (size_t)B * (T - 1) + (size_t)B * nrates : 0);
	PHBC_CHECK(cudaMemcpyAsync(h + B, d_logjac, tail * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
	PHBC_CHECK(cudaGetLastError());
	memcpy(lnl, h, B * sizeof(double));
	if (logjac) memcpy(logjac, h + B, B * sizeof(double));
	if (want_gradient) {
		if (grad_ratios) memcpy(grad_ratios, h + 2 * B, (size_t)B * (T - 1) * sizeof(double));
		if (grad_rates) memcpy(grad_rates, h + 2 * B + (size_t)B * (T - 1), (size_t)B * nrates * sizeof(double));
	}
	return 0;
}
