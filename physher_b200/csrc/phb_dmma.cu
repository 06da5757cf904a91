// phb_dmma.cu -- FP64 tensor-core (DMMA) kernels for 20-state (amino-acid) and 61-state (codon) models.
//
// Replaces the reference's SSE paths update_partials_20_SSE (treelikelihood20.c:114-647),
// update_partials_codon_SSE (treelikelihoodCodon.c, stale) / update_partials_general_even_SSE
// (treelikelihoodX.c:1207), calculate_branch_partials_20_SSE (treelikelihood20.c:834-1025) and the
// reductions of gradient_cat_branch_lengths (treelikelihood.c:2793-2941).
//
// For one (node, category) the partial update  out[k,i] = sum_j P[i,j] x[k,j]  is the dense product
// [patterns x S] . [S x S]^T.  The reference's layouts make it an `mma.sync.m8n8k4.row.col.f64` as is:
//   A (row-major 8x4)  = 8 patterns x 4 states of a partials buffer ([pattern][state], treelikelihood.c:1028)
//   B (col-major 4x8)  = P[i][j] row-major (substmodel.c:547-555) read as B[k=j][n=i]
//   D (8x8)            = 8 patterns x 8 parent states, two adjacent states per thread -> 16-byte stores
// (tcgen05 has no f64 kind; DMMA through mma.sync is Blackwell's FP64 tensor path.)
// K is padded to a multiple of 4 and N to a multiple of 8 in the shared-memory copy of the matrices only
// (zero rows / columns); partials in HBM keep stride S.  Shared-memory rows use a leading dimension
// LD = 4 (mod 8) so the B-fragment loads hit all 16 eight-byte banks twice (the minimum for 256 bytes).
//
// Two kernels, launched per tree level on the buffers and schedules of the node-at-a-time path:
//   k_dmma_lower  K1-K4: both children's products in registers, Hadamard product, one store.
//   k_dmma_upper  K8-K10 fused per PARENT: W = P_n U_n once, M_x = P_x L_x and D_x = dP_x L_x from the
//                 same A fragments, U_a = W o M_b, U_b = W o M_a stored for internal children only,
//                 branch-gradient terms sum_i f_i U_x[i] D_x[i] * w_k / L_k reduced in the same pass.
// Rescaling reuses the generic K5 kernel between levels; under rescaling the gradient reductions (which need
// cross-category denominators) run through the generic K9/K10 kernel on the uppers this path stored.
#include "phb_ctx.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

template <int S_>
struct DmmaShape {
	static constexpr int S = S_;
	static constexpr int KP = (S + 3) / 4 * 4;  // padded contraction length
	static constexpr int NP = (S + 7) / 8 * 8;  // padded output states
	static constexpr int KT = KP / 4, NT = NP / 8;
	static constexpr int LD = (KP % 8 == 4) ? KP : KP + 4;  // leading dimension of a staged matrix, = 4 (mod 8)
	static constexpr int MAT = NP * LD;                     // doubles per staged matrix
};

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ const double *dm_partial_ptr(const Bufs &b, int idx, int c) {
	const size_t PS = (size_t)b.P * b.S;
	if (idx < b.T) return b.tip_partials + (size_t)idx * PS;
	if (idx < b.N) return b.lower + ((size_t)(idx - b.T) * b.C + c) * PS;
	return b.upper + ((size_t)(idx - b.N) * b.C + c) * PS;
}
__device__ __forceinline__ bool dm_state_tip(const Bufs &b, int idx) { return idx < b.T && b.tip_kind == PHBC_TIP_STATES; }

// stage one S x S row-major matrix into shared memory, zero padded: as [NP][LD] (B-fragment layout) for operands that are
// partials, or TRANSPOSED as [S][NP] for state tips so that one tip state selects a contiguous column of M (16-byte gathers);
// optional row sums (first NP entries of rs)
template <class Sh>
__device__ __forceinline__ void stage_matrix(double *dst, const double *__restrict__ src, double *rs, bool transposed) {
	if (transposed) {
		for (int e = threadIdx.x; e < Sh::S * Sh::NP; e += blockDim.x) {
			const int s = e / Sh::NP, i = e - s * Sh::NP;
			dst[e] = i < Sh::S ? __ldg(src + i * Sh::S + s) : 0.0;
		}
	} else {
		for (int e = threadIdx.x; e < Sh::MAT; e += blockDim.x) {
			const int i = e / Sh::LD, j = e - i * Sh::LD;
			dst[e] = (i < Sh::S && j < Sh::S) ? __ldg(src + i * Sh::S + j) : 0.0;
		}
	}
	if (rs) {
		for (int i = threadIdx.x; i < Sh::NP; i += blockDim.x) {
			double acc = 0.0;
			if (i < Sh::S)
				for (int j = 0; j < Sh::S; j++) acc += __ldg(src + i * Sh::S + j);
			rs[i] = acc;
		}
	}
}

// A fragments of MT m-tiles (8 patterns each) starting at pattern p0: a[m][t] = X[p0 + 8m + lane/4][4t + lane%4]
template <class Sh, int MT>
__device__ __forceinline__ void load_a(const double *__restrict__ X, int p0, int P, int lane, double (&a)[MT][Sh::KT]) {
	const int r = lane >> 2, q = lane & 3;
#pragma unroll
	for (int m = 0; m < MT; m++) {
		const int p = p0 + 8 * m + r;
		const double *row = X + (size_t)p * Sh::S;
#pragma unroll
		for (int t = 0; t < Sh::KT; t++) {
			const int k = 4 * t + q;
			a[m][t] = (p < P && (Sh::KP == Sh::S || k < Sh::S)) ? __ldg(row + k) : 0.0;
		}
	}
}

// acc[m][j] (+)= A . M^T for the NTW n-tiles starting at n-tile n0
template <class Sh, int MT, int NTW>
__device__ __forceinline__ void gemm(const double *__restrict__ M, int n0, int lane, const double (&a)[MT][Sh::KT], double (&acc)[MT][NTW][2]) {
	const int r = lane >> 2, q = lane & 3;
#pragma unroll
	for (int m = 0; m < MT; m++)
#pragma unroll
		for (int j = 0; j < NTW; j++) acc[m][j][0] = acc[m][j][1] = 0.0;
#pragma unroll
	for (int t = 0; t < Sh::KT; t++)
#pragma unroll
		for (int j = 0; j < NTW; j++) {
			const double bf = M[((n0 + j) * 8 + r) * Sh::LD + 4 * t + q];
#pragma unroll
			for (int m = 0; m < MT; m++) dmma_m8n8k4(acc[m][j][0], acc[m][j][1], a[m][t], bf);
		}
}

// message of a state tip: column s of M, or `unknown_value(i)` when s >= S (probability matrices: 1, treelikelihood20.c:125-131;
// derivative matrices: the row sum, treelikelihoodX.c:878-1001)
template <class Sh, int MT, int NTW, bool PROB>
__device__ __forceinline__ void tip_gather(const uint8_t *__restrict__ states, const double *__restrict__ M, const double *__restrict__ rs,
                                           int p0, int P, int n0, int lane, double (&acc)[MT][NTW][2]) {
	const int r = lane >> 2, q = lane & 3;
#pragma unroll
	for (int m = 0; m < MT; m++) {
		const int p = p0 + 8 * m + r;
		const int s = p < P ? states[p] : Sh::S;
#pragma unroll
		for (int j = 0; j < NTW; j++) {
			const int i = (n0 + j) * 8 + 2 * q;
			if (s < Sh::S) {
				const double2 v = *reinterpret_cast<const double2 *>(M + s * Sh::NP + i);  // transposed staging: M[s][i]
				acc[m][j][0] = v.x;
				acc[m][j][1] = v.y;
			} else {
				acc[m][j][0] = PROB ? 1.0 : rs[i];
				acc[m][j][1] = PROB ? 1.0 : rs[i + 1];
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// K1-K4: out = (P_a x_a) o (P_b x_b);  grid (pattern chunks, C, ops of the level)
// warps: WM m-groups x NSPLIT n-groups; a warp owns MT m-tiles and NT / NSPLIT n-tiles
// ---------------------------------------------------------------------------------------------
template <int S, int MT, int NSPLIT, int WM>
__global__ void __launch_bounds__(32 * WM * NSPLIT) k_dmma_lower(Bufs b, const phbc_op *__restrict__ ops, const double *__restrict__ Pm) {
	using Sh = DmmaShape<S>;
	constexpr int NTW = Sh::NT / NSPLIT;
	static_assert(Sh::NT % NSPLIT == 0, "n-tiles must split evenly");
	extern __shared__ double sm[];
	double *mA = sm, *mB = sm + Sh::MAT;
	const phbc_op op = ops[blockIdx.z];
	const int c = blockIdx.y;
	const bool a_tip = dm_state_tip(b, op.a), b_tip = dm_state_tip(b, op.b);
	stage_matrix<Sh>(mA, Pm + ((size_t)op.a_mat * b.C + c) * S * S, nullptr, a_tip);
	stage_matrix<Sh>(mB, Pm + ((size_t)op.b_mat * b.C + c) * S * S, nullptr, b_tip);
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int wm = warp / NSPLIT, n0 = (warp % NSPLIT) * NTW;
	const int r = lane >> 2, q = lane & 3;
	constexpr int TP = WM * MT * 8;
	double *out = (double *)dm_partial_ptr(b, op.out, c);
	const double *xa = dm_partial_ptr(b, op.a, c), *xb = dm_partial_ptr(b, op.b, c);
	const int ntiles = (b.P + TP - 1) / TP;
	for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const int p0 = tile * TP + wm * MT * 8;
		double acc[MT][NTW][2], acc2[MT][NTW][2];
		double fa[MT][Sh::KT], fb[MT][Sh::KT];
		// every global load of the tile is issued before the first product
		if (!a_tip) load_a<Sh, MT>(xa, p0, b.P, lane, fa);
		if (!b_tip) load_a<Sh, MT>(xb, p0, b.P, lane, fb);
		if (a_tip) tip_gather<Sh, MT, NTW, true>(b.tip_states + (size_t)op.a * b.P, mA, nullptr, p0, b.P, n0, lane, acc);
		else gemm<Sh, MT, NTW>(mA, n0, lane, fa, acc);
		if (b_tip) tip_gather<Sh, MT, NTW, true>(b.tip_states + (size_t)op.b * b.P, mB, nullptr, p0, b.P, n0, lane, acc2);
		else gemm<Sh, MT, NTW>(mB, n0, lane, fb, acc2);
#pragma unroll
		for (int m = 0; m < MT; m++) {
			const int p = p0 + 8 * m + r;
			if (p >= b.P) continue;
			double *row = out + (size_t)p * S;
#pragma unroll
			for (int j = 0; j < NTW; j++) {
				const int i = (n0 + j) * 8 + 2 * q;
				const double v0 = acc[m][j][0] * acc2[m][j][0], v1 = acc[m][j][1] * acc2[m][j][1];
				if (S % 2 == 0) {
					if (i < S) *reinterpret_cast<double2 *>(row + i) = make_double2(v0, v1);  // S even: 16-byte aligned
				} else {
					if (i < S) row[i] = v0;
					if (i + 1 < S) row[i + 1] = v1;
				}
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// K8-K10 fused per parent;  grid (pattern chunks, C, parent ops of the level)
// GRAD: reduce the branch-gradient terms (unscaled form, site likelihood from pattern_lnl);
// !GRAD: only the upper partials, tips included (the caller reduces with the generic K9/K10 kernel).
// partial: [N][C][gridDim.x] per-CTA sums of the children's gradient terms.
// ---------------------------------------------------------------------------------------------
template <int S, int MT, int NSPLIT, int WM, bool GRAD>
__global__ void __launch_bounds__(32 * WM * NSPLIT) k_dmma_upper(Bufs b, const phbc_parent_op *__restrict__ ops, const double *__restrict__ Pm,
                                                               const double *__restrict__ dPm, const double *__restrict__ freqs,
                                                               const double *__restrict__ weights, const double *__restrict__ pattern_lnl,
                                                               int include_root_freqs, double *__restrict__ partial) {
	using Sh = DmmaShape<S>;
	constexpr int NTW = Sh::NT / NSPLIT;
	constexpr int NWARPS = WM * NSPLIT;
	extern __shared__ double sm[];
	double *mP = sm, *mA = sm + Sh::MAT, *mB = sm + 2 * Sh::MAT;
	double *dA = GRAD ? sm + 3 * Sh::MAT : nullptr, *dB = GRAD ? sm + 4 * Sh::MAT : nullptr;
	double *aux = sm + (GRAD ? 5 : 3) * Sh::MAT;  // rsA[NP], rsB[NP], fq[NP], wroot[NP], red[2 * NWARPS]
	double *rsA = aux, *rsB = aux + Sh::NP, *fq = aux + 2 * Sh::NP, *wroot = aux + 3 * Sh::NP, *red = aux + 4 * Sh::NP;
	const phbc_parent_op op = ops[blockIdx.z];
	const int c = blockIdx.y;
	const bool is_root = op.flags & 1;
	const bool a_tip = dm_state_tip(b, op.a), b_tip = dm_state_tip(b, op.b);
	if (!is_root) stage_matrix<Sh>(mP, Pm + ((size_t)op.node * b.C + c) * S * S, nullptr, false);
	stage_matrix<Sh>(mA, Pm + ((size_t)op.a * b.C + c) * S * S, nullptr, a_tip);
	stage_matrix<Sh>(mB, Pm + ((size_t)op.b * b.C + c) * S * S, nullptr, b_tip);
	if (GRAD) {
		stage_matrix<Sh>(dA, dPm + ((size_t)op.a * b.C + c) * S * S, rsA, a_tip);
		stage_matrix<Sh>(dB, dPm + ((size_t)op.b * b.C + c) * S * S, rsB, b_tip);
	}
	for (int i = threadIdx.x; i < Sh::NP; i += blockDim.x) {
		const double f = i < S ? freqs[i] : 0.0;
		fq[i] = i < S ? (include_root_freqs ? 1.0 : f) : 0.0;     // weights of the gradient numerator
		wroot[i] = i < S ? (include_root_freqs ? f : 1.0) : 0.0;  // message entering the root's children (treelikelihood.c:2145-2154)
	}
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int wm = warp / NSPLIT, n0 = (warp % NSPLIT) * NTW;
	const int r = lane >> 2, q = lane & 3;
	constexpr int TP = WM * MT * 8;
	const bool a_leaf = op.a < b.T, b_leaf = op.b < b.T;
	const double *xw = b.upper + ((size_t)op.node * b.C + c) * (size_t)b.P * S;
	const double *xa = dm_partial_ptr(b, op.a, c), *xb = dm_partial_ptr(b, op.b, c);
	double *Ua = b.upper + ((size_t)op.a * b.C + c) * (size_t)b.P * S;
	double *Ub = b.upper + ((size_t)op.b * b.C + c) * (size_t)b.P * S;
	const int ntiles = (b.P + TP - 1) / TP;
	double tot_a = 0.0, tot_b = 0.0;

	auto store = [&](double *U, int p0, const double (&v)[MT][NTW][2]) {
#pragma unroll
		for (int m = 0; m < MT; m++) {
			const int p = p0 + 8 * m + r;
			if (p >= b.P) continue;
			double *row = U + (size_t)p * S;
#pragma unroll
			for (int j = 0; j < NTW; j++) {
				const int i = (n0 + j) * 8 + 2 * q;
				if (S % 2 == 0) {
					if (i < S) *reinterpret_cast<double2 *>(row + i) = make_double2(v[m][j][0], v[m][j][1]);
				} else {
					if (i < S) row[i] = v[m][j][0];
					if (i + 1 < S) row[i + 1] = v[m][j][1];
				}
			}
		}
	};

	for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const int p0 = tile * TP + wm * MT * 8;
		double W[MT][NTW][2];
		double fw[MT][Sh::KT], fa[MT][Sh::KT], fb[MT][Sh::KT];
		// every global load of the tile is issued before the first product
		if (!is_root) load_a<Sh, MT>(xw, p0, b.P, lane, fw);
		if (!b_tip) load_a<Sh, MT>(xb, p0, b.P, lane, fb);
		if (!a_tip) load_a<Sh, MT>(xa, p0, b.P, lane, fa);
		if (is_root) {
#pragma unroll
			for (int m = 0; m < MT; m++)
#pragma unroll
				for (int j = 0; j < NTW; j++) {
					const int i = (n0 + j) * 8 + 2 * q;
					W[m][j][0] = wroot[i], W[m][j][1] = wroot[i + 1];
				}
		} else {
			gemm<Sh, MT, NTW>(mP, n0, lane, fw, W);
		}
		// child b: M_b, D_b  ->  U_a = W o M_b (stored), Y_b = f o W o D_b (kept for g_b = sum Y_b M_a)
		double X[MT][NTW][2], Y[MT][NTW][2];
		if (b_tip) {
			const uint8_t *st = b.tip_states + (size_t)op.b * b.P;
			tip_gather<Sh, MT, NTW, true>(st, mB, nullptr, p0, b.P, n0, lane, X);
			if (GRAD) tip_gather<Sh, MT, NTW, false>(st, dB, rsB, p0, b.P, n0, lane, Y);
		} else {
			gemm<Sh, MT, NTW>(mB, n0, lane, fb, X);
			if (GRAD) gemm<Sh, MT, NTW>(dB, n0, lane, fb, Y);
		}
#pragma unroll
		for (int m = 0; m < MT; m++)
#pragma unroll
			for (int j = 0; j < NTW; j++) {
				const int i = (n0 + j) * 8 + 2 * q;
				X[m][j][0] *= W[m][j][0], X[m][j][1] *= W[m][j][1];  // X = U_a
				if (GRAD) Y[m][j][0] *= fq[i] * W[m][j][0], Y[m][j][1] *= fq[i + 1] * W[m][j][1];
			}
		if (!a_leaf || !GRAD) store(Ua, p0, X);
		// child a: M_a, D_a  ->  U_b = W o M_a (stored), g_a = sum f U_a D_a, g_b = sum Y_b M_a
		double Ma[MT][NTW][2], Da[MT][NTW][2];
		if (a_tip) {
			const uint8_t *st = b.tip_states + (size_t)op.a * b.P;
			tip_gather<Sh, MT, NTW, true>(st, mA, nullptr, p0, b.P, n0, lane, Ma);
			if (GRAD) tip_gather<Sh, MT, NTW, false>(st, dA, rsA, p0, b.P, n0, lane, Da);
		} else {
			gemm<Sh, MT, NTW>(mA, n0, lane, fa, Ma);
			if (GRAD) gemm<Sh, MT, NTW>(dA, n0, lane, fa, Da);
		}
		if (GRAD) {
#pragma unroll
			for (int m = 0; m < MT; m++) {
				double ga = 0.0, gb = 0.0;
#pragma unroll
				for (int j = 0; j < NTW; j++) {
					const int i = (n0 + j) * 8 + 2 * q;
					ga = fma(fq[i] * X[m][j][0], Da[m][j][0], fma(fq[i + 1] * X[m][j][1], Da[m][j][1], ga));
					gb = fma(Y[m][j][0], Ma[m][j][0], fma(Y[m][j][1], Ma[m][j][1], gb));
				}
				// the four lanes of a quad hold one pattern's columns
				ga += __shfl_xor_sync(0xffffffffu, ga, 1), gb += __shfl_xor_sync(0xffffffffu, gb, 1);
				ga += __shfl_xor_sync(0xffffffffu, ga, 2), gb += __shfl_xor_sync(0xffffffffu, gb, 2);
				const int p = p0 + 8 * m + r;
				const double wl = p < b.P ? weights[p] / exp(pattern_lnl[p]) : 0.0;  // w_k / L_k (treelikelihood.c:3207-3210)
				tot_a = fma(ga, wl, tot_a);
				tot_b = fma(gb, wl, tot_b);
			}
		}
#pragma unroll
		for (int m = 0; m < MT; m++)
#pragma unroll
			for (int j = 0; j < NTW; j++) Ma[m][j][0] *= W[m][j][0], Ma[m][j][1] *= W[m][j][1];  // U_b
		if (!b_leaf || !GRAD) store(Ub, p0, Ma);
	}
	if (GRAD) {
		// every quad lane carries the same row value: sum lanes with q == 0 over the 8 rows, then across warps (fixed order)
		tot_a = q == 0 ? tot_a : 0.0, tot_b = q == 0 ? tot_b : 0.0;
		tot_a = phb_warp_sum(tot_a), tot_b = phb_warp_sum(tot_b);
		if (lane == 0) red[2 * warp] = tot_a, red[2 * warp + 1] = tot_b;
		__syncthreads();
		if (threadIdx.x == 0) {
			double sa = 0.0, sb = 0.0;
			for (int w = 0; w < NWARPS; w++) sa += red[2 * w], sb += red[2 * w + 1];
			partial[((size_t)op.a * b.C + c) * gridDim.x + blockIdx.x] = sa;
			partial[((size_t)op.b * b.C + c) * gridDim.x + blockIdx.x] = sb;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int S>
struct DmmaConfig;
template <>
struct DmmaConfig<20> {  // 128 threads; 64 patterns per tile in the lower pass, 32 in the upper pass
	static constexpr int MT = 2, NSPLIT = 1, WM = 4;
	static constexpr int UMT = 1, UNSPLIT = 1, UWM = 4;
};
template <>
struct DmmaConfig<61> {  // 256 threads, 32 patterns per tile, n-tiles split over two warps
	static constexpr int MT = 1, NSPLIT = 2, WM = 4;
	static constexpr int UMT = 1, UNSPLIT = 2, UWM = 4;
};

bool phbc_dmma_supported(const phbc_ctx *ctx, const phbc_eval_opts *o) {
	(void)o;
	return ctx->S == 20 || ctx->S == 61;
}

// pattern chunks per launch: enough CTAs for ~4 waves, never more than the tile count
static int chunk_count(const phbc_ctx *ctx, int ntiles, int ctas_per_chunk) {
	int want = (4 * ctx->num_sms + ctas_per_chunk - 1) / ctas_per_chunk;
	if (want < 1) want = 1;
	return want < ntiles ? want : ntiles;
}

template <int S>
static int dmma_evaluate(phbc_ctx *ctx, const phbc_eval_opts *o) {
	using Sh = DmmaShape<S>;
	using Cf = DmmaConfig<S>;
	const int C = ctx->C, P = ctx->P, N = ctx->N;
	int rc;
	if ((rc = phbc_generic_prepare(ctx, o))) return rc;
	Bufs b = phbc_make_bufs(ctx);
	auto lower = k_dmma_lower<S, Cf::MT, Cf::NSPLIT, Cf::WM>;
	const size_t lsmem = 2 * Sh::MAT * sizeof(double);
	PHBC_CHECK(cudaFuncSetAttribute(lower, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem));
	const int lthreads = 32 * Cf::WM * Cf::NSPLIT, ltiles = (P + Cf::WM * Cf::MT * 8 - 1) / (Cf::WM * Cf::MT * 8);
	if ((rc = phbc_time_begin(ctx))) return rc;
	for (int l = 0; l < ctx->n_lower_levels; l++) {
		const int beg = ctx->h_lower_level_off[l], cnt = ctx->h_lower_level_off[l + 1] - beg;
		if (cnt <= 0) continue;
		for (int z0 = 0; z0 < cnt; z0 += 65535) {
			const int zc = cnt - z0 < 65535 ? cnt - z0 : 65535;
			lower<<<dim3(chunk_count(ctx, ltiles, C * zc), C, zc), lthreads, lsmem, ctx->stream>>>(b, ctx->d_lower_ops + beg + z0, ctx->d_P);
			ctx->launches++;
		}
		if (o->scale && (rc = phbc_generic_scale_ops(ctx, ctx->d_lower_ops + beg, cnt, o->scaling_threshold))) return rc;
	}
	double *result = ctx->d_result + (size_t)o->batch_index * (1 + N);
	if ((rc = phbc_generic_root(ctx, o, result))) return rc;
	if (o->want_gradient) {
		const bool grad = !o->scale;  // fused reductions use the unscaled form
		auto upper_g = k_dmma_upper<S, Cf::UMT, Cf::UNSPLIT, Cf::UWM, true>;
		auto upper_u = k_dmma_upper<S, Cf::UMT, Cf::UNSPLIT, Cf::UWM, false>;
		const int uwarps = Cf::UWM * Cf::UNSPLIT;
		const size_t usmem = ((grad ? 5 : 3) * Sh::MAT + 4 * Sh::NP + 2 * uwarps) * sizeof(double);
		PHBC_CHECK(cudaFuncSetAttribute(grad ? upper_g : upper_u, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)usmem));
		const int uthreads = 32 * uwarps, utiles = (P + Cf::UWM * Cf::UMT * 8 - 1) / (Cf::UWM * Cf::UMT * 8);
		// one chunk count for the whole pass: the per-CTA gradient partials are laid out [N][C][chunks]
		int chunks = chunk_count(ctx, utiles, C);
		if ((rc = phbc_ensure_scratch(ctx, (size_t)N * C * chunks * sizeof(double)))) return rc;
		for (int l = 0; l < ctx->n_upper_levels; l++) {
			const int beg = ctx->h_parent_level_off[l], cnt = ctx->h_parent_level_off[l + 1] - beg;
			if (cnt > 0) {
				for (int z0 = 0; z0 < cnt; z0 += 65535) {
					const int zc = cnt - z0 < 65535 ? cnt - z0 : 65535;
					(grad ? upper_g : upper_u)<<<dim3(chunks, C, zc), uthreads, usmem, ctx->stream>>>(
					    b, ctx->d_parent_ops + beg + z0, ctx->d_P, ctx->d_dP, ctx->d_freqs, ctx->d_weights, ctx->d_pattern_lnl, o->include_root_freqs,
					    ctx->d_scratch);
					ctx->launches++;
				}
			}
			// the children written by this level are the per-child ops of depth l + 1
			const int ubeg = ctx->h_upper_level_off[l], ucnt = ctx->h_upper_level_off[l + 1] - ubeg;
			if (o->scale && ucnt > 0 && (rc = phbc_generic_scale_ops(ctx, ctx->d_upper_ops + ubeg, ucnt, o->scaling_threshold))) return rc;
		}
		if (grad) {
			if ((rc = phbc_gradient_from_partials(ctx, chunks, result))) return rc;
		} else {
			if ((rc = phbc_generic_gradient(ctx, o, result))) return rc;
		}
	}
	if ((rc = phbc_time_end(ctx))) return rc;
	PHBC_CHECK(cudaGetLastError());
	return 0;
}

int phbc_dmma_evaluate(phbc_ctx *ctx, const phbc_eval_opts *o) {
	if (ctx->S == 20) return dmma_evaluate<20>(ctx, o);
	if (ctx->S == 61) return dmma_evaluate<61>(ctx, o);
	snprintf(phbc_errbuf, sizeof(phbc_errbuf), "tensor-core kernels are instantiated for 20 and 61 states, not %d", ctx->S);
	return -1;
}
