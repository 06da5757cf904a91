// phb_nuc4.cu -- fused whole-tree walk kernels for 4-state (nucleotide) models on sm_100a.
//
// Replaces, for nstate == 4, the reference's one-node-at-a-time SSE path
// (update_partials_4_SSE treelikelihood4.c:1409, integrate_partials_4_SSE :822,
// node_log_likelihoods_4_SSE :882, update_upper_partials treelikelihood.c:2129,
// calculate_branch_partials_4_SSE treelikelihood4.c:1752, gradient_cat_branch_lengths
// treelikelihood.c:2793) by ONE kernel per evaluation:
//
//   * a thread owns one (pattern, rate category) pair and walks the WHOLE tree: post-order for the
//     lower partials, root integration, then pre-order for upper partials and branch gradients;
//   * intermediate partials live in thread-private shared-memory slots; the host orders the walks so
//     that the number of live slots is the tree's Strahler number (<= log2(T) + 1);
//   * lower partials of internal nodes are streamed once to a CTA-private HBM scratch row (32 B per
//     thread per node, coalesced 1 KB per warp) and read back once by the pre-order pass; upper
//     partials never leave the SM.  HBM traffic is ~2(T-1)*C*32 B per pattern instead of the
//     ~(5T-9)*C*32 B of node-at-a-time streaming;
//   * transition matrices arrive in walk order through TMA bulk copies (cp.async.bulk + mbarrier,
//     double-buffered chunks of 8 ops) and are read as warp-uniform broadcasts;
//   * dP/dt L is evaluated as Q (P L): dP/dt = Q P(t) for any rate matrix, so derivative matrices are
//     never built or staged; Q sits in the kernel parameter (constant) bank;
//   * per-branch gradient terms are reduced with a paired warp butterfly (two branches per 5
//     shuffles) and accumulated with no-return reductions into warp-private rows, which a second
//     tiny kernel sums in a fixed order (deterministic).
//
// CTAs are persistent: grid = min(tiles, resident CTAs), each CTA loops over pattern tiles and owns
// its scratch, so device memory is independent of the pattern count.
#include "phb_ctx.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NUC4_CHUNK 8  // ops per TMA chunk
#define NUC4_NT 256   // threads per CTA

struct Nuc4Params {
	int T, N, C, P, PB, root;
	int n_post, n_pre, nslots, ntiles;
	int include_root_freqs, compat;
	double threshold;
	const uint8_t *tip_codes;  // [T][P]
	const double *weights;
	const double *props;
	const phbc_post_op *post_ops;
	const phbc_pre_op *pre_ops;
	const double *post_mats;  // [n_post][2][C][16]
	const double *pre_mats;   // [n_pre][3][C][16]
	double *lower;            // [grid][n_post][C][PB][4]
	double *gacc;             // [grid][warps][N]
	double *cta_lnl;          // [grid]
	double *pattern_lnl;      // [P]
	double freqs[4];
	double Q[16];
};

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (global -> shared)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
	             "l"(src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	uint32_t done;
	do {
		asm volatile(
		    "{\n"
		    ".reg .pred p;\n"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
		    "selp.u32 %0, 1, 0, p;\n"
		    "}\n"
		    : "=r"(done)
		    : "r"(smem_u32(bar)), "r"(parity)
		    : "memory");
	} while (!done);
}
__device__ __forceinline__ void red_add_f64(double *addr, double v) {
	asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// small dense helpers (everything in registers)
// ---------------------------------------------------------------------------------------------
// y = M x with M a 4x4 row-major matrix in shared memory (warp-uniform address => broadcast)
__device__ __forceinline__ void matvec_smem(const double *__restrict__ M, const double (&x)[4], double (&y)[4]) {
	const double2 *M2 = reinterpret_cast<const double2 *>(M);
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const double2 a = M2[2 * i], b = M2[2 * i + 1];
		y[i] = fma(a.x, x[0], fma(a.y, x[1], fma(b.x, x[2], b.y * x[3])));
	}
}

// tip message: code bits 0-3 = set of compatible states, bit 4 = "missing, factor exactly 1"
// (state tips with state >= 4, treelikelihood4.c:1107-1150).  One-hot codes gather a column.
__device__ __forceinline__ void tip_message(const double *__restrict__ M, unsigned code, double (&y)[4]) {
	const bool simple = (code & 0x10u) || __popc(code & 0xfu) == 1;
	if (__all_sync(0xffffffffu, simple)) {
		const int s = (code & 0x10u) ? 0 : (__ffs(code) - 1);
#pragma unroll
		for (int i = 0; i < 4; i++) {
			const double v = M[4 * i + s];
			y[i] = (code & 0x10u) ? 1.0 : v;
		}
	} else {
		double x[4];
#pragma unroll
		for (int j = 0; j < 4; j++) x[j] = ((code >> j) & 1u) ? 1.0 : 0.0;
		if (code & 0x10u) x[0] = x[1] = x[2] = x[3] = 0.0;
		matvec_smem(M, x, y);
		if (code & 0x10u) y[0] = y[1] = y[2] = y[3] = 1.0;
	}
}

struct Slots {
	double2 *base;  // [slot][half][NT]
	__device__ __forceinline__ void load(int slot, double (&x)[4]) const {
		const double2 lo = base[(slot * 2 + 0) * NUC4_NT + threadIdx.x];
		const double2 hi = base[(slot * 2 + 1) * NUC4_NT + threadIdx.x];
		x[0] = lo.x, x[1] = lo.y, x[2] = hi.x, x[3] = hi.y;
	}
	__device__ __forceinline__ void store(int slot, const double (&x)[4]) const {
		base[(slot * 2 + 0) * NUC4_NT + threadIdx.x] = make_double2(x[0], x[1]);
		base[(slot * 2 + 1) * NUC4_NT + threadIdx.x] = make_double2(x[2], x[3]);
	}
};

__device__ __forceinline__ void load_lower(const double *__restrict__ p, double (&x)[4]) {
	const double2 lo = __ldcs(reinterpret_cast<const double2 *>(p));
	const double2 hi = __ldcs(reinterpret_cast<const double2 *>(p) + 1);
	x[0] = lo.x, x[1] = lo.y, x[2] = hi.x, x[3] = hi.y;
}
__device__ __forceinline__ void store_lower(double *__restrict__ p, const double (&x)[4]) {
	__stcs(reinterpret_cast<double2 *>(p), make_double2(x[0], x[1]));
	__stcs(reinterpret_cast<double2 *>(p) + 1, make_double2(x[2], x[3]));
}

// ---------------------------------------------------------------------------------------------
// the walk kernel
// ---------------------------------------------------------------------------------------------
// shared memory map (dynamic):
//   [0, 16)                      two mbarriers
//   stage[2]: each NUC4_CHUNK * (48 + 3*C*128) bytes (descriptors then matrices)
//   slots:    nslots * 2 * NT * 16 bytes
//   xch:      exchange area for cross-category sums (4 * C * PB doubles) + invLw[PB] + sfslot[nslots][NT]
template <bool SCALE, bool GRAD>
__global__ void __launch_bounds__(NUC4_NT, 2) k_nuc4_walk(const Nuc4Params prm) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int C = prm.C, PB = prm.PB;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int c = tid / PB, pl = tid - c * PB;
	const size_t stage_bytes = (size_t)NUC4_CHUNK * (48 + 3 * C * 128);
	uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
	unsigned char *stage0 = smem_raw + 128;
	Slots slots;
	slots.base = reinterpret_cast<double2 *>(stage0 + 2 * stage_bytes);
	double *xch = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(slots.base) + (size_t)prm.nslots * 2 * NUC4_NT * 16);
	double *invLw = xch + 4 * C * PB;
	double *sfslot = invLw + PB;  // [nslots][NT] thread-private copies (SCALE only)

	if (tid == 0) {
		mbar_init(&bars[0], 1);
		mbar_init(&bars[1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	uint32_t loads = 0;  // chunk loads consumed so far by this CTA (uniform): stage = loads & 1, parity = (loads >> 1) & 1
	const double prop_c = (C == 1) ? 1.0 : prm.props[c];
	double cta_lnl = 0.0;
	double *my_lower = GRAD ? prm.lower + (size_t)blockIdx.x * prm.n_post * C * PB * 4 : nullptr;
	double *my_gacc = GRAD ? prm.gacc + ((size_t)blockIdx.x * (NUC4_NT / 32) + warp) * prm.N : nullptr;

	for (int tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x) {
		const int p = tile * PB + pl;
		const bool live = p < prm.P;
		const int pc = live ? p : prm.P - 1;  // clamp loads of the ragged last tile
		const uint8_t *codes = prm.tip_codes + pc;

		// ------------------------------------------------------------------ post-order
		double out[4] = {1.0, 1.0, 1.0, 1.0};
		double sf_acc = 0.0;  // SCALE: log scaling factor of the value currently in `out`
		{
			const int nchunks = (prm.n_post + NUC4_CHUNK - 1) / NUC4_CHUNK;
			auto issue = [&](int ch, uint32_t ld) {
				const int first = ch * NUC4_CHUNK;
				const int cnt = min(NUC4_CHUNK, prm.n_post - first);
				unsigned char *dst = stage0 + (ld & 1) * stage_bytes;
				const uint32_t dbytes = cnt * (uint32_t)sizeof(phbc_post_op), mbytes = cnt * 2 * C * 128;
				mbar_expect_tx(&bars[ld & 1], dbytes + mbytes);
				bulk_g2s(dst, prm.post_ops + first, dbytes, &bars[ld & 1]);
				bulk_g2s(dst + NUC4_CHUNK * 48, prm.post_mats + (size_t)first * 2 * C * 16, mbytes, &bars[ld & 1]);
			};
			if (tid == 0) issue(0, loads);
			for (int ch = 0; ch < nchunks; ch++) {
				if (tid == 0 && ch + 1 < nchunks) issue(ch + 1, loads + 1);
				mbar_wait(&bars[loads & 1], (loads >> 1) & 1);
				const unsigned char *st = stage0 + (loads & 1) * stage_bytes;
				const phbc_post_op *desc = reinterpret_cast<const phbc_post_op *>(st);
				const double *mats = reinterpret_cast<const double *>(st + NUC4_CHUNK * 48);
				const int first = ch * NUC4_CHUNK;
				const int cnt = min(NUC4_CHUNK, prm.n_post - first);
				// prefetch the tip codes of the first op of the chunk
				unsigned codeA = 0, codeB = 0;
				if (desc[0].a_kind == PHBC_W_TIP) codeA = codes[(size_t)desc[0].a_idx * prm.P];
				if (desc[0].b_kind == PHBC_W_TIP) codeB = codes[(size_t)desc[0].b_idx * prm.P];
				for (int j = 0; j < cnt; j++) {
					const phbc_post_op d = desc[j];
					const unsigned curA = codeA, curB = codeB;
					if (j + 1 < cnt) {  // software prefetch of the next op's tip codes
						if (desc[j + 1].a_kind == PHBC_W_TIP) codeA = codes[(size_t)desc[j + 1].a_idx * prm.P];
						if (desc[j + 1].b_kind == PHBC_W_TIP) codeB = codes[(size_t)desc[j + 1].b_idx * prm.P];
					}
					const double *MA = mats + ((size_t)(j * 2 + 0) * C + c) * 16;
					const double *MB = mats + ((size_t)(j * 2 + 1) * C + c) * 16;
					double ma[4], mb[4], x[4];
					double sf_in = 0.0;
					if (d.a_kind == PHBC_W_TIP) {
						tip_message(MA, curA, ma);
					} else {
						slots.load(d.a_idx, x);
						matvec_smem(MA, x, ma);
						if (SCALE) sf_in += sfslot[d.a_idx * NUC4_NT + tid];
					}
					if (d.b_kind == PHBC_W_TIP) {
						tip_message(MB, curB, mb);
					} else {
						slots.load(d.b_idx, x);
						matvec_smem(MB, x, mb);
						if (SCALE) sf_in += sfslot[d.b_idx * NUC4_NT + tid];
					}
#pragma unroll
					for (int i = 0; i < 4; i++) out[i] = ma[i] * mb[i];
					if (SCALE) {
						// SingleTreeLikelihood_scalePartials (treelikelihood.c:1790-1836): max over categories and states
						double m = fmax(fmax(out[0], out[1]), fmax(out[2], out[3]));
						double *mx = xch + ((first + j) & 1) * C * PB;
						mx[c * PB + pl] = m;
						__syncthreads();
						m = mx[pl];
						for (int cc = 1; cc < C; cc++) m = fmax(m, mx[cc * PB + pl]);
						double sf = 0.0;
						if (m < prm.threshold) {
#pragma unroll
							for (int i = 0; i < 4; i++) out[i] /= m;
							sf = log(m);
						}
						sf_acc = sf + sf_in;
						sfslot[d.dst_slot * NUC4_NT + tid] = sf_acc;
					}
					slots.store(d.dst_slot, out);
					if (GRAD) store_lower(my_lower + (((size_t)(first + j) * C + c) * PB + pl) * 4, out);
				}
				loads++;
				__syncthreads();  // everyone is done with this stage before it is refilled
			}
		}

		// ------------------------------------------------------------------ root integration
		// integrate_partials_4_SSE + node_log_likelihoods_4_SSE + weighted sum
		{
			double site = prm.freqs[0] * out[0] + prm.freqs[1] * out[1] + prm.freqs[2] * out[2] + prm.freqs[3] * out[3];
			site *= prop_c;
			xch[c * PB + pl] = site;
			__syncthreads();
			if (c == 0) {
				double L = xch[pl];
				for (int cc = 1; cc < C; cc++) L += xch[cc * PB + pl];
				double plk = log(L);
				if (SCALE) plk += sf_acc;
				const double w = live ? prm.weights[pc] : 0.0;
				if (live) prm.pattern_lnl[p] = plk;
				invLw[pl] = w / L;  // unscaled path: w_k / L_k; scaled paths use ratios instead
				double v = live ? plk * w : 0.0;
				v = phb_warp_sum(v);
				if (lane == 0) cta_lnl += v;  // warps of category 0 only; summed below
			}
			__syncthreads();
		}

		// ------------------------------------------------------------------ pre-order + gradients
		if (GRAD) {
			const double sgrad = invLw[pl];
			const double wk = live ? prm.weights[pc] : 0.0;
			const int nchunks = (prm.n_pre + NUC4_CHUNK - 1) / NUC4_CHUNK;
			auto issue = [&](int ch, uint32_t ld) {
				const int first = ch * NUC4_CHUNK;
				const int cnt = min(NUC4_CHUNK, prm.n_pre - first);
				unsigned char *dst = stage0 + (ld & 1) * stage_bytes;
				const uint32_t dbytes = cnt * (uint32_t)sizeof(phbc_pre_op), mbytes = cnt * 3 * C * 128;
				mbar_expect_tx(&bars[ld & 1], dbytes + mbytes);
				bulk_g2s(dst, prm.pre_ops + first, dbytes, &bars[ld & 1]);
				bulk_g2s(dst + NUC4_CHUNK * 48, prm.pre_mats + (size_t)first * 3 * C * 16, mbytes, &bars[ld & 1]);
			};
			if (tid == 0) issue(0, loads);
			for (int ch = 0; ch < nchunks; ch++) {
				if (tid == 0 && ch + 1 < nchunks) issue(ch + 1, loads + 1);
				mbar_wait(&bars[loads & 1], (loads >> 1) & 1);
				const unsigned char *st = stage0 + (loads & 1) * stage_bytes;
				const phbc_pre_op *desc = reinterpret_cast<const phbc_pre_op *>(st);
				const double *mats = reinterpret_cast<const double *>(st + NUC4_CHUNK * 48);
				const int first = ch * NUC4_CHUNK;
				const int cnt = min(NUC4_CHUNK, prm.n_pre - first);
				// operands of the first op of the chunk
				double La[4], Lb[4];
				unsigned codeA = 0, codeB = 0;
				auto fetch = [&](const phbc_pre_op &d, double (&A)[4], double (&B)[4], unsigned &ca, unsigned &cb) {
					if (d.a_tip) ca = codes[(size_t)d.a_node * prm.P];
					else load_lower(my_lower + (((size_t)d.a_row * C + c) * PB + pl) * 4, A);
					if (d.b_tip) cb = codes[(size_t)d.b_node * prm.P];
					else load_lower(my_lower + (((size_t)d.b_row * C + c) * PB + pl) * 4, B);
				};
				fetch(desc[0], La, Lb, codeA, codeB);
				for (int j = 0; j < cnt; j++) {
					const phbc_pre_op d = desc[j];
					double xa[4], xb[4];
#pragma unroll
					for (int i = 0; i < 4; i++) xa[i] = La[i], xb[i] = Lb[i];
					const unsigned curA = codeA, curB = codeB;
					if (j + 1 < cnt) fetch(desc[j + 1], La, Lb, codeA, codeB);  // prefetch next operands
					const double *MP = mats + ((size_t)(j * 3 + 0) * C + c) * 16;
					const double *MA = mats + ((size_t)(j * 3 + 1) * C + c) * 16;
					const double *MB = mats + ((size_t)(j * 3 + 2) * C + c) * 16;
					double W[4], ma[4], mb[4];
					if (d.u_kind == PHBC_W_ROOT) {
						// children of the root: u = P_s L_s [o pi] (treelikelihood.c:2145-2154)
#pragma unroll
						for (int i = 0; i < 4; i++) W[i] = prm.include_root_freqs ? prm.freqs[i] : 1.0;
					} else {
						double up[4];
						slots.load(d.u_slot, up);
						matvec_smem(MP, up, W);  // P_p u_p
					}
					if (d.a_tip) tip_message(MA, curA, ma);
					else matvec_smem(MA, xa, ma);
					if (d.b_tip) tip_message(MB, curB, mb);
					else matvec_smem(MB, xb, mb);
					double ua[4], ub[4];
#pragma unroll
					for (int i = 0; i < 4; i++) ua[i] = W[i] * mb[i], ub[i] = W[i] * ma[i];
					// numerators: sum_i f_i u_n[i] (dP_n L_n)[i] with dP_n L_n = Q (P_n L_n)
					double na = 0.0, nb = 0.0, da = 0.0, db = 0.0;
#pragma unroll
					for (int i = 0; i < 4; i++) {
						const double f = prm.include_root_freqs ? 1.0 : prm.freqs[i];
						const double qa = fma(prm.Q[4 * i], ma[0], fma(prm.Q[4 * i + 1], ma[1], fma(prm.Q[4 * i + 2], ma[2], prm.Q[4 * i + 3] * ma[3])));
						const double qb = fma(prm.Q[4 * i], mb[0], fma(prm.Q[4 * i + 1], mb[1], fma(prm.Q[4 * i + 2], mb[2], prm.Q[4 * i + 3] * mb[3])));
						na = fma(f * ua[i], qa, na);
						nb = fma(f * ub[i], qb, nb);
						if (SCALE) {
							da = fma(f * ua[i], ma[i], da);
							db = fma(f * ub[i], mb[i], db);
						}
					}
					double va, vb;
					if (!SCALE) {
						va = na * sgrad;
						vb = nb * sgrad;
					} else {
						// rescale the upper partials like the reference (their scale cancels in the ratios below)
						double *plane = xch;  // 4 planes: max_a, max_b, den_a, den_b
						__syncthreads();      // previous op's readers are done
						plane[0 * C * PB + c * PB + pl] = fmax(fmax(ua[0], ua[1]), fmax(ua[2], ua[3]));
						plane[1 * C * PB + c * PB + pl] = fmax(fmax(ub[0], ub[1]), fmax(ub[2], ub[3]));
						plane[2 * C * PB + c * PB + pl] = da * prop_c;
						plane[3 * C * PB + c * PB + pl] = db * prop_c;
						__syncthreads();
						double mxa = 0.0, mxb = 0.0, dta = 0.0, dtb = 0.0;
						for (int cc = 0; cc < C; cc++) {
							mxa = fmax(mxa, plane[0 * C * PB + cc * PB + pl]);
							mxb = fmax(mxb, plane[1 * C * PB + cc * PB + pl]);
							dta += plane[2 * C * PB + cc * PB + pl];
							dtb += plane[3 * C * PB + cc * PB + pl];
						}
						if (mxa < prm.threshold) {
#pragma unroll
							for (int i = 0; i < 4; i++) ua[i] /= mxa;
						}
						if (mxb < prm.threshold) {
#pragma unroll
							for (int i = 0; i < 4; i++) ub[i] /= mxb;
						}
						// exact: one site denominator shared by the categories; compat: per-category ratio
						// (gradient_cat_branch_lengths_aux, treelikelihood.c:2721-2738)
						va = prm.compat ? na / da * wk : na / dta * wk;
						vb = prm.compat ? nb / db * wk : nb / dtb * wk;
					}
					if (d.a_slot >= 0) slots.store(d.a_slot, ua);
					if (d.b_slot >= 0) slots.store(d.b_slot, ub);
					// paired butterfly: lanes 0-15 end with branch a, lanes 16-31 with branch b
					const bool hi = lane & 16;
					double r = (hi ? vb : va) + __shfl_xor_sync(0xffffffffu, hi ? va : vb, 16);
					r += __shfl_xor_sync(0xffffffffu, r, 8);
					r += __shfl_xor_sync(0xffffffffu, r, 4);
					r += __shfl_xor_sync(0xffffffffu, r, 2);
					r += __shfl_xor_sync(0xffffffffu, r, 1);
					if (lane == 0) red_add_f64(my_gacc + d.a_node, r);
					if (lane == 16) red_add_f64(my_gacc + d.b_node, r);
				}
				loads++;
				__syncthreads();
			}
		}
	}
	if (c == 0 && lane == 0) {
		// one partial lnL per category-0 warp
		prm.cta_lnl[(size_t)blockIdx.x * (NUC4_NT / 32) + warp] = cta_lnl;
	} else if (lane == 0) {
		prm.cta_lnl[(size_t)blockIdx.x * (NUC4_NT / 32) + warp] = 0.0;
	}
}

// ---------------------------------------------------------------------------------------------
// walk-ordered transition matrices: P(t) = |V exp(L t) V^-1| (substmodel.c:518-557), one thread per
// (matrix, category).  post entries: [op][a|b]; pre entries: [op][parent|a|b].
// ---------------------------------------------------------------------------------------------
__global__ void k_nuc4_matrices(int C, int root, int n_post, int n_pre, const phbc_post_op *__restrict__ post_ops,
                                const phbc_pre_op *__restrict__ pre_ops, const double *__restrict__ evec,
                                const double *__restrict__ eval, const double *__restrict__ ivec, const double *__restrict__ bl,
                                const double *__restrict__ rates, double *__restrict__ post_mats, double *__restrict__ pre_mats) {
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	const int total = (2 * n_post + 3 * n_pre) * C;
	if (e >= total) return;
	const int c = e % C;
	const int m = e / C;
	int node;
	double *dst;
	if (m < 2 * n_post) {
		const phbc_post_op op = post_ops[m >> 1];
		node = (m & 1) ? op.b_node : op.a_node;
		dst = post_mats + (size_t)e * 16;
	} else {
		const int mm = m - 2 * n_post;
		const phbc_pre_op op = pre_ops[mm / 3];
		const int which = mm % 3;
		node = which == 0 ? op.node : (which == 1 ? op.a_node : op.b_node);
		dst = pre_mats + ((size_t)mm * C + c) * 16;
	}
	if (node == root) {
#pragma unroll
		for (int k = 0; k < 16; k++) dst[k] = 0.0;
		return;
	}
	const double t = bl[node] * rates[c];
	double ex[4];
#pragma unroll
	for (int k = 0; k < 4; k++) ex[k] = exp(eval[k] * t);
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) {
			double acc = 0.0;
#pragma unroll
			for (int k = 0; k < 4; k++) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(ivec[k * 4 + j], ex[k]), evec[i * 4 + k]));  // substmodel.c:539-555 order
			dst[i * 4 + j] = fabs(acc);
		}
}

// tip encodings -> 5-bit codes.  states: s < 4 -> 1 << s, else 0x10 (missing).  partials: bit j set when
// partial[j] == 1; anything that is not a 0/1 vector raises *bad (the fused path then declines).
__global__ void k_nuc4_encode_states(size_t n, const uint8_t *__restrict__ states, uint8_t *__restrict__ codes) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const unsigned s = states[i];
	codes[i] = s < 4 ? (uint8_t)(1u << s) : (uint8_t)0x10;
}
__global__ void k_nuc4_encode_partials(size_t n, const double *__restrict__ partials, uint8_t *__restrict__ codes, int *__restrict__ bad) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	unsigned code = 0;
	for (int j = 0; j < 4; j++) {
		const double v = partials[i * 4 + j];
		if (v == 1.0) code |= 1u << j;
		else if (v != 0.0) *bad = 1;
	}
	codes[i] = (uint8_t)code;
}

// fixed-order final sums: lnL and cat_grad[n][c] from the per-CTA / per-warp partials
__global__ void k_nuc4_finalize(int N, int C, int PB, int grid, int root, const double *__restrict__ cta_lnl,
                                const double *__restrict__ gacc, int want_grad, double *__restrict__ cat_grad,
                                double *__restrict__ result) {
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	const int warps = NUC4_NT / 32, wpc = PB / 32;
	if (e == 0) {
		double s = 0.0;
		for (int i = 0; i < grid * warps; i++) s += cta_lnl[i];
		result[0] = s;
	}
	if (!want_grad || e >= N * C) return;
	const int c = e / N, n = e % N;  // consecutive threads -> consecutive nodes (coalesced rows)
	double s = 0.0;
	if (n != root)
		for (int b = 0; b < grid; b++)
			for (int w = 0; w < wpc; w++) s += gacc[((size_t)b * warps + c * wpc + w) * N + n];
	cat_grad[(size_t)n * C + c] = s;
}

__global__ void k_collapse_categories_nuc4(int N, int C, const double *__restrict__ cat_grad, const double *__restrict__ props,
                                           const double *__restrict__ rates, double *__restrict__ result) {
	const int n = blockIdx.x * blockDim.x + threadIdx.x;
	if (n >= N) return;
	double g;
	if (C == 1) {
		g = cat_grad[n];
	} else {
		g = cat_grad[(size_t)n * C] * props[0] * rates[0];
		for (int c = 1; c < C; c++) g += cat_grad[(size_t)n * C + c] * props[c] * rates[c];
	}
	result[1 + n] = g;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static size_t nuc4_smem_bytes(int C, int PB, int nslots, bool scale) {
	return 128 + 2 * (size_t)NUC4_CHUNK * (48 + 3 * C * 128) + (size_t)nslots * 2 * NUC4_NT * 16 +
	       (size_t)(4 * C * PB + PB + (scale ? nslots * NUC4_NT : 0)) * sizeof(double);
}

static int pattern_block(int C) {
	int pb = (NUC4_NT / C) & ~31;
	return pb;
}

bool phbc_nuc4_supported(const phbc_ctx *ctx, const phbc_eval_opts *o) {
	if (ctx->S != 4 || o->explicit_matrices || !ctx->have_eigen) return false;
	if (ctx->C > NUC4_NT / 32) return false;
	const int PB = pattern_block(ctx->C);
	if (PB < 32 || PB * ctx->C != NUC4_NT) return false;  // categories must tile the CTA exactly
	const int nslots = ctx->post_slots > ctx->pre_slots ? ctx->post_slots : ctx->pre_slots;
	if (nuc4_smem_bytes(ctx->C, PB, nslots, o->scale != 0) > ctx->smem_optin) return false;
	return true;
}


int phbc_nuc4_evaluate(phbc_ctx *ctx, const phbc_eval_opts *o) {
	PHBC_CHECK(cudaSetDevice(ctx->device));
	const int C = ctx->C, N = ctx->N, T = ctx->T, P = ctx->P;
	const int PB = pattern_block(C);
	const int nslots = ctx->post_slots > ctx->pre_slots ? ctx->post_slots : ctx->pre_slots;
	const size_t smem = nuc4_smem_bytes(C, PB, nslots, o->scale != 0);
	const int ntiles = (P + PB - 1) / PB;
	// tip codes (once per tip upload)
	if (!ctx->d_nuc4_codes) {
		PHBC_CHECK(cudaMalloc((void **)&ctx->d_nuc4_codes, (size_t)T * P));
		PHBC_CHECK(cudaMalloc((void **)&ctx->d_nuc4_bad, sizeof(int)));
	}
	if (!ctx->nuc4_codes_valid) {
		const size_t n = (size_t)T * P;
		PHBC_CHECK(cudaMemsetAsync(ctx->d_nuc4_bad, 0, sizeof(int), ctx->stream));
		if (ctx->tip_kind == PHBC_TIP_STATES)
			k_nuc4_encode_states<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, ctx->d_tip_states, ctx->d_nuc4_codes);
		else
			k_nuc4_encode_partials<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, ctx->d_tip_partials, ctx->d_nuc4_codes, ctx->d_nuc4_bad);
		ctx->launches++;
		int bad = 0;
		PHBC_CHECK(cudaMemcpyAsync(&bad, ctx->d_nuc4_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
		PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
		ctx->nuc4_codes_bad = bad != 0;
		ctx->nuc4_codes_valid = true;
	}
	if (ctx->nuc4_codes_bad) return phbc_generic_evaluate(ctx, o);  // non 0/1 tip partials: node-at-a-time kernels

	// launch geometry: persistent CTAs, two per SM when shared memory allows
	auto kern = o->scale ? (o->want_gradient ? k_nuc4_walk<true, true> : k_nuc4_walk<true, false>)
	                     : (o->want_gradient ? k_nuc4_walk<false, true> : k_nuc4_walk<false, false>);
	PHBC_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int per_sm = 0;
	PHBC_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NUC4_NT, smem));
	if (per_sm < 1) {
		snprintf(phbc_errbuf, sizeof(phbc_errbuf), "nuc4 walk kernel does not fit on an SM (smem %zu)", smem);
		return -1;
	}
	int grid = per_sm * ctx->num_sms;
	if (grid > ntiles) grid = ntiles;
	const int warps = NUC4_NT / 32;
	// scratch: walk matrices, per-CTA lower rows, per-warp gradient rows, per-warp lnL
	const size_t mats_bytes = (size_t)(2 * ctx->n_post + 3 * ctx->n_pre) * C * 16 * sizeof(double);
	if (mats_bytes > ctx->walk_mats_bytes) {
		PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
		if (ctx->d_walk_mats) cudaFree(ctx->d_walk_mats);
		ctx->d_walk_mats = NULL;
		PHBC_CHECK(cudaMalloc((void **)&ctx->d_walk_mats, mats_bytes));
		ctx->walk_mats_bytes = mats_bytes;
	}
	if (o->want_gradient) {
		const size_t lower_bytes = (size_t)grid * ctx->n_post * C * PB * 4 * sizeof(double);
		if (lower_bytes > ctx->walk_lower_bytes) {
			PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
			if (ctx->d_walk_lower) cudaFree(ctx->d_walk_lower);
			ctx->d_walk_lower = NULL;
			PHBC_CHECK(cudaMalloc((void **)&ctx->d_walk_lower, lower_bytes));
			ctx->walk_lower_bytes = lower_bytes;
		}
		const size_t gacc_bytes = (size_t)grid * warps * N * sizeof(double);
		if (gacc_bytes > ctx->walk_gacc_bytes) {
			PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
			if (ctx->d_walk_gacc) cudaFree(ctx->d_walk_gacc);
			ctx->d_walk_gacc = NULL;
			PHBC_CHECK(cudaMalloc((void **)&ctx->d_walk_gacc, gacc_bytes));
			ctx->walk_gacc_bytes = gacc_bytes;
		}
		PHBC_CHECK(cudaMemsetAsync(ctx->d_walk_gacc, 0, gacc_bytes, ctx->stream));
	}
	if (!ctx->d_nuc4_cta_lnl || ctx->nuc4_grid < grid) {
		PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
		if (ctx->d_nuc4_cta_lnl) cudaFree(ctx->d_nuc4_cta_lnl);
		ctx->d_nuc4_cta_lnl = NULL;
		PHBC_CHECK(cudaMalloc((void **)&ctx->d_nuc4_cta_lnl, (size_t)grid * warps * sizeof(double)));
		ctx->nuc4_grid = grid;
	}
	double *post_mats = ctx->d_walk_mats;
	double *pre_mats = ctx->d_walk_mats + (size_t)2 * ctx->n_post * C * 16;
	{
		const int total = (2 * ctx->n_post + 3 * ctx->n_pre) * C;
		k_nuc4_matrices<<<(total + 127) / 128, 128, 0, ctx->stream>>>(C, ctx->root, ctx->n_post, ctx->n_pre, ctx->d_post_ops, ctx->d_pre_ops,
		                                                             ctx->d_evec, ctx->d_eval, ctx->d_ivec,
		                                                             ctx->d_bl + (size_t)o->batch_index * N, ctx->d_rates, post_mats, pre_mats);
		ctx->launches++;
	}
	Nuc4Params prm;
	memset(&prm, 0, sizeof(prm));
	prm.T = T, prm.N = N, prm.C = C, prm.P = P, prm.PB = PB, prm.root = ctx->root;
	prm.n_post = ctx->n_post, prm.n_pre = ctx->n_pre, prm.nslots = nslots, prm.ntiles = ntiles;
	prm.include_root_freqs = o->include_root_freqs, prm.compat = o->compat_scaled_gradient;
	prm.threshold = o->scaling_threshold;
	prm.tip_codes = ctx->d_nuc4_codes;
	prm.weights = ctx->d_weights;
	prm.props = ctx->d_props;
	prm.post_ops = ctx->d_post_ops;
	prm.pre_ops = ctx->d_pre_ops;
	prm.post_mats = post_mats;
	prm.pre_mats = pre_mats;
	prm.lower = ctx->d_walk_lower;
	prm.gacc = ctx->d_walk_gacc;
	prm.cta_lnl = ctx->d_nuc4_cta_lnl;
	prm.pattern_lnl = ctx->d_pattern_lnl;
	memcpy(prm.freqs, ctx->h_freqs, sizeof(prm.freqs));  // small model constants travel in the kernel parameter bank
	memcpy(prm.Q, ctx->h_qmat, sizeof(prm.Q));
	int trc;
	if ((trc = phbc_time_begin(ctx))) return trc;
	kern<<<grid, NUC4_NT, smem, ctx->stream>>>(prm);
	ctx->launches++;
	if ((trc = phbc_time_end(ctx))) return trc;
	double *result = ctx->d_result + (size_t)o->batch_index * (1 + N);
	k_nuc4_finalize<<<(N * C + 127) / 128, 128, 0, ctx->stream>>>(N, C, PB, grid, ctx->root, ctx->d_nuc4_cta_lnl, ctx->d_walk_gacc,
	                                                             o->want_gradient, ctx->d_cat_grad, result);
	ctx->launches++;
	if (o->want_gradient) {
		k_collapse_categories_nuc4<<<(N + 127) / 128, 128, 0, ctx->stream>>>(N, C, ctx->d_cat_grad, ctx->d_props, ctx->d_rates, result);
		ctx->launches++;
	}
	PHBC_CHECK(cudaGetLastError());
	return 0;
}
