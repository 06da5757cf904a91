// phb_dwalk.cu -- whole-tree walk on the FP64 tensor cores for 20-state (amino-acid) models.
//
// Replaces, for one pattern tile at a time and without materialising a partial per node, the reference's update_partials_20_SSE
// (treelikelihood20.c:114-647), update_upper_partials (treelikelihood.c:2129-2162), calculate_branch_partials_20_SSE
// (treelikelihood20.c:834-1025) and the reductions of gradient_cat_branch_lengths (treelikelihood.c:2793-2941).
//
// The level-batched kernels of phb_dmma.cu stream every operand through HBM: 8 rows of S doubles per (pattern, category, internal
// node) and evaluation, which makes 20 states HBM-bound (AI = 4 flop / byte).  Here a CTA owns a tile of TP patterns of ONE rate
// category and walks the whole tree in the DFS orders the host compiled for the 4-state walk (phb_treelikelihood.c,
// build_walk_schedules): what travels between consecutive ops stays in shared memory, first-visited children are parked in K
// shared-memory slots (Strahler number of the tree; further slots spill to HBM), and HBM sees each message M_n = P_n L_n exactly
// twice -- written by the post-order pass, read by the pre-order pass -- through 1-D TMA bulk copies issued per warp (a warp's 16
// pattern rows are contiguous in the [pattern][state] layout of the lower buffers).  Upper partials never leave the SM.
//
//   k_dwalk_post   L_n = M_a o M_b formed straight into the A fragments (tips: columns of the transposed matrix image),
//                  M_n = L_n P_n^T on the tensor pipe, result to shared memory (next op / parked) and to HBM (bulk store).
//   [phbc_generic_root: site likelihoods from the root's row, all categories]
//   k_dwalk_pre    per internal node n: W = U_n P_n^T and Z = U_n (pi o dP_n) from the same A fragments, the children's messages
//                  from HBM (bulk load, one op ahead) or the tips' images, U_a = W o M_b, U_b = W o M_a to shared memory, branch
//                  gradients in adjoint form (see phb_dmma.cu) reduced per warp and added to warp-private rows (deterministic).
//
// Work items are (pattern tile, category) pairs on a persistent grid of one CTA per SM; warps own disjoint pattern rows and share
// only the matrix images, which a PRODUCER WARP stages op by op through a ring of TMA bulk copies (full / empty mbarriers).
// Bound after the change: the FP64 pipe (DMMA and the element-wise products share it); HBM traffic falls from the streaming model's
// 8 rows per node to ~2 (profiles/).
#include "phb_dmma_common.cuh"

#include <stdlib.h>
#include <string.h>

// ---------------------------------------------------------------------------------------------
// descriptors (device copies of the host's walk schedules with GLOBAL tip positions)
// ---------------------------------------------------------------------------------------------
struct DwPost {  // 32 bytes: one internal node, DFS post-order
	int node, a_node, b_node;
	int a_tip, b_tip;          // position of a tip child's code row in post-order walk order (-1: internal)
	int16_t kind;              // 0 tip-tip, 1 tip-internal (b = the preceding op's result), 2 internal-internal (a parked, b preceding)
	int16_t a_slot, dst_slot;  // slot holding a (kind 2) / slot receiving the result (-1: only the hand-over buffer)
	int16_t pad0;
	int pad1;
};
struct DwPre {  // 48 bytes: one internal node acting as parent, DFS pre-order
	int node, a_node, b_node;
	int a_tip, b_tip;          // pre-order walk positions of tip children's code rows
	int16_t kind;              // 0 tip-tip, 1 tip-internal (a is the tip), 2 internal-internal
	int16_t u_kind;            // PHBC_W_ROOT / PHBC_W_REG (left in the hand-over buffer by the preceding op) / PHBC_W_SLOT
	int16_t u_slot, a_slot;    // slot holding U_node / slot receiving U_a (kind 2)
	int pf_a, pf_b;            // internal children of the op DW_PF ops later: L2 prefetch targets (-1: none)
	int pad[3];
};
static_assert(sizeof(DwPost) == 32 && sizeof(DwPre) == 48, "descriptor sizes are TMA granules");
#define DW_PF 3

struct DwParams {
	const double *img;  // packed images [P | dP][N][C][IMG]
	double *lower;      // message rows [N - T][C][P][S]
	const uint8_t *codes;
	const void *ops;
	int nops, T, N, C, P, ntiles, nitems, root;
	int K, nstg, nw, spill;  // shared-memory slots per warp, ring depth, consumer warps, slots beyond K exist
	double *spillbuf;        // [slots - K][C][P][S]
	const double *freqs, *weights, *pattern_lnl;
	int include_root_freqs, pstride;
	double *gacc;            // [N][C][pstride] warp-private gradient rows
};

// ---------------------------------------------------------------------------------------------
// PTX helpers beyond phb_ctx.cuh: bulk stores, L2 prefetch, mbarrier arrive
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
	asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
	asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_prefetch_l2(const void *g, uint32_t bytes) {
	asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// shared-memory geometry, the same arithmetic on host and device
template <int S, int MT>
struct DwGeom {
	using Sh = DmmaShape<S>;
	static_assert(S % 8 == 4, "row-major tiles of stride S are bank-conflict free (and need no k padding) only for S = 4 (mod 8)");
	static constexpr int ROWS = MT * 8;                 // pattern rows per warp
	static constexpr int SLICE = ROWS * S;              // doubles per warp and tile slot
	static constexpr int IMGB = Sh::IMG * 8;            // bytes per matrix image
	static constexpr int TPMAX = 8 * ROWS;              // patterns per tile at 8 consumer warps
	static constexpr int CODE_OFF = 64, IMG_OFF = CODE_OFF + 2 * TPMAX;
	static constexpr int HDR = 1024;                    // barriers, frequency vectors
	__host__ __device__ static constexpr int stage_bytes(int nimg) { return (IMG_OFF + nimg * IMGB + 127) / 128 * 128; }
};

// ---------------------------------------------------------------------------------------------
// tip codes, tile-major in walk order: codes[pass][tile][k][TP], code = state (< S) or S (unknown / padding pattern)
// tip partials are accepted when every vector is one-hot or all ones (what SitePattern_get_partials produces for unambiguous data)
// ---------------------------------------------------------------------------------------------
__global__ void k_dwalk_codes(int T, int P, int S, int TP, int ntiles, int tip_kind, const uint8_t *__restrict__ states,
                              const double *__restrict__ partials, const int *__restrict__ post_order, const int *__restrict__ pre_order,
                              uint8_t *__restrict__ codes, int *__restrict__ bad) {
	const size_t per_walk = (size_t)ntiles * T * TP;
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 2 * per_walk) return;
	const int walk = i >= per_walk;
	const size_t r = i - walk * per_walk;
	const int pl = (int)(r % TP);
	const int k = (int)((r / TP) % T);
	const int tile = (int)(r / ((size_t)TP * T));
	const int p = tile * TP + pl;
	const int tip = walk ? pre_order[k] : post_order[k];
	int code = S;
	if (p < P) {
		if (tip_kind == PHBC_TIP_STATES) {
			const int s = states[(size_t)tip * P + p];
			code = s < S ? s : S;
		} else {
			const double *v = partials + ((size_t)tip * P + p) * S;
			int ones = 0, first = -1;
			for (int j = 0; j < S; j++) {
				if (v[j] == 1.0) {
					if (first < 0) first = j;
					ones++;
				} else if (v[j] != 0.0) *bad = 1;
			}
			if (ones == 1) code = first;
			else if (ones == S) code = S;
			else *bad = 1;  // an ambiguity set: its message is a sum of columns, not a gather
		}
	}
	codes[i] = (uint8_t)code;
}

// ---------------------------------------------------------------------------------------------
// producer warp: stages descriptor, tip codes and matrix images of every op of every item of this CTA
// ---------------------------------------------------------------------------------------------
template <int S, int MT, bool PRE>
__device__ __forceinline__ void dwalk_producer(const DwParams &p, unsigned char *smraw, uint64_t *full, uint64_t *empty) {
	using G = DwGeom<S, MT>;
	using Sh = DmmaShape<S>;
	constexpr int NIMG = PRE ? 6 : 3;
	const int stgb = G::stage_bytes(NIMG);
	const int TP = p.nw * G::ROWS;
	const size_t dimg = (size_t)p.N * p.C * Sh::IMG;
	int s = 0;
	uint32_t ph = 1;  // a fresh barrier passes a wait on the phase "before" its first one
	for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
		const int tile = item / p.C, c = item - tile * p.C;
		const uint8_t *codes = p.codes + (size_t)tile * p.T * TP;
		for (int k = 0; k < p.nops; k++) {
			unsigned char *stg = smraw + G::HDR + (size_t)s * stgb;
			double *im = reinterpret_cast<double *>(stg + G::IMG_OFF);
			mbar_wait(&empty[s], ph);
			if (PRE) {
				const DwPre *d = reinterpret_cast<const DwPre *>(p.ops) + k;
				const int node = d->node, a_node = d->a_node, b_node = d->b_node, a_tip = d->a_tip, b_tip = d->b_tip;
				const bool root = d->u_kind == PHBC_W_ROOT;
				const int nimg = (root ? 0 : 2) + (a_tip >= 0 ? 2 : 0) + (b_tip >= 0 ? 2 : 0);
				mbar_expect_tx(&full[s], (uint32_t)(sizeof(DwPre) + nimg * G::IMGB + ((a_tip >= 0) + (b_tip >= 0)) * TP));
				bulk_g2s(stg, d, sizeof(DwPre), &full[s]);
				if (!root) {
					bulk_g2s(im, p.img + ((size_t)node * p.C + c) * Sh::IMG, G::IMGB, &full[s]);
					bulk_g2s(im + Sh::IMG, p.img + dimg + ((size_t)node * p.C + c) * Sh::IMG, G::IMGB, &full[s]);
				}
				if (a_tip >= 0) {
					bulk_g2s(im + 2 * Sh::IMG, p.img + ((size_t)a_node * p.C + c) * Sh::IMG, G::IMGB, &full[s]);
					bulk_g2s(im + 3 * Sh::IMG, p.img + dimg + ((size_t)a_node * p.C + c) * Sh::IMG, G::IMGB, &full[s]);
					bulk_g2s(stg + G::CODE_OFF, codes + (size_t)a_tip * TP, TP, &full[s]);
				}
				if (b_tip >= 0) {
					bulk_g2s(im + 4 * Sh::IMG, p.img + ((size_t)b_node * p.C + c) * Sh::IMG, G::IMGB, &full[s]);
					bulk_g2s(im + 5 * Sh::IMG, p.img + dimg + ((size_t)b_node * p.C + c) * Sh::IMG, G::IMGB, &full[s]);
					bulk_g2s(stg + G::CODE_OFF + G::TPMAX, codes + (size_t)b_tip * TP, TP, &full[s]);
				}
			} else {
				const DwPost *d = reinterpret_cast<const DwPost *>(p.ops) + k;
				const int node = d->node, a_node = d->a_node, b_node = d->b_node, a_tip = d->a_tip, b_tip = d->b_tip;
				const int nimg = 1 + (a_tip >= 0) + (b_tip >= 0);
				mbar_expect_tx(&full[s], (uint32_t)(sizeof(DwPost) + nimg * G::IMGB + ((a_tip >= 0) + (b_tip >= 0)) * TP));
				bulk_g2s(stg, d, sizeof(DwPost), &full[s]);
				bulk_g2s(im, p.img + ((size_t)node * p.C + c) * Sh::IMG, G::IMGB, &full[s]);
				if (a_tip >= 0) {
					bulk_g2s(im + Sh::IMG, p.img + ((size_t)a_node * p.C + c) * Sh::IMG, G::IMGB, &full[s]);
					bulk_g2s(stg + G::CODE_OFF, codes + (size_t)a_tip * TP, TP, &full[s]);
				}
				if (b_tip >= 0) {
					bulk_g2s(im + 2 * Sh::IMG, p.img + ((size_t)b_node * p.C + c) * Sh::IMG, G::IMGB, &full[s]);
					bulk_g2s(stg + G::CODE_OFF + G::TPMAX, codes + (size_t)b_tip * TP, TP, &full[s]);
				}
			}
			if (++s == p.nstg) s = 0, ph ^= 1;
		}
	}
}

// D-layout accumulators (row 8 m + r, columns 8 j + 2 q, + 1) to a row-major tile slice; padding columns are dropped
template <int S, int MT>
__device__ __forceinline__ void dw_store_slice(double *slice, int r, int q, const double (&v)[MT][DmmaShape<S>::NT][2]) {
#pragma unroll
	for (int m = 0; m < MT; m++)
#pragma unroll
		for (int j = 0; j < DmmaShape<S>::NT; j++) {
			const int col = 8 * j + 2 * q;
			if (col < S) *reinterpret_cast<double2 *>(slice + (8 * m + r) * S + col) = make_double2(v[m][j][0], v[m][j][1]);
		}
}

// ---------------------------------------------------------------------------------------------
// post-order pass
// ---------------------------------------------------------------------------------------------
template <int S, int MT>
__global__ void __launch_bounds__(288, 1) k_dwalk_post(const DwParams p) {
	using G = DwGeom<S, MT>;
	using Sh = DmmaShape<S>;
	extern __shared__ __align__(128) unsigned char smraw[];
	uint64_t *full = reinterpret_cast<uint64_t *>(smraw), *empty = full + 8, *lbar = full + 16;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int nw = p.nw, nstg = p.nstg, K = p.K;
	const int stgb = G::stage_bytes(3);
	const int nslice = K + 2 + (p.spill ? 1 : 0);
	double *tiles = reinterpret_cast<double *>(smraw + G::HDR + (size_t)nstg * stgb);
	if (threadIdx.x == 0) {
		for (int i = 0; i < nstg; i++) mbar_init(&full[i], 1), mbar_init(&empty[i], nw);
		for (int w = 0; w < nw; w++) mbar_init(&lbar[w], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	for (int i = threadIdx.x; i < nw * nslice * G::SLICE; i += blockDim.x) tiles[i] = 0.0;  // rows past P keep finite values
	__syncthreads();
	if (warp == nw) {
		if (lane == 0) dwalk_producer<S, MT, false>(p, smraw, full, empty);
		return;
	}
	const int r = lane >> 2, q = lane & 3;
	double *mine = tiles + (size_t)warp * nslice * G::SLICE;
	double *cur0 = mine + (size_t)K * G::SLICE, *cur1 = cur0 + G::SLICE, *land = cur1 + G::SLICE;
	const int TP = nw * G::ROWS;
	const size_t PS = (size_t)p.P * S;
	int s = 0, cb = 0;
	uint32_t ph = 0, lph = 0;
	for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
		const int tile = item / p.C, c = item - tile * p.C;
		const int p0 = tile * TP + warp * G::ROWS;
		const int rows = min(max(p.P - p0, 0), G::ROWS);
		const uint32_t rbytes = (uint32_t)rows * S * 8;
		for (int k = 0; k < p.nops; k++) {
			const unsigned char *stg = smraw + G::HDR + (size_t)s * stgb;
			mbar_wait(&full[s], ph);
			const DwPost *d = reinterpret_cast<const DwPost *>(stg);
			const int kind = d->kind, a_slot = d->a_slot, dst_slot = d->dst_slot, node = d->node, a_node = d->a_node;  // the stage is recycled after the arrive below
			const double *mN = reinterpret_cast<const double *>(stg + G::IMG_OFF), *mA = mN + Sh::IMG, *mB = mA + Sh::IMG;
			const double *prev = cb ? cur1 : cur0;
			const double *ta = prev;
			if (kind == 2) {
				if (a_slot < K) ta = mine + (size_t)a_slot * G::SLICE;
				else {  // parked beyond the shared-memory slots: its message row is in HBM already
					if (rows > 0) {
						if (lane == 0) {
							bulk_wait_all<0>();  // the row was written by this thread's own bulk store
							mbar_expect_tx(&lbar[warp], rbytes);
							bulk_g2s(land, p.lower + ((size_t)(a_node - p.T) * p.C + c) * PS + (size_t)p0 * S, rbytes, &lbar[warp]);
						}
						mbar_wait(&lbar[warp], lph);
						lph ^= 1;
					}
					ta = land;
				}
			}
			// A fragments: L_n = M_a o M_b
			double a[MT][Sh::KT];
#pragma unroll
			for (int m = 0; m < MT; m++) {
				const int row = 8 * m + r;
				const int sa = kind < 2 ? stg[G::CODE_OFF + warp * G::ROWS + row] : 0;
				const int sb = kind == 0 ? stg[G::CODE_OFF + G::TPMAX + warp * G::ROWS + row] : 0;
#pragma unroll
				for (int tt = 0; tt < Sh::KT; tt++) {
					const int col = 4 * tt + q;
					const double va = kind < 2 ? mA[sa * Sh::NP + col] : ta[row * S + col];
					const double vb = kind == 0 ? mB[sb * Sh::NP + col] : prev[row * S + col];
					a[m][tt] = va * vb;
				}
			}
			// M_n = L_n P_n^T
			double acc[MT][Sh::NT][2];
#pragma unroll
			for (int m = 0; m < MT; m++)
#pragma unroll
				for (int j = 0; j < Sh::NT; j++) acc[m][j][0] = acc[m][j][1] = 0.0;
			double bN = mN[r * Sh::LD + q];
#pragma unroll
			for (int tt = 0; tt < Sh::KT; tt++)
#pragma unroll
				for (int j = 0; j < Sh::NT; j++) {
					const int nj = j + 1 < Sh::NT ? j + 1 : 0, nt = j + 1 < Sh::NT ? tt : (tt + 1 < Sh::KT ? tt + 1 : 0);
					const double nN = mN[(nj * 8 + r) * Sh::LD + 4 * nt + q];
#pragma unroll
					for (int m = 0; m < MT; m++) dmma_m8n8k4(acc[m][j][0], acc[m][j][1], a[m][tt], bN);
					bN = nN;
				}
			__syncwarp();
			if (lane == 0) mbar_arrive(&empty[s]);  // images and codes of this op are consumed
			// result: parked slot or the other hand-over buffer, then to HBM from there
			double *dst;
			if (dst_slot >= 0 && dst_slot < K) dst = mine + (size_t)dst_slot * G::SLICE;
			else dst = cb ? cur0 : cur1, cb ^= 1;
			if (lane == 0) bulk_wait_read<1>();  // every store but the preceding op's has left shared memory
			__syncwarp();
			dw_store_slice<S, MT>(dst, r, q, acc);
			fence_async_smem();
			__syncwarp();
			if (lane == 0 && rows > 0) {
				bulk_s2g(p.lower + ((size_t)(node - p.T) * p.C + c) * PS + (size_t)p0 * S, dst, rbytes);
				bulk_commit();
			}
			if (++s == nstg) s = 0, ph ^= 1;
		}
	}
	if (lane == 0) bulk_wait_all<0>();
}

// ---------------------------------------------------------------------------------------------
// pre-order pass with branch gradients
// ---------------------------------------------------------------------------------------------
template <int S, int MT>
__global__ void __launch_bounds__(288, 1) k_dwalk_pre(const DwParams p) {
	using G = DwGeom<S, MT>;
	using Sh = DmmaShape<S>;
	extern __shared__ __align__(128) unsigned char smraw[];
	uint64_t *full = reinterpret_cast<uint64_t *>(smraw), *empty = full + 8, *lbar = full + 16;
	double *fq = reinterpret_cast<double *>(smraw + 256), *wroot = fq + Sh::NP;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int nw = p.nw, nstg = p.nstg, K = p.K;
	const int stgb = G::stage_bytes(6);
	const int nslice = K + 3 + (p.spill ? 1 : 0);
	double *tiles = reinterpret_cast<double *>(smraw + G::HDR + (size_t)nstg * stgb);
	if (threadIdx.x == 0) {
		for (int i = 0; i < nstg; i++) mbar_init(&full[i], 1), mbar_init(&empty[i], nw);
		for (int w = 0; w < nw; w++) mbar_init(&lbar[w], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	for (int i = threadIdx.x; i < Sh::NP; i += blockDim.x) {
		const double f = i < S ? p.freqs[i] : 0.0;
		fq[i] = i < S ? (p.include_root_freqs ? 1.0 : f) : 0.0;
		wroot[i] = i < S ? (p.include_root_freqs ? f : 1.0) : 0.0;
	}
	for (int i = threadIdx.x; i < nw * nslice * G::SLICE; i += blockDim.x) tiles[i] = 0.0;
	__syncthreads();
	if (warp == nw) {
		if (lane == 0) dwalk_producer<S, MT, true>(p, smraw, full, empty);
		return;
	}
	const int r = lane >> 2, q = lane & 3;
	double *mine = tiles + (size_t)warp * nslice * G::SLICE;
	double *cur = mine + (size_t)K * G::SLICE, *landA = cur + G::SLICE, *landB = landA + G::SLICE, *stage_out = landB + G::SLICE;
	const int TP = nw * G::ROWS;
	const size_t PS = (size_t)p.P * S, CPS = (size_t)p.C * PS;
	int s = 0;
	uint32_t ph = 0, lph = 0;
	double *grow = p.gacc + (size_t)blockIdx.x * nw + warp;

	// message rows (and a spilled U) of op d for (c, p0): one bulk load each into the landing slices, one mbarrier phase for all
	auto issue_loads = [&](const DwPre *d, int c, int p0, uint32_t rbytes) -> bool {
		const bool la = d->kind == 2, lb = d->kind >= 1, lu = d->u_kind == PHBC_W_SLOT && d->u_slot >= K;
		if (!(la || lb || lu) || rbytes == 0) return false;
		if (lane == 0) {
			if (lu) bulk_wait_all<0>();  // the spilled row was written by this thread's own bulk store
			mbar_expect_tx(&lbar[warp], ((la ? 1u : 0u) + (lb ? 1u : 0u) + (lu ? 1u : 0u)) * rbytes);
			if (la) bulk_g2s(landA, p.lower + ((size_t)(d->a_node - p.T) * p.C + c) * PS + (size_t)p0 * S, rbytes, &lbar[warp]);
			if (lb) bulk_g2s(landB, p.lower + ((size_t)(d->b_node - p.T) * p.C + c) * PS + (size_t)p0 * S, rbytes, &lbar[warp]);
			if (lu) bulk_g2s(cur, p.spillbuf + (size_t)(d->u_slot - K) * CPS + (size_t)c * PS + (size_t)p0 * S, rbytes, &lbar[warp]);
		}
		return true;
	};

	bool armed = false;
	{  // loads of the first op of the first item (the descriptor comes from global memory: its stage may not have landed)
		const int item = blockIdx.x;
		if (item < p.nitems) {
			const int tile = item / p.C, c = item - tile * p.C, p0 = tile * TP + warp * G::ROWS;
			const int rows = min(max(p.P - p0, 0), G::ROWS);
			armed = issue_loads(reinterpret_cast<const DwPre *>(p.ops), c, p0, (uint32_t)rows * S * 8);
		}
	}
	for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
		const int tile = item / p.C, c = item - tile * p.C;
		const int p0 = tile * TP + warp * G::ROWS;
		const int rows = min(max(p.P - p0, 0), G::ROWS);
		const uint32_t rbytes = (uint32_t)rows * S * 8;
		double wl[MT];
#pragma unroll
		for (int m = 0; m < MT; m++) {
			const int pp = p0 + 8 * m + r;
			wl[m] = pp < p.P ? __ldg(p.weights + pp) / exp(__ldg(p.pattern_lnl + pp)) : 0.0;
		}
		for (int k = 0; k < p.nops; k++) {
			const unsigned char *stg = smraw + G::HDR + (size_t)s * stgb;
			mbar_wait(&full[s], ph);
			const DwPre *d = reinterpret_cast<const DwPre *>(stg);
			const int kind = d->kind, u_kind = d->u_kind, u_slot = d->u_slot, a_slot = d->a_slot;
			const int node = d->node, a_node = d->a_node, b_node = d->b_node, pf_a = d->pf_a, pf_b = d->pf_b;  // the stage is recycled after the arrive below
			const bool root = u_kind == PHBC_W_ROOT;
			const double *mP = reinterpret_cast<const double *>(stg + G::IMG_OFF), *mZ = mP + Sh::IMG;
			const double *tA = mZ + Sh::IMG, *dA = tA + Sh::IMG, *tB = dA + Sh::IMG, *dB = tB + Sh::IMG;
			if (armed) {
				mbar_wait(&lbar[warp], lph);
				lph ^= 1;
			}
			// A fragments of U_n
			const double *usrc = (u_kind == PHBC_W_SLOT && u_slot < K) ? mine + (size_t)u_slot * G::SLICE : cur;
			double u[MT][Sh::KT];
			if (!root) {
#pragma unroll
				for (int m = 0; m < MT; m++)
#pragma unroll
					for (int tt = 0; tt < Sh::KT; tt++) u[m][tt] = usrc[(8 * m + r) * S + 4 * tt + q];
			}
			// the children's messages in accumulator layout
			int sa[MT], sb[MT];
#pragma unroll
			for (int m = 0; m < MT; m++) {
				sa[m] = kind < 2 ? stg[G::CODE_OFF + warp * G::ROWS + 8 * m + r] : 0;
				sb[m] = kind == 0 ? stg[G::CODE_OFF + G::TPMAX + warp * G::ROWS + 8 * m + r] : 0;
			}
			double Ma[MT][Sh::NT][2], Mb[MT][Sh::NT][2];
#pragma unroll
			for (int m = 0; m < MT; m++)
#pragma unroll
				for (int j = 0; j < Sh::NT; j++) {
					const int col = 8 * j + 2 * q;
					double2 va = make_double2(0.0, 0.0), vb = va;
					if (col < S) {
						va = *reinterpret_cast<const double2 *>(kind < 2 ? tA + sa[m] * Sh::NP + col : landA + (8 * m + r) * S + col);
						vb = *reinterpret_cast<const double2 *>(kind == 0 ? tB + sb[m] * Sh::NP + col : landB + (8 * m + r) * S + col);
					}
					Ma[m][j][0] = va.x, Ma[m][j][1] = va.y;
					Mb[m][j][0] = vb.x, Mb[m][j][1] = vb.y;
				}
			__syncwarp();  // every lane has read the landing slices and the hand-over buffer
			{  // rows of the next op (of the next item after the last op), and an L2 hint DW_PF ops ahead
				int nitem = item, nk = k + 1, ns = s + 1 == nstg ? 0 : s + 1;
				uint32_t nph = s + 1 == nstg ? ph ^ 1 : ph;
				if (nk == p.nops) nk = 0, nitem = item + gridDim.x;
				armed = false;
				if (nitem < p.nitems) {
					mbar_wait(&full[ns], nph);
					const DwPre *nd = reinterpret_cast<const DwPre *>(smraw + G::HDR + (size_t)ns * stgb);
					const int ntile = nitem / p.C, nc = nitem - ntile * p.C, np0 = ntile * TP + warp * G::ROWS;
					const int nrows = min(max(p.P - np0, 0), G::ROWS);
					armed = issue_loads(nd, nc, np0, (uint32_t)nrows * S * 8);
				}
				if (lane == 0 && rbytes) {
					if (pf_a >= 0) bulk_prefetch_l2(p.lower + ((size_t)(pf_a - p.T) * p.C + c) * PS + (size_t)p0 * S, rbytes);
					if (pf_b >= 0) bulk_prefetch_l2(p.lower + ((size_t)(pf_b - p.T) * p.C + c) * PS + (size_t)p0 * S, rbytes);
				}
			}
			// W = U_n P_n^T, Z = U_n (pi o dP_n)
			double W[MT][Sh::NT][2], Z[MT][Sh::NT][2];
#pragma unroll
			for (int m = 0; m < MT; m++)
#pragma unroll
				for (int j = 0; j < Sh::NT; j++) W[m][j][0] = W[m][j][1] = Z[m][j][0] = Z[m][j][1] = 0.0;
			if (!root) {
				double bP = mP[r * Sh::LD + q], bZ = mZ[r * Sh::LD + q];
#pragma unroll
				for (int tt = 0; tt < Sh::KT; tt++)
#pragma unroll
					for (int j = 0; j < Sh::NT; j++) {
						const int nj = j + 1 < Sh::NT ? j + 1 : 0, nt = j + 1 < Sh::NT ? tt : (tt + 1 < Sh::KT ? tt + 1 : 0);
						const int noff = (nj * 8 + r) * Sh::LD + 4 * nt + q;
						const double nP = mP[noff], nZ = mZ[noff];
#pragma unroll
						for (int m = 0; m < MT; m++) {
							dmma_m8n8k4(W[m][j][0], W[m][j][1], u[m][tt], bP);
							dmma_m8n8k4(Z[m][j][0], Z[m][j][1], u[m][tt], bZ);
						}
						bP = nP, bZ = nZ;
					}
			} else {
#pragma unroll
				for (int m = 0; m < MT; m++)
#pragma unroll
					for (int j = 0; j < Sh::NT; j++) W[m][j][0] = wroot[8 * j + 2 * q], W[m][j][1] = wroot[8 * j + 2 * q + 1];
			}
			// n's own branch (adjoint form), U_a = W o M_b, U_b = W o M_a, the branches of tip children
			double gn = 0.0, ga = 0.0, gb = 0.0;
#pragma unroll
			for (int m = 0; m < MT; m++) {
				double g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
				for (int j = 0; j < Sh::NT; j++) {
					const int col = 8 * j + 2 * q;
					g0 = fma(Ma[m][j][0] * Mb[m][j][0], Z[m][j][0], fma(Ma[m][j][1] * Mb[m][j][1], Z[m][j][1], g0));
					const double ua0 = W[m][j][0] * Mb[m][j][0], ua1 = W[m][j][1] * Mb[m][j][1];
					const double ub0 = W[m][j][0] * Ma[m][j][0], ub1 = W[m][j][1] * Ma[m][j][1];
					if (kind < 2 && col < S) {
						const double2 da = *reinterpret_cast<const double2 *>(dA + sa[m] * Sh::NP + col);
						g1 = fma(fq[col] * ua0, da.x, fma(fq[col + 1] * ua1, da.y, g1));
					}
					if (kind == 0 && col < S) {
						const double2 db = *reinterpret_cast<const double2 *>(dB + sb[m] * Sh::NP + col);
						g2 = fma(fq[col] * ub0, db.x, fma(fq[col + 1] * ub1, db.y, g2));
					}
					Mb[m][j][0] = ua0, Mb[m][j][1] = ua1;
					Ma[m][j][0] = ub0, Ma[m][j][1] = ub1;
				}
				const bool live = p0 + 8 * m + r < p.P;
				gn = live ? fma(g0, wl[m], gn) : gn;
				ga = live ? fma(g1, wl[m], ga) : ga;
				gb = live ? fma(g2, wl[m], gb) : gb;
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&empty[s]);  // images and codes of this op are consumed
			if (kind >= 1) dw_store_slice<S, MT>(cur, r, q, Ma);  // U_b: child b's op is the next one
			if (kind == 2) {
				if (a_slot < K) dw_store_slice<S, MT>(mine + (size_t)a_slot * G::SLICE, r, q, Mb);
				else {  // parked beyond the shared-memory slots: through a staging slice to HBM
					if (lane == 0) bulk_wait_read<0>();
					__syncwarp();
					dw_store_slice<S, MT>(stage_out, r, q, Mb);
					fence_async_smem();
					__syncwarp();
					if (lane == 0 && rows > 0) {
						bulk_s2g(p.spillbuf + (size_t)(a_slot - K) * CPS + (size_t)c * PS + (size_t)p0 * S, stage_out, rbytes);
						bulk_commit();
					}
				}
			}
			// one value per branch and warp, added to this warp's own rows (fixed order: deterministic)
			if (!root) gn = phb_warp_sum(gn);
			if (kind < 2) ga = phb_warp_sum(ga);
			if (kind == 0) gb = phb_warp_sum(gb);
			if (lane == 0) {
				if (!root) red_add_f64(grow + ((size_t)node * p.C + c) * p.pstride, gn);
				if (kind < 2) red_add_f64(grow + ((size_t)a_node * p.C + c) * p.pstride, ga);
				if (kind == 0) red_add_f64(grow + ((size_t)b_node * p.C + c) * p.pstride, gb);
			}
			if (++s == nstg) s = 0, ph ^= 1;
		}
	}
	if (lane == 0) bulk_wait_all<0>();
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#define DW_S 20
#define DW_MT 2

template <class T>
static int dw_upload(phbc_ctx *ctx, void **dst, const T *src, size_t n) {
	if (*dst) cudaFree(*dst);
	*dst = NULL;
	if (n == 0) return 0;
	PHBC_CHECK(cudaMalloc(dst, n * sizeof(T)));
	PHBC_CHECK(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
	PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

// descriptors of both walks from the host schedules (phb_treelikelihood.c, build_walk_schedules); the 4-state walk's cherry
// recomputation flags are not used here
int phbc_dwalk_set_schedule(phbc_ctx *ctx, const phbc_schedule *s) {
	ctx->dw_nops = 0;
	ctx->dw_codes_tp = 0;
	if (ctx->S != DW_S || s->n_post <= 0 || s->n_post != s->n_pre) return 0;
	const int n = s->n_post;
	DwPost *po = (DwPost *)calloc(n, sizeof(DwPost));
	DwPre *pr = (DwPre *)calloc(n, sizeof(DwPre));
	if (!po || !pr) {
		free(po), free(pr);
		return -3;
	}
	int k = 0, q = 0;
	for (int i = 0; i < n; i++) {
		const phbc_post_op *h = &s->post_ops[i];
		const int a_kind = h->a_kind & 0xff, b_kind = h->b_kind & 0xff;
		DwPost *d = &po[i];
		d->node = h->node, d->a_node = h->a_node, d->b_node = h->b_node;
		d->kind = (int16_t)((a_kind == PHBC_W_TIP ? 0 : 1) + (b_kind == PHBC_W_TIP ? 0 : 1));
		d->a_tip = a_kind == PHBC_W_TIP ? k++ : -1;
		d->b_tip = b_kind == PHBC_W_TIP ? k++ : -1;
		d->a_slot = (int16_t)(d->kind == 2 ? h->a_idx : -1);
		d->dst_slot = (int16_t)h->dst_slot;
		const phbc_pre_op *g = &s->pre_ops[i];
		DwPre *e = &pr[i];
		e->node = g->node, e->a_node = g->a_node, e->b_node = g->b_node;
		e->kind = g->kind, e->u_kind = g->u_kind, e->u_slot = g->u_slot, e->a_slot = g->a_slot;
		e->a_tip = g->kind != 2 ? q++ : -1;
		e->b_tip = g->kind == 0 ? q++ : -1;
		const phbc_pre_op *f = i + DW_PF < n ? &s->pre_ops[i + DW_PF] : NULL;
		e->pf_a = f && f->kind == 2 ? f->a_node : -1;
		e->pf_b = f && f->kind >= 1 ? f->b_node : -1;
	}
	int rc = (k == ctx->T && q == ctx->T) ? 0 : -1;
	if (rc) snprintf(phbc_errbuf, sizeof(phbc_errbuf), "tensor-core walk: schedules consume %d / %d of %d tips", k, q, ctx->T);
	if (!rc) rc = dw_upload(ctx, &ctx->d_dw_post, po, (size_t)n);
	if (!rc) rc = dw_upload(ctx, &ctx->d_dw_pre, pr, (size_t)n);
	free(po), free(pr);
	if (!rc) ctx->dw_nops = n;
	return rc;
}

struct DwPlan {
	int nw, TP, ntiles, nitems, grid;
	int Kpost, nstg_post, spill_post;
	int Kpre, nstg_pre, spill_pre;
	size_t smem_post, smem_pre;
};

// launch geometry: 8 consumer warps (128-pattern tiles) when that still gives every SM several items, else 4; as many
// shared-memory slots as fit beside a ring of 3 stages (2 when that avoids spilling)
static bool dw_plan(const phbc_ctx *ctx, DwPlan *pl) {
	using G = DwGeom<DW_S, DW_MT>;
	// PHB_OPT_TUNE (profiling / tests): 11 = one shared-memory slot (parked values spill to HBM), 12 = 8 consumer warps whatever the
	// pattern count, 13 = both, 14 = 4 consumer warps
	const bool one_slot = ctx->tune == 11 || ctx->tune == 13, force8 = ctx->tune == 12 || ctx->tune == 13, force4 = ctx->tune == 14;
	const size_t cap = ctx->smem_optin;
	for (int nw = force4 ? 4 : 8; nw >= 4; nw -= 4) {
		const int TP = nw * G::ROWS, ntiles = (ctx->P + TP - 1) / TP;
		const long long nitems = (long long)ntiles * ctx->C;
		if (nw == 8 && !force8 && nitems < 6LL * ctx->num_sms) continue;
		if (nitems > 0x7fffffffLL) return false;
		const size_t slice = (size_t)nw * G::SLICE * 8;
		auto fit = [&](int nstg, int stage, int fixed, int want, int *K, int *spill) -> bool {
			const size_t base = G::HDR + (size_t)nstg * stage;
			if (one_slot && want > 1) {
				*K = 1, *spill = 1;
				return base + (size_t)(fixed + 2) * slice <= cap;
			}
			if (base + (size_t)(fixed + want) * slice <= cap) {
				*K = want, *spill = 0;
				return true;
			}
			if (base + (size_t)(fixed + 1 + 1) * slice > cap) return false;
			*K = (int)((cap - base) / slice) - fixed - 1, *spill = 1;
			return *K >= 1;
		};
		pl->nw = nw, pl->TP = TP, pl->ntiles = ntiles, pl->nitems = (int)nitems;
		pl->grid = nitems < ctx->num_sms ? (int)nitems : ctx->num_sms;
		int K3, sp3, K2, sp2;
		// post-order: slices = K slots + 2 hand-over buffers (+ 1 landing slice when spilling)
		bool ok3 = fit(3, G::stage_bytes(3), 2, ctx->post_slots, &K3, &sp3), ok2 = fit(2, G::stage_bytes(3), 2, ctx->post_slots, &K2, &sp2);
		if (ok3 && (!sp3 || !ok2 || sp2)) pl->nstg_post = 3, pl->Kpost = K3, pl->spill_post = sp3;
		else if (ok2) pl->nstg_post = 2, pl->Kpost = K2, pl->spill_post = sp2;
		else continue;
		// pre-order: K slots + hand-over + two landing slices (+ 1 staging slice when spilling)
		ok3 = fit(3, G::stage_bytes(6), 3, ctx->pre_slots, &K3, &sp3), ok2 = fit(2, G::stage_bytes(6), 3, ctx->pre_slots, &K2, &sp2);
		if (ok3 && (!sp3 || !ok2 || sp2)) pl->nstg_pre = 3, pl->Kpre = K3, pl->spill_pre = sp3;
		else if (ok2) pl->nstg_pre = 2, pl->Kpre = K2, pl->spill_pre = sp2;
		else continue;
		pl->smem_post = G::HDR + (size_t)pl->nstg_post * G::stage_bytes(3) + (size_t)(pl->Kpost + 2 + pl->spill_post) * slice;
		pl->smem_pre = G::HDR + (size_t)pl->nstg_pre * G::stage_bytes(6) + (size_t)(pl->Kpre + 3 + pl->spill_pre) * slice;
		return true;
	}
	return false;
}

bool phbc_dwalk_supported(const phbc_ctx *ctx, const phbc_eval_opts *o) {
	if (ctx->S != DW_S || ctx->dw_nops <= 0 || o->scale || o->materialize_uppers || !ctx->have_eigen || o->explicit_matrices) return false;
	if (ctx->tune == 9) return false;  // PHB_OPT_TUNE 9: the level-batched kernels (A/B runs)
	if (ctx->N > 32000 || (size_t)ctx->P * DW_S * 8 > 0x7fffffffu) return false;
	DwPlan pl;
	return dw_plan(ctx, &pl);
}

// tile-major tip codes for the plan's tile size; *usable = 0 when tip partials are not single states / all ones
static int dw_prepare_codes(phbc_ctx *ctx, const DwPlan &pl, int *usable) {
	if (ctx->dw_codes_tp != pl.TP) {
		const size_t n = 2 * (size_t)pl.ntiles * ctx->T * pl.TP;
		if (n > ctx->dw_codes_bytes) {
			PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
			if (ctx->d_dw_codes) cudaFree(ctx->d_dw_codes);
			ctx->d_dw_codes = NULL, ctx->dw_codes_bytes = 0;
			PHBC_CHECK(cudaMalloc((void **)&ctx->d_dw_codes, n));
			ctx->dw_codes_bytes = n;
		}
		if (!ctx->d_dw_bad) PHBC_CHECK(cudaMalloc((void **)&ctx->d_dw_bad, sizeof(int)));
		PHBC_CHECK(cudaMemsetAsync(ctx->d_dw_bad, 0, sizeof(int), ctx->stream));
		k_dwalk_codes<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->T, ctx->P, ctx->S, pl.TP, pl.ntiles, ctx->tip_kind, ctx->d_tip_states,
		                                                                  ctx->d_tip_partials, ctx->d_post_tip_order, ctx->d_pre_tip_order,
		                                                                  ctx->d_dw_codes, ctx->d_dw_bad);
		ctx->launches++;
		int bad = 0;
		PHBC_CHECK(cudaMemcpyAsync(&bad, ctx->d_dw_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
		PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
		ctx->dw_codes_bad = bad != 0;
		ctx->dw_codes_tp = pl.TP;
	}
	*usable = !ctx->dw_codes_bad;
	return 0;
}

// 1 = declined (tips that are not single states): the caller takes the level-batched kernels
int phbc_dwalk_usable(phbc_ctx *ctx, const phbc_eval_opts *o) {
	if (!phbc_dwalk_supported(ctx, o)) return 0;
	if (ctx->tip_kind == PHBC_TIP_STATES) return 1;
	DwPlan pl;
	int usable = 0;
	if (!dw_plan(ctx, &pl) || dw_prepare_codes(ctx, pl, &usable)) return 0;
	return usable;
}

// the passes of one evaluation on the ctx stream; matrices (d_P, d_dP) are current, the lower buffers exist
int phbc_dwalk_passes(phbc_ctx *ctx, const phbc_eval_opts *o, double *result) {
	using G = DwGeom<DW_S, DW_MT>;
	DwPlan pl;
	int rc, usable = 0;
	if (!dw_plan(ctx, &pl)) {
		snprintf(phbc_errbuf, sizeof(phbc_errbuf), "tensor-core walk: no launch geometry fits");
		return -1;
	}
	if ((rc = dw_prepare_codes(ctx, pl, &usable))) return rc;
	if (!usable) {
		snprintf(phbc_errbuf, sizeof(phbc_errbuf), "tensor-core walk: tip partials are not single states");
		return -1;
	}
	if ((rc = phbc_dmma_pack_images(ctx, true, o->include_root_freqs, true))) return rc;
	const size_t per_walk = (size_t)pl.ntiles * ctx->T * pl.TP;
	DwParams p;
	memset(&p, 0, sizeof(p));
	p.img = ctx->d_dmma_img, p.lower = ctx->d_lower;
	p.T = ctx->T, p.N = ctx->N, p.C = ctx->C, p.P = ctx->P, p.ntiles = pl.ntiles, p.nitems = pl.nitems, p.root = ctx->root;
	p.nops = ctx->dw_nops, p.nw = pl.nw;
	p.freqs = ctx->d_freqs, p.weights = ctx->d_weights, p.pattern_lnl = ctx->d_pattern_lnl, p.include_root_freqs = o->include_root_freqs;
	const int threads = 32 * (pl.nw + 1);
	auto post = k_dwalk_post<DW_S, DW_MT>;
	auto pre = k_dwalk_pre<DW_S, DW_MT>;
	PHBC_CHECK(cudaFuncSetAttribute(post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_post));
	p.codes = ctx->d_dw_codes, p.ops = ctx->d_dw_post;
	p.K = pl.Kpost, p.nstg = pl.nstg_post, p.spill = pl.spill_post;
	post<<<pl.grid, threads, pl.smem_post, ctx->stream>>>(p);
	ctx->launches++;
	PHBC_CHECK(cudaGetLastError());
	if ((rc = phbc_generic_root(ctx, o, result))) return rc;
	if (!o->want_gradient) return 0;
	if (pl.spill_pre) {
		const size_t need = (size_t)(ctx->pre_slots - pl.Kpre) * ctx->C * ctx->P * DW_S * sizeof(double);
		if (need > ctx->dw_spill_bytes) {
			PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
			if (ctx->d_dw_spill) cudaFree(ctx->d_dw_spill);
			ctx->d_dw_spill = NULL, ctx->dw_spill_bytes = 0;
			PHBC_CHECK(cudaMalloc((void **)&ctx->d_dw_spill, need));
			ctx->dw_spill_bytes = need;
		}
	}
	const int pstride = pl.grid * pl.nw;
	const size_t gbytes = (size_t)ctx->N * ctx->C * pstride * sizeof(double);
	if ((rc = phbc_ensure_scratch(ctx, gbytes))) return rc;
	PHBC_CHECK(cudaMemsetAsync(ctx->d_scratch, 0, gbytes, ctx->stream));
	PHBC_CHECK(cudaFuncSetAttribute(pre, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_pre));
	p.codes = ctx->d_dw_codes + per_walk, p.ops = ctx->d_dw_pre;
	p.K = pl.Kpre, p.nstg = pl.nstg_pre, p.spill = pl.spill_pre, p.spillbuf = ctx->d_dw_spill;
	p.pstride = pstride, p.gacc = ctx->d_scratch;
	pre<<<pl.grid, threads, pl.smem_pre, ctx->stream>>>(p);
	ctx->launches++;
	PHBC_CHECK(cudaGetLastError());
	(void)G::HDR;
	return phbc_gradient_from_partials(ctx, pstride, result);
}
