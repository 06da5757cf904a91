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
//                  M_n = L_n P_n^T on the tensor pipe, result to shared memory (next op / parked) and from the accumulators to HBM.
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

#include <type_traits>

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
	int turns;               // warp pairs of a scheduler take turns on the tensor pipe
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
// red.global.add.f64 when addr is not null, as ONE predicated instruction (no branch: the surrounding DMMAs stay in one basic block)
__device__ __forceinline__ void red_add_f64_if(double *addr, double v) {
	asm volatile("{\n.reg .pred p;\nsetp.ne.u64 p, %0, 0;\n@p red.global.add.f64 [%0], %1;\n}" ::"l"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// shared-memory geometry, the same arithmetic on host and device
template <int S, int MT, int NWMAX_>
struct DwGeom {
	using Sh = DmmaShape<S>;
	static_assert(S % 8 == 4, "row-major tiles of stride S are bank-conflict free (and need no k padding) only for S = 4 (mod 8)");
	static constexpr int ROWS = MT * 8;                 // pattern rows per warp
	static constexpr int SLICE = ROWS * S;              // doubles per warp and tile slot
	static constexpr int IMGB = Sh::IMG * 8;            // bytes per matrix image
	static constexpr int NWMAX = NWMAX_;                // consumer warps of the widest launch
	static constexpr int TPMAX = NWMAX * ROWS;          // patterns per tile then
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
// Shared-memory plan of both kernels
//   [0, HDR)        mbarriers (image ring full / empty, landing full / empty), frequency vectors
//   image ring      nstg stages of { descriptor | tip codes of a, b | matrix images }
//   tiles           slot-major: tile t = [nw warps][ROWS rows][S], so that a tile is ONE contiguous block of the [pattern][state]
//                   layout in HBM (tile-wide bulk copies by the producer) and a warp's slice of it is contiguous too (per-warp stores)
// ---------------------------------------------------------------------------------------------
struct DwBars {
	uint64_t imgfull[4], imgempty[4], landfull, landempty;
	uint64_t turn[8];  // tensor-pipe turns of the warp pairs that share a scheduler: [pair][member]
};

// The two consumer warps of a scheduler (warps w and w + 4) take TURNS on the tensor pipe: left alone they run in lockstep -- both
// issue DMMAs at half rate, then both run their scalar epilogue with the pipe idle (ncu: 48 % pipe utilisation, `wait` stalls on
// every DMMA).  With the turn, one warp's DMMA phase runs at full rate inside the other's epilogue / prologue (the ping-pong
// schedule of warp-specialised GEMMs).  Member 0 goes first; each side waits for the other's hand-over of the SAME op.
struct DwTurn {
	uint64_t *mine, *other;
	uint32_t ph;
	bool on;
	__device__ __forceinline__ void init(DwBars *bars, int warp, int nw, int turns) {
		on = nw == 8 && turns;  // warps w and w + 4 share a scheduler
		const int pair = warp & 3, member = warp >> 2;
		mine = &bars->turn[2 * pair + (member & 1)], other = &bars->turn[2 * pair + ((member & 1) ^ 1)];
		ph = member == 0 ? 1 : 0;  // member 0's first wait passes on the fresh barrier
	}
	__device__ __forceinline__ void acquire() {
		if (on) {
			mbar_wait(mine, ph);
			ph ^= 1;
		}
	}
	__device__ __forceinline__ void release(int lane) {
		if (on) {
			__syncwarp();
			if (lane == 0) mbar_arrive(other);
		}
	}
};

// ---------------------------------------------------------------------------------------------
// producer warp (one lane): every asynchronous load of the CTA.  Images, codes and the descriptor of op g run nstg - 1 ops ahead
// through the ring; the pre-order pass also gets the message tiles of the children (and a spilled U) of op h into the single landing
// buffers as soon as every consumer warp has taken op h - 1's copy into registers (landempty).
// ---------------------------------------------------------------------------------------------
template <int S, int MT, int NWM, bool PRE>
__device__ __forceinline__ void dwalk_producer(const DwParams &p, unsigned char *smraw, DwBars *bars, double *tiles) {
	using G = DwGeom<S, MT, NWM>;
	using Sh = DmmaShape<S>;
	constexpr int NIMG = PRE ? 6 : 3;
	const int stgb = G::stage_bytes(NIMG);
	const int TP = p.nw * G::ROWS, nstg = p.nstg, LA = nstg - 1;
	const size_t dimg = (size_t)p.N * p.C * Sh::IMG;
	const size_t PS = (size_t)p.P * S, CPS = (size_t)p.C * PS;
	const size_t tile_d = (size_t)p.nw * G::SLICE;  // doubles per tile
	double *cur = tiles + (size_t)p.K * tile_d, *landA = cur + tile_d, *landB = landA + tile_d;
	const int nmine = blockIdx.x < p.nitems ? (p.nitems - 1 - blockIdx.x) / gridDim.x + 1 : 0;
	const long long total = (long long)nmine * p.nops;
	int s = 0, ik = 0, iitem = blockIdx.x, lk = 0, litem = blockIdx.x;
	uint32_t eph = 1, lfree = 0;  // a fresh barrier passes a wait on the phase "before" its first one
	for (long long g = 0; g < total + LA; g++) {
		if (PRE && g >= LA) {  // landing of op h = g - LA
			const long long h = g - LA;
			if (h > 0) {
				mbar_wait(&bars->landempty, lfree);
				lfree ^= 1;
			}
			const DwPre *d = reinterpret_cast<const DwPre *>(p.ops) + lk;
			const int kind = d->kind, a_node = d->a_node, b_node = d->b_node;
			const bool lu = d->u_kind == PHBC_W_SLOT && d->u_slot >= p.K;
			// landfull completes once per op, loads or not: a warp cannot start op h + 1 before every warp has released op h's landing
			// buffers, so the arrivals of two ops never meet in one phase of landempty
			if (!(kind >= 1 || lu)) mbar_arrive(&bars->landfull);
			else {
				const int tile = litem / p.C, c = litem - tile * p.C;
				const int p0 = tile * TP;
				const uint32_t bytes = (uint32_t)min(p.P - p0, TP) * S * 8;
				const size_t off = (size_t)c * PS + (size_t)p0 * S;
				mbar_expect_tx(&bars->landfull, ((kind == 2) + (kind >= 1) + (lu ? 1 : 0)) * bytes);
				if (kind == 2) bulk_g2s(landA, p.lower + (size_t)(a_node - p.T) * CPS + off, bytes, &bars->landfull);
				if (kind >= 1) bulk_g2s(landB, p.lower + (size_t)(b_node - p.T) * CPS + off, bytes, &bars->landfull);
				if (lu) bulk_g2s(cur, p.spillbuf + (size_t)(d->u_slot - p.K) * CPS + off, bytes, &bars->landfull);
			}
			{  // the tiles of the op DW_PF ops later, hinted into L2 so that their landing is an L2 hit
				const int tile = litem / p.C, c = litem - tile * p.C, p0 = tile * TP;
				const uint32_t bytes = (uint32_t)min(p.P - p0, TP) * S * 8;
				const size_t off = (size_t)c * PS + (size_t)p0 * S;
				if (d->pf_a >= 0) bulk_prefetch_l2(p.lower + (size_t)(d->pf_a - p.T) * CPS + off, bytes);
				if (d->pf_b >= 0) bulk_prefetch_l2(p.lower + (size_t)(d->pf_b - p.T) * CPS + off, bytes);
			}
			if (++lk == p.nops) lk = 0, litem += gridDim.x;
		}
		if (g < total) {  // ring stage of op g
			const int tile = iitem / p.C, c = iitem - tile * p.C;
			const uint8_t *codes = p.codes + (size_t)tile * p.T * TP;
			unsigned char *stg = smraw + G::HDR + (size_t)s * stgb;
			double *im = reinterpret_cast<double *>(stg + G::IMG_OFF);
			mbar_wait(&bars->imgempty[s], eph);
			if (PRE) {
				const DwPre *d = reinterpret_cast<const DwPre *>(p.ops) + ik;
				const int node = d->node, a_node = d->a_node, b_node = d->b_node, a_tip = d->a_tip, b_tip = d->b_tip;
				const bool root = d->u_kind == PHBC_W_ROOT;
				const int nimg = (root ? 0 : 2) + (a_tip >= 0 ? 2 : 0) + (b_tip >= 0 ? 2 : 0);
				mbar_expect_tx(&bars->imgfull[s], (uint32_t)(sizeof(DwPre) + nimg * G::IMGB + ((a_tip >= 0) + (b_tip >= 0)) * TP));
				bulk_g2s(stg, d, sizeof(DwPre), &bars->imgfull[s]);
				if (!root) {
					bulk_g2s(im, p.img + ((size_t)node * p.C + c) * Sh::IMG, G::IMGB, &bars->imgfull[s]);
					bulk_g2s(im + Sh::IMG, p.img + dimg + ((size_t)node * p.C + c) * Sh::IMG, G::IMGB, &bars->imgfull[s]);
				}
				if (a_tip >= 0) {
					bulk_g2s(im + 2 * Sh::IMG, p.img + ((size_t)a_node * p.C + c) * Sh::IMG, G::IMGB, &bars->imgfull[s]);
					bulk_g2s(im + 3 * Sh::IMG, p.img + dimg + ((size_t)a_node * p.C + c) * Sh::IMG, G::IMGB, &bars->imgfull[s]);
					bulk_g2s(stg + G::CODE_OFF, codes + (size_t)a_tip * TP, TP, &bars->imgfull[s]);
				}
				if (b_tip >= 0) {
					bulk_g2s(im + 4 * Sh::IMG, p.img + ((size_t)b_node * p.C + c) * Sh::IMG, G::IMGB, &bars->imgfull[s]);
					bulk_g2s(im + 5 * Sh::IMG, p.img + dimg + ((size_t)b_node * p.C + c) * Sh::IMG, G::IMGB, &bars->imgfull[s]);
					bulk_g2s(stg + G::CODE_OFF + G::TPMAX, codes + (size_t)b_tip * TP, TP, &bars->imgfull[s]);
				}
			} else {
				const DwPost *d = reinterpret_cast<const DwPost *>(p.ops) + ik;
				const int node = d->node, a_node = d->a_node, b_node = d->b_node, a_tip = d->a_tip, b_tip = d->b_tip;
				const int nimg = 1 + (a_tip >= 0) + (b_tip >= 0);
				mbar_expect_tx(&bars->imgfull[s], (uint32_t)(sizeof(DwPost) + nimg * G::IMGB + ((a_tip >= 0) + (b_tip >= 0)) * TP));
				bulk_g2s(stg, d, sizeof(DwPost), &bars->imgfull[s]);
				bulk_g2s(im, p.img + ((size_t)node * p.C + c) * Sh::IMG, G::IMGB, &bars->imgfull[s]);
				if (a_tip >= 0) {
					bulk_g2s(im + Sh::IMG, p.img + ((size_t)a_node * p.C + c) * Sh::IMG, G::IMGB, &bars->imgfull[s]);
					bulk_g2s(stg + G::CODE_OFF, codes + (size_t)a_tip * TP, TP, &bars->imgfull[s]);
				}
				if (b_tip >= 0) {
					bulk_g2s(im + 2 * Sh::IMG, p.img + ((size_t)b_node * p.C + c) * Sh::IMG, G::IMGB, &bars->imgfull[s]);
					bulk_g2s(stg + G::CODE_OFF + G::TPMAX, codes + (size_t)b_tip * TP, TP, &bars->imgfull[s]);
				}
			}
			if (++s == nstg) s = 0, eph ^= 1;
			if (++ik == p.nops) ik = 0, iitem += gridDim.x;
		}
	}
}

// D-layout accumulators (row 8 m + r, columns 8 j + 2 q, + 1) to this lane's places of a row-major slice (`at` = slice + r S + 2 q);
// padding columns are dropped
template <int S, int MT>
__device__ __forceinline__ void dw_store_slice(double *at, int q, const double (&v)[MT][DmmaShape<S>::NT][2]) {
#pragma unroll
	for (int m = 0; m < MT; m++)
#pragma unroll
		for (int j = 0; j < DmmaShape<S>::NT; j++)
			if (8 * j + 6 < S || 8 * j + 2 * q < S) *reinterpret_cast<double2 *>(at + 8 * m * S + 8 * j) = make_double2(v[m][j][0], v[m][j][1]);
}

// ---------------------------------------------------------------------------------------------
// post-order pass
// ---------------------------------------------------------------------------------------------
template <int S, int MT, int NWM>
__global__ void __launch_bounds__(32 * (NWM + 1), 2) k_dwalk_post(const DwParams p) {
	using G = DwGeom<S, MT, NWM>;
	using Sh = DmmaShape<S>;
	extern __shared__ __align__(128) unsigned char smraw[];
	DwBars *bars = reinterpret_cast<DwBars *>(smraw);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int nw = p.nw, nstg = p.nstg, K = p.K;
	const int stgb = G::stage_bytes(3);
	const int ntile = K + 1;
	const size_t tile_d = (size_t)nw * G::SLICE;
	double *tiles = reinterpret_cast<double *>(smraw + G::HDR + (size_t)nstg * stgb);
	if (threadIdx.x == 0) {
		for (int i = 0; i < nstg; i++) mbar_init(&bars->imgfull[i], 1), mbar_init(&bars->imgempty[i], nw);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	for (int i = threadIdx.x; i < ntile * (int)tile_d; i += blockDim.x) tiles[i] = 0.0;  // rows past P keep finite values
	__syncthreads();
	if (warp == nw) {
		if (lane == 0) dwalk_producer<S, MT, NWM, false>(p, smraw, bars, tiles);
		return;
	}
	const int r = lane >> 2, q = lane & 3;
	double *mine = tiles + (size_t)warp * G::SLICE;  // this warp's slice of tile 0
	double *cur = mine + (size_t)K * tile_d;         // hand-over buffer: the preceding op's result
	const int TP = nw * G::ROWS;
	const size_t PS = (size_t)p.P * S, CPS = (size_t)p.C * PS;
	// Fragment row r of a DMMA works on pattern row rr of the warp's slice, rr = r with its two low bits swapped.  Rows are independent,
	// so any fixed assignment is valid; this one makes the 128-bit accesses in accumulator layout conflict-free: a quarter-warp (fragment
	// rows 2k, 2k + 1) then touches slice rows two apart, whose 64-byte quads are 320 bytes = 16 banks (mod 32) apart, where adjacent rows
	// (160 bytes = 8 banks) overlap in half of their banks (ncu: 8 wavefronts per STS.128 / LDS.128 instead of 4).  The 64-bit
	// A-fragment reads see the same set of rows per half-warp as before and stay conflict-free.
	const int rr = (r & 4) | ((r & 1) << 1) | ((r >> 1) & 1);
	const int lofA = rr * S + q, lofD = rr * S + 2 * q;  // this lane's place in a slice, A-fragment / accumulator layout
	const int bof = r * Sh::LD + q;
	int s = 0;
	uint32_t ph = 0;
	for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
		const int tile = item / p.C, c = item - tile * p.C;
		const int p0 = tile * TP + warp * G::ROWS;
		const int rows = min(max(p.P - p0, 0), G::ROWS);
		double *rowbase = p.lower + (size_t)c * PS + (size_t)p0 * S;  // + (node - T) CPS
		for (int k = 0; k < p.nops; k++) {
			const unsigned char *stg = smraw + G::HDR + (size_t)s * stgb;
			mbar_wait(&bars->imgfull[s], ph);
			const DwPost *d = reinterpret_cast<const DwPost *>(stg);
			const int kind = d->kind, a_slot = d->a_slot, dst_slot = d->dst_slot, node = d->node, a_node = d->a_node;  // the stage is recycled after the arrive below
			const double *mN = reinterpret_cast<const double *>(stg + G::IMG_OFF), *mA = mN + Sh::IMG, *mB = mA + Sh::IMG;
			// A fragments: L_n = M_a o M_b; an operand is a row of a tile slice, the column a tip's state selects in its image, or -- a
			// child parked beyond the shared-memory slots -- the row of its message in HBM, which this warp wrote itself
			const bool a_far = kind == 2 && a_slot >= K && rows > 0;
			const double *ta = a_far ? rowbase + (size_t)(a_node - p.T) * CPS + q : (kind == 2 && a_slot < K ? mine + (size_t)a_slot * tile_d : cur) + lofA;
			double a[MT][Sh::KT];
#pragma unroll
			for (int m = 0; m < MT; m++) {
				const int row = warp * G::ROWS + 8 * m + rr;
				const double *pa = kind < 2 ? mA + stg[G::CODE_OFF + row] * Sh::NP + q : ta + (a_far ? min(8 * m + rr, rows - 1) * S : 8 * m * S);
				const double *pb = kind == 0 ? mB + stg[G::CODE_OFF + G::TPMAX + row] * Sh::NP + q : cur + lofA + 8 * m * S;
#pragma unroll
				for (int tt = 0; tt < Sh::KT; tt++) a[m][tt] = pa[4 * tt] * pb[4 * tt];
			}
			// M_n = L_n P_n^T
			double acc[MT][Sh::NT][2];
#pragma unroll
			for (int m = 0; m < MT; m++)
#pragma unroll
				for (int j = 0; j < Sh::NT; j++) acc[m][j][0] = acc[m][j][1] = 0.0;
#pragma unroll
			for (int tt = 0; tt < Sh::KT; tt++)
#pragma unroll
				for (int j = 0; j < Sh::NT; j++) {
					const double bN = mN[bof + j * 8 * Sh::LD + 4 * tt];
#pragma unroll
					for (int m = 0; m < MT; m++) dmma_m8n8k4(acc[m][j][0], acc[m][j][1], a[m][tt], bN);
				}
			__syncwarp();  // every lane has its A fragments out of the hand-over buffer / the slot
			if (lane == 0) mbar_arrive(&bars->imgempty[s]);  // images and codes of this op are consumed
			// result: to the parked slot or the hand-over buffer for the ops to come, and straight from the accumulators to its row in HBM
			// (a quad writes 64 contiguous bytes of a row; streaming stores: the pre-order pass reads the row 50 GB later)
			dw_store_slice<S, MT>(((dst_slot >= 0 && dst_slot < K) ? mine + (size_t)dst_slot * tile_d : cur) + lofD, q, acc);
			double *grow = rowbase + (size_t)(node - p.T) * CPS + lofD;
#pragma unroll
			for (int m = 0; m < MT; m++)
#pragma unroll
				for (int j = 0; j < Sh::NT; j++)
					if ((8 * j + 6 < S || 8 * j + 2 * q < S) && 8 * m + rr < rows) __stcs(reinterpret_cast<double2 *>(grow + 8 * m * S + 8 * j), make_double2(acc[m][j][0], acc[m][j][1]));
			__syncwarp();  // the result is visible to the lanes that read it in A-fragment layout
			if (++s == nstg) s = 0, ph ^= 1;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// pre-order pass with branch gradients
// ---------------------------------------------------------------------------------------------
template <int S, int MT, int NWM>
__global__ void __launch_bounds__(32 * (NWM + 1), 1) k_dwalk_pre(const DwParams p) {
	using G = DwGeom<S, MT, NWM>;
	using Sh = DmmaShape<S>;
	extern __shared__ __align__(128) unsigned char smraw[];
	DwBars *bars = reinterpret_cast<DwBars *>(smraw);
	double *fq = reinterpret_cast<double *>(smraw + 256), *wroot = fq + Sh::NP, *zeros = wroot + Sh::NP;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int nw = p.nw, nstg = p.nstg, K = p.K;
	const int stgb = G::stage_bytes(6);
	const int ntile = K + 3 + (p.spill ? 1 : 0);
	const size_t tile_d = (size_t)nw * G::SLICE;
	double *tiles = reinterpret_cast<double *>(smraw + G::HDR + (size_t)nstg * stgb);
	if (threadIdx.x == 0) {
		for (int i = 0; i < nstg; i++) mbar_init(&bars->imgfull[i], 1), mbar_init(&bars->imgempty[i], nw);
		mbar_init(&bars->landfull, 1), mbar_init(&bars->landempty, nw);
		for (int i = 0; i < 8; i++) mbar_init(&bars->turn[i], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	for (int i = threadIdx.x; i < Sh::NP; i += blockDim.x) {
		const double f = i < S ? p.freqs[i] : 0.0;
		fq[i] = i < S ? (p.include_root_freqs ? 1.0 : f) : 0.0;
		wroot[i] = i < S ? (p.include_root_freqs ? f : 1.0) : 0.0;
		zeros[i] = 0.0;
	}
	for (int i = threadIdx.x; i < ntile * (int)tile_d; i += blockDim.x) tiles[i] = 0.0;
	__syncthreads();
	if (warp == nw) {
		if (lane == 0) dwalk_producer<S, MT, NWM, true>(p, smraw, bars, tiles);
		return;
	}
	const int r = lane >> 2, q = lane & 3;
	double *mine = tiles + (size_t)warp * G::SLICE;
	double *cur = mine + (size_t)K * tile_d, *landA = cur + tile_d, *landB = landA + tile_d, *stage_out = landB + tile_d;
	const int TP = nw * G::ROWS;
	const size_t PS = (size_t)p.P * S, CPS = (size_t)p.C * PS;
	const int rr = (r & 4) | ((r & 1) << 1) | ((r >> 1) & 1);  // pattern row of fragment row r (see k_dwalk_post)
	const int lofA = rr * S + q, lofD = rr * S + 2 * q;
	const int bof = r * Sh::LD + q;
	int s = 0;
	uint32_t ph = 0, lph = 0;
	DwTurn turn;
	turn.init(bars, warp, nw, p.turns);
	double pend_n = 0.0, pend_a = 0.0, pend_b = 0.0;  // the preceding op's per-lane branch sums and where they go
	double *pend_tn = nullptr, *pend_ta = nullptr, *pend_tb = nullptr;
	double *grow = p.gacc + (size_t)blockIdx.x * nw + warp;
	for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
		const int tile = item / p.C, c = item - tile * p.C;
		const int p0 = tile * TP + warp * G::ROWS;
		const int rows = min(max(p.P - p0, 0), G::ROWS);
		const uint32_t rbytes = (uint32_t)rows * S * 8;
		double wl[MT];
#pragma unroll
		for (int m = 0; m < MT; m++) {
			const int pp = p0 + 8 * m + rr;
			wl[m] = pp < p.P ? __ldg(p.weights + pp) / exp(__ldg(p.pattern_lnl + pp)) : 0.0;
		}
		double *gitem = grow + (size_t)c * p.pstride;  // + node C pstride
		for (int k = 0; k < p.nops; k++) {
			const unsigned char *stg = smraw + G::HDR + (size_t)s * stgb;
			mbar_wait(&bars->imgfull[s], ph);
			const DwPre *d = reinterpret_cast<const DwPre *>(stg);
			const int kind = d->kind, u_kind = d->u_kind, u_slot = d->u_slot, a_slot = d->a_slot;
			const int node = d->node, a_node = d->a_node, b_node = d->b_node;  // the stage is recycled after the arrive below
			const bool root = u_kind == PHBC_W_ROOT;
			const double *mP = reinterpret_cast<const double *>(stg + G::IMG_OFF), *mZ = mP + Sh::IMG;
			const double *tA = mZ + Sh::IMG, *dA = tA + Sh::IMG, *tB = dA + Sh::IMG, *dB = tB + Sh::IMG;
			mbar_wait(&bars->landfull, lph);  // the message tiles of the children (and a spilled U) have landed; completes for every op
			lph ^= 1;
			// A fragments of U_n
			const double *usrc = ((u_kind == PHBC_W_SLOT && u_slot < K) ? mine + (size_t)u_slot * tile_d : cur) + lofA;
			double u[MT][Sh::KT];
#pragma unroll
			for (int m = 0; m < MT; m++)
#pragma unroll
				for (int tt = 0; tt < Sh::KT; tt++) u[m][tt] = usrc[8 * m * S + 4 * tt];
			// the children's messages in accumulator layout: a row of the landing tile, or the column a tip's state selects
			int sa[MT], sb[MT];
			double Ma[MT][Sh::NT][2], Mb[MT][Sh::NT][2];
#pragma unroll
			for (int m = 0; m < MT; m++) {
				const int row = warp * G::ROWS + 8 * m + rr;
				sa[m] = stg[G::CODE_OFF + row] * Sh::NP + 2 * q;
				sb[m] = stg[G::CODE_OFF + G::TPMAX + row] * Sh::NP + 2 * q;
				const double *pa = kind < 2 ? tA + sa[m] : landA + lofD + 8 * m * S;
				const double *pb = kind == 0 ? tB + sb[m] : landB + lofD + 8 * m * S;
#pragma unroll
				for (int j = 0; j < Sh::NT; j++) {
					double2 va = make_double2(0.0, 0.0), vb = va;
					if (8 * j + 6 < S || 8 * j + 2 * q < S) {
						va = *reinterpret_cast<const double2 *>(pa + 8 * j);
						vb = *reinterpret_cast<const double2 *>(pb + 8 * j);
					}
					Ma[m][j][0] = va.x, Ma[m][j][1] = va.y;
					Mb[m][j][0] = vb.x, Mb[m][j][1] = vb.y;
				}
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&bars->landempty);  // landing tiles and the hand-over buffer are in registers: the next op's may land
			// the derivative columns of tip children (a zero vector for internal ones: their branches are reduced at their own ops)
			const double *pda[MT], *pdb[MT];
#pragma unroll
			for (int m = 0; m < MT; m++) pda[m] = kind < 2 ? dA + sa[m] : zeros + 2 * q, pdb[m] = kind == 0 ? dB + sb[m] : zeros + 2 * q;
			turn.acquire();
			// W = U_n P_n^T and Z = U_n (pi o dP_n), one n-tile at a time (four accumulator chains in flight); the element-wise work of a
			// finished tile -- n's own branch sum_j L[j] Z[j] (adjoint form), U_a = W o M_b, U_b = W o M_a, the tips' branches -- is issued
			// under the NEXT tile's DMMAs, and so is the shuffle chain that sums the PREVIOUS op's three branch terms over the warp: the
			// FP64 pipe serves both kinds of work, and outside this warp's turn it belongs to the partner's DMMAs.
			// Butterfly: lanes 16.. take over (ga, gb), lanes ..15 (gn, 0), then halves again; lane 0 ends with gn, lane 16 ga, lane 24 gb --
			// one value per branch and warp, added to this warp's own rows (fixed order: deterministic).
			double g0[MT], g1[MT], g2[MT];
#pragma unroll
			for (int m = 0; m < MT; m++) g0[m] = g1[m] = g2[m] = 0.0;
			auto tile_dmma = [&](int j, double (&W)[MT][2], double (&Z)[MT][2]) {
#pragma unroll
				for (int m = 0; m < MT; m++) W[m][0] = W[m][1] = Z[m][0] = Z[m][1] = 0.0;
#pragma unroll
				for (int tt = 0; tt < Sh::KT; tt++) {
					const double bP = mP[bof + j * 8 * Sh::LD + 4 * tt], bZ = mZ[bof + j * 8 * Sh::LD + 4 * tt];
#pragma unroll
					for (int m = 0; m < MT; m++) {
						dmma_m8n8k4(W[m][0], W[m][1], u[m][tt], bP);
						dmma_m8n8k4(Z[m][0], Z[m][1], u[m][tt], bZ);
					}
				}
			};
			// KIND (compile time): the number of internal children; their derivative columns are not read (shared memory runs at more
			// than half of its bandwidth beside the tensor pipe, every wavefront counts)
			auto combine = [&](auto KIND, int j, const double (&W)[MT][2], const double (&Z)[MT][2]) {
				constexpr int kd = decltype(KIND)::value;
				const bool colok = 8 * j + 6 < S || 8 * j + 2 * q < S;
#pragma unroll
				for (int m = 0; m < MT; m++) {
					g0[m] = fma(Ma[m][j][0] * Mb[m][j][0], Z[m][0], fma(Ma[m][j][1] * Mb[m][j][1], Z[m][1], g0[m]));
					const double ua0 = W[m][0] * Mb[m][j][0], ua1 = W[m][1] * Mb[m][j][1];
					const double ub0 = W[m][0] * Ma[m][j][0], ub1 = W[m][1] * Ma[m][j][1];
					Mb[m][j][0] = ua0, Mb[m][j][1] = ua1;
					Ma[m][j][0] = ub0, Ma[m][j][1] = ub1;
					if (kd < 2) {
						const double2 da = *reinterpret_cast<const double2 *>((colok ? pda[m] : zeros) + 8 * j);
						g1[m] = fma(ua0, da.x, fma(ua1, da.y, g1[m]));
					}
					if (kd == 0) {
						const double2 db = *reinterpret_cast<const double2 *>((colok ? pdb[m] : zeros) + 8 * j);
						g2[m] = fma(ub0, db.x, fma(ub1, db.y, g2[m]));
					}
				}
			};
			const bool hi = lane & 16, hi2 = lane & 8;
			double bx, by, bz;
			auto bf1 = [&]() {
				bx = hi ? pend_a : pend_n, by = hi ? pend_b : 0.0;
				bx += __shfl_xor_sync(0xffffffffu, hi ? pend_n : pend_a, 16), by += __shfl_xor_sync(0xffffffffu, hi ? 0.0 : pend_b, 16);
			};
			auto bf2 = [&]() {
				bz = hi2 ? by : bx;
				bz += __shfl_xor_sync(0xffffffffu, hi2 ? bx : by, 8);
				bz += __shfl_xor_sync(0xffffffffu, bz, 4);
			};
			auto bf3 = [&]() {
				bz += __shfl_xor_sync(0xffffffffu, bz, 2);
				bz += __shfl_xor_sync(0xffffffffu, bz, 1);
				red_add_f64_if(lane == 0 ? pend_tn : (lane == 16 ? pend_ta : (lane == 24 ? pend_tb : nullptr)), bz);
			};
			static_assert(Sh::NT == 3, "the software pipeline below is written out for three n-tiles");
			auto dphase = [&](auto KIND) {
				double W0[MT][2], Z0[MT][2], W1[MT][2], Z1[MT][2];
				bf1();
				tile_dmma(0, W0, Z0);
				bf2();
				tile_dmma(1, W1, Z1);
				combine(KIND, 0, W0, Z0);
				bf3();
				tile_dmma(2, W0, Z0);
				combine(KIND, 1, W1, Z1);
				combine(KIND, 2, W0, Z0);
			};
			if (!root) {
				if (kind == 2) dphase(std::integral_constant<int, 2>{});
				else if (kind == 1) dphase(std::integral_constant<int, 1>{});
				else dphase(std::integral_constant<int, 0>{});
			} else {  // once per item: the general form (derivative columns of internal children read as zeros)
				bf1(), bf2(), bf3();
#pragma unroll
				for (int j = 0; j < Sh::NT; j++) {
					double W[MT][2], Z[MT][2];
#pragma unroll
					for (int m = 0; m < MT; m++) W[m][0] = wroot[8 * j + 2 * q], W[m][1] = wroot[8 * j + 2 * q + 1], Z[m][0] = Z[m][1] = 0.0;
					combine(std::integral_constant<int, 0>{}, j, W, Z);
				}
			}
			turn.release(lane);
			if (kind >= 1) dw_store_slice<S, MT>(cur + lofD, q, Ma);  // U_b: child b's op is the next one
			if (kind == 2) {
				if (a_slot < K) dw_store_slice<S, MT>(mine + (size_t)a_slot * tile_d + lofD, q, Mb);
				else {  // parked beyond the shared-memory slots: through a staging slice to HBM; complete before the producer may read it back
					dw_store_slice<S, MT>(stage_out + lofD, q, Mb);
					fence_async_smem();
					__syncwarp();
					if (lane == 0) {
						if (rows > 0) bulk_s2g(p.spillbuf + (size_t)(a_slot - K) * CPS + (size_t)c * PS + (size_t)p0 * S, stage_out, rbytes);
						bulk_commit();
						bulk_wait_all<0>();
					}
				}
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&bars->imgempty[s]);  // images and codes of this op are consumed
			pend_n = pend_a = pend_b = 0.0;
#pragma unroll
			for (int m = 0; m < MT; m++) pend_n = fma(g0[m], wl[m], pend_n), pend_a = fma(g1[m], wl[m], pend_a), pend_b = fma(g2[m], wl[m], pend_b);
			pend_tn = root ? nullptr : gitem + (size_t)node * p.C * p.pstride;
			pend_ta = kind < 2 ? gitem + (size_t)a_node * p.C * p.pstride : nullptr;
			pend_tb = kind == 0 ? gitem + (size_t)b_node * p.C * p.pstride : nullptr;
			if (++s == nstg) s = 0, ph ^= 1;
		}
	}
	{  // the last op's sums
		const bool hi = lane & 16, hi2 = lane & 8;
		double x = hi ? pend_a : pend_n, y = hi ? pend_b : 0.0;
		x += __shfl_xor_sync(0xffffffffu, hi ? pend_n : pend_a, 16), y += __shfl_xor_sync(0xffffffffu, hi ? 0.0 : pend_b, 16);
		double z = hi2 ? y : x;
		z += __shfl_xor_sync(0xffffffffu, hi2 ? x : y, 8);
		z += __shfl_xor_sync(0xffffffffu, z, 4);
		z += __shfl_xor_sync(0xffffffffu, z, 2);
		z += __shfl_xor_sync(0xffffffffu, z, 1);
		red_add_f64_if(lane == 0 ? pend_tn : (lane == 16 ? pend_ta : (lane == 24 ? pend_tb : nullptr)), z);
	}
	if (lane == 0) bulk_wait_all<0>();
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#define DW_S 20

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
	int nw, TP, ntiles, nitems, grid, ctas_post;
	int Kpost, nstg_post, spill_post;
	int Kpre, nstg_pre, spill_pre;
	size_t smem_post, smem_pre;
};

// launch geometry: 8 consumer warps (128-pattern tiles) when that still gives every SM several items, else 4; as many
// shared-memory slots as fit beside a ring of 3 stages (2 when that avoids spilling)
template <class G>
static bool dw_plan_g(const phbc_ctx *ctx, DwPlan *pl) {
	// PHB_OPT_TUNE (profiling / tests): 11 = one shared-memory slot (parked values spill to HBM), 12 = 8 consumer warps whatever the
	// pattern count, 13 = both, 14 = 4 consumer warps
	const bool one_slot = ctx->tune == 11 || ctx->tune == 13, force8 = ctx->tune == 12 || ctx->tune == 13, force4 = ctx->tune == 14;
	const size_t cap = ctx->smem_optin;
	for (int nw = force4 ? G::NWMAX / 2 : G::NWMAX; nw >= G::NWMAX / 2; nw -= G::NWMAX / 2) {
		const int TP = nw * G::ROWS, ntiles = (ctx->P + TP - 1) / TP;
		const long long nitems = (long long)ntiles * ctx->C;
		if (nw == G::NWMAX && !force8 && nitems < ctx->num_sms) continue;  // measured: 8 warps win down to 25k patterns (profiles/r2_l_tune_c4.jsonl)
		if (nitems > 0x7fffffffLL) return false;
		const size_t slice = (size_t)nw * G::SLICE * 8;
		// `extra` = tiles a spilling launch needs on top of the fixed ones (pre-order: a staging tile for the bulk store of a parked U)
		auto fit = [&](size_t cap, int nstg, int stage, int fixed, int extra, int want, int *K, int *spill) -> bool {
			const size_t base = G::HDR + (size_t)nstg * stage;
			if (one_slot && want > 1) {
				*K = 1, *spill = 1;
				return base + (size_t)(fixed + extra + 1) * slice <= cap;
			}
			if (base + (size_t)(fixed + want) * slice <= cap) {
				*K = want, *spill = 0;
				return true;
			}
			if (base + (size_t)(fixed + extra + 1) * slice > cap) return false;
			*K = (int)((cap - base) / slice) - fixed - extra, *spill = 1;
			return *K >= 1;
		};
		pl->nw = nw, pl->TP = TP, pl->ntiles = ntiles, pl->nitems = (int)nitems;
		pl->grid = nitems < ctx->num_sms ? (int)nitems : ctx->num_sms;
		int K3, sp3, K2, sp2;
		// post-order: tiles = K slots + the hand-over buffer; TWO CTAs per SM (16 consumer warps hide the scalar phases of the ops and the
		// latency of re-reading a child parked beyond K from its message row in HBM), so each gets half of the SM's shared memory
		const size_t cap2 = (ctx->smem_sm - 2 * 1024) / 2;
		pl->ctas_post = ctx->tune == 19 ? 1 : 2;  // PHB_OPT_TUNE 19: one CTA per SM (A/B runs)
		bool ok3 = fit(pl->ctas_post == 2 ? cap2 : cap, 3, G::stage_bytes(3), 1, 0, ctx->post_slots, &K3, &sp3), ok2 = false;
		if (ok3) pl->nstg_post = 3, pl->Kpost = K3, pl->spill_post = sp3;
		else continue;
		// pre-order: K slots + hand-over + two landing slices (+ 1 staging slice when spilling)
		ok3 = fit(cap, 3, G::stage_bytes(6), 3, 1, ctx->pre_slots, &K3, &sp3), ok2 = fit(cap, 2, G::stage_bytes(6), 3, 1, ctx->pre_slots, &K2, &sp2);
		if (ok3 && (!sp3 || !ok2 || sp2)) pl->nstg_pre = 3, pl->Kpre = K3, pl->spill_pre = sp3;
		else if (ok2) pl->nstg_pre = 2, pl->Kpre = K2, pl->spill_pre = sp2;
		else continue;
		pl->smem_post = G::HDR + (size_t)pl->nstg_post * G::stage_bytes(3) + (size_t)(pl->Kpost + 1) * slice;
		pl->smem_pre = G::HDR + (size_t)pl->nstg_pre * G::stage_bytes(6) + (size_t)(pl->Kpre + 3 + pl->spill_pre) * slice;
		return true;
	}
	return false;
}

// geometries: 8 consumer warps x 16 patterns (two m-tiles per warp: every B fragment feeds two DMMAs), or -- PHB_OPT_TUNE 16 -- 12 consumer
// warps x 8 patterns (three warps per scheduler overlap better, at twice the shared-memory reads per DMMA)
using DwGeomA = DwGeom<DW_S, 2, 8>;
using DwGeomB = DwGeom<DW_S, 1, 12>;
static bool dw_plan(const phbc_ctx *ctx, DwPlan *pl) {
	return ctx->tune == 16 ? dw_plan_g<DwGeomB>(ctx, pl) : dw_plan_g<DwGeomA>(ctx, pl);
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
template <class G, int MT>
static int dw_passes_g(phbc_ctx *ctx, const phbc_eval_opts *o, double *result, int phases) {
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
	if ((rc = phbc_dmma_pack_images(ctx, true, o->include_root_freqs, 2))) return rc;
	const size_t per_walk = (size_t)pl.ntiles * ctx->T * pl.TP;
	DwParams p;
	memset(&p, 0, sizeof(p));
	p.img = ctx->d_dmma_img, p.lower = ctx->d_lower;
	p.T = ctx->T, p.N = ctx->N, p.C = ctx->C, p.P = ctx->P, p.ntiles = pl.ntiles, p.nitems = pl.nitems, p.root = ctx->root;
	p.nops = ctx->dw_nops, p.nw = pl.nw;
	p.turns = ctx->tune == 15;  // PHB_OPT_TUNE 15: warp pairs take turns on the tensor pipe (A/B runs; measured slower than free-running warps)
	p.freqs = ctx->d_freqs, p.weights = ctx->d_weights, p.pattern_lnl = ctx->d_pattern_lnl, p.include_root_freqs = o->include_root_freqs;
	const int threads = 32 * (pl.nw + 1);
	auto post = k_dwalk_post<DW_S, MT, G::NWMAX>;
	auto pre = k_dwalk_pre<DW_S, MT, G::NWMAX>;
	if (phases & 1) {  // forward: post-order walk, root integration
		PHBC_CHECK(cudaFuncSetAttribute(post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_post));
		p.codes = ctx->d_dw_codes, p.ops = ctx->d_dw_post;
		p.K = pl.Kpost, p.nstg = pl.nstg_post, p.spill = pl.spill_post;
		const long long gpost = (long long)pl.ctas_post * ctx->num_sms;
		post<<<(int)(pl.nitems < gpost ? pl.nitems : gpost), threads, pl.smem_post, ctx->stream>>>(p);
		ctx->launches++;
		PHBC_CHECK(cudaGetLastError());
		if ((rc = phbc_generic_root(ctx, o, result))) return rc;
	}
	if (!(phases & 2) || !o->want_gradient) return 0;
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
	return phbc_gradient_from_partials(ctx, pstride, result);
}
int phbc_dwalk_passes(phbc_ctx *ctx, const phbc_eval_opts *o, double *result, int phases) {
	return ctx->tune == 16 ? dw_passes_g<DwGeomB, 1>(ctx, o, result, phases) : dw_passes_g<DwGeomA, 2>(ctx, o, result, phases);
}
