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
#include "phb_dmma_common.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

// acc[m][j] (+)= A-fragments[m][tt] x bf for all MT m-tiles of one (k-step, n-tile).
// (m16n8k4 on pairs of m-tiles was measured on B200: same results, 5-12 % SLOWER end to end than two m8n8k4 -- round 1, s16.)
template <int MT, int NTW, int KCH>
__device__ __forceinline__ void dmma_mtiles(double (&acc)[MT][NTW][2], const double (&a)[MT][KCH], int j, int tt, double bf) {
#pragma unroll
	for (int m = 0; m < MT; m++) dmma_m8n8k4(acc[m][j][0], acc[m][j][1], a[m][tt], bf);
}

__device__ __forceinline__ const double *dm_partial_ptr(const Bufs &b, int idx, int c) {
	const size_t PS = (size_t)b.P * b.S;
	if (idx < b.T) return b.tip_partials + (size_t)idx * PS;
	if (idx < b.N) return b.lower + ((size_t)(idx - b.T) * b.C + c) * PS;
	return b.upper + ((size_t)(idx - b.N) * b.C + c) * PS;
}
__device__ __forceinline__ bool dm_state_tip(const Bufs &b, int idx) { return idx < b.T && b.tip_kind == PHBC_TIP_STATES; }

// Packed shared-memory images of the transition matrices, written once per evaluation so that a CTA stages a matrix with ONE
// TMA bulk copy instead of an element-wise gather: zero padded [NP][LD] (B-fragment layout) for nodes whose lower partials are
// an operand, or TRANSPOSED [S][NP] followed by the row sums [NP] for state tips, so that one tip state selects a contiguous
// column of M (16-byte gathers) and an unknown state selects the row sums (treelikelihoodX.c:878-1001).
// grid (N, C, 2: P | dP), images laid out [which][node][category][IMG].
// adjoint (message form): the dP image of an INTERNAL node is stored transposed and weighted, [j][i] = f_i dP[i][j], the B operand of
// Z[j] = sum_i U[i] f_i dP[i][j] -- the node's own branch gradient is then sum_j L[j] Z[j] with the reference's dP entries themselves.
template <class Sh>
__global__ void k_dmma_pack(int T, int N, int C, int root, int tip_states, const double *__restrict__ Pm, const double *__restrict__ dPm,
                            double *__restrict__ img, int adjoint, const double *__restrict__ freqs, int include_root_freqs) {
	const int n = blockIdx.x, c = blockIdx.y, which = blockIdx.z;
	const double *src = (which ? dPm : Pm) + ((size_t)n * C + c) * Sh::S * Sh::S;
	double *dst = img + (((size_t)which * N + n) * C + c) * Sh::IMG;
	if (n == root) {  // the root has no branch: its "matrix" is the identity (message form: L_root passes through unchanged)
		for (int e = threadIdx.x; e < Sh::IMG; e += blockDim.x) {
			const int i = e / Sh::LD, j = e - i * Sh::LD;
			dst[e] = (e < Sh::MAT && i < Sh::S && i == j && which == 0) ? 1.0 : 0.0;
		}
		return;
	}
	if (n < T && tip_states) {
		// tip_states == 2 (whole-tree walk): derivative images carry the frequency weight of the gradient sum, [s][i] = f_i dP[i][s]
		const bool weighted = which == 1 && tip_states == 2 && !include_root_freqs;
		for (int e = threadIdx.x; e < Sh::S * Sh::NP; e += blockDim.x) {
			const int s = e / Sh::NP, i = e - s * Sh::NP;
			dst[e] = i < Sh::S ? (weighted ? freqs[i] : 1.0) * src[i * Sh::S + s] : 0.0;
		}
		for (int i = threadIdx.x; i < Sh::NP; i += blockDim.x) {
			// row S = what an unknown state selects: 1 for probabilities (treelikelihood20.c:125-131), the real row sum for derivatives
			double acc = 0.0;
			if (i < Sh::S) {
				if (which == 0) acc = 1.0;
				else
					for (int j = 0; j < Sh::S; j++) acc += src[i * Sh::S + j];
				if (weighted) acc *= freqs[i];
			}
			dst[Sh::S * Sh::NP + i] = acc;
		}
		for (int e = Sh::TIP_IMG + threadIdx.x; e < Sh::IMG; e += blockDim.x) dst[e] = 0.0;
	} else if (which == 1 && adjoint && n >= T) {
		for (int e = threadIdx.x; e < Sh::IMG; e += blockDim.x) {
			const int j = e / Sh::LD, i = e - j * Sh::LD;
			dst[e] = (e < Sh::MAT && i < Sh::S && j < Sh::S) ? (include_root_freqs ? 1.0 : freqs[i]) * src[i * Sh::S + j] : 0.0;
		}
	} else {
		for (int e = threadIdx.x; e < Sh::IMG; e += blockDim.x) {
			const int i = e / Sh::LD, j = e - i * Sh::LD;
			dst[e] = (e < Sh::MAT && i < Sh::S && j < Sh::S) ? src[i * Sh::S + j] : 0.0;
		}
	}
}

// A operands (patterns x states slices of partials buffers) travel HBM -> shared memory by cp.async in k-chunks, a ring of
// NSTAGE chunk buffers per m-group (the NSPLIT warps that share the same patterns), two chunks ahead of the tensor pipe:
// no register is tied up by data in flight, so the prefetch distance does not depend on instruction scheduling.
// One chunk buffer = NOPS operands x (MT * 8 patterns) rows x KC doubles, row stride RS = 4 (mod 8) doubles so that the
// fragment reads (lane (r, q) -> row r, column 4 tt + q) are bank-conflict free.
template <class Sh, int MT, int NOPS, int NST = 0>
struct AStage {
	static constexpr int KC = Sh::KCH * 4;
	static constexpr int RS = KC + ((4 - KC % 8) + 8) % 8;
	static constexpr int OPB = MT * 8 * RS;
	static constexpr int STG = NOPS * OPB;
	// ring depth: the fills run NSTAGE - 1 chunks ahead of the tensor pipe.  A contraction that is ONE chunk (20 states) stages whole tiles:
	// one tile ahead is enough there and the smaller ring buys a fourth resident CTA per SM (round 1, h4 profile: 12 -> 16 warps per SM)
	static constexpr int NSTAGE = NST > 0 ? NST : (Sh::NCH == 1 ? 2 : 3);
};

__device__ __forceinline__ void cp_async8(double *dst, const double *src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(double *dst, const double *src) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
	asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Per-thread plan of the cp.async granules of one operand chunk: thread gl of the m-group copies elements e = j * GT + gl of the
// [MT * 8 rows][KC doubles] chunk.  Row / column / shared-memory offsets are computed ONCE per kernel; a fill is then one min, one
// multiply-add and one cp.async per granule.  No predicates: rows past the pattern count read the last pattern again (their
// results are never stored and carry weight 0), columns past S read the first doubles of the NEXT row / block -- finite values
// that meet zero-padded matrix columns (the partials buffers are zero-initialised with a zeroed tail for exactly this reason).
// GB = doubles per granule: 2 (16-byte cp.async, half the LSU instructions and address arithmetic) wherever the HBM row stride S * 8, the
// chunk width and the staged row stride are multiples of 16 bytes (20 states: rows of 160 bytes); 1 otherwise (61 states: 488 bytes)
template <class Sh, int MT, int NOPS, int GT, int NST = 0, int GB = 1>
struct AFill {
	using A = AStage<Sh, MT, NOPS, NST>;
	static_assert(GB == 1 || (GB == 2 && Sh::S % 2 == 0 && A::KC % 2 == 0 && A::RS % 2 == 0), "16-byte granules need 16-byte aligned rows");
	static constexpr int KG = A::KC / GB;  // granules per row
	static constexpr int PER = MT * 8 * KG;
	static constexpr int NE = (PER + GT - 1) / GT;
	int row[NE], col[NE], soff[NE];
	__device__ __forceinline__ void init(int gl) {
#pragma unroll
		for (int j = 0; j < NE; j++) {
			const int e = j * GT + gl;
			row[j] = e / KG;
			col[j] = (e - row[j] * KG) * GB;
			soff[j] = row[j] * A::RS + col[j];
			if (PER % GT != 0 && e >= PER) row[j] = -1;  // this thread has no granule j
		}
	}
	__device__ __forceinline__ void fill(double *opbuf, const double *__restrict__ X, int p0, int P, int ch) const {
#pragma unroll
		for (int j = 0; j < NE; j++) {
			if (PER % GT != 0 && row[j] < 0) continue;
			const int p = min(p0 + row[j], P - 1);
			if (GB == 2) cp_async16(opbuf + soff[j], X + (size_t)p * Sh::S + (ch * A::KC + col[j]));
			else cp_async8(opbuf + soff[j], X + (size_t)p * Sh::S + (ch * A::KC + col[j]));
		}
	}
};

// fragments of one staged operand: a[m][tt] = chunk[8 m + lane / 4][4 tt + lane % 4]
template <class Sh, int MT, int NOPS, int NST = 0>
__device__ __forceinline__ void read_frags(const double *opbuf, int lane, double (&a)[MT][Sh::KCH]) {
	using A = AStage<Sh, MT, NOPS, NST>;
	const int r = lane >> 2, q = lane & 3;
#pragma unroll
	for (int m = 0; m < MT; m++)
#pragma unroll
		for (int tt = 0; tt < Sh::KCH; tt++) a[m][tt] = opbuf[(8 * m + r) * A::RS + 4 * tt + q];
}

// CTA-uniform op properties (child is a state tip, parent is the root) become compile-time constants of the tile loop:
// a predicated-off DMMA still occupies the FP64 tensor pipe, so the products an op does not need must not be in its code path.
template <class F>
__device__ __forceinline__ void dispatch2(bool x, bool y, F &&f) {
	using T = std::true_type;
	using N = std::false_type;
	if (x) {
		if (y) f(T{}, T{});
		else f(T{}, N{});
	} else {
		if (y) f(N{}, T{});
		else f(N{}, N{});
	}
}
template <class F>
__device__ __forceinline__ void dispatch3(bool x, bool y, bool z, F &&f) {
	using T = std::true_type;
	using N = std::false_type;
	if (x) dispatch2(y, z, [&](auto Y, auto Z) { f(T{}, Y, Z); });
	else dispatch2(y, z, [&](auto Y, auto Z) { f(N{}, Y, Z); });
}

template <int NSPLIT>
__device__ __forceinline__ void group_sync(int wm) {
	if (NSPLIT == 1) __syncwarp();
	else asm volatile("bar.sync %0, %1;" ::"r"(1 + wm), "n"(32 * NSPLIT) : "memory");
}

template <int MT, int NTW>
__device__ __forceinline__ void zero_acc(double (&acc)[MT][NTW][2]) {
#pragma unroll
	for (int m = 0; m < MT; m++)
#pragma unroll
		for (int j = 0; j < NTW; j++) acc[m][j][0] = acc[m][j][1] = 0.0;
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

template <class Sh, int MT, int NTW>
__device__ __forceinline__ void store_tile(double *__restrict__ U, int p0, int P, int n0, int lane, const double (&v)[MT][NTW][2]) {
	const int r = lane >> 2, q = lane & 3;
#pragma unroll
	for (int m = 0; m < MT; m++) {
		const int p = p0 + 8 * m + r;
		if (p >= P) continue;
		double *row = U + (size_t)p * Sh::S;
#pragma unroll
		for (int j = 0; j < NTW; j++) {
			const int i = (n0 + j) * 8 + 2 * q;
			if (Sh::S % 2 == 0) {
				if (i < Sh::S) *reinterpret_cast<double2 *>(row + i) = make_double2(v[m][j][0], v[m][j][1]);  // S even: 16-byte aligned
			} else {
				if (i < Sh::S) row[i] = v[m][j][0];
				if (i + 1 < Sh::S) row[i + 1] = v[m][j][1];
			}
		}
	}
}

// Rescaling without re-reading the partials: the kernel that produces a partial also leaves, per (op slot, category, n-split, pattern),
// the largest entry of what it stored (padding columns hold exact zeros).  k_dmma_scale_from_max then decides per pattern from C x
// NSPLIT numbers instead of C x S, and touches the partial only where a pattern really is rescaled (SingleTreeLikelihood_scalePartials,
// treelikelihood.c:1790-1836).  The one-thread-per-pattern K5 kernel cost 19 of the 33.8 ms of a rescaled LG+G4 400 x 50k evaluation.
template <class Sh, int MT, int NTW>
__device__ __forceinline__ void store_rowmax(double *__restrict__ rowmax, size_t row0 /* ((slot * C + c) * NSPLIT + split) * P */, int p0, int P,
                                             int n0, int lane, const double (&v)[MT][NTW][2], bool stored) {
	const int r = lane >> 2, q = lane & 3;
#pragma unroll
	for (int m = 0; m < MT; m++) {
		double mx = 0.0;
#pragma unroll
		for (int j = 0; j < NTW; j++) {
			const int i = (n0 + j) * 8 + 2 * q;  // padding columns are not part of the partial (a gap tip's message is 1 there)
			if (i < Sh::S) mx = fmax(mx, v[m][j][0]);
			if (i + 1 < Sh::S) mx = fmax(mx, v[m][j][1]);
		}
		mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
		mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
		const int p = p0 + 8 * m + r;
		if (q == 0 && p < P) rowmax[row0 + p] = stored ? mx : 0.0;  // an upper partial that is not stored: nothing to rescale
	}
}

// one thread per (pattern, category); slot_of == NULL: the op's slot is its index in the launch (lower passes), else slot_of[child]
// (upper passes: 2 x the parent op's index in its level + which child)
__global__ void k_dmma_scale_from_max(Bufs b, const phbc_op *__restrict__ ops, const double *__restrict__ rowmax, const int *__restrict__ slot_of,
                                      int slot0, int nslots, int nsplit, double threshold) {
	const phbc_op op = ops[blockIdx.y];
	__shared__ double m_s[128];
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int C = b.C;
	const int p = (int)(t / C), c = (int)(t - (size_t)p * C);
	const bool live = p < b.P;
	const int slot = slot_of ? slot_of[op.out - b.N] - slot0 : (int)blockIdx.y;
	if (slot < 0 || slot >= nslots) return;  // uniform over the block: this op's producer ran in another chunk of the level
	double m = 0.0;
	if (live)
		for (int k = 0; k < nsplit; k++) m = fmax(m, rowmax[(((size_t)slot * C + c) * nsplit + k) * b.P + p]);
	for (int off = C >> 1; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
	const bool rescaled = live && m < threshold && m > 0.0;
	if (c == 0) m_s[threadIdx.x / C] = rescaled ? m : 0.0;
	if (__syncthreads_or(rescaled)) phbc_rescale_block(b, op.out, (int)((size_t)blockIdx.x * blockDim.x / C), blockDim.x / C, m_s);
	if (live && c == 0) {
		double sf = rescaled ? log(m) : 0.0;
		if (!is_state_tip(b, op.a)) sf += b.sf[(size_t)op.a * b.P + p];
		if (op.b >= 0 && !is_state_tip(b, op.b)) sf += b.sf[(size_t)op.b * b.P + p];
		b.sf[(size_t)op.out * b.P + p] = sf;
	}
}

// ---------------------------------------------------------------------------------------------
// K1-K4: out = (P_a x_a) o (P_b x_b);  grid (pattern chunks, C, ops of the level)
// warps: WM m-groups x NSPLIT n-groups; a warp owns MT m-tiles and NT / NSPLIT n-tiles.
// Both products advance through ONE k-loop (2 MT NTW independent accumulator chains per warp keep the FP64 tensor pipe
// busy across its latency) and the A fragments of the next chunk / next tile are in flight while it runs.
// ---------------------------------------------------------------------------------------------
template <int S, int MT, int NSPLIT, int WM>
__global__ void __launch_bounds__(32 * WM * NSPLIT) k_dmma_lower(Bufs b, const phbc_op *__restrict__ ops, const double *__restrict__ img,
                                                               double *__restrict__ rowmax /* NULL: not rescaling */) {
	using Sh = DmmaShape<S>;
	constexpr int NTW = Sh::NT / NSPLIT;
	static_assert(Sh::NT % NSPLIT == 0, "n-tiles must split evenly");
	extern __shared__ __align__(128) unsigned char smraw[];
	uint64_t *bar = reinterpret_cast<uint64_t *>(smraw);
	double *sm = reinterpret_cast<double *>(smraw + 128);
	double *mA = sm, *mB = sm + Sh::IMG;
	const phbc_op op = ops[blockIdx.z];
	const int c = blockIdx.y;
	const bool a_tip_rt = dm_state_tip(b, op.a), b_tip_rt = dm_state_tip(b, op.b);
	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		mbar_expect_tx(bar, 2 * Sh::IMG * 8);
		bulk_g2s(mA, img + ((size_t)op.a_mat * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
		bulk_g2s(mB, img + ((size_t)op.b_mat * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
	}
	__syncthreads();  // the barrier is initialised before anyone polls it
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int wm = warp / NSPLIT, n0 = (warp % NSPLIT) * NTW;
	const int r = lane >> 2, q = lane & 3;
	constexpr int TP = WM * MT * 8;
	double *out = (double *)dm_partial_ptr(b, op.out, c);
	// state tips have no partial operand: point the (fully predicated-off) loads at a valid address
	const double *xa = a_tip_rt ? nullptr : dm_partial_ptr(b, op.a, c), *xb = b_tip_rt ? nullptr : dm_partial_ptr(b, op.b, c);
	const int Pa = a_tip_rt ? 0 : b.P, Pb = b_tip_rt ? 0 : b.P;  // P = 0 switches the operand's loads off
	const int ntiles = (b.P + TP - 1) / TP;
	using AS = AStage<Sh, MT, 2>;
	constexpr int GT = 32 * NSPLIT;
	double *abuf = sm + 2 * Sh::IMG + wm * AS::NSTAGE * AS::STG;  // this m-group's ring
	const int gl = (warp % NSPLIT) * 32 + lane;
	mbar_wait(bar, 0);  // matrices have landed (the first A chunks are already in flight)
	dispatch2(a_tip_rt, b_tip_rt, [&](auto ATIP, auto BTIP) {
	constexpr bool a_tip = decltype(ATIP)::value, b_tip = decltype(BTIP)::value;
	AFill<Sh, MT, 2, GT> plan;
	plan.init(gl);
	int f_tile = blockIdx.x, f_ch = 0, f_stage = 0;
	auto fill_next = [&]() {  // stage the next chunk of the (tile, chunk) sequence; always commits, so group counts stay uniform
		if (f_tile < ntiles) {
			double *stg = abuf + f_stage * AS::STG;
			const int fp0 = f_tile * TP + wm * MT * 8;
			if (!a_tip) plan.fill(stg, xa, fp0, b.P, f_ch);
			if (!b_tip) plan.fill(stg + AS::OPB, xb, fp0, b.P, f_ch);
		}
		cp_async_commit();
		f_stage = f_stage + 1 == AS::NSTAGE ? 0 : f_stage + 1;
		if (++f_ch == Sh::NCH) f_ch = 0, f_tile += gridDim.x;
	};
#pragma unroll
	for (int pf = 0; pf < AS::NSTAGE - 1; pf++) fill_next();
	int c_stage = 0;
	// B fragments run one (k-step, n-tile) ahead of the tensor pipe: the LDS of step u + 1 is issued before the DMMAs of step u
	double bA = 0.0, bB = 0.0;
	{
		const int off = (n0 * 8 + r) * Sh::LD + q;
		if (!a_tip) bA = mA[off];
		if (!b_tip) bB = mB[off];
	}
	for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const int p0 = tile * TP + wm * MT * 8;
		double accA[MT][NTW][2], accB[MT][NTW][2];
		zero_acc<MT, NTW>(accA);
		zero_acc<MT, NTW>(accB);
		if (!(a_tip && b_tip))  // a cherry is two gathers: no contraction, no staged operand
#pragma unroll
		for (int ch = 0; ch < Sh::NCH; ch++) {
			cp_async_wait<AS::NSTAGE - 2>();      // this chunk has landed (the newest group may still be in flight)
			group_sync<NSPLIT>(wm);  // ... for every thread of the group, and the buffer refilled below is no longer read
			fill_next();
			const double *st = abuf + c_stage * AS::STG;
			c_stage = c_stage + 1 == AS::NSTAGE ? 0 : c_stage + 1;
			double ca[MT][Sh::KCH], cb[MT][Sh::KCH];
			if (!a_tip) read_frags<Sh, MT, 2>(st, lane, ca);
			if (!b_tip) read_frags<Sh, MT, 2>(st + AS::OPB, lane, cb);
#pragma unroll
			for (int tt = 0; tt < Sh::KCH; tt++) {
				const int t = ch * Sh::KCH + tt;
				if (t < Sh::KT) {
#pragma unroll
					for (int j = 0; j < NTW; j++) {
						const int nj = j + 1 < NTW ? j + 1 : 0, nt = j + 1 < NTW ? t : (t + 1 < Sh::KT ? t + 1 : 0);  // wraps into the next tile
						const int noff = ((n0 + nj) * 8 + r) * Sh::LD + 4 * nt + q;
						double nA = 0.0, nB = 0.0;
						if (!a_tip) nA = mA[noff];
						if (!b_tip) nB = mB[noff];
						if (!a_tip) {
dmma_mtiles<MT, NTW, Sh::KCH>(accA, ca, j, tt, bA);
						}
						if (!b_tip) {
dmma_mtiles<MT, NTW, Sh::KCH>(accB, cb, j, tt, bB);
						}
						bA = nA, bB = nB;
					}
				}
			}
		}
		if (a_tip) tip_gather<Sh, MT, NTW, true>(b.tip_states + (size_t)op.a * b.P, mA, nullptr, p0, b.P, n0, lane, accA);
		if (b_tip) tip_gather<Sh, MT, NTW, true>(b.tip_states + (size_t)op.b * b.P, mB, nullptr, p0, b.P, n0, lane, accB);
#pragma unroll
		for (int m = 0; m < MT; m++)
#pragma unroll
			for (int j = 0; j < NTW; j++) accA[m][j][0] *= accB[m][j][0], accA[m][j][1] *= accB[m][j][1];
		store_tile<Sh, MT, NTW>(out, p0, b.P, n0, lane, accA);
		if (rowmax) store_rowmax<Sh, MT, NTW>(rowmax, (((size_t)blockIdx.z * b.C + c) * NSPLIT + warp % NSPLIT) * b.P, p0, b.P, n0, lane, accA, true);
	}
	});
}

// ---------------------------------------------------------------------------------------------
// K8-K10 fused per parent;  grid (pattern chunks, C, parent ops of the level)
// GRAD: reduce the branch-gradient terms (unscaled form, site likelihood from pattern_lnl);
// !GRAD: only the upper partials, tips included (the caller reduces with the generic K9/K10 kernel).
// partial: [N][C][gridDim.x] per-CTA sums of the children's gradient terms.
// The five products W = P_n U_n, M_b = P_b L_b, D_b = dP_b L_b, M_a = P_a L_a, D_a = dP_a L_a share one k-loop
// (5 MT NTW independent accumulator chains per warp) with the next chunk's A fragments in flight.
// ---------------------------------------------------------------------------------------------
template <int S, int MT, int NSPLIT, int WM, bool GRAD>
__global__ void __launch_bounds__(32 * WM * NSPLIT) k_dmma_upper(Bufs b, const phbc_parent_op *__restrict__ ops, const double *__restrict__ img,
                                                               const double *__restrict__ freqs, const double *__restrict__ weights,
                                                               const double *__restrict__ pattern_lnl, int include_root_freqs, int pstride,
                                                               double *__restrict__ partial, int scaled, double *__restrict__ rowmax) {
	using Sh = DmmaShape<S>;
	constexpr int NTW = Sh::NT / NSPLIT;
	constexpr int NWARPS = WM * NSPLIT;
	extern __shared__ __align__(128) unsigned char smraw[];
	uint64_t *bar = reinterpret_cast<uint64_t *>(smraw);
	double *sm = reinterpret_cast<double *>(smraw + 128);
	double *mP = sm, *mA = sm + Sh::IMG, *mB = sm + 2 * Sh::IMG;
	double *dA = GRAD ? sm + 3 * Sh::IMG : nullptr, *dB = GRAD ? sm + 4 * Sh::IMG : nullptr;
	double *aux = sm + (GRAD ? 5 : 3) * Sh::IMG;  // fq[NP], wroot[NP], red[2 * NWARPS]
	double *fq = aux, *wroot = aux + Sh::NP, *red = aux + 2 * Sh::NP;
	const phbc_parent_op op = ops[blockIdx.z];
	const int c = blockIdx.y;
	const bool is_root_rt = op.flags & 1;
	const bool a_tip_rt = dm_state_tip(b, op.a), b_tip_rt = dm_state_tip(b, op.b);
	const double *rsA = dA + Sh::S * Sh::NP, *rsB = dB + Sh::S * Sh::NP;  // row sums ride behind a tip's transposed image
	if (threadIdx.x == 0) {
		const size_t dimg = (size_t)b.N * b.C * Sh::IMG;  // offset of the dP images
		mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		mbar_expect_tx(bar, ((is_root_rt ? 0 : 1) + (GRAD ? 4 : 2)) * Sh::IMG * 8);
		if (!is_root_rt) bulk_g2s(mP, img + ((size_t)op.node * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
		bulk_g2s(mA, img + ((size_t)op.a * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
		bulk_g2s(mB, img + ((size_t)op.b * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
		if (GRAD) {
			bulk_g2s(dA, img + dimg + ((size_t)op.a * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
			bulk_g2s(dB, img + dimg + ((size_t)op.b * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
		}
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
	const double *xa = a_tip_rt ? nullptr : dm_partial_ptr(b, op.a, c), *xb = b_tip_rt ? nullptr : dm_partial_ptr(b, op.b, c);
	const int Pw = is_root_rt ? 0 : b.P, Pa = a_tip_rt ? 0 : b.P, Pb = b_tip_rt ? 0 : b.P;  // P = 0 switches an operand's loads off
	double *Ua = b.upper + ((size_t)op.a * b.C + c) * (size_t)b.P * S;
	double *Ub = b.upper + ((size_t)op.b * b.C + c) * (size_t)b.P * S;
	const int ntiles = (b.P + TP - 1) / TP;
	double tot_a = 0.0, tot_b = 0.0;
	using AS = AStage<Sh, MT, 3>;
	constexpr int GT = 32 * NSPLIT;
	double *abuf = red + 2 * NWARPS + wm * AS::NSTAGE * AS::STG;  // this m-group's ring: operands U_n | L_b | L_a
	const int gl = (warp % NSPLIT) * 32 + lane;
	mbar_wait(bar, 0);  // matrices have landed (the first A chunks are already in flight)
	dispatch3(is_root_rt, a_tip_rt, b_tip_rt, [&](auto ROOT, auto ATIP, auto BTIP) {
	constexpr bool is_root = decltype(ROOT)::value, a_tip = decltype(ATIP)::value, b_tip = decltype(BTIP)::value;
	AFill<Sh, MT, 3, GT> plan;
	plan.init(gl);
	int f_tile = blockIdx.x, f_ch = 0, f_stage = 0;
	auto fill_next = [&]() {  // stage the next chunk of the (tile, chunk) sequence; always commits, so group counts stay uniform
		if (f_tile < ntiles) {
			double *stg = abuf + f_stage * AS::STG;
			const int fp0 = f_tile * TP + wm * MT * 8;
			if (!is_root) plan.fill(stg, xw, fp0, b.P, f_ch);
			if (!b_tip) plan.fill(stg + AS::OPB, xb, fp0, b.P, f_ch);
			if (!a_tip) plan.fill(stg + 2 * AS::OPB, xa, fp0, b.P, f_ch);
		}
		cp_async_commit();
		f_stage = f_stage + 1 == AS::NSTAGE ? 0 : f_stage + 1;
		if (++f_ch == Sh::NCH) f_ch = 0, f_tile += gridDim.x;
	};
#pragma unroll
	for (int pf = 0; pf < AS::NSTAGE - 1; pf++) fill_next();
	int c_stage = 0;
	// B fragments run one (k-step, n-tile) ahead of the tensor pipe: the LDS of step u + 1 is issued before the DMMAs of step u
	double bP = 0.0, bA = 0.0, bB = 0.0, bdA = 0.0, bdB = 0.0;
	{
		const int off = (n0 * 8 + r) * Sh::LD + q;
		if (!is_root) bP = mP[off];
		if (!b_tip) {
			bB = mB[off];
			if (GRAD) bdB = dB[off];
		}
		if (!a_tip) {
			bA = mA[off];
			if (GRAD) bdA = dA[off];
		}
	}
	for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const int p0 = tile * TP + wm * MT * 8;
		// w_k and log L_k of this tile's patterns: requested now, consumed in the epilogue
		double wk[MT], lk[MT];
		if (GRAD) {
#pragma unroll
			for (int m = 0; m < MT; m++) {
				const int p = p0 + 8 * m + r;
				wk[m] = p < b.P ? __ldg(weights + p) : 0.0;
				lk[m] = p < b.P ? __ldg(pattern_lnl + p) : 0.0;
				// Rescaling: the partials this op combines carry their CUMULATIVE log scaling factors (k_generic_scale: sf[out] = log m +
				// sf[a] + sf[b]).  U_a = W o M_b is formed here from the parent's upper partial (scale sf[N + n]) and the sibling's lower
				// partial (sf[b]) and meets L_a (sf[a]) -- and the same three factors in the other order for branch b -- so both branch
				// terms of this op are the scaled sums times exp(sf[N + n] + sf[a] + sf[b] - lnL_k): the exact gradient
				// (gradient_cat_branch_lengths with dlikelihood / likelihood, treelikelihood.c:2715-2789, 2464-2474), one exp per pattern.
				if (scaled && p < b.P) {
					double e = is_root ? 0.0 : b.sf[(size_t)(b.N + op.node) * b.P + p];
					if (!a_tip) e += b.sf[(size_t)op.a * b.P + p];
					if (!b_tip) e += b.sf[(size_t)op.b * b.P + p];
					lk[m] -= e;
				}
			}
		}
		double W[MT][NTW][2], Mb[MT][NTW][2], Db[MT][NTW][2], Ma[MT][NTW][2], Da[MT][NTW][2];
		zero_acc<MT, NTW>(W);
		zero_acc<MT, NTW>(Mb);
		zero_acc<MT, NTW>(Ma);
		if (GRAD) {
			zero_acc<MT, NTW>(Db);
			zero_acc<MT, NTW>(Da);
		}
#pragma unroll
		for (int ch = 0; ch < Sh::NCH; ch++) {
			cp_async_wait<AS::NSTAGE - 2>();      // this chunk has landed (the newest group may still be in flight)
			group_sync<NSPLIT>(wm);  // ... for every thread of the group, and the buffer refilled below is no longer read
			fill_next();
			const double *st = abuf + c_stage * AS::STG;
			c_stage = c_stage + 1 == AS::NSTAGE ? 0 : c_stage + 1;
			double cw[MT][Sh::KCH], cb[MT][Sh::KCH], ca[MT][Sh::KCH];
			if (!is_root) read_frags<Sh, MT, 3>(st, lane, cw);
			if (!b_tip) read_frags<Sh, MT, 3>(st + AS::OPB, lane, cb);
			if (!a_tip) read_frags<Sh, MT, 3>(st + 2 * AS::OPB, lane, ca);
#pragma unroll
			for (int tt = 0; tt < Sh::KCH; tt++) {
				const int t = ch * Sh::KCH + tt;
				if (t < Sh::KT) {
#pragma unroll
					for (int j = 0; j < NTW; j++) {
						// B fragments of the NEXT (k-step, n-tile) first, then this step's DMMAs (wraps into the next tile)
						const int nj = j + 1 < NTW ? j + 1 : 0, nt = j + 1 < NTW ? t : (t + 1 < Sh::KT ? t + 1 : 0);
						const int noff = ((n0 + nj) * 8 + r) * Sh::LD + 4 * nt + q;
						double nP = 0.0, nA = 0.0, nB = 0.0, ndA = 0.0, ndB = 0.0;
						if (!is_root) nP = mP[noff];
						if (!b_tip) {
							nB = mB[noff];
							if (GRAD) ndB = dB[noff];
						}
						if (!a_tip) {
							nA = mA[noff];
							if (GRAD) ndA = dA[noff];
						}
						if (!is_root) {
dmma_mtiles<MT, NTW, Sh::KCH>(W, cw, j, tt, bP);
						}
						if (!b_tip) {
dmma_mtiles<MT, NTW, Sh::KCH>(Mb, cb, j, tt, bB);
							if (GRAD) {
dmma_mtiles<MT, NTW, Sh::KCH>(Db, cb, j, tt, bdB);
							}
						}
						if (!a_tip) {
dmma_mtiles<MT, NTW, Sh::KCH>(Ma, ca, j, tt, bA);
							if (GRAD) {
dmma_mtiles<MT, NTW, Sh::KCH>(Da, ca, j, tt, bdA);
							}
						}
						bP = nP, bA = nA, bB = nB, bdA = ndA, bdB = ndB;
					}
				}
			}
		}
		if (is_root) {
#pragma unroll
			for (int m = 0; m < MT; m++)
#pragma unroll
				for (int j = 0; j < NTW; j++) {
					const int i = (n0 + j) * 8 + 2 * q;
					W[m][j][0] = wroot[i], W[m][j][1] = wroot[i + 1];
				}
		}
		if (b_tip) {
			const uint8_t *st = b.tip_states + (size_t)op.b * b.P;
			tip_gather<Sh, MT, NTW, true>(st, mB, nullptr, p0, b.P, n0, lane, Mb);
			if (GRAD) tip_gather<Sh, MT, NTW, false>(st, dB, rsB, p0, b.P, n0, lane, Db);
		}
		if (a_tip) {
			const uint8_t *st = b.tip_states + (size_t)op.a * b.P;
			tip_gather<Sh, MT, NTW, true>(st, mA, nullptr, p0, b.P, n0, lane, Ma);
			if (GRAD) tip_gather<Sh, MT, NTW, false>(st, dA, rsA, p0, b.P, n0, lane, Da);
		}
		// U_a = W o M_b, U_b = W o M_a (in place), g_a = sum f U_a D_a, g_b = sum f U_b D_b
#pragma unroll
		for (int m = 0; m < MT; m++)
#pragma unroll
			for (int j = 0; j < NTW; j++) {
				const double ua0 = W[m][j][0] * Mb[m][j][0], ua1 = W[m][j][1] * Mb[m][j][1];
				const double ub0 = W[m][j][0] * Ma[m][j][0], ub1 = W[m][j][1] * Ma[m][j][1];
				Mb[m][j][0] = ua0, Mb[m][j][1] = ua1;
				Ma[m][j][0] = ub0, Ma[m][j][1] = ub1;
			}
		if (!a_leaf || !GRAD) store_tile<Sh, MT, NTW>(Ua, p0, b.P, n0, lane, Mb);
		if (!b_leaf || !GRAD) store_tile<Sh, MT, NTW>(Ub, p0, b.P, n0, lane, Ma);
		if (rowmax) {  // slots 2 z (child a) and 2 z + 1 (child b) of this launch
			const size_t row = (((size_t)(2 * blockIdx.z) * b.C + c) * NSPLIT + warp % NSPLIT) * b.P;
			store_rowmax<Sh, MT, NTW>(rowmax, row, p0, b.P, n0, lane, Mb, !a_leaf || !GRAD);
			store_rowmax<Sh, MT, NTW>(rowmax, row + (size_t)b.C * NSPLIT * b.P, p0, b.P, n0, lane, Ma, !b_leaf || !GRAD);
		}
		if (GRAD) {
#pragma unroll
			for (int m = 0; m < MT; m++) {
				double ga = 0.0, gb = 0.0;
#pragma unroll
				for (int j = 0; j < NTW; j++) {
					const int i = (n0 + j) * 8 + 2 * q;
					ga = fma(fq[i] * Mb[m][j][0], Da[m][j][0], fma(fq[i + 1] * Mb[m][j][1], Da[m][j][1], ga));
					gb = fma(fq[i] * Ma[m][j][0], Db[m][j][0], fma(fq[i + 1] * Ma[m][j][1], Db[m][j][1], gb));
				}
				// the four lanes of a quad hold one pattern's columns
				ga += __shfl_xor_sync(0xffffffffu, ga, 1), gb += __shfl_xor_sync(0xffffffffu, gb, 1);
				ga += __shfl_xor_sync(0xffffffffu, ga, 2), gb += __shfl_xor_sync(0xffffffffu, gb, 2);
				const double wl = wk[m] / exp(lk[m]);  // w_k / L_k (treelikelihood.c:3207-3210); padding patterns carry w = 0
				tot_a = fma(ga, wl, tot_a);
				tot_b = fma(gb, wl, tot_b);
			}
		}
	}
	});
	if (GRAD) {
		// every quad lane carries the same row value: sum lanes with q == 0 over the 8 rows, then across warps (fixed order)
		tot_a = q == 0 ? tot_a : 0.0, tot_b = q == 0 ? tot_b : 0.0;
		tot_a = phb_warp_sum(tot_a), tot_b = phb_warp_sum(tot_b);
		if (lane == 0) red[2 * warp] = tot_a, red[2 * warp + 1] = tot_b;
		__syncthreads();
		if (threadIdx.x == 0) {
			double sa = 0.0, sb = 0.0;
			for (int w = 0; w < NWARPS; w++) sa += red[2 * w], sb += red[2 * w + 1];
			partial[((size_t)op.a * b.C + c) * pstride + blockIdx.x] = sa;
			partial[((size_t)op.b * b.C + c) * pstride + blockIdx.x] = sb;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Message form (unscaled evaluations with state tips -- the fast path).  What a node hands to its parent is stored instead of its
// lower partial: M_n = P_n L_n with L_n = M_a o M_b.  The product P_n L_n is needed by the parent's lower partial and by the sibling's
// upper partial, and the node-at-a-time formulation above computes it twice (lower pass and upper pass); here it is computed once.
// The branch gradient of n is reduced at n's OWN pre-order op from L_n = M_a o M_b and U_n (adjoint form, see k_dmma_upper_msg), so an
// evaluation runs 3 dense products per internal node (P_n L_n, P_n U_n, U_n (f o dP_n)) instead of 4.
// HBM traffic is unchanged: M_n takes the place of L_n in the lower buffers.
//
// k_dmma_lower_msg: A fragments are formed on the fly as products of the children's messages (staged rows for internal children,
// columns of the transposed matrix image for state tips), B is the node's OWN matrix (the identity at the root, whose L then feeds
// the root integration unchanged).
// ---------------------------------------------------------------------------------------------
#define PHBC_OP_TABLE 0x100  // phbc_op.flags of a cherry-table op: the row block of the table is flags >> 9
template <class Sh>
__device__ __forceinline__ double tip_value(const double *__restrict__ MT, int s, int col) {
	return s < Sh::S ? MT[s * Sh::NP + col] : 1.0;  // unknown state: factor 1 (treelikelihood20.c:125-131)
}

template <int S, int MT, int NSPLIT, int WM, int NST = 0, int GB = 1, bool RESCALE = false>
__global__ void __launch_bounds__(32 * WM * NSPLIT) k_dmma_lower_msg(Bufs b, const phbc_op *__restrict__ ops, const double *__restrict__ img, int nimg,
                                                                   double *__restrict__ rowmax /* rescaling: [op of the launch][C][P] max of L_n, else NULL */) {
	using Sh = DmmaShape<S>;
	constexpr int NTW = Sh::NT / NSPLIT;
	static_assert(Sh::NT % NSPLIT == 0, "n-tiles must split evenly");
	extern __shared__ __align__(128) unsigned char smraw[];
	uint64_t *bar = reinterpret_cast<uint64_t *>(smraw);
	double *sm = reinterpret_cast<double *>(smraw + 128);
	const phbc_op op = ops[blockIdx.z];
	const int c = blockIdx.y;
	const bool a_tip_rt = op.a < b.T, b_tip_rt = op.b < b.T;
	// image slots: own matrix, then one transposed image per tip child; the launch provides nimg >= 1 + tip children slots
	double *mN = sm, *mA = sm + Sh::IMG, *mB = sm + (a_tip_rt ? 2 : 1) * Sh::IMG;
	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		mbar_expect_tx(bar, (1 + (a_tip_rt ? 1 : 0) + (b_tip_rt ? 1 : 0)) * Sh::IMG * 8);
		bulk_g2s(mN, img + ((size_t)op.out * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
		if (a_tip_rt) bulk_g2s(mA, img + ((size_t)op.a_mat * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
		if (b_tip_rt) bulk_g2s(mB, img + ((size_t)op.b_mat * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
	}
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int wm = warp / NSPLIT, n0 = (warp % NSPLIT) * NTW;
	const int r = lane >> 2, q = lane & 3;
	constexpr int TP = WM * MT * 8;
	// a cherry-table op (see k_dmma_cherry_gather) writes row block flags >> 9 of the table b.lower points at
	double *out = (op.flags & PHBC_OP_TABLE) ? b.lower + ((size_t)(op.flags >> 9) * b.C + c) * (size_t)b.P * S : (double *)dm_partial_ptr(b, op.out, c);
	const double *xa = a_tip_rt ? nullptr : dm_partial_ptr(b, op.a, c), *xb = b_tip_rt ? nullptr : dm_partial_ptr(b, op.b, c);
	const int ntiles = (b.P + TP - 1) / TP;
	using AS = AStage<Sh, MT, 2, NST>;
	constexpr int GT = 32 * NSPLIT;
	double *abuf = sm + nimg * Sh::IMG + wm * AS::NSTAGE * AS::STG;
	const int gl = (warp % NSPLIT) * 32 + lane;
	mbar_wait(bar, 0);
	dispatch2(a_tip_rt, b_tip_rt, [&](auto ATIP, auto BTIP) {
	constexpr bool a_tip = decltype(ATIP)::value, b_tip = decltype(BTIP)::value;
	AFill<Sh, MT, 2, GT, NST, GB> plan;
	plan.init(gl);
	int f_tile = blockIdx.x, f_ch = 0, f_stage = 0;
	auto fill_next = [&]() {
		if (f_tile < ntiles) {
			double *stg = abuf + f_stage * AS::STG;
			const int fp0 = f_tile * TP + wm * MT * 8;
			if (!a_tip) plan.fill(stg, xa, fp0, b.P, f_ch);
			if (!b_tip) plan.fill(stg + AS::OPB, xb, fp0, b.P, f_ch);
		}
		cp_async_commit();
		f_stage = f_stage + 1 == AS::NSTAGE ? 0 : f_stage + 1;
		if (++f_ch == Sh::NCH) f_ch = 0, f_tile += gridDim.x;
	};
#pragma unroll
	for (int pf = 0; pf < AS::NSTAGE - 1; pf++) fill_next();
	int c_stage = 0;
	double bN = mN[(n0 * 8 + r) * Sh::LD + q];
	const uint8_t *sta = b.tip_states + (size_t)(a_tip ? op.a : 0) * b.P, *stb = b.tip_states + (size_t)(b_tip ? op.b : 0) * b.P;
	for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const int p0 = tile * TP + wm * MT * 8;
		int sa[MT], sb[MT];
#pragma unroll
		for (int m = 0; m < MT; m++) {
			const int p = p0 + 8 * m + r;
			sa[m] = (a_tip && p < b.P) ? sta[p] : Sh::S;
			sb[m] = (b_tip && p < b.P) ? stb[p] : Sh::S;
		}
		double acc[MT][NTW][2];
		zero_acc<MT, NTW>(acc);
		double lmax[MT];  // rescaling: the largest entry of L_n per row, what SingleTreeLikelihood_scalePartials compares with the threshold
#pragma unroll
		for (int m = 0; m < MT; m++) lmax[m] = 0.0;
#pragma unroll
		for (int ch = 0; ch < Sh::NCH; ch++) {
			cp_async_wait<AS::NSTAGE - 2>();
			group_sync<NSPLIT>(wm);
			fill_next();
			const double *st = abuf + c_stage * AS::STG;
			c_stage = c_stage + 1 == AS::NSTAGE ? 0 : c_stage + 1;
			double ca[MT][Sh::KCH], cb[MT][Sh::KCH];
			if (!a_tip) read_frags<Sh, MT, 2, NST>(st, lane, ca);
			if (!b_tip) read_frags<Sh, MT, 2, NST>(st + AS::OPB, lane, cb);
#pragma unroll
			for (int m = 0; m < MT; m++)
#pragma unroll
				for (int tt = 0; tt < Sh::KCH; tt++) {
					const int col = ch * AS::KC + 4 * tt + q;
					if (ch * Sh::KCH + tt < Sh::KT) {
						if (a_tip) ca[m][tt] = tip_value<Sh>(mA, sa[m], col);
						if (b_tip) cb[m][tt] = tip_value<Sh>(mB, sb[m], col);
						ca[m][tt] *= cb[m][tt];  // L_n = M_a o M_b, straight into the A fragment
						if (RESCALE && col < S) lmax[m] = fmax(lmax[m], ca[m][tt]);  // compiled out of the unscaled kernel: DMNMX shares the FP64 pipe with DMMA
					} else ca[m][tt] = 0.0;
				}
#pragma unroll
			for (int tt = 0; tt < Sh::KCH; tt++) {
				const int t = ch * Sh::KCH + tt;
				if (t < Sh::KT) {
#pragma unroll
					for (int j = 0; j < NTW; j++) {
						const int nj = j + 1 < NTW ? j + 1 : 0, nt = j + 1 < NTW ? t : (t + 1 < Sh::KT ? t + 1 : 0);
						const double nN = mN[((n0 + nj) * 8 + r) * Sh::LD + 4 * nt + q];
						dmma_mtiles<MT, NTW, Sh::KCH>(acc, ca, j, tt, bN);
						bN = nN;
					}
				}
			}
		}
		store_tile<Sh, MT, NTW>(out, p0, b.P, n0, lane, acc);
		if (RESCALE && warp % NSPLIT == 0) {  // every n-split warp of the group formed the same L_n
#pragma unroll
			for (int m = 0; m < MT; m++) {
				double mx = fmax(lmax[m], __shfl_xor_sync(0xffffffffu, lmax[m], 1));
				mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
				const int p = p0 + 8 * m + r;
				if (q == 0 && p < b.P) rowmax[((size_t)blockIdx.z * b.C + c) * b.P + p] = mx;
			}
		}
	}
	});
}

// Cherries (both children state tips).  The message of a cherry depends on a pattern only through the PAIR of tip states, and there
// are (S + 1)^2 pairs (S = unknown) however many patterns there are: 3,844 at 61 states against the 10^6 patterns of C5, where the 34
// cherry ops were 16.1 of the 38.6 ms of the post-order pass (two shared-memory gathers per A-fragment element, 47 % tensor-pipe use).
// So the SAME kernel runs on the enumeration of the pairs (k_dmma_cherry_ops re-points the ops at two enumerated "tips" and at a
// table), and the per-pattern message is a copy of the pair's row: same arithmetic in the same order, bit-identical messages, an
// HBM-rate kernel instead of a tensor-pipe one.
__global__ void k_dmma_cherry_ops(const phbc_op *__restrict__ ops, int n, phbc_op *__restrict__ out) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	phbc_op op = ops[k];
	op.a = 0, op.b = 1;  // rows of the enumeration [2][(S + 1)^2]; a_mat / b_mat keep the real tips' matrices, out the node's
	op.flags = PHBC_OP_TABLE | (k << 9);
	out[k] = op;
}
__global__ void k_dmma_cherry_enum(int S1, uint8_t *__restrict__ e) {
	const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= S1 * S1) return;
	e[q] = (uint8_t)(q / S1), e[S1 * S1 + q] = (uint8_t)(q % S1);
}
template <int S, bool RESCALE>
__global__ void __launch_bounds__(256) k_dmma_cherry_gather(Bufs b, const phbc_op *__restrict__ ops, const double *__restrict__ table,
                                                            const double *__restrict__ pairmax, double *__restrict__ rowmax) {
	const phbc_op op = ops[blockIdx.z];
	const int c = blockIdx.y;
	constexpr int S1 = S + 1, U = 8;  // U independent (states -> row -> store) chains per thread, resident CTAs looping: one chain per
	                                 // thread ran at 3 TB/s, and so did two million short-lived CTAs (1,700 waves of one latency chain each);
	                                 // pairs of elements per thread with 128-bit stores measured slower (5.8 against 4.1 ms at C5)
	const uint8_t *sta = b.tip_states + (size_t)op.a * b.P, *stb = b.tip_states + (size_t)op.b * b.P;
	const double *t = table + ((size_t)blockIdx.z * b.C + c) * (size_t)(S1 * S1) * S;
	double *out = (double *)dm_partial_ptr(b, op.out, c);
	const size_t n = (size_t)b.P * S, stride = (size_t)gridDim.x * (256 * U);
	for (size_t e0 = (size_t)blockIdx.x * (256 * U) + threadIdx.x; e0 < n; e0 += stride) {
		double v[U];
#pragma unroll
		for (int u = 0; u < U; u++) {
			const size_t e = e0 + (size_t)u * 256;
			if (e < n) {
				const int p = (int)(e / S), i = (int)(e - (size_t)p * S);
				const int sa = min((int)sta[p], S), sb = min((int)stb[p], S);
				v[u] = t[(size_t)(sa * S1 + sb) * S + i];
				// rescaling: the row maximum of L_n is the pair's as well ([op of the launch][C][pairs] -> [op][C][P])
				if (RESCALE && i == 0) rowmax[((size_t)blockIdx.z * b.C + c) * b.P + p] = pairmax[((size_t)blockIdx.z * b.C + c) * (S1 * S1) + sa * S1 + sb];
			}
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
			const size_t e = e0 + (size_t)u * 256;
			if (e < n) __stcs(out + e, v[u]);
		}
	}
}

// The pre-order op of a cherry n (children a, b state tips; n not the root) by the same pairs.  Its three branch terms are
//   g_a = sum_i f_i W[i] M_b[i] dP_a[i][s_a],  g_b likewise,  g_n = sum_j L_n[j] Z[j]     (k_dmma_upper_msg below)
// with W = P_n U_n and Z = U_n (f o dP_n): every one is LINEAR in U_n with a coefficient row that depends on the pattern only through
// (s_a, s_b),  g_x = sum_k U_n[k] R_x[pair][k]:   R_a = P_n^T (f o M_b o dP_a[:, s_a]),  R_b likewise,  R_n = f o (dP_n (M_a o M_b)).
// k_dmma_cherry_upper_tables builds the three rows of every pair (plain FP64: 3 (S + 1)^2 S^2 FMAs per cherry), k_dmma_cherry_upper is
// then three dot products per pattern against L2-resident rows -- no dense product per pattern at all, where the op ran two.
// An unknown state selects 1 for a probability column and the row sum for a derivative column, as tip_gather does.
template <int S>
__global__ void __launch_bounds__(192) k_dmma_cherry_upper_tables(int C, const phbc_parent_op *__restrict__ ops, const double *__restrict__ Pm,
                                                                  const double *__restrict__ dPm, const double *__restrict__ freqs,
                                                                  int include_root_freqs, double *__restrict__ tab) {
	constexpr int S1 = S + 1, PT = S1 * S1, QB = 2;  // pairs per CTA: the kernel is a latency chain per pair, a level's few cherries need the CTAs
	__shared__ double v[3][64];
	const phbc_parent_op op = ops[blockIdx.z];
	const int c = blockIdx.y;
	constexpr size_t SS = (size_t)S * S;
	const double *Pn = Pm + ((size_t)op.node * C + c) * SS, *dPn = dPm + ((size_t)op.node * C + c) * SS;
	const double *Pa = Pm + ((size_t)op.a * C + c) * SS, *dPa = dPm + ((size_t)op.a * C + c) * SS;
	const double *Pb = Pm + ((size_t)op.b * C + c) * SS, *dPb = dPm + ((size_t)op.b * C + c) * SS;
	const int which = threadIdx.x >> 6, k = threadIdx.x & 63;
	double *t = tab + (((size_t)blockIdx.z * C + c) * 3 + which) * (size_t)PT * S;
	const int q1 = min(PT, (int)(blockIdx.x + 1) * QB);
	for (int q = blockIdx.x * QB; q < q1; q++) {
		const int sa = q / S1, sb = q - sa * S1;
		__syncthreads();
		if (threadIdx.x < S) {
			const int i = threadIdx.x;
			const double f = include_root_freqs ? 1.0 : freqs[i];
			double ma = 1.0, mb = 1.0, da = 0.0, db = 0.0;
			if (sa < S) ma = Pa[i * S + sa], da = dPa[i * S + sa];
			else
				for (int j = 0; j < S; j++) da += dPa[i * S + j];
			if (sb < S) mb = Pb[i * S + sb], db = dPb[i * S + sb];
			else
				for (int j = 0; j < S; j++) db += dPb[i * S + j];
			v[0][i] = ma * mb, v[1][i] = f * mb * da, v[2][i] = f * ma * db;
		}
		__syncthreads();
		if (k < S) {
			double acc[4] = {0.0, 0.0, 0.0, 0.0};  // four chains, summed in a fixed order
			if (which == 0) {
#pragma unroll 4
				for (int j = 0; j < S; j++) acc[j & 3] = fma(dPn[k * S + j], v[0][j], acc[j & 3]);
			} else {
#pragma unroll 4
				for (int i = 0; i < S; i++) acc[i & 3] = fma(Pn[i * S + k], v[which][i], acc[i & 3]);
			}
			const double r = (acc[0] + acc[1]) + (acc[2] + acc[3]);
			t[(size_t)q * S + k] = which == 0 ? r * (include_root_freqs ? 1.0 : freqs[k]) : r;
		}
	}
}
template <int S>
__global__ void __launch_bounds__(256) k_dmma_cherry_upper(Bufs b, const phbc_parent_op *__restrict__ ops, const double *__restrict__ tab,
                                                           const double *__restrict__ weights, const double *__restrict__ pattern_lnl, int pstride,
                                                           double *__restrict__ partial, int scaled) {
	constexpr int S1 = S + 1, PT = S1 * S1;
	__shared__ double red[8][3];
	const phbc_parent_op op = ops[blockIdx.z];
	const int c = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	// eight lanes to a pattern, four patterns to a warp.  Measured alternatives (C5, ms per evaluation for the 34 cherries): a warp to a
	// pattern 11.3; this 6.3; every load hoisted above the products (98 registers, 16 warps per SM) 6.8; sixteen lanes to a pattern with
	// hoisted loads 7.3 -- the kernel lives on resident warps (64 per SM at 48 registers), not on loads in flight per warp
	const int l8 = lane & 7, sub = lane >> 3;
	const double *u = b.upper + ((size_t)op.node * b.C + c) * (size_t)b.P * S;
	const double *tn = tab + ((size_t)blockIdx.z * b.C + c) * 3 * (size_t)PT * S, *ta = tn + (size_t)PT * S, *tb = ta + (size_t)PT * S;
	const uint8_t *sta = b.tip_states + (size_t)op.a * b.P, *stb = b.tip_states + (size_t)op.b * b.P;
	double tot_n = 0.0, tot_a = 0.0, tot_b = 0.0;
	// the states of the NEXT turn are requested a turn ahead: one dependent latency (states -> rows -> sums) less per turn
	int nsa = S, nsb = S;
	{
		const int p = (blockIdx.x * 8 + warp) * 4 + sub;
		if (p < b.P) nsa = sta[p], nsb = stb[p];
	}
	for (int p0 = (blockIdx.x * 8 + warp) * 4; p0 < b.P; p0 += gridDim.x * 32) {
		const int p = p0 + sub;
		const bool live = p < b.P;
		const int sa = nsa, sb = nsb;
		{
			const int np = p + gridDim.x * 32;
			if (np < b.P) nsa = sta[np], nsb = stb[np];
		}
		double gn = 0.0, ga = 0.0, gb = 0.0, wl = 0.0;
		if (live) {
			const size_t row = (size_t)(min(sa, S) * S1 + min(sb, S)) * S;
			// rescaling: U_n carries exp(-sf[N + n]); the children are tips (see k_dmma_upper)
			if (l8 == 0) wl = __ldg(weights + p) / exp(__ldg(pattern_lnl + p) - (scaled ? b.sf[(size_t)(b.N + op.node) * b.P + p] : 0.0));
			const double *up = u + (size_t)p * S;
#pragma unroll
			for (int k0 = 0; k0 < S; k0 += 8) {
				const int k = k0 + l8;
				if (k < S) {
					const double x = __ldcs(up + k);
					gn = fma(x, tn[row + k], gn), ga = fma(x, ta[row + k], ga), gb = fma(x, tb[row + k], gb);
				}
			}
		}
#pragma unroll
		for (int off = 1; off < 8; off <<= 1) {
			gn += __shfl_xor_sync(0xffffffffu, gn, off), ga += __shfl_xor_sync(0xffffffffu, ga, off), gb += __shfl_xor_sync(0xffffffffu, gb, off);
		}
		tot_n = fma(gn, wl, tot_n), tot_a = fma(ga, wl, tot_a), tot_b = fma(gb, wl, tot_b);  // wl = 0 off the group's first lane
	}
	tot_n = phb_warp_sum(tot_n), tot_a = phb_warp_sum(tot_a), tot_b = phb_warp_sum(tot_b);
	if (lane == 0) red[warp][0] = tot_n, red[warp][1] = tot_a, red[warp][2] = tot_b;
	__syncthreads();
	if (threadIdx.x == 0) {
		double sn = 0.0, sa = 0.0, sb = 0.0;
		for (int w = 0; w < 8; w++) sn += red[w][0], sa += red[w][1], sb += red[w][2];
		partial[((size_t)op.node * b.C + c) * pstride + blockIdx.x] = sn;
		partial[((size_t)op.a * b.C + c) * pstride + blockIdx.x] = sa;
		partial[((size_t)op.b * b.C + c) * pstride + blockIdx.x] = sb;
	}
}

// ---------------------------------------------------------------------------------------------
// k_dmma_upper_msg: per PARENT n with children a, b.  W = P_n U_n (DMMA), the children's messages M_a, M_b come straight from
// the lower buffers (or the tips' matrix columns), U_a = W o M_b and U_b = W o M_a are stored for internal children.  Branch gradients
// in ADJOINT form: the op of n reduces n's OWN branch, sum_i f_i U_n[i] (dP_n L_n)[i] = sum_j L_n[j] Z[j] with L_n = M_a o M_b and
// Z = U_n (f o dP_n) -- a second DMMA product that shares the A fragments of W -- so an internal branch sees the reference's dP
// entries themselves (Q (P L) differs from dP L by 1e-9 relative where codon probabilities of order t^3 dominate a pattern) at two
// products per op instead of up to three; the branches of TIP children are reduced here too, from the tips' transposed dP images.
// The accumulator-layout copies of M_a, M_b are picked out of the staged chunks as they pass, so each message is read from HBM once.
// Shared-memory image slots: 0 P_n | 1 a: tip image | 2 a: tip dP image | 3 b: tip image | 4 b: tip dP image; the adjoint image of n
// sits in slot 1 when a is internal, else in slot 3 when b is internal, else (both tips) in a sixth slot at 20 states and in slot 1 at
// 61 states, where five images are all an SM can hold -- a's tip image is then read from global memory (one L2-resident column per pattern).
// ---------------------------------------------------------------------------------------------
template <int S, int MT, int NSPLIT, int WM, int NST = 0, int GB = 1, bool RESCALE = false>
__global__ void __launch_bounds__(32 * WM * NSPLIT) k_dmma_upper_msg(Bufs b, const phbc_parent_op *__restrict__ ops, const double *__restrict__ img,
                                                                   const double *__restrict__ freqs, const double *__restrict__ weights,
                                                                   const double *__restrict__ pattern_lnl, int include_root_freqs, int pstride,
                                                                   double *__restrict__ partial, int nslots, int scaled, double *__restrict__ rowmax) {
	using Sh = DmmaShape<S>;
	constexpr int NTW = Sh::NT / NSPLIT;
	constexpr int NWARPS = WM * NSPLIT;
	extern __shared__ __align__(128) unsigned char smraw[];
	uint64_t *bar = reinterpret_cast<uint64_t *>(smraw);
	double *sm = reinterpret_cast<double *>(smraw + 128);
	const phbc_parent_op op = ops[blockIdx.z];
	const int c = blockIdx.y;
	const bool is_root_rt = op.flags & 1;
	const bool a_tip_rt = op.a < b.T, b_tip_rt = op.b < b.T;
	double *mP = sm, *mA = sm + Sh::IMG, *dA = sm + 2 * Sh::IMG, *mB = sm + 3 * Sh::IMG, *dB = sm + 4 * Sh::IMG;
	const bool cherry_rt = a_tip_rt && b_tip_rt;
	const bool sixth = nslots >= 6;  // room for a slot of its own (20 states); at 61 states five images are all an SM can hold
	const double *mZ = !a_tip_rt ? mA : (!b_tip_rt ? mB : (sixth ? sm + 5 * Sh::IMG : mA));  // adjoint image (f o dP_n) transposed; not staged at the root
	const double *tipA = (cherry_rt && !is_root_rt && !sixth) ? img + ((size_t)op.a * b.C + c) * Sh::IMG : mA;  // a's transposed P image
	double *aux = sm + nslots * Sh::IMG;  // 2 slots when both children are internal (P_n, Q), else 5
	double *fq = aux, *wroot = aux + Sh::NP, *red = aux + 2 * Sh::NP;
	const double *rsA = dA + Sh::S * Sh::NP, *rsB = dB + Sh::S * Sh::NP;
	if (threadIdx.x == 0) {
		const size_t dimg = (size_t)b.N * b.C * Sh::IMG;
		mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		const bool a_img = a_tip_rt && tipA == mA;  // a's tip image goes to shared memory unless the adjoint image needs its slot
		const int nimg = (is_root_rt ? 0 : 2) + (a_tip_rt ? 1 : 0) + (a_img ? 1 : 0) + (b_tip_rt ? 2 : 0);
		mbar_expect_tx(bar, nimg * Sh::IMG * 8);
		if (!is_root_rt) {
			bulk_g2s(mP, img + ((size_t)op.node * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
			bulk_g2s(const_cast<double *>(mZ), img + dimg + ((size_t)op.node * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
		}
		if (a_tip_rt) {
			if (a_img) bulk_g2s(mA, img + ((size_t)op.a * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
			bulk_g2s(dA, img + dimg + ((size_t)op.a * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
		}
		if (b_tip_rt) {
			bulk_g2s(mB, img + ((size_t)op.b * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
			bulk_g2s(dB, img + dimg + ((size_t)op.b * b.C + c) * Sh::IMG, Sh::IMG * 8, bar);
		}
	}
	for (int i = threadIdx.x; i < Sh::NP; i += blockDim.x) {
		const double f = i < S ? freqs[i] : 0.0;
		fq[i] = i < S ? (include_root_freqs ? 1.0 : f) : 0.0;
		wroot[i] = i < S ? (include_root_freqs ? f : 1.0) : 0.0;
	}
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int wm = warp / NSPLIT, n0 = (warp % NSPLIT) * NTW;
	const int r = lane >> 2, q = lane & 3;
	constexpr int TP = WM * MT * 8;
	const double *xw = b.upper + ((size_t)op.node * b.C + c) * (size_t)b.P * S;
	const double *xa = a_tip_rt ? nullptr : dm_partial_ptr(b, op.a, c), *xb = b_tip_rt ? nullptr : dm_partial_ptr(b, op.b, c);
	double *Ua = b.upper + ((size_t)op.a * b.C + c) * (size_t)b.P * S;
	double *Ub = b.upper + ((size_t)op.b * b.C + c) * (size_t)b.P * S;
	const int ntiles = (b.P + TP - 1) / TP;
	double tot_a = 0.0, tot_b = 0.0, tot_n = 0.0;
	using AS = AStage<Sh, MT, 3, NST>;
	constexpr int GT = 32 * NSPLIT;
	double *abuf = red + 4 * NWARPS + wm * AS::NSTAGE * AS::STG;  // ring operands: U_n | M_b | M_a
	const int gl = (warp % NSPLIT) * 32 + lane;
	mbar_wait(bar, 0);
	dispatch3(is_root_rt, a_tip_rt, b_tip_rt, [&](auto ROOT, auto ATIP, auto BTIP) {
	constexpr bool is_root = decltype(ROOT)::value, a_tip = decltype(ATIP)::value, b_tip = decltype(BTIP)::value;
	AFill<Sh, MT, 3, GT, NST, GB> plan;
	plan.init(gl);
	int f_tile = blockIdx.x, f_ch = 0, f_stage = 0;
	auto fill_next = [&]() {
		if (f_tile < ntiles) {
			double *stg = abuf + f_stage * AS::STG;
			const int fp0 = f_tile * TP + wm * MT * 8;
			if (!is_root) plan.fill(stg, xw, fp0, b.P, f_ch);
			if (!b_tip) plan.fill(stg + AS::OPB, xb, fp0, b.P, f_ch);
			if (!a_tip) plan.fill(stg + 2 * AS::OPB, xa, fp0, b.P, f_ch);
		}
		cp_async_commit();
		f_stage = f_stage + 1 == AS::NSTAGE ? 0 : f_stage + 1;
		if (++f_ch == Sh::NCH) f_ch = 0, f_tile += gridDim.x;
	};
#pragma unroll
	for (int pf = 0; pf < AS::NSTAGE - 1; pf++) fill_next();
	int c_stage = 0;
	double bP = 0.0, bZ = 0.0;
	{
		const int off = (n0 * 8 + r) * Sh::LD + q;
		if (!is_root) bP = mP[off], bZ = mZ[off];
	}
	for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const int p0 = tile * TP + wm * MT * 8;
		double wk[MT], lk[MT];
#pragma unroll
		for (int m = 0; m < MT; m++) {
			const int p = p0 + 8 * m + r;
			wk[m] = p < b.P ? __ldg(weights + p) : 0.0;
			lk[m] = p < b.P ? __ldg(pattern_lnl + p) : 0.0;
			if (RESCALE && scaled && p < b.P) {  // every branch term of this op carries exp(-(sf[N + n] + sf[a] + sf[b])), see k_dmma_upper
				double e = is_root ? 0.0 : b.sf[(size_t)(b.N + op.node) * b.P + p];
				if (!a_tip) e += b.sf[(size_t)op.a * b.P + p];
				if (!b_tip) e += b.sf[(size_t)op.b * b.P + p];
				lk[m] -= e;
			}
		}
		double W[MT][NTW][2], Z[MT][NTW][2], Mb[MT][NTW][2], Ma[MT][NTW][2];
		zero_acc<MT, NTW>(W);
		zero_acc<MT, NTW>(Z);
		zero_acc<MT, NTW>(Mb);
		zero_acc<MT, NTW>(Ma);
		if (!(is_root && a_tip && b_tip))
#pragma unroll
		for (int ch = 0; ch < Sh::NCH; ch++) {
			cp_async_wait<AS::NSTAGE - 2>();
			group_sync<NSPLIT>(wm);
			fill_next();
			const double *st = abuf + c_stage * AS::STG;
			c_stage = c_stage + 1 == AS::NSTAGE ? 0 : c_stage + 1;
			double cw[MT][Sh::KCH];
			if (!is_root) read_frags<Sh, MT, 3, NST>(st, lane, cw);
			// the messages again in accumulator layout (row 8 m + r, columns (n0 + j) 8 + 2 q, + 1) while their chunk is staged
#pragma unroll
			for (int j = 0; j < NTW; j++) {
				const int col = (n0 + j) * 8 + 2 * q, lc = col - ch * AS::KC;
				if (lc >= 0 && lc < AS::KC && col < S) {
#pragma unroll
					for (int m = 0; m < MT; m++) {
						if (!b_tip) {
							const double2 v = *reinterpret_cast<const double2 *>(st + AS::OPB + (8 * m + r) * AS::RS + lc);
							Mb[m][j][0] = v.x, Mb[m][j][1] = v.y;
						}
						if (!a_tip) {
							const double2 v = *reinterpret_cast<const double2 *>(st + 2 * AS::OPB + (8 * m + r) * AS::RS + lc);
							Ma[m][j][0] = v.x, Ma[m][j][1] = v.y;
						}
					}
				}
			}
#pragma unroll
			for (int tt = 0; tt < Sh::KCH; tt++) {
				const int t = ch * Sh::KCH + tt;
				if (t < Sh::KT) {
#pragma unroll
					for (int j = 0; j < NTW; j++) {
						const int nj = j + 1 < NTW ? j + 1 : 0, nt = j + 1 < NTW ? t : (t + 1 < Sh::KT ? t + 1 : 0);
						const int noff = ((n0 + nj) * 8 + r) * Sh::LD + 4 * nt + q;
						double nP = 0.0, nZ = 0.0;
						if (!is_root) {
							nP = mP[noff], nZ = mZ[noff];
dmma_mtiles<MT, NTW, Sh::KCH>(W, cw, j, tt, bP);
dmma_mtiles<MT, NTW, Sh::KCH>(Z, cw, j, tt, bZ);
						}
						bP = nP, bZ = nZ;
					}
				}
			}
		}
		if (is_root) {
#pragma unroll
			for (int m = 0; m < MT; m++)
#pragma unroll
				for (int j = 0; j < NTW; j++) {
					const int i = (n0 + j) * 8 + 2 * q;
					W[m][j][0] = wroot[i], W[m][j][1] = wroot[i + 1];
				}
		}
		double Da[MT][NTW][2], Db[MT][NTW][2];  // dP columns of the TIP children (dead code for internal children)
		if (b_tip) {
			const uint8_t *st = b.tip_states + (size_t)op.b * b.P;
			tip_gather<Sh, MT, NTW, true>(st, mB, nullptr, p0, b.P, n0, lane, Mb);
			tip_gather<Sh, MT, NTW, false>(st, dB, rsB, p0, b.P, n0, lane, Db);
		}
		if (a_tip) {
			const uint8_t *st = b.tip_states + (size_t)op.a * b.P;
			tip_gather<Sh, MT, NTW, true>(st, tipA, nullptr, p0, b.P, n0, lane, Ma);
			tip_gather<Sh, MT, NTW, false>(st, dA, rsA, p0, b.P, n0, lane, Da);
		}
		double gn[MT];
#pragma unroll
		for (int m = 0; m < MT; m++) {
			gn[m] = 0.0;
#pragma unroll
			for (int j = 0; j < NTW; j++) {
				// n's own branch: sum_j L_n[j] Z[j], L_n = M_a o M_b (the weights f sit in the adjoint image; padding columns are zero)
				if (!is_root) gn[m] = fma(Ma[m][j][0] * Mb[m][j][0], Z[m][j][0], fma(Ma[m][j][1] * Mb[m][j][1], Z[m][j][1], gn[m]));
				const double ua0 = W[m][j][0] * Mb[m][j][0], ua1 = W[m][j][1] * Mb[m][j][1];
				const double ub0 = W[m][j][0] * Ma[m][j][0], ub1 = W[m][j][1] * Ma[m][j][1];
				Mb[m][j][0] = ua0, Mb[m][j][1] = ua1;
				Ma[m][j][0] = ub0, Ma[m][j][1] = ub1;
			}
		}
		if (!a_tip) store_tile<Sh, MT, NTW>(Ua, p0, b.P, n0, lane, Mb);
		if (!b_tip) store_tile<Sh, MT, NTW>(Ub, p0, b.P, n0, lane, Ma);
		if (RESCALE && rowmax) {  // rescaling: slots 2 z (child a) and 2 z + 1 (child b) of this launch, as k_dmma_upper
			const size_t row = (((size_t)(2 * blockIdx.z) * b.C + c) * NSPLIT + warp % NSPLIT) * b.P;
			store_rowmax<Sh, MT, NTW>(rowmax, row, p0, b.P, n0, lane, Mb, !a_tip);
			store_rowmax<Sh, MT, NTW>(rowmax, row + (size_t)b.C * NSPLIT * b.P, p0, b.P, n0, lane, Ma, !b_tip);
		}
#pragma unroll
		for (int m = 0; m < MT; m++) {
			double ga = 0.0, gb = 0.0, g0 = gn[m];
#pragma unroll
			for (int j = 0; j < NTW; j++) {
				const int i = (n0 + j) * 8 + 2 * q;
				if (a_tip) ga = fma(fq[i] * Mb[m][j][0], Da[m][j][0], fma(fq[i + 1] * Mb[m][j][1], Da[m][j][1], ga));
				if (b_tip) gb = fma(fq[i] * Ma[m][j][0], Db[m][j][0], fma(fq[i + 1] * Ma[m][j][1], Db[m][j][1], gb));
			}
			ga += __shfl_xor_sync(0xffffffffu, ga, 1), gb += __shfl_xor_sync(0xffffffffu, gb, 1), g0 += __shfl_xor_sync(0xffffffffu, g0, 1);
			ga += __shfl_xor_sync(0xffffffffu, ga, 2), gb += __shfl_xor_sync(0xffffffffu, gb, 2), g0 += __shfl_xor_sync(0xffffffffu, g0, 2);
			const double wl = wk[m] / exp(lk[m]);
			tot_a = fma(ga, wl, tot_a);
			tot_b = fma(gb, wl, tot_b);
			tot_n = fma(g0, wl, tot_n);
		}
	}
	});
	tot_a = q == 0 ? tot_a : 0.0, tot_b = q == 0 ? tot_b : 0.0, tot_n = q == 0 ? tot_n : 0.0;
	tot_a = phb_warp_sum(tot_a), tot_b = phb_warp_sum(tot_b), tot_n = phb_warp_sum(tot_n);
	if (lane == 0) red[3 * warp] = tot_a, red[3 * warp + 1] = tot_b, red[3 * warp + 2] = tot_n;
	__syncthreads();
	if (threadIdx.x == 0) {
		double sa = 0.0, sb = 0.0, sn = 0.0;
		for (int w = 0; w < NWARPS; w++) sa += red[3 * w], sb += red[3 * w + 1], sn += red[3 * w + 2];
		// every non-root node is written exactly once: a tip by its parent's op, an internal node by its own
		if (a_tip_rt) partial[((size_t)op.a * b.C + c) * pstride + blockIdx.x] = sa;
		if (b_tip_rt) partial[((size_t)op.b * b.C + c) * pstride + blockIdx.x] = sb;
		if (!is_root_rt) partial[((size_t)op.node * b.C + c) * pstride + blockIdx.x] = sn;
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
	static constexpr int UWM_II = 4;  // message-form upper kernel, parents without tip children
};
struct DmmaConfigCodon {  // 256 threads, 32 patterns per tile, n-tiles split over two warps
	static constexpr int MT = 1, NSPLIT = 2, WM = 4;
	static constexpr int UMT = 1, UNSPLIT = 2, UWM = 4;
	static constexpr int UWM_II = 8;  // two image slots instead of five leave room for a 16-warp CTA sharing them
};
// codon models: 64 codons minus the stop codons of the genetic code (1 ... 4 of them), the reference's ">= 60 states" dispatch
// (treelikelihood.c:1086-1090); all four state counts pad to the same 64 x 64 tile grid
template <> struct DmmaConfig<60> : DmmaConfigCodon {};
template <> struct DmmaConfig<61> : DmmaConfigCodon {};
template <> struct DmmaConfig<62> : DmmaConfigCodon {};
template <> struct DmmaConfig<63> : DmmaConfigCodon {};

// Geometry of the MESSAGE-form kernels, by state count and tuning variant (PHB_OPT_TUNE; 0 = what ships, chosen from the measured
// table in profiles/): m-tiles per warp, n-split, m-groups per CTA, cp.async ring depth (0 = AStage's default) and granule size.
template <int S, int VAR>
struct MsgCfg;
template <int MT_, int NSPLIT_, int WM_, int LNST_, int LGB_, int UMT_, int UNSPLIT_, int UWM_, int UWM_II_, int UNST_, int UGB_>
struct MsgGeom {
	static constexpr int MT = MT_, NSPLIT = NSPLIT_, WM = WM_, LNST = LNST_, LGB = LGB_;
	static constexpr int UMT = UMT_, UNSPLIT = UNSPLIT_, UWM = UWM_, UWM_II = UWM_II_, UNST = UNST_, UGB = UGB_;
};
//                                      lower: MT NS WM NST GB | upper: MT NS WM WM_II NST GB
template <> struct MsgCfg<20, 0> : MsgGeom<2, 1, 4, 0, 2, 1, 1, 4, 4, 0, 2> {};  // 16-byte granules (rows of 160 bytes)
template <> struct MsgCfg<20, 1> : MsgGeom<2, 1, 4, 0, 1, 1, 1, 4, 4, 0, 1> {};  // round-1 geometry: 8-byte granules
template <> struct MsgCfg<20, 2> : MsgGeom<2, 1, 4, 3, 2, 1, 1, 4, 4, 3, 2> {};  // ring of 3: two tiles ahead
template <> struct MsgCfg<20, 3> : MsgGeom<1, 1, 4, 0, 2, 1, 1, 4, 4, 0, 2> {};  // 32-pattern tiles in the lower pass
template <> struct MsgCfg<20, 4> : MsgGeom<2, 1, 8, 0, 2, 1, 1, 8, 8, 0, 2> {};  // 256-thread CTAs
template <> struct MsgCfg<20, 5> : MsgGeom<2, 1, 4, 0, 2, 2, 1, 4, 4, 0, 2> {};  // 64-pattern tiles in the upper pass
template <> struct MsgCfg<20, 6> : MsgGeom<4, 1, 2, 0, 2, 2, 1, 2, 2, 0, 2> {};  // 64-thread CTAs, wide warps
template <> struct MsgCfg<61, 0> : MsgGeom<1, 2, 4, 0, 1, 1, 2, 4, 8, 0, 1> {};
template <> struct MsgCfg<60, 0> : MsgCfg<61, 0> {};
template <> struct MsgCfg<62, 0> : MsgCfg<61, 0> {};
template <> struct MsgCfg<63, 0> : MsgCfg<61, 0> {};
template <> struct MsgCfg<61, 1> : MsgGeom<1, 2, 4, 2, 1, 1, 2, 4, 8, 2, 1> {};  // ring of 2 (more CTAs per SM)
template <> struct MsgCfg<61, 2> : MsgGeom<1, 2, 4, 0, 1, 1, 2, 4, 4, 0, 1> {};  // no wide variant for parents of two internal nodes
template <> struct MsgCfg<61, 3> : MsgGeom<2, 2, 2, 0, 1, 2, 2, 2, 4, 0, 1> {};  // two m-tiles per warp: every B fragment feeds two DMMAs (half the shared-memory reads)
template <> struct MsgCfg<61, 4> : MsgGeom<2, 2, 2, 0, 1, 1, 2, 4, 8, 0, 1> {};  // ... in the lower pass only
#define PHBC_MSG_VARIANTS_20 7
#define PHBC_MSG_VARIANTS_61 5

bool phbc_dmma_supported(const phbc_ctx *ctx, const phbc_eval_opts *o) {
	(void)o;
	return ctx->S == 20 || (ctx->S >= 60 && ctx->S <= 63);
}

// Pattern chunks for one launch of `units` = C x ops (op, category) pairs on `slots` resident CTA slots: the chunk count whose
// CTA total fills whole waves best, at most ~4 waves (every CTA pays the matrix staging once), never more than the tile count.
static int pick_chunks(int slots, int units, int ntiles) {
	int kmax = (4 * slots) / units;
	if (kmax < 1) kmax = 1;
	if (kmax > ntiles) kmax = ntiles;
	int best = 1;
	double best_eff = 0.0;
	for (int k = 1; k <= kmax; k++) {
		const long long ctas = (long long)units * k;
		const long long waves = (ctas + slots - 1) / slots;
		const double eff = (double)ctas / (double)(waves * slots);
		if (eff > best_eff + 1e-9) best_eff = eff, best = k;
	}
	return best;
}


// packed matrix images [P | dP][node][category][IMG] from the per-node matrices
template <int S>
static int dmma_pack(phbc_ctx *ctx, bool adjoint = false, int include_root_freqs = 0, int tip_images = -1) {
	using Sh = DmmaShape<S>;
	const int C = ctx->C, N = ctx->N;
	const size_t img_bytes = (size_t)2 * N * C * Sh::IMG * sizeof(double);
	if (img_bytes > ctx->dmma_img_bytes) {
		PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
		if (ctx->d_dmma_img) cudaFree(ctx->d_dmma_img);
		ctx->d_dmma_img = NULL;
		ctx->dmma_img_bytes = 0;
		PHBC_CHECK(cudaMalloc((void **)&ctx->d_dmma_img, img_bytes));
		ctx->dmma_img_bytes = img_bytes;
	}
	ctx->dmma_pack_adjoint = adjoint, ctx->dmma_pack_irf = include_root_freqs;
	ctx->dmma_pack_tips = tip_images < 0 ? (ctx->tip_kind == PHBC_TIP_STATES ? 1 : 0) : tip_images;
	k_dmma_pack<Sh><<<dim3(N, C, 2), 128, 0, ctx->stream>>>(ctx->T, N, C, ctx->root, ctx->dmma_pack_tips, ctx->d_P, ctx->d_dP, ctx->d_dmma_img,
	                                                         adjoint ? 1 : 0, ctx->d_freqs, include_root_freqs);
	ctx->launches++;
	PHBC_CHECK(cudaGetLastError());
	return 0;
}

// one level of K1-K4 ops (device list): wave-filling launch geometry, every CTA stages its two matrices once
// Row maxima scratch of the rescaled passes: `per_op` slots per op of C x nsplit x P doubles, at most 512 MB (a level wider than
// that is launched in chunks).  *zmax: ops per launch.  Categories must tile a warp (k_dmma_scale_from_max's butterfly).
static bool dmma_rowmax_usable(const phbc_ctx *ctx) { return ctx->C >= 1 && ctx->C <= 32 && (ctx->C & (ctx->C - 1)) == 0; }
static int dmma_rowmax_reserve(phbc_ctx *ctx, int nsplit, int per_op, int widest, int *zmax) {
	const size_t slot = (size_t)ctx->C * nsplit * ctx->P * sizeof(double), cap = (size_t)512 << 20;
	size_t z = cap / (slot * per_op);
	if (z < 1) z = 1;
	if (z > (size_t)widest) z = widest > 0 ? widest : 1;
	if (z > 65535) z = 65535;
	const size_t bytes = z * per_op * slot;
	if (bytes > ctx->rowmax_bytes) {
		PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
		if (ctx->d_rowmax) cudaFree(ctx->d_rowmax);
		ctx->d_rowmax = NULL, ctx->rowmax_bytes = 0;
		PHBC_CHECK(cudaMalloc((void **)&ctx->d_rowmax, bytes));
		ctx->rowmax_bytes = bytes;
	}
	*zmax = (int)z;
	return 0;
}
// the message-form passes scale a whole level at once: [slots][C][n-split][P], 1 / S of the level's partials
static int dmma_rowmax_reserve_level(phbc_ctx *ctx, size_t bytes) {
	if (bytes <= ctx->rowmax_bytes) return 0;
	PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
	if (ctx->d_rowmax) cudaFree(ctx->d_rowmax);
	ctx->d_rowmax = NULL, ctx->rowmax_bytes = 0;
	PHBC_CHECK(cudaMalloc((void **)&ctx->d_rowmax, bytes));
	ctx->rowmax_bytes = bytes;
	return 0;
}
static int dmma_scale_from_max(phbc_ctx *ctx, const Bufs &b, const phbc_op *d_ops, int cnt, const int *slot_of, int slot0, int nslots, int nsplit,
                               double threshold) {
	const unsigned gx = (unsigned)(((size_t)ctx->P * ctx->C + 127) / 128);
	for (int y0 = 0; y0 < cnt; y0 += 65535) {
		const int yc = cnt - y0 < 65535 ? cnt - y0 : 65535;
		k_dmma_scale_from_max<<<dim3(gx, yc), 128, 0, ctx->stream>>>(b, d_ops + y0, ctx->d_rowmax, slot_of, slot_of ? slot0 : 0, nslots, nsplit, threshold);
		ctx->launches++;
	}
	return 0;
}

template <int S>
static int dmma_lower_ops(phbc_ctx *ctx, const phbc_op *d_ops, int cnt, bool scale = false, double threshold = 0.0) {
	using Sh = DmmaShape<S>;
	using Cf = DmmaConfig<S>;
	const int C = ctx->C, P = ctx->P;
	Bufs b = phbc_make_bufs(ctx);
	auto lower = k_dmma_lower<S, Cf::MT, Cf::NSPLIT, Cf::WM>;
	const size_t lsmem = 128 + (2 * Sh::IMG + Cf::WM * AStage<Sh, Cf::MT, 2>::NSTAGE * AStage<Sh, Cf::MT, 2>::STG) * sizeof(double);
	PHBC_CHECK(cudaFuncSetAttribute(lower, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem));
	const int lthreads = 32 * Cf::WM * Cf::NSPLIT, ltiles = (P + Cf::WM * Cf::MT * 8 - 1) / (Cf::WM * Cf::MT * 8);
	int lper_sm = 1;
	PHBC_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lper_sm, lower, lthreads, lsmem));
	if (lper_sm < 1) lper_sm = 1;
	int zmax = 65535, rc;
	const bool fused = scale && dmma_rowmax_usable(ctx);
	if (fused && (rc = dmma_rowmax_reserve(ctx, Cf::NSPLIT, 1, cnt, &zmax))) return rc;
	for (int z0 = 0; z0 < cnt; z0 += zmax) {
		const int zc = cnt - z0 < zmax ? cnt - z0 : zmax;
		lower<<<dim3(pick_chunks(lper_sm * ctx->num_sms, C * zc, ltiles), C, zc), lthreads, lsmem, ctx->stream>>>(b, d_ops + z0, ctx->d_dmma_img,
		                                                                                                         fused ? ctx->d_rowmax : nullptr);
		ctx->launches++;
		if (fused && (rc = dmma_scale_from_max(ctx, b, d_ops + z0, zc, nullptr, 0, zc, Cf::NSPLIT, threshold))) return rc;
	}
	PHBC_CHECK(cudaGetLastError());
	if (scale && !fused) return phbc_generic_scale_ops(ctx, d_ops, cnt, threshold);
	return 0;
}

// 0/1 tip partials as state codes (what SitePattern_get_partials produces for unambiguous data and gaps: phycpp's default tip mode,
// examples/fluA/*.json): one-hot -> the state, all ones -> S (unknown).  Anything else -- an ambiguity SET -- raises *bad: its
// message is a sum of columns, not a gather, and the evaluation stays on the kernels that read the partials.
__global__ void k_dmma_encode_tips(size_t n, int S, const double *__restrict__ partials, uint8_t *__restrict__ codes, int *__restrict__ bad) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double *v = partials + i * S;
	int ones = 0, first = -1;
	for (int j = 0; j < S; j++) {
		if (v[j] == 1.0) {
			if (first < 0) first = j;
			ones++;
		} else if (v[j] != 0.0) *bad = 1;
	}
	if (ones != 1 && ones != S) *bad = 1;
	codes[i] = (uint8_t)(ones == 1 ? first : S);
}

// 1 when the tips are states or encode as states (cached until the next tip upload)
static int dmma_tips_as_states(phbc_ctx *ctx, int *usable) {
	*usable = 1;
	if (ctx->tip_kind == PHBC_TIP_STATES) return 0;
	if (!ctx->enc_states_valid) {
		const size_t n = (size_t)ctx->T * ctx->P;
		if (!ctx->d_enc_states) PHBC_CHECK(cudaMalloc((void **)&ctx->d_enc_states, n));
		if (!ctx->d_dw_bad) PHBC_CHECK(cudaMalloc((void **)&ctx->d_dw_bad, sizeof(int)));
		PHBC_CHECK(cudaMemsetAsync(ctx->d_dw_bad, 0, sizeof(int), ctx->stream));
		k_dmma_encode_tips<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, ctx->S, ctx->d_tip_partials, ctx->d_enc_states, ctx->d_dw_bad);
		ctx->launches++;
		int bad = 0;
		PHBC_CHECK(cudaMemcpyAsync(&bad, ctx->d_dw_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
		PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
		ctx->enc_states_bad = bad != 0;
		ctx->enc_states_valid = true;
	}
	*usable = !ctx->enc_states_bad;
	return 0;
}
// the buffers as the message-form kernels see them: tips are states, encoded from 0/1 partials where the caller uploaded those
static Bufs dmma_msg_bufs(phbc_ctx *ctx) {
	Bufs b = phbc_make_bufs(ctx);
	if (ctx->tip_kind != PHBC_TIP_STATES) b.tip_states = ctx->d_enc_states, b.tip_kind = PHBC_TIP_STATES;
	return b;
}

// cherry tables (k_dmma_cherry_gather, k_dmma_cherry_upper): used where there are (4 x) more patterns than state pairs;
// PHB_OPT_TUNE 20 / 21: always / never.  One buffer serves both passes, at most 256 MB: wider levels go in chunks of *zmax ops.
template <int S>
static bool dmma_cherry_tables_on(const phbc_ctx *ctx) {
	return ctx->tune != 21 && (ctx->tune == 20 || ctx->P >= 4 * (S + 1) * (S + 1));
}
static int dmma_cherry_reserve(phbc_ctx *ctx, size_t per_op, int count, int *zmax) {
	size_t z = ((size_t)256 << 20) / per_op;
	if (z < 1) z = 1;
	if (z > (size_t)count) z = count;
	if (z > 32767) z = 32767;  // gridDim.z, and phbc_op.flags >> 9 stays positive
	if (z * per_op > ctx->cherry_tab_bytes) {
		PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
		if (ctx->d_cherry_tab) cudaFree(ctx->d_cherry_tab);
		ctx->d_cherry_tab = NULL, ctx->cherry_tab_bytes = 0;
		PHBC_CHECK(cudaMalloc((void **)&ctx->d_cherry_tab, z * per_op));
		ctx->cherry_tab_bytes = z * per_op;
	}
	if ((int)z > ctx->cherry_ops_cap) {
		PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
		if (ctx->d_cherry_ops) cudaFree(ctx->d_cherry_ops);
		ctx->d_cherry_ops = NULL, ctx->cherry_ops_cap = 0;
		PHBC_CHECK(cudaMalloc((void **)&ctx->d_cherry_ops, z * sizeof(phbc_op)));
		ctx->cherry_ops_cap = (int)z;
	}
	*zmax = (int)z;
	return 0;
}

// one level of message-form lower ops: three launches by the number of tip children (the device op list is sorted that way), so
// that ops without tip children do not pay shared memory for tip images (61 states: 3 CTAs per SM instead of 1)
template <int S, int VAR>
static int dmma_lower_msg_level(phbc_ctx *ctx, int level, const phbc_eval_opts *o) {
	using Sh = DmmaShape<S>;
	using Cf = MsgCfg<S, VAR>;
	const int C = ctx->C, P = ctx->P;
	Bufs b = dmma_msg_bufs(ctx);
	// rescaling (SingleTreeLikelihood_scalePartials, treelikelihood.c:1790-1836, on the message form): the kernels leave the largest
	// entry of L_n per (op of the level, category, pattern), k_dmma_scale_from_max decides per pattern and divides the MESSAGE --
	// P_n (L_n / m) = (P_n L_n) / m -- with the cumulative factors in sf as on the node-at-a-time path
	const int lbeg = ctx->h_lower_level_off[level], lend = ctx->h_lower_level_off[level + 1];
	const bool scale = o->scale != 0;
	if (scale) {
		int rc;
		if ((rc = dmma_rowmax_reserve_level(ctx, (size_t)(lend - lbeg) * C * P * sizeof(double)))) return rc;
	}
	// the rescaling instantiation exists for the shipped geometry only (dmma_evaluate keeps rescaled evaluations off the tuning variants)
	constexpr bool RS = VAR == 0;
	auto lower = (RS && o->scale) ? k_dmma_lower_msg<S, Cf::MT, Cf::NSPLIT, Cf::WM, Cf::LNST, Cf::LGB, RS> : k_dmma_lower_msg<S, Cf::MT, Cf::NSPLIT, Cf::WM, Cf::LNST, Cf::LGB, false>;
	if (o->scale && !RS) return -1;
	const size_t ring = (size_t)Cf::WM * AStage<Sh, Cf::MT, 2, Cf::LNST>::NSTAGE * AStage<Sh, Cf::MT, 2, Cf::LNST>::STG;
	const int lthreads = 32 * Cf::WM * Cf::NSPLIT, ltiles = (P + Cf::WM * Cf::MT * 8 - 1) / (Cf::WM * Cf::MT * 8);
	PHBC_CHECK(cudaFuncSetAttribute(lower, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(128 + (3 * Sh::IMG + ring) * sizeof(double))));
	const bool split = true;  // per-kind launches (measured faster than one five-slot variant per level, round 1)
	for (int kind = 0; kind < 3; kind++) {
		int beg = ctx->h_lower_kind_off[4 * level + kind], end = ctx->h_lower_kind_off[4 * level + kind + 1];
		int nimg = 1 + kind;
		if (!split) {
			if (kind) break;
			end = ctx->h_lower_kind_off[4 * level + 3];
			nimg = 3;
		}
		if (end <= beg) continue;
		const size_t lsmem = 128 + (nimg * Sh::IMG + ring) * sizeof(double);
		int lper_sm = 1;
		PHBC_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lper_sm, lower, lthreads, lsmem));
		if (lper_sm < 1) lper_sm = 1;
		// cherries through the table of state pairs
		constexpr int S1 = S + 1, PT = S1 * S1;
		if (kind == 2 && split && dmma_cherry_tables_on<S>(ctx)) {
			int zmax = 1, rc;
			if ((rc = dmma_cherry_reserve(ctx, (size_t)C * PT * S * sizeof(double), end - beg, &zmax))) return rc;
			if (scale && (size_t)zmax * C * PT * sizeof(double) > ctx->cherry_pairmax_bytes) {
				PHBC_CHECK(cudaStreamSynchronize(ctx->stream));
				if (ctx->d_cherry_pairmax) cudaFree(ctx->d_cherry_pairmax);
				ctx->d_cherry_pairmax = NULL, ctx->cherry_pairmax_bytes = 0;
				PHBC_CHECK(cudaMalloc((void **)&ctx->d_cherry_pairmax, (size_t)zmax * C * PT * sizeof(double)));
				ctx->cherry_pairmax_bytes = (size_t)zmax * C * PT * sizeof(double);
			}
			if (!ctx->d_cherry_enum) {
				PHBC_CHECK(cudaMalloc((void **)&ctx->d_cherry_enum, 2 * PT));
				k_dmma_cherry_enum<<<(PT + 255) / 256, 256, 0, ctx->stream>>>(S1, ctx->d_cherry_enum);
				ctx->launches++;
			}
			Bufs bt = b;  // the pairs as a pattern set of their own: two enumerated tips, the table in place of the lower buffers
			bt.P = PT, bt.tip_states = ctx->d_cherry_enum, bt.tip_kind = PHBC_TIP_STATES, bt.lower = ctx->d_cherry_tab;
			const int ttiles = (PT + Cf::WM * Cf::MT * 8 - 1) / (Cf::WM * Cf::MT * 8);
			for (int z0 = beg; z0 < end; z0 += zmax) {
				const int zc = end - z0 < zmax ? end - z0 : zmax;
				k_dmma_cherry_ops<<<(zc + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_lower_ops + z0, zc, ctx->d_cherry_ops);
				lower<<<dim3(pick_chunks(lper_sm * ctx->num_sms, C * zc, ttiles), C, zc), lthreads, lsmem, ctx->stream>>>(bt, ctx->d_cherry_ops, ctx->d_dmma_img, nimg,
				                                                                                                          scale ? ctx->d_cherry_pairmax : nullptr);
				const size_t gtiles = ((size_t)P * S + 8 * 256 - 1) / (8 * 256);
				const int gx = pick_chunks(8 * ctx->num_sms, C * zc, gtiles > 65535 ? 65535 : (int)gtiles);
				if (scale)
					k_dmma_cherry_gather<S, true><<<dim3((unsigned)gx, C, zc), 256, 0, ctx->stream>>>(b, ctx->d_lower_ops + z0, ctx->d_cherry_tab, ctx->d_cherry_pairmax,
					                                                                                  ctx->d_rowmax + (size_t)(z0 - lbeg) * C * P);
				else
					k_dmma_cherry_gather<S, false><<<dim3((unsigned)gx, C, zc), 256, 0, ctx->stream>>>(b, ctx->d_lower_ops + z0, ctx->d_cherry_tab, nullptr, nullptr);
				ctx->launches += 3;
			}
			continue;
		}
		for (int z0 = beg; z0 < end; z0 += 65535) {
			const int zc = end - z0 < 65535 ? end - z0 : 65535;
			lower<<<dim3(pick_chunks(lper_sm * ctx->num_sms, C * zc, ltiles), C, zc), lthreads, lsmem, ctx->stream>>>(
			    b, ctx->d_lower_ops + z0, ctx->d_dmma_img, nimg, scale ? ctx->d_rowmax + (size_t)(z0 - lbeg) * C * P : nullptr);
			ctx->launches++;
		}
	}
	PHBC_CHECK(cudaGetLastError());
	if (scale) return dmma_scale_from_max(ctx, b, ctx->d_lower_ops + lbeg, lend - lbeg, nullptr, 0, lend - lbeg, 1, o->scaling_threshold);
	return 0;
}

// the state counts the kernels are instantiated for
#define PHBC_DMMA_BY_STATES(S_, CALL) \
	switch (S_) {                     \
	case 20: return CALL(20);         \
	case 60: return CALL(60);         \
	case 61: return CALL(61);         \
	case 62: return CALL(62);         \
	case 63: return CALL(63);         \
	default: break;                   \
	}

int phbc_dmma_pack_images(phbc_ctx *ctx, bool adjoint, int include_root_freqs, int tip_images) {
#define CALL(S) dmma_pack<S>(ctx, adjoint, include_root_freqs, tip_images)
	PHBC_DMMA_BY_STATES(ctx->S, CALL)
#undef CALL
	return -1;
}
int phbc_dmma_pack(phbc_ctx *ctx) {
#define CALL(S) dmma_pack<S>(ctx)
	PHBC_DMMA_BY_STATES(ctx->S, CALL)
#undef CALL
	return -1;
}
int phbc_dmma_lower_ops(phbc_ctx *ctx, const phbc_op *d_ops, int cnt) {
#define CALL(S) dmma_lower_ops<S>(ctx, d_ops, cnt)
	PHBC_DMMA_BY_STATES(ctx->S, CALL)
#undef CALL
	return -1;
}

// message-form pre-order pass: per depth, one launch per kind of parent (0, 1, 2 tip children; the device op list is sorted that
// way).  Parents without tip children need two image slots (P_n, Q) and run as wider CTAs where that pays (DmmaConfig::UWM_II).
template <int S, int VAR>
static int dmma_upper_msg(phbc_ctx *ctx, const phbc_eval_opts *o, double *result) {
	using Sh = DmmaShape<S>;
	using Cf = MsgCfg<S, VAR>;
	const int C = ctx->C, P = ctx->P, N = ctx->N;
	int rc;
	Bufs b = dmma_msg_bufs(ctx);
	const bool split = true;  // per-kind launches (measured faster than one five-slot variant per level, round 1)
	typedef void (*upper_fn)(Bufs, const phbc_parent_op *, const double *, const double *, const double *, const double *, int, int, double *, int, int, double *);
	struct Variant {
		upper_fn fn;
		int wm, nslots, threads, tiles, slots;
		size_t smem;
	} var[3];
	for (int kind = 0; kind < 3; kind++) {
		Variant &v = var[kind];
		const bool wide = split && kind == 0 && Cf::UWM_II != Cf::UWM;
		constexpr bool RS = VAR == 0;  // see dmma_lower_msg_level
		if (o->scale && !RS) return -1;
		if (o->scale)
			v.fn = wide ? (upper_fn)k_dmma_upper_msg<S, Cf::UMT, Cf::UNSPLIT, Cf::UWM_II, Cf::UNST, Cf::UGB, RS> : (upper_fn)k_dmma_upper_msg<S, Cf::UMT, Cf::UNSPLIT, Cf::UWM, Cf::UNST, Cf::UGB, RS>;
		else
			v.fn = wide ? (upper_fn)k_dmma_upper_msg<S, Cf::UMT, Cf::UNSPLIT, Cf::UWM_II, Cf::UNST, Cf::UGB, false> : (upper_fn)k_dmma_upper_msg<S, Cf::UMT, Cf::UNSPLIT, Cf::UWM, Cf::UNST, Cf::UGB, false>;
		v.wm = wide ? Cf::UWM_II : Cf::UWM;
		v.nslots = (split && kind == 0) ? 2 : (S <= 32 ? 6 : 5);
		const int warps = v.wm * Cf::UNSPLIT;
		v.threads = 32 * warps;
		v.tiles = (P + v.wm * Cf::UMT * 8 - 1) / (v.wm * Cf::UMT * 8);
		v.smem = 128 + ((size_t)v.nslots * Sh::IMG + 2 * Sh::NP + 4 * warps + (size_t)v.wm * AStage<Sh, Cf::UMT, 3, Cf::UNST>::NSTAGE * AStage<Sh, Cf::UMT, 3, Cf::UNST>::STG) * sizeof(double);
	}
	for (int kind = 0; kind < 3; kind++) {
		Variant &v = var[kind];
		size_t mx = 0;
		for (int k2 = 0; k2 < 3; k2++)
			if (var[k2].fn == v.fn && var[k2].smem > mx) mx = var[k2].smem;
		PHBC_CHECK(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
		int per_sm = 1;
		PHBC_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v.fn, v.threads, v.smem));
		v.slots = (per_sm < 1 ? 1 : per_sm) * ctx->num_sms;
	}
	// group [beg, end) of (level, kind); without the split one group per level takes every op with the five-slot variant
	auto group = [&](int l, int kind, int &beg, int &end) {
		beg = ctx->h_parent_kind_off[4 * l + kind], end = ctx->h_parent_kind_off[4 * l + kind + 1];
		if (!split) {
			if (kind) end = beg;
			else end = ctx->h_parent_kind_off[4 * l + 3];
		}
	};
	// cherries (never the root: a tree of more than two taxa) by the table of state pairs
	const bool cherries = split && ctx->T > 2 && dmma_cherry_tables_on<S>(ctx);
	int pstride = 1;
	for (int l = 0; l < ctx->n_upper_levels; l++)
		for (int kind = 0; kind < 3; kind++) {
			int beg, end;
			group(l, kind, beg, end);
			for (int z0 = beg; z0 < end; z0 += 65535) {
				const int k = pick_chunks(var[kind].slots, C * (end - z0 < 65535 ? end - z0 : 65535), var[kind].tiles);
				if (k > pstride) pstride = k;
			}
			if (kind == 2 && cherries && end > beg) {  // k_dmma_cherry_upper: 8 resident CTAs of 8 warps per SM, 32 patterns a turn
				int zmax = 1;
				if ((rc = dmma_cherry_reserve(ctx, (size_t)3 * C * (S + 1) * (S + 1) * S * sizeof(double), end - beg, &zmax))) return rc;
				for (int z0 = beg; z0 < end; z0 += zmax) {
					const int k = pick_chunks(8 * ctx->num_sms, C * (end - z0 < zmax ? end - z0 : zmax), (P + 31) / 32);
					if (k > pstride) pstride = k;
				}
			}
		}
	if ((rc = phbc_ensure_scratch(ctx, (size_t)N * C * pstride * sizeof(double)))) return rc;
	PHBC_CHECK(cudaMemsetAsync(ctx->d_scratch, 0, (size_t)N * C * pstride * sizeof(double), ctx->stream));
	// rescaling: the upper partials a level stores are rescaled like the lower ones (slots 2 z, 2 z + 1 of the level's sorted parent
	// ops, d_upper_slot); tip children and the children of table cherries store nothing and report 0
	const bool scale = o->scale != 0;
	for (int l = 0; l < ctx->n_upper_levels; l++) {
		const int pbeg = ctx->h_parent_level_off[l], pend = ctx->h_parent_level_off[l + 1];
		const size_t rm_level = (size_t)2 * (pend - pbeg) * C * Cf::UNSPLIT * P;
		if (scale && pend > pbeg) {
			if ((rc = dmma_rowmax_reserve_level(ctx, rm_level * sizeof(double)))) return rc;
			PHBC_CHECK(cudaMemsetAsync(ctx->d_rowmax, 0, rm_level * sizeof(double), ctx->stream));
		}
		for (int kind = 0; kind < 3; kind++) {
			int beg, end;
			group(l, kind, beg, end);
			const Variant &v = var[kind];
			if (kind == 2 && cherries && end > beg) {
				constexpr int PT = (S + 1) * (S + 1);
				int zmax = 1;
				if ((rc = dmma_cherry_reserve(ctx, (size_t)3 * C * PT * S * sizeof(double), end - beg, &zmax))) return rc;
				for (int z0 = beg; z0 < end; z0 += zmax) {
					const int zc = end - z0 < zmax ? end - z0 : zmax;
					k_dmma_cherry_upper_tables<S><<<dim3((PT + 1) / 2, C, zc), 192, 0, ctx->stream>>>(C, ctx->d_parent_ops + z0, ctx->d_P, ctx->d_dP, ctx->d_freqs,
					                                                                                      o->include_root_freqs, ctx->d_cherry_tab);
					k_dmma_cherry_upper<S><<<dim3(pick_chunks(8 * ctx->num_sms, C * zc, (P + 31) / 32), C, zc), 256, 0, ctx->stream>>>(
					    b, ctx->d_parent_ops + z0, ctx->d_cherry_tab, ctx->d_weights, ctx->d_pattern_lnl, pstride, ctx->d_scratch, scale ? 1 : 0);
					ctx->launches += 2;
				}
				continue;
			}
			for (int z0 = beg; z0 < end; z0 += 65535) {
				const int zc = end - z0 < 65535 ? end - z0 : 65535;
				v.fn<<<dim3(pick_chunks(v.slots, C * zc, v.tiles), C, zc), v.threads, v.smem, ctx->stream>>>(
				    b, ctx->d_parent_ops + z0, ctx->d_dmma_img, ctx->d_freqs, ctx->d_weights, ctx->d_pattern_lnl, o->include_root_freqs, pstride, ctx->d_scratch,
				    v.nslots, scale ? 1 : 0, scale ? ctx->d_rowmax + (size_t)2 * (z0 - pbeg) * C * Cf::UNSPLIT * P : nullptr);
				ctx->launches++;
			}
		}
		const int ubeg = ctx->h_upper_level_off[l], ucnt = ctx->h_upper_level_off[l + 1] - ubeg;
		if (scale && ucnt > 0 && pend > pbeg &&
		    (rc = dmma_scale_from_max(ctx, b, ctx->d_upper_ops + ubeg, ucnt, ctx->d_upper_slot, 0, 2 * (pend - pbeg), Cf::UNSPLIT, o->scaling_threshold)))
			return rc;
	}
	PHBC_CHECK(cudaGetLastError());
	return phbc_gradient_from_partials(ctx, pstride, result);
}

// lower and upper passes of the message form in tuning variant VAR
// phases: PHBC_PH_FORWARD = post-order pass + root integration, PHBC_PH_GRADIENT = pre-order pass with the branch reductions (run alone
// by phbc_dmma_matrix_gradient, once per matrix set, on the messages the forward phase left)
#define PHBC_PH_FORWARD 1
#define PHBC_PH_GRADIENT 2
template <int S, int VAR>
static int dmma_msg_passes(phbc_ctx *ctx, const phbc_eval_opts *o, double *result, int phases) {
	int rc;
	if (phases & PHBC_PH_FORWARD) {
		for (int l = 0; l < ctx->n_lower_levels; l++) {
			if (ctx->h_lower_level_off[l + 1] - ctx->h_lower_level_off[l] <= 0) continue;
			if ((rc = dmma_lower_msg_level<S, VAR>(ctx, l, o))) return rc;
		}
		if ((rc = phbc_generic_root(ctx, o, result))) return rc;
	}
	if ((phases & PHBC_PH_GRADIENT) && o->want_gradient && (rc = dmma_upper_msg<S, VAR>(ctx, o, result))) return rc;
	return 0;
}
template <int S>
static int dmma_msg_dispatch(phbc_ctx *ctx, const phbc_eval_opts *o, double *result, int phases) {
	return dmma_msg_passes<S, 0>(ctx, o, result, phases);  // the shipped geometry; tuning variants exist for 20 and 61 states
}
template <>
int dmma_msg_dispatch<20>(phbc_ctx *ctx, const phbc_eval_opts *o, double *result, int phases) {
	switch (ctx->tune) {
	case 1: return dmma_msg_passes<20, 1>(ctx, o, result, phases);
	case 2: return dmma_msg_passes<20, 2>(ctx, o, result, phases);
	case 3: return dmma_msg_passes<20, 3>(ctx, o, result, phases);
	case 4: return dmma_msg_passes<20, 4>(ctx, o, result, phases);
	case 5: return dmma_msg_passes<20, 5>(ctx, o, result, phases);
	case 6: return dmma_msg_passes<20, 6>(ctx, o, result, phases);
	default: return dmma_msg_passes<20, 0>(ctx, o, result, phases);
	}
}
template <>
int dmma_msg_dispatch<61>(phbc_ctx *ctx, const phbc_eval_opts *o, double *result, int phases) {
	switch (ctx->tune) {
	case 1: return dmma_msg_passes<61, 1>(ctx, o, result, phases);
	case 2: return dmma_msg_passes<61, 2>(ctx, o, result, phases);
	case 3: return dmma_msg_passes<61, 3>(ctx, o, result, phases);
	case 4: return dmma_msg_passes<61, 4>(ctx, o, result, phases);
	default: return dmma_msg_passes<61, 0>(ctx, o, result, phases);
	}
}

template <int S>
static int dmma_evaluate(phbc_ctx *ctx, const phbc_eval_opts *o, int phases = PHBC_PH_FORWARD | PHBC_PH_GRADIENT) {
	using Sh = DmmaShape<S>;
	using Cf = DmmaConfig<S>;
	const int C = ctx->C, P = ctx->P, N = ctx->N;
	const bool fwd = phases & PHBC_PH_FORWARD;  // a gradient-only call keeps the transition matrices as they are (d_dP is the caller's)
	int rc;
	if (phbc_dwalk_usable(ctx, o)) {
		// whole-tree walk (phb_dwalk.cu): messages in the lower buffers as in the message form below, upper partials never materialised
		phbc_eval_opts e = *o;
		e.want_gradient = 0;  // no upper buffers
		if (fwd && (rc = phbc_generic_prepare(ctx, &e))) return rc;
		ctx->lower_is_message = true;
		if (fwd) ctx->node_evals++;
		if ((rc = phbc_time_begin(ctx))) return rc;
		if ((rc = phbc_dwalk_passes(ctx, o, ctx->d_result + (size_t)o->batch_index * (1 + ctx->N), phases))) return rc;
		if ((rc = phbc_time_end(ctx))) return rc;
		PHBC_CHECK(cudaGetLastError());
		return 0;
	}
	if ((rc = fwd ? phbc_generic_prepare(ctx, o) : phbc_generic_buffers(ctx, o))) return rc;
	Bufs b = phbc_make_bufs(ctx);
	// message form: the fast path (unscaled, state tips, eigen system, upper partials not needed as such afterwards)
	// rescaled evaluations too (round 2, last step), unless the reference-compatible per-category normalisation is asked for, the
	// categories do not tile a warp (k_dmma_scale_from_max) or PHB_OPT_TUNE 22 keeps them on the node-at-a-time form for comparison
	bool msg = !o->materialize_uppers && ctx->have_eigen && !o->explicit_matrices &&
	           (!o->scale || (!o->compat_scaled_gradient && dmma_rowmax_usable(ctx) && ctx->d_upper_slot && ctx->tune != 22 && !(ctx->tune >= 1 && ctx->tune <= 6)));
	if (msg) {
		int usable = 0;
		if ((rc = dmma_tips_as_states(ctx, &usable))) return rc;
		msg = usable != 0;
	}
	if ((rc = dmma_pack<S>(ctx, msg, o->include_root_freqs, msg ? 1 : -1))) return rc;
	ctx->lower_is_message = msg;
	if (fwd) ctx->node_evals++;
	if ((rc = phbc_time_begin(ctx))) return rc;
	double *result = ctx->d_result + (size_t)o->batch_index * (1 + N);
	if (msg) {
		if ((rc = dmma_msg_dispatch<S>(ctx, o, result, phases))) return rc;
		if ((rc = phbc_time_end(ctx))) return rc;
		PHBC_CHECK(cudaGetLastError());
		return 0;
	}
	for (int l = 0; fwd && l < ctx->n_lower_levels; l++) {
		const int beg = ctx->h_lower_level_off[l], cnt = ctx->h_lower_level_off[l + 1] - beg;
		if (cnt <= 0) continue;
		if ((rc = dmma_lower_ops<S>(ctx, ctx->d_lower_ops + beg, cnt, o->scale != 0, o->scaling_threshold))) return rc;
	}
	if (fwd && (rc = phbc_generic_root(ctx, o, result))) return rc;
	if ((phases & PHBC_PH_GRADIENT) && o->want_gradient) {
		// fused reductions skip the tips' uppers; under rescaling they give the exact gradient, the reference-compatible per-category
		// normalisation (PHB_OPT_COMPAT_SCALED_GRADIENT) needs every category's denominator and stays with the generic K9 / K10 kernel
		const bool grad = !o->materialize_uppers && !(o->scale && o->compat_scaled_gradient);
		auto upper = grad ? k_dmma_upper<S, Cf::UMT, Cf::UNSPLIT, Cf::UWM, true> : k_dmma_upper<S, Cf::UMT, Cf::UNSPLIT, Cf::UWM, false>;
		const int uwarps = Cf::UWM * Cf::UNSPLIT;
		const size_t usmem = 128 + ((grad ? 5 : 3) * Sh::IMG + 2 * Sh::NP + 2 * uwarps + Cf::UWM * AStage<Sh, Cf::UMT, 3>::NSTAGE * AStage<Sh, Cf::UMT, 3>::STG) * sizeof(double);
		PHBC_CHECK(cudaFuncSetAttribute(upper, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)usmem));
		const int uthreads = 32 * uwarps, utiles = (P + Cf::UWM * Cf::UMT * 8 - 1) / (Cf::UWM * Cf::UMT * 8);
		int uper_sm = 1;
		PHBC_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&uper_sm, upper, uthreads, usmem));
		if (uper_sm < 1) uper_sm = 1;
		const int uslots = uper_sm * ctx->num_sms;
		// per-CTA gradient partials are laid out [N][C][pstride]; levels launch different chunk counts, unused entries stay zero
		int pstride = 1;
		for (int l = 0; l < ctx->n_upper_levels; l++) {
			const int cnt = ctx->h_parent_level_off[l + 1] - ctx->h_parent_level_off[l];
			for (int z0 = 0; z0 < cnt; z0 += 65535) {
				const int k = pick_chunks(uslots, C * (cnt - z0 < 65535 ? cnt - z0 : 65535), utiles);
				if (k > pstride) pstride = k;
			}
		}
		if ((rc = phbc_ensure_scratch(ctx, (size_t)N * C * pstride * sizeof(double)))) return rc;
		if (grad) PHBC_CHECK(cudaMemsetAsync(ctx->d_scratch, 0, (size_t)N * C * pstride * sizeof(double), ctx->stream));
		const bool fused = o->scale && dmma_rowmax_usable(ctx) && ctx->d_upper_slot;
		for (int l = 0; l < ctx->n_upper_levels; l++) {
			const int beg = ctx->h_parent_level_off[l], cnt = ctx->h_parent_level_off[l + 1] - beg;
			// the children written by this level are the per-child ops of depth l + 1
			const int ubeg = ctx->h_upper_level_off[l], ucnt = ctx->h_upper_level_off[l + 1] - ubeg;
			int zmax = 65535;
			if (fused && cnt > 0 && (rc = dmma_rowmax_reserve(ctx, Cf::UNSPLIT, 2, cnt, &zmax))) return rc;
			for (int z0 = 0; z0 < cnt; z0 += zmax) {
				const int zc = cnt - z0 < zmax ? cnt - z0 : zmax;
				upper<<<dim3(pick_chunks(uslots, C * zc, utiles), C, zc), uthreads, usmem, ctx->stream>>>(
				    b, ctx->d_parent_ops + beg + z0, ctx->d_dmma_img, ctx->d_freqs, ctx->d_weights, ctx->d_pattern_lnl, o->include_root_freqs, pstride,
				    ctx->d_scratch, o->scale ? 1 : 0, fused ? ctx->d_rowmax : nullptr);
				ctx->launches++;
				if (fused && ucnt > 0 &&
				    (rc = dmma_scale_from_max(ctx, b, ctx->d_upper_ops + ubeg, ucnt, ctx->d_upper_slot, 2 * z0, 2 * zc, Cf::UNSPLIT, o->scaling_threshold)))
					return rc;
			}
			if (o->scale && !fused && ucnt > 0 && (rc = phbc_generic_scale_ops(ctx, ctx->d_upper_ops + ubeg, ucnt, o->scaling_threshold))) return rc;
		}
		if (grad) {
			if ((rc = phbc_gradient_from_partials(ctx, pstride, result))) return rc;
		} else {
			if ((rc = phbc_generic_gradient(ctx, o, result))) return rc;
		}
	}
	if ((rc = phbc_time_end(ctx))) return rc;
	PHBC_CHECK(cudaGetLastError());
	return 0;
}

// the packed images k_dmma_pack wrote last, brought back to [N][C][S][S]: what the tensor-core kernels actually stage
template <int S>
static int dmma_download(phbc_ctx *ctx, double *P, double *dP) {
	using Sh = DmmaShape<S>;
	const int C = ctx->C, N = ctx->N, T = ctx->T;
	if (!ctx->d_dmma_img) return -4;
	const size_t n = (size_t)2 * N * C * Sh::IMG;
	double *img = (double *)malloc(n * sizeof(double)), *f = (double *)malloc(S * sizeof(double));
	if (!img || !f) {
		free(img), free(f);
		return -3;
	}
	cudaError_t e = cudaMemcpyAsync(img, ctx->d_dmma_img, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(f, ctx->d_freqs, S * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	if (e == cudaSuccess)
		for (int which = 0; which < 2; which++) {
			double *dst = which ? dP : P;
			if (!dst) continue;
			for (int nd = 0; nd < N; nd++)
				for (int c = 0; c < C; c++) {
					const double *src = img + (((size_t)which * N + nd) * C + c) * Sh::IMG;
					double *M = dst + ((size_t)nd * C + c) * S * S;
					const bool tip = nd < T && ctx->dmma_pack_tips, adj = which == 1 && ctx->dmma_pack_adjoint && nd >= T;
					for (int i = 0; i < S; i++)
						for (int j = 0; j < S; j++) {
							if (nd == ctx->root) M[i * S + j] = src[i * Sh::LD + j];
							else if (tip) M[i * S + j] = src[j * Sh::NP + i] / ((which == 1 && ctx->dmma_pack_tips == 2 && !ctx->dmma_pack_irf) ? f[i] : 1.0);
							else if (adj) M[i * S + j] = src[j * Sh::LD + i] / (ctx->dmma_pack_irf ? 1.0 : f[i]);
							else M[i * S + j] = src[i * Sh::LD + j];
						}
				}
		}
	free(img), free(f);
	PHBC_CHECK(e);
	return 0;
}

int phbc_dmma_download_matrices(phbc_ctx *ctx, double *P, double *dP) {
	PHBC_CHECK(cudaSetDevice(ctx->device));
#define CALL(S) dmma_download<S>(ctx, P, dP)
	PHBC_DMMA_BY_STATES(ctx->S, CALL)
#undef CALL
	return -1;
}

// Node terms of calculate_dlnl_dQ (treelikelihood.c:2337-2583) for nsets per-node matrix sets on the tensor cores: the forward phase
// once, then ONE gradient phase per set with the set standing in for dP/dt -- the reductions that give d lnL / d t for dP/dt give
// sum_p w_p / L_p sum_i f_i U_n[i] (M_k L_n)[i] for M_k -- where the node-at-a-time sweep needs every upper partial in HBM and S^2
// scalar operations per (pattern, node, set).  d_cat [nsets][N][C]: the per-(node, category) sums, before any collapse.
int phbc_dmma_matrix_gradient(phbc_ctx *ctx, const phbc_eval_opts *o, int nsets, const double *d_M, double *d_cat) {
	phbc_eval_opts e = *o;
	e.want_gradient = 1;
	e.materialize_uppers = 0;
	int rc = 0;
	const size_t set = (size_t)ctx->N * ctx->C * ctx->S * ctx->S, nc = (size_t)ctx->N * ctx->C;
#define CALL(S) dmma_evaluate<S>(ctx, &e, PHBC_PH_FORWARD)
	switch (ctx->S) {
	case 20: rc = CALL(20); break;
	case 60: rc = CALL(60); break;
	case 61: rc = CALL(61); break;
	case 62: rc = CALL(62); break;
	case 63: rc = CALL(63); break;
	default: return -1;
	}
#undef CALL
	double *own_dP = ctx->d_dP;
	for (int k = 0; k < nsets && !rc; k++) {
		ctx->d_dP = const_cast<double *>(d_M) + (size_t)k * set;
#define CALL(S) dmma_evaluate<S>(ctx, &e, PHBC_PH_GRADIENT)
		switch (ctx->S) {
		case 20: rc = CALL(20); break;
		case 60: rc = CALL(60); break;
		case 61: rc = CALL(61); break;
		case 62: rc = CALL(62); break;
		default: rc = CALL(63); break;
		}
#undef CALL
		if (!rc && cudaMemcpyAsync(d_cat + (size_t)k * nc, ctx->d_cat_grad, nc * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) rc = -2;
	}
	ctx->d_dP = own_dP;
	return rc;
}

int phbc_dmma_evaluate(phbc_ctx *ctx, const phbc_eval_opts *o) {
#define CALL(S) dmma_evaluate<S>(ctx, o)
	PHBC_DMMA_BY_STATES(ctx->S, CALL)
#undef CALL
	snprintf(phbc_errbuf, sizeof(phbc_errbuf), "tensor-core kernels are instantiated for 20 and 60 ... 63 states, not %d", ctx->S);
	return -1;
}
