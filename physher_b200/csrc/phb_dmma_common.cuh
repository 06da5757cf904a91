// phb_dmma_common.cuh -- shapes and the DMMA wrapper shared by the FP64 tensor-core translation units (phb_dmma.cu: level-batched
// kernels and the packed matrix images; phb_dwalk.cu: whole-tree walk).  Internal.
#pragma once

#include "phb_ctx.cuh"

template <int S_>
struct DmmaShape {
	static constexpr int S = S_;
	static constexpr int KP = (S + 3) / 4 * 4;  // padded contraction length
	static constexpr int NP = (S + 7) / 8 * 8;  // padded output states
	static constexpr int KT = KP / 4, NT = NP / 8;
	static constexpr int LD = (KP % 8 == 4) ? KP : KP + 4;  // leading dimension of a staged matrix, = 4 (mod 8)
	static constexpr int MAT = NP * LD;                     // doubles per staged matrix
	// the contraction runs in chunks of KCH k-steps; the A fragments of chunk i + 1 are fetched from HBM while the tensor
	// pipe works on chunk i (register double buffering).  Short contractions are one chunk: the prefetch then spans tiles.
	static constexpr int KCH = KT <= 6 ? KT : 4;
	static constexpr int NCH = (KT + KCH - 1) / KCH;
	// packed matrix image (one TMA bulk copy): [NP][LD] for partial operands, or TRANSPOSED [S][NP] + row sums [NP] for state tips
	static constexpr int TIP_IMG = S * NP + NP;
	static constexpr int IMG = ((MAT > TIP_IMG ? MAT : TIP_IMG) + 1) / 2 * 2;  // doubles, 16-byte multiple
};

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// the same without `volatile`: a pure function of its operands, so the compiler may schedule independent work between the DMMAs
__device__ __forceinline__ void dmma_m8n8k4_nv(double &d0, double &d1, double a, double b) {
	asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// packed matrix images of every (node, category) on the ctx stream (phb_dmma.cu, k_dmma_pack): `adjoint` stores the dP image of an
// internal node transposed and frequency-weighted, `tip_images` (1) gives tips the transposed [S + 1][NP] gather layout, (2) with the
// derivative images frequency-weighted as well
int phbc_dmma_pack_images(phbc_ctx *ctx, bool adjoint, int include_root_freqs, int tip_images);
