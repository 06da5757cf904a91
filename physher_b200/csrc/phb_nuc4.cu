// phb_nuc4.cu -- fused whole-tree walk kernels for 4-state models (placeholder until the walk lands)
#include "phb_ctx.cuh"

bool phbc_nuc4_supported(const phbc_ctx *ctx, const phbc_eval_opts *o) { return false; }
int phbc_nuc4_evaluate(phbc_ctx *ctx, const phbc_eval_opts *o) { return -1; }
