"""Pattern-sharded tree likelihood: one process per GPU, one all-reduce of [lnL, grad[N]] per evaluation.

Site patterns are independent units: rank g owns the contiguous pattern range
[P*g/G, P*(g+1)/G) with its slice of tip states and weights; topology, eigen system, category
rates and all transition matrices are replicated and rebuilt redundantly on every rank from the
same branch-length vector.  The only exchange is one SUM all-reduce of N+1 doubles.

What must happen AFTER the reduction (it needs the reduced lnL) lives here, mirroring the
single-GPU host code in csrc/phb_treelikelihood.c:
  * -inf (or +inf) lnL  => rescaling switched on on every rank and everything recomputed
    (treelikelihood.c:1496-1519).  The decision is global because a -inf shard makes the reduced
    lnL -inf on every rank, so all ranks take the same branch without a second collective;
  * NaN / inf lnL       => NaN-filled gradient (treelikelihood.c:328-332);
  * unrooted convention => root's right child gradient forced to 0 (treelikelihood.c:3249-3255).

Substitution-model parameter gradients shard the same way: the node sweep of calculate_dlnl_dQ
(treelikelihood.c:2337-2583) is a sum over patterns of w_k / L_k (...) with L_k local to the pattern, so
the per-shard values of `phb_tlk_matrix_gradient` add up -- one more SUM all-reduce of [lnL, out[nsets]].

The shard evaluator is the `phb_tlk_gradient_device` entry point of the C ABI; tests inject a
different evaluator to exercise this host logic with the gloo backend on CPU.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(npatterns: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous pattern range [begin, end) of `rank` (all patterns cost the same)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    return (npatterns * rank) // world, (npatterns * (rank + 1)) // world


def shard_inputs(tip_states_or_partials: np.ndarray, weights: np.ndarray, rank: int, world: int):
    """Slice [T][P](...) tip data and [P] weights to the rank's pattern range (copies, contiguous)."""
    b, e = shard_range(weights.shape[0], rank, world)
    return np.ascontiguousarray(tip_states_or_partials[:, b:e]), np.ascontiguousarray(weights[b:e])


class ShardedTreeLikelihood:
    """lnL + branch gradients of a pattern-sharded alignment.

    evaluate_shard(bl, rescaling) -> tensor [1+N] (float64, on `device`) holding this rank's raw partial sums
    [lnL_shard, grad_shard[0..N)].  With `tlk` given (a physher_b200.SingleTreeLikelihood built on this rank's
    shard) the evaluator is tlk.gradient_device on the tlk's own stream.
    """

    def __init__(self, nnodes: int, root: int, root_right: int, tlk=None, evaluate_shard: Optional[Callable] = None,
                 unrooted: bool = True, group=None, device=None, evaluate_matrix_shard: Optional[Callable] = None):
        import torch
        import torch.distributed as dist

        if (tlk is None) == (evaluate_shard is None):
            raise ValueError("give exactly one of tlk / evaluate_shard")
        self.torch, self.dist = torch, dist
        self.N, self.root, self.root_right = int(nnodes), int(root), int(root_right)
        self.unrooted = bool(unrooted)
        self.group = group
        self.tlk = tlk
        self.rescaling = bool(tlk.rescaling()) if tlk is not None else False
        self.evaluations = 0
        self._evaluate_matrix = evaluate_matrix_shard if tlk is None else self._evaluate_matrix_tlk
        if tlk is not None:
            self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
            self.out = torch.zeros(1 + self.N, dtype=torch.float64, device=self.device)
            self.stream = torch.cuda.ExternalStream(tlk.stream(), device=self.device)
            self._evaluate = self._evaluate_tlk
        else:
            self.device = torch.device("cpu") if device is None else torch.device(device)
            self.out = None
            self._evaluate = evaluate_shard

    # -- shard evaluators ---------------------------------------------------------------------
    def _evaluate_tlk(self, bl, rescaling):
        if bl is not None:
            self.tlk.set_branch_lengths(bl)
        if rescaling != self.tlk.rescaling():
            self.tlk.use_rescaling(rescaling)
        self.tlk.gradient_device(self.out.data_ptr())
        # the collective runs on torch's current stream: order it after the tlk's stream
        self.torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return self.out

    def _evaluate_matrix_tlk(self, M, rescaling):
        if rescaling != self.tlk.rescaling():
            self.tlk.use_rescaling(rescaling)
        out = self.tlk.matrix_gradient(M)  # host values of this shard (the C ABI returns them to the host)
        lnl = self.tlk.calculate()          # cached by the sweep
        return self.torch.from_numpy(np.concatenate([[lnl], out])).to(self.device)

    # -- reduction + post-reduction policy ------------------------------------------------------
    def reduce_device(self, bl=None):
        """One evaluation, result left on the device: reduced raw sums [lnL, grad[N]] (no policy applied)."""
        out = self._evaluate(bl, self.rescaling)
        if self.dist.is_available() and self.dist.is_initialized() and self.dist.get_world_size(self.group) > 1:
            self.dist.all_reduce(out, op=self.dist.ReduceOp.SUM, group=self.group)
        self.evaluations += 1
        return out

    def gradient(self, bl=None) -> Tuple[float, np.ndarray]:
        """(lnL, d lnL / d bl by node id) with the reference's NaN / inf / unrooted conventions."""
        h = self.reduce_device(bl).detach().cpu().numpy().copy()
        lnl = float(h[0])
        if math.isinf(lnl) and not self.rescaling:
            # every rank sees the same reduced lnL, so every rank switches (treelikelihood.c:1496-1519)
            self.rescaling = True
            h = self.reduce_device(None if self.tlk is not None else bl).detach().cpu().numpy().copy()
            lnl = float(h[0])
        g = h[1:]
        if math.isnan(lnl) or math.isinf(lnl):
            g[:] = np.nan  # treelikelihood.c:328-332
        else:
            g[self.root] = 0.0
            if self.unrooted:
                g[self.root_right] = 0.0  # treelikelihood.c:3249-3255
        return lnl, g

    def matrix_gradient(self, M) -> Tuple[float, np.ndarray]:
        """(lnL, out[nsets]) of phb_tlk_matrix_gradient over ALL shards for per-node matrix sets M [nsets][N][C][S][S] (the same on
        every rank): per-shard sweeps, one SUM all-reduce, then the reference's inf / NaN policy on the reduced lnL."""
        if self._evaluate_matrix is None:
            raise ValueError("no matrix-gradient evaluator (give tlk or evaluate_matrix_shard)")

        def once():
            t = self._evaluate_matrix(M, self.rescaling)
            if self.dist.is_available() and self.dist.is_initialized() and self.dist.get_world_size(self.group) > 1:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
            self.evaluations += 1
            return t.detach().cpu().numpy().copy()

        h = once()
        if math.isinf(float(h[0])) and not self.rescaling:
            self.rescaling = True  # every rank sees the same reduced lnL (treelikelihood.c:1496-1519)
            h = once()
        lnl, out = float(h[0]), h[1:]
        if math.isnan(lnl) or math.isinf(lnl):
            out[:] = np.nan
        return lnl, out
