"""Pattern sharding: host logic on CPU with world_size-2 gloo, and the real path on GPUs.

CPU tests inject the oracle as the shard evaluator: what is under test is physher_b200/sharded.py --
the shard ranges, the single SUM all-reduce of [lnL, grad[N]], and the policy applied after the
reduction (rescaling switch on -inf, NaN fill, unrooted convention).
"""
import os
import socket

import numpy as np
import pytest

from oracle import oracle as O
from physher_b200 import models, sharded, synthetic as syn
from tests.util import RTOL, grad_err, rel_err


def _problem(T=14, P=203, C=4, seed=31, deep=False):
    topo = syn.caterpillar_topology(T) if deep else syn.random_topology(T, seed)
    m = models.gtr([0.05, 0.3, 0.1, 0.15, 0.3, 0.1], [0.1, 0.2, 0.3, 0.4])
    rates, props = models.discrete_gamma(0.5, C)
    pb = O.Problem(left=topo.left, right=topo.right, parent=topo.parent, root=topo.root, nstate=4,
                   tip_states=syn.random_patterns(T, P, 4, 0.75 if deep else 0.15, seed + 1, unknown_frac=0.0 if deep else 0.02),
                   weights=np.random.default_rng(seed + 2).integers(1, 5, P).astype(np.float64),
                   freqs=m.freqs, rates=rates, props=props,
                   bl=syn.random_branch_lengths(topo, seed + 3, 0.8, 1.6) if deep else syn.random_branch_lengths(topo, seed + 3),
                   evec=m.evec, eval=m.eval, ivec=m.ivec)
    return pb


def test_shard_ranges_partition_the_patterns():
    for P in (1, 2, 7, 100, 100_000, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            edges = [sharded.shard_range(P, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == P
            for (b0, e0), (b1, e1) in zip(edges, edges[1:]):
                assert e0 == b1 and b0 <= e0
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharded.shard_range(10, 2, 2)


def test_shard_inputs_slices_are_contiguous_copies():
    pb = _problem()
    a, w = sharded.shard_inputs(pb.tip_states, pb.weights, 1, 2)
    b, e = sharded.shard_range(pb.npatterns, 1, 2)
    assert a.flags.c_contiguous and a.shape == (pb.ntips, e - b) and np.array_equal(a, pb.tip_states[:, b:e])
    assert np.array_equal(w, pb.weights[b:e])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pb = _problem(**case["problem"])
        full_states, full_w = pb.tip_states, pb.weights
        pb.tip_states, pb.weights = sharded.shard_inputs(full_states, full_w, rank, world)
        calls = []

        def evaluate_shard(bl, rescaling):
            # the oracle stands in for phb_tlk_gradient_device: raw shard sums, no unrooted convention
            calls.append(bool(rescaling))
            pb.scale = bool(rescaling)
            pb.unrooted = False
            if bl is not None:
                pb.bl = np.asarray(bl, dtype=np.float64)
            out = O.evaluate(pb)
            if case.get("poison_rank") == rank:
                out["lnl"] = float("nan")
            return torch.from_numpy(np.concatenate([[out["lnl"]], out["grad"]]))

        def evaluate_matrix_shard(M, rescaling):
            # stands in for phb_tlk_matrix_gradient + the cached lnL of the shard
            calls.append(bool(rescaling))
            pb.scale = bool(rescaling)
            pb.unrooted = True  # the sweep skips the root's right child itself (treelikelihood.c:2408)
            lnl_shard = O.evaluate(pb, gradient=False)["lnl"]
            return torch.from_numpy(np.concatenate([[lnl_shard], O.matrix_gradient(pb, M)]))

        st = sharded.ShardedTreeLikelihood(pb.nnodes, pb.root, int(pb.right[pb.root]), evaluate_shard=evaluate_shard,
                                           evaluate_matrix_shard=evaluate_matrix_shard)
        if case.get("matrix_sets"):
            M = np.random.default_rng(case["matrix_seed"]).normal(size=(case["matrix_sets"], pb.nnodes, pb.ncat, pb.nstate, pb.nstate))
            lnl, g = st.matrix_gradient(M)
        else:
            lnl, g = st.gradient(pb.bl)
        q.put((rank, lnl, g, calls, st.rescaling, st.evaluations))
    finally:
        dist.destroy_process_group()


def _run(case, world=2):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


@pytest.mark.timeout(300)
def test_two_rank_allreduce_matches_unsharded():
    case = {"problem": dict(T=14, P=203, C=4, seed=31)}
    res = _run(case)
    pb = _problem(**case["problem"])
    want = O.evaluate(pb)
    for rank, lnl, g, calls, rescaling, evals in res:
        assert rel_err(lnl, want["lnl"]) < RTOL
        assert grad_err(g, want["grad"]) < RTOL
        assert g[pb.root] == 0.0 and g[pb.right[pb.root]] == 0.0
        assert calls == [False] and not rescaling and evals == 1
    # both ranks hold bit-identical reduced results
    assert res[0][1] == res[1][1] and np.array_equal(res[0][2], res[1][2])


@pytest.mark.timeout(300)
def test_two_rank_underflow_switches_rescaling_on_everywhere():
    """A shard that underflows makes the REDUCED lnL -inf, so both ranks switch together (treelikelihood.c:1496-1519)."""
    case = {"problem": dict(T=600, P=24, C=1, seed=3, deep=True)}
    pb = _problem(**case["problem"])
    assert np.isinf(O.evaluate(pb, gradient=False)["lnl"])
    pb.scale = True
    want = O.evaluate(pb)
    res = _run(case)
    for rank, lnl, g, calls, rescaling, evals in res:
        assert calls == [False, True] and rescaling and evals == 2
        assert np.isfinite(lnl) and rel_err(lnl, want["lnl"]) < RTOL
        assert grad_err(g, want["grad"]) < 1e-9


@pytest.mark.timeout(300)
def test_two_rank_nan_fills_the_gradient_on_every_rank():
    case = {"problem": dict(T=10, P=64, C=2, seed=5), "poison_rank": 1}
    for rank, lnl, g, calls, rescaling, evals in _run(case):
        assert np.isnan(lnl) and np.isnan(g).all() and not rescaling  # treelikelihood.c:328-332


@pytest.mark.timeout(300)
def test_two_rank_matrix_gradient_matches_unsharded():
    """substitution-model parameter gradients are sums over patterns too: per-shard sweeps + one all-reduce of [lnL, out[nsets]]"""
    case = {"problem": dict(T=12, P=151, C=4, seed=41), "matrix_sets": 4, "matrix_seed": 42}
    res = _run(case)
    pb = _problem(**case["problem"])
    M = np.random.default_rng(42).normal(size=(4, pb.nnodes, pb.ncat, pb.nstate, pb.nstate))
    want, want_lnl = O.matrix_gradient(pb, M), O.evaluate(pb, gradient=False)["lnl"]
    for rank, lnl, out, calls, rescaling, evals in res:
        assert rel_err(lnl, want_lnl) < RTOL and grad_err(out, want) < RTOL
        assert calls == [False] and evals == 1
    assert np.array_equal(res[0][2], res[1][2])


# -------------------------------------------------------------------------------------------------
# the real thing: C-ABI shards on GPUs (single process drives both shards when one GPU is visible)
# -------------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_gpu_shards_sum_to_the_unsharded_result():
    import torch

    import physher_b200 as phb

    pb = _problem(T=40, P=1001, C=4, seed=77)
    want = O.evaluate(pb)
    full = phb.SingleTreeLikelihood.from_problem(pb)
    lnl_full, g_full = full.calculate(), full.gradient()
    full.close()
    world = 3
    acc = torch.zeros(1 + pb.nnodes, dtype=torch.float64, device="cuda")
    for rank in range(world):
        states, w = sharded.shard_inputs(pb.tip_states, pb.weights, rank, world)
        sub = O.Problem(**{**pb.__dict__, "tip_states": states, "weights": w})
        tlk = phb.SingleTreeLikelihood.from_problem(sub, device=rank % max(torch.cuda.device_count(), 1) if torch.cuda.device_count() >= world else 0)
        st = sharded.ShardedTreeLikelihood(pb.nnodes, pb.root, int(pb.right[pb.root]), tlk=tlk, device="cuda:0")
        acc += st.reduce_device(pb.bl).to("cuda:0")
        torch.cuda.synchronize()
        tlk.close()
    h = acc.cpu().numpy()
    g = h[1:].copy()
    g[pb.root] = 0.0
    g[pb.right[pb.root]] = 0.0
    assert rel_err(float(h[0]), want["lnl"]) < RTOL and rel_err(float(h[0]), lnl_full) < 1e-12
    assert grad_err(g, want["grad"]) < RTOL and grad_err(g, g_full) < 1e-11


@pytest.mark.gpu
def test_gpu_matrix_gradient_shards_sum_to_the_unsharded_result():
    import physher_b200 as phb

    pb = _problem(T=33, P=777, C=4, seed=91)
    M = np.random.default_rng(92).normal(size=(3, pb.nnodes, pb.ncat, 4, 4))
    want = O.matrix_gradient(pb, M)
    world, acc, lnl = 3, np.zeros(3), 0.0
    for rank in range(world):
        states, w = sharded.shard_inputs(pb.tip_states, pb.weights, rank, world)
        sub = O.Problem(**{**pb.__dict__, "tip_states": states, "weights": w})
        tlk = phb.SingleTreeLikelihood.from_problem(sub)
        st = sharded.ShardedTreeLikelihood(pb.nnodes, pb.root, int(pb.right[pb.root]), tlk=tlk, device="cuda:0")
        shard_lnl, out = st.matrix_gradient(M)  # no process group: the shard's own values
        acc += out
        lnl += shard_lnl
        tlk.close()
    assert grad_err(acc, want) < RTOL
    assert rel_err(lnl, O.evaluate(pb, gradient=False)["lnl"]) < RTOL
