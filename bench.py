#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 tree-likelihood path (driver contract, task ④).

One "step" = one full lnL + branch-gradient evaluation (protocol of the reference's
examples/benchmarking.c:498-503: every node dirty, new branch lengths in, transition matrices
rebuilt, post-order + pre-order passes, lnL and grad[N] back on the host).

Workload at N=1: BASELINE.json configs[1] -- GTR+Γ4 nucleotide, synthetic 1000 taxa × 100k site
patterns.  Multi-GPU: patterns sharded across ranks (one process per GPU), one NCCL all-reduce of
[lnL, grad[N]] per step; weak scaling (100k patterns per GPU).

    python bench.py --gpus 1 --steps 20 --warmup 3
    torchrun --nproc-per-node N bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own CPU path (oracle/_ref) on host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from physher_b200 import models, synthetic as syn  # noqa: E402

CONFIGS = {
    # name: (taxa, patterns per GPU, states, categories, model)
    "c2": dict(taxa=1000, patterns=100_000, states=4, cats=4, model="GTR+G4", mu=0.04),
    "c2_1m": dict(taxa=1000, patterns=1_000_000, states=4, cats=4, model="GTR+G4", mu=0.04),
    "c3": dict(taxa=500, patterns=50_000, states=4, cats=4, model="HKY+G4", mu=0.04, batch=128),
    "c4": dict(taxa=200, patterns=200_000, states=20, cats=4, model="LG+G4", mu=0.08),
    "c5": dict(taxa=100, patterns=1_000_000, states=61, cats=1, model="GY94", mu=0.1),
}
# FP64 tensor-core throughput measured on this pool's B200 with tools/dmma_peak.cu (profiles/r1_c_dmma_peak.md); nominal 40
DMMA_PEAK_TFLOPS_FILE = os.path.join(ROOT, "profiles", "dmma_peak.json")
CODON_BASES = "TCAG"
CODON_AA = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
METRIC = "lnL+gradient throughput (site patterns x tree nodes per second)"
UNIT = "pattern*node/s"


def make_inputs(cfg: dict, rank: int, seed: int = 20261017):
    """Seeded synthetic inputs. Topology, model and branch lengths are identical on
    every rank; each rank draws its own pattern shard."""
    T, P, S, C = cfg["taxa"], cfg["patterns"], cfg["states"], cfg["cats"]
    topo = syn.random_topology(T, seed)
    bl = syn.random_branch_lengths(topo, seed + 1)
    if S == 4 and cfg["model"].startswith("HKY"):
        m = models.hky(3.0, [0.1, 0.2, 0.3, 0.4])
    elif S == 4:
        m = models.gtr([0.05, 0.3, 0.1, 0.15, 0.3, 0.1], [0.1, 0.2, 0.3, 0.4])
    elif S == 61:
        m = models.gy94(2.5, 0.3)
    elif S == 20:
        m = lg_model()
    else:
        m = models.random_reversible(S, seed + 2)
    rates, props = models.discrete_gamma(0.5, C)
    # data evolved down the same tree at 0.35x the evaluation branch lengths: all columns unique, per-pattern lnL
    # around -240 (min > -400), so neither arm ever switches rescaling on (SURVEY.md 8c caveat 2)
    patterns = syn.simulate_patterns(topo, bl * 0.35, P, S, seed + 100 + rank)
    weights = np.ones(P)
    return topo, bl, m, rates, props, patterns, weights


def inputs_sha256(topo, bl, m, rates, props, patterns=None, weights=None):
    """SHA-256 of the workload as both arms see it (recorded next to every result).  `tree_model` covers topology,
    branch lengths, eigen system, frequencies and the site model -- identical in both arms and on every rank; `patterns` covers the
    pattern shard of rank 0 (our arm; the reference arm times bounded shards of the same generator, see cpu_baseline.sample)."""
    import hashlib

    h = hashlib.sha256()
    for a in (topo.left, topo.right, [topo.root], bl, m.evec, m.eval, m.ivec, m.freqs, rates, props):
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    out = {"tree_model": h.hexdigest()}
    if patterns is not None:
        g = hashlib.sha256()
        g.update(np.ascontiguousarray(patterns, dtype=np.uint8).tobytes())
        g.update(np.ascontiguousarray(weights, dtype=np.float64).tobytes())
        out["patterns_rank0"] = g.hexdigest()
    return out


def lg_model():
    """LG eigen system as the reference's host code produced it (lg.c + eigen.c), stored in the committed golden fixture."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "synth_lg_g4_tipstates.npz"))
    return models.SubstitutionModel("LG", 20, z["freqs"], z["evec"], z["eval"], z["ivec"])


def algorithmic_flops(cfg: dict) -> float:
    """8 S^2 (T-2) flop per pattern per category (one S x S mat-vec per internal operand)."""
    T, P, S, C = cfg["taxa"], cfg["patterns"], cfg["states"], cfg["cats"]
    return 8.0 * S * S * (T - 2) * C * P


def dmma_peak():
    try:
        d = json.load(open(DMMA_PEAK_TFLOPS_FILE))
        return float(d["dmma_tflops"]), f"measured (tools/dmma_peak.cu, {d.get('shape', 'm8n8k4')}, profiles/dmma_peak.json)"
    except Exception:
        return 40.0, "nominal B200 FP64 tensor (no measured file)"


def algorithmic_bytes(cfg: dict) -> float:
    """Streaming-model bytes per evaluation: P * [(5T-9) * C*S*8 + 10T]."""
    T, P, S, C = cfg["taxa"], cfg["patterns"], cfg["states"], cfg["cats"]
    return P * ((5 * T - 9) * C * S * 8.0 + 10 * T)


class ClockSampler:
    """SM clock / throttle-reason sampling DURING the timed region (B200_PROFILING.md clocks line): NVML polled from a
    thread every few ms (the timed region of a short run is shorter than one `nvidia-smi -lms` period), with the
    nvidia-smi loop as the fallback when the NVML binding is missing."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int, period_s: float = 0.004):
        self.index, self.period = index, period_s
        self.proc, self.lines, self.samples, self.stop = None, [], [], threading.Event()
        self.nvml = None

    def _nvml_loop(self):
        n = self.nvml
        h = n.nvmlDeviceGetHandleByIndex(self.index)
        bits = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        while not self.stop.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((float(sm), float(mx), [k for k, b in bits.items() if r & b]))
            except Exception:
                pass
            self.stop.wait(self.period)

    def __enter__(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        self.stop.set()
        if self.nvml is not None:
            self.thread.join(timeout=2)
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        for s_, m_, r_ in self.samples:
            sm.append(s_)
            mx.append(m_)
            reasons.update(r_)
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml" if self.samples else "nvidia-smi"}


def measured_traffic(config_name, kernels):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu --set full
    capture of this workload (profiles/traffic.json, written by tools/ncu_summary.py); None when there is no capture."""
    if config_name is None:
        return None
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return float(d[f"{config_name}:{kernels}"]["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# -------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference (oracle/_ref) on host cores
# -------------------------------------------------------------------------------------------------

def _reference_worker(args):
    """One process = one single-threaded reference tree likelihood on a pattern shard."""
    cfg, shard_patterns, shard_index, iters, warm = args
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)  # the reference prints alignment statistics to stdout
    try:
        from oracle import oracle as O

        sub = dict(cfg, patterns=shard_patterns)
        topo, bl, m, rates, props, patterns, weights = make_inputs(sub, rank=1000 + shard_index)
        names = [f"t{i}" for i in range(sub["taxa"])]
        S = sub["states"]
        if S == 4:
            seqs = dict(zip(names, syn.sequences_from_patterns(patterns, syn.NUCLEOTIDES)))
            mspec = (O.nucleotide_model_spec("hky", [0.1, 0.2, 0.3, 0.4], kappa=3.0) if sub["model"].startswith("HKY") else
                     O.nucleotide_model_spec("gtr", [0.1, 0.2, 0.3, 0.4], [0.05, 0.3, 0.1, 0.15, 0.3, 0.1]))
            spec = O.treelikelihood_spec(syn.to_newick(topo, bl, names), seqs, mspec, categories=sub["cats"], alpha=0.5, tipstates=False)
            ref = O.Reference(spec)
        elif S == 20:
            seqs = dict(zip(names, syn.sequences_from_patterns(patterns, syn.AMINO_ACIDS)))
            lg = {"id": "sm", "type": "substitutionmodel", "model": "lg", "datatype": "aa",
                  "frequencies": {"id": "freqs", "type": "Simplex", "values": [float(x) for x in m.freqs]}}
            spec = O.treelikelihood_spec(syn.to_newick(topo, bl, names), seqs, lg, categories=sub["cats"], alpha=0.5, tipstates=False,
                                         datatype="aa")
            ref = O.Reference(spec)
        elif S == 61:
            codons = [a + b + c for a in CODON_BASES for b in CODON_BASES for c in CODON_BASES]
            sense = [c for c, a in zip(codons, CODON_AA) if a != "*"]
            seqs = {n: "".join(sense[s_] for s_ in row) for n, row in zip(names, patterns)}
            ref = O.Reference(codon=dict(newick=syn.to_newick(topo, bl, names), sequences=seqs, kappa=2.5, omega=0.3))
        else:
            raise NotImplementedError(f"reference arm has no model for {S} states")
        # gradient request as in BASELINE.md §4.3: tree model flag, include_root_freqs = false
        ref.time_gradient(warm, O.FLAG_TREE_MODEL, 0)
        sec = ref.time_gradient(iters, O.FLAG_TREE_MODEL, 0)
        npat = ref.P
        nodes = ref.N
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
    return sec, npat, nodes


def run_reference(cfg: dict, cores: int, sample_patterns: int, iters: int, warm: int = 1):
    """Throughput of the reference's SSE path: `cores` independent single-threaded processes (the path has no
    intra-likelihood threading) on disjoint pattern shards of a bounded sample."""
    import multiprocessing as mp

    per = max(64, sample_patterns // cores)
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_reference_worker, [(cfg, per, i, iters, warm) for i in range(cores)])
    wall = time.perf_counter() - t0
    # aggregate: every process evaluates its shard at its own rate
    pn_per_s = sum(npat * nodes / sec for sec, npat, nodes in res)
    return dict(value=pn_per_s, sec_per_eval=[r[0] for r in res], patterns=[r[1] for r in res], nodes=res[0][2], wall=wall, per=per)


def reference_main(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O

    if not O.reference_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_harness.so missing (built where /root/reference exists)"}))
        return 0
    cores = os.cpu_count() or 1
    cores = min(cores, 32)
    sample = sample_per_core(cfg) * cores
    iters = max(1, min(args.steps, 10))
    topo_, bl_, m_, rates_, props_, _, _ = make_inputs(dict(cfg, patterns=64), 0)
    t0 = time.perf_counter()
    r = run_reference(cfg, cores, sample, iters, warm=min(args.warmup, 1) or 1)
    total_patterns = sum(r["patterns"])
    ms = 1e3 * float(np.mean(r["sec_per_eval"]))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": iters, "warmup": 1,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg), "taxa": cfg["taxa"], "patterns_per_gpu": cfg["patterns"], "states": cfg["states"],
                   "categories": cfg["cats"], "model": cfg["model"], "inputs_sha256": inputs_sha256(topo_, bl_, m_, rates_, props_)},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{cores} single-threaded processes x {r['per']} patterns each ({total_patterns} unique patterns total) of the same "
                                   f"{cfg['taxa']}-taxon workload, {iters} lnL+gradient evaluations each, protocol examples/benchmarking.c:498-503"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "evals_per_s_at_workload": r["value"] / (cfg["patterns"] * r["nodes"]),
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))
    return 0


def sample_per_core(cfg) -> int:
    """Patterns per reference process: ~1 s of CPU work per evaluation and process at each state count, i.e. 10-30 core-seconds
    per core over the warm-up + timed evaluations (the reference allocates every node's partials: ~1.3 GB per process at C2)."""
    return {4: 5000, 20: 2400, 61: 800}.get(cfg["states"], 200)


def workload_name(cfg):
    batch = f" x {cfg['batch']} branch-length samples per step" if cfg.get("batch", 1) > 1 else ""
    return f"{cfg['model']} {cfg['taxa']} taxa x {cfg['patterns']} patterns per GPU{batch}, lnL + branch gradients"


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--patterns", type=int, default=0, help="override patterns per GPU")
    ap.add_argument("--kernels", default="auto", choices=["auto", "generic", "fused"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.patterns:
        cfg["patterns"] = args.patterns
    if args.impl == "reference":
        return reference_main(args, cfg)

    import torch
    import torch.distributed as dist

    import physher_b200 as phb
    from physher_b200.treelikelihood import OPT_KERNELS, OPT_TIMING

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the tree-likelihood path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W, K = max(args.warmup, 3), args.steps

    topo, bl, m, rates, props, patterns, weights = make_inputs(cfg, rank)
    sha = inputs_sha256(topo, bl, m, rates, props, patterns, weights) if rank == 0 else None
    T, P, S, C = cfg["taxa"], cfg["patterns"], cfg["states"], cfg["cats"]
    N = 2 * T - 1
    tlk = phb.SingleTreeLikelihood(topo.left, topo.right, topo.root, S, C, P, use_tip_states=True, device=local_rank)
    tlk.set_tip_states(patterns)
    tlk.set_pattern_weights(weights)
    tlk.set_eigen(m.evec, m.eval, m.ivec)
    tlk.set_frequencies(m.freqs)
    tlk.set_site_model(rates, props)
    tlk.set_option(OPT_KERNELS, {"auto": phb.KERNELS_AUTO, "generic": phb.KERNELS_GENERIC, "fused": phb.KERNELS_FUSED}[args.kernels])
    tlk.initialize_gradient(phb.FLAG_TREE_MODEL)
    ext = torch.cuda.ExternalStream(tlk.stream(), device=torch.device("cuda", local_rank))
    out_dev = torch.zeros(1 + N, dtype=torch.float64, device="cuda")
    rng = np.random.default_rng(7)

    def new_bl():
        # every step sees new branch lengths (host buffer), like an optimiser / VI iteration would produce
        b = bl * rng.uniform(0.98, 1.02, size=bl.shape)
        b[topo.root] = 0.0
        b[topo.right[topo.root]] = 0.0
        return b

    B = int(cfg.get("batch", 1))

    def step_batch():
        """BASELINE config 3: B branch-length samples (base x LogNormal(0, 0.1)) per step through phb_tlk_gradient_batch --
        host buffers in ([B][N] doubles) and out (lnl[B], grad[B][N]); one fused launch for the whole batch."""
        bls = bl[None, :] * rng.lognormal(0.0, 0.1, size=(B, N))
        bls[:, topo.root] = 0.0
        bls[:, topo.right[topo.root]] = 0.0
        lnls, grads = tlk.gradient_batch(bls)
        if world > 1:
            t = torch.from_numpy(np.concatenate([lnls[:, None], grads], axis=1)).cuda()
            dist.all_reduce(t)
            h = t.cpu().numpy()
            lnls, grads = h[:, 0], h[:, 1:]
        return float(lnls[-1]), grads[-1]

    def step_e2e():
        """Public API, host in / host out: H2D of the branch lengths, full evaluation, D2H of lnL + gradient."""
        if B > 1:
            return step_batch()
        tlk.set_branch_lengths(new_bl())
        if world == 1:
            g = tlk.gradient()
            return tlk.calculate(), g
        tlk.gradient_device(out_dev.data_ptr())
        tlk.synchronize()
        dist.all_reduce(out_dev)
        h = out_dev.cpu().numpy()
        g = h[1:].copy()
        g[topo.root] = 0.0
        g[topo.right[topo.root]] = 0.0
        return float(h[0]), g

    def step_device():
        """Device-resident step: inputs already in HBM, result left on the device."""
        if B > 1:  # the batched entry point takes host buffers (2 x B x N doubles per step, ~1 MB each way at C3)
            step_batch()
            return
        tlk.gradient_device(out_dev.data_ptr())
        if world > 1:
            tlk.synchronize()
            dist.all_reduce(out_dev)

    def sync_all():
        tlk.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    tlk.set_branch_lengths(bl)
    for _ in range(W):
        lnl, g = step_e2e()
    if not np.isfinite(lnl):
        raise SystemExit(f"non-finite lnL {lnl}")
    for _ in range(W):
        step_device()
    sync_all()

    # ---- timed region 1: device-resident throughput ("value"), CUDA events on the launching stream
    tlk.set_option(OPT_TIMING, 1)
    launches0 = tlk.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        sync_all()
        e0.record(ext)
        t0 = time.perf_counter()
        for _ in range(K):
            step_device()
        if world > 1:
            ext.wait_stream(torch.cuda.current_stream())
        e1.record(ext)
        sync_all()
        wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    launches = tlk.launch_count() - launches0
    kern_ms, kern_n = tlk.kernel_time()
    tlk.set_option(OPT_TIMING, 0)
    step_ms = max(dev_ms, 0.0) / K
    if world > 1:
        t = torch.tensor([step_ms, wall * 1e3 / K], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, wall_ms = float(t[0]), float(t[1])
        step_ms = max(step_ms, wall_ms) if step_ms <= 0 else step_ms
    # ---- timed region 2: end to end through the public API with host buffers
    sync_all()
    t0 = time.perf_counter()
    for _ in range(K):
        lnl, g = step_e2e()
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / K
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])

    total_pn = float(P) * world * N * B
    value = total_pn / (step_ms * 1e-3)
    e2e_value = total_pn / (e2e_ms * 1e-3)
    peak, peak_src = measured_peak_gbs()
    alg = algorithmic_bytes(cfg) * B  # one launch processes the whole batch
    kms = kern_ms / max(kern_n, 1)
    achieved = alg / (kms * 1e-3) / 1e9 if kms > 0 else None
    fused = args.kernels != "generic" and S == 4
    tensor = args.kernels != "generic" and S in (20, 61)
    traffic = measured_traffic(args.config if not args.patterns else None, args.kernels)
    hbm = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
           "traffic": traffic, "peak_source": peak_src, "kernel_ms": kms, "algorithmic_bytes_per_launch": alg}
    if traffic and kms > 0:
        hbm["traffic_GBs"] = traffic / (kms * 1e-3) / 1e9
        hbm["traffic_frac"] = hbm["traffic_GBs"] / peak
    if fused:
        fused_bytes = float(P) * B * (2 * (T - 1) * C * S * 8 + 2 * T)  # what the fused walk must move: lower rows out and back, tip codes twice
        hbm.update(kernel="k_nuc4_walk<C=%d,scale=0,grad=1>" % C, fused_min_bytes_per_launch=fused_bytes,
                   note="achieved = SURVEY.md 8d streaming-model bytes / kernel time; the fused walk keeps upper partials on chip, so it moves "
                        f"~{fused_bytes/1e9:.1f} GB per launch (traffic = ncu dram bytes) and frac can exceed 1; traffic_frac is the real DRAM utilisation")
        roof = hbm
    elif tensor:
        flops = algorithmic_flops(cfg)
        tpeak, tsrc = dmma_peak()
        tf = flops / (kms * 1e-3) / 1e12 if kms > 0 else None
        roof = {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": (tf / tpeak) if tf else None, "traffic": traffic,
                "peak_source": tsrc, "kernel": "k_dmma_lower_msg + k_dmma_upper_msg (FP64 mma.sync m8n8k4, message form, branch gradients in adjoint form: 3 dense products per internal node), all levels of one evaluation",
                "kernel_ms": kms, "algorithmic_flops_per_launch": flops, "hbm": hbm,
                "note": "launch = the level-batched kernel sequence of one evaluation; S=20 sits at the FP64 ridge so the HBM view is given too"}
    else:
        hbm.update(kernel="generic node-at-a-time kernels, all levels of one evaluation")
        roof = hbm
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg), "taxa": T, "patterns_per_gpu": P, "states": S, "categories": C, "model": cfg["model"],
                   "kernels": args.kernels, "l2": "per-evaluation working set (>= 2 GB of partials) exceeds the 126 MB L2; no explicit flush",
                   "sharding": f"patterns x{world}" if world > 1 else "single GPU", "samples_per_step": B, "inputs_sha256": sha},
        "evals_per_s": B * 1e3 / step_ms, "lnl": lnl,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": 8 * N * B, "d2h_bytes_per_step": 8 * (N + 1) * B,
                "evals_per_s": B * 1e3 / e2e_ms},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": roof,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import oracle as O

            if O.reference_available():
                cores = min(os.cpu_count() or 1, 32)
                r = run_reference(cfg, cores, sample_per_core(cfg) * cores, iters=3)
                line["cpu_baseline"] = {
                    "value": r["value"], "unit": UNIT, "cores": cores, "kind": "reference",
                    "sample": f"{cores} single-threaded reference processes x {r['per']} patterns each, 3 lnL+gradient evaluations each "
                              f"({cfg['model']}, {T} taxa, tip partials, SSE on), {r['wall']:.1f} s wall",
                    "one_core_value": float(np.mean([p * r['nodes'] / s for s, p in zip(r['sec_per_eval'], r['patterns'])])),
                }
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
        except Exception as exc:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {exc!r}"}
    if rank == 0:
        print(json.dumps(line))
    tlk.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
