#!/usr/bin/env python
"""bench.py — alignment columns/sec of the SubRecon hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # product arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU algorithm on host cores
    python bench.py --config 4                                # BASELINE configs 1..5 (default 3, the headline)
    python bench.py --config 5 --scaling strong --gpus 8      # fixed total size, contiguous column blocks per rank

A "step" is one pass of the hot path (P matrices -> clade tables -> pruning -> root-branch joint) over one batch of
synthetic alignment columns. Default = BASELINE config #3: 1,000 taxa x 1,000,000 sites, WAG, K=4, alpha=0.5, Yule tree
(seed 3), weak scaling (every rank owns a block of that size). Columns never interact, so there is no data-path
collective — only the barrier, the max-over-ranks of the time and a reduce of the total lnL.

  value     whole-job columns/s with the alignment resident in HBM and results (lnl + 400 probabilities) left in HBM
            (CUDA events on the launching stream)
  e2e       the same metric through the call a host makes, with HOST buffers: pinned states H2D and the per-column
            result D2H inside the timed region. The result is the COMPACT record at the reference's default -threshold
            0.5 — everything SiteResult keeps of a column (recon/SiteResult.java:55,69-83), 112 B — which is what the Java
            integration reads (java/subrecon/recon/NativeRecon.java). One process per GPU (torchrun).
  e2e_single_process   the same through sr_multi_run_streamed: ONE host process (rank 0) drives all N GPUs, the route a
            single JVM takes; the other ranks idle meanwhile
  e2e_full_joint       the same with all 400 probabilities per column coming back (3 208 B/column): bound by the host's
            D2H ingest on multi-GPU boxes; `host_d2h_ceiling` is that limit measured in the same run
  roofline  the pruning kernel's algorithmic FP64 flops (K*(820*E_int + 20*E_leaf + 1200) per column) over
            its CUDA-event time, against the FP64 tensor-pipe (DMMA) peak probed live on the same GPU
  cpu_baseline  the oracle's jar-faithful work mode (what PAL 1.51 really executes per column) on all host cores, on a
                bounded sample of the same workload
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

METRIC = "alignment_columns_per_sec"
UNIT = "columns/s"

# BASELINE.json configs. `base` = simulated columns that are tiled up to `columns`.
CONFIGS = {
    1: dict(label="example/ primate lysozyme (19 taxa x 130 sites), WAG + mle.wag.freqs, K=4", kind="example", ncat=4, model="wag"),
    2: dict(label="synthetic 100 taxa x 10k sites, JTT, K=4, balanced tree", taxa=100, columns=10_000, model="jtt", ncat=4,
            tree="balanced", tree_seed=2, aln_seed=102, base=10_000),
    3: dict(label="synthetic 1,000 taxa x 1M sites, WAG, K=4, Yule tree seed 3", taxa=1000, columns=1_000_000, model="wag", ncat=4,
            tree="yule", tree_seed=3, aln_seed=103, base=8192),
    4: dict(label="synthetic 5,000-taxon caterpillar x 200k sites, BLOSUM62, K=8, custom -pi", taxa=5000, columns=200_000,
            model="blosum62", ncat=8, tree="caterpillar", tree_seed=4, aln_seed=104, base=1024, custom_pi=True),
    5: dict(label="synthetic 2,000 taxa x 8M sites, Dayhoff, K=4, Yule tree seed 5", taxa=2000, columns=8_000_000, model="dayhoff",
            ncat=4, tree="yule", tree_seed=5, aln_seed=105, base=4096),
}
TREES = {"yule": "yule_tree", "balanced": "balanced_tree", "caterpillar": "caterpillar_tree"}
ORACLE_MODEL = {"wag": "WAG", "jtt": "JTT", "dayhoff": "Dayhoff", "blosum62": "BLOSUM62"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU gets the config's column count; strong: the config's columns are split over the GPUs")
    ap.add_argument("--columns", type=int, default=0, help="override the config's column count")
    ap.add_argument("--alpha", type=float, default=0.5)
    ap.add_argument("--chunk-cols", type=int, default=0, help="0 = library default")
    ap.add_argument("--tiled", action="store_true",
                    help="fill the alignment by repeating a block of `base` simulated columns instead of simulating every column")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-full-joint", action="store_true", help="skip the 3 208 B/column end-to-end variant")
    ap.add_argument("--no-single-process", action="store_true", help="skip the one-process-N-GPUs end-to-end variant")
    ap.add_argument("--single-process-devices", type=int, default=0,
                    help="without torchrun: let the one-process leg (sr_multi_*) drive this many GPUs, the config's columns on each "
                         "(no other rank's process shares the GPUs, as under a single JVM)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def config_of(a):
    c = dict(CONFIGS[a.config])
    if a.columns:
        c["columns"] = a.columns
    return c


def workload_config(a, c, columns_per_gpu, world):
    data = ("the reference's own example alignment" if c.get("kind") == "example" else
            f"columns simulated down the tree ({str(min(c['base'], c['columns'])) + ' distinct, tiled' if a.tiled else 'every column its own draw'})")
    big = columns_per_gpu * c.get("taxa", 19) > 2 ** 28
    return {"workload": f"BASELINE config #{a.config}: {c['label']}, alpha={a.alpha}; {data}",
            "baseline_config": a.config, "taxa": c.get("taxa", 19),
            "columns_total": columns_per_gpu * world if a.scaling == "weak" else c["columns"],
            "columns_per_gpu": columns_per_gpu, "model": c["model"], "rate_classes": c["ncat"], "alpha": a.alpha,
            "l2_policy": ("inputs larger than L2: the state matrix and the result arrays of one step are each far beyond the 126 MB L2"
                          if big else "correctness-sized config: its inputs fit in L2"),
            "parallelism": "columns sharded in contiguous blocks, one process per GPU, no data-path collective"}


def make_model(a, c):
    from subrecon_b200 import palmodel
    freqs = None
    if c.get("custom_pi"):
        freqs = palmodel.normalise_cli_frequencies(np.random.default_rng(4).dirichlet(5.0 * np.ones(20)))
    elif c.get("kind") == "example":
        ex = os.path.join(ROOT, "tests", "golden", "example")
        freqs = palmodel.normalise_cli_frequencies(open(os.path.join(ex, "freqs")).read().strip().split(","))
    model = palmodel.host_model(c["model"], freqs)
    model.custom_frequencies = freqs is not None
    return model


def make_tree(c):
    from subrecon_b200 import synth, treeio
    if c.get("kind") == "example":
        return treeio.read_tree(os.path.join(ROOT, "tests", "golden", "example", "root.tre"))
    return getattr(synth, TREES[c["tree"]])(c["taxa"], seed=c["tree_seed"])


def make_problem(a, c, seed_shift=0):
    """-> (tree, flat, model, rates, base states [taxa][base columns])"""
    from subrecon_b200 import palmodel, synth, treeio
    rates = palmodel.gamma_rates(c["ncat"], a.alpha)
    model = make_model(a, c)
    tree = make_tree(c)
    if c.get("kind") == "example":
        names, seqs = treeio.read_fasta(os.path.join(ROOT, "tests", "golden", "example", "prot.fasta"))
        c.setdefault("columns", len(seqs[0]))
        c.setdefault("base", len(seqs[0]))
        return tree, treeio.flatten_tree(tree, {n: i for i, n in enumerate(names)}), model, rates, treeio.encode_alignment(seqs)
    base = synth.simulate_alignment(tree, model, rates, min(c["base"], c["columns"]), seed=c["aln_seed"] + seed_shift)
    return tree, treeio.flatten_tree(tree), model, rates, base


def fill_states(a, c, dst, base, rank, device, tree, model, rates):
    """dst[taxa][L] <- L columns each simulated down the tree (torch ops on this rank's GPU: data generation, outside every
    timed region), or (--tiled) the numpy-simulated base block repeated. The table-row gathers of the pruning kernel see
    what real data would show them only when the columns are distinct (8 192 tiled columns keep the hot rows in L1/L2)."""
    L = dst.shape[1]
    if a.tiled or c.get("kind") == "example":
        for c0 in range(0, L, base.shape[1]):
            w = min(base.shape[1], L - c0)
            dst[:, c0:c0 + w] = base[:, :w]
        return
    import torch
    from subrecon_b200 import synth
    synth.simulate_alignment_torch(tree, model, rates, L, seed=c["aln_seed"] + 7919 * rank, device=device, out=dst)
    torch.cuda.empty_cache()


class ClockSampler(threading.Thread):
    """SM clock, power and clock-event (throttle) reasons of this rank's GPU while the timed regions run. Read in-process
    through NVML (the same source `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons...` reads — B200_PROFILING.md's
    clocks line): spawning nvidia-smi five times a second per rank makes eight of them queue on the driver at N=8 and stalls
    the ranks' own kernel launches, which is what a benchmark must not do to itself. Falls back to nvidia-smi when NVML
    cannot be loaded. Samples carry a time stamp so that a leg's own window can be reported (`window`)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, bus_id=None, period=0.1):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_ev, self.period = index, [], threading.Event(), period
        self.nvml, self.handle, self.max_mhz = None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode()) if bus_id else pynvml.nvmlDeviceGetHandleByIndex(index)
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.masks = [pynvml.nvmlClocksThrottleReasonHwSlowdown, pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                          pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, pynvml.nvmlClocksThrottleReasonSwPowerCap]
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample(self):
        if self.nvml is not None:
            n = self.nvml
            sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            watts = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
            bits = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            return sm, self.max_mhz, watts, [bool(bits & m) for m in self.masks]
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        f = [x.strip() for x in out.split(",")]
        return float(f[0]), float(f[1]), float(f[2]), [x.lower() == "active" for x in f[3:7]]

    def run(self):
        while not self._stop_ev.is_set():
            try:
                self.samples.append((time.perf_counter(),) + self._sample())
            except Exception:
                pass
            self._stop_ev.wait(self.period if self.nvml is not None else 0.5)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=6)
        return self.window(None, None)

    def window(self, t0, t1):
        """Statistics over the samples taken in [t0, t1] (perf_counter times; None = everything)."""
        ss = [x for x in self.samples if (t0 is None or x[0] >= t0) and (t1 is None or x[0] <= t1)]
        if not ss and self.samples and t0 is not None:          # a region shorter than the sampling period: nearest sample
            ss = [min(self.samples, key=lambda x: abs(x[0] - 0.5 * (t0 + t1)))]
        sm = [x[1] for x in ss]
        reasons = sorted({self.NAMES[i] for x in ss for i in range(4) if x[4][i]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": max((x[2] for x in ss), default=None), "power_w_max": max((x[3] for x in ss), default=None),
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def bind_to_gpu_numa_node(local):
    """Pin this rank (and therefore its page-locked buffers, first touch) to the NUMA node its GPU hangs off."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def cpu_reference_run(columns, threads, mode, flat, model, model_name, rates, states):
    """Time the oracle (the CPU restatement of the reference's per-column task) on `columns` columns. The oracle builds its
    own model from PAL's tables and the same frequencies the product arm was given."""
    from oracle import oracle
    om = oracle.OracleModel(ORACLE_MODEL[model_name], np.asarray(model.pi) if getattr(model, "custom_frequencies", False) else None)
    t0 = time.perf_counter()
    oracle.run(flat, np.ascontiguousarray(states[:, :columns]), om, rates, mode=mode, threads=threads, want_joint=True)
    return time.perf_counter() - t0


def calibrate_cpu_sample(threads, mode, flat, model, model_name, rates, states, target_s):
    n = min(threads, states.shape[1])
    dt = cpu_reference_run(n, threads, mode, flat, model, model_name, rates, states)
    per_round = max(dt, 1e-3)                       # `threads` columns take ~per_round seconds
    rounds = max(1, int(target_s / per_round))
    return min(states.shape[1], n * rounds)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    c = config_of(a)
    tree, flat, model, rates, base = make_problem(a, c)
    threads = os.cpu_count() or 1
    sample = calibrate_cpu_sample(threads, 2, flat, model, c["model"], rates, base, target_s=max(2.0, 40.0 / max(1, a.steps + a.warmup)))
    for _ in range(a.warmup):
        cpu_reference_run(sample, threads, 2, flat, model, c["model"], rates, base)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cpu_reference_run(sample, threads, 2, flat, model, c["model"], rates, base)
    dt = time.perf_counter() - t0
    v = sample * a.steps / dt
    desc = (f"{sample} columns of the workload per step; oracle work mode 2 = per column, per edge, per rate class: rate-matrix "
            f"rebuild + elmhes/eltran/hqr2/luinverse + 20^3 P product, as PAL 1.51 executes it (SURVEY B.2); C port, -O2, "
            f"no JVM in this image")
    cols, world = c["columns"], max(1, a.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": a.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic" if c.get("kind") != "example" else "reference example files",
            "config": workload_config(a, c, cols if a.scaling == "weak" else -(-cols // world), world),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def run_b200(a):
    import torch
    import torch.distributed as dist
    from subrecon_b200 import native, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product arm has no CPU fallback")
    torch.cuda.set_device(local)
    affinity0 = os.sched_getaffinity(0)
    numa_node = bind_to_gpu_numa_node(local)          # page-locked buffers land on the GPU's NUMA node (first touch)
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")     # host-side rendezvous that leaves the GPUs idle (single-process leg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    c = config_of(a)
    tree, flat, model, rates, base = make_problem(a, c, seed_shift=rank)
    L_total = c["columns"]
    L = sharding.column_block(L_total, world, rank)[1] if a.scaling == "strong" else L_total
    N = base.shape[0]
    total_columns = L_total if a.scaling == "strong" else L * world
    ctx = native.Context(local)
    if a.chunk_cols:
        ctx.set_option("chunk_cols", a.chunk_cols)
    ctx.set_model(model.Eval, model.Evec, model.Ievc, model.pi)
    ctx.set_rates(rates)
    ctx.set_tree(flat)
    info = ctx.schedule_info()
    flops_per_col = info["flops_per_column_rate"] * len(rates)
    # Clade tables: a collapsed cherry's two tip products and its 20x20 contraction (2*20 + 820 algorithmic flops per
    # column and rate class) are replaced by one 20-flop gather product. `achieved` below stays ALGORITHMIC (SURVEY
    # §8d: what the reference's arithmetic costs); `executed` counts only what the kernel still computes per column.
    _, node_slot, _ = native.schedule_dump(flat)
    n_cherry = n_fork = 0
    for v in np.nonzero(node_slot <= -2)[0]:
        kids = flat.child_idx[flat.child_off[v]:flat.child_off[v + 1]]
        if all(flat.child_off[x] == flat.child_off[x + 1] for x in kids):
            n_cherry += 1                  # two tip products + one contraction -> one gather product
        else:
            n_fork += 1                    # (cherry, tip): three tip products + two contractions -> one gather product
    executed_per_col = flops_per_col - len(rates) * (n_cherry * (820.0 + 20.0) + n_fork * (2 * 820.0 + 2 * 20.0))

    # host (pinned) copies of the step's inputs and outputs
    h_states = native.PinnedArray((N, L), np.uint8)
    fill_states(a, c, h_states.array, base, rank, f"cuda:{local}", tree, model, rates)
    h_lnl = native.PinnedArray(L, np.float64)

    # ---- value: alignment resident in HBM, results stay in HBM
    ctx.set_alignment(h_states.array)
    d_lnl = torch.empty(L, dtype=torch.float64, device="cuda")
    d_joint = torch.empty((L, 400), dtype=torch.float64, device="cuda")
    tstream = torch.cuda.Stream()                      # a real (non-default) stream: the library launches on it, the
    stream = tstream.cuda_stream                       # events below are recorded on it

    def step_resident():
        ctx.set_rates(rates)            # marks the matrices dirty: K1 + K1b run inside every step
        ctx.run_device(0, L, d_lnl.data_ptr(), d_joint.data_ptr(), None, 0.5, stream)

    for _ in range(a.warmup):
        step_resident()
    peak_tflops = ctx.measure_fp64_peak(150.0)
    ctx.set_option("profile", 1)
    ctx.profile(reset=True)
    props = torch.cuda.get_device_properties(local)
    bus_id = "%08x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), getattr(props, "pci_bus_id", 0), getattr(props, "pci_device_id", 0))
    sampler = ClockSampler(local, bus_id if hasattr(props, "pci_bus_id") else None)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_res0 = time.perf_counter()
    ev0.record(tstream)
    for _ in range(a.steps):
        step_resident()
    ev1.record(tstream)
    barrier()
    t_res1 = time.perf_counter()
    ms_rank = ev0.elapsed_time(ev1)
    ms = max_over_ranks(ms_rank)
    prof = ctx.profile(reset=True)
    ctx.set_option("profile", 0)
    value = total_columns * a.steps / (ms * 1e-3)
    total_lnl = float(d_lnl.sum().item())
    if world > 1:
        t = torch.tensor([total_lnl], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)          # the path's only cross-GPU value (SubRecon.java:137,143)
        total_lnl = float(t.item())
    k2_ms = prof["ms_prune"] / max(1, prof["n_prune"])
    k2_cols = prof["columns_pruned"] / max(1, prof["n_prune"])
    achieved = flops_per_col * k2_cols / (k2_ms * 1e-3) * 1e-12
    launches = int(prof["n_pbuild"] + prof["n_prune"] + prof["n_root"] + prof["n_index"])
    d_lnl_host = d_lnl.cpu().numpy()
    del d_joint
    torch.cuda.empty_cache()

    def time_e2e(step, warm):
        for _ in range(warm):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        return dt

    # ---- e2e: host buffers, H2D + D2H inside the timed region, through the C ABI call a host makes; the compact record
    #      at the default -threshold 0.5 is the result the reference's consumer keeps (SiteResult.java:55,69-83)
    e2e = e2e_full = e2e_single = None
    comp_dtype = native.compact_dtype(8)
    comp_ref = None
    if not a.no_e2e:
        h_comp = native.PinnedArray(L, comp_dtype)
        outc = {"lnl": h_lnl.array, "compact": h_comp.array}

        def step_e2e():
            ctx.set_rates(rates)
            ctx.run_streamed(h_states.array, 0, L, want_joint=False, want_compact=True, threshold=0.5, out=outc)

        dtc = time_e2e(step_e2e, max(1, min(a.warmup, 2)))
        e2e = {"value": total_columns * a.steps / dtc, "unit": UNIT, "h2d_bytes_per_step": int(N) * L,
               "d2h_bytes_per_step": L * (8 + comp_dtype.itemsize), "ms_per_step": dtc / a.steps * 1e3,
               "call": "sr_run_streamed, one process per GPU: pinned host states in; lnl + compact record (lnl, maxII, max, entries >= "
                       "0.5: what SiteResult keeps) per column out",
               "threshold": 0.5, "lnl_bit_identical_to_resident": bool(np.array_equal(h_comp.array["lnl"], d_lnl_host))}
        comp_ref = h_comp.array[:min(L, 65536)].copy()
        h_comp.free()

    # ---- the same with all 400 probabilities coming back, and the host's D2H ceiling next to it
    if not a.no_e2e and not a.no_full_joint:
        h_joint = native.PinnedArray((L, 20, 20), np.float64)
        out = {"lnl": h_lnl.array, "joint": h_joint.array}

        def step_full():
            ctx.set_rates(rates)
            ctx.run_streamed(h_states.array, 0, L, want_joint=True, out=out)

        dt = time_e2e(step_full, 1)
        # every rank drains a device buffer into its pinned result array at the same time: what the host can ingest
        nbytes = int(min(h_joint.array.nbytes, 1 << 30))
        src = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        dst = torch.from_numpy(h_joint.array.reshape(-1).view(np.uint8)[:nbytes])
        dst.copy_(src)
        barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dtp = max_over_ranks(time.perf_counter() - t0)
        ceiling = world * 4 * nbytes / dtp
        e2e_full = {"value": total_columns * a.steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(N) * L,
                    "d2h_bytes_per_step": L * (8 + 3200), "ms_per_step": dt / a.steps * 1e3,
                    "call": "sr_run_streamed with joint output: lnl + 400 joint probabilities per column out",
                    "host_d2h_ceiling": {"GB_per_s_all_ranks": ceiling * 1e-9, "columns_per_s": ceiling / 3208.0,
                                         "how": f"{world} rank(s) each copying {nbytes >> 20} MiB device->pinned host 4x concurrently"},
                    "bit_identical_to_resident": bool(np.array_equal(h_lnl.array, d_lnl_host))}
        del src, dst
        h_joint.free()

    # ---- one host process drives all N GPUs (sr_multi_*: the route a single JVM takes); ranks > 0 idle on the host meanwhile
    if not a.no_e2e and not a.no_single_process:
        if world > 1:
            dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                n_multi = world
                if world == 1 and a.single_process_devices > 1:
                    n_multi = min(a.single_process_devices, torch.cuda.device_count())
                Lm = total_columns if n_multi == world else L * n_multi
                m = native.MultiContext(list(range(n_multi)))
                m.set_model(model.Eval, model.Evec, model.Ievc, model.pi)
                m.set_rates(rates)
                m.set_tree(flat)
                own_states = Lm != L
                ms_states = native.PinnedArray((N, Lm), np.uint8) if own_states else h_states
                if own_states:
                    fill_states(a, c, ms_states.array, base, 0, f"cuda:{local}", tree, model, rates)
                ms_lnl = native.PinnedArray(Lm, np.float64)
                ms_comp = native.PinnedArray(Lm, comp_dtype)
                outm = {"lnl": ms_lnl.array, "compact": ms_comp.array}

                def step_multi():
                    m.set_rates(rates)
                    m.run_streamed(ms_states.array, 0, Lm, want_joint=False, want_compact=True, threshold=0.5, out=outm)

                step_multi()
                t0 = time.perf_counter()
                for _ in range(a.steps):
                    step_multi()
                dtm = time.perf_counter() - t0
                same = bool(ms_comp.array[:len(comp_ref)].tobytes() == comp_ref.tobytes())
                e2e_single = {"value": Lm * a.steps / dtm, "unit": UNIT, "h2d_bytes_per_step": int(N) * Lm,
                              "d2h_bytes_per_step": Lm * (8 + comp_dtype.itemsize), "ms_per_step": dtm / a.steps * 1e3, "devices": m.count,
                              "call": "sr_multi_run_streamed from ONE host process (rank 0): a context and a host thread per GPU inside the "
                                      "library, contiguous column blocks, compact records land in place",
                              "first_records_identical_to_per_process_run": same}
                m.close()
                for h in (ms_lnl, ms_comp) + ((ms_states,) if own_states else ()):
                    h.free()
            except Exception as exc:                        # never lose the main line over the extra leg
                e2e_single = {"error": str(exc)}
        if world > 1:
            dist.barrier(group=cpu_group)

    clocks_all = sampler.stop()                      # every leg (resident and end-to-end)
    clocks = sampler.window(t_res0, t_res1)          # the region `value` was timed over
    clocks["whole_run"] = {k: clocks_all[k] for k in ("sm_mhz", "sm_min_mhz", "power_w_max", "reasons", "samples")}
    # every rank's own view of the resident region: step time, kernel time, clocks, power (the reported time is their max)
    mine = {"rank": rank, "ms_per_step": ms_rank / a.steps, "prune_ms": prof["ms_prune"] / max(1, prof["n_prune"]),
            "sm_mhz": clocks["sm_mhz"], "sm_min_mhz": clocks["sm_min_mhz"], "power_w_max": clocks["power_w_max"], "reasons": clocks["reasons"]}
    per_rank = [mine]
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine, group=cpu_group)
        reasons_any = sorted({r for x in per_rank for r in x["reasons"]})
        clocks["reasons"] = reasons_any             # any rank's throttle reason counts for the run
        clocks["sm_mhz"] = float(np.median([x["sm_mhz"] for x in per_rank if x["sm_mhz"] is not None]))
        clocks["sm_min_mhz"] = min(x["sm_min_mhz"] for x in per_rank if x["sm_min_mhz"] is not None)

    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (same launch size), else null
    traffic, traffic_note = None, None
    try:
        if a.config == 3 and L == 1_000_000:
            prof_json = json.load(open(os.path.join(ROOT, "profiles", "r02_prune_kernel_ncu_full_1Mcols.json")))
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            traffic = sum(float(prof_json[k]["value"]) * scale[prof_json[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            traffic_note = ("bytes per launch, dram__bytes_read.sum + dram__bytes_write.sum of prune_kernel from "
                            "profiles/r02_prune_kernel_ncu_full_1Mcols.json (same launch size)")
    except Exception:
        traffic = None

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic" if c.get("kind") != "example" else "reference example files",
            "config": workload_config(a, c, L, world),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / peak_tflops, "traffic": traffic, "traffic_note": traffic_note,
                         "kernel": "prune_kernel (FP64 DMMA m8n8k4)", "kernel_ms": k2_ms,
                         "kernel_share_of_step": prof["ms_prune"] / ms if world == 1 else None,
                         "flops_per_column": flops_per_col,
                         "executed": achieved * executed_per_col / flops_per_col,
                         "frac_executed": achieved * executed_per_col / flops_per_col / peak_tflops,
                         "executed_note": f"{n_cherry} cherries and {n_fork} pitchforks (cherry + tip) of the tree are tabulated per state "
                                          "combination once per rate class (K1b) and gathered instead of contracted per column: `achieved`/`frac` count the algorithmic "
                                          "flops of SURVEY 8d, `executed`/`frac_executed` only the flops prune_kernel still performs "
                                          "per column (its tensor-pipe efficiency; ceiling 0.83 from the 20->24 padding)",
                         "peak_source": "live DMMA probe (sr_measure_fp64_peak) on this GPU; MEASURED_PEAKS.json has no FP64 entry; "
                                        "nominal 148 SM x 64 FMA/clk x 1.965 GHz = 37.2 TFLOP/s; cuBLAS DGEMM beside it: profiles/fp64_peak_r02.json"},
            "e2e": e2e, "e2e_full_joint": e2e_full, "e2e_single_process": e2e_single, "numa_node": numa_node, "gpu_launches": launches,
            "kernel_ms": {"pbuild_and_clade_tables_per_step": prof["ms_pbuild"] / max(1, a.steps),
                          "clade_index": prof["ms_index"] / max(1, prof["n_index"]), "prune": k2_ms,
                          "root": prof["ms_root"] / max(1, prof["n_root"])},
            "clocks": clocks, "per_rank": per_rank, "total_lnl": total_lnl,
            "schedule": {"steps": info["n_steps"], "stack_depth": info["stack_depth"], "cherries": n_cherry, "pitchforks": n_fork}}

    # ---- cpu_baseline: the reference's CPU algorithm on this box's host cores (rank 0, N=1 only)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        os.sched_setaffinity(0, affinity0)               # the CPU baseline gets every host core back
        threads = os.cpu_count() or 1
        ctx.close()
        res = {}
        for mode, key, tgt in ((2, "A_jar_faithful", a.cpu_seconds), (1, "B_p_rebuild_only", a.cpu_seconds / 3),
                               (0, "C_hoisted_honest_cpu", a.cpu_seconds / 3)):
            n = calibrate_cpu_sample(threads, mode, flat, model, c["model"], rates, base, tgt)
            dt = cpu_reference_run(n, threads, mode, flat, model, c["model"], rates, base)
            res[key] = {"columns_per_s": n / dt, "sample_columns": n, "seconds": dt}
        line["cpu_baseline"] = {
            "value": res["A_jar_faithful"]["columns_per_s"], "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{res['A_jar_faithful']['sample_columns']} columns of the same workload; oracle work mode 2 "
                      "(jar-faithful: PAL 1.51 redoes the eigendecomposition on every setDistance, SURVEY B.2); "
                      "C port with -O2, so it flatters the JVM",
            "variants": res}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else any library prints to fd 1 during the run
    (NCCL's version banner, for one) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    a = parse_args()
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)                          # C-level and Python-level stdout of this process now land on stderr
    if a.impl == "reference":
        return run_reference(a)
    return run_b200(a)


if __name__ == "__main__":
    sys.exit(main())
