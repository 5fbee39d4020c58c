#!/usr/bin/env python
"""bench.py — alignment columns/sec of the SubRecon hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # product arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU algorithm on host cores

A "step" is one pass of the hot path (P matrices -> pruning -> root-branch joint) over one batch of synthetic
alignment columns: BASELINE config #3, 1,000 taxa x 1,000,000 sites, WAG, K=4, alpha=0.5, Yule tree (seed 3).
Weak scaling: every rank (one process per GPU) owns a block of that size; columns never interact, so there is no
data-path collective — only the barrier, the max-over-ranks of the time and a reduce of the total lnL.

  value   whole-job columns/s with the alignment resident in HBM and results left in HBM (CUDA events)
  e2e     the same metric through sr_run_streamed with HOST buffers: pinned states H2D and lnl+400 probabilities
          D2H inside the timed region, pipelined in column chunks
  roofline  the pruning kernel's algorithmic FP64 flops (K*(820*E_int + 20*E_leaf + 1200) per column) over
            its CUDA-event time, against the FP64 tensor-pipe (DMMA) peak probed live on the same GPU
            (MEASURED_PEAKS.json has no FP64 entry; see profiles/fp64_peak_r01.json)
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--taxa", type=int, default=1000)
    ap.add_argument("--columns", type=int, default=1_000_000, help="columns per GPU")
    ap.add_argument("--model", default="wag")
    ap.add_argument("--ncat", type=int, default=4)
    ap.add_argument("--alpha", type=float, default=0.5)
    ap.add_argument("--chunk-cols", type=int, default=0, help="0 = library default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def workload_config(a):
    return {"workload": f"BASELINE config #3: synthetic {a.taxa} taxa x {a.columns} sites per GPU, {a.model.upper()}, "
                        f"K={a.ncat}, alpha={a.alpha}, Yule tree seed 3, columns simulated down the tree (8192 distinct, tiled)",
            "taxa": a.taxa, "columns_per_gpu": a.columns, "model": a.model, "rate_classes": a.ncat, "alpha": a.alpha,
            "l2_policy": "inputs larger than L2: 1 GB of states in, 3.2 GB of results out per step (L2 = 126 MB)",
            "parallelism": "columns sharded in contiguous blocks, one process per GPU, no data-path collective"}


def make_problem(a, seed_shift=0):
    from subrecon_b200 import palmodel, synth, treeio
    tree = synth.yule_tree(a.taxa, seed=3)
    model = palmodel.host_model(a.model)
    rates = palmodel.gamma_rates(a.ncat, a.alpha)
    base = synth.simulate_alignment(tree, model, rates, min(8192, a.columns), seed=103 + seed_shift)
    return tree, treeio.flatten_tree(tree), model, rates, base


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_ev = index, [], threading.Event()

    def run(self):
        while not self._stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_ev.wait(0.2)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) >= 7 and s[3 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


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
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def cpu_reference_run(a, columns, threads, mode, flat, model_name, rates, states):
    """Time the oracle (the CPU restatement of the reference's per-column task) on `columns` columns."""
    from oracle import oracle
    om = oracle.OracleModel({"wag": "WAG", "jtt": "JTT", "dayhoff": "Dayhoff", "blosum62": "BLOSUM62"}[model_name])
    t0 = time.perf_counter()
    oracle.run(flat, states[:, :columns], om, rates, mode=mode, threads=threads, want_joint=True)
    return time.perf_counter() - t0


def calibrate_cpu_sample(a, threads, mode, flat, rates, states, target_s):
    n = threads
    dt = cpu_reference_run(a, n, threads, mode, flat, a.model, rates, states)
    per_round = max(dt, 1e-3)                       # `threads` columns take ~per_round seconds
    rounds = max(1, int(target_s / per_round))
    return min(states.shape[1], n * rounds)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    tree, flat, model, rates, base = make_problem(a)
    threads = os.cpu_count() or 1
    sample = calibrate_cpu_sample(a, threads, 2, flat, rates, base, target_s=max(2.0, 40.0 / max(1, a.steps + a.warmup)))
    for _ in range(a.warmup):
        cpu_reference_run(a, sample, threads, 2, flat, a.model, rates, base)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cpu_reference_run(a, sample, threads, 2, flat, a.model, rates, base)
    dt = time.perf_counter() - t0
    v = sample * a.steps / dt
    desc = (f"{sample} columns of the workload per step; oracle work mode 2 = per column, per edge, per rate class: rate-matrix "
            f"rebuild + elmhes/eltran/hqr2/luinverse + 20^3 P product, as PAL 1.51 executes it (SURVEY B.2); C port, -O2, "
            f"no JVM in this image")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(a),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def run_b200(a):
    import torch
    import torch.distributed as dist
    from subrecon_b200 import native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product arm has no CPU fallback")
    torch.cuda.set_device(local)
    affinity0 = os.sched_getaffinity(0)
    numa_node = bind_to_gpu_numa_node(local)          # page-locked buffers land on the GPU's NUMA node (first touch)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

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

    tree, flat, model, rates, base = make_problem(a, seed_shift=rank)
    L, N = a.columns, base.shape[0]
    ctx = native.Context(local)
    if a.chunk_cols:
        ctx.set_option("chunk_cols", a.chunk_cols)
    ctx.set_model(model.Eval, model.Evec, model.Ievc, model.pi)
    ctx.set_rates(rates)
    ctx.set_tree(flat)
    info = ctx.schedule_info()
    flops_per_col = info["flops_per_column_rate"] * len(rates)
    # Cherry tables: a collapsed cherry's two tip products and its 20x20 contraction (2*20 + 820 algorithmic flops per
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
    for c0 in range(0, L, base.shape[1]):
        w = min(base.shape[1], L - c0)
        h_states.array[:, c0:c0 + w] = base[:, :w]
    h_lnl = native.PinnedArray(L, np.float64)
    h_joint = native.PinnedArray((L, 20, 20), np.float64)

    # ---- value: alignment resident in HBM, results stay in HBM
    ctx.set_alignment(h_states.array)
    d_lnl = torch.empty(L, dtype=torch.float64, device="cuda")
    d_joint = torch.empty((L, 400), dtype=torch.float64, device="cuda")
    tstream = torch.cuda.Stream()                      # a real (non-default) stream: the library launches on it, the
    stream = tstream.cuda_stream                       # events below are recorded on it

    def step_resident():
        ctx.set_rates(rates)            # marks the matrices dirty: K1 runs inside every step
        ctx.run_device(0, L, d_lnl.data_ptr(), d_joint.data_ptr(), None, 0.5, stream)

    for _ in range(a.warmup):
        step_resident()
    peak_tflops = ctx.measure_fp64_peak(150.0)
    ctx.set_option("profile", 1)
    ctx.profile(reset=True)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(tstream)
    for _ in range(a.steps):
        step_resident()
    ev1.record(tstream)
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    prof = ctx.profile(reset=True)
    ctx.set_option("profile", 0)
    value = world * L * a.steps / (ms * 1e-3)
    total_lnl = float(d_lnl.sum().item())
    if world > 1:
        t = torch.tensor([total_lnl], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)          # the path's only cross-GPU value (SubRecon.java:137,143)
        total_lnl = float(t.item())
    k2_ms = prof["ms_prune"] / max(1, prof["n_prune"])
    k2_cols = prof["columns_pruned"] / max(1, prof["n_prune"])
    achieved = flops_per_col * k2_cols / (k2_ms * 1e-3) * 1e-12
    launches = int(prof["n_pbuild"] + prof["n_prune"] + prof["n_root"])

    # ---- e2e: host buffers, H2D + D2H inside the timed region, through the C ABI call a host makes
    e2e = None
    if not a.no_e2e:
        out = {"lnl": h_lnl.array, "joint": h_joint.array}

        def step_e2e():
            ctx.set_rates(rates)
            ctx.run_streamed(h_states.array, 0, L, want_joint=True, out=out)

        for _ in range(max(1, min(a.warmup, 2))):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step_e2e()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e = {"value": world * L * a.steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(N) * L,
               "d2h_bytes_per_step": L * (8 + 3200), "ms_per_step": dt / a.steps * 1e3,
               "call": "sr_run_streamed (pinned host states in; lnl + 400 joint probabilities per column out)",
               "bit_identical_to_resident": bool(np.array_equal(h_lnl.array, d_lnl.cpu().numpy())) and
                                            bool(np.array_equal(h_joint.array[:4096].reshape(-1, 400), d_joint[:4096].cpu().numpy()))}

    # ---- e2e, compact result: what the reference's SiteResult keeps per column (lnL, maxII, max, entries >= threshold;
    #      recon/SiteResult.java:55,69-83) at the default -threshold 0.5 — 112 B instead of 3 208 B per column back to the host
    e2e_compact = None
    if not a.no_e2e:
        h_comp = native.PinnedArray(L, native.COMPACT_DTYPE)
        outc = {"lnl": h_lnl.array, "compact": h_comp.array}

        def step_e2e_compact():
            ctx.set_rates(rates)
            ctx.run_streamed(h_states.array, 0, L, want_joint=False, want_compact=True, threshold=0.5, out=outc)

        step_e2e_compact()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step_e2e_compact()
        torch.cuda.synchronize()
        dtc = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e_compact = {"value": world * L * a.steps / dtc, "unit": UNIT, "h2d_bytes_per_step": int(N) * L,
                       "d2h_bytes_per_step": L * (8 + native.COMPACT_DTYPE.itemsize), "ms_per_step": dtc / a.steps * 1e3,
                       "lnl_identical_to_full": bool(np.array_equal(h_comp.array["lnl"], d_lnl.cpu().numpy()))}

    clocks = sampler.stop()      # sampled over the resident AND the end-to-end timed regions

    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (same launch size), else null
    traffic = None
    try:
        if (a.taxa, a.columns, a.ncat, a.model) == (1000, 1_000_000, 4, "wag"):
            prof_json = json.load(open(os.path.join(ROOT, "profiles", "r01_prune_kernel_clade_tables_ncu_full_1Mcols.json")))
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            traffic = sum(float(prof_json[k]["value"]) * scale[prof_json[k]["unit"]]
                          for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except Exception:
        traffic = None

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(a),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / peak_tflops, "traffic": traffic,
                         "traffic_note": "bytes per launch, dram__bytes_read.sum + dram__bytes_write.sum of prune_kernel from "
                                         "profiles/r01_prune_kernel_clade_tables_ncu_full_1Mcols.json; the kernel's own HBM need is "
                                         "K*N bytes of states in + K*328 bytes of root partials out per column = 5.3 GB per launch, plus "
                                         "the rows of the clade tables that miss L2 (0.8 GB measured; L2 hit rate 92 %) "
                                         "(0.9 % of the HBM roof: the path is FP64-tensor-bound, not HBM-bound)",
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
                                        "nominal 148 SM x 64 FMA/clk x 1.965 GHz = 37.2 TFLOP/s"},
            "e2e": e2e, "e2e_compact": e2e_compact, "numa_node": numa_node, "gpu_launches": launches,
            "kernel_ms": {"pbuild_and_clade_tables_per_step": prof["ms_pbuild"] / max(1, a.steps), "prune": k2_ms,
                          "root": prof["ms_root"] / max(1, prof["n_root"])},
            "clocks": clocks, "total_lnl": total_lnl,
            "schedule": {"steps": info["n_steps"], "stack_depth": info["stack_depth"]}}

    # ---- cpu_baseline: the reference's CPU algorithm on this box's host cores (rank 0, N=1 only)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        os.sched_setaffinity(0, affinity0)               # the CPU baseline gets every host core back
        threads = os.cpu_count() or 1
        ctx.close()
        res = {}
        for mode, key, tgt in ((2, "A_jar_faithful", a.cpu_seconds), (1, "B_p_rebuild_only", a.cpu_seconds / 3),
                               (0, "C_hoisted_honest_cpu", a.cpu_seconds / 3)):
            n = calibrate_cpu_sample(a, threads, mode, flat, rates, base, tgt)
            dt = cpu_reference_run(a, n, threads, mode, flat, a.model, rates, base)
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
