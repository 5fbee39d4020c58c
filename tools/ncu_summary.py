"""Summarise one kernel of an .ncu-rep (a `--set full --import-source on` capture) into the JSON files kept under profiles/.
usage: python tools/ncu_summary.py <rep> <out.json> "<command that was profiled>" [--mix <mix.json> --warp-steps N]"""
import csv
import io
import json
import subprocess
import sys
import collections

KEEP = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active sm__warps_active.avg.pct_of_peak_sustained_active
launch__registers_per_thread launch__block_size launch__grid_size launch__shared_mem_per_block_dynamic
sm__issue_active.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum smsp__warps_eligible.avg.per_cycle_active
smsp__warps_active.avg.per_cycle_active smsp__inst_executed.sum sm__cycles_elapsed.avg.per_second sm__cycles_elapsed.max lts__t_sector_hit_rate.pct
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active smsp__issue_active.avg.pct_of_peak_sustained_active
lts__t_sectors_srcunit_tex_op_read.sum l1tex__t_sector_hit_rate.pct sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed lts__t_bytes.sum""".split()

rep, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
names, units, vals = rows[0], rows[1], rows[2]
d = {"_command": cmd, "_kernel": vals[names.index("Kernel Name")] if "Kernel Name" in names else ""}
for n, u, v in zip(names, units, vals):
    if n in KEEP or n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"):
        try:
            d[n] = {"unit": u, "value": float(v.replace(",", ""))}
        except ValueError:
            pass
json.dump(d, open(out, "w"), indent=1)
print("wrote", out, len(d), "entries; kernel:", d["_kernel"])
if "--mix" in sys.argv:
    mix_out = sys.argv[sys.argv.index("--mix") + 1]
    ws = float(sys.argv[sys.argv.index("--warp-steps") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ia, isrc, isam = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    ops, samples, stalls = collections.Counter(), collections.Counter(), collections.Counter()
    tot = tots = 0
    for r in rows[2:]:
        if len(r) <= ia or not r[ia]:
            continue
        n, sm = int(r[ia]), int(r[isam] or 0)
        toks = r[isrc].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        ops[op] += n; samples[op] += sm; tot += n; tots += sm
        for i in stall_cols:
            if r[i]:
                stalls[hdr[i]] += int(r[i])
    mix = {"_command": cmd, "_source": "ncu --page source: Instructions Executed per SASS line, summed per opcode, divided by the consumer warp-steps of the launch",
           "consumer_warp_steps": ws, "instructions_per_warp_step": tot / ws,
           "per_opcode": {op: {"per_warp_step": n / ws, "pct_of_samples": 100.0 * samples[op] / max(1, tots)} for op, n in ops.most_common(40)},
           "stall_samples_pct": {k: 100.0 * v / max(1, tots) for k, v in stalls.most_common(12)}}
    json.dump(mix, open(mix_out, "w"), indent=1)
    print("wrote", mix_out, "instructions per warp-step:", round(tot / ws, 1))
