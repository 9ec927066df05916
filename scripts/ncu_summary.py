"""Summarise an .ncu-rep: per-kernel headline metrics and (optionally) the dynamic instruction mix / top stall lines.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--mix regex] [--n-particles N]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
mix = sys.argv[sys.argv.index("--mix") + 1] if "--mix" in sys.argv else None
npart = int(sys.argv[sys.argv.index("--n-particles") + 1]) if "--n-particles" in sys.argv else 1 << 24
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:90])
    for w in WANT:
        if w in hdr:
            print(f"   {w:88s} {r[hdr.index(w)]}")
if mix:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{mix}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, samp, tot, lines = collections.Counter(), collections.Counter(), 0, []
    for r in rows[2:]:
        if len(r) < len(hdr) or r[0] in ("Kernel Name", "Address"):
            break
        try:
            e, s = int(r[iE]), int(r[iN])
        except ValueError:
            continue
        op = r[iS].split()
        if not op:
            continue
        o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
        ops[o] += e; samp[o] += s; tot += e
        lines.append((s, e, r[iS].strip()[:100]))
    print(f"-- {mix}: warp instructions {tot}, per particle (lane instr) {tot * 32 / npart:.1f}")
    for o, c in ops.most_common(22):
        print(f"   {o:10s} {c * 32 / npart:7.1f}/particle   stall samples {samp[o]}")
    print("-- top stall-sample instructions")
    for s, e, t in sorted(lines, reverse=True)[:14]:
        print(f"   {s:6d} samples  {e:10d} exec  {t}")
