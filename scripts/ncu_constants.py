"""Writes profiles/ncu_constants.json from committed-name ncu reports: the per-launch figures only a profiler can give (DRAM bytes,
executed instructions, issue-slot utilisation), which bench.py quotes as provenance next to its own live timings.
usage: python scripts/ncu_constants.py gpurun_out/r2_prof_nested.ncu-rep gpurun_out/r2_prof_is.ncu-rep gpurun_out/r2_prof_mh.ncu-rep"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    return hdr, units, r[2:]


def val(hdr, units, row, name):
    v = float(row[hdr.index(name)].replace(",", ""))
    u = units[hdr.index(name)]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(u, 1.0)


out = {}
for rep in sys.argv[1:]:
    hdr, units, rs = rows(rep)
    for row in rs:
        name = row[hdr.index("Kernel Name")]
        g = lambda k: val(hdr, units, row, k)
        common = {"kernel": name[:120], "duration_us_under_ncu": g("gpu__time_duration.sum") * 1e6, "warp_instructions_per_launch": g("smsp__inst_executed.sum"),
                  "issue_slot_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"), "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
                  "registers_per_thread": g("launch__registers_per_thread"), "source": "profiles/" + os.path.basename(rep).replace(".ncu-rep", "") + " (ncu --set full --clock-control none)"}
        dram = g("dram__bytes_read.sum") + g("dram__bytes_write.sum")
        if "pf_extend_kernel" in name and "Lgssm4<float>" in name and ", 2, 0, 1>" in name:
            out["extend_2p24_nested"] = dict(common, dram_bytes_per_launch=dram, algorithmic_bytes_per_particle=40, lane_instructions_per_particle=common["warp_instructions_per_launch"] * 32 / (1 << 24))
        elif "nested_expand_kernel" in name:
            out["expand_2p24_nested"] = dict(common, dram_bytes_per_launch=dram, algorithmic_bytes_per_particle=8, lane_instructions_per_particle=common["warp_instructions_per_launch"] * 32 / (1 << 24))
        elif "nested_sections_kernel" in name:
            out["sections_2p24_nested"] = dict(common, dram_bytes_per_launch=dram)
        elif "nested_plan_kernel" in name:
            out["plan_2p24_nested"] = dict(common, dram_bytes_per_launch=dram)
        elif "is_kernel" in name and "LineModel" in name:
            out["is_kernel"] = dict(common, unit="% of issue slots", achieved=common["issue_slot_pct"], frac=common["issue_slot_pct"] / 100.0,
                                    lane_instructions_per_proposal=common["warp_instructions_per_launch"] * 32 / (1 << 20), dram_bytes_per_launch=dram)
        elif "is_kernel" in name and "Hierarchical" in name:
            out["is_kernel_hierarchical"] = dict(common, lane_instructions_per_proposal=common["warp_instructions_per_launch"] * 32 / (1 << 20), dram_bytes_per_launch=dram)
        elif "weight_reduce_kernel<double>" in name:
            out["is_weight_reduce"] = dict(common, dram_bytes_per_launch=dram)
        elif "mh_schedule_kernel" in name:
            steps = (1 << 20) * 8 * 18
            out["mh_sweep_kernel"] = dict(common, unit="% of issue slots", achieved=common["issue_slot_pct"], frac=common["issue_slot_pct"] / 100.0,
                                          lane_instructions_per_chain_step=common["warp_instructions_per_launch"] * 32 / steps,
                                          note="register-resident: occupancy is limited by 112 registers per thread (25 % of the warp slots)")
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_constants.json"), "w"), indent=1)
print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk not in ("kernel", "source")} for k, v in out.items()}, indent=1)[:3000])
