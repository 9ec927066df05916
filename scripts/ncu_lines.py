"""Per-source-line instruction counts and stall samples of one kernel in an .ncu-rep (needs -lineinfo and --import-source on).
usage: python scripts/ncu_lines.py report.ncu-rep kernel_regex [n_units]   (n_units: particles / slots per launch, for per-unit figures)"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else float(1 << 24)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows, cur_file, hdr = [], "", None
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif r[0] == "Function Name":
        if rows and "--first" in sys.argv:
            break
    elif hdr and r[0].isdigit():
        try:
            rows.append((cur_file, int(r[0]), r[1].strip()[:110], int(r[hdr.index("# Samples")]), int(r[hdr.index("Thread Instructions Executed")])))
        except ValueError:
            pass
tot = sum(x[4] for x in rows)
print(f"total lane instr/unit {tot / units:.1f}")
for f, ln, src, smp, ti in sorted(rows, key=lambda x: -x[4])[:45]:
    print(f"{ti / units:7.2f}/unit {smp:6d} smp  {f}:{ln}  {src}")
