"""Top source lines of a kernel by stall samples / executed instructions from an ncu report captured with --import-source on.
Usage: python tools/ncu_lines.py gpurun_out/prof_mc.ncu-rep [n]"""
import csv, io, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if r and r[0] == "Line No")
ci, cs = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
lines = []
for r in rows:
    if r and r[0].isdigit() and len(r) > ci:
        try:
            lines.append((int(r[0]), r[1][:100], int(r[cs] or 0), int(r[ci] or 0)))
        except ValueError:
            pass
ts, ti = sum(l[2] for l in lines) or 1, sum(l[3] for l in lines) or 1
print(f"total stall samples {ts}, warp instructions {ti}")
for l in sorted(lines, key=lambda x: -x[2])[:top]:
    print(f"{l[0]:4d} {100 * l[2] / ts:5.1f}% smp {100 * l[3] / ti:5.1f}% inst  {l[1]}")
