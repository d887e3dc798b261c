#!/usr/bin/env python
"""Per-CUDA-source-line executed instructions and stall samples from an .ncu-rep
(needs nvcc -lineinfo and ncu --import-source on).
usage: python profiles/ncu_source_hot.py rep.ncu-rep [topN]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur_file = ""
rows = []
total_inst = total_samp = 0
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) < 10 or r[0] in ("Line No", ""):
        continue
    try:
        line, src, samp, inst = int(r[0]), r[1], int(r[4] or 0), int(r[7] or 0)
    except ValueError:
        continue
    rows.append((inst, samp, cur_file, line, src.strip()[:100]))
    total_inst += inst
    total_samp += samp
print(f"total warp-instructions {total_inst}, stall samples {total_samp}")
print("   instr    %inst  %samples  file:line  source")
for inst, samp, f, line, src in sorted(rows, key=lambda x: -x[1])[:top]:
    print(f"{inst:11d} {100.0 * inst / max(total_inst, 1):6.2f} {100.0 * samp / max(total_samp, 1):7.2f}   {f}:{line}  {src}")
