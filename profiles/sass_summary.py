#!/usr/bin/env python
"""SASS listings of the kernels DESIGN.md discusses: full listings (gzip) + opcode histograms.
usage: python profiles/sass_summary.py [TAG]   (runs cuobjdump on mptc_b200/libmptc_b200.so; no GPU needed)"""
import collections
import gzip
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"
KERNELS = ["k_inter_search_wide", "k_inter_search_tiled", "k_intra_rows", "k_dxt1_to_rgb", "k_inter_pixel_search", "k_dxt1_fit", "k_endpoint_planes"]
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "mptc_b200", "libmptc_b200.so")], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
out = [f"opcode histograms of the sm_100a SASS (cuobjdump -sass mptc_b200/libmptc_b200.so), {TAG}\n"]
for k in KERNELS:
    for body in funcs[1:]:
        name = body.split("\n", 1)[0]
        if k not in name:
            continue
        lines = [ln for ln in body.split("\n") if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", ln)]
        ops = collections.Counter()
        for ln in lines:
            m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if m:
                ops[m.group(1).split(".")[0] if not m.group(1).startswith(("UBLKCP", "STG", "LDG", "ATOM", "RED", "MEMBAR", "LDS", "STS")) else m.group(1)] += 1
        tag = name.split("(")[0].strip()
        short = re.sub(r"^_ZN4mptc\d*", "", tag)
        with gzip.open(os.path.join(ROOT, "profiles", f"{TAG}_sass_{k}{''.join('_mode' + m for m in re.findall(r'ILi(\d)E', name))}.txt.gz"), "wt") as f:
            f.write("Function : " + body)
        out.append(f"== {name.strip()}  ({len(lines)} instructions)")
        out.append("   " + "  ".join(f"{o}x{n}" for o, n in ops.most_common(28)))
        special = {o: n for o, n in ops.items() if o.startswith(("UBLKCP", "VABSDIFF4", "IDP", "PRMT", "FADD2", "SHFL", "MEMBAR", "ATOM", "RED", "STG.E.64.STRONG", "LDG.E.64.STRONG", "UTMA", "UTC"))}
        out.append("   of interest: " + ", ".join(f"{o} x{n}" for o, n in sorted(special.items())))
open(os.path.join(ROOT, "profiles", f"{TAG}_sass_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
