#!/usr/bin/env python
"""Per-opcode executed warp-instruction counts (sass__inst_executed_per_opcode) of every launch in an .ncu-rep, as text:
    python tools/ncu_opcodes.py gpurun_out/x.ncu-rep > profiles/r02_ncu_opcodes_cfg2.txt
One block per launch:  == kernel name ==  then  OPCODE count  lines, largest first."""
import re
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--print-metric-instances", "details", "--page", "raw", "--metrics",
                      "sass__inst_executed_per_opcode"], capture_output=True, text=True).stdout
print("# source: %s (ncu --set full --clock-control none); warp instructions executed per opcode, one block per launch" % sys.argv[1])
name = None
buf = ""
for line in raw.splitlines() + ["  END (0, 0, 0)x(0, 0, 0), Context"]:
    m = re.match(r"^\s\s(\S.*?) \(\d+, \d+, \d+\)x\(\d+, \d+, \d+\), ", line)
    if m:
        if name is not None:
            pairs = re.findall(r"([A-Z][A-Z0-9_.]*):\s*(\d+)", buf.replace("\n", " "))
            total = re.search(r"sass__inst_executed_per_opcode\s+(\d+)", buf)
            print("\n== %s ==" % name)
            if total:
                print("%-12s %d" % ("TOTAL", int(total.group(1))))
            for op, n in sorted(pairs, key=lambda p: -int(p[1])):
                print("%-12s %d" % (op, int(n)))
        name, buf = m.group(1), ""
    else:
        buf += re.sub(r"\s+", " ", line) + " "
