#!/usr/bin/env python
"""Static SASS instruction count of one kernel per source line (innermost inlined frame), from
`nvdisasm --print-line-info` of the library's cubin -- where a kernel's code size comes from.

    python tools/sass_regions.py [mangled-kernel-name] [top-n]
"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kern = sys.argv[1] if len(sys.argv) > 1 else "_ZN5tapir18profile_mom_kernelILi8ELb0ELi16EEEvNS_8ProfArgsE"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "tapir_b200", "libtapir_cuda.so")], cwd=tmp,
                   check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "ingest" not in f][0]
    text = subprocess.run(["nvdisasm", "--print-line-info", cubin], cwd=tmp, check=True, capture_output=True, text=True).stdout
lines = text.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text." + kern + ":")][0]
end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("//--------------------- ."))
cur, agg, n = None, collections.Counter(), 0
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        agg[cur] += 1
        n += 1
print("%s: %d instructions (%.1f KB)" % (kern, n, n * 16 / 1024))
src = {}
for (k, v) in agg.most_common(top):
    if k is None:
        print("%5d  (no line)" % v)
        continue
    if k[0] not in src:
        path = os.path.join(ROOT, "tapir_b200", "csrc", k[0])
        src[k[0]] = open(path).read().split("\n") if os.path.exists(path) else None
    code = src[k[0]][k[1] - 1].strip()[:110] if src[k[0]] else ""
    print("%5d  %s:%d  %s" % (v, k[0], k[1], code))
