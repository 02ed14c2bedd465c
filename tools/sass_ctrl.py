#!/usr/bin/env python
"""Decode the scheduling control fields of sm_100a SASS (cuobjdump -sass) for one function.
usage: sass_ctrl.py lib.so <function-substring> [start_addr end_addr]"""
import re, subprocess, sys
so, name = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi_ = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
funcs = re.split(r'\n\s*Function : ', txt)
body = next(f for f in funcs if name in f.split('\n')[0])
lines = body.split('\n')
pat = re.compile(r'^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/')
pat2 = re.compile(r'^\s+/\* (0x[0-9a-f]{16}) \*/')
i = 0
while i < len(lines) - 1:
    m, m2 = pat.match(lines[i]), pat2.match(lines[i + 1])
    if m and m2:
        h = int(m2.group(1), 16)
        a = int(m.group(1), 16)
        if lo <= a <= hi_:
            print('%05x st=%2d y=%d wr=%d rd=%d wait=%02x ru=%x  %s' % (a, (h >> 41) & 0xf, (h >> 45) & 1, (h >> 46) & 7,
                  (h >> 49) & 7, (h >> 52) & 0x3f, (h >> 58) & 0xf, m.group(2)))
        i += 2
    else:
        i += 1
