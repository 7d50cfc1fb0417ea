#!/usr/bin/env python
"""Static SASS instruction mix of one kernel of the built library (substring match on the name).
Usage: python tools/sass_mix.py sh_render_kernelILi512 [lib.so]"""
import collections
import re
import subprocess
import sys

name = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "retisar_b200/lib/libretisar_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, mix = None, collections.Counter()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and name in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m:
            mix[m.group(2)] += 1
print(sum(mix.values()), "instructions")
print(mix.most_common(30))
