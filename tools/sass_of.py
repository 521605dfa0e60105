#!/usr/bin/env python
"""Plain SASS listing of one kernel of libcube_b200.so:  tools/sass_of.py <regex> [lib] [--mix]"""
import collections
import re
import subprocess
import sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
pat = re.compile(args[0])
lib = args[1] if len(args) > 1 else "rubiks_cube_solver_b200/libcube_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
on, seen, lines = False, 0, []
for line in out.splitlines():
    if "Function :" in line:
        on = bool(pat.search(line)) and seen == 0
        seen += on
        continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*/\*", line)
        if m:
            lines.append(m.group(2).rstrip(" ;"))
if "--mix" in sys.argv:
    c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", l).split()[0].split(".")[0] for l in lines)
    print(len(lines), "instructions:", ", ".join("%s %d" % kv for kv in c.most_common(16)))
else:
    print("\n".join(lines))
