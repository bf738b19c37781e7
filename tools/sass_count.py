#!/usr/bin/env python3
"""Static SASS analysis helper: per-kernel instruction histogram, split by basic-block loops (development tool).
usage: sass_count.py <binary> <kernel-substring>"""
import re, subprocess, sys, collections
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
cur, fn = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); fn[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        fn[cur].append((int(m.group(1), 16), m.group(2)))
for name, ins in fn.items():
    if sys.argv[2] not in name:
        continue
    print("==", name, len(ins), "instructions")
    # find backward branches = loops
    loops = []
    for addr, txt in ins:
        m = re.search(r"BRA\S*\s+(?:!?U?P\d,\s*)?`?\(?0x([0-9a-f]+)", txt)
        if m and int(m.group(1), 16) <= addr:
            loops.append((int(m.group(1), 16), addr))
    print("loops (start,end):", [(hex(a), hex(b)) for a, b in loops])
    def hist(lo, hi):
        h = collections.Counter()
        for addr, txt in ins:
            if lo <= addr <= hi:
                t = txt.split()
                op = t[1] if t[0].startswith("@") else t[0]
                h[op.split(".")[0] + (".WIDE" if ".WIDE" in op else "")] += 1
        return h
    for a, b in loops:
        h = hist(a, b)
        print(" loop %x-%x: %d instr:" % (a, b, sum(h.values())), dict(h.most_common(12)))
    h = hist(0, 1 << 40)
    print(" total:", dict(h.most_common(14)))
