"""Developer tool: list the loops (backward branches) of a kernel's SASS with instruction counts per loop body.
usage: cuobjdump -sass -fun <mangled> file.o | python tools/sass_loops.py [min_len]"""
import re
import sys
from collections import Counter

min_len = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
print("instructions:", len(ins))
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)*(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            j = addr_index[tgt]
            n = i - j + 1
            if n >= min_len:
                ops = Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for _, x in ins[j:i + 1])
                top = ", ".join("%s %d" % kv for kv in ops.most_common(14))
                print("loop 0x%x..0x%x: %d instr | %s" % (tgt, a, n, top))
