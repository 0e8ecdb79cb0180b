#!/usr/bin/env python
"""Development aid: list the backward-branch loops of one kernel in a cuobjdump -sass dump with their
instruction count and opcode histogram (is the 1 kHz body spill-free? how many instructions per tick?).

    cuobjdump -sass -fun <mangled> lib.so > k.sass ; python tools/sass_loops.py k.sass [min_len]
"""
import collections
import re
import sys

ins = []
for ln in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
min_len = int(sys.argv[2]) if len(sys.argv) > 2 else 20
print("total instructions", len(ins))
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if not m:
        continue
    tgt = int(m.group(1), 16)
    if tgt <= a and tgt in addr_index and i - addr_index[tgt] >= min_len:
        body = ins[addr_index[tgt]:i + 1]
        ops = collections.Counter()
        for _, s in body:
            s = re.sub(r"^@!?U?P\d\s+", "", s)
            ops[s.split()[0].split(".")[0]] += 1
        print(f"loop {tgt:#x}..{a:#x}: {len(body)} instructions")
        print("   ", ", ".join(f"{k} {v}" for k, v in ops.most_common()))
