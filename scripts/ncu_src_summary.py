"""Condense `ncu -i X.ncu-rep --page source --csv` (stdin) to a few KB: per kernel, the opcodes and the
individual SASS lines that collect most warp-stall samples, with their dominant stall reasons.

usage: ncu -i X.ncu-rep --page source --csv | python scripts/ncu_src_summary.py [top_lines]"""
import collections
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rd = csv.reader(sys.stdin)
kern, hdr = None, None
lines = collections.defaultdict(list)
for row in rd:
    if row and row[0] == "Kernel Name":
        kern, hdr = row[1][:120], None
        continue
    if row and row[0] == "Address":
        hdr = row
        continue
    if hdr is None or len(row) < len(hdr):
        continue
    lines[kern].append(dict(zip(hdr, row)))
for kern, rows in lines.items():
    print("==== kernel", kern)
    tot = sum(float(r["# Samples"] or 0) for r in rows) or 1.0
    stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
    agg = collections.defaultdict(float)
    inst = collections.defaultdict(float)
    for r in rows:
        toks = r["Source"].split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "")
        agg[op] += float(r["# Samples"] or 0)
        inst[op] += float(r["Instructions Executed"] or 0)
    print("-- opcodes by samples")
    for op, v in sorted(agg.items(), key=lambda kv: -kv[1])[:18]:
        print(f"  {op:<28} {100 * v / tot:6.2f}%   inst {inst[op]:14.0f}")
    print("-- stall reasons (all lines)")
    st = {c: sum(float(r[c] or 0) for r in rows) for c in stall_cols}
    ssum = sum(st.values()) or 1.0
    for c, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  {c:<24} {100 * v / ssum:6.2f}%")
    print("-- hottest SASS lines")
    for i, r in sorted(enumerate(rows), key=lambda ir: -float(ir[1]["# Samples"] or 0))[:top]:
        s = float(r["# Samples"] or 0)
        if s <= 0:
            break
        why = sorted(((float(r[c] or 0), c) for c in stall_cols), reverse=True)[:2]
        print(f"  #{i:<6} {100 * s / tot:5.2f}%  {r['Source'].strip()[:70]:<70}  " + ", ".join(f"{c[6:]}={v:.0f}" for v, c in why if v > 0))
