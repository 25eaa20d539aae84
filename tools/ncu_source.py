#!/usr/bin/env python
"""Top stall sites of an .ncu-rep source page (SASS view): tools/ncu_source.py rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(out)))
# a report with several kernels repeats the two header lines; pick kernel number sys.argv[3] (default 0)
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lo = starts[which]; hi = starts[which + 1] if which + 1 < len(starts) else len(allrows)
rows = allrows[lo:hi]
print("kernel:", rows[0][1][:140])
hdr = rows[1]; body = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print("total samples", tot, " instrs", len(body))
print("by reason:", ", ".join("%s=%.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
# cumulative by opcode class
byop = {}
for r in body:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"): op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    byop[op] = byop.get(op, 0) + int(r[ix["# Samples"]] or 0)
print("by opcode:", ", ".join("%s=%.1f%%" % (k, 100.0 * v / tot) for k, v in sorted(byop.items(), key=lambda kv: -kv[1])[:14]))
cnt = {}
for r in body:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"): op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    cnt[op] = cnt.get(op, 0) + 1
print("static instr mix:", ", ".join("%s=%d" % kv for kv in sorted(cnt.items(), key=lambda kv: -kv[1])[:16]))
print("top sites:")
for n, r in sorted(enumerate(body), key=lambda nr: -int(nr[1][ix["# Samples"]] or 0))[:top]:
    s = int(r[ix["# Samples"]] or 0)
    why = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print("  #%4d %5.1f%%  %-60s %s" % (n, 100.0 * s / tot, r[ix["Source"]].strip()[:60], " ".join("%s:%d" % (k, v) for v, k in why if v)))
# timeline: samples per window of 128 instructions, with the window's dominant stall reasons and its memory / barrier instructions
print("timeline (window of 128 instructions: share of samples, top stall reasons, LDG/STG/LDS/STS/BAR/FP64 instruction counts):")
W = 128
for w0 in range(0, len(body), W):
    win = body[w0:w0 + W]
    sm = sum(int(r[ix["# Samples"]] or 0) for r in win)
    why = sorted(((sum(int(r[ix[k]] or 0) for r in win), k[6:]) for k in stalls), reverse=True)[:3]
    ops = {}
    for r in win:
        f = r[ix["Source"]].split()
        op = (f[1] if f and f[0].startswith("@") and len(f) > 1 else (f[0] if f else "?")).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    fp = ops.get("DFMA", 0) + ops.get("DADD", 0) + ops.get("DMUL", 0)
    print("  @%5d %5.1f%%  %-44s LDG %2d STG %2d LDS %2d STS %2d BAR %d FP64 %3d LDL/STL %d" % (
        w0, 100.0 * sm / tot, " ".join("%s:%d" % (k, v) for v, k in why if v), ops.get("LDG", 0), ops.get("STG", 0), ops.get("LDS", 0),
        ops.get("STS", 0), ops.get("BAR", 0), fp, ops.get("LDL", 0) + ops.get("STL", 0)))

