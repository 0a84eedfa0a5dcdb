"""Summarise an `ncu --page source --csv` export: executed warp-instructions per opcode, hot stall lines."""
import csv, sys, collections, re
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iw = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
tot = 0; by_op = collections.Counter(); samp_op = collections.Counter(); lines = []
for r in rows[2:]:
    if len(r) <= iex: continue
    src = r[isrc].strip(); ex = int(r[iex] or 0); sm = int(r[ismp] or 0)
    op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0] if src else "?"
    by_op[op] += ex; samp_op[op] += sm; tot += ex
    lines.append((ex, sm, src, int(r[iw] or 0) if iw is not None else 0))
print("total warp-instructions executed:", tot)
launched = lines[0][0]
print("warps launched (first instr):", launched, " => instr per warp:", tot / max(1, launched))
print("\n-- by opcode (executed, share, stall samples)")
for op, ex in by_op.most_common(28):
    print(f"{op:28s} {ex:14d} {ex/tot:6.3f}  samples {samp_op[op]}")
print("\n-- top stall-sample lines")
for ex, sm, src, wf in sorted(lines, key=lambda t: -t[1])[:25]:
    print(f"samples {sm:7d} exec {ex:12d} smem_wavefronts {wf:12d}  {src}")
