"""Summarise an `ncu --page source --csv` export (first kernel section): executed warp-instructions per opcode,
hot stall lines.  python tools/ncu_src_summary.py file.csv [n_opcodes]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 28
sec, n = [], 0
for r in rows:
    if r and r[0] == "Kernel Name":
        n += 1
        if n == 2:
            break
        continue
    sec.append(r)
hdr = sec[0]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iw = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
tot = 0; by_op = collections.Counter(); samp_op = collections.Counter(); lines = []
for r in sec[1:]:
    if len(r) <= iex: continue
    src = r[isrc].strip(); ex = int(r[iex] or 0); sm = int(r[ismp] or 0)
    op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0] if src else "?"
    by_op[op] += ex; samp_op[op] += sm; tot += ex
    lines.append((ex, sm, src, int(r[iw] or 0) if iw is not None else 0))
print("total warp-instructions executed:", tot)
launched = lines[0][0]
print("warps launched (first instr):", launched, " => instr per warp:", tot / max(1, launched))
print("\n-- by opcode (executed, share, stall samples)")
for op, ex in by_op.most_common(top):
    print(f"{op:28s} {ex:14d} {ex/tot:6.3f}  samples {samp_op[op]}")
print("\n-- top stall-sample lines")
for ex, sm, src, wf in sorted(lines, key=lambda t: -t[1])[:15]:
    print(f"samples {sm:7d} exec {ex:12d} smem_wavefronts {wf:12d}  {src}")
