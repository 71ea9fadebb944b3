"""Join an ncu SASS source page (csv) with nvdisasm --print-line-info to aggregate samples / executed
instructions per CUDA source line and per function region.
usage: ncu_by_line.py <sass_page.csv> <disasm.txt> <kernel-symbol> [source-file-for-function-map]"""
import csv, re, sys, collections
sass_csv, disasm, sym = sys.argv[1:4]
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) > ci["Instructions Executed"]]
base = int(data[0][0], 16)
per_off = {int(r[0], 16) - base: (float(r[ci["# Samples"]] or 0), float(r[ci["Instructions Executed"]] or 0), r[ci["Source"]]) for r in data}
# parse disasm
lines = open(disasm).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + sym + ":"))
cur = None; off2line = {}
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
by_line = collections.Counter(); inst_line = collections.Counter()
for off, (s, n, src) in per_off.items():
    key = off2line.get(off, ("?", 0))
    by_line[key] += s; inst_line[key] += n
tot_s = sum(by_line.values()); tot_n = sum(inst_line.values())
print(f"total samples {tot_s:.0f}  total warp-instructions {tot_n:.3e}")
# function map from source file (mpc_wpp.cuh): function start lines
if len(sys.argv) > 4:
    src = open(sys.argv[4]).read().split("\n")
    funcs = []
    for i, l in enumerate(src, 1):
        m = re.match(r"\s*(?:template <[^>]*>\s*)?__device__ (?:__forceinline__ )?[\w:<>&\* ]+?\b(\w+)\(", l)
        if m:
            funcs.append((i, m.group(1)))
    fname = sys.argv[4].split("/")[-1]
    agg_s = collections.Counter(); agg_n = collections.Counter()
    for (f, ln), s in by_line.items():
        name = f + ":?"
        if f == fname:
            cand = [n for (st, n) in funcs if st <= ln]
            name = cand[-1] if cand else "top"
        agg_s[name] += s; agg_n[name] += inst_line[(f, ln)]
    print("\nper function region (samples %, instr %):")
    for name, s in agg_s.most_common(25):
        print(f"  {name:28s} {100*s/tot_s:6.2f}%  {100*agg_n[name]/tot_n:6.2f}%")
print("\ntop lines by samples:")
for key, s in by_line.most_common(40):
    print(f"  {key[0]}:{key[1]:<5d} samples {100*s/tot_s:5.2f}%  instr {100*inst_line[key]/tot_n:5.2f}%")
