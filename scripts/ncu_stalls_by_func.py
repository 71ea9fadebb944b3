"""per source function: share of pc samples split by stall reason (ncu SASS source page csv + nvdisasm line info)
usage: ncu_stalls_by_func.py <sass_page.csv> <disasm.txt> <kernel-symbol> <source.cuh>"""
import csv, re, collections, sys
sass_csv, disasm, sym, srcf = sys.argv[1:5]
rows=list(csv.reader(open(sass_csv)))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="Address")
hdr=rows[hi]; ci={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[hi+1:] if len(r)>ci["stall_wait"]]
base=int(data[0][0],16)
lines=open(disasm).read().split("\n")
start=next(i for i,l in enumerate(lines) if l.startswith(".text."+sym+":"))
cur=None; off2line={}
for l in lines[start+1:]:
    if l.startswith(".text.") or l.startswith("\t.section"): break
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split("/")[-1],int(m.group(2))); continue
    m=re.match(r"\s*/\*([0-9a-f]{4,})\*/",l)
    if m: off2line[int(m.group(1),16)]=cur
src=open(srcf).read().split("\n")
fname=srcf.split("/")[-1]
funcs=[]
for i,l in enumerate(src,1):
    m=re.match(r"\s*(?:template <[^>]*>\s*)?__device__ (?:__forceinline__ )?[\w:<>&\* ]+?\b(\w+)\(",l)
    if m: funcs.append((i,m.group(1)))
reasons=["stall_no_inst","stall_long_sb","stall_wait","stall_short_sb","stall_selected","stall_branch_resolving","stall_lg","stall_math"]
agg=collections.defaultdict(collections.Counter)
for r in data:
    off=int(r[0],16)-base
    key=off2line.get(off,("?",0))
    name=key[0]+":?"
    if key[0]==fname:
        cand=[n for (st,n) in funcs if st<=key[1]]; name=cand[-1] if cand else "top"
    for rs in reasons: agg[name][rs]+=float(r[ci[rs]] or 0)
    agg[name]["all"]+=float(r[ci["# Samples"]] or 0)
    agg[name]["inst"]+=float(r[ci["Instructions Executed"]] or 0)
tot=sum(a["all"] for a in agg.values()); toti=sum(a["inst"] for a in agg.values())
print(f"total samples {tot:.0f}, warp-instructions {toti:.3e}")
print(f"{'function':26s} {'all%':>6s} {'inst%':>6s} " + " ".join(f"{r[6:]:>9s}" for r in reasons))
for name,a in sorted(agg.items(), key=lambda x:-x[1]["all"])[:20]:
    print(f"{name:26s} {100*a['all']/tot:6.2f} {100*a['inst']/toti:6.2f} " + " ".join(f"{100*a[r]/tot:9.2f}" for r in reasons))
print(f"{'TOTAL':26s} {100.0:6.2f} {100.0:6.2f} " + " ".join(f"{100*sum(a[r] for a in agg.values())/tot:9.2f}" for r in reasons))
