"""static SASS size per source function of one kernel: nvdisasm --print-line-info output joined with the function
start lines of mpc_wpp.cuh.  usage: sass_size_by_func.py <lib.so> <kernel-symbol-substring>"""
import collections, re, subprocess, sys, os
lib, sym = sys.argv[1:3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cub = "/tmp/_k.cubin"
subprocess.check_call(f"cd /tmp && cuobjdump -xelf all {lib} > /dev/null && ls /tmp/*.cubin | head -1", shell=True, stdout=subprocess.DEVNULL)
import glob
cubin = sorted(glob.glob("/tmp/*sm_100a*.cubin"), key=os.path.getmtime)[-1]
txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(txt) if l.startswith(".text.") and sym in l)
srcs = {}
def funcs_of(fname):
    path = os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200", "csrc", fname)
    if fname not in srcs:
        out = []
        if os.path.exists(path):
            for i, l in enumerate(open(path).read().split("\n"), 1):
                m = re.match(r"\s*(?:template <[^>]*>\s*)?__device__ (?:__forceinline__ )?[\w:<>&\* ]+?\b(\w+)\(", l)
                if m: out.append((i, m.group(1)))
        srcs[fname] = out
    return srcs[fname]
cur = None; cnt = collections.Counter(); total = 0
for l in txt[start + 1:]:
    if l.startswith(".text.") or l.startswith("\t.section"): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        total += 1
        if cur is None: cnt["?"] += 1; continue
        f, ln = cur
        cand = [n for (st, n) in funcs_of(f) if st <= ln]
        cnt[f"{f}:{cand[-1] if cand else '?'}"] += 1
print(f"{txt[start][:80]}  total {total} instr = {total*16/1024:.1f} KB")
for k, v in cnt.most_common(30): print(f"  {k:45s} {v:6d}  {v*16/1024:6.1f} KB  {100*v/total:5.1f}%")
