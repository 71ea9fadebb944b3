"""Summarise one `ncu --set full --import-source on` capture of the warp kernel into markdown (profiles/*.md).
usage: ncu_summary.py <report.ncu-rep> <lib.so> <kernel-symbol> <title> [source file under csrc/, default mpc_wpp.cuh] > out.md
Needs ncu, cuobjdump and nvdisasm on PATH (no GPU)."""
import csv, glob, io, os, subprocess, sys, tempfile
rep, lib, sym, title = sys.argv[1:5]
src_name = sys.argv[5] if len(sys.argv) > 5 else "mpc_wpp.cuh"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.per_cycle_active", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# {title}\n")
print("| metric | unit | value |\n|---|---|---|")
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"| {w} | {units[i]} | {vals[i]} |")
st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(vals[i]) for i, h in enumerate(hdr)
      if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
tot = sum(st.values())
print("\n## warp stall reasons (pc sampling, all samples)\n\n| share | reason |\n|---|---|")
for k, v in sorted(st.items(), key=lambda x: -x[1])[:9]:
    print(f"| {100*v/tot:.1f} % | {k} |")
with tempfile.TemporaryDirectory() as td:
    sass = os.path.join(td, "sass.csv")
    open(sass, "w").write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout)
    subprocess.check_call(f"cd {td} && cuobjdump -xelf all {os.path.abspath(lib)} > /dev/null", shell=True)
    cubin = sorted(glob.glob(os.path.join(td, "*sm_100a*.cubin")))[-1]
    dis = os.path.join(td, "dis.txt")
    open(dis, "w").write(subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_stalls_by_func.py"), sass, dis, sym,
                          os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200", "csrc", src_name)], capture_output=True, text=True).stdout
print("\n## pc samples per source function, split by stall reason (% of all samples; ncu SASS page joined with nvdisasm line info)\n\n```")
print(out.rstrip())
print("```")
