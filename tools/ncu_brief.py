"""Brief markdown summary of an .ncu-rep (headline + tensor-pipe metrics) for profiles/. usage: ncu_brief.py rep title out.md"""
import csv, io, subprocess, sys, os
rep, title, out = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hd = rr[0]
keys = [k for k in hd if any(s in k for s in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct", "sm__throughput.avg.pct", "pipe_tensor", "sm__inst_executed_pipe_tensor",
        "sm__pipe_fma_cycles_active.avg.pct", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "Kernel Name"))]
with open(out, "w") as f:
    f.write(f"# {title}\n\n| metric | unit | value |\n|---|---|---|\n")
    for k in keys:
        i = hd.index(k)
        f.write(f"| {k} | {rr[1][i]} | {', '.join(r[i] for r in rr[2:])} |\n")
    summ = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    f.write("\n```\n" + "\n".join(l for l in summ.splitlines() if l.startswith(("mix:", "stalls:", "static SASS"))) + "\n```\n")
print(open(out).read())
