"""Turn the scratch artefacts of one gpurun visit (gpurun_out/) into the tracked summaries under profiles/.
usage: python tools/make_profiles.py <tag> <round label>"""
import collections, csv, io, json, os, subprocess, sys
tag, label = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
os.makedirs(pr, exist_ok=True)
# ---- launch list --------------------------------------------------------------------------
rows = list(csv.reader(l for l in open(os.path.join(go, f"launches_{tag}.csv")) if l.startswith('"')))
h = rows[0]; iK, iV, iM = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[iM] == "gpu__time_duration.sum":
        name = r[iK].split("(")[0][:70]
        agg.setdefault(name, []).append(float(r[iV].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(pr, f"{label}_launches.md"), "w") as f:
    f.write(f"# {label}: ncu launch list of the engine's kernels (namespace tcb), one bench step\n\n")
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:tcb:: "
            "python bench.py --steps 1 --warmup 1` (cold-cache, serialised: compare SHARES, not absolutes).\n\n")
    f.write("| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| `{k}` | {len(v)} | {sum(v)/1e6:.3f} | {sum(v)/len(v)/1e6:.3f} | {100*sum(v)/tot:.1f}% |\n")
# ---- full capture of the pass kernel --------------------------------------------------------
rep = os.path.join(go, f"prof_pass_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hd = rr[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return None
unit = {w: rr[1][hd.index(w)] for w in want if w in hd}
vals = {w: [num(r[hd.index(w)]) for r in rr[2:]] for w in want if w in hd}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
rd = [v * scale.get(unit["dram__bytes_read.sum"], 1.0) for v in vals["dram__bytes_read.sum"]]
wr = [v * scale.get(unit["dram__bytes_write.sum"], 1.0) for v in vals["dram__bytes_write.sum"]]
traffic = sum(a + b for a, b in zip(rd, wr)) / len(rd)
json.dump({"dram_bytes_per_launch": traffic, "launches_captured": len(rd), "source": f"profiles/{label}_pass_kernel_ncu.md"},
          open(os.path.join(pr, "pass_kernel_traffic.json"), "w"))
summ = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
with open(os.path.join(pr, f"{label}_pass_kernel_ncu.md"), "w") as f:
    f.write(f"# {label}: `ncu --set full --clock-control none --import-source on -k regex:pass_kernel` on bench.py (30-qubit QAOA)\n\n")
    f.write("| metric | unit | per captured launch |\n|---|---|---|\n")
    for w in want:
        if w in vals:
            f.write(f"| {w} | {unit[w]} | {', '.join(f'{v:.4g}' for v in vals[w])} |\n")
    f.write(f"\nDRAM traffic per launch (read + write): {traffic/1e9:.3f} GB; algorithmic bytes 17.18 GB (16 B x 2^30).\n")
    f.write("\n## instruction mix, stall reasons, hottest SASS regions (tools/ncu_summary.py)\n\n```\n" + summ + "```\n")
print(open(os.path.join(pr, f"{label}_launches.md")).read())
print(open(os.path.join(pr, f"{label}_pass_kernel_ncu.md")).read()[:3000])
