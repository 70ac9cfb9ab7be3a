"""Summarise an .ncu-rep: headline metrics, instruction mix, stall reasons, hottest SASS regions."""
import sys, subprocess, csv, io, collections, re
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
 'smsp__inst_executed.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__cycles_elapsed.avg',
 'smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum','l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum','lts__t_bytes.sum',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum','smsp__warps_eligible.avg.per_cycle_active']
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"{w:75s} {[r[i] for r in rows[2:]]} {rows[1][i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
h = rows[hi]; iS = h.index('Source'); iE = h.index('Instructions Executed'); iN = h.index('# Samples')
data = [(r[iS].strip(), int(r[iE]), int(r[iN] or 0), r) for r in rows[hi+1:] if len(r) > iE and r[iE].isdigit()]
tot_e = sum(d[1] for d in data); tot_s = sum(d[2] for d in data)
ops = collections.Counter(); 
for s_, e, n_, _ in data:
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', s_); ops[m.group(2) if m else '?'] += e
print("static SASS", len(data), "executed warp-instrs", tot_e)
print("mix:", ", ".join(f"{k} {100*v/tot_e:.1f}%" for k, v in ops.most_common(14)))
stall = collections.Counter()
cols = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
for *_, r in data:
    for i in cols:
        try: stall[h[i]] += int(r[i] or 0)
        except: pass
ts = sum(stall.values())
print("stalls:", ", ".join(f"{k[6:]} {100*v/ts:.1f}%" for k, v in stall.most_common(9)))
B = 128
reg = []
for b in range(0, len(data), B):
    ch = data[b:b+B]; reg.append((sum(d[2] for d in ch), sum(d[1] for d in ch), b))
for s_, e, b in sorted(reg, reverse=True)[:8]:
    if tot_s == 0: break
    ch = data[b:b+B]; o = collections.Counter((re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", d[0]) or re.match("()(.*)", "x ?")).group(2) for d in ch)
    print(f"  sass[{b:5d}:{b+B:5d}] samples {100*s_/tot_s:4.1f}% exec {100*e/tot_e:4.1f}%  {o.most_common(6)}")
