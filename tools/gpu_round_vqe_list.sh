#!/bin/bash
# ncu launch list (durations only) of one VQE value_and_grad step at batch 64
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:tcb:: --csv \
  --log-file gpurun_out/launches_vqe_r2.csv python bench.py --workload vqe --steps 1 --warmup 1 > gpurun_out/ncu_vqe_r2.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_vqe_r2.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    ms = v / 1e6 if u == "ns" else (v / 1e3 if u == "us" else v)
    k = r[ki].split("(")[0][:60]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} {a[0]:5d} {a[1]:9.3f} ms {a[1]/a[0]:8.3f} avg {100*a[1]/tot:5.1f}%")
PY
