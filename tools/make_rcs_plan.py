"""Search and store the contraction plan of bench.py's configs[4] workload (plans/*.pkl, tree_data schema)."""
import math, os, pickle, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import tensorcircuit_ng_b200 as tc
from tensorcircuit_ng_b200 import planner
from tensorcircuit_ng_b200.experimental import DistributedContractor
rows = cols = int(sys.argv[1]) if len(sys.argv) > 1 else 7
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lt = int(sys.argv[3]) if len(sys.argv) > 3 else 30
t0 = time.time()
nodes_fn = lambda _: bench.build_rcs(tc, rows, cols, depth).amplitude_before("0" * (rows * cols))
inp, out, sd, _, groups = DistributedContractor._network(nodes_fn, None, True)
td = planner.search_sites(inp, out, sd, groups, target_size=2**lt, max_slices_log2=80)
td["hyper_diagonal"] = True
st = planner.path_stats(td["inputs"], td["output"], td["size_dict"], td["path"], list(td["sliced_inds"]))
print(f"{rows}x{cols} d{depth}: tensors {len(inp)} log10 cmacs/slice {math.log10(st['flops']):.2f} log2 size {math.log2(st['size']):.0f} "
      f"log2 write {math.log2(st['write']):.1f} sliced {len(td['sliced_inds'])}  ({time.time()-t0:.0f} s)")
os.makedirs(os.path.join(ROOT, "plans"), exist_ok=True)
pickle.dump(td, open(bench.rcs_plan_path(rows, cols, depth, lt), "wb"))
