"""Developer tool: run one pairwise contraction shape for ncu (nm nn nk)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tensorcircuit_ng_b200 import tnengine
nm, nn, nk = (int(x) for x in sys.argv[1:4])
letters = [chr(ord("a") + i) for i in range(nm + nn + nk)]
ms, ns, ks = letters[:nm], letters[nm:nm + nn], letters[nm + nn:]
a = torch.randn([2] * (nm + nk), dtype=torch.complex64, device="cuda")
b = torch.randn([2] * (nk + nn), dtype=torch.complex64, device="cuda")
out = torch.empty([2] * (nm + nn), dtype=torch.complex64, device="cuda")
for _ in range(3):
    tnengine.contract_raw(a, ms + ks, b, ks + ns, ms + ns, out=out)
torch.cuda.synchronize(); print("done")
