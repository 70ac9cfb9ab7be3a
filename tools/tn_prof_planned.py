import os, sys
sys.path.insert(0, "/root/repo")
import torch
from tensorcircuit_ng_b200 import tnengine
nm, nn, nk = 18, 10, 10
letters = [chr(ord("a") + i) for i in range(nm + nn + nk)]
ms, ns, ks = letters[:nm], letters[nm:nm + nn], letters[nm + nn:]
ma, mb, mc = ms + ks, ns + ks, ms + ns
a = torch.randn([2] * len(ma), dtype=torch.complex64, device="cuda")
b = torch.randn([2] * len(mb), dtype=torch.complex64, device="cuda")
out = torch.empty([2] * len(mc), dtype=torch.complex64, device="cuda")
for _ in range(3):
    tnengine.contract_raw(a, ma, b, mb, mc, out=out)
torch.cuda.synchronize()
