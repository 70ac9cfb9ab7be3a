"""Helper run in a subprocess by tests/test_gpu_tc.py (and by hand): pairwise contractions through the
C ABI vs numpy einsum, for whatever kernel TCB_TN_KERNEL selects."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorcircuit_ng_b200 import tnengine  # noqa: E402


def check(nb, nm, nn, nk, seed, conj_a=False, conj_b=False, tol=None):
    # FP32 SIMT kernel: ~1e-6; 3xTF32 keeps ~22 mantissa bits per operand: ~1e-5 at K = 512
    tol = tol or (1.5e-5 if os.environ.get("TCB_TN_KERNEL", "auto") != "simt" else 2e-6)
    rng = np.random.default_rng(seed)
    letters = [chr(ord("a") + i) for i in range(nb + nm + nn + nk)]
    bat, ms, ns, ks = letters[:nb], letters[nb:nb + nm], letters[nb + nm:nb + nm + nn], letters[nb + nm + nn:]
    ma = list(rng.permutation(bat + ms + ks))
    mb = list(rng.permutation(bat + ns + ks))
    mc = list(rng.permutation(bat + ms + ns))
    a = (rng.normal(size=[2] * len(ma)) + 1j * rng.normal(size=[2] * len(ma))).astype(np.complex64)
    b = (rng.normal(size=[2] * len(mb)) + 1j * rng.normal(size=[2] * len(mb))).astype(np.complex64)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = tnengine.contract_raw(ta, ma, tb, mb, mc, conj_a=conj_a, conj_b=conj_b)
    torch.cuda.synchronize()
    a64 = (a.conj() if conj_a else a).astype(np.complex128)
    b64 = (b.conj() if conj_b else b).astype(np.complex128)
    ref = np.einsum(a64, [letters.index(x) for x in ma], b64, [letters.index(x) for x in mb], [letters.index(x) for x in mc])
    err = np.abs(out.cpu().numpy() - ref).max() / max(1.0, np.abs(ref).max())
    # accumulate into an existing output
    out2 = tnengine.contract_raw(ta, ma, tb, mb, mc, conj_a=conj_a, conj_b=conj_b, out=out.clone(), accumulate=True)
    err2 = np.abs(out2.cpu().numpy() - 2 * ref).max() / max(1.0, np.abs(ref).max())
    print(f"nb={nb} nm={nm} nn={nn} nk={nk} conj=({int(conj_a)},{int(conj_b)}): rel err {err:.2e} / acc {err2:.2e}", flush=True)
    assert err <= tol and err2 <= 2 * tol, (err, err2)


if __name__ == "__main__":
    cases = [(0, 7, 4, 3), (0, 8, 6, 5), (1, 7, 3, 4), (0, 5, 2, 1), (2, 9, 5, 6), (0, 10, 7, 7), (0, 7, 0, 9), (3, 7, 4, 0),
             (0, 12, 6, 6), (0, 3, 8, 2),
             # skinny steps: the streaming kernel (auto / simt modes)
             (0, 14, 1, 1), (0, 12, 3, 2), (2, 11, 2, 1), (3, 10, 0, 3), (1, 13, 3, 0), (0, 10, 2, 3), (2, 9, 1, 2)]
    for i, c in enumerate(cases):
        check(*c, seed=i, conj_a=bool(i % 2), conj_b=bool((i // 2) % 2))
    print("ok", os.environ.get("TCB_TN_KERNEL", "auto"))
