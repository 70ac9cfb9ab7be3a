"""
Sampling bitstrings from a resident statevector.

The reference's `sample(allow_state=True)` materialises p = |psi|^2, its cumulative sum and an arange of
2^n indices and searches them (`tensorcircuit/basecircuit.py:1490-1512`,
`tensorcircuit/backends/abstract_backend.py:1828-1861`); its perfect sampling contracts the network once per
qubit per shot (`:449-558`).  With the state in HBM both are the same two kernels: `tcb_sv_sample_prepare`
reads the state once and leaves the float64 CDF of its 4096-amplitude segments (2 MiB at 30 qubits), and
`tcb_sv_sample` resolves each shot inside the one segment it lands in — by CDF inversion with one uniform
(mode 0, `probability_sample`'s rule) or by the qubit-by-qubit conditional walk with n uniforms (mode 1,
`measure_jit`'s rule).
"""

from __future__ import annotations

from typing import Any, Optional, Tuple

import torch

from . import _lib

SEG_BITS = 12


class StateSampler:
    """The segment CDF of one state, reusable for any number of draws."""

    def __init__(self, psi: torch.Tensor, nbits: int) -> None:
        _lib.require_cuda(psi, "state")
        self.psi = psi.detach().to(torch.complex64).resolve_conj().contiguous().reshape(-1)
        if self.psi.numel() != (1 << nbits):
            raise ValueError(f"state has {self.psi.numel()} amplitudes, expected 2^{nbits}")
        self.nbits = nbits
        self.seg_bits = min(SEG_BITS, nbits)
        self.cdf = torch.empty(1 << (nbits - self.seg_bits), dtype=torch.float64, device=self.psi.device)
        _lib.call("tcb_sv_sample_prepare", self.psi.data_ptr(), nbits, self.seg_bits, self.cdf.data_ptr(),
                  _lib.stream_ptr())  # fmt: skip

    def draw(self, status: torch.Tensor, mode: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """status: [shots] (mode 0) or [shots, nbits] (mode 1) uniforms in [0, 1).
        Returns (int64 flat indices [shots], float64 probabilities [shots])."""
        st = torch.as_tensor(status).detach().to(device=self.psi.device, dtype=torch.float64).contiguous()
        want = 1 if mode == 0 else 2
        if st.dim() != want or (mode == 1 and st.shape[1] != self.nbits):
            raise ValueError(
                "status must have shape [shots]" if mode == 0 else f"status must have shape [shots, {self.nbits}]"
            )
        shots = int(st.shape[0])
        idx = torch.empty(shots, dtype=torch.int64, device=self.psi.device)
        prob = torch.empty(shots, dtype=torch.float64, device=self.psi.device)
        _lib.call("tcb_sv_sample", self.psi.data_ptr(), self.nbits, self.seg_bits, self.cdf.data_ptr(), st.data_ptr(),
                  shots, mode, idx.data_ptr(), prob.data_ptr(), _lib.stream_ptr())  # fmt: skip
        return idx, prob


def uniforms(shape: Any, random_generator: Optional[torch.Generator], device: torch.device) -> torch.Tensor:
    """float64 uniforms in [0, 1) from `random_generator` (a torch.Generator, on any device) or the global one."""
    gdev = random_generator.device if random_generator is not None else device
    return torch.rand(shape, generator=random_generator, dtype=torch.float64, device=gdev).to(device)
