"""
`tensorcircuit/torchnn.py:16-99`: a `torch.nn.Module` around a quantum function f(inputs..., weights...).
The batch dimension is evaluated by `backend.vmap` (a loop; each sample is a fused-pass evolution) and the
weights are ordinary `nn.Parameter`s, so any torch optimiser trains the circuit through the engine's vjps.
"""

from __future__ import annotations

from typing import Any, Callable, Sequence, Tuple, Union

import torch

from . import backend
from .interfaces import torch_interface


class QuantumNet(torch.nn.Module):
    def __init__(self, f: Callable[..., Any], weights_shape: Sequence[Tuple[int, ...]],
                 initializer: Union[Any, Sequence[Any]] = None, use_vmap: bool = True,
                 vectorized_argnums: Union[int, Sequence[int]] = 0, use_interface: bool = True,
                 use_jit: bool = True, enable_dlpack: bool = False) -> None:  # fmt: skip
        super().__init__()
        if use_vmap:
            f = backend.vmap(f, vectorized_argnums=vectorized_argnums)
        if use_interface:
            f = torch_interface(f, jit=use_jit, enable_dlpack=enable_dlpack)
        self.f = f
        self.q_weights = torch.nn.ParameterList()
        if isinstance(weights_shape[0], int):
            weights_shape = [weights_shape]  # type: ignore[list-item]
        if not isinstance(initializer, (list, tuple)):
            initializer = [initializer] * len(weights_shape)
        for ws, initf in zip(weights_shape, initializer):
            initf = torch.randn if initf is None else initf
            self.q_weights.append(torch.nn.Parameter(initf(tuple(ws))))

    def forward(self, *inputs: torch.Tensor) -> torch.Tensor:
        return self.f(*inputs, *self.q_weights)


TorchLayer = QuantumNet
